"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle restatement on the same
inputs, against the committed reference fixtures, and -- when oracle/_ref travelled -- against the
reference's own code.  Integer/ids bit-exact; fp32 t/u/v bit-exact (tolerance: 0 ulp)."""
import hashlib

import numpy as np
import pytest

from dod_raytracer_b200 import capi
from gpu_util import assert_hits_equal, upload
from oracle_api import (CLS_BOX, CLS_CYLINDER, CLS_PLANE, CLS_SPHERE, CLS_TREE, MISS, RAY_ANY, RefLib, Scene, have_ref,
                        reference_cylinder, reference_planes)
from scenes import (GOLDEN, LIGHT0, analytic_scene_arrays, hit_points, load_teapot_arrays, make_rays, teapot_scene)

pytestmark = pytest.mark.gpu
ALL = CLS_SPHERE | CLS_PLANE | CLS_CYLINDER | CLS_TREE
CLASSES = {"tree": CLS_TREE, "all": ALL, "sphere_tree": CLS_SPHERE | CLS_TREE}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# the variants THIS build holds (product: 0, 3, 7; libdodrt_cuda_exp.so, run by tests/test_gpu_experiments.py: 0-8)
VARIANTS = [v for v in range(9) if capi.variant_available(v)]


@pytest.fixture(scope="module", params=VARIANTS, ids=lambda v: f"variant{v}")
def teapot(request):
    """every kernel variant must return the same bits"""
    scene = teapot_scene(full=True)
    g = upload(scene)
    g.set_kernel_variant(request.param)
    yield scene, g
    g.close()


def test_explicit_ray_batches_match_oracle(teapot, oracle):
    scene, g = teapot
    rays = oracle.primary_rays(640, 360)
    for cls in (CLS_TREE, CLS_SPHERE, CLS_PLANE, CLS_CYLINDER, ALL, CLS_SPHERE | CLS_TREE, CLS_PLANE | CLS_CYLINDER):
        want = oracle.intersect(scene, rays, cls, nthreads=8)
        got = g.intersect(rays, cls)
        assert_hits_equal(got, want, what=f"classes={cls}")


def test_full_1080p_frame_matches_reference_fixture(teapot):
    """BASELINE.json config 1 geometry: teapot, 1920x1080, srand(1) spheres + planes + cylinder, light0.
    Checksums of the complete hit / visibility arrays produced by the reference itself."""
    _, g = teapot
    z = np.load(f"{GOLDEN}/teapot_frame.npz")
    w, h, step = int(z["width"]), int(z["height"]), int(z["step"])
    sel = (np.arange(0, h, step)[:, None] * w + np.arange(0, w, step)[None, :]).ravel()
    from oracle_api import Oracle
    xs, ys = Oracle().ray_tables(w, h)
    for name, cls in CLASSES.items():
        frame = capi.Frame.make(w, h, classes=cls)
        hits = g.trace_primary(frame, xs, ys)
        assert_hits_equal(hits[sel], z[f"{name}_hits"], what=name)
        assert int((hits["prim"] != MISS).sum()) == int(z[f"{name}_num_hits"])
        assert sha(hits) == str(z[f"{name}_hits_sha"]), name
        vis = g.trace_shadow(frame, xs, ys, hits, LIGHT0)
        assert int(vis.sum()) == int(z[f"{name}_num_visible"])
        assert sha(vis) == str(z[f"{name}_vis_sha"]), name
        # fused primary + shadow call gives the same bytes
        h2, v2 = g.trace_frame(frame, xs, ys, LIGHT0[None, :])
        assert h2.tobytes() == hits.tobytes() and v2[0].tobytes() == vis.tobytes()


def test_edge_case_rays_match_reference_fixture(teapot):
    _, g = teapot
    z = np.load(f"{GOLDEN}/teapot_edge.npz")
    rays = z["rays"]
    for name, cls in (("tree", CLS_TREE), ("all", ALL)):
        assert_hits_equal(g.intersect(rays, cls), z[name], rays=rays, what=name)


def test_random_rays_with_clips_and_any_hit(teapot, oracle):
    scene, g = teapot
    rng = np.random.RandomState(5)
    n = 200000
    o = (rng.rand(n, 3).astype(np.float32) - np.float32(0.5)) * np.float32(9.0)
    d = rng.randn(n, 3).astype(np.float32)
    d /= np.sqrt((d * d).sum(axis=1, dtype=np.float32))[:, None].astype(np.float32)
    rays = make_rays(o, d)
    rays["clip"][::3] = rng.rand(len(rays["clip"][::3])).astype(np.float32) * 8
    rays["flags"][1::5] = RAY_ANY
    for cls in (ALL, CLS_TREE):
        want = oracle.intersect(scene, rays, cls, nthreads=8)
        assert_hits_equal(g.intersect(rays, cls), want, rays=rays, what=f"random classes={cls}")


def test_analytic_spheres_and_boxes(oracle):
    z = np.load(f"{GOLDEN}/analytic10k.npz")
    spheres, boxes = analytic_scene_arrays(4, 10000)
    scene = Scene(spheres=spheres, boxes=boxes)
    w, h = int(z["width"]), int(z["height"])
    xs, ys = oracle.ray_tables(w, h)
    with upload(scene) as g:
        hits = g.trace_primary(capi.Frame.make(w, h, classes=CLS_SPHERE), xs, ys)
        assert_hits_equal(hits, z["sphere_hits"], what="10k spheres")
        vis = g.trace_shadow(capi.Frame.make(w, h, classes=CLS_SPHERE), xs, ys, hits, LIGHT0)
        assert vis.tobytes() == z["sphere_vis"].tobytes()
        bh = g.trace_primary(capi.Frame.make(w, h, classes=CLS_SPHERE | CLS_BOX), xs, ys)
        assert_hits_equal(bh, z["sphere_box_hits_oracle"], what="spheres+boxes")
    # partial last lane: 10 spheres, 3 boxes
    small = Scene(spheres=spheres[:10], boxes=boxes[:3])
    rays = oracle.primary_rays(96, 54)
    with upload(small) as g:
        assert_hits_equal(g.intersect(rays, CLS_SPHERE | CLS_BOX), oracle.intersect(small, rays, CLS_SPHERE | CLS_BOX))


def test_tile_split_is_byte_identical_to_single_gpu(teapot, oracle):
    """image-tile split: N ranks' compact results re-assembled == the one-GPU frame (no cross-tile math)"""
    _, g = teapot
    w, h = 500, 277  # not a multiple of the tile size: partial edge tiles
    xs, ys = oracle.ray_tables(w, h)
    full = g.trace_primary(capi.Frame.make(w, h, classes=ALL), xs, ys)
    fvis = g.trace_shadow(capi.Frame.make(w, h, classes=ALL), xs, ys, full, LIGHT0)
    for world, tile in ((2, (32, 32)), (3, (16, 8)), (8, (64, 16))):
        hits = np.zeros(w * h, capi.HIT_DT)
        vis = np.zeros(w * h, np.uint8)
        for rank in range(world):
            f = capi.Frame.make(w, h, classes=ALL, tile=tile, first_tile=rank, tile_stride=world, compact=1)
            m = capi.frame_pixel_map(f)
            lh, lv = g.trace_frame(f, xs, ys, LIGHT0[None, :])
            ok = m != 0xFFFFFFFF
            assert (lh["prim"][~ok] == MISS).all() and (lv[0][~ok] == 0).all()
            hits[m[ok]] = lh[ok]
            vis[m[ok]] = lv[0][ok]
        assert hits.tobytes() == full.tobytes() and vis.tobytes() == fvis.tobytes()


def test_banded_host_path_at_1080p(teapot, oracle):
    """dodrt_trace_frame copies results band by band while the next band is traced: full-frame bands (whole
    tile rows) and compact bands (runs of local tiles) must reproduce the un-banded device path exactly."""
    import torch
    _, g = teapot
    w, h = 1920, 1080
    xs, ys = oracle.ray_tables(w, h)
    dev = torch.device("cuda:0")
    d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
    d_hits = torch.empty((w * h, 16), dtype=torch.uint8, device=dev)
    d_vis = torch.zeros(w * h, dtype=torch.uint8, device=dev)
    frame = capi.Frame.make(w, h, classes=ALL)
    g.trace_primary_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr())
    g.trace_shadow_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), LIGHT0, d_vis.data_ptr())
    torch.cuda.synchronize()
    want_h, want_v = d_hits.cpu().numpy().tobytes(), d_vis.cpu().numpy().tobytes()
    lights = np.stack([LIGHT0, np.array([4.0, 4.3, 3.3], np.float32)])  # lights[0], lights[1] of main.cpp:284-285
    hits, vis = g.trace_frame(frame, xs, ys, lights)
    assert hits.tobytes() == want_h and vis[0].tobytes() == want_v
    assert vis[1].sum() > 0 and vis[1].tobytes() != vis[0].tobytes()
    full_h = np.frombuffer(want_h, capi.HIT_DT)
    for world in (2,):
        for rank in range(world):
            f = capi.Frame.make(w, h, classes=ALL, first_tile=rank, tile_stride=world, compact=1)
            m = capi.frame_pixel_map(f)
            lh, lv = g.trace_frame(f, xs, ys, LIGHT0[None, :])
            ok = m != 0xFFFFFFFF
            assert lh[ok].tobytes() == full_h[m[ok]].tobytes()
            assert lv[0][ok].tobytes() == np.frombuffer(want_v, np.uint8)[m[ok]].tobytes()
            assert (lh["prim"][~ok] == MISS).all()


def test_device_resident_entry_points(teapot, oracle):
    import torch
    scene, g = teapot
    w, h = 320, 180
    rays = oracle.primary_rays(w, h)
    want = oracle.intersect(scene, rays, ALL, nthreads=8)
    dev = torch.device("cuda:0")
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32)).to(dev)
    d_hits = torch.empty((len(rays), 16), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    before = g.launch_count()
    g.intersect_device(d_rays.data_ptr(), len(rays), ALL, d_hits.data_ptr(), stream)
    torch.cuda.synchronize()
    assert g.launch_count() == before + 1
    assert_hits_equal(d_hits.cpu().numpy().reshape(-1).view(capi.HIT_DT), want)
    xs, ys = oracle.ray_tables(w, h)
    d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
    d_hits.zero_()
    frame = capi.Frame.make(w, h, classes=ALL)
    g.trace_primary_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), stream)
    d_vis = torch.zeros(w * h, dtype=torch.uint8, device=dev)
    g.trace_shadow_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), LIGHT0, d_vis.data_ptr(), stream)
    torch.cuda.synchronize()
    got = d_hits.cpu().numpy().reshape(-1).view(capi.HIT_DT)
    assert_hits_equal(got, want)
    assert d_vis.cpu().numpy().tobytes() == oracle.trace_shadow(scene, w, h, ALL, want, LIGHT0, nthreads=8).tobytes()


def test_empty_inputs_and_empty_tree(oracle):
    with capi.Scene(0) as g:
        rays = oracle.primary_rays(16, 8)
        hits = g.intersect(rays, ALL)  # nothing registered: everything misses
        assert (hits["prim"] == MISS).all() and np.isinf(hits["t"]).all()
        assert len(g.intersect(rays[:0], ALL)) == 0
        # the reference's tree for a missing mesh: one empty leaf (mesh.cpp:17-21, kdtree.cpp:106-110)
        g.set_kdtree(np.array([3], np.uint64), np.zeros((0, 72), np.float32), np.array([np.inf] * 3 + [-np.inf] * 3))
        assert (g.intersect(rays, CLS_TREE)["prim"] == MISS).all()
        with pytest.raises(capi.DodrtError):  # leaf pointing past the lane array
            g.set_kdtree(np.array([3 | (2 << 2)], np.uint64), np.zeros((1, 72), np.float32), np.zeros(6))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_against_the_reference_code_directly():
    """GPU vs the reference's own translation units (no restatement in between)."""
    import os
    ref = RefLib()
    w, h = 480, 270
    ref.set_config(w, h)
    ref.add_reference_spheres(1, 16)
    ref.add_reference_planes()
    ref.add_reference_cylinder()
    ref.add_mesh(os.path.join(GOLDEN, "teapot.dodm"))
    ref.build_tree()
    nodes, lanes, prim, bounds = ref.export_tree()
    scene = Scene(nodes, lanes, bounds, spheres=ref.export_spheres()[:, :4], planes=reference_planes(),
                  cylinders=reference_cylinder())
    rays = ref.primary_rays(w, h)
    with upload(scene) as g:
        for cls in (CLS_TREE, ALL):
            want = ref.intersect(rays, cls, 8)
            got = g.intersect(rays, cls)
            assert_hits_equal(got, want, what=f"vs reference classes={cls}")
            pts = hit_points(rays, want["t"])
            sr = ref.shadow_rays(pts, LIGHT0)
            assert (g.intersect(sr, cls)["prim"] == ref.intersect(sr, cls, 8)["prim"]).all()


# ---- shading + 10-bounce loop on the GPU vs the reference's own rayTrace --------------------
def _render_scene():
    from dod_raytracer_b200 import host
    hs = host.HostScene()
    hs.add_reference_scene(1, 16)
    hs.add_mesh_file(f"{GOLDEN}/teapot.dodm")
    hs.build_tree()
    return hs


def test_full_reference_frame_pixels_within_one_255th():
    """north_star: 'final 8-bit pixels must be within 1/255'.  Fixture = the UNMODIFIED reference rayTrace
    (9 lights, 10 mirror bounces, spheres + planes + cylinder + teapot kd-tree) rendered by oracle/_ref."""
    from dod_raytracer_b200 import host, workloads
    want = np.load(f"{GOLDEN}/teapot_render_240x135.npz")["rgb"].astype(np.int32)
    h, w = want.shape[:2]
    xs, ys = host.ray_tables(w, h)
    for variant in [v for v in (3, 0, 4, 6, 7) if capi.variant_available(v)]:
        with _render_scene().upload(0, shading=True) as g:
            g.set_kernel_variant(variant)
            got = g.render(capi.Frame.make(w, h, classes=ALL), xs, ys, workloads.REFERENCE_LIGHTS,
                           workloads.REFERENCE_DEPTH).astype(np.int32)
        diff = np.abs(got - want)
        assert diff.max() <= 1, f"variant {variant}: max pixel difference {diff.max()} at {np.argwhere(diff > 1)[:5]}"
        assert (diff == 0).mean() > 0.999, f"variant {variant}: only {(diff == 0).mean():.5f} of the channels identical"


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_render_against_live_reference_other_sizes_and_depths():
    from dod_raytracer_b200 import host, workloads
    ref = RefLib()
    ref.add_reference_spheres(1, 16)
    ref.add_reference_planes()
    ref.add_reference_cylinder()
    ref.add_mesh(f"{GOLDEN}/teapot.dodm")
    ref.build_tree()
    with _render_scene().upload(0, shading=True) as g:
        for (w, h) in ((160, 90), (97, 61)):
            want = ref.render(w, h, nthreads=1).astype(np.int32)  # one band: canonical raster tables
            xs, ys = host.ray_tables(w, h)
            got = g.render(capi.Frame.make(w, h, classes=ALL), xs, ys, workloads.REFERENCE_LIGHTS, 10).astype(np.int32)
            assert np.abs(got - want).max() <= 1
        # depth 1 (primary + 9 shadow passes) is still a sane image and differs from depth 10
        d1 = g.render(capi.Frame.make(160, 90, classes=ALL), *host.ray_tables(160, 90), workloads.REFERENCE_LIGHTS, 1)
        assert d1.mean() > 10 and np.abs(d1.astype(np.int32) - ref.render(160, 90, nthreads=1)).max() > 1


def test_gpu_side_lane_reorder_matches_host_reorder():
    """f-4: dodrt_scene_set_kdtree_indexed / _set_shading_indexed (Triangle::reorderLanesByIndices, triangle.cpp:349-367,
    done by the upload kernels from the lanes in creation order + m_primNums) give the same scene as the host-side
    re-order: identical hit records, visibility bytes and rendered pixels."""
    from dod_raytracer_b200 import host, workloads

    def scene(keep):
        hs = host.HostScene()
        hs.add_reference_scene(1, 16)
        hs.add_mesh_file(f"{GOLDEN}/teapot.dodm")
        hs.build_tree(keep_creation_order=keep)
        return hs

    w, h = 480, 270
    xs, ys = host.ray_tables(w, h)
    out = []
    for keep in (False, True):
        hs = scene(keep)
        assert hs.sizes().num_lanes == 3021 and hs.sizes().num_orig_lanes == 790  # teapot pins
        with hs.upload(0, shading=True) as g:
            frame = capi.Frame.make(w, h, classes=ALL)
            hits, vis = g.trace_frame(frame, xs, ys, LIGHT0[None, :])
            rgb = g.render(frame, xs, ys, workloads.REFERENCE_LIGHTS, 3)
            out.append((hits.tobytes(), vis.tobytes(), rgb.tobytes()))
    assert out[0][0] == out[1][0], "hit records differ"
    assert out[0][1] == out[1][1], "visibility differs"
    assert out[0][2] == out[1][2], "rendered pixels differ"
    assert (np.frombuffer(out[0][0], capi.HIT_DT)["prim"] >> 29 == 0).sum() > 1000  # the teapot is in view


def test_donated_rays_resume_bit_identically(monkeypatch, oracle):
    """Variant 7 (dodrt_donate.inl).  DODRT_DONATE_ALWAYS makes every warp suspend its live rays at every poll, so
    nearly every ray that traverses more than a few nodes is finished by the 32-lane resume path; all four finishers
    (hit record, any-hit record, visibility byte of both shadow modes) must give the bits of the plain kernel."""
    from dod_raytracer_b200 import host, workloads
    scene = teapot_scene(full=True)
    rays = oracle.primary_rays(480, 270)
    anyrays = rays.copy()
    anyrays["flags"] = RAY_ANY
    anyrays["clip"] = 6.0
    w, h = 640, 360
    xs, ys = host.ray_tables(w, h)
    frame = capi.Frame.make(w, h, classes=ALL)
    results = []
    for always in ("0", "1"):
        monkeypatch.setenv("DODRT_DONATE_ALWAYS", always)
        with upload(scene) as g:
            g.set_kernel_variant(7 if always == "1" else 3)
            hits, vis = g.trace_frame(frame, xs, ys, LIGHT0[None, :])
            results.append((g.intersect(rays, ALL).tobytes(), g.intersect(anyrays, ALL).tobytes(), hits.tobytes(), vis.tobytes()))
    for a, b, what in zip(results[0], results[1], ("closest", "any-hit", "frame hits", "frame visibility")):
        assert a == b, what
    monkeypatch.setenv("DODRT_DONATE_ALWAYS", "1")
    with _render_scene().upload(0, shading=True) as g:  # kModeShadowRays finisher (bounce loop)
        g.set_kernel_variant(7)
        got = g.render(capi.Frame.make(160, 90, classes=ALL), *host.ray_tables(160, 90), workloads.REFERENCE_LIGHTS, 4)
    monkeypatch.setenv("DODRT_DONATE_ALWAYS", "0")
    with _render_scene().upload(0, shading=True) as g:
        g.set_kernel_variant(3)
        want = g.render(capi.Frame.make(160, 90, classes=ALL), *host.ray_tables(160, 90), workloads.REFERENCE_LIGHTS, 4)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("kind", ["lattice", "duplicates"])
def test_tie_rules_on_lattice_and_duplicated_triangles(kind, oracle, monkeypatch):
    """SURVEY A.5 on the GPU: equal t inside a lane, between the lanes of a leaf and between leaves (lattice triangles
    with shared edges; every triangle present twice, some three times) -- every variant of this build, and the donating
    kernel with every ray through the queue, must report the id, t, u, v of the oracle (which tests/test_oracle_vs_ref.py
    pins against the reference's own code on the same scenes and rays)."""
    from dod_raytracer_b200 import host
    from gpu_util import oracle_scene
    from scenes import tie_scene_rays, tie_scene_triangles
    pos, idx = tie_scene_triangles(kind)
    hs = host.HostScene()
    hs.add_mesh(pos, idx)
    hs.build_tree()
    scene = oracle_scene(hs)
    rays = tie_scene_rays()
    want = oracle.intersect(scene, rays, CLS_TREE, nthreads=8)
    assert ((want["prim"] != MISS) & ((rays["flags"] & RAY_ANY) == 0)).sum() > 3000
    for variant, always in [(v, "0") for v in VARIANTS] + [(7, "1")]:
        monkeypatch.setenv("DODRT_DONATE_ALWAYS", always)
        with upload(scene) as g:
            g.set_kernel_variant(variant)
            assert_hits_equal(g.intersect(rays, CLS_TREE), want, rays, f"{kind}, variant {variant}, donate-always {always}")


@pytest.mark.parametrize("nthreads", [4, 8])
def test_concurrent_donating_launches(oracle, nthreads):
    """Several host threads trace frames with the donating kernel (variant 7) at the same time, each on its own CUDA
    stream: the grids together exceed what the GPU can keep resident, so blocks of one launch start only when blocks
    of another retire.  The helper loop must not wait for blocks that have not started (it counts warps at kernel
    entry); every call has to return, with the single-launch bits.  The threads issue their FIRST launch on a fresh
    scene together (barrier): that is when the persistent donation queues are created and handed out, and a queue
    must never serve two launches in flight (8 threads > 4 queues: the rest take the per-launch pool path)."""
    import threading
    import torch
    from dod_raytracer_b200 import host
    scene = teapot_scene(full=True)
    w, h = 512, 288
    xs, ys = host.ray_tables(w, h)
    frame = capi.Frame.make(w, h, classes=ALL)
    dev = torch.device("cuda:0")
    d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
    light = np.asarray(LIGHT0, np.float32)
    with upload(scene) as g0:
        g0.set_kernel_variant(3)
        want_hits, want_vis = g0.trace_frame(frame, xs, ys, LIGHT0[None, :])
    with upload(scene) as g:
        g.set_kernel_variant(7)
        results, errors = {}, []
        gate = threading.Barrier(nthreads)

        def work(k):
            try:
                st = torch.cuda.Stream(device=dev)
                d_hits = torch.empty((w * h, 16), dtype=torch.uint8, device=dev)
                d_vis = torch.empty(w * h, dtype=torch.uint8, device=dev)
                gate.wait(timeout=60)
                for rep in range(8):
                    g.trace_primary_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), st.cuda_stream)
                    g.trace_shadow_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), d_hits.data_ptr(), light, d_vis.data_ptr(),
                                          st.cuda_stream)
                st.synchronize()
                results[k] = (d_hits.cpu().numpy().tobytes(), d_vis.cpu().numpy().tobytes())
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        threads = [threading.Thread(target=work, args=(k,), daemon=True) for k in range(nthreads)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=120)
            assert not t.is_alive(), "a donating launch did not return (helpers waiting for non-resident blocks?)"
        assert not errors, errors
        assert len(results) == nthreads
        for k, (hits, vis) in results.items():
            assert hits == want_hits.tobytes(), k
            assert vis == want_vis[0].tobytes(), k
