"""GPU (-m gpu): randomised scenes and rays -- triangle soups with degenerate / duplicated / axis-aligned
triangles, ragged lane counts, tiny and huge coordinates, rays with zero direction components and origins on
vertices -- built with the product's host builder and answered by the CUDA path and by the oracle.  Bit-exact."""
import numpy as np
import pytest

from dod_raytracer_b200 import capi, host
from gpu_util import assert_hits_equal
from oracle_api import CLS_BOX, CLS_CYLINDER, CLS_PLANE, CLS_SPHERE, CLS_TREE, CYL_DT, RAY_ANY, Scene, pack_lanes
from scenes import make_rays

pytestmark = pytest.mark.gpu
EVERYTHING = CLS_SPHERE | CLS_PLANE | CLS_CYLINDER | CLS_TREE | CLS_BOX


def random_soup(rng, n, scale):
    """n triangles around the origin; ~10 % degenerate (zero area), ~10 % duplicates, ~10 % axis-aligned."""
    c = (rng.rand(n, 1, 3).astype(np.float32) - np.float32(0.5)) * np.float32(4.0 * scale)
    tri = c + (rng.rand(n, 3, 3).astype(np.float32) - np.float32(0.5)) * np.float32(0.8 * scale)
    k = max(1, n // 10)
    tri[:k, 2] = tri[:k, 1]                      # degenerate: two equal vertices
    tri[k:2 * k] = tri[2 * k:3 * k]              # exact duplicates
    tri[3 * k:4 * k, :, 2] = tri[3 * k:4 * k, :1, 2]  # flat in z (axis-aligned plane)
    pos = tri.reshape(-1, 3)
    idx = np.arange(n * 3, dtype=np.uint32).reshape(n, 3)
    return pos, idx


def random_rays(rng, n, scale, verts):
    o = (rng.rand(n, 3).astype(np.float32) - np.float32(0.5)) * np.float32(8.0 * scale)
    d = rng.randn(n, 3).astype(np.float32)
    d[::7, 0] = 0.0                               # axis-parallel components: inv = +-inf, NaN slabs
    d[::11, 1] = 0.0
    d[::13, 2] = -0.0
    d[5::97] = [0.0, 0.0, 1.0]
    norm = np.sqrt((d * d).sum(axis=1, dtype=np.float32)).astype(np.float32)
    norm[norm == 0] = 1
    d = (d / norm[:, None]).astype(np.float32)
    o[3::17] = verts[rng.randint(0, len(verts), size=len(o[3::17]))]  # origins exactly on mesh vertices
    rays = make_rays(o, d)
    rays["clip"][::3] = (rng.rand(len(rays["clip"][::3])) * 6 * scale).astype(np.float32)
    rays["clip"][4::29] = 0.0
    rays["flags"][1::5] = RAY_ANY
    return rays


@pytest.mark.parametrize("seed,ntri,scale", [(1, 5, 1.0), (2, 37, 1.0), (3, 400, 1.0), (4, 3000, 1.0), (5, 1500, 1e-3),
                                             (6, 1500, 1e4), (7, 9, 1.0), (8, 20000, 1.0)])
def test_random_scene_matches_oracle(oracle, seed, ntri, scale):
    rng = np.random.RandomState(seed)
    pos, idx = random_soup(rng, ntri, scale)
    hs = host.HostScene()
    hs.add_mesh(pos, idx)
    nsph, nbox, npl = rng.randint(0, 20), rng.randint(0, 20), rng.randint(0, 10)
    spheres = np.concatenate([(rng.rand(nsph, 3) - 0.5) * 6 * scale, rng.rand(nsph, 1) * 0.5 * scale], axis=1).astype(np.float32)
    lo = ((rng.rand(nbox, 3) - 0.5) * 6 * scale).astype(np.float32)
    boxes = np.concatenate([lo, lo + (rng.rand(nbox, 3) * scale).astype(np.float32)], axis=1).astype(np.float32)
    nrm = rng.randn(npl, 3).astype(np.float32)
    nrm /= np.sqrt((nrm * nrm).sum(axis=1, keepdims=True)).astype(np.float32) if npl else 1
    planes = np.concatenate([((rng.rand(npl, 3) - 0.5) * 8 * scale).astype(np.float32), nrm], axis=1).astype(np.float32)
    for s in spheres:
        hs.add_sphere(s[:3], float(s[3]))
    for b in boxes:
        hs.add_box(b[:3], b[3:])
    hs.build_tree()
    a = hs.arrays()
    cyl = np.zeros(1, CYL_DT)
    cyl["base"], cyl["axis"], cyl["radius_sq"], cyl["height"] = [0.3 * scale, -scale, 0], [0, 1, 0], (0.4 * scale) ** 2, 2 * scale
    scene = Scene(a["nodes"], a["tri_lanes"], a["bounds"], spheres=spheres, planes=planes, cylinders=cyl, boxes=boxes)
    assert scene.sphere_lanes.tobytes() == a["sphere_lanes"].tobytes() and scene.box_lanes.tobytes() == a["box_lanes"].tobytes()
    rays = random_rays(rng, 40000, scale, pos)
    g = hs.upload(0)
    g.set_planes(pack_lanes(planes), npl)
    g.set_cylinders(cyl)
    try:
        for variant in [v for v in (3, 0, 4, 5, 6, 7, 8) if capi.variant_available(v)]:
            g.set_kernel_variant(variant)
            for cls in (EVERYTHING, CLS_TREE):
                want = oracle.intersect(scene, rays, cls, nthreads=8)
                assert_hits_equal(g.intersect(rays, cls), want, rays=rays, what=f"seed {seed} variant {variant} classes {cls}")
    finally:
        g.close()
    hit_rate = (want["prim"] != 0xFFFFFFFF).mean()
    assert 0.0 <= hit_rate <= 1.0


@pytest.mark.parametrize("seed,count,scale", [(11, 64, 1.0), (12, 777, 1.0), (13, 5000, 1.0), (14, 3000, 1e-2), (15, 3000, 1e3)])
def test_sphere_box_culling_bvh_is_exact(oracle, seed, count, scale):
    """>= 64 spheres / boxes switch the CUDA path from the reference's brute force to the culling BVH
    (dodrt_prim_bvh.cuh); the answer must stay the brute-force answer (lexicographic min of (t, id)), including
    overlapping and nested primitives, origins inside primitives, ties between identical primitives."""
    rng = np.random.RandomState(seed)
    c = ((rng.rand(count, 3) - 0.5) * 9 * scale).astype(np.float32)
    r = ((0.03 + rng.rand(count, 1) * 0.3) * scale).astype(np.float32)
    spheres = np.concatenate([c, r], axis=1).astype(np.float32)
    spheres[1::50] = spheres[0::50][: len(spheres[1::50])]  # exact duplicates: the lower id must win
    spheres[7::90, 3] *= 12  # a few big ones that contain many others (and many ray origins)
    lo = ((rng.rand(count, 3) - 0.5) * 9 * scale).astype(np.float32)
    boxes = np.concatenate([lo, lo + ((0.02 + rng.rand(count, 3) * 0.5) * scale).astype(np.float32)], axis=1).astype(np.float32)
    boxes[3::40] = boxes[2::40][: len(boxes[3::40])]
    boxes[5::120, 3:] += np.float32(4 * scale)
    scene = Scene(spheres=spheres, boxes=boxes)
    rays = random_rays(rng, 60000, scale, c)
    # dodrt_ray.d need not be normalised (include/dodrt.h): with |D| != 1 the reference's sphere arithmetic
    # (tca = L.D, d2 = |L|^2 - tca^2, sphere.cpp:62-90) accepts spheres the geometric line misses, which a culling
    # structure would skip -- such rays must get the brute-force answer too.  Lengths 1e-3 .. 1e3, a hair off 1, 0.
    loose = random_rays(rng, 30000, scale, c)
    f = np.float32(10.0) ** rng.uniform(-3, 3, len(loose)).astype(np.float32)
    f[::4] = np.float32(1.0) + rng.uniform(-1e-3, 1e-3, len(f[::4])).astype(np.float32)
    f[1::64] = 0.0
    loose["d"] = (loose["d"] * f[:, None]).astype(np.float32)
    with capi.Scene(0) as g:
        g.set_spheres(scene.sphere_lanes, count)
        g.set_boxes(scene.box_lanes, count)
        for cls in (CLS_SPHERE, CLS_BOX, CLS_SPHERE | CLS_BOX):
            want = oracle.intersect(scene, rays, cls, nthreads=8)
            assert_hits_equal(g.intersect(rays, cls), want, rays=rays, what=f"seed {seed} classes {cls}")
            assert (want["prim"] != 0xFFFFFFFF).mean() > 0.02
            want = oracle.intersect(scene, loose, cls, nthreads=8)
            assert_hits_equal(g.intersect(loose, cls), want, rays=loose, what=f"seed {seed} classes {cls} un-normalised")
    # a NaN / inf primitive cannot be boxed: the class falls back to the reference's loop, answers unchanged
    bad = spheres.copy()
    bad[5, 0], bad[9, 3] = np.nan, np.inf
    bscene = Scene(spheres=bad, boxes=boxes)
    with capi.Scene(0) as g:
        g.set_spheres(bscene.sphere_lanes, count)
        want = oracle.intersect(bscene, rays[:20000], CLS_SPHERE, nthreads=8)
        assert_hits_equal(g.intersect(rays[:20000], CLS_SPHERE), want, rays=rays[:20000], what=f"seed {seed} NaN sphere")


@pytest.mark.parametrize("kind", ["unnormalised", "axis-parallel", "nan-inf-zero"])
def test_irregular_rays_all_classes(kind, oracle):
    """Un-normalised directions against the kd-tree and every analytic class, axis-parallel rays, NaN / inf / zero
    directions and NaN origins (tests/scenes.py::irregular_rays; the oracle is pinned against the reference's own code
    on exactly these rays in tests/test_oracle_vs_ref.py): every variant of this build, ids and t, u, v bit-exact --
    NaN results included."""
    from gpu_util import upload
    from scenes import irregular_rays, teapot_scene
    scene = teapot_scene(full=True)
    rays = irregular_rays(kind)
    everything = CLS_SPHERE | CLS_PLANE | CLS_CYLINDER | CLS_TREE
    with upload(scene) as g:
        for cls in (CLS_TREE, CLS_SPHERE, CLS_PLANE, CLS_CYLINDER, everything):
            want = oracle.intersect(scene, rays, cls, nthreads=8)
            for variant in [v for v in (3, 0, 4, 5, 6, 7, 8) if capi.variant_available(v)]:
                g.set_kernel_variant(variant)
                assert_hits_equal(g.intersect(rays, cls), want, rays=rays, what=f"{kind} variant {variant} classes {cls}")
