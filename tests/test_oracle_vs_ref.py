"""CPU: the restatement against the reference's own code, live (needs oracle/_ref/libdodrt_ref.so, which is
built from /root/reference by oracle/Makefile in the build container and travels to the GPU box)."""
import os

import numpy as np
import pytest

from oracle_api import (CLS_CYLINDER, CLS_PLANE, CLS_SPHERE, CLS_TREE, MISS, RAY_ANY, RefLib, Scene, have_ref,
                        reference_cylinder, reference_planes, same_bits)
from scenes import GOLDEN, LIGHT0, analytic_scene_arrays, hit_points, make_rays

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference here)")]
ALL = CLS_SPHERE | CLS_PLANE | CLS_CYLINDER | CLS_TREE


@pytest.fixture(scope="module")
def teapot_ref():
    ref = RefLib()
    ref.set_config(640, 360)
    ref.add_reference_spheres(1, 16)
    ref.add_reference_planes()
    ref.add_reference_cylinder()
    assert ref.add_mesh(os.path.join(GOLDEN, "teapot.dodm")) == 6320
    ref.build_tree()
    nodes, lanes, prim, bounds = ref.export_tree()
    scene = Scene(nodes, lanes, bounds, spheres=ref.export_spheres()[:, :4], planes=reference_planes(),
                  cylinders=reference_cylinder())
    return ref, scene


def test_restatement_matches_reference_per_class(teapot_ref, oracle):
    ref, scene = teapot_ref
    w, h = 640, 360
    rr = ref.primary_rays(w, h)
    assert rr.tobytes() == oracle.primary_rays(w, h).tobytes()
    for cls in (CLS_TREE, CLS_SPHERE, CLS_PLANE, CLS_CYLINDER, ALL, CLS_SPHERE | CLS_TREE, CLS_PLANE | CLS_TREE):
        want = ref.intersect(rr, cls, 4)
        got = oracle.intersect(scene, rr, cls, nthreads=4)
        assert got.tobytes() == want.tobytes(), f"classes={cls}"
        # shadow rays built from the reference's hit points, answered by both
        pts = hit_points(rr, want["t"])
        sr = ref.shadow_rays(pts, LIGHT0)
        hitmask = want["prim"] != MISS
        assert sr[hitmask][:2000].tobytes() == oracle.shadow_rays(pts[hitmask][:2000], LIGHT0).tobytes()
        occ_ref = ref.intersect(sr[hitmask], cls, 4)
        occ = oracle.intersect(scene, sr[hitmask], cls, nthreads=4)
        assert (occ["prim"] == occ_ref["prim"]).all()


def test_random_rays_all_classes(teapot_ref, oracle):
    ref, scene = teapot_ref
    rng = np.random.RandomState(11)
    n = 60000
    o = (rng.rand(n, 3).astype(np.float32) - np.float32(0.5)) * np.float32(9.0)
    d = rng.randn(n, 3).astype(np.float32)
    d /= np.sqrt((d * d).sum(axis=1, dtype=np.float32))[:, None].astype(np.float32)
    rays = make_rays(o, d)
    rays["clip"][::3] = rng.rand(len(rays["clip"][::3])).astype(np.float32) * 8
    rays["flags"][1::5] = RAY_ANY
    want = ref.intersect(rays, ALL, 4)
    got = oracle.intersect(scene, rays, ALL, nthreads=4)
    assert (got["prim"] == want["prim"]).all()
    closest = (rays["flags"] & RAY_ANY) == 0
    assert got[closest].tobytes() == want[closest].tobytes()
    assert (want["prim"][closest] != MISS).sum() > 10000


@pytest.mark.parametrize("kind", ["unnormalised", "axis-parallel", "nan-inf-zero"])
def test_irregular_rays_all_classes(kind, teapot_ref, oracle):
    """Rays the reference's arithmetic was not written for, but answers deterministically: directions of length 1e-3 ...
    1e3 (dodrt_ray.d need not be normalised), one or two zero components (infinite slab inverses, 0 * inf = NaN in
    box.cpp:38-47), NaN / inf / all-zero directions and NaN origins.  The restatement must follow the reference's
    compares bit for bit on every class (the GPU path is compared with the restatement on the same kinds of rays in
    tests/test_gpu_fuzz.py)."""
    from scenes import irregular_rays
    ref, scene = teapot_ref
    rays = irregular_rays(kind)
    closest = (rays["flags"] & RAY_ANY) == 0
    for cls in (CLS_TREE, CLS_SPHERE, CLS_PLANE, CLS_CYLINDER, ALL):
        want = ref.intersect(rays, cls, 4)
        got = oracle.intersect(scene, rays, cls, nthreads=4)
        assert (got["prim"] == want["prim"]).all(), f"{kind}, classes={cls}"
        assert got[closest].tobytes() == want[closest].tobytes(), f"{kind}, classes={cls}"
        assert (want["prim"] != MISS).sum() > 400


@pytest.mark.parametrize("kind", ["lattice", "duplicates"])
def test_tie_rules_on_lattice_and_duplicated_triangles(kind, oracle, tmp_path):
    """equal t inside a lane, between lanes of a leaf and between leaves (a triangle referenced by several
    leaves, or present twice) must resolve to the id the reference's slot-by-slot loop and leaf order leave behind.
    Triangles on a coarse lattice (shared edges and vertices, axis-parallel faces) or every triangle twice; rays along
    lattice lines, through lattice points and at random; ids, t, u, v of the restatement against the reference's code."""
    from dod_raytracer_b200 import host
    from scenes import tie_scene_rays, tie_scene_triangles
    pos, idx = tie_scene_triangles(kind)
    path = str(tmp_path / f"{kind}.dodm")
    host.write_dodm(path, pos, idx)
    ref = RefLib()
    ref.set_config(64, 64)
    assert ref.add_mesh(path) == len(idx)
    ref.build_tree()
    nodes, lanes, prim, bounds = ref.export_tree()
    scene = Scene(nodes, lanes, bounds)
    rays = tie_scene_rays()
    want = ref.intersect(rays, CLS_TREE, 4)
    got = oracle.intersect(scene, rays, CLS_TREE, nthreads=4)
    assert (got["prim"] == want["prim"]).all()
    closest = (rays["flags"] & RAY_ANY) == 0
    assert got[closest].tobytes() == want[closest].tobytes()
    assert (want["prim"][closest] != MISS).sum() > 3000


def test_reference_hit_record_is_rebuilt_from_prim_u_v(teapot_ref, oracle):
    """The C ABI returns (t, prim, u, v); the reference's HitRecord (hitrecord.h) must follow from it:
    hitPoint = o + d*t and hitNormal = mat3(AN,BN,CN) * (1-(u+v), u, v) (triangle.cpp:170-174)."""
    ref, scene = teapot_ref
    rr = ref.primary_rays(640, 360)
    hits = ref.intersect(rr, CLS_TREE, 4)
    recs = ref.intersect_records(rr, CLS_TREE, 4)
    normals = ref.export_normals()
    m = hits["prim"] != MISS
    assert (recs["hit"][m] == 1).all() and (recs["hit"][~m] == 0).all()
    assert same_bits(recs["t"][m], hits["t"][m]).all()
    assert same_bits(recs["point"][m], hit_points(rr, hits["t"])[m]).all()
    ids = hits["prim"][m] & ((1 << 29) - 1)
    u, v = hits["u"][m], hits["v"][m]
    b0 = np.float32(1.0) - (u + v)
    n9 = normals[ids]
    for k in range(3):
        want = (n9[:, 0 + k] * b0 + n9[:, 3 + k] * u) + n9[:, 6 + k] * v
        assert same_bits(recs["normal"][m][:, k], want.astype(np.float32)).all()
    assert np.allclose(recs["color"][m], [0.1, 0.8, 0.3])  # mesh.cpp:23


def test_ten_thousand_spheres_live(oracle):
    spheres, _ = analytic_scene_arrays(4, 10000)
    ref = RefLib()
    ref.add_spheres(spheres)
    rays = ref.primary_rays(128, 72)
    want = ref.intersect(rays, CLS_SPHERE, 4)
    got = oracle.intersect(Scene(spheres=spheres), rays, CLS_SPHERE, nthreads=4)
    assert got.tobytes() == want.tobytes()
    # partial last lane (sphere.cpp:31-37): 10 spheres
    ref2 = RefLib()
    ref2.add_spheres(spheres[:10])
    r2 = ref2.primary_rays(64, 36)
    assert oracle.intersect(Scene(spheres=spheres[:10]), r2, CLS_SPHERE).tobytes() == ref2.intersect(r2, CLS_SPHERE).tobytes()


def test_empty_tree_and_empty_scene(oracle):
    """mesh.cpp:17-21: a missing mesh still builds an (empty, one-leaf) tree that never hits"""
    ref = RefLib()
    ref.build_tree()
    nodes, lanes, prim, bounds = ref.export_tree()
    assert len(nodes) == 1 and len(lanes) == 0
    rays = ref.primary_rays(32, 18)
    want = ref.intersect(rays, CLS_TREE)
    got = oracle.intersect(Scene(nodes, lanes, bounds), rays, CLS_TREE)
    assert (want["prim"] == MISS).all() and got.tobytes() == want.tobytes()


def test_reference_frame_loop_matches_golden_visibility():
    """ref_trace_frame runs the reference's per-pixel order (closest chain, then canSeeLight from the record's OWN
    hitPoint, main.cpp:182-219); its visibility must equal the fixture that was built from o + d*t hit points,
    i.e. every shape class's hitPoint really is o + d*t bit for bit."""
    z = np.load(os.path.join(GOLDEN, "teapot_frame.npz"))
    w, h, step = int(z["width"]), int(z["height"]), int(z["step"])
    ref = RefLib()
    ref.add_reference_spheres(1, 16)
    ref.add_reference_planes()
    ref.add_reference_cylinder()
    ref.add_mesh(os.path.join(GOLDEN, "teapot.dodm"))
    ref.build_tree()
    t, vis = ref.trace_frame(w, h, ALL, LIGHT0, 8)
    sel = (np.arange(0, h, step)[:, None] * w + np.arange(0, w, step)[None, :]).ravel()
    assert vis[sel].tobytes() == z["all_vis"].tobytes()
    assert same_bits(t[sel], z["all_hits"]["t"]).all()
    assert int(vis.sum()) == int(z["all_num_visible"])
