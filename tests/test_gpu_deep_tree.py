"""GPU (-m gpu): parity on the DEEP trees the headline numbers are quoted on (BASELINE.json configs[1], [2], [4]).

The teapot tree is 10 levels deep; the stand-in dragon (871,200 triangles) builds 17 levels / 68,601 nodes with leaves
of up to 99 lanes, and the 16-instance grid 21 levels / ~1.0 M nodes.  Only there do the `bigTree` auto-variant switch
(kBigTreeNodes), the donation queue's stack limit (kDonateMaxStack = 16 entries: deeper rays are not donated) and the
long leaves come into play.  Every comparison is on ids AND t, u, v bit patterns (tolerance 0 ulp) against the oracle
restatement, and -- where oracle/_ref travelled -- against the reference's own translation units
(KDTree::intersect, /root/reference/src/accelerators/kdtree.cpp:263-361; Triangle::intersectInRange,
src/shapes/triangle.cpp:22-177; the analytic chain of main.cpp:314-321 / 198-217)."""
import os

import numpy as np
import pytest

from dod_raytracer_b200 import capi, host, workloads
from gpu_util import assert_hits_equal, oracle_scene
from oracle_api import CLS_CYLINDER, CLS_PLANE, CLS_SPHERE, CLS_TREE, RAY_ANY, RefLib, have_ref
from scenes import LIGHT0, make_rays

pytestmark = pytest.mark.gpu
ALL = CLS_SPHERE | CLS_PLANE | CLS_CYLINDER | CLS_TREE
# (variant, DODRT_DONATE_ALWAYS): auto picks per launch (donating kernel for short passes over a big tree); 7 + ALWAYS
# suspends every live ray with <= 16 stack entries at every poll, so the deep rays are the ones that stay behind
MODES = [(-1, "0"), (3, "0"), (7, "0"), (7, "1")]


def _build(name):
    w = workloads.WORKLOADS[name]
    hs = workloads.build_host_scene(w)
    return w, hs, oracle_scene(hs)


@pytest.fixture(scope="module")
def dragon():
    return _build("dragon4k")


@pytest.fixture(scope="module")
def dragon16():
    return _build("dragon16_8k")


def _inside_rays(rng, bounds, n):
    """random rays between points of the (slightly inflated) tree bounds: long paths through many leaves, deep stacks"""
    lo, hi = bounds[:3], bounds[3:]
    ext = (hi - lo) * np.float32(0.55)
    ctr = (lo + hi) * np.float32(0.5)
    a = (ctr + (rng.rand(n, 3).astype(np.float32) * 2 - 1) * ext).astype(np.float32)
    b = (ctr + (rng.rand(n, 3).astype(np.float32) * 2 - 1) * ext).astype(np.float32)
    d = b - a
    dist = np.sqrt((d * d).sum(axis=1, dtype=np.float32)).astype(np.float32)
    d = (d / dist[:, None]).astype(np.float32)
    rays = make_rays(a, d)
    rays["clip"][::4] = dist[::4]            # clipped at the target point
    rays["flags"][1::4] = RAY_ANY            # any-hit (shadow-like), unclipped
    rays["flags"][2::8] = RAY_ANY
    rays["clip"][2::8] = dist[2::8]          # any-hit with clip = light distance
    return rays


def _check_scene(monkeypatch, oracle, w, hs, scene, width, height, min_depth, want_deep_stack):
    z = hs.sizes()
    assert z.max_depth >= min_depth, f"tree depth {z.max_depth}: not the deep tree this test is about"
    xs, ys = host.ray_tables(width, height)
    want_hits, ctr = oracle.trace_primary(scene, width, height, w.classes, counters=True, nthreads=8)
    want_vis = oracle.trace_shadow(scene, width, height, w.classes, want_hits, LIGHT0, nthreads=8)
    tri_hits = int(((want_hits["prim"] >> 29) == 0).sum())
    assert tri_hits > 1000, "the mesh must be in view"
    rng = np.random.RandomState(17)
    inside = _inside_rays(rng, hs.arrays(raw=True)["bounds"], 60000)
    want_inside, ctr_in = oracle.intersect(scene, inside, CLS_TREE, counters=True, nthreads=8)
    assert int(ctr_in["max_stack"].max()) >= want_deep_stack, int(ctr_in["max_stack"].max())
    frame = capi.Frame.make(width, height, classes=w.classes)
    for variant, always in MODES:
        monkeypatch.setenv("DODRT_DONATE_ALWAYS", always)
        with hs.upload(0) as g:
            g.set_kernel_variant(variant)
            hits, vis = g.trace_frame(frame, xs, ys, LIGHT0[None, :])
            what = f"{w.name} variant {variant} always {always}"
            assert_hits_equal(hits, want_hits, what=what + " primary")
            assert vis[0].tobytes() == want_vis.tobytes(), what + " shadow"
            assert_hits_equal(g.intersect(inside, CLS_TREE), want_inside, rays=inside, what=what + " inside rays")
    return want_hits, want_vis


def test_standin_dragon_depth17(monkeypatch, oracle, dragon, tmp_path):
    """configs[1]/[2] geometry: stand-in dragon n=660 (871,200 tris) + the reference scene, 960x540, primary + light0."""
    w, hs, scene = dragon
    assert hs.sizes().num_triangles == 871200 or os.environ.get("DODRT_DRAGON_OBJ")
    want_hits, want_vis = _check_scene(monkeypatch, oracle, w, hs, scene, 960, 540, 17, 10)
    # the 1920x1080 raster of configs[1], decimated: every 3rd pixel of every 3rd row as an explicit batch
    rays = oracle.primary_rays(1920, 1080).reshape(1080, 1920)[::3, ::3].ravel()
    want = oracle.intersect(scene, rays, CLS_TREE, nthreads=8)
    with hs.upload(0) as g:
        assert_hits_equal(g.intersect(rays, CLS_TREE), want, what="dragon 1080p decimated, tree only")
    if not have_ref():
        return
    # ... and against the reference's own code on the same mesh file
    files = workloads.write_mesh_files(w, str(tmp_path))
    ref = RefLib()
    ref.set_config(960, 540)
    ref.add_reference_spheres(1, 16)
    ref.add_reference_planes()
    ref.add_reference_cylinder()
    for f in files:
        ref.add_mesh(f)
    ref.build_tree()
    rr = ref.primary_rays(960, 540)
    assert_hits_equal(want_hits, ref.intersect(rr, w.classes, 8), what="oracle vs reference, dragon primary")
    t, vis = ref.trace_frame(960, 540, w.classes, LIGHT0, 8)
    assert vis.tobytes() == want_vis.tobytes()
    hit = want_hits["prim"] != 0xFFFFFFFF
    assert (t[hit].view(np.uint32) == want_hits["t"][hit].view(np.uint32)).all()


def test_sixteen_dragon_grid_depth21(monkeypatch, oracle, dragon16):
    """configs[4] geometry: 16 stand-in dragons flattened into ONE kd-tree (13.9 M tris, depth 21), 960x540."""
    w, hs, scene = dragon16
    _check_scene(monkeypatch, oracle, w, hs, scene, 960, 540, 20, 12)


def _chain_tree(depth):
    """A hand-made kd-tree whose LEFT spine is `depth` interior nodes deep (DFS pre-order, kdtree.h:16-48): node i splits
    x at depth - i, its right child is a one-lane leaf holding a triangle across the slab [depth-i, depth-i+1].  A ray
    travelling +x from x < 0 takes the near (left) child at every level and stacks the far one (kdtree.cpp:320-329): the
    short stack grows to `depth` entries -- more than a donation slot holds (kDonateMaxStack = 16), which the real
    meshes never reach (depth-21 grid: 13)."""
    n_nodes = 2 * depth + 1
    w0 = np.zeros(n_nodes, np.uint32)
    w1 = np.zeros(n_nodes, np.uint32)
    lanes = np.zeros((depth + 1, 9, 8), np.float32)
    rng = np.random.RandomState(3)

    def tri(lane, x):  # a triangle perpendicular to x covering part of the unit yz square, slot chosen at random
        j = rng.randint(0, 8)
        a = np.array([x, rng.uniform(-0.2, 0.3), rng.uniform(-0.2, 0.3)], np.float32)
        b = a + np.array([0.0, rng.uniform(0.6, 1.2), 0.0], np.float32)
        c = a + np.array([0.0, 0.0, rng.uniform(0.6, 1.2)], np.float32)
        lanes[lane, 0:3, j], lanes[lane, 3:6, j], lanes[lane, 6:9, j] = a, b, c

    for i in range(depth):
        w0[i] = 0 | ((2 * depth - i) << 2)                      # axis x, right child index
        w1[i] = np.float32(depth - i).view(np.uint32)           # split offset
    w0[depth], w1[depth] = 3 | (1 << 2), 0                       # leftmost leaf: slab [0, 1], lane 0
    tri(0, 0.5)
    for k in range(1, depth + 1):                                # right child of node depth-k: slab [k, k+1], lane k
        w0[depth + k], w1[depth + k] = 3 | (1 << 2), k
        tri(k, k + 0.5)
    nodes = w0.astype(np.uint64) | (w1.astype(np.uint64) << np.uint64(32))
    bounds = np.array([0, -0.5, -0.5, depth + 1, 1.5, 1.5], np.float32)
    return nodes, lanes.reshape(-1, 72), bounds


def test_stack_deeper_than_a_donation_slot(monkeypatch, oracle):
    from oracle_api import Scene
    from gpu_util import upload
    depth = 28
    nodes, lanes, bounds = _chain_tree(depth)
    scene = Scene(nodes, lanes, bounds)
    rng = np.random.RandomState(9)
    n = 50000
    o = np.stack([np.full(n, -1.0), rng.uniform(-0.3, 1.3, n), rng.uniform(-0.3, 1.3, n)], axis=1).astype(np.float32)
    d = np.stack([np.ones(n), rng.uniform(-0.02, 0.02, n), rng.uniform(-0.02, 0.02, n)], axis=1).astype(np.float32)
    d /= np.sqrt((d * d).sum(axis=1, dtype=np.float32))[:, None].astype(np.float32)
    rays = make_rays(o, d)
    rays["flags"][1::3] = RAY_ANY
    rays["clip"][2::5] = rng.uniform(1.0, depth, len(rays["clip"][2::5])).astype(np.float32)
    want, ctr = oracle.intersect(scene, rays, CLS_TREE, counters=True, nthreads=8)
    assert int(ctr["max_stack"].max()) >= depth - 1 and 0.2 < (want["prim"] != 0xFFFFFFFF).mean() < 1.0
    for variant, always in MODES:
        monkeypatch.setenv("DODRT_DONATE_ALWAYS", always)
        with upload(scene) as g:
            g.set_kernel_variant(variant)
            assert_hits_equal(g.intersect(rays, CLS_TREE), want, rays=rays, what=f"chain tree variant {variant} always {always}")
