"""CPU: the oracle restatement (oracle/dodrt_oracle.c) against the committed fixtures that were generated
from the reference's own translation units (tests/golden/make_golden.py).  Bit-exact everywhere."""
import hashlib

import numpy as np

from oracle_api import (CLS_BOX, CLS_CYLINDER, CLS_PLANE, CLS_SPHERE, CLS_TREE, MISS, RAY_ANY, Scene)
from scenes import GOLDEN, LIGHT0, analytic_scene_arrays, edge_rays, load_teapot_arrays, teapot_scene

ALL = CLS_SPHERE | CLS_PLANE | CLS_CYLINDER | CLS_TREE
CLASSES = {"tree": CLS_TREE, "all": ALL, "sphere_tree": CLS_SPHERE | CLS_TREE}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_teapot_scene_fixture_is_the_reference_tree():
    a = load_teapot_arrays()
    z = np.load(f"{GOLDEN}/teapot_scene.npz")
    # structural pins measured on the reference
    assert len(a["nodes"]) == 649 and len(a["tri_lanes"]) == 3021 and a["max_depth"] == 10
    w0 = (a["nodes"] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    leaves = (w0 & 3) == 3
    assert leaves.sum() == 325 and ((w0[leaves] >> 2) == 0).sum() == 32 and (w0[leaves] >> 2).max() == 38
    assert np.allclose(a["bounds"], [-3, 0, -2, 3.434, 3.15, 2])
    assert sha(a["tri_lanes"]) == str(z["lanes_sha"])


def test_ray_tables_and_primary_rays(oracle):
    z = np.load(f"{GOLDEN}/teapot_frame.npz")
    w, h, step = int(z["width"]), int(z["height"]), int(z["step"])
    rays = oracle.primary_rays(w, h)
    assert sha(rays) == str(z["rays_sha"])
    sel = (np.arange(0, h, step)[:, None] * w + np.arange(0, w, step)[None, :]).ravel()
    assert rays[sel].tobytes() == z["rays_sel"].tobytes()
    xs, ys = oracle.ray_tables(w, h)
    assert xs[0] == np.float32(-(np.float32(w) / np.float32(h))) and ys[0] == np.float32(1.0)


def test_full_frame_hits_and_shadows_match_reference(oracle):
    """1080p teapot frame: every hit record (t bits, id, u, v) and every shadow bit, via checksums of the
    full arrays plus the committed decimated vectors."""
    z = np.load(f"{GOLDEN}/teapot_frame.npz")
    w, h, step = int(z["width"]), int(z["height"]), int(z["step"])
    sel = (np.arange(0, h, step)[:, None] * w + np.arange(0, w, step)[None, :]).ravel()
    scene = teapot_scene(full=True)
    for name, cls in CLASSES.items():
        hits = oracle.trace_primary(scene, w, h, cls, nthreads=8)
        assert int((hits["prim"] != MISS).sum()) == int(z[f"{name}_num_hits"])
        assert hits[sel].tobytes() == z[f"{name}_hits"].tobytes(), name
        assert sha(hits) == str(z[f"{name}_hits_sha"]), name
        vis = oracle.trace_shadow(scene, w, h, cls, hits, LIGHT0, nthreads=8)
        assert int(vis.sum()) == int(z[f"{name}_num_visible"])
        assert vis[sel].tobytes() == z[f"{name}_vis"].tobytes(), name
        assert sha(vis) == str(z[f"{name}_vis_sha"]), name


def test_explicit_ray_batch_equals_fused_primary(oracle):
    scene = teapot_scene(full=True)
    w, h = 160, 90
    rays = oracle.primary_rays(w, h)
    a = oracle.intersect(scene, rays, ALL)
    b = oracle.trace_primary(scene, w, h, ALL)
    assert a.tobytes() == b.tobytes()


def test_edge_case_rays(oracle):
    """axis-parallel rays (inf/NaN slabs), origins on split planes, clipped and any-hit queries"""
    z = np.load(f"{GOLDEN}/teapot_edge.npz")
    a = load_teapot_arrays()
    rays = z["rays"]
    assert edge_rays(a["nodes"], a["bounds"]).tobytes() == rays.tobytes()
    scene = teapot_scene(full=True)
    anyray = (rays["flags"] & RAY_ANY) != 0
    for name, cls in (("tree", CLS_TREE), ("all", ALL)):
        got, want = oracle.intersect(scene, rays, cls), z[name]
        assert (got["prim"] == want["prim"]).all(), name
        closest = ~anyray
        assert got[closest].tobytes() == want[closest].tobytes(), name
    # the set exercises what it claims to
    d = rays["d"]
    assert ((d == 0).sum(axis=1) == 2).sum() > 1000 and anyray.sum() > 100
    assert np.isfinite(rays["clip"]).sum() > 300


def test_ten_thousand_spheres(oracle):
    z = np.load(f"{GOLDEN}/analytic10k.npz")
    spheres, boxes = analytic_scene_arrays(4, 10000)
    assert sha(spheres) == str(z["spheres_sha"]) and sha(boxes) == str(z["boxes_sha"])
    w, h = int(z["width"]), int(z["height"])
    scene = Scene(spheres=spheres, boxes=boxes)
    hits = oracle.trace_primary(scene, w, h, CLS_SPHERE, nthreads=8)
    assert hits.tobytes() == z["sphere_hits"].tobytes()
    vis = oracle.trace_shadow(scene, w, h, CLS_SPHERE, hits, LIGHT0, nthreads=8)
    assert vis.tobytes() == z["sphere_vis"].tobytes()
    # extension class (regression guard; no reference counterpart)
    bh = oracle.trace_primary(scene, w, h, CLS_SPHERE | CLS_BOX, nthreads=8)
    assert bh.tobytes() == z["sphere_box_hits_oracle"].tobytes()
    kinds = bh["prim"][bh["prim"] != MISS] >> 29
    assert (kinds == 4).sum() > 1000 and (kinds == 1).sum() > 1000


def test_box_extension_semantics(oracle):
    """unit box at the origin: entry distance, origin-inside rejection, lower id wins ties"""
    boxes = np.array([[-1, -1, -1, 1, 1, 1], [-1, -1, -1, 1, 1, 1], [2, -1, -1, 3, 1, 1]], np.float32)
    scene = Scene(boxes=boxes)
    from scenes import make_rays
    rays = make_rays([[-5, 0, 0], [0, 0, 0], [5, 0.5, 0.5], [-5, 0, 0]], [[1, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0]])
    hits = oracle.intersect(scene, rays, CLS_BOX)
    assert hits["t"][0] == 4.0 and hits["prim"][0] == (4 << 29) | 0  # tie between box 0 and 1 -> lower id
    assert hits["prim"][1] == (4 << 29) | 2 and hits["t"][1] == 2.0  # inside boxes 0/1: rejected, box 2 ahead
    assert hits["prim"][2] == (4 << 29) | 2 and hits["t"][2] == 2.0
    assert hits["prim"][3] == MISS


def test_counters_give_algorithmic_bytes(oracle):
    """teapot tree-only primary rays at 1080p = 3.70 nodes + 3.18 lanes = 945 B/ray"""
    scene = teapot_scene(full=False)
    hits, ctr = oracle.trace_primary(scene, 1920, 1080, CLS_TREE, counters=True, nthreads=8)
    nodes, lanes = ctr["nodes"].mean(), ctr["lanes"].mean()
    assert abs(nodes - 3.70) < 0.01 and abs(lanes - 3.18) < 0.01
    assert abs(8 * nodes + 288 * lanes - 944.9) < 1.0
    assert ctr["max_stack"].max() == 6
