#!/usr/bin/env python
"""Regenerates the fixtures in tests/golden/ from the reference itself.

Run in the build container only (needs /root/reference and oracle/_ref/libdodrt_ref.so):
    python tests/golden/make_golden.py
Everything written here comes out of the reference's own translation units (oracle/ref_harness.cpp
around the unmodified /root/reference/src/**/*.cpp), never out of the restatement or the CUDA path.
The one exception is boxes_*: renderable boxes are an extension with no reference counterpart, so
their fixture is produced by the oracle restatement and only guards against regressions.
"""
import ctypes
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_api import (CLS_BOX, CLS_CYLINDER, CLS_PLANE, CLS_SPHERE, CLS_TREE, MISS, Oracle, RefLib, Scene,  # noqa: E402
                        ensure_oracle_built)
from scenes import LIGHT0, analytic_scene_arrays, edge_rays, hit_points  # noqa: E402

REF_TEAPOT = "/root/reference/assets/teapot.obj"
W, H = 1920, 1080
STEP = 12  # decimation of the committed per-ray vectors: every 12th column/row of the 1080p frame
ALL = CLS_SPHERE | CLS_PLANE | CLS_CYLINDER | CLS_TREE


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def obj_to_dodm(obj_path: str, out_path: str):
    """Re-encode the OBJ as the binary mesh the stand-in importer also reads; floats parsed with libc
    strtof exactly like oracle/shim/assimp_shim.cpp does."""
    libc = ctypes.CDLL("libc.so.6")
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    pos, idx = [], []
    for line in open(obj_path, "rb"):
        tok = line.split()
        if not tok:
            continue
        if tok[0] == b"v":
            pos.append([libc.strtof(t, None) for t in tok[1:4]])
        elif tok[0] == b"f":
            c = [int(t.split(b"/")[0]) for t in tok[1:]]
            c = [v - 1 if v > 0 else len(pos) + v for v in c]
            for k in range(1, len(c) - 1):
                idx.append([c[0], c[k], c[k + 1]])
    pos = np.array(pos, np.float32)
    idx = np.array(idx, np.uint32)
    with open(out_path, "wb") as f:
        f.write(b"DODM")
        f.write(np.array([len(pos), len(idx)], np.uint32).tobytes())
        f.write(pos.tobytes())
        f.write(idx.tobytes())
    return pos, idx


def build_ref(mesh_path):
    ref = RefLib()
    ref.set_config(W, H)
    ref.add_reference_spheres(1, 16)  # srand(1) instead of srand(time(NULL)), main.cpp:351
    ref.add_reference_planes()
    ref.add_reference_cylinder()
    assert ref.add_mesh(mesh_path) == 6320
    ref.build_tree()
    return ref


def main():
    ensure_oracle_built()
    dodm = os.path.join(HERE, "teapot.dodm")
    obj_to_dodm(REF_TEAPOT, dodm)
    ref_obj = build_ref(REF_TEAPOT)
    ref = build_ref(dodm)
    nodes, lanes, prim, bounds = ref.export_tree()
    n2, l2, p2, b2 = ref_obj.export_tree()
    assert nodes.tobytes() == n2.tobytes() and lanes.tobytes() == l2.tobytes() and prim.tobytes() == p2.tobytes()
    assert bounds.tobytes() == b2.tobytes()
    sizes = ref.tree_sizes()
    print("teapot tree", sizes)
    # pins from SURVEY.md section 4 / appendix D
    assert sizes == dict(nodes=649, lanes=3021, prim_nums=3021, max_depth=10, triangles=6320)

    num_orig = (sizes["triangles"] + 7) // 8
    orig = np.zeros((num_orig, 72), np.float32)
    seen = np.zeros(num_orig, bool)
    for k, p in enumerate(prim):
        if not seen[p]:
            orig[p] = lanes[k]
            seen[p] = True
    assert seen.all() and (orig[prim] == lanes).all()
    spheres = ref.export_spheres()
    normals = ref.export_normals()
    np.savez_compressed(os.path.join(HERE, "teapot_scene.npz"), nodes=nodes, orig_lanes=orig, prim_nums=prim,
                        bounds=bounds, spheres=spheres, max_depth=np.uint32(sizes["max_depth"]),
                        lanes_sha=np.array(sha(lanes)), normals_sha=np.array(sha(normals)))

    # ---- primary + shadow rays, full 1080p frame through the reference ------------------------------
    rays = ref.primary_rays(W, H)
    sel = (np.arange(0, H, STEP)[:, None] * W + np.arange(0, W, STEP)[None, :]).ravel()
    out = dict(step=np.uint32(STEP), width=np.uint32(W), height=np.uint32(H), light=LIGHT0)
    for name, cls in (("tree", CLS_TREE), ("all", ALL), ("sphere_tree", CLS_SPHERE | CLS_TREE)):
        hits = ref.intersect(rays, cls, 8)
        hit_mask = hits["prim"] != MISS
        pts = hit_points(rays, hits["t"])
        sh = ref.shadow_rays(pts, LIGHT0)
        occ = ref.intersect(sh, cls, 8)
        vis = ((occ["prim"] == MISS) & hit_mask).astype(np.uint8)
        print(name, "hits", int(hit_mask.sum()), "visible", int(vis.sum()))
        out[f"{name}_hits"] = hits[sel]
        out[f"{name}_vis"] = vis[sel]
        out[f"{name}_num_hits"] = np.uint64(hit_mask.sum())
        out[f"{name}_num_visible"] = np.uint64(vis.sum())
        out[f"{name}_hits_sha"] = np.array(sha(hits))
        out[f"{name}_vis_sha"] = np.array(sha(vis))
    assert int(out["tree_num_hits"]) == 170522
    out["rays_sha"] = np.array(sha(rays))
    out["rays_sel"] = rays[sel]
    np.savez_compressed(os.path.join(HERE, "teapot_frame.npz"), **out)

    # ---- edge cases -----------------------------------------------------------------------------------
    er = edge_rays(nodes, bounds)
    eh_tree = ref.intersect(er, CLS_TREE, 1)
    eh_all = ref.intersect(er, ALL, 1)
    print("edge rays", len(er), "tree hits", int((eh_tree["prim"] != MISS).sum()))
    np.savez_compressed(os.path.join(HERE, "teapot_edge.npz"), rays=er, tree=eh_tree, all=eh_all)

    # ---- 10k analytic spheres (config 4), brute force through Sphere::intersect ----------------------
    spheres10k, boxes10k = analytic_scene_arrays(4, 10000)
    ref2 = RefLib()
    ref2.add_spheres(spheres10k)
    sw, shh = 192, 108
    ref2.set_config(sw, shh)
    srays = ref2.primary_rays(sw, shh)
    shits = ref2.intersect(srays, CLS_SPHERE, 8)
    spts = hit_points(srays, shits["t"])
    ssh = ref2.shadow_rays(spts, LIGHT0)
    socc = ref2.intersect(ssh, CLS_SPHERE, 8)
    svis = ((socc["prim"] == MISS) & (shits["prim"] != MISS)).astype(np.uint8)
    print("10k spheres hits", int((shits["prim"] != MISS).sum()), "visible", int(svis.sum()))
    # boxes: extension, oracle-generated (regression guard only)
    orc = Oracle()
    bscene = Scene(spheres=spheres10k, boxes=boxes10k)
    bhits = orc.intersect(bscene, srays, CLS_SPHERE | CLS_BOX)
    np.savez_compressed(os.path.join(HERE, "analytic10k.npz"), width=np.uint32(sw), height=np.uint32(shh),
                        spheres_sha=np.array(sha(spheres10k)), boxes_sha=np.array(sha(boxes10k)),
                        sphere_hits=shits, sphere_vis=svis, sphere_box_hits_oracle=bhits)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))




def render_fixture():
    """The reference's own rayTrace (main.cpp:273-347: 9 lights, 10 bounces, all shape classes) on the srand(1)
    teapot scene, one band starting at row 0 (canonical raster tables), 240x135."""
    ref = build_ref(os.path.join(HERE, "teapot.dodm"))
    img = ref.render(240, 135, nthreads=1)
    np.savez_compressed(os.path.join(HERE, "teapot_render_240x135.npz"), rgb=img)
    print("render fixture", img.shape, float(img.mean()))


if __name__ == "__main__":
    if not os.environ.get("DODRT_GOLDEN_RENDER_ONLY"):
        main()
    render_fixture()
