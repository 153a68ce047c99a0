#!/usr/bin/env python
"""f-4 scene ingest timing on the GPU box: mesh registration, kd build (1 thread vs all cores, same
bits), host-side lane re-order + plain upload vs creation-order lanes + GPU-side re-order (dodrt_scene_set_kdtree_indexed).
    python tests/tools/ingest_bench.py [workload]      (default dragon16_8k)"""
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

CHILD = r"""
import hashlib, json, sys, time
import numpy as np
sys.path.insert(0, %r)
from dod_raytracer_b200 import capi, host, workloads
w = workloads.WORKLOADS[%r]
keep = %r
capi.Scene(0).close()  # CUDA context creation is not part of the ingest
t0 = time.perf_counter()
hs = host.HostScene()
if w.reference_scene:
    hs.add_reference_scene(1, 16)
pos, idx = host.standin_dragon(w.dragon_n)
for scale, tr in workloads._dragon_instances(w):
    hs.add_mesh((pos * np.float32(scale) + np.asarray(tr, np.float32)).astype(np.float32), idx)
t1 = time.perf_counter()
hs.build_tree(keep_creation_order=keep)
t2 = time.perf_counter()
g = hs.upload(0)
t3 = time.perf_counter()
z = hs.sizes()
a = hs.arrays(raw=True)
h = hashlib.sha256()
h.update(a["nodes"].tobytes()); h.update(a["prim_nums"].tobytes())
xs, ys = host.ray_tables(640, 360)
hits, vis = g.trace_frame(capi.Frame.make(640, 360, classes=w.classes), xs, ys, np.array(w.lights[0], np.float32)[None, :])
print(json.dumps(dict(add_mesh_s=t1 - t0, build_tree_s=t2 - t1, upload_s=t3 - t2, nodes=z.num_nodes, lanes=z.num_lanes,
                      orig_lanes=z.num_orig_lanes, tree_sha=h.hexdigest()[:16],
                      frame_sha=hashlib.sha256(hits.tobytes() + vis.tobytes()).hexdigest()[:16])))
"""


def run(workload, threads, keep):
    env = dict(os.environ, DODRT_HOST_THREADS=str(threads), DODRT_HOST_VERBOSE="1")
    p = subprocess.run([sys.executable, "-c", CHILD % (ROOT, workload, keep)], capture_output=True, text=True, env=env)
    if p.returncode:
        raise SystemExit(p.stderr[-2000:])
    out = json.loads(p.stdout.strip().splitlines()[-1])
    out["threads"], out["gpu_reorder"] = threads, keep
    out["host_log"] = [l for l in p.stderr.splitlines() if "kd build" in l]
    return out


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "dragon16_8k"
    cores = os.cpu_count() or 1
    rows = [run(workload, 1, False), run(workload, cores, False), run(workload, cores, True)]
    assert len({r["tree_sha"] for r in rows}) == 1, "kd-tree differs between thread counts"
    assert len({r["frame_sha"] for r in rows}) == 1, "traced frame differs between ingest paths"
    for r in rows:
        print(json.dumps(r))
    print(f"{workload}: identical tree and identical 640x360 frame on all three paths ({cores} host cores)")


if __name__ == "__main__":
    main()
