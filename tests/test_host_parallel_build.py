"""f-4: the task-parallel kd build and the deferred lane re-order must reproduce the sequential
reference algorithm bit for bit.  tests/test_host_vs_ref.py pins the builder against the reference's own
KDTree::buildTree output; here every thread count and the creation-order mode are compared with the 1-thread build
(the 1-thread path never forks, i.e. it IS the recursion of kdtree.cpp:95-250 in the reference's order)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from dod_raytracer_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, {root!r})
from dod_raytracer_b200 import host
pos, idx = host.standin_dragon({n})
hs = host.HostScene()
hs.add_reference_scene(1, 16)
for k in range({copies}):
    hs.add_mesh((pos * np.float32(0.5) + np.float32(k * 0.7)).astype(np.float32), idx)
hs.build_tree(keep_creation_order={keep})
a = hs.arrays(normals=True)
z = hs.sizes()
h = hashlib.sha256()
for key in ("nodes", "prim_nums", "bounds", "tri_lanes", "tri_normals", "tri_attributes"):
    h.update(np.ascontiguousarray(a[key]).tobytes())
print(z.num_nodes, z.num_lanes, z.num_orig_lanes, z.max_depth, h.hexdigest())
"""


def _build(threads: int, keep: bool, n: int = 160, copies: int = 3) -> str:
    env = dict(os.environ, DODRT_HOST_THREADS=str(threads))
    code = _CHILD.format(root=ROOT, n=n, copies=copies, keep=keep)
    return subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, text=True, env=env).stdout.strip()


def test_parallel_build_is_bit_identical_to_sequential():
    want = _build(1, False)
    assert int(want.split()[0]) > 1000, want  # a real tree, deep enough to fork (>= 4096 lanes per task)
    for threads in (2, 3, 8):
        assert _build(threads, False) == want, f"{threads} threads"


def test_creation_order_mode_defers_only_the_gather():
    """DODRT_HOST_BUILD_KEEP_CREATION_ORDER: same nodes / primNums; lanes[primNums] == the re-ordered lanes."""
    assert _build(8, True) == _build(1, False)


def test_sizes_in_creation_order_mode():
    pos, idx = host.standin_dragon(40)
    hs = host.HostScene()
    hs.add_mesh(pos, idx)
    hs.build_tree(keep_creation_order=True)
    z = hs.sizes()
    a = hs.arrays(raw=True)
    assert a["tri_lanes"].shape[0] == z.num_orig_lanes == (2 * 40 * 40 + 7) // 8
    assert a["prim_nums"].shape[0] == z.num_lanes >= z.num_orig_lanes
    assert a["prim_nums"].max() < z.num_orig_lanes
    assert hs.arrays()["tri_lanes"].shape[0] == z.num_lanes  # default view: gathered through prim_nums
    with pytest.raises(RuntimeError):
        hs.build_tree()


_OBJ_CHILD = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, {root!r})
from dod_raytracer_b200 import host
hs = host.HostScene()
hs.add_mesh_file({path!r})
a = hs.arrays(normals=True)
print(hs.sizes().num_triangles, hashlib.sha256(a["tri_lanes"].tobytes() + a["tri_normals"].tobytes()).hexdigest())
"""


def test_chunked_obj_parse_equals_sequential(tmp_path):
    """the OBJ text is parsed in per-thread chunks cut at line ends; negative (relative) indices and polygons must
    resolve exactly as in one sequential pass"""
    pos, idx = host.standin_dragon(180)
    path = tmp_path / "big.obj"
    rng = np.random.default_rng(7)
    with open(path, "w") as f:
        # vertices and faces interleaved in blocks, so relative indices cross chunk borders
        done_v = 0
        order = np.argsort(idx.max(axis=1), kind="stable")
        faces = idx[order]
        need = faces.max(axis=1)
        k = 0
        for block in range(0, len(pos), 997):
            hi = min(len(pos), block + 997)
            for p in pos[block:hi]:
                f.write(f"v {p[0]:.9g} {p[1]:.9g} {p[2]:.9g}\n")
            done_v = hi
            while k < len(faces) and need[k] < done_v:
                a, b, c = (int(x) for x in faces[k])
                if rng.random() < 0.5:  # relative form: -1 = last vertex seen
                    f.write(f"f {a - done_v} {b - done_v}/1/1 {c - done_v}//3\n")
                else:
                    f.write(f"f {a + 1} {b + 1} {c + 1}\n")
                k += 1
            f.write("# comment\nvn 0 0 1\n")
        assert k == len(faces)
    assert os.path.getsize(path) > 2 << 20  # at least two 1-MiB chunks
    out = []
    for threads in (1, 8):
        env = dict(os.environ, DODRT_HOST_THREADS=str(threads))
        code = _OBJ_CHILD.format(root=ROOT, path=str(path))
        out.append(subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, text=True, env=env).stdout)
    assert out[0] == out[1]
    assert int(out[0].split()[0]) == len(idx)
