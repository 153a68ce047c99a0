"""Deterministic test scenes and ray sets shared by the tests and tests/golden/make_golden.py."""
from __future__ import annotations

import os

import numpy as np

from oracle_api import RAY_DT, Scene, reference_cylinder, reference_planes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LIGHT0 = np.array([0.0, 0.0, -2.0], np.float32)  # lights[0], main.cpp:284


def lcg_uniform(seed: int, n: int) -> np.ndarray:
    """n fp32 values in [0,1): x <- 1664525 x + 1013904223 (mod 2^32), value = (x >> 8) / 2^24."""
    vals = np.empty(n, np.uint32)
    x = seed & 0xFFFFFFFF
    for i in range(n):
        x = (1664525 * x + 1013904223) & 0xFFFFFFFF
        vals[i] = x >> 8
    return vals.astype(np.float32) / np.float32(16777216.0)  # both exact in fp32


def analytic_scene_arrays(seed: int = 4, count: int = 10000):
    """BASELINE.json config 4 law: centres uniform in [-4.5,4.5]^3, radius / half-extent uniform in
    [0.03,0.12].  Returns (spheres [N,4] x y z r, boxes [N,6] min xyz max xyz)."""
    u = lcg_uniform(seed, count * 8).reshape(count, 8)
    nine, lo, span = np.float32(9.0), np.float32(4.5), np.float32(0.09)
    sc = u[:, 0:3] * nine - lo
    sr = u[:, 3] * span + np.float32(0.03)
    bc = u[:, 4:7] * nine - lo
    bh = (u[:, 7] * span + np.float32(0.03))[:, None]
    spheres = np.concatenate([sc, sr[:, None]], axis=1).astype(np.float32)
    boxes = np.concatenate([bc - bh, bc + bh], axis=1).astype(np.float32)
    return spheres, boxes


def load_teapot_arrays():
    """The flattened teapot scene exactly as the reference builds it (fixture made by make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "teapot_scene.npz"))
    lanes = z["orig_lanes"][z["prim_nums"]]  # Triangle::reorderLanesByIndices, triangle.cpp:349-367
    return dict(nodes=z["nodes"], tri_lanes=np.ascontiguousarray(lanes), bounds=z["bounds"], prim_nums=z["prim_nums"],
                spheres=z["spheres"], max_depth=int(z["max_depth"]))


def teapot_scene(full: bool = True) -> Scene:
    """Teapot kd-tree + (if full) the reference's 16 srand(1) spheres, 6 planes and the cylinder."""
    a = load_teapot_arrays()
    if not full:
        return Scene(a["nodes"], a["tri_lanes"], a["bounds"])
    return Scene(a["nodes"], a["tri_lanes"], a["bounds"], spheres=a["spheres"][:, :4], planes=reference_planes(),
                 cylinders=reference_cylinder())


def make_rays(o, d, clip=np.inf, flags=0) -> np.ndarray:
    o = np.asarray(o, np.float32).reshape(-1, 3)
    d = np.asarray(d, np.float32).reshape(-1, 3)
    n = max(len(o), len(d))
    rays = np.zeros(n, RAY_DT)
    rays["o"], rays["d"] = o, d
    rays["clip"] = clip
    rays["flags"] = flags
    return rays


def edge_rays(nodes: np.ndarray, bounds: np.ndarray) -> np.ndarray:
    """Edge cases of SURVEY.md section 4: axis-parallel rays (inv = +-inf, NaN from 0*inf in the
    slabs), origins exactly on split planes and on the bounds, origins inside the tree, clipped
    queries (clip < tmin at entry), any-hit rays."""
    w0 = (nodes & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    w1 = (nodes >> np.uint64(32)).astype(np.uint32)
    interior = (w0 & 3) != 3
    splits = w1[interior].view(np.float32)
    axes = (w0[interior] & 3).astype(np.int64)
    lo, hi = bounds[:3], bounds[3:]
    ctr = (lo + hi) * np.float32(0.5)
    rays = []
    # axis-parallel grids through the bounds, both directions, from outside and from the centre planes
    g = np.linspace(0.03, 0.97, 12, dtype=np.float32)
    for ax in range(3):
        a1, a2 = (ax + 1) % 3, (ax + 2) % 3
        for s in (-1.0, 1.0):
            for u in g:
                for v in g:
                    o = np.zeros(3, np.float32)
                    o[a1] = lo[a1] + u * (hi[a1] - lo[a1])
                    o[a2] = lo[a2] + v * (hi[a2] - lo[a2])
                    d = np.zeros(3, np.float32)
                    d[ax] = s
                    for start in (lo[ax] - 1.0 if s > 0 else hi[ax] + 1.0, ctr[ax], lo[ax] if s > 0 else hi[ax]):
                        o2 = o.copy()
                        o2[ax] = start
                        rays.append((o2, d.copy(), np.inf, 0))
    # origins exactly on split planes, with zero / negative / positive direction along the split axis
    rng = np.random.RandomState(7)
    pick = rng.choice(len(splits), size=min(200, len(splits)), replace=False)
    for k in pick:
        ax = int(axes[k])
        o = (lo + rng.rand(3).astype(np.float32) * (hi - lo)).astype(np.float32)
        o[ax] = splits[k]
        for comp in (0.0, -0.5, 0.5):
            d = rng.randn(3).astype(np.float32)
            d[ax] = comp
            d /= np.float32(np.sqrt(np.float32((d * d).sum())))
            rays.append((o.copy(), d, np.inf, 0))
    # random rays from a shell around the tree towards points inside it, some clipped short,
    # some any-hit with a finite clip
    for i in range(1500):
        p = (lo + rng.rand(3).astype(np.float32) * (hi - lo)).astype(np.float32)
        o = (ctr + rng.randn(3).astype(np.float32) * np.float32(6.0)).astype(np.float32)
        d = p - o
        dist = np.float32(np.sqrt(np.float32((d * d).sum())))
        d = (d / dist).astype(np.float32)
        mode = i % 4
        if mode == 0:
            rays.append((o, d, np.inf, 0))
        elif mode == 1:
            rays.append((o, d, np.float32(dist * rng.rand()), 0))  # clipped closest-hit
        elif mode == 2:
            rays.append((o, d, dist, 1))  # any-hit to a point inside
        else:
            rays.append((p, -d, np.inf, 0))  # origin inside the bounds
    out = np.zeros(len(rays), RAY_DT)
    for i, (o, d, c, f) in enumerate(rays):
        out[i]["o"], out[i]["d"], out[i]["clip"], out[i]["flags"] = o, d, c, f
    return out


def hit_points(rays: np.ndarray, t: np.ndarray) -> np.ndarray:
    """hitPoint = o + d*t, mul then add in fp32 (triangle.cpp:170)."""
    p = np.empty((len(rays), 3), np.float32)
    for k in range(3):
        p[:, k] = rays["o"][:, k] + (rays["d"][:, k] * t.astype(np.float32)).astype(np.float32)
    return p


# ---- meshes for the CPU arms WITHOUT the product's libraries (bench.py --impl reference) ------------------------------
def standin_dragon_py(n: int = 660):
    """The deterministic stand-in for the missing assets/dragon.obj, restated with Python's libm-backed math.sin/cos in
    the expression order of dodrt_host_standin_dragon (dod_raytracer_b200/host/dodrt_host.cpp): bit-identical positions
    (tests/test_host_vs_ref.py::test_python_standin_dragon_is_the_host_librarys), so the reference arm of bench.py
    builds the very same mesh without mapping libdodrt_host.so."""
    import math
    pi = 3.14159265358979323846
    pos = np.empty(((n + 1) * (n + 1), 3), np.float32)
    k = 0
    for j in range(n + 1):
        v = 0.02 + (pi - 0.04) * float(j) / float(n)
        sv, cv, s5v, s29v = math.sin(v), math.cos(v), math.sin(5.0 * v), math.sin(29.0 * v)
        for i in range(n + 1):
            u = 2.0 * pi * float(i) / float(n)
            r = 2.2 + 0.25 * math.sin(7.0 * u) * s5v + 0.08 * math.sin(31.0 * u + 3.0) * s29v
            pos[k, 0] = r * sv * math.cos(u)
            pos[k, 1] = r * cv
            pos[k, 2] = r * sv * math.sin(u)
            k += 1
    stride = n + 1
    j, i = np.meshgrid(np.arange(n, dtype=np.uint32), np.arange(n, dtype=np.uint32), indexing="ij")
    a = (j * stride + i).ravel()
    b, c, d = a + 1, a + stride + 1, a + stride
    idx = np.stack([a, b, c, c, d, a], axis=1).reshape(-1, 3).astype(np.uint32)  # quad a b / d c -> (a,b,c), (c,d,a)
    return pos, idx


def write_dodm_py(path: str, positions: np.ndarray, indices: np.ndarray):
    """'DODM' + u32 vertices + u32 triangles + positions + indices: the binary mesh both loaders read."""
    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    indices = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
    with open(path, "wb") as f:
        f.write(b"DODM")
        f.write(np.array([len(positions), len(indices)], np.uint32).tobytes())
        f.write(positions.tobytes())
        f.write(indices.tobytes())


def workload_mesh_files_py(w, directory: str):
    """dod_raytracer_b200.workloads.write_mesh_files without the host library (same bytes)."""
    from dod_raytracer_b200 import workloads  # plain-Python constants only; nothing is dlopen'ed by this import
    if w.mesh == "teapot":
        return [workloads.TEAPOT_FIXTURE]
    if w.mesh not in ("dragon", "dragon16"):
        return []
    real = os.environ.get("DODRT_DRAGON_OBJ")
    if real:
        return [real]
    pos, idx = standin_dragon_py(w.dragon_n)
    out = []
    for k, (scale, tr) in enumerate(workloads._dragon_instances(w)):
        p = (pos * np.float32(scale) + np.asarray(tr, np.float32)).astype(np.float32)
        path = os.path.join(directory, f"{w.name}_{k}.dodm")
        write_dodm_py(path, p, idx)
        out.append(path)
    return out


def tie_scene_triangles(kind: str, n: int = 6000, seed: int = 5):
    """Triangles for the tie-rule tests: vertices on a coarse lattice -- shared edges and vertices,
    axis-parallel faces -- neighbours in space consecutive in the file; kind "duplicates": every triangle twice, the
    copies adjacent (same lane), and every 16th once more at the end of the file (another lane, other leaves).
    Returns (positions [3n, 3] float32, indices [n, 3] uint32)."""
    rng = np.random.default_rng(seed)
    c = rng.integers(-4, 5, (n, 1, 3)).astype(np.float64) * 0.5
    c = c[np.lexsort((c[:, 0, 2], c[:, 0, 1], c[:, 0, 0]))]
    tri = c + rng.integers(0, 2, (n, 3, 3)) * 0.5
    if kind == "duplicates":
        tri = np.repeat(tri[: n // 2], 2, axis=0)
        tri = np.concatenate([tri, tri[::16]])
    pos = tri.reshape(-1, 3).astype(np.float32)
    return pos, np.arange(len(pos), dtype=np.uint32).reshape(-1, 3)


def tie_scene_rays(m: int = 40000, seed: int = 6):
    """Rays along lattice lines, through lattice points (slightly perturbed) and at random; every 4th is any-hit."""
    rng = np.random.default_rng(seed)
    o = np.empty((m, 3), np.float32)
    d = np.empty((m, 3), np.float32)
    third = m // 3
    o[:third] = rng.integers(-5, 6, (third, 3)) * 0.5
    o[:third, 2] = -4.0
    d[:third] = (0.0, 0.0, 1.0)
    o[third:2 * third] = rng.integers(-10, 11, (third, 3)) * 0.25
    tgt = rng.integers(-5, 6, (third, 3)) * 0.5
    d[third:2 * third] = tgt - o[third:2 * third] + np.float32(1e-3) * rng.standard_normal((third, 3))
    o[2 * third:] = rng.uniform(-4, 4, (m - 2 * third, 3))
    d[2 * third:] = rng.standard_normal((m - 2 * third, 3))
    norm = np.sqrt((d.astype(np.float64) ** 2).sum(axis=1))
    norm[norm == 0] = 1.0
    rays = make_rays(o, (d / norm[:, None]).astype(np.float32))
    rays["flags"][1::4] = 1  # DODRT_RAY_ANY
    return rays


def irregular_rays(kind: str, n: int = 30000, seed: int = 3):
    """Rays the reference's arithmetic was not written for but answers deterministically: "unnormalised" = directions
    of length 1e-3 ... 1e3; "axis-parallel" = one or two zero components (infinite slab inverses, 0 * inf = NaN in
    box.cpp:38-47); "nan-inf-zero" = NaN / inf / all-zero directions and NaN origins.  A third carry a finite clip,
    every 4th is any-hit."""
    rng = np.random.default_rng(seed)
    o = rng.uniform(-4.5, 4.5, (n, 3)).astype(np.float32)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    if kind == "unnormalised":
        d = (d * np.float32(10.0) ** rng.uniform(-3, 3, (n, 1)).astype(np.float32)).astype(np.float32)
    elif kind == "axis-parallel":
        d[np.arange(n), rng.integers(0, 3, n)] = 0.0
        d[::2, 1] = 0.0
    else:
        d[::5, 0] = np.nan
        d[1::5, 1] = np.inf
        d[2::5] = 0.0
        o[3::5, 2] = np.nan
    rays = make_rays(o, d)
    rays["flags"][1::4] = 1  # DODRT_RAY_ANY
    rays["clip"][::3] = rng.uniform(0, 8, len(rays["clip"][::3])).astype(np.float32)
    return rays
