"""CPU: the product's host side (libdodrt_host.so: lanes, kd-tree builder, scene generators, ray tables)
against the reference's own output -- every array bit-identical."""
import ctypes
import os
import tempfile

import numpy as np
import pytest

from dod_raytracer_b200 import host
from oracle_api import RefLib, have_ref, reference_cylinder, reference_planes, sphere_lanes
from scenes import GOLDEN, analytic_scene_arrays, load_teapot_arrays

needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def test_teapot_tree_matches_reference_fixture():
    """fixture = the reference's KDTree::buildTree output for teapot (tests/golden/make_golden.py)"""
    want = load_teapot_arrays()
    hs = host.HostScene()
    hs.add_mesh_file(os.path.join(GOLDEN, "teapot.dodm"))
    hs.add_reference_scene(1, 16)
    hs.build_tree()
    got = hs.arrays()
    assert got["max_depth"] == want["max_depth"] and got["num_triangles"] == 6320
    assert got["nodes"].tobytes() == want["nodes"].tobytes()
    assert got["prim_nums"].tobytes() == want["prim_nums"].tobytes()
    assert got["tri_lanes"].tobytes() == want["tri_lanes"].tobytes()
    assert got["bounds"].tobytes() == want["bounds"].tobytes()
    # srand(1) spheres (positions, radius^2) as the reference created them; planes; cylinder
    assert got["sphere_lanes"].tobytes() == sphere_lanes(want["spheres"][:, :4]).tobytes()
    assert got["sphere_colors"].tobytes() == np.ascontiguousarray(want["spheres"][:, 4:7]).tobytes()
    from oracle_api import pack_lanes
    assert got["plane_lanes"].tobytes() == pack_lanes(reference_planes()).tobytes()
    assert got["cylinders"].tobytes() == reference_cylinder().tobytes()


def test_rand_restatement_matches_libc():
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 4, 12345):
        hs = host.HostScene()
        hs.add_reference_scene(seed, 40)
        a = hs.arrays()
        libc.srand(seed)
        cols, pos = [], []
        rmax = np.float32(2147483647)
        for _ in range(40):
            cols.append([np.float32(libc.rand()) / rmax for _ in range(3)])
            pos.append([np.float32(libc.rand()) / rmax * np.float32(10.0) - np.float32(5.0) for _ in range(3)])
        assert a["sphere_colors"].tobytes() == np.array(cols, np.float32).tobytes()
        lanes = a["sphere_lanes"]
        idx = np.arange(40)
        assert lanes[idx // 8, :3, idx % 8].tobytes() == np.array(pos, np.float32).tobytes()


def test_ray_tables_match_oracle(oracle):
    for w, h in ((1920, 1080), (3840, 2160), (7680, 4320), (33, 9)):
        xs, ys = host.ray_tables(w, h)
        oxs, oys = oracle.ray_tables(w, h)
        assert xs.tobytes() == oxs.tobytes() and ys.tobytes() == oys.tobytes()


def test_analytic_scene_matches_test_generator():
    spheres, boxes = analytic_scene_arrays(4, 1000)
    hs = host.HostScene()
    hs.add_analytic_scene(4, 1000)
    a = hs.arrays()
    from oracle_api import pack_lanes
    assert a["sphere_lanes"].tobytes() == sphere_lanes(spheres).tobytes()
    assert a["box_lanes"].tobytes() == pack_lanes(boxes).tobytes()


def test_config_ini(tmp_path):
    p = tmp_path / "config.ini"
    p.write_text("Width: 1920\nHeight: 1080  \nMaxPrims : 4\nEpsilon: 0.001\nWidth: 7\n")
    cfg = host.load_config(str(p))
    assert (cfg.width, cfg.height, cfg.max_prims) == (1920, 1080, 4) and abs(cfg.epsilon - 1e-3) < 1e-9
    assert (cfg.intersect_cost, cfg.traversal_cost) == (80, 80)  # config.h:11-12 defaults
    with pytest.raises(RuntimeError):
        host.load_config(str(tmp_path / "missing.ini"))


def _compare_with_ref(mesh_path, scale=None, translate=None, normals=True):
    ref = RefLib()
    assert scale is None and translate is None
    ntri = ref.add_mesh(mesh_path)
    ref.build_tree()
    rn, rl, rp, rb = ref.export_tree()
    hs = host.HostScene()
    hs.add_mesh_file(mesh_path)
    hs.build_tree()
    a = hs.arrays(normals=normals)
    assert a["num_triangles"] == ntri
    assert a["max_depth"] == ref.tree_sizes()["max_depth"]
    assert a["nodes"].tobytes() == rn.tobytes(), "kd nodes differ"
    assert a["prim_nums"].tobytes() == rp.tobytes()
    assert a["tri_lanes"].tobytes() == rl.tobytes()
    assert a["bounds"].tobytes() == rb.tobytes()
    if normals:
        assert a["tri_normals"].tobytes() == ref.export_normals().tobytes(), "smooth normals differ"
    return a


@needs_ref
def test_teapot_live_including_normals():
    _compare_with_ref(os.path.join(GOLDEN, "teapot.dodm"))
    if os.path.exists("/root/reference/assets/teapot.obj"):  # the OBJ text path of both loaders
        _compare_with_ref("/root/reference/assets/teapot.obj")


@needs_ref
@pytest.mark.parametrize("n", [24, 150])
def test_standin_dragon_tree_matches_reference_builder(n, tmp_path):
    pos, idx = host.standin_dragon(n)
    assert len(idx) == 2 * n * n
    path = str(tmp_path / f"dragon{n}.dodm")
    host.write_dodm(path, pos * np.float32(0.68), idx)
    a = _compare_with_ref(path)
    assert a["num_triangles"] == 2 * n * n


def _soup(seed, n, kind):
    rng = np.random.default_rng(seed)
    if kind == "small":    # small triangles, neighbours in space consecutive in the file (a lane is 8 consecutive
        c = rng.uniform(-2, 2, (n, 1, 3))  # triangles, triangle.cpp:238-262): the SAH splits many times
        cell = np.floor((c[:, 0] + 2) * 4).astype(np.int64)
        c = c[np.lexsort((cell[:, 2], cell[:, 1], cell[:, 0]))]
        tri = c + rng.normal(0, 0.02, (n, 3, 3))
    elif kind == "big":    # heavily overlapping triangles: bad-refine and cost exits
        tri = rng.uniform(-2, 2, (n, 1, 3)) + rng.normal(0, 0.15, (n, 3, 3))
    elif kind == "grid":   # vertices on a coarse lattice: equal offsets everywhere (ties in the sweep and in the sorts)
        tri = rng.integers(-4, 5, (n, 1, 3)) * 0.5 + rng.integers(0, 2, (n, 3, 3)) * 0.5
    else:                  # "dup": duplicated and degenerate triangles
        base = rng.uniform(-1, 1, (max(n // 4, 1), 3, 3))
        tri = base[rng.integers(0, len(base), n)]
        tri[::7, 1] = tri[::7, 0]
    return tri.reshape(-1, 3).astype(np.float32), np.arange(3 * n, dtype=np.uint32).reshape(n, 3)


@needs_ref
@pytest.mark.parametrize("kind,n", [("small", 1), ("small", 8), ("small", 9), ("small", 20000), ("big", 5000), ("grid", 65),
                                    ("grid", 5000), ("dup", 500), ("dup", 5000)])
def test_triangle_soups_match_reference_builder(kind, n, tmp_path):
    """Random soups through both builders (the reference's KDTree::buildTree compiled from its own sources, and the host
    library): nodes, primNums, re-ordered lanes and bounds bit for bit -- also with ties in every sort and degenerate
    triangles.  (Scenes that lie in one axis-aligned plane are left out: the reference's own builder asserts on them,
    kdtree.cpp:245, through the getMaxElementIndex quirk of SURVEY appendix B.)"""
    pos, idx = _soup(1000 + n, n, kind)
    path = str(tmp_path / f"{kind}{n}.dodm")
    host.write_dodm(path, pos, idx)
    a = _compare_with_ref(path, normals=False)
    assert a["num_triangles"] == n
    if kind == "small" and n == 20000:
        assert len(a["nodes"]) > 500, len(a["nodes"])


@needs_ref
def test_obj_text_loader_with_polygons_and_slashes(tmp_path):
    p = tmp_path / "quad.obj"
    p.write_text("# comment\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0.5\nv 0.5 0.5 1\nvn 0 0 1\n"
                 "f 1/1/1 2/2/1 3/3/1 4/4/1\nf -1 1 2\nf 3//1 4//1 5//1\n")
    a = _compare_with_ref(str(p))
    assert a["num_triangles"] == 4


def test_python_standin_dragon_is_the_host_librarys(tmp_path):
    """bench.py's reference arm generates its meshes without the product's libraries (scenes.standin_dragon_py):
    the files must be byte-identical to the ones the GPU arm's host library writes."""
    import filecmp
    from dod_raytracer_b200 import workloads
    from scenes import standin_dragon_py, workload_mesh_files_py
    for n in (24, 150, 660):
        pos, idx = host.standin_dragon(n)
        ppos, pidx = standin_dragon_py(n)
        assert pos.tobytes() == ppos.tobytes() and idx.tobytes() == pidx.tobytes(), n
    w = workloads.WORKLOADS["dragon4k"]
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    fa, fb = workloads.write_mesh_files(w, str(tmp_path / "a")), workload_mesh_files_py(w, str(tmp_path / "b"))
    assert len(fa) == len(fb) == 1 and filecmp.cmp(fa[0], fb[0], shallow=False)
