"""Helpers for the -m gpu tests: upload an oracle_api.Scene through the C ABI."""
import numpy as np

from dod_raytracer_b200 import capi


def upload(scene, device=0) -> capi.Scene:
    """oracle_api.Scene (host arrays in the reference's layouts) -> capi.Scene (resident on the GPU)."""
    g = capi.Scene(device)
    if len(scene.nodes):
        g.set_kdtree(scene.nodes, scene.tri_lanes, scene.bounds)
    if len(scene.spheres):
        g.set_spheres(scene.sphere_lanes, len(scene.spheres))
    if len(scene.planes):
        g.set_planes(scene.plane_lanes, len(scene.planes))
    if len(scene.cylinders):
        g.set_cylinders(scene.cylinders.view(capi.CYL_DT))
    if len(scene.boxes):
        g.set_boxes(scene.box_lanes, len(scene.boxes))
    g.set_epsilon(scene.epsilon)
    return g


def oracle_scene(hs):
    """dod_raytracer_b200.host.HostScene (after build_tree) -> oracle_api.Scene over copies of its arrays."""
    from oracle_api import Scene
    return Scene.from_host_arrays(hs.arrays())


def assert_hits_equal(got: np.ndarray, want: np.ndarray, rays=None, what=""):
    """Bit-exact comparison of dodrt_hit arrays; any-hit rays compare hit/miss only."""
    assert len(got) == len(want)
    if rays is not None:
        closest = (rays["flags"] & 1) == 0
    else:
        closest = np.ones(len(got), bool)
    bad_prim = got["prim"] != want["prim"]
    assert not bad_prim.any(), f"{what}: {int(bad_prim.sum())} prim mismatches, first at {int(np.argmax(bad_prim))}: " \
                               f"{got[np.argmax(bad_prim)]} vs {want[np.argmax(bad_prim)]}"
    g, w = got[closest], want[closest]
    for field in ("t", "u", "v"):
        bad = g[field].view(np.uint32) != w[field].view(np.uint32)
        assert not bad.any(), f"{what}: {int(bad.sum())} `{field}` bit mismatches, first {g[np.argmax(bad)]} vs {w[np.argmax(bad)]}"
