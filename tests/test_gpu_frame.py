"""GPU (-m gpu): dodrt_trace_frame_device (a frame share: primary pass + shadow passes; with libdodrt_cuda_exp.so also
the one-launch frame kernels, DODRT_FUSED=1), result mirrors (a peer GPU's frame buffer, pinned host memory), frame
buffers shared between processes (CUDA IPC) and the several-GPUs-one-process API.  Everything must reproduce the bytes of
the separate primary / shadow entry points, which tests/test_gpu_parity.py pins to the oracle, the reference fixtures and
the reference itself."""
import os
import subprocess
import sys

import numpy as np
import pytest

from dod_raytracer_b200 import capi, host
from gpu_util import upload
from oracle_api import CLS_CYLINDER, CLS_PLANE, CLS_SPHERE, CLS_TREE, MISS
from scenes import GOLDEN, LIGHT0, teapot_scene

pytestmark = pytest.mark.gpu
ALL = CLS_SPHERE | CLS_PLANE | CLS_CYLINDER | CLS_TREE
LIGHTS2 = np.stack([LIGHT0, np.array([4.0, 4.3, 3.3], np.float32)])  # lights[0], lights[1] of main.cpp:284-285
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_frame(g, frame, xs, ys, lights, mirror=None, fused=False):
    """dodrt_trace_frame_device on torch buffers -> (hits, vis[lights, slots]) as numpy; fused: the A/B frame kernel"""
    import torch
    dev = torch.device("cuda", g.device)
    with torch.cuda.device(dev):
        d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
        slots = capi.frame_local_pixels(frame) if frame.compact else frame.width * frame.height
        d_hits = torch.full((slots, 16), 0xAB, dtype=torch.uint8, device=dev)
        d_vis = torch.full((len(lights), slots), 0xCD, dtype=torch.uint8, device=dev)
        os.environ["DODRT_FUSED"] = "1" if fused else "0"
        try:
            g.trace_frame_device(frame, d_xs.data_ptr(), d_ys.data_ptr(), lights, d_hits.data_ptr(), d_vis.data_ptr(), mirror,
                                 torch.cuda.current_stream().cuda_stream)
        finally:
            os.environ.pop("DODRT_FUSED", None)
        torch.cuda.synchronize(dev)
        return d_hits.cpu().numpy().reshape(-1).view(capi.HIT_DT), d_vis.cpu().numpy()


class _RawDeviceBytes:
    """a device address as something torch.as_tensor understands"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _read_frame_buffer(fb, w, h, nl, device=0):
    import torch
    hp, vp = fb.pointers()
    with torch.cuda.device(device):
        torch.cuda.synchronize()
        hits = torch.as_tensor(_RawDeviceBytes(hp, w * h * 16), device=f"cuda:{device}").cpu().numpy().view(capi.HIT_DT).copy()
        vis = (torch.as_tensor(_RawDeviceBytes(vp, nl * w * h), device=f"cuda:{device}").cpu().numpy().reshape(nl, w * h).copy()
               if nl else np.zeros((0, w * h), np.uint8))
    return hits, vis


@pytest.fixture(scope="module", params=[v for v in (-1, 3, 7, 0) if capi.variant_available(v)], ids=lambda v: f"variant{v}")
def teapot(request):
    """auto / plain fused / donating fused / a variant without a fused form (falls back to separate passes)"""
    g = upload(teapot_scene(full=True))
    g.set_kernel_variant(request.param)
    g.variant = request.param
    yield g
    g.close()


needs_experiments = pytest.mark.skipif(not capi.experiments_build(), reason="one-launch frame kernels are A/B experiments: "
                                       "libdodrt_cuda_exp.so only (run by tests/test_gpu_experiments.py)")


@needs_experiments
@pytest.mark.parametrize("w,h,tile", [(640, 360, (32, 32)), (500, 277, (16, 8)), (1920, 1080, (32, 32))])
def test_fused_launch_equals_separate_passes(teapot, w, h, tile):
    g = teapot
    xs, ys = host.ray_tables(w, h)
    frame = capi.Frame.make(w, h, classes=ALL, tile=tile)
    want_h, want_v = _device_frame(g, frame, xs, ys, LIGHTS2, fused=False)
    before = g.launch_count()
    got_h, got_v = _device_frame(g, frame, xs, ys, LIGHTS2, fused=True)
    assert got_h.tobytes() == want_h.tobytes() and got_v.tobytes() == want_v.tobytes()
    if g.variant in (-1, 3, 7):
        assert g.launch_count() - before <= 2  # one frame kernel (+ the tile-order helper), not 1 + lights passes
    assert want_v[0].sum() > 0 and want_v[1].tobytes() != want_v[0].tobytes()
    # primary only (no shadow queue)
    got_h0, _ = _device_frame(g, frame, xs, ys, LIGHTS2[:0])
    assert got_h0.tobytes() == want_h.tobytes()
    # A/B form of the frame kernel: per-tile ready queues instead of block-fused batches
    os.environ["DODRT_FRAME_QUEUES"] = "1"
    try:
        q_h, q_v = _device_frame(g, frame, xs, ys, LIGHTS2, fused=True)
    finally:
        os.environ.pop("DODRT_FRAME_QUEUES", None)
    assert q_h.tobytes() == want_h.tobytes() and q_v.tobytes() == want_v.tobytes()


def test_fused_tile_split_with_frame_buffer_mirror(teapot):
    """every rank of an image-tile split writes its pixels straight into ONE row-major frame buffer (here all ranks
    run one after the other on the same GPU; tests/test_gpu_multi.py does it across GPUs and processes)"""
    g = teapot
    w, h = 500, 277
    xs, ys = host.ray_tables(w, h)
    full_h, full_v = _device_frame(g, capi.Frame.make(w, h, classes=ALL), xs, ys, LIGHTS2, fused=False)
    # dodrt_trace_frame_device == the classic entry points (which test_gpu_parity.py pins to the oracle and the reference)
    ph = g.trace_primary(capi.Frame.make(w, h, classes=ALL), xs, ys)
    assert ph.tobytes() == full_h.tobytes()
    for l in range(2):
        assert g.trace_shadow(capi.Frame.make(w, h, classes=ALL), xs, ys, ph, LIGHTS2[l]).tobytes() == full_v[l].tobytes()
    for world, tile in ((2, (32, 32)), (5, (16, 8))):
        with capi.FrameBuffer.create(g, w, h, 2) as fb:
            hits0, vis0 = _read_frame_buffer(fb, w, h, 2)
            assert (hits0["prim"] == MISS).all() and not vis0.any()
            for rank in range(world):
                f = capi.Frame.make(w, h, classes=ALL, tile=tile, first_tile=rank, tile_stride=world, compact=1)
                m = capi.frame_pixel_map(f)
                ok = m != 0xFFFFFFFF
                # (the experiments build alternates between the product path and the one-launch frame kernel)
                lh, lv = _device_frame(g, f, xs, ys, LIGHTS2, mirror=fb, fused=capi.experiments_build() and rank % 2 == 1)
                assert lh[ok].tobytes() == full_h[m[ok]].tobytes() and (lh["prim"][~ok] == MISS).all()
                assert lv[:, ok].tobytes() == full_v[:, m[ok]].tobytes() and not lv[:, ~ok].any()
            hits, vis = _read_frame_buffer(fb, w, h, 2)
            assert hits.tobytes() == full_h.tobytes() and vis.tobytes() == full_v.tobytes()
    with capi.FrameBuffer.create(g, w, h, 1) as small:  # a mirror with fewer lights than the call is refused
        with pytest.raises(capi.DodrtError):
            _device_frame(g, capi.Frame.make(w, h, classes=ALL), xs, ys, LIGHTS2, mirror=small)


@pytest.mark.parametrize("chunks", [2, 4])
def test_frame_share_in_chunks_on_concurrent_streams(teapot, monkeypatch, chunks):
    """A rank's share is cut into chunks whose primary / shadow passes run on concurrent streams so that one chunk's tail
    overlaps the other's main phase.  Same bytes, locally and in the mirror, full frame and share."""
    g = teapot
    w, h = 1920, 1080
    xs, ys = host.ray_tables(w, h)
    for kw in (dict(), dict(first_tile=1, tile_stride=2, compact=1)):
        frame = capi.Frame.make(w, h, classes=ALL, **kw)
        monkeypatch.setenv("DODRT_FRAME_CHUNKS", "1")
        want_h, want_v = _device_frame(g, frame, xs, ys, LIGHTS2)
        monkeypatch.setenv("DODRT_FRAME_CHUNKS", str(chunks))
        with capi.FrameBuffer.create(g, w, h, 2) as fb:
            got_h, got_v = _device_frame(g, frame, xs, ys, LIGHTS2, mirror=fb)
            assert got_h.tobytes() == want_h.tobytes() and got_v.tobytes() == want_v.tobytes()
            fh, fv = _read_frame_buffer(fb, w, h, 2)
            if frame.compact:
                m = capi.frame_pixel_map(frame)
                ok = m != 0xFFFFFFFF
                assert fh[m[ok]].tobytes() == want_h[ok].tobytes() and fv[:, m[ok]].tobytes() == want_v[:, ok].tobytes()
            else:
                assert fh.tobytes() == want_h.tobytes() and fv.tobytes() == want_v.tobytes()


def test_host_buffers_zero_copy_equals_staged(teapot, monkeypatch):
    """dodrt_trace_frame with PINNED host buffers and DODRT_ZEROCOPY=1 lets the kernels store the results into them (no D2H
    copies; opt-in, measured slower than staging on this pool).  Same bytes, full-frame and compact (padded slots included)."""
    import torch
    g = teapot
    w, h = 1000, 562
    xs, ys = host.ray_tables(w, h)
    for kw in (dict(), dict(first_tile=1, tile_stride=3, compact=1)):
        frame = capi.Frame.make(w, h, classes=ALL, **kw)
        slots = capi.frame_local_pixels(frame) if frame.compact else w * h
        monkeypatch.setenv("DODRT_ZEROCOPY", "0")
        want_h, want_v = g.trace_frame(frame, xs, ys, LIGHTS2)
        monkeypatch.setenv("DODRT_ZEROCOPY", "1")
        ph = torch.full((slots, 16), 0x5A, dtype=torch.uint8).pin_memory().numpy().reshape(-1).view(capi.HIT_DT)
        pv = torch.full((2, slots), 0x5A, dtype=torch.uint8).pin_memory().numpy()
        before = g.launch_count()
        g.trace_frame(frame, xs, ys, LIGHTS2, ph, pv)
        assert ph.tobytes() == want_h.tobytes() and pv.tobytes() == want_v.tobytes()
        # primary only into pinned memory
        ph[:] = np.zeros(1, capi.HIT_DT)
        g.trace_frame(frame, xs, ys, LIGHTS2[:0], ph, None)
        assert ph.tobytes() == want_h.tobytes()
        assert g.launch_count() > before
        # default (no knob): staged copies into the pinned buffers
        monkeypatch.delenv("DODRT_ZEROCOPY")
        ph[:] = np.zeros(1, capi.HIT_DT)
        pv[:] = 7
        g.trace_frame(frame, xs, ys, LIGHTS2, ph, pv)
        assert ph.tobytes() == want_h.tobytes() and pv.tobytes() == want_v.tobytes()


def test_frame_buffer_shared_with_another_process(tmp_path):
    """CUDA IPC: a second PROCESS opens the exported descriptor and its kernels fill the owner's frame buffer
    (what every rank >= 1 of bench.py does with rank 0's frame)."""
    w, h = 480, 270
    scene = teapot_scene(full=True)
    xs, ys = host.ray_tables(w, h)
    with upload(scene) as g:
        full_h, full_v = _device_frame(g, capi.Frame.make(w, h, classes=ALL), xs, ys, LIGHT0[None, :], fused=False)
        with capi.FrameBuffer.create(g, w, h, 1) as fb:
            desc = tmp_path / "desc.bin"
            desc.write_bytes(fb.export().to_bytes())
            f0 = capi.Frame.make(w, h, classes=ALL, first_tile=0, tile_stride=2, compact=1)
            _device_frame(g, f0, xs, ys, LIGHT0[None, :], mirror=fb)  # rank 0 = this process
            env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests")]))
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "ipc_peer.py"), str(desc), "1", "2"],
                               env=env, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
            hits, vis = _read_frame_buffer(fb, w, h, 1)
            assert hits.tobytes() == full_h.tobytes() and vis.tobytes() == full_v.tobytes()


def test_multi_with_one_gpu_and_with_all_gpus():
    """dodrt_multi: N GPUs of this process fill ONE host frame (pinned: straight from the kernels; pageable: assembled
    in scenes[0]'s HBM over peer stores first).  Runs with however many GPUs the box has (1 is a valid group)."""
    import torch
    w, h = 1000, 562
    scene = teapot_scene(full=True)
    xs, ys = host.ray_tables(w, h)
    frame = capi.Frame.make(w, h, classes=ALL)
    with upload(scene) as g:
        want_h, want_v = _device_frame(g, frame, xs, ys, LIGHTS2, fused=False)
    ngpu = capi.device_count()
    for n in sorted({1, ngpu}):
        scenes = [upload(scene, d) for d in range(n)]
        try:
            with capi.Multi(scenes) as m:
                hits, vis = m.trace_frame(frame, xs, ys, LIGHTS2)  # pageable
                assert hits.tobytes() == want_h.tobytes() and vis.tobytes() == want_v.tobytes(), f"{n} GPUs pageable"
                ph = torch.zeros((w * h, 16), dtype=torch.uint8).pin_memory().numpy().reshape(-1).view(capi.HIT_DT)
                pv = torch.zeros((2, w * h), dtype=torch.uint8).pin_memory().numpy()
                for _ in range(2):
                    m.trace_frame(frame, xs, ys, LIGHTS2, ph, pv)
                assert ph.tobytes() == want_h.tobytes() and pv.tobytes() == want_v.tobytes(), f"{n} GPUs pinned"
        finally:
            for s in scenes:
                s.close()


@pytest.mark.parametrize("w,h", [(1003, 77), (4104, 40), (16, 3)])
def test_multi_pinned_bands_ragged_and_wide_frames(w, h, monkeypatch):
    """The staged-band form of dodrt_multi_trace_frame (pinned host frame): widths that are no multiple of 8, frames
    wider than the tile limit (two bands per row, 2-D copies), fewer bands than GPUs; and the zero-copy form (opt-in)."""
    import torch
    scene = teapot_scene(full=True)
    xs, ys = host.ray_tables(w, h)
    frame = capi.Frame.make(w, h, classes=ALL)
    with upload(scene) as g:
        want_h, want_v = _device_frame(g, frame, xs, ys, LIGHTS2, fused=False)
    scenes = [upload(scene, d) for d in range(capi.device_count())]
    try:
        with capi.Multi(scenes) as m:
            for zero_copy in ("0", "1"):
                monkeypatch.setenv("DODRT_ZEROCOPY", zero_copy)
                ph = torch.zeros((w * h, 16), dtype=torch.uint8).pin_memory().numpy().reshape(-1).view(capi.HIT_DT)
                pv = torch.zeros((2, w * h), dtype=torch.uint8).pin_memory().numpy()
                for _ in range(2):
                    m.trace_frame(frame, xs, ys, LIGHTS2, ph, pv)
                assert ph.tobytes() == want_h.tobytes() and pv.tobytes() == want_v.tobytes(), f"zero-copy {zero_copy}"
    finally:
        for s in scenes:
            s.close()


def test_scene_on_another_device_than_the_current_one():
    """streams / pools must be created on the scene's device, not on the caller's current one."""
    import torch
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    scene = teapot_scene(full=True)
    w, h = 320, 180
    xs, ys = host.ray_tables(w, h)
    frame = capi.Frame.make(w, h, classes=ALL)
    torch.cuda.set_device(0)
    with upload(scene, 0) as g0, upload(scene, 1) as g1:
        want = g0.trace_frame(frame, xs, ys, LIGHT0[None, :])
        got = g1.trace_frame(frame, xs, ys, LIGHT0[None, :])  # current device is still 0
        assert want[0].tobytes() == got[0].tobytes() and want[1].tobytes() == got[1].tobytes()
        rays = np.zeros(100, capi.RAY_DT)
        rays["d"][:, 2] = 1
        rays["o"][:, 2] = -4.9
        rays["clip"] = np.inf
        assert g1.intersect(rays, ALL).tobytes() == g0.intersect(rays, ALL).tobytes()


def _render_host_scene():
    hs = host.HostScene()
    hs.add_reference_scene(1, 16)
    hs.add_mesh_file(f"{GOLDEN}/teapot.dodm")
    hs.build_tree()
    return hs


def test_render_tile_split_and_multi_gpu():
    """rayTrace on the GPU (dodrt_render) renders only the call's tiles; the tiles of all ranks together -- compact or
    full-frame, one GPU after the other or all GPUs of the process at once (dodrt_multi_render, one rgb frame filled in
    place) -- are the pixels of the one-GPU render, which test_gpu_parity.py pins to the unmodified reference."""
    import torch
    from dod_raytracer_b200 import workloads
    w, h, depth = 250, 141, 4
    xs, ys = host.ray_tables(w, h)
    hs = _render_host_scene()
    with hs.upload(0, shading=True) as g:
        want = g.render(capi.Frame.make(w, h, classes=ALL), xs, ys, workloads.REFERENCE_LIGHTS, depth)
        assert want.mean() > 10
        # the per-bounce spatial sort of the hit points only changes the order in which rays are traced
        for knobs in ({"DODRT_RENDER_SORT": "0"}, {"DODRT_RENDER_CELL_BITS": "6", "DODRT_RENDER_SORT_FROM": "0"}):
            os.environ.update(knobs)
            try:
                other = g.render(capi.Frame.make(w, h, classes=ALL), xs, ys, workloads.REFERENCE_LIGHTS, depth)
            finally:
                for k in knobs:
                    os.environ.pop(k, None)
            assert other.tobytes() == want.tobytes(), knobs
        for world, tile in ((3, (32, 32)), (4, (16, 8))):
            full = np.zeros((h * w, 3), np.uint8)
            comp = np.zeros((h * w, 3), np.uint8)
            for rank in range(world):
                f0 = capi.Frame.make(w, h, classes=ALL, tile=tile, first_tile=rank, tile_stride=world, compact=0)
                part = g.render(f0, xs, ys, workloads.REFERENCE_LIGHTS, depth).reshape(-1, 3)
                f1 = capi.Frame.make(w, h, classes=ALL, tile=tile, first_tile=rank, tile_stride=world, compact=1)
                m = capi.frame_pixel_map(f1)
                ok = m != 0xFFFFFFFF
                mine = np.zeros(h * w, bool)
                mine[m[ok]] = True
                assert not part[~mine].any()  # other ranks' pixels stay black
                full[mine] = part[mine]
                c = g.render(f1, xs, ys, workloads.REFERENCE_LIGHTS, depth)
                assert c.shape == (len(m), 3) and not c[~ok].any()
                comp[m[ok]] = c[ok]
            assert full.tobytes() == want.tobytes() and comp.tobytes() == want.tobytes(), (world, tile)
    for n in sorted({1, capi.device_count()}):
        scenes = [hs.upload(d, shading=True) for d in range(n)]
        try:
            with capi.Multi(scenes) as m:
                got = m.render(capi.Frame.make(w, h, classes=ALL), xs, ys, workloads.REFERENCE_LIGHTS, depth)  # pageable
                assert got.tobytes() == want.tobytes(), f"{n} GPUs pageable"
                pinned = torch.zeros((h, w, 3), dtype=torch.uint8).pin_memory().numpy()
                m.render(capi.Frame.make(w, h, classes=ALL), xs, ys, workloads.REFERENCE_LIGHTS, depth, pinned)
                assert pinned.tobytes() == want.tobytes(), f"{n} GPUs pinned"
        finally:
            for s in scenes:
                s.close()


def test_ragged_and_degenerate_frames(oracle):
    """Edge cases of the frame description: 1x1 and 7x3 images (smaller than one 8x4 block), tiles larger than the image,
    the largest tile the ABI accepts, a share with no tile at all, 16 lights, no light -- against the oracle."""
    scene = teapot_scene(full=True)
    lights16 = np.stack([np.array([np.cos(k) * 3.0, 2.0 + 0.1 * k, np.sin(k) * 3.0], np.float32) for k in range(16)])
    with upload(scene) as g:
        for (w, h, tile) in ((1, 1, (8, 4)), (7, 3, (8, 4)), (33, 17, (64, 64)), (100, 50, (4096, 4096)), (97, 61, (16, 4))):
            xs, ys = host.ray_tables(w, h)
            want = oracle.trace_primary(scene, w, h, ALL, nthreads=4)
            frame = capi.Frame.make(w, h, classes=ALL, tile=tile)
            hits, vis = g.trace_frame(frame, xs, ys, lights16)
            assert hits.tobytes() == want.tobytes(), (w, h, tile)
            for l in (0, 7, 15):
                assert vis[l].tobytes() == oracle.trace_shadow(scene, w, h, ALL, want, lights16[l], nthreads=4).tobytes(), (w, h, l)
            h0, v0 = g.trace_frame(frame, xs, ys, lights16[:0])
            assert h0.tobytes() == want.tobytes() and v0.size == 0
            # a rank beyond the last tile has nothing to do and says so
            empty = capi.Frame.make(w, h, classes=ALL, tile=tile, first_tile=10 ** 6, tile_stride=2, compact=1)
            assert capi.frame_local_pixels(empty) == 0
            eh, ev = g.trace_frame(empty, xs, ys, lights16[:1])
            assert len(eh) == 0
        with pytest.raises(capi.DodrtError):  # tile sides are bounded (checkFrame)
            g.trace_frame(capi.Frame.make(64, 64, classes=ALL, tile=(8192, 4)), *host.ray_tables(64, 64), lights16[:1])
        with pytest.raises(capi.DodrtError):  # more lights than the render / frame kernels carry
            import torch
            d = torch.zeros(64, dtype=torch.uint8, device="cuda")
            g.trace_frame_device(capi.Frame.make(8, 4, classes=ALL), d.data_ptr(), d.data_ptr(), np.zeros((17, 3), np.float32),
                                 d.data_ptr(), d.data_ptr())
