"""GPU (-m gpu): the COMPILED integration -- integration/_build/dod_raytracer_gpu = the reference's own host side
(Config::Load, Sphere/Plane/Cylinder::create, Mesh::Create, KDTree::buildTree, stbi_write_png; translation units compiled
unmodified from /root/reference/src by integration/Makefile) with rayTrace (main.cpp:273-347) replaced by one call into
libdodrt_cuda.so through integration/dodrt_adapter.hpp.  The image must match the reference's own rayTrace, run by the
same binary on the same scene, within 1/255 per channel (north_star), and the one-ray wrapper of include/dodrt.hpp must
agree bit for bit with KDTree::intersect (kdtree.h:13)."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from dod_raytracer_b200 import capi
from scenes import GOLDEN

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "integration", "_build", "dod_raytracer_gpu")


def _png_size(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    assert raw[12:16] == b"IHDR"
    w, h, depth, ctype = struct.unpack(">IIBB", raw[16:26])
    assert depth == 8 and ctype == 2  # 8-bit RGB, what stbi_write_png(..., STBI_rgb, ...) writes (main.cpp:396)
    # the IDAT stream must inflate to h * (1 + 3 w) bytes (one filter byte per row)
    idat, pos = b"", 8
    while pos < len(raw):
        n, tag = struct.unpack(">I4s", raw[pos:pos + 8])
        if tag == b"IDAT":
            idat += raw[pos + 8:pos + 8 + n]
        pos += 12 + n
    assert len(zlib.decompress(idat)) == h * (1 + 3 * w)
    return w, h


@pytest.mark.skipif(not os.path.exists(BINARY), reason="integration/_build/dod_raytracer_gpu not built (needs /root/reference)")
@pytest.mark.parametrize("gpus", [1, 0])
def test_reference_program_with_the_gpu_path(tmp_path, gpus):
    if gpus == 0:
        gpus = capi.device_count()
        if gpus < 2:
            pytest.skip("needs 2 GPUs for the multi-GPU run")
    w, h = 320, 180
    (tmp_path / "config.ini").write_text(f"Width: {w}\nHeight: {h}\n")
    cmd = [BINARY, "--config", str(tmp_path / "config.ini"), "--mesh", os.path.join(GOLDEN, "teapot.dodm"), "--seed", "1",
           "--out", str(tmp_path / "output.png"), "--raw", str(tmp_path / "gpu.rgb"), "--cpu-raw", str(tmp_path / "cpu.rgb"),
           "--gpus", str(gpus)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "0 of 64 differ" in r.stdout, r.stdout
    assert _png_size(tmp_path / "output.png") == (w, h)
    gpu = np.fromfile(tmp_path / "gpu.rgb", np.uint8).astype(np.int32)
    cpu = np.fromfile(tmp_path / "cpu.rgb", np.uint8).astype(np.int32)
    assert gpu.size == cpu.size == w * h * 3
    diff = np.abs(gpu - cpu)
    assert diff.max() <= 1, f"max pixel difference {diff.max()}"
    assert (diff == 0).mean() > 0.999
    assert cpu.reshape(h, w, 3).mean() > 10  # a real image, not black
