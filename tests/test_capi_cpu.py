"""CPU: the C-ABI library loads, exports every symbol include/dodrt.h declares, its host-only helpers are
right, and without a GPU every compute entry point fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dod_raytracer_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dodrt.h")).read()
    return sorted(set(re.findall(r"DODRT_API\s+[\w\s\*]+?\b(dodrt_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = header_symbols()
    assert len(names) >= 21
    assert sorted(capi.EXPORTED_SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.dodrt_abi_version() == capi.ABI_VERSION == 2
    assert C.sizeof(capi.FrameBufferDesc) == 96


def test_struct_layouts_match_header():
    assert capi.RAY_DT.itemsize == 32 and capi.HIT_DT.itemsize == 16 and capi.CYL_DT.itemsize == 32
    assert C.sizeof(capi.Frame) == 44


def py_pixel_map(w, h, tw, th, first, stride):
    tiles_x, tiles_y = -(-w // tw), -(-h // th)
    out = []
    for tile in range(first, tiles_x * tiles_y, stride):
        tx, ty = tile % tiles_x, tile // tiles_x
        for i in range(tw * th):
            block, lane = i // 32, i % 32
            bpr = tw // 8
            col = tx * tw + (block % bpr) * 8 + lane % 8
            row = ty * th + (block // bpr) * 4 + lane // 8
            out.append(row * w + col if col < w and row < h else 0xFFFFFFFF)
    return np.array(out, np.uint32)


@pytest.mark.parametrize("w,h,tile", [(64, 32, (32, 32)), (100, 37, (32, 32)), (1920, 1080, (32, 32)), (50, 50, (8, 4)),
                                      (33, 9, (16, 8))])
def test_frame_pixel_map_partitions_the_image(w, h, tile):
    world = 3
    seen = np.zeros(w * h, np.int32)
    for rank in range(world):
        f = capi.Frame.make(w, h, tile=tile, first_tile=rank, tile_stride=world, compact=1)
        m = capi.frame_pixel_map(f)
        assert len(m) == capi.frame_local_pixels(f)
        assert (m == py_pixel_map(w, h, tile[0], tile[1], rank, world)).all()
        valid = m[m != 0xFFFFFFFF]
        np.add.at(seen, valid, 1)
    assert (seen == 1).all()  # every pixel belongs to exactly one rank


def test_bad_arguments_are_rejected():
    f = capi.Frame.make(64, 64, tile=(12, 4))
    with pytest.raises(capi.DodrtError) as e:
        capi.frame_local_pixels(f)
    assert e.value.code == -1 and "multiple of 8x4" in str(e.value)
    f = capi.Frame.make(64, 64, tile_stride=0)
    with pytest.raises(capi.DodrtError):
        capi.frame_local_pixels(f)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.DodrtError) as e:
        capi.Scene(0)
    assert e.value.code == -2  # DODRT_E_CUDA


def test_host_library_exports_every_declared_symbol():
    """include/dodrt_host.h vs libdodrt_host.so (the host side above the C ABI)"""
    from dod_raytracer_b200 import host
    text = open(os.path.join(ROOT, "include", "dodrt_host.h")).read()
    names = sorted(set(re.findall(r"DODRT_API\s+[\w\s\*]+?\b(dodrt_host_\w+)\s*\(", text)))
    assert len(names) >= 30 and "dodrt_host_build_tree_ex" in names
    lib = host.load()
    for n in names:
        assert getattr(lib, n) is not None, n


def test_headers_are_plain_c_and_cxx(tmp_path):
    """The drop-in boundary is a C ABI: include/dodrt.h and include/dodrt_host.h must compile as C99 on their own
    (plain pointers and sizes, no C++ or CUDA types), include/dodrt.hpp as C++17 with nothing but the C header."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("gcc") or not shutil.which("g++"):
        pytest.skip("no host compiler")
    for header in ("dodrt.h", "dodrt_host.h"):
        src = tmp_path / (header + ".c")
        src.write_text(f'#include "{header}"\nint main(void) {{ return 0; }}\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(root, "include"),
                        str(src)], check=True)
    src = tmp_path / "hpp.cpp"
    src.write_text('#include "dodrt.hpp"\nint main() { return 0; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "include"), str(src)],
                   check=True)
