"""SASS audit of the packed-fp32 first stage (variant 6, csrc/dodrt_device.cuh "packed fp32 pairs"; an A/B experiment that
lives in lib/libdodrt_cuda_exp.so since round 2 -- the product library must not contain packed fp32 arithmetic at all).

ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false, which would change the bits of
det / dot(T, pvec) relative to the reference's un-fused AVX arithmetic (triangle.cpp:66-93).  The kernels therefore
multiply with fma.rn.f32x2(a, b, -0.0 pair from a kernel parameter).  This test disassembles the built library and
checks the property that makes that exact: every packed multiply-add in the library has the uniform-register
-0.0 pair as its addend (i.e. it is a plain IEEE multiply), and no packed multiply was left for ptxas to fuse."""
import os
import re
import shutil
import subprocess

import pytest

from dod_raytracer_b200 import capi


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_packed_multiplies_are_not_contracted():
    if not os.path.exists(capi.EXP_LIB_PATH) or not os.path.exists(capi.LIB_PATH):
        pytest.skip("libraries not built")
    product = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], check=True, capture_output=True, text=True).stdout
    assert not re.search(r"\b(FFMA2|FMUL2|FADD2)\b", product), "packed fp32 arithmetic in the product library"
    sass = subprocess.run(["cuobjdump", "-sass", capi.EXP_LIB_PATH], check=True, capture_output=True, text=True).stdout
    ffma2 = [l for l in sass.splitlines() if re.search(r"\bFFMA2\b", l)]
    fadd2 = [l for l in sass.splitlines() if re.search(r"\bFADD2\b", l)]
    fmul2 = [l for l in sass.splitlines() if re.search(r"\bFMUL2\b", l)]
    assert len(ffma2) >= 48 and len(fadd2) >= 40, "the packed first stage is missing from the library"
    assert not fmul2, f"FMUL2 present (ptxas may fuse it with its consumer): {fmul2[:3]}"
    for line in ffma2:
        ops = line.split("FFMA2", 1)[1].split(";")[0].split(",")
        assert len(ops) == 4, line
        addend = ops[3].strip()
        assert re.fullmatch(r"UR\d+\.F32x2\.HI_LO", addend), f"FFMA2 with a live addend (a contraction): {line.strip()}"
