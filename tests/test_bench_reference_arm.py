"""CPU: bench.py's reference arm (`--impl reference`) -- the reference's own code (oracle/_ref) on the host cores, printed in
the bench contract's JSON form -- runs without a GPU and without any of the product's libraries."""
import json
import os
import subprocess
import sys

import pytest

from oracle_api import have_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference here)")
def test_reference_arm_prints_the_contract_line_without_the_product_libraries(tmp_path):
    # the product's libraries are made unloadable: the arm must not need them (it maps oracle/_ref only)
    env = dict(os.environ, DODRT_LIB=str(tmp_path / "missing_cuda.so"), DODRT_HOST_LIB=str(tmp_path / "missing_host.so"),
               CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "teapot1080"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s (primary+shadow)" and line["unit"] == "Mrays/s"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["higher_is_better"] is True
    assert line["config"]["workload"] == "teapot1080" and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and abs(line["ms_per_step"] * 1e-3 * line["value"] * 1e6 - 2 * 1920 * 1080) < 1e-3 * 2 * 1920 * 1080


def test_gpu_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a CUDA device the product arm of the bench stops with a message and prints no
    metric line."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--workload", "teapot1080"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr
    assert '"metric"' not in r.stdout
