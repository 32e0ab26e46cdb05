"""CPU test of the bench.py contract: the reference arm (`--impl reference`, the oracle port on the host cores) must print ONE JSON line with
the keys the driver reads; the GPU arm must refuse to run without a CUDA device (no CPU fallback in the product path)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-pairs", "4"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "fragment_pairs_per_sec_match_ransac_svd" and j["unit"] == "pairs/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["higher_is_better"] is True and j["vs_baseline"] is None and j["value"] > 0
    assert j["config"]["workload"].startswith("BASELINE.json configs[1]")
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "4 pairs" in j["cpu_baseline"]["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a GPU")
def test_gpu_arm_refuses_to_run_without_cuda():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)
