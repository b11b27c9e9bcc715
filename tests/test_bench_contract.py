"""CPU: bench.py's reference arm honours the driver's JSON contract (it is the one bench leg that runs without a GPU), and
the native arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                          "--cpu-envs", "2"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_native_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return                                     # covered by the real bench run on the GPU box
    out = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "3"], cwd=ROOT, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode != 0 and not any(ln.startswith("{") for ln in out.stdout.splitlines())
