"""bench.py contract on CPU: the reference arm (`--impl reference`) is the one leg that runs without a GPU, so its JSON
line is checked here; the GPU arm's line is written by the same `emit` helper and checked by the driver on the B200."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--batch", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["unit"] == "crops/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "SwinIR-M" in d["metric"] and "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None  # BASELINE.md has no published number for this metric
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    from oracle import ref_shim
    if ref_shim.available():  # /root/reference here, baseline/_ref on the GPU box: the arm must then be the real reference
        assert cb["kind"] == "reference" and "image.optimize_parameters" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_gpu_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback behind the product arm: without CUDA it must fail, not print a number."""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert not any(ln.startswith("{") and '"value"' in ln for ln in out.stdout.splitlines())


def test_reference_arm_is_rank0_only():
    """Under torchrun the other ranks of the reference arm exit 0 without work or output."""
    import os
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
