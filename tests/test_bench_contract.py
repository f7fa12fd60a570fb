"""bench.py's contract checked on the CPU: the reference arm (`--impl reference`: the oracle port on host cores, no GPU
needed) prints exactly one JSON line with the keys the driver reads, also under a 2-rank launch (rank 0 alone works and
prints); the CUDA arm refuses to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run_bench(["--impl", "reference", "--workload", "c1-minimal", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pixel-frames/s" and d["unit"] == "pixel-frames/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert d["config"]["workload"] == "c1-minimal" and d["config"]["frames"] == 25
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "rows" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pixel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None  # BASELINE.md publishes no number for this metric


def test_reference_arm_under_a_multi_rank_launch_only_rank_0_works():
    r1 = run_bench(["--impl", "reference", "--gpus", "2", "--workload", "c1-minimal", "--steps", "1", "--warmup", "0"],
                   env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    r0 = run_bench(["--impl", "reference", "--gpus", "2", "--workload", "c1-minimal", "--steps", "1", "--warmup", "0"],
                   env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"})
    assert r0.returncode == 0, r0.stderr
    d = json.loads(r0.stdout.strip())
    assert d["impl"] == "reference" and d["n_gpus"] == 2


def test_cuda_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = run_bench(["--workload", "c1-minimal", "--steps", "1", "--warmup", "0"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
