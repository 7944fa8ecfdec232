"""bench.py's reference arm (no GPU needed): exactly ONE JSON line on stdout with the keys the driver reads; ranks > 0 of a
torchrun launch print nothing and exit 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env, *args):
    env = dict(os.environ, PMCTF_BENCH_REF_HW="128x192", **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], env=env, cwd=ROOT,
                          capture_output=True, text=True, timeout=300)


def test_reference_arm_prints_one_json_line(conv_mode):
    if conv_mode != "tensor":
        return  # once is enough: the arm chooses its own arithmetic
    r = run({}, "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "1080p GOP-16 MCTF frames/s" and d["unit"] == "frames/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["value"] - 2.0 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["config"]["workload"]


def test_reference_arm_other_ranks_are_silent(conv_mode):
    if conv_mode != "tensor":
        return
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_our_arm_json_contract_on_gpu(conv_mode):
    """python bench.py (one GOP per step to keep it short): exactly one stdout line with every key of the driver's contract."""
    if conv_mode != "tensor":
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--frames", "16",
                        "--no-cpu-baseline"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["metric"] == "1080p GOP-16 MCTF frames/s" and d["unit"] == "frames/s" and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert d["value"] > 0 and abs(d["value"] - 16 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "frames/s" and e["h2d_bytes_per_step"] > 16 * 1080 * 1920 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    rf = d["roofline"]
    assert rf["bound"] in ("hbm", "tensor") and rf["unit"] in ("GB/s", "TFLOP/s")
    assert rf["achieved"] > 0 and rf["peak"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and "traffic" in rf
