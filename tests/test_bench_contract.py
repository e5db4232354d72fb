"""bench.py contract checks that need no GPU: the reference arm's JSON line, and that the GPU arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "3", "--cpu-fraction", "0.004")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    d = json.loads(lines[-1])                                   # the JSON line is the last thing on stdout
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["config"]["workload"] == "sixteen_armadillos" and d["config"]["width"] == 3840 and d["config"]["height"] == 2160
    assert d["warmup"] >= 3 and d["steps"] == 1 and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "3")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_committed_bench_lines_carry_the_contract():
    # the bench lines kept under profiles/ (what DESIGN.md's tables are printed from) hold every key of the contract
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02f_bench_*_n1.json")))
    assert len(paths) == 5
    for p in paths:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                    "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
            assert key in d, (p, key)
        assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
        r = d["roofline"]
        assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r) and 0.0 < r["frac"] <= 1.0
        e = d["e2e"]
        assert e["d2h_bytes_per_step"] >= d["config"]["width"] * d["config"]["height"] * 4 and e["value"] < d["value"]
        assert e["two_frames_in_flight"]["last_frame_equals_render"] is True
        c = d["clocks"]
        assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        cb = d["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0
