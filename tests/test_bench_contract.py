"""CPU: the reference arm of bench.py prints exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "vh/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("virtual heights/sec") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("configs[1]")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "vh/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


import pytest  # noqa: E402


@pytest.mark.gpu
def test_b200_arm_prints_one_json_line_with_roofline_and_e2e():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3",
                          "--no-batched"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["unit"] == "vh/s" and d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["dtype"] == "f64"
    assert d["value"] > 1e6 and d["data"] == "synthetic" and d["vs_baseline"] is None
    r = d["roofline"]
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert e["unit"] == "vh/s" and 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] >= d["steps"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0
    assert d["parity"]["nan_mask_mismatches"] == 0 and d["parity"]["max_rel_err_vs_numpy_port"] <= 1e-9
