"""bench.py prints exactly one JSON line on stdout with the keys the driver reads; both arms print the SAME config."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(argv, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + argv, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0]), out


def test_reference_arm_prints_one_json_line():
    d, out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 900)
    assert d["impl"] == "reference" and d["unit"] == "vh/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("virtual heights/sec") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["scaling"] == "strong" and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("BASELINE configs[3] unit") and d["config"]["n_profiles"] == 65536
    cb = d["cpu_baseline"]
    # in the dev container the reference is mounted, so the arm must time the real thing from oracle/_ref
    want_kind = "reference" if os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "PyRayHF", "library.py")) else "port"
    assert cb["kind"] == want_kind and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert cb["one_process_value"] > 0                     # 1 process AND all cores
    assert d["e2e"] == {"value": d["value"], "unit": "vh/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_never_loads_the_product_package():
    """The driver records every .so the reference arm maps: none of the repo's may be among them, so the arm must not
    import pyrayhf_b200 (its __init__ loads the CUDA extension) -- not even for the input generator."""
    code = ("import sys, runpy; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0'];\n"
            "import bench\n"
            "bench.quiet_stdout(); s = bench.load_synth(); bench.workload_parameters(s); bench.config_dict(1)\n"
            "from oracle import cpu_baseline\n"
            "assert not [m for m in sys.modules if m.startswith('pyrayhf_b200')], sorted(sys.modules)\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libpyrayhf_b200' not in maps and '_prhf_fast' not in maps\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    src = open(os.path.join(ROOT, "bench.py")).read()
    ref_arm = src[src.index("def run_reference_arm"):src.index("# B200 arm")]
    assert "pyrayhf_b200" not in ref_arm
    assert "NCCL_DEBUG" not in src.replace("NCCL_DEBUG=INFO", "")     # the driver's NCCL log level is left alone


def test_both_arms_share_one_config():
    import bench
    assert bench.config_dict(4) == bench.config_dict(4)
    assert set(bench.config_dict(1)) == set(bench.config_dict(8))


@pytest.mark.gpu
def test_b200_arm_prints_one_json_line_with_roofline_and_e2e():
    d, out = _run(["--steps", "3", "--warmup", "3", "--no-extras"], 1500)
    assert d["unit"] == "vh/s" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["dtype"] == "f64"
    assert d["value"] > 1e7 and d["data"] == "synthetic" and d["vs_baseline"] is None and d["scaling"] == "strong"
    import bench
    assert d["config"] == bench.config_dict(1)
    r = d["roofline"]
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert e["unit"] == "vh/s" and 0 < e["value"] and e["h2d_bytes_per_step"] > 65536 * 620 * 3 * 8 - 1
    assert e["d2h_bytes_per_step"] >= 65536 * 174 * 8
    assert d["gpu_launches"] >= 2 * d["steps"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["one_process_value"] > 0
    par = d["parity"]
    assert par["resident_equals_e2e_bitwise"] is True
    assert par["x_mode"]["nan_mask_mismatches"] == 0 and par["x_mode"]["max_rel_err_vs_reference"] <= 1e-9
    assert par["o_mode"]["nan_mask_mismatches"] == 0 and par["o_mode"]["max_rel_err_vs_long_double_truth"] <= 1e-9
    lat = d["latency"]
    assert lat["workload"].startswith("BASELINE configs[1]") and 0 < lat["roofline"]["frac"] < 1
    assert lat["e2e_us_per_call"] > lat["device_us_per_call"] > 0
