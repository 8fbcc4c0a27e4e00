"""bench.py's output contract (one JSON line on stdout with the keys the driver reads), checked on the CPU with the
reference arm (the oracle port) and on the committed line of the last GPU run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                        "--ref-rows", "16"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # exactly one line on stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["warmup"] >= 3 and d["steps"] == 1
    assert d["metric"].startswith("tree-pairs/sec") and d["unit"] == "tree-pairs/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and "model" not in d["config"]


def _check_gpu_line(d):
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "cpu_baseline"} <= set(d)
    assert d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and d["roofline"]["bound"] in ("hbm", "tensor")
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]                    # the end-to-end number is not a copy of the resident one
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    for name in ("i8", "popc"):                              # the two ways the north star names, measured beside the main line
        assert d["other_methods"][name]["bit_identical_means"] is True


@pytest.mark.gpu
def test_live_gpu_line_has_the_contract_keys():
    """The line the CURRENT bench.py and kernels print on this box (short run: cfg2 with the int8 / popcount
    comparison and a small CPU sample; the cfg3 and cfg5 blocks have their own tests)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-ess", "--no-cfg5",
                        "--no-live-peaks", "--cpu-rows", "64"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_gpu_line(d)
    assert d["steps"] == 3 and d["warmup"] == 3 and d["config"]["workload"].startswith("cfg2")
    assert d["e2e"]["bit_identical_to_resident"] is True and len(d["S_checksum"]) == 16


def test_committed_gpu_line_is_a_record_of_the_same_contract():
    # an ARTIFACT check (documentation of the last full GPU run), not a check of the current code: see the live test above
    path = os.path.join(ROOT, "profiles", "r02_bench_n1_final.json")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    _check_gpu_line(d)
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert {"ess", "cfg5"} <= set(d) and d["ess"]["roofline"]["bound"] == "fp64" and d["cfg5"]["scaling"] == "strong"
