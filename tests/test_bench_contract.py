"""The bench.py JSON contract (task statement "Measurement"): the reference arm is run here for real on a small read count;
the GPU arm's line is checked on the committed lines of the last GPU run (profiles/), which bench.py printed verbatim."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def test_reference_arm_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                   "--cpu-seconds", "1", "--reads", "30000"], cwd=ROOT, timeout=600)
    lines = [ln for ln in out.decode().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "barcode_pairs_scored_per_s" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "query rows" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_gpu_lines_follow_the_contract():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1h_bench_t*.json")))
    assert files
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d), f
        assert d["metric"] == "barcode_pairs_scored_per_s" and d["n_gpus"] == 1 and d["gpu_launches"] > 0
        assert d["warmup"] >= 3 and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "u32"
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
        assert 0 < d["e2e"]["value"] < d["value"]                         # copies inside the timed region cost something
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert "workload" in d["config"] and "l2" in d["config"]
        if "cpu_baseline" in d:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
        assert d["cli"]["outputs_identical"] is True and d["pipeline"]["reads_per_s"] > 0
