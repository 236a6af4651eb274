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


def test_both_arms_print_the_same_config():
    """The driver compares the `config` dicts of the two arms: they come from one function."""
    sys.path.insert(0, ROOT)
    import bench
    from badger_b200 import synth
    cfg = dict(synth.CONFIGS["C4"])
    c = bench.config_dict("C4", cfg, 4567717)
    assert c["workload"].startswith("C4: 20000000 simulated ONT reads") and c["threshold"] == 2 and c["distinct"] == 4567717
    sys.argv = ["bench.py"]
    a = bench.parse()
    assert a.config == "C4" and a.gpus == 1 and a.mode == "auto"


def test_committed_gpu_lines_follow_the_contract():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2*_bench_c4*.json")))
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d), f
        assert d["metric"] == "barcode_pairs_scored_per_s" and d["gpu_launches"] > 0 and d["scaling"] == "strong"
        assert d["warmup"] >= 3 and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "u32"
        assert d["config"]["workload"].startswith("C4") and d["config"]["threshold"] == 2
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
        assert 0 < d["e2e"]["value"] < d["value"]                         # copies inside the timed region cost something
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] == 1 and "c2" in d:
            assert d["c2"]["cli"]["outputs_identical"] is True and d["pipeline"]["reads_per_s"] > 0
        if "cpu_baseline" in d:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])


def test_round1_gpu_lines_follow_the_contract():
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
