"""bench.py's reference arm (the one leg that runs without a GPU): exactly one JSON line on stdout carrying the keys the driver
reads, the CPU legs' sample sizing, and rank != 0 staying silent under a multi-rank launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ); env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "impl", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["value"] > 0
    assert d["metric"].startswith("GN iterations/sec") and d["unit"] == "iterations/s" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "100000 states" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ); env.update(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env,
                       timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_sample_sizing():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.cpu_sample_states(4) == 100000            # the default 1 + 3 iterations: the whole workload
    assert bench.cpu_sample_states(23) == 100000           # 20 + 3
    assert 10000 <= bench.cpu_sample_states(103) < 100000  # long runs are cut to stay under the budget
    assert bench.cpu_sample_states(10 ** 6) == 10000
