"""The JSON line bench.py prints (the driver's contract), checked on the committed lines of
the last measured runs (profiles/r01_bench_*.json) and on bench.py's argument surface."""
import json
import os
import subprocess
import sys

import pytest
from conftest import REPO

PROFILES = os.path.join(REPO, "profiles")


def _load(name):
    with open(os.path.join(PROFILES, name)) as f:
        return json.load(f)


@pytest.mark.parametrize("name,n", [("r01_bench_n1.json", 1), ("r01_bench_n2.json", 2), ("r01_bench_n4.json", 4),
                                    ("r01_bench_n8.json", 8)])
def test_our_arm_line_has_the_contract_keys(name, n):
    d = _load(name)
    with open(os.path.join(REPO, "BASELINE.json")) as f:
        base = json.load(f)
    assert d["metric"] == "patristic_distance_pairs_per_sec" and "pairs/sec" in base["metric"]
    assert d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["n_gpus"] == n and d["steps"] >= 1 and d["warmup"] >= 3
    assert d["vs_baseline"] is None and base["published"] == {}
    assert d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "cfg2" in d["config"]["workload"] and "model" not in d["config"]
    assert "larger than L2" in d["config"]["l2"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["traffic"] is None or 0.9e9 < r["traffic"] < 2.0e9  # ~ the 1.6 GB algorithmic bytes per launch
    per_gpu = d["value"] / n
    assert r["achieved"] == pytest.approx(16 * per_gpu / 1e9, rel=0.05)
    e = d["e2e"]
    assert e["unit"] == "pairs/s" and e["h2d_bytes_per_step"] == 16 * e["pairs_per_step_per_gpu"]
    assert e["d2h_bytes_per_step"] == 8 * e["pairs_per_step_per_gpu"] and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] == d["steps"]
    c = d["clocks"]
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["unit"] == "pairs/s" and b["sample"]
        assert d["parity_vs_reference_sample"] is True
    else:
        assert d["cpu_baseline"] is None
    # weak scaling: N GPUs do N times the work in about the same time
    assert per_gpu > 5e10


@pytest.mark.parametrize("name,n", [("r02_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n4.json", 4),
                                    ("r02_bench_n8.json", 8)])
def test_round2_line_reports_the_reference_signature_and_its_roofline(name, n):
    """Round 2: e2e is the reference's own call (pageable numpy in, fresh array out, no out=
    kwarg, no pinned buffers handed in), measured over `steps` calls, with a copy roofline
    measured in the same run; at N > 1 the cfg4 collective is inside the timed call and the
    all-reduced r is checked."""
    d = _load(name)
    e = d["e2e"]
    assert "out=" not in e["call"] and "pinned" not in e["call"] and "fresh" in e["call"]
    assert e["steps"] == d["steps"] and e["matches_device_path"] is True
    r = e["roofline"]
    assert r["unit"] == "GB/s" and r["frac"] == pytest.approx(e["value"] / r["peak_pairs_per_s"])
    assert r["frac"] >= 0.85 and ("%d rank" % n) in r["what"]
    assert e["value"] >= 3e9
    c4 = d["other_workloads"]["cfg4_pearson_sampler"]
    if n > 1:
        assert "ncclAllReduce" in c4["workload"] and c4["allreduce_check"]["ok_1e-12"] is True
        assert c4["samples_in_r"] == n * 125_000_000
    else:
        assert d["other_workloads"]["cfg5_matrix"]["parity_vs_oracle"]["bit_exact"] is True
        assert d["other_workloads"]["cfg5_matrix"]["parity_vs_oracle"]["sampled_elements"] >= 1_000_000


def test_round2_reference_arm_runs_the_full_step():
    d = _load("r02_bench_reference_arm.json")
    ours = _load("r02_bench_n1.json")
    assert d["impl"] == "reference" and d["config"]["workload"] == ours["config"]["workload"]
    per_step = d["config"].get("pairs_per_step_per_gpu", d["config"].get("pairs_per_step"))
    assert per_step == ours["config"]["pairs_per_step_per_gpu"] == 100_000_000
    assert d["cpu_baseline"]["kind"] == "reference"


def test_reference_arm_line():
    d = _load("r01_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "patristic_distance_pairs_per_sec"
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["gpu_launches"] == 0


def test_bench_cli_surface():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
    # N > 1 without torchrun must refuse rather than silently run one rank
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, env={**os.environ, "WORLD_SIZE": "1"})
    assert out.returncode == 2 and "torchrun" in out.stderr
