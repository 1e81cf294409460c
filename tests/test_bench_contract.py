"""The bench lines committed under profiles/ keep the driver's contract (keys, types, consistency); CPU only."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.isfile(path):
        pytest.skip(name + " not recorded")
    return json.load(open(path))


@pytest.mark.parametrize("name,n", [("r2_bench_n1.json", 1), ("r2_bench_n2_scale.json", 2), ("r2_bench_n8_scale.json", 8)])
def test_our_arm_line(name, n):
    d = _line(name)
    assert d["metric"].startswith("queries/sec") and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["steps"] >= 10 and d["warmup"] >= 3
    Q = 3368 * n
    assert "Q=%d " % Q in d["config"]["workload"] and "model" not in d["config"]
    # value is whole-job throughput: all queries of a step / time of a step
    assert abs(d["value"] - Q / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == "queries/s" and e["h2d_bytes_per_step"] > 100e6 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                                   # host copies inside the timed region
    assert d["gpu_launches"] >= 10 * d["steps"]
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and isinstance(c["reasons"], list)
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    p = d["result"]["parity"]
    assert p["ok"] is True and p["oracle_subsample"]["cmc_bit_exact"] is True and p["oracle_subsample"]["mAP_abs_err"] < 1e-6
    if n > 1:
        assert p["sharded_equals_single_gpu"] is True and d["cpu_baseline"] is None
    else:
        b = d["cpu_baseline"]
        assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]


def test_reference_arm_line_matches_our_config():
    ref, ours = _line("r2_bench_reference_arm.json"), _line("r2_bench_n1.json")
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["higher_is_better"] == ours["higher_is_better"]
    assert ref["config"]["workload"] == ours["config"]["workload"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
    assert ref["cpu_baseline"]["value"] == ref["value"] and "512" in ref["cpu_baseline"]["sample"]


def test_both_arms_build_the_same_config():
    import bench
    for n in (1, 2, 8):
        assert bench.bench_config(n) == bench.bench_config(n) and "Q=%d " % (3368 * n) in bench.bench_config(n)["workload"]
