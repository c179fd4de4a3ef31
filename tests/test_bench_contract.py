"""The committed bench records (profiles/r02s3_bench_*gpu.json: the lines `bench.py` printed on the GPU boxes at the end of
round 2) carry every key of the measurement contract, and their figures are consistent with one another."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(n):
    path = os.path.join(ROOT, "profiles", f"r02s3_bench_{n}gpu.json")
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_record_has_the_contract_keys_and_consistent_numbers(n):
    d = _load(n)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "roofline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert "workload" in d["config"] and d["vs_baseline"] is None
    assert d["gpu_launches"] > 0
    # value = individuals of all ranks per second of the max-over-ranks step time
    M_total = d["config"]["M_total"]
    assert abs(d["value"] - M_total / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                      # copies inside the timed region cost something
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.0 < r["frac"] < 1.0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        c = d["cpu_baseline"]
        for key in ("value", "unit", "cores", "kind", "sample"):
            assert key in c, key
        assert c["kind"] in ("port", "reference") and c["unit"] == d["unit"]
    else:
        assert d["sharded_vs_single"]["ok"] is True and d["sharded_vs_single"]["hard_assignments_identical"] is True
