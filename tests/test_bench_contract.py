"""The bench line contract (what the driver parses), checked on the committed final line of the round: the keys of the base
contract, the `roofline` / `cpu_baseline` / `e2e` objects of the hot-path tier, and their internal consistency."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _final_line():
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r02_bf16_v8_*wps_e2e*.json")))
    assert paths, "the final bench line of the round is committed under profiles/"
    for line in open(paths[-1]):
        if line.startswith("{"):
            return json.loads(line)
    raise AssertionError("no JSON line")


def test_bench_line_has_the_contract_keys_and_is_self_consistent():
    d = _final_line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "train windows/sec" and d["unit"] == "windows/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = windows of all ranks / device time per step
    windows = d["config"]["windows_per_step_per_gpu"] * d["n_gpus"]
    assert abs(d["value"] - windows / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac"] <= 1.0
    assert len(r["top_kernels"]) == 8 and all(0.0 < t["frac"] <= 1.0 for t in r["top_kernels"])
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] <= d["value"] * 1.02               # end to end cannot beat the device-timed step
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
