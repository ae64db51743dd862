"""The committed bench lines (profiles/r01_bench_*.json, written by bench.py on the GPU box) carry every key the
bench contract asks for.  Pure CPU: guards bench.py's output format against regressions."""
import json
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LINES = sorted((ROOT / "profiles").glob("r01_bench_*.json"))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


@pytest.mark.parametrize("path", LINES, ids=lambda p: p.stem)
def test_bench_line_has_contract_keys(path):
    text = path.read_text().strip()
    assert text.count("\n") == 0, "one JSON line"
    d = json.loads(text)
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["unit"] == "out-ch*samples/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    if d.get("impl") == "reference":
        assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
        assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
        return
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]                       # end-to-end is measured, not copied
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    if d["n_gpus"] == 1 and path.stem.endswith("c4_n1"):
        cb = d["cpu_baseline"]
        assert cb and {"value", "unit", "cores", "kind", "sample"} <= set(cb)


def test_default_workload_is_the_metric_config():
    d = json.loads((ROOT / "profiles" / "r01_bench_c4_n1.json").read_text())
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert "64-in x 64-out" in d["config"]["workload"] and "configs[3]" in d["config"]["workload"]
    assert "64-in x 64-out" in base["configs"][3]
