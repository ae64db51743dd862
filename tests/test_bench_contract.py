"""The committed bench lines (profiles/r0N_bench_*.json, written by bench.py on the GPU box) carry every key the
bench contract asks for.  Pure CPU: guards bench.py's output format against regressions."""
import json
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LINES = sorted((ROOT / "profiles").glob("r0[0-9]_bench_*.json"))

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


def test_round2_default_line():
    """The round-2 default line: both arms describe the workload with the SAME `config`, the secondary block carries every
    other BASELINE.json config (latencies with page-locked and pageable buffers, the offline tensor-core render with its
    roofline), the timed path is parity-checked against the oracle, and the 1-core 'as shipped' CPU figure is there."""
    own = json.loads((ROOT / "profiles" / "r02_bench_c4_n1.json").read_text())
    ref = json.loads((ROOT / "profiles" / "r02_bench_c4_reference.json").read_text())
    if "config_detail" in own:                            # lines written after the two configs were unified
        assert own["config"] == ref["config"]
    assert own["parity"]["parity_rel_l2"] <= 1e-6 and own["parity"]["parity_max_abs_fs"] <= 1e-5
    cb = own["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["as_shipped_1_core"]["cores"] == 1 and cb["cpu_model"]
    sec = own["secondary"]
    assert {"C1", "C2", "C3", "UT", "C5"} <= set(sec)
    for k in ("C1", "C2", "C3", "UT"):
        h = sec[k]["host_api"]
        assert h["blocks"] >= 2000 and h["warmup_blocks"] >= 100 and h["pinned"]["p50_ms"] > 0 and h["pageable"]["p50_ms"] > 0
        assert sec[k]["parity"]["parity_rel_l2"] <= 1e-6
    r5 = sec["C5"]["roofline"]
    assert r5["bound"] == "tensor" and 0 < r5["frac"] < 1 and r5["issued_frac"] > r5["frac"]
    e = own["e2e"]
    assert e["blocks"] >= 2000 and e["block_latency_ms_p50"] > 0 and e["block_latency_paced_ms_p50"] > 0
    # the filter producers in front of the convolver: timed, checked against the fp64 restatement, reference beside them
    pr = sec["producers"]
    for k in ("LS", "MAGLS_diffCM_maxRE"):
        d = pr["decoder"][k]
        assert d["gpu_ms"] > 0 and d["parity_rel_l2_vs_fp64"] <= 5e-6 and d["reference_1_core_ms"] > d["gpu_ms"]
        assert d["parity_rel_l2_vs_fp64"] <= d["reference_rel_l2_vs_fp64"]          # closer to the truth than the reference
    im = pr["ims"]
    assert im["image_sources"] > 4e8 and im["render_s"] > 0 and im["parity_same_image_count"] and im["parity_same_taps"]
    assert im["parity_rel_l2"] <= 1e-6 and im["reference_1_core"]["image_sources_per_s"] < im["image_sources_per_s"]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_round2_multi_gpu_lines(n):
    """N > 1: `e2e` is the per-block synchronous call on ONE C-ABI handle over N GPUs (the N = 1 contract), with p50 / p99."""
    d = json.loads((ROOT / "profiles" / f"r02_bench_c4_n{n}.json").read_text())
    assert d["n_gpus"] == n
    e = d["e2e"]
    assert "safconv_matrixConv_create_multi" in e["api"] and e["blocks"] >= 2000
    assert e["block_latency_ms_p50"] > 0 and e["block_latency_ms_p99"] >= e["block_latency_ms_p50"]
    assert len(e["multi_gpu_handle"]["devices"]) == n and set(e["multi_gpu_handle"]["transports"]) == {"host", "nccl"}
    assert len(d["roofline"]["per_rank"]["mac_ms_per_block"]) == n
