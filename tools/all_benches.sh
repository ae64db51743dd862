# bench lines of every workload at N = 1 (profiles/r01_bench_*_n1.json)
for w in UT C1 C2 C3; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r01_bench_${w}_n1.json 2>/dev/null; done
python bench.py --workload C5 --steps 10 --warmup 3 > gpurun_out/r01_bench_c5_n1.json 2>/dev/null
python bench.py --workload C4o --steps 5 --warmup 3 --no-cpu > gpurun_out/r01_bench_c4o_n1.json 2>/dev/null
python bench.py --impl reference > gpurun_out/r01_bench_c4_reference.json 2>/dev/null
python bench.py > gpurun_out/r01_bench_c4_n1.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r01_bench_*_n1.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; e=d["e2e"]
        print(f.split("/")[-1], "%.4g"%d["value"], "ms/step %.4g"%d["ms_per_step"], "frac %.3f"%r["frac"], "e2e %.4g"%e["value"], "p50", e.get("block_latency_ms_p50"), "paced", e.get("block_latency_paced_ms_p50"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex: print(f, "FAILED", ex)
PY
