set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in C4 UT C3 C2 C1 C5; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/q_$w.json")); print("$w", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"], d.get("phases"), d.get("latency"))
except Exception as e: print("$w failed", e)
PY
done
