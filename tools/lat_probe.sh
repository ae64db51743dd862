# latency configs: GPU parity suite, then C1 / C2 host-API latency with and without the cluster kernel
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for W in C1 C2; do
  for CL in 1 0; do
    SAFCONV_SMALL_CLUSTER=$CL timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/lat_${W}_cl$CL.json 2> gpurun_out/lat_${W}_cl$CL.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/lat_${W}_cl$CL.json")); e=d["e2e"]
    print("$W cluster=$CL: p50 %.2f us p99 %.2f us paced p50 %.2f us; parity %s"%(1e3*e["block_latency_ms_p50"],1e3*e["block_latency_ms_p99"],1e3*e.get("block_latency_paced_ms_p50",0), d.get("parity_rel_l2")))
except Exception as ex: print("$W cluster=$CL FAILED", ex)
PY
  done
done
timeout 300 python tools/offline_check.py 1024 8192 121 64 360 2>&1 | head -2
