# C5 (offline tensor-core render) probe: parity tests of the offline path, then the C5 bench line and two knobs
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "offline" 2>&1 | tail -5
for FL in 1 2; do
  SAFCONV_OFF_FLUSH=$FL timeout 300 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu --no-check > gpurun_out/c5_flush$FL.json 2> gpurun_out/c5_flush$FL.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c5_flush$FL.json")); r=d["roofline"]
    print("flush $FL: ms/step %.3f issued_frac %.3f"%(d["ms_per_step"], r["issued_frac"]), r["kernel_ms_per_render"], "e2e %.3g"%d["e2e"]["value"])
except Exception as ex: print("flush $FL FAILED", ex)
PY
done
timeout 300 python tools/offline_check.py 2>&1 | tail -8
