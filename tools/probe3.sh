timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "offline" 2>&1 | tail -3
for FP in 1 0; do
  SAFCONV_OFF_FFTP=$FP timeout 300 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu --no-check > gpurun_out/c5_fftp$FP.json 2> gpurun_out/c5_fftp$FP.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c5_fftp$FP.json")); r=d["roofline"]
    print("fftp $FP: ms/step %.3f issued_frac %.3f"%(d["ms_per_step"], r["issued_frac"]), r["kernel_ms_per_render"], "e2e %.3g"%d["e2e"]["value"])
except Exception as ex: print("fftp $FP FAILED", ex)
PY
done
for W in C1 C2; do
  SAFCONV_HOSTTRACE=1 timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-secondary --no-check > gpurun_out/lat_${W}.json 2> gpurun_out/lat_${W}.err
  grep "host trace" gpurun_out/lat_${W}.err | tail -3
  python -c "
import json; d=json.load(open('gpurun_out/lat_${W}.json')); e=d['e2e']; print('$W p50 %.2f us'%(1e3*e['block_latency_ms_p50']))"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:small_cluster -s 50 -c 10 --csv python bench.py --workload C2 --steps 2 --warmup 3 --no-cpu --no-secondary --no-check --e2e-blocks 200 2>/dev/null | grep small_cluster | cut -d, -f5,12- | head -10
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:small_cluster -s 50 -c 4 --csv python bench.py --workload C1 --steps 2 --warmup 3 --no-cpu --no-secondary --no-check --e2e-blocks 200 2>/dev/null | grep small_cluster | cut -d, -f5,12- | head -4
