timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for W in C1 C2; do
  for FW in 1 0; do
    SAFCONV_FLAG_WAIT=$FW SAFCONV_HOSTTRACE=1 timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/lat_${W}_fw$FW.json 2> gpurun_out/lat_${W}_fw$FW.err
    grep "host trace" gpurun_out/lat_${W}_fw$FW.err | tail -1
    python -c "
import json; d=json.load(open('gpurun_out/lat_${W}_fw$FW.json')); e=d['e2e']; print('$W flag_wait=$FW p50 %.2f us p99 %.2f paced %.2f parity %s'%(1e3*e['block_latency_ms_p50'],1e3*e['block_latency_ms_p99'],1e3*e.get('block_latency_paced_ms_p50',0),d.get('parity_rel_l2')))"
  done
done
timeout 300 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu --no-check > gpurun_out/c5_zpad.json 2> gpurun_out/c5_zpad.err
python - <<PY
import json
d=json.load(open("gpurun_out/c5_zpad.json")); r=d["roofline"]
print("C5: ms/step %.3f issued_frac %.3f"%(d["ms_per_step"], r["issued_frac"]), r["kernel_ms_per_render"], "e2e %.3g"%d["e2e"]["value"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:small_cluster -s 50 -c 6 --csv python bench.py --workload C2 --steps 2 --warmup 3 --no-cpu --no-secondary --no-check --e2e-blocks 200 2>/dev/null | grep small_cluster | cut -d, -f5,18- | head -6
