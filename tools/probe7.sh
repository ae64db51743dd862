timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for W in C1 C2; do
  SAFCONV_KSTAMPS=1 timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-secondary --no-check > gpurun_out/ks_${W}.json 2> gpurun_out/ks_${W}.err
  grep "small_cluster phases" gpurun_out/ks_${W}.err | tail -1
  SAFCONV_HOSTTRACE=1 timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/lat_${W}.json 2> gpurun_out/lat_${W}.err
  grep "host trace" gpurun_out/lat_${W}.err | tail -1
  python -c "
import json; d=json.load(open('gpurun_out/lat_${W}.json')); e=d['e2e']; print('$W p50 %.2f us p99 %.2f paced %.2f parity %s'%(1e3*e['block_latency_ms_p50'],1e3*e['block_latency_ms_p99'],1e3*e.get('block_latency_paced_ms_p50',0),d.get('parity_rel_l2')))"
done
