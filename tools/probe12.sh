for RS in 0 8 16 32; do
  export SAFCONV_TAIL_RESERVE_SMS=$RS
  for W in C4g8 C4; do
  timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-secondary --no-check --e2e-blocks 1000 > gpurun_out/rs_${W}_$RS.json 2> gpurun_out/rs_${W}_$RS.err
  python -c "
import json; d=json.load(open('gpurun_out/rs_${W}_$RS.json')); e=d['e2e']; print('reserve $RS $W: e2e %.4g p50 %.4f ms p99 %.4f paced p50 %.4f'%(e['value'], e['block_latency_ms_p50'], e['block_latency_ms_p99'], e.get('block_latency_paced_ms_p50',0)))"
  done
done
