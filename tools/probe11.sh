for W in C4 C4g8; do
  timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-secondary --no-check --e2e-blocks 1000 > gpurun_out/h_${W}.json 2> gpurun_out/h_${W}.err
  python -c "
import json; d=json.load(open('gpurun_out/h_${W}.json')); e=d['e2e']; print('$W: e2e %.4g p50 %.4f ms p99 %.4f paced p50 %.4f p99 %.4f'%(e['value'], e['block_latency_ms_p50'], e['block_latency_ms_p99'], e.get('block_latency_paced_ms_p50',0), e.get('block_latency_paced_ms_p99',0)))"
done
timeout 120 python tools/timeline.py C4g8 16 2>&1 | grep timeline | tail -8
