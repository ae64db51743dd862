for D in 1 2; do
  export SAFCONV_LA_DEPTH=$D
  for W in C4g8 C4; do
    timeout 300 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu --no-secondary --no-check > gpurun_out/la_${W}_d$D.json 2> gpurun_out/la_${W}_d$D.err
    python -c "
import json; d=json.load(open('gpurun_out/la_${W}_d$D.json')); e=d['e2e']; print('depth $D $W: ms/block %.4f e2e p50 %.4f ms p99 %.4f paced p50 %.4f'%(d['ms_per_block'], e['block_latency_ms_p50'], e['block_latency_ms_p99'], e.get('block_latency_paced_ms_p50',0)))"
  done
  timeout 120 python tools/trace_lookahead.py C4g8 2>&1 | tail -9
done
