timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --workload C4g8 --steps 5 --warmup 3 --no-cpu --no-secondary > gpurun_out/c4g8.json 2> gpurun_out/c4g8.err
python -c "
import json; d=json.load(open('gpurun_out/c4g8.json')); e=d['e2e']; print('C4g8 value %.4g ms/block %.4f e2e %.4g p50 %.4f ms p99 %.4f paced %.4f parity %s'%(d['value'], d['ms_per_block'], e['value'], e['block_latency_ms_p50'], e['block_latency_ms_p99'], e.get('block_latency_paced_ms_p50',0), d.get('parity_rel_l2')))"
timeout 120 python tools/trace_lookahead.py C4g8 2>&1 | tail -16
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-secondary > gpurun_out/c4_quick.json 2> gpurun_out/c4_quick.err
python -c "
import json; d=json.load(open('gpurun_out/c4_quick.json')); e=d['e2e']; print('C4 value %.4g ms/block %.4f e2e %.4g p50 %.4f ms paced %.4f parity %s'%(d['value'], d['ms_per_block'], e['value'], e['block_latency_ms_p50'], e.get('block_latency_paced_ms_p50',0), d.get('parity_rel_l2')))"
