python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
for w in UT C2 C1 C3; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r01_bench_${w}_n1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r01_bench_${w}_n1.json')); e=d['e2e']; print('$w', d['value'], d['ms_per_block'], d['roofline']['kernel_ms_per_block'], 'p50', e['block_latency_ms_p50'], 'paced', e.get('block_latency_paced_ms_p50'))"; done
