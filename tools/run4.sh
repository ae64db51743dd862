python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for g in 0 4 16; do SAFCONV_MULTI_G=$g python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C3 G=$g', d['value'], d['ms_per_block'], d['roofline']['kernel_ms_per_block'], d['e2e']['block_latency_ms_p50'], d['e2e'].get('block_latency_paced_ms_p50'))"; done
