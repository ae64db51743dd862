python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/q_C4.json 2> gpurun_out/q_C4.err; python -c "
import json; d=json.load(open('gpurun_out/q_C4.json')); print(d['value'], d['ms_per_block'], d['e2e'])"
python bench.py --workload C4o --steps 5 --warmup 3 --no-cpu > gpurun_out/q_C4o.json 2> gpurun_out/q_C4o.err; python -c "
import json; d=json.load(open('gpurun_out/q_C4o.json')); print('C4o', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'], d['roofline']['frac'], d['e2e']['value'])" || tail -5 gpurun_out/q_C4o.err
for t in 128; do for o in 4 8; do SAFCONV_OFF_THREADS=$t SAFCONV_OFF_OPC=$o python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C5 thr $t opc $o', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'])"; done; done
SAFCONV_OFF_OPC=8 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C5 thr 256 opc 8', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'])"
