python -m pytest tests -m gpu -x -q -k "offline or smoke" -s 2>&1 | grep -v "^$" | tail -25
for k in f16 tf32; do SAFCONV_OFF_KIND=$k python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu > gpurun_out/q_C5_$k.json 2> gpurun_out/q_C5_$k.err; python -c "
import json; d=json.load(open('gpurun_out/q_C5_$k.json')); print('$k', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'], d['roofline']['frac'], d['e2e']['value'])" || tail -5 gpurun_out/q_C5_$k.err; done
SAFCONV_OFF_FPC=4 python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('f16 fpc4', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'])"
python tools/offline_check.py 2>&1 | tail -4
