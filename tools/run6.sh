python -m pytest tests -m gpu -x -q -k "offline or smoke" -s 2>&1 | grep -v "^$" | tail -16
for wf in 1 0; do SAFCONV_OFF_WFFT=$wf python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu > gpurun_out/q_C5_w$wf.json 2> gpurun_out/q_C5_w$wf.err; python -c "
import json; d=json.load(open('gpurun_out/q_C5_w$wf.json')); print('wfft=$wf', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'], d['e2e']['value'])" || tail -5 gpurun_out/q_C5_w$wf.err; done
