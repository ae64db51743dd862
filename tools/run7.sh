python -m pytest tests -m gpu -x -q -k "offline or smoke" 2>&1 | tail -2
python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu > gpurun_out/q_C5_w1.json 2> gpurun_out/q_C5_w1.err; python -c "
import json; d=json.load(open('gpurun_out/q_C5_w1.json')); print('wfft', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'], d['e2e']['value'])" || tail -5 gpurun_out/q_C5_w1.err
ncu --set full --clock-control none --import-source on -k regex:_w_kernel -s 4 -c 2 -f -o gpurun_out/prof_wfft python bench.py --workload C5s --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_wfft.log 2>&1; tail -2 gpurun_out/ncu_wfft.log
