# GPU tests + default bench (C4) ; everything important goes to gpurun_out/
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/q_C4.json 2> gpurun_out/q_C4.err; python -c "
import json; d=json.load(open('gpurun_out/q_C4.json')); print(d['value'], d['ms_per_block'], d['e2e'])"
SAFCONV_LOOKAHEAD=0 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/q_C4_nola.json 2> gpurun_out/q_C4_nola.err; python -c "
import json; d=json.load(open('gpurun_out/q_C4_nola.json')); print(d['value'], d['ms_per_block'], d['e2e'])"
