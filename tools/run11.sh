python -m pytest tests -m gpu -x -q -k "offline or smoke" -s 2>&1 | grep -E "passed|failed|Error|offline" | tail -14
for nt in 1 0; do SAFCONV_OFF_NT=$nt python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C5 NT=$nt', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'], d['roofline']['frac'], d['e2e']['value'])"; done
SAFCONV_OFF_KIND=tf32 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C5 tf32 NT', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'])"
python bench.py --workload C4o --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C4o', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'], d['roofline']['issued_frac'])"
