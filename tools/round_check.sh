# what the driver runs at round end: gpu tests, smoke, both bench arms (timed)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
S=$(date +%s); python bench.py --impl reference > gpurun_out/${R:-r02}_bench_c4_reference.json 2> gpurun_out/ref.err; echo "reference arm wall $(( $(date +%s) - S )) s"; cat gpurun_out/${R:-r02}_bench_c4_reference.json | cut -c1-600
S=$(date +%s); python bench.py > gpurun_out/${R:-r02}_bench_c4_n1.json 2> gpurun_out/own.err; echo "own arm wall $(( $(date +%s) - S )) s"; python -c "
import json; d=json.load(open('gpurun_out/${R:-r02}_bench_c4_n1.json')); print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','clocks','cpu_baseline')}); print(d['e2e']); print(d['roofline'])"
