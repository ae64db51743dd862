#!/bin/bash
# final GPU-box check of a round: every GPU test, smoke, producers timings + ncu, then the default bench line
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests_final.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${TAG}_gputests_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 120 python tools/producers_bench.py --sources 64 --cpu --out gpurun_out/${TAG}_producers_bench_s64.json > gpurun_out/${TAG}_producers_bench_s64.log 2>&1
echo "producers bench rc=$?"; tail -1 gpurun_out/${TAG}_producers_bench_s64.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_producers.csv \
    python tools/producers_bench.py --sources 8 --max-time 1.0 > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"ims_render" --launch-skip 1 -c 1 -o gpurun_out/${TAG}_ims_render_full -f \
    python tools/producers_bench.py --sources 8 --max-time 1.0 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
S=$(date +%s); timeout 240 python bench.py > gpurun_out/${TAG}_bench_c4_n1_final.json 2> gpurun_out/${TAG}_bench_final.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_c4_n1_final.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['value'], d['roofline']['frac']); print(json.dumps(d['secondary'].get('producers'))[:1200])"
