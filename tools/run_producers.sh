#!/bin/bash
# GPU-box run of the filter producers: parity tests, timings, launch list, one --set full capture of the render kernel.
#   gpurun --timeout 600 -- 'bash tools/run_producers.sh r02'
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_producers.py -m gpu -q -s --tb=short > gpurun_out/${TAG}_producers_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/${TAG}_producers_tests.log
tail -5 gpurun_out/${TAG}_producers_tests.log
timeout 120 python tools/producers_bench.py --sources 4 --cpu --out gpurun_out/${TAG}_producers_bench_s4.json > gpurun_out/${TAG}_producers_bench_s4.log 2>&1
echo "bench s4 rc=$?"; tail -2 gpurun_out/${TAG}_producers_bench_s4.log
timeout 150 python tools/producers_bench.py --sources 64 --out gpurun_out/${TAG}_producers_bench_s64.json > gpurun_out/${TAG}_producers_bench_s64.log 2>&1
echo "bench s64 rc=$?"; tail -2 gpurun_out/${TAG}_producers_bench_s64.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_producers.csv \
    python tools/producers_bench.py --sources 8 --max-time 1.0 > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:ims_render -c 1 -o gpurun_out/${TAG}_ims_render_full -f \
    python tools/producers_bench.py --sources 8 --max-time 1.0 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
