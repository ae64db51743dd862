#!/bin/bash
# GPU-box run of the filter producers: parity tests, smoke, the bench's producers block, timings of both render kernels,
# launch list, --set full captures of the windowed render kernel and the MagLS cluster kernel.
#   gpurun --timeout 600 -- 'bash tools/run_producers.sh r02'
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_producers.py -m gpu -q -s --tb=short > gpurun_out/${TAG}_producers_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/${TAG}_producers_tests.log
tail -4 gpurun_out/${TAG}_producers_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 120 python -c "
import argparse, json, torch, bench
print(json.dumps(bench.producers_block(argparse.Namespace(), torch.device('cuda:0'))))" > gpurun_out/${TAG}_bench_producers_block.json 2> gpurun_out/${TAG}_bench_producers_block.err
echo "bench block rc=$?"; cut -c1-1500 gpurun_out/${TAG}_bench_producers_block.json; tail -3 gpurun_out/${TAG}_bench_producers_block.err
timeout 120 python tools/producers_bench.py --sources 64 --cpu --out gpurun_out/${TAG}_producers_bench_s64.json > gpurun_out/${TAG}_producers_bench_s64.log 2>&1
echo "bench s64 (windows) rc=$?"; tail -1 gpurun_out/${TAG}_producers_bench_s64.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_producers.csv \
    python tools/producers_bench.py --sources 8 --max-time 1.0 > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"ims_window|magls_cluster" --launch-skip 1 -c 2 -o gpurun_out/${TAG}_producers_full -f \
    python tools/producers_bench.py --sources 8 --max-time 1.0 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
