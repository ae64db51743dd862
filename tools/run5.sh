python bench.py --workload C3 --steps 3 --warmup 2 --no-cpu > gpurun_out/plain_c3.log 2>&1 || tail -3 gpurun_out/plain_c3.log
ncu --set full --clock-control none --import-source on -k regex:multi --launch-skip 6 --launch-count 3 -o gpurun_out/prof_multi_c3_b -f python bench.py --workload C3 --steps 3 --warmup 2 --no-cpu > gpurun_out/ncu_c3.log 2>&1
tail -3 gpurun_out/ncu_c3.log
