# ncu evidence for the round (B200_PROFILING.md recipe): plain run first, then launch list, then --set full of the top kernels
set -x
R=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_c4.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_c4.csv $CMD > gpurun_out/ncu_l_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mac_kernel -s 3 -c 2 -f -o gpurun_out/${R}_mac_c4_full $CMD > gpurun_out/ncu_f_c4.log 2>&1
CMD5="python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu"
$CMD5 > gpurun_out/plain_c5.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${R}_launches_c5.csv $CMD5 > gpurun_out/ncu_l_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:offline_ -s 6 -c 4 -f -o gpurun_out/${R}_offline_c5_full $CMD5 > gpurun_out/ncu_f_c5.log 2>&1
ls -la gpurun_out/${R}_*
