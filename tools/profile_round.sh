# ncu evidence for the round (B200_PROFILING.md recipe): plain run first, then launch list, then --set full of the top kernels.
#   bash tools/profile_round.sh r02        (on the GPU box, from the repo root; outputs under gpurun_out/)
# The SASS summary (tools/sass_summary.py) is generated on the build host: python tools/sass_summary.py > profiles/rNN_sass_summary.txt
set -x
R=${1:-r02}
FAST="--no-cpu --no-secondary --no-check --e2e-blocks 200"
CMD="python bench.py --steps 2 --warmup 3 $FAST"
$CMD > gpurun_out/plain_c4.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_c4.csv $CMD > gpurun_out/ncu_l_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mac_kernel -s 3 -c 2 -f -o gpurun_out/${R}_mac_c4_full $CMD > gpurun_out/ncu_f_c4.log 2>&1
CMD5="python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu --no-check"
$CMD5 > gpurun_out/plain_c5.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${R}_launches_c5.csv $CMD5 > gpurun_out/ncu_l_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:offline_ -s 6 -c 4 -f -o gpurun_out/${R}_offline_c5_full $CMD5 > gpurun_out/ncu_f_c5.log 2>&1
# latency configs: the fused one-launch kernels (C2 small_fused_kernel, C3 multi_fused_kernel + the batched multiConv kernels)
for W in C2 C3; do
  CMDW="python bench.py --workload $W --steps 2 --warmup 3 $FAST"
  $CMDW > gpurun_out/plain_$W.log 2>&1 || continue
  ncu --set full --clock-control none --import-source on -k regex:'small_|multi_' -s 8 -c 6 -f -o gpurun_out/${R}_${W}_full $CMDW > gpurun_out/ncu_f_$W.log 2>&1
done
for f in gpurun_out/${R}_*_full.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
done
# gpurun copies at most 64 MiB back: keep the raw CSVs of all captures, the reports of the two headline kernels only
rm -f gpurun_out/${R}_C2_full.ncu-rep gpurun_out/${R}_C3_full.ncu-rep
ls -la gpurun_out/${R}_*
