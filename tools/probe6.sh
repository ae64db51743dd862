for W in C1 C2; do
  SAFCONV_KSTAMPS=1 timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-secondary --no-check > gpurun_out/ks_${W}.json 2> gpurun_out/ks_${W}.err
  grep "small_cluster phases" gpurun_out/ks_${W}.err | tail -2
done
