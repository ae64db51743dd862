for D in 1 2; do for HK in 1 0; do
  export SAFCONV_LA_DEPTH=$D SAFCONV_HEAD_IN_K3=$HK
  for W in C4 C4g8; do
  timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-secondary --no-check --e2e-blocks 1000 > gpurun_out/la_${W}_d${D}_h$HK.json 2> gpurun_out/la_${W}_d${D}_h$HK.err
  python -c "
import json; d=json.load(open('gpurun_out/la_${W}_d${D}_h$HK.json')); e=d['e2e']; print('depth $D head_in_k3 $HK $W: e2e p50 %.4f ms p99 %.4f paced p50 %.4f p99 %.4f'%(e['block_latency_ms_p50'], e['block_latency_ms_p99'], e.get('block_latency_paced_ms_p50',0), e.get('block_latency_paced_ms_p99',0)))"
  done
done; done
