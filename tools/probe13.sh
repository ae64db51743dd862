export SAFCONV_TAIL_RESERVE_SMS=32
for D in 1 2; do for HK in 0 1; do
  export SAFCONV_LA_DEPTH=$D SAFCONV_HEAD_IN_K3=$HK
  for W in C4g8 C4; do
  timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-secondary --no-check --e2e-blocks 1000 > gpurun_out/m_${W}_${D}_$HK.json 2> gpurun_out/m_${W}_${D}_$HK.err
  python -c "
import json; d=json.load(open('gpurun_out/m_${W}_${D}_$HK.json')); e=d['e2e']; print('depth $D hk $HK $W: e2e %.4g p50 %.4f ms p99 %.4f paced p50 %.4f p99 %.4f'%(e['value'], e['block_latency_ms_p50'], e['block_latency_ms_p99'], e.get('block_latency_paced_ms_p50',0), e.get('block_latency_paced_ms_p99',0)))"
  done
done; done
