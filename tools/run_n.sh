N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r01_bench_c4_n$N.json 2> gpurun_out/r01_bench_c4_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/r01_bench_c4_n$N.json')); print('C4 N=$N', d['value'], d['ms_per_block'], d['roofline']['achieved'], d['e2e']['value'])" || tail -5 gpurun_out/r01_bench_c4_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload C5 --steps 10 --warmup 3 > gpurun_out/r01_bench_c5_n$N.json 2> gpurun_out/r01_bench_c5_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/r01_bench_c5_n$N.json')); print('C5 N=$N', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_render'], d['e2e']['value'])" || tail -5 gpurun_out/r01_bench_c5_n$N.err
