python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/check_sharded_host.py 2>&1 | grep -E "step|sharded|Error|error" | tail -8
for sb in 4 1; do SAFCONV_SHARD_SUBBATCHES=$sb python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$sb bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2 subbatches=$sb', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'])"; done
