run() {
echo "== $*"
env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-sweep --comm overlap 2>&1 | grep '^{"metric' | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('FAILED', t[-600:])"
}
run NCCL_MAX_NCHANNELS=2
run NCCL_MAX_NCHANNELS=4
run NCCL_PROTO=LL
run NCCL_MAX_NCHANNELS=4 NCCL_PROTO=LL128
run NCCL_MAX_CTAS=4
python bench.py --steps 30 --warmup 5 --no-sweep --no-cpu-baseline 2>&1 | grep '^{"metric' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('single', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'])"
