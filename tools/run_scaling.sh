#!/bin/bash
# One N of the round's scaling table: bench.py (config 2), config 4 and config 3 under torchrun on N GPUs of one box.
#   gpurun --gpus N -- 'bash tools/run_scaling.sh N'
N=$1
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 "$@" 2>&1 | grep '^{' ; }
mkdir -p gpurun_out
echo "== bench N=$N"; run bench.py --gpus $N --steps 30 --warmup 5 --no-sweep | python -c "
import json,sys
d=json.loads(sys.stdin.read()); open('gpurun_out/r2_bench_${N}gpu.json','w').write(json.dumps(d)); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])"
echo "== config 4 N=$N"; run tools/bench_configs.py --config 4 | tee gpurun_out/r2_config4_${N}gpu.json | cut -c150-420
echo "== config 3 N=$N"; run tools/bench_configs.py --config 3 --steps 10 --warmup 3 | tee gpurun_out/r2_config3_${N}gpu.json | cut -c300-600
