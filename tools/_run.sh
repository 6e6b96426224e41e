timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
CMD1="python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu-baseline --no-graph"
timeout 300 python bench.py --steps 30 --warmup 5 --no-sweep --no-cpu-baseline 2>&1 | grep '^{"metric' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'])"
$CMD1 > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv $CMD1 > gpurun_out/ncu1.log 2>&1
