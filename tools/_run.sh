timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for t in ""; do echo "== tune '$t'"; timeout 300 python bench.py --steps 30 --warmup 5 --no-sweep --no-cpu-baseline --tune "$t" 2>&1 | grep '^{"metric' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'])"; done
