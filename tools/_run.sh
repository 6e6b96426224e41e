python -m pytest tests -m gpu -q -x 2>&1 | tail -30
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -4 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['cpu_baseline'])
for r in d['roofline_more']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k not in ('note',)})
for r in d['scan_sweep']: print(r)
PY
