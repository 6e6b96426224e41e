python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'], d['roofline']['kernels_ms'])"
python tools/time_block.py
