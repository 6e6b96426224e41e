python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python tools/bench_configs.py --config 4 2>&1 | grep "^{" | cut -c150-500
python __graft_entry__.py smoke 2>&1 | tail -1
