#!/bin/bash
# One GPU call that re-validates the final build of a round and refreshes the artefacts under profiles/:
#   gpurun --timeout 780 -- 'bash tools/final_check.sh'
# 1. pytest -m gpu   2. python bench.py (the driver's line)   3. smoke()   4. ncu launch list of the benchmark step
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/final_pytest.log
tail -5 gpurun_out/final_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/final_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["clocks"])
PY
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/final_smoke.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/final_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu-baseline --no-graph > gpurun_out/final_ncu.log 2>&1; echo "ncu rc $?"
