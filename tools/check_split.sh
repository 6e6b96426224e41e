#!/bin/bash
# GPU check of the time-parallel forward scan: parity tests, then A/B of the serial walk (knob 6 = 1) against the split at
# config 5's long points, then the full bench line.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/split_pytest.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/split_pytest.log
tail -4 gpurun_out/split_pytest.log
python tools/ab_split.py 2>&1 | grep -v Warning | tee gpurun_out/split_ab.log
python bench.py --steps 30 --warmup 5 > gpurun_out/split_bench.json 2> gpurun_out/split_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/split_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"])
for r in d["scan_sweep"]:
    if r["L"] >= 4096: print(r)
PY
