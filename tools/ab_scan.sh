#!/bin/bash
# Kernel-only A/B of experiment builds (build.py --variant NAME -D...) for the scan kernels on one GPU:
#   gpurun -- 'bash tools/ab_scan.sh main branch half split2:6=2'
# name[:knob=value,...]; per build: one block's kernels at config 2 (eager, L2 flushed) and the op at 2048 x 256 (bf16, fp32).
V=robust-audio-deepfake-evolution_b200/_variants
for spec in "$@"; do
  name=${spec%%:*}; tune=""; [ "$spec" != "$name" ] && tune=${spec#*:}
  if [ "$name" = main ]; then unset BIMAMBA_LIB; else export BIMAMBA_LIB=$PWD/$V/libbimamba_sm100_$name.so; fi
  echo "== build $name tune '$tune'"
  python tools/time_block.py --tune "$tune" 2>&1 | grep -E "scan|sum of" | tr '\n' ' '; echo
  for dt in bf16 f32; do python tools/time_scan.py --batch 2048 --L 256 --dtype $dt --tune "$tune" 2>&1 | tail -2 | tr '\n' ' '; echo; done
done
