#!/bin/bash
# A/B timing of experiment builds of the library (build.py --variant NAME -D...) against the shipped build, on one GPU:
#   gpurun -- 'bash tools/ab_variants.sh nopipe late poly2'
# Prints, per build, the bench step (config 2, graph replay), the scan kernels' mean launch time inside it, and the
# kernel-only scan times at config-5 size (2048 x 256) in bf16 and fp32.
mkdir -p gpurun_out
V=robust-audio-deepfake-evolution_b200/_variants
for name in main "$@"; do
  if [ "$name" = main ]; then unset BIMAMBA_LIB; else export BIMAMBA_LIB=$PWD/$V/libbimamba_sm100_$name.so; fi
  echo "== build $name"
  python bench.py --steps 30 --warmup 5 --no-sweep --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({'ms_per_step':round(d['ms_per_step'],4),'e2e_ms':round(d['e2e']['ms_per_step'],4),'scan_bwd_ms':round(d['roofline'].get('avg_launch_ms',0),4),'scan_fwd_ms':round(d['roofline_more'][0]['avg_launch_ms'],4),'gemm_in_proj_ms':d['roofline_more'][1].get('avg_launch_ms'),'kernels_ms':d['roofline'].get('kernels_ms')})"
  for dt in bf16 f32; do python tools/time_scan.py --batch 2048 --L 256 --dtype $dt 2>/dev/null | tr '\n' ' '; echo; done
done
