set -x
CMD1="python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu-baseline --no-graph"
$CMD1 > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv $CMD1 > gpurun_out/ncu1.log 2>&1
CMD2="python tools/prof_block.py --iters 2"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"scan_bwd_lane|scan_fwd_warp|gemm_nt_kernel|gemm_tn_kernel|conv_bwd_seg" -s 30 -c 24 -o gpurun_out/r2_kernels $CMD2 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | tail -8
