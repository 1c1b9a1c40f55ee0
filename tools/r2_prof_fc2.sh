#!/bin/bash
# ncu: the fc2 GEMM with the plain reduce-add epilogue vs the residual + statistics epilogue (serial staging), same shape
VFM_RS_MODE=${1:-2} timeout 120 python tools/bench_kernels.py --crops 36 --only fc2 --iters 1 --warmup 1 > gpurun_out/prof_fc2_plain.log 2>&1 || { tail -5 gpurun_out/prof_fc2_plain.log; exit 1; }
VFM_RS_MODE=${1:-2} timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -c 4 -o gpurun_out/r2_prof_fc2_stats -f \
  python tools/bench_kernels.py --crops 36 --only fc2 --iters 1 --warmup 1 > gpurun_out/prof_fc2_ncu.log 2>&1
ls -la gpurun_out/r2_prof_fc2_stats.ncu-rep; tail -3 gpurun_out/prof_fc2_plain.log
