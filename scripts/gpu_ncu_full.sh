#!/bin/bash
# one ncu --set full capture of the fused GEMM+GroupNorm kernel and of the plain GEMM kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 60 -c 6 -o gpurun_out/${1:-r2}_gemm_full -f python bench.py --steps 1 --warmup 1 --batch 4096 --timesteps 2 --no-cpu-baseline > gpurun_out/${1:-r2}_ncu_full.log 2>&1
tail -2 gpurun_out/${1:-r2}_ncu_full.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
