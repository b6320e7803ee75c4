#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r17}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gemm_gnt -s 30 -c 3 -o gpurun_out/${T}_gnt_full -f python bench.py --steps 1 --warmup 1 --batch 4096 --timesteps 2 --no-cpu-baseline > gpurun_out/${T}_ncu_gnt.log 2>&1
tail -2 gpurun_out/${T}_ncu_gnt.log | cut -c1-200
ls -la gpurun_out/${T}_*.ncu-rep
echo done
