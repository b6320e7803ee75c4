#!/bin/bash
# ncu --set full captures (with source) of the three kernels that dominate the step after the GEMMs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r13}
for K in k_linattn_mma k_layernorm512; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 10 -c 1 -o gpurun_out/${T}_${K}_full -f python bench.py --steps 1 --warmup 1 --batch 4096 --timesteps 2 --no-cpu-baseline > gpurun_out/${T}_ncu_${K}.log 2>&1
  tail -1 gpurun_out/${T}_ncu_${K}.log | cut -c1-200
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 62 -c 4 -o gpurun_out/${T}_gemm_full -f python bench.py --steps 1 --warmup 1 --batch 4096 --timesteps 2 --no-cpu-baseline > gpurun_out/${T}_ncu_gemm.log 2>&1
tail -1 gpurun_out/${T}_ncu_gemm.log | cut -c1-200
ls -la gpurun_out/${T}_*.ncu-rep
echo done
