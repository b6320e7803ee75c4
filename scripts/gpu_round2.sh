#!/bin/bash
# fused GroupNorm epilogue bring-up: tests, bench, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== bf16 tests"; timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x 2>&1 | tail -30 | tee gpurun_out/r2_t_bf16.log
echo "=== fp32 tests"; timeout 900 python -m pytest tests/test_gpu_fp32.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_t_fp32.log
echo "=== bench 4096 fused"; timeout 900 python bench.py --steps 1 --warmup 1 --batch 4096 --profile-ops > gpurun_out/r2_bench_4096.json 2> gpurun_out/r2_bench_4096.err; tail -1 gpurun_out/r2_bench_4096.json; head -3 gpurun_out/r2_bench_4096.err
echo "=== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --batch 4096 --timesteps 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1; tail -1 gpurun_out/r2_ncu_bench.log | cut -c1-200
echo done
