#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r9}
echo "=== fp32 tests"; timeout 900 python -m pytest tests/test_gpu_fp32.py -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${T}_t_fp32.log
echo "=== bf16 tests"; timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/${T}_t_bf16.log
echo "=== trace"; timeout 300 python tests/gpu_trace_gemm.py 2>&1 | grep -A1 "GN\|plain 512x512 " | cut -c1-200 | tee gpurun_out/${T}_trace.log
echo "=== bench 4096"; timeout 900 python bench.py --steps 1 --warmup 1 --batch 4096 --profile-ops > gpurun_out/${T}_bench_4096.json 2> gpurun_out/${T}_bench_4096.err; tail -1 gpurun_out/${T}_bench_4096.json | cut -c1-330; head -1 gpurun_out/${T}_bench_4096.err
echo "=== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --batch 4096 --timesteps 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
echo done
