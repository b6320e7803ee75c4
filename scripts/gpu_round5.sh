#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r5}
echo "=== trace"; timeout 300 python tests/gpu_trace_gemm.py 2>&1 | tail -22 | tee gpurun_out/${T}_trace.log
echo "=== bf16 tests"; timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/${T}_t_bf16.log
echo "=== fp32 tests"; timeout 900 python -m pytest tests/test_gpu_fp32.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/${T}_t_fp32.log
echo "=== bench 4096"; timeout 900 python bench.py --steps 1 --warmup 1 --batch 4096 --profile-ops > gpurun_out/${T}_bench_4096.json 2> gpurun_out/${T}_bench_4096.err; tail -1 gpurun_out/${T}_bench_4096.json | cut -c1-330; head -1 gpurun_out/${T}_bench_4096.err
for c in 2048 1480 1024; do echo "=== bench chunk $c"; timeout 900 python bench.py --steps 1 --warmup 1 --batch 4096 --chunk $c --no-cpu-baseline 2>/dev/null | cut -c1-330; done
echo "=== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --batch 4096 --timesteps 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
echo done
