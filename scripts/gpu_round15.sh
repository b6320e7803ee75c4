#!/bin/bash
# bring-up of the channels-on-lanes GroupNorm GEMM: numeric probe, bf16 tests, bench at fuse 1 and 2 (all time-boxed)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r15}
echo "=== GNT probe"; GNT_ONLY=1 timeout 150 python tests/gpu_trace_gemm.py 2>&1 | tail -40 | tee gpurun_out/${T}_probe.log
echo "=== bf16 tests"; timeout 420 python -m pytest tests/test_gpu_bf16.py -x -q -m gpu --timeout 100 2>&1 | tail -6 | tee gpurun_out/${T}_t_bf16.log
echo "=== bench fuse 1"; timeout 200 python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline --fuse 1 --profile-ops > gpurun_out/${T}_bench_f1.json 2> gpurun_out/${T}_bench_f1.err; tail -1 gpurun_out/${T}_bench_f1.json | cut -c1-200; grep -E "core|mid_attn" gpurun_out/${T}_bench_f1.err | head -4
echo "=== bench fuse 2"; timeout 200 python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline --fuse 2 --profile-ops > gpurun_out/${T}_bench_f2.json 2> gpurun_out/${T}_bench_f2.err; tail -1 gpurun_out/${T}_bench_f2.json | cut -c1-200; head -1 gpurun_out/${T}_bench_f2.err; grep -E "downs.0.0|downs.0.1" gpurun_out/${T}_bench_f2.err | head -4; tail -3 gpurun_out/${T}_bench_f2.err | cut -c1-300
echo done
