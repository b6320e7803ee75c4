#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r19}
echo "=== GNT probe"; GNT_ONLY=1 timeout 200 python tests/gpu_trace_gemm.py 2>&1 | grep -vE "^   prod_wait" | tail -70 | tee gpurun_out/${T}_probe.log
echo "=== bf16 tests"; timeout 420 python -m pytest tests/test_gpu_bf16.py -x -q -m gpu --timeout 100 2>&1 | tail -4 | tee gpurun_out/${T}_t_bf16.log
for F in 2 3; do
echo "=== bench fuse $F"; timeout 200 python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline --fuse $F --profile-ops > gpurun_out/${T}_bench_f$F.json 2> gpurun_out/${T}_bench_f$F.err; tail -1 gpurun_out/${T}_bench_f$F.json | cut -c1-200; head -1 gpurun_out/${T}_bench_f$F.err
done
echo done
