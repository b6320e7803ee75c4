#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r16}
echo "=== GNT probe"; GNT_ONLY=1 timeout 150 python tests/gpu_trace_gemm.py 2>&1 | grep -vE "^   prod_wait" | tail -24 | tee gpurun_out/${T}_probe.log
echo "=== bf16 tests"; timeout 420 python -m pytest tests/test_gpu_bf16.py -x -q -m gpu --timeout 100 2>&1 | tail -4 | tee gpurun_out/${T}_t_bf16.log
echo "=== bench default"; timeout 200 python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline --profile-ops > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -1 gpurun_out/${T}_bench.json | cut -c1-200; head -1 gpurun_out/${T}_bench.err; grep -E "downs.0.0|downs.0.1|downs.0.4|mid_attn" gpurun_out/${T}_bench.err | head -12
echo done
