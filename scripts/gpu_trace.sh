#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 1 2 4; do
  echo "=== trace cluster $c"; DS_TC_CLUSTER=$c timeout 300 python tests/gpu_trace_gemm.py 2>&1 | tail -25 | tee gpurun_out/trace_c$c.log
  echo "=== gemm unit tests cluster $c"; DS_TC_CLUSTER=$c timeout 600 python -m pytest tests/test_gpu_bf16.py -q -m gpu -k "gemm_matches_torch and tcgen05" 2>&1 | tail -4
done
echo "=== full bf16 tests with cluster 2"; DS_TC_CLUSTER=2 timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu 2>&1 | tail -6
echo "=== bench cluster 2"; DS_TC_CLUSTER=2 timeout 900 python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline 2>/dev/null | cut -c1-420
echo "=== bench cluster 4"; DS_TC_CLUSTER=4 timeout 900 python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline 2>/dev/null | cut -c1-420
echo "=== api tests"; timeout 900 python -m pytest tests/test_gpu_api.py -q -m gpu 2>&1 | tail -15
echo done
