#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r26}
echo "=== all gpu tests (default)"; timeout 500 python -m pytest tests/ -x -q -m gpu --timeout 120 2>&1 | tail -3 | tee gpurun_out/${T}_t_all.log
echo "=== bf16 tests, DS_GNT_CLUSTER=2"; DS_GNT_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_bf16.py -x -q -m gpu --timeout 100 2>&1 | tail -3 | tee gpurun_out/${T}_t_bf16_c2.log
for c in 1 2 1 2; do echo "GNT_CLUSTER=$c"; DS_GNT_CLUSTER=$c timeout 150 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'])"; done
echo done
