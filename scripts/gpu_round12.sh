#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r12}
echo "=== all gpu tests"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${T}_t_all.log
echo "=== bench 4096"; timeout 900 python bench.py --steps 2 --warmup 1 --batch 4096 --profile-ops > gpurun_out/${T}_bench_4096.json 2> gpurun_out/${T}_bench_4096.err; tail -1 gpurun_out/${T}_bench_4096.json | cut -c1-330; head -1 gpurun_out/${T}_bench_4096.err; grep -E "core|prenorm|mid_attn" gpurun_out/${T}_bench_4096.err | head -4
echo done
