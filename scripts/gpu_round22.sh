#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r22}
echo "=== all gpu tests"; timeout 600 python -m pytest tests/ -x -q -m gpu --timeout 120 2>&1 | tail -4 | tee gpurun_out/${T}_t_all.log
echo "=== bench default (driver command)"; timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -1 gpurun_out/${T}_bench.json | cut -c1-260; tail -2 gpurun_out/${T}_bench.err | cut -c1-200
echo "=== bench PDL off"; DS_TC_PDL=0 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | cut -c1-200
echo done
