#!/bin/bash
# robustness pass: the driver's exact test command, the reference arm, default bench, ncu of the pointwise kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r10}
echo "=== pytest -m gpu (driver command)"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${T}_t_all.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-600
echo "=== bench default"; timeout 1500 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -1 gpurun_out/${T}_bench_default.json | cut -c1-1500
echo "=== ncu full pointwise"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_linattn|k_layernorm512|k_softattn" -s 20 -c 4 -o gpurun_out/${T}_pw_full -f python bench.py --steps 1 --warmup 1 --batch 4096 --timesteps 2 --no-cpu-baseline > gpurun_out/${T}_ncu_pw.log 2>&1
echo done
