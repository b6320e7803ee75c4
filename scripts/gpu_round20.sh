#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r20}
echo "=== all gpu tests"; timeout 600 python -m pytest tests/ -x -q -m gpu --timeout 120 2>&1 | tail -4 | tee gpurun_out/${T}_t_all.log
echo "=== ncu launch list + dram bytes"; timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 250 -c 260 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --batch 4096 --timesteps 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; tail -1 gpurun_out/${T}_ncu_bench.log | cut -c1-100
echo "=== bench"; timeout 300 python bench.py --steps 2 --warmup 1 --batch 4096 --no-cpu-baseline --profile-ops > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -1 gpurun_out/${T}_bench.json | cut -c1-200; head -1 gpurun_out/${T}_bench.err
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo done
