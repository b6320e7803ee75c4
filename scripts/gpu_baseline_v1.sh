#!/bin/bash
# v1 (unfused) baseline numbers + ncu launch list, for profiles/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 1 --warmup 1 --batch 4096 --profile-ops > gpurun_out/v1_bench_4096.json 2> gpurun_out/v1_bench_4096.err
tail -1 gpurun_out/v1_bench_4096.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/v1_launches.csv python bench.py --steps 1 --warmup 1 --batch 4096 --timesteps 3 --no-cpu-baseline > gpurun_out/v1_ncu_bench.log 2>&1
tail -2 gpurun_out/v1_ncu_bench.log
wc -l gpurun_out/v1_launches.csv
echo done
