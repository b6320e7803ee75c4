#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r24}
echo "=== all gpu tests"; timeout 600 python -m pytest tests/ -x -q -m gpu --timeout 120 2>&1 | tail -3 | tee gpurun_out/${T}_t_all.log
echo "=== bench default (driver command)"; timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -1 gpurun_out/${T}_bench.json | cut -c1-260
for c in "1 0" "0 1"; do set -- $c; echo "TC_PDL=$1 PW_PDL=$2"; DS_TC_PDL=$1 DS_PW_PDL=$2 timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'])"; done
echo done
