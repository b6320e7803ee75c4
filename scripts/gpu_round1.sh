#!/bin/bash
# First GPU bring-up: fp32 parity, tcgen05 probe, bf16 tests, a short bench.  Logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== fp32 parity tests"; timeout 900 python -m pytest tests/test_gpu_fp32.py -q -m gpu 2>&1 | tail -60 | tee gpurun_out/t_fp32.log
echo "=== tcgen05 probe (default descriptors)"; timeout 300 python tests/gpu_probe_tc.py default 2>&1 | tail -20 | tee gpurun_out/probe_default.log
if grep -q "FAILED\|max_err=[1-9]\|max_err=0\.[1-9]\|nan" gpurun_out/probe_default.log; then
  echo "=== probe variant: no version bit"; DS_TC_DESC_HI=0x4000004000010000 timeout 300 python tests/gpu_probe_tc.py nover 2>&1 | tail -14 | tee gpurun_out/probe_nover.log
  echo "=== probe variant: lbo 0"; DS_TC_DESC_HI=0x4000404000000000 timeout 300 python tests/gpu_probe_tc.py lbo0 2>&1 | tail -14 | tee gpurun_out/probe_lbo0.log
fi
echo "=== bf16 tests (simt backend first)"; timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -k "simt" 2>&1 | tail -30 | tee gpurun_out/t_bf16_simt.log
echo "=== bf16 tests (tcgen05)"; timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -k "not simt" 2>&1 | tail -40 | tee gpurun_out/t_bf16_tc.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench (short: 1024 scenes, T=1000)"; timeout 900 python bench.py --steps 1 --warmup 1 --batch 1024 --profile-ops > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; tail -3 gpurun_out/bench_1024.json; tail -150 gpurun_out/bench_1024.err | head -170
echo "=== bench fp32 (short: 256 scenes, T=100)"; timeout 600 python bench.py --steps 1 --warmup 1 --batch 256 --timesteps 100 --precision fp32 --no-cpu-baseline > gpurun_out/bench_fp32.json 2>&1; tail -2 gpurun_out/bench_fp32.json
echo done
