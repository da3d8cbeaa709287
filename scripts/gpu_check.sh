#!/bin/bash
# One gpurun call: build check, GPU parity tests, smoke, bench, launch list.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1; echo "build rc=$?" | tee -a gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -rA --durations=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -n 60 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "${1:-}" != "quick" ]; then
  timeout 600 python bench.py --chunks-per-gpu 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2>> gpurun_out/bench.err
  cat gpurun_out/bench_b8.json
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_(conv|attn|gn_|splitk|linear|temb|pack|ddim|apply|tc_)" -c 3000 --csv --log-file gpurun_out/launches.csv \
      env MTV_NO_GRAPH=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
fi
