#!/bin/bash
# quick GPU check: parity tests + B=1 / B=8 bench + in-kernel timing
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1; echo "build rc=$?" | tee gpurun_out/summary.txt
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --cpu-baseline-steps 2 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
timeout 300 python bench.py --chunks-per-gpu 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2>> gpurun_out/bench.err
timeout 120 python scripts/tc_timing.py 1 > gpurun_out/tc_timing_b1.log 2>&1; timeout 120 python scripts/tc_timing.py 8 > gpurun_out/tc_timing_b8.log 2>&1
