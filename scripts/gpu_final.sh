#!/bin/bash
# What the driver runs at round end, in one call: full GPU test-suite (-x), smoke(), bench (both arms).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?" | tee gpurun_out/summary.txt
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench.err; echo "bench ref rc=$?" | tee -a gpurun_out/summary.txt
cut -c1-300 gpurun_out/bench_ref.json
timeout 600 python bench.py > gpurun_out/bench.json 2>> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json
timeout 300 python bench.py --chunks-per-gpu 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2>> gpurun_out/bench.err
cut -c1-200 gpurun_out/bench_b8.json
timeout 120 python scripts/tc_timing.py 1 > gpurun_out/tc_timing_b1.log 2>&1
timeout 120 python scripts/tc_timing.py 8 > gpurun_out/tc_timing_b8.log 2>&1
