#!/bin/bash
# GPU run focused on the tensor-core path: bisect script first (cheap, isolates op classes), then tests, then bench.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1; echo "build rc=$?" | tee gpurun_out/summary.txt
timeout 240 python scripts/tc_bisect.py base new > gpurun_out/bisect_base.log 2>&1; echo "bisect base rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bisect_base.log | tail -12
timeout 240 python scripts/tc_bisect.py tiny new > gpurun_out/bisect_tiny.log 2>&1; echo "bisect tiny rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bisect_tiny.log | tail -12
timeout 600 python -m pytest tests -q -m gpu -rA --durations=5 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
grep -E "passed|failed|PASSED|FAILED|rel-L2|stage errors|Error" gpurun_out/pytest_gpu.log | tail -40
timeout 300 python bench.py --cpu-baseline-steps 2 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 300 python bench.py --chunks-per-gpu 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_b8.json
timeout 120 python scripts/tc_timing.py 1 > gpurun_out/tc_timing_b1.log 2>&1; tail -30 gpurun_out/tc_timing_b1.log | head -5
timeout 200 env MTV_PDL=1 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_pdl.json | cut -c1-300
timeout 200 env MTV_PDL=1 python bench.py --chunks-per-gpu 8 --no-cpu-baseline > gpurun_out/bench_b8_pdl.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_b8_pdl.json | cut -c1-300
