#!/bin/bash
# chain kernel bring-up: equality vs stand-alone launches, GPU parity suite, A/B bench
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1; echo "build rc=$?" | tee gpurun_out/summary.txt
timeout 300 python scripts/chain_check.py tiny 2 > gpurun_out/chain_check.log 2>&1; echo "chain_check tiny rc=$?" | tee -a gpurun_out/summary.txt
timeout 300 python scripts/chain_check.py base 1 3 8 >> gpurun_out/chain_check.log 2>&1; echo "chain_check base rc=$?" | tee -a gpurun_out/summary.txt
tail -8 gpurun_out/chain_check.log
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
for m in 0x1fff 0xfff; do
  for b in 1 8; do
    timeout 200 env MTV_TC_MASK=$m python bench.py --steps 100 --chunks-per-gpu $b --no-cpu-baseline > gpurun_out/ab_${m}_b${b}.json 2>> gpurun_out/bench.err
    python -c "import json;d=json.load(open('gpurun_out/ab_${m}_b${b}.json'));print('MASK=$m B=$b', round(d['ms_per_step'],3), round(d['value'],1), d['gpu_launches'], d['kernel_families_us'])"
  done
done
timeout 120 python scripts/chain_timing.py 1 > gpurun_out/chain_timing_b1.log 2>&1; tail -5 gpurun_out/chain_timing_b1.log
