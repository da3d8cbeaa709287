#!/bin/bash
# direct-A bring-up: GPU parity suite in direct mode, A/B bench against the apply+TMA path
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1; echo "build rc=$?" | tee gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log
for m in ${MASKS:-0x2fff 0xfff}; do
  for b in 1 8; do
    timeout 200 env MTV_TC_MASK=$m python bench.py --steps 100 --chunks-per-gpu $b --no-cpu-baseline > gpurun_out/ab_${m}_b${b}.json 2>> gpurun_out/bench.err
    python -c "import json;d=json.load(open('gpurun_out/ab_${m}_b${b}.json'));print('MASK=$m B=$b', round(d['ms_per_step'],3), round(d['value'],1), d['gpu_launches'], d['kernel_families_us'])"
  done
done
tail -3 gpurun_out/bench.err
