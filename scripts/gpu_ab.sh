#!/bin/bash
# A/B of MTV_TC_MASK feature bits on the B=1 / B=8 bench (usage: gpu_ab.sh mask1 mask2 ...)
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "fixture or launch_modes" > gpurun_out/pytest_ab.log 2>&1; tail -2 gpurun_out/pytest_ab.log
for m in "$@"; do
  for b in 1 8; do
    timeout 200 env MTV_TC_MASK=$m python bench.py --steps 100 --chunks-per-gpu $b --no-cpu-baseline > gpurun_out/ab_${m}_b${b}.json 2>> gpurun_out/bench.err
    python -c "import json;d=json.load(open('gpurun_out/ab_${m}_b${b}.json'));print('MASK=$m B=$b', round(d['ms_per_step'],3), round(d['value'],1), d['kernel_families_us'])"
  done
done
