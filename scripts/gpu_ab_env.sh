#!/bin/bash
# generic A/B: each argument is a quoted env assignment list, e.g. "MTV_PDL=5" "MTV_PDL=5 MTV_TC_MASK=0xfdf"
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
i=0
for cfg in "$@"; do
  for b in ${BATCHES:-1 8}; do
    timeout 200 env $cfg python bench.py --steps 100 --chunks-per-gpu $b --no-cpu-baseline > gpurun_out/abenv_${i}_b${b}.json 2>> gpurun_out/bench.err
    python -c "import json;d=json.load(open('gpurun_out/abenv_${i}_b${b}.json'));print('[$cfg] B=$b', round(d['ms_per_step'],3), round(d['value'],1), d['gpu_launches']//d['steps'], d['kernel_families_us'])"
  done
  i=$((i+1))
done
