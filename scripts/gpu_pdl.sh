#!/bin/bash
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for m in 1 5 9 13; do
  for b in 1; do
    timeout 200 env MTV_PDL=$m python bench.py --steps 100 --chunks-per-gpu $b --no-cpu-baseline > gpurun_out/pdl_${m}_b${b}.json 2>> gpurun_out/bench.err
    python -c "import json;d=json.load(open('gpurun_out/pdl_${m}_b${b}.json'));print('MTV_PDL=$m B=$b', round(d['ms_per_step'],3), round(d['value'],1))"
  done
done
