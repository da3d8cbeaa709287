#!/bin/bash
# A/B of programmatic-dependent-launch class masks (MTV_PDL) on the default (apply + TMA) path
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for pdl in ${PDLS:-0 1 3 5 9 13 15 31}; do
  for b in 1 8; do
    timeout 200 env MTV_TC_MASK=${MASK:-0xfff} MTV_PDL=$pdl python bench.py --steps 100 --chunks-per-gpu $b --no-cpu-baseline > gpurun_out/pdl_${pdl}_b${b}.json 2>> gpurun_out/bench.err
    python -c "import json;d=json.load(open('gpurun_out/pdl_${pdl}_b${b}.json'));print('PDL=$pdl B=$b', round(d['ms_per_step'],3), round(d['value'],1))"
  done
done
