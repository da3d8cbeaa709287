#!/bin/bash
# A/B of environment settings: bash scripts/gpu_ab_env.sh TAG "ENV1=.. ENV2=.." "ENV=.." ...   ("-" = defaults)
set -u
TAG=$1; shift
O=gpurun_out/$TAG; mkdir -p $O
python __graft_entry__.py > $O/build.log 2>&1 || { echo build failed; tail $O/build.log; exit 1; }
i=0
for cfg in "$@"; do
  i=$((i+1))
  [ "$cfg" = "-" ] && cfg=""
  for B in ${AB_BATCHES:-1 8}; do
    env $cfg timeout 300 python bench.py --no-cpu-baseline --steps ${AB_STEPS:-60} --chunks-per-gpu $B > $O/bench_${i}_b$B.json 2>> $O/bench.err
    python - <<PY
import json
try:
    d = json.loads(open("$O/bench_${i}_b$B.json").read().strip().splitlines()[-1])
    print("[$cfg] B=$B ms/step", round(d["ms_per_step"], 4), "value", round(d["value"], 1), "launches", d["gpu_launches"] / d["steps"], d["kernel_families_us"])
except Exception as e:
    print("[$cfg] B=$B unreadable", e)
PY
  done
done
tail -n 3 $O/bench.err
