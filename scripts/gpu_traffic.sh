#!/bin/bash
# whole-step DRAM traffic (ncu range replay around one eager step) with and without the weight L2 prefetch, + timing A/B
set -u
O=gpurun_out/${1:-t1}; mkdir -p $O
python __graft_entry__.py > $O/build.log 2>&1
for cfg in default nopf; do
  M=""; [ $cfg = nopf ] && M="MTV_TC_MASK=0x1cbff"
  for b in 1 8; do
    env $M MTV_NO_GRAPH=1 timeout 600 ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
        --csv --log-file $O/range_${cfg}_b${b}.csv python scripts/step_traffic.py $b > $O/range_${cfg}_b${b}.log 2>&1
    echo "range $cfg b$b rc=$?"; grep -E "dram__bytes|gpu__time" $O/range_${cfg}_b${b}.csv | cut -d, -f10-
  done
done
AB_BATCHES="1 8" bash scripts/gpu_ab_env.sh ${1:-t1}_ab - MTV_TC_MASK=0x1cbff 2>&1 | grep -v sampling
