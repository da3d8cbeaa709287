#!/bin/bash
# final ncu launch lists (warm L2), B=1 and B=8, one forward each
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none \
    -k regex:"k_(conv|attn|gn_|splitk|linear|temb|pack|ddim|apply|tc_|qkv)" -s 1000 -c 340 --csv --log-file gpurun_out/launches_b1.csv \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
echo "ncu b1 rc=$?" | tee gpurun_out/summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none \
    -k regex:"k_(conv|attn|gn_|splitk|linear|temb|pack|ddim|apply|tc_|qkv)" -s 970 -c 330 --csv --log-file gpurun_out/launches_b8.csv \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --chunks-per-gpu 8 > gpurun_out/ncu_b8.log 2>&1
echo "ncu b8 rc=$?" | tee -a gpurun_out/summary.txt
