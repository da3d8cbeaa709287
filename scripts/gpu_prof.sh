#!/bin/bash
# ncu evidence run (1 GPU): launch list of one forward + full-set captures of the two tensor-core kernels.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none \
    -k regex:"k_(conv|attn|gn_|splitk|linear|temb|pack|ddim|apply|tc_|qkv)" -s 1100 -c 360 --csv --log-file gpurun_out/launches.csv \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?" | tee gpurun_out/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_conv_tc" -s 350 -c 4 -o gpurun_out/prof_conv_tc \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --chunks-per-gpu 8 > gpurun_out/ncu_full.log 2>&1
echo "ncu full conv rc=$?" | tee -a gpurun_out/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_attn_tc" -s 40 -c 3 -o gpurun_out/prof_attn_tc \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --chunks-per-gpu 8 > gpurun_out/ncu_full2.log 2>&1
echo "ncu full attn rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out | tail -6
