#!/bin/bash
# ncu evidence run (1 GPU): launch list of ~2 forwards + full-set capture of the tensor-core tap-GEMM.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"k_(conv|attn|gn_|splitk|linear|temb|pack|ddim|apply|tc_)" -c 1200 --csv --log-file gpurun_out/launches.csv \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?" | tee gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_conv_tc" -s 40 -c 6 -o gpurun_out/prof_conv_tc \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out | tail -5
