#!/bin/bash
# Round-2 closing evidence in one call: ncu captures from HEAD (scripts/gpu_prof_r2.sh), in-kernel timing, and the bench lines
# (default B=1 with every baseline leg, B=8, longvid, CPU reference arm).  Usage: bash scripts/gpu_final_r2.sh [tag]
set -u
TAG=${1:-final}
O=gpurun_out/$TAG; mkdir -p $O
bash scripts/gpu_prof_r2.sh $TAG
timeout 120 python scripts/tc_timing.py 1 > $O/tc_timing_b1.log 2>&1
timeout 300 python bench.py --impl reference --gpus 1 --steps 4 --warmup 1 > $O/bench_ref.json 2> $O/bench.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py > $O/bench_b1.json 2>> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 300 python bench.py --chunks-per-gpu 8 --steps 100 --no-cpu-baseline > $O/bench_b8.json 2>> $O/bench.err; echo "bench b8 rc=$?" | tee -a $O/summary.txt
timeout 300 python bench.py --config longvid --steps 100 --no-cpu-baseline > $O/bench_longvid_b1.json 2>> $O/bench.err; echo "bench longvid rc=$?" | tee -a $O/summary.txt
for f in bench_ref bench_b1 bench_b8 bench_longvid_b1; do cut -c1-160 $O/$f.json; done
tail -3 $O/bench.err
