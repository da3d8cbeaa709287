#!/bin/bash
# Round-2 ncu evidence from HEAD: launch lists with DRAM bytes (B=1, B=8), whole-step DRAM traffic (range replay), --set full
# captures of the dominant kernels.  Usage: bash scripts/gpu_prof_r2.sh [tag]
set -u
TAG=${1:-prof}
O=gpurun_out/$TAG; mkdir -p $O
python __graft_entry__.py > $O/build.log 2>&1
K='regex:k_(conv|attn|gn_|splitk|linear|temb|pack|ddim|apply|tc_|qkv)'
for b in 1 8; do
  n=$(python -c "import torch;from moditalker_b200 import BASE_UNET_CONFIG as C,DiffusionWrapper as W,UNetModel as U;from moditalker_b200.synth import synth_state_dict as S;m=W(U(**C));m.load_state_dict(S(C,0,'diffusion_model.'));m=m.cuda().eval();print(m.diffusion_model.plan_info($b)['launches'])" 2>/dev/null | tail -1)
  echo "launches per forward (B=$b): $n" | tee -a $O/summary.txt
  # skip the warm-up steps (n plan launches + pack_in + ddim_step each), capture one whole step (forward + ddim_step)
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
      -k "$K" -s $((3 * (n + 1))) -c $((n + 1)) --csv --log-file $O/launches_b${b}.csv \
      env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --chunks-per-gpu $b > $O/ncu_b${b}.log 2>&1
  echo "ncu list b$b rc=$?" | tee -a $O/summary.txt
done
# whole-step DRAM traffic without serialising the kernels: range replay around ONE step (graph launch + ddim step)
for b in 1 8; do
  env MTV_NO_GRAPH=1 timeout 600 ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
      --csv --log-file $O/range_b${b}.csv python scripts/step_traffic.py $b > $O/range_b${b}.log 2>&1
  echo "ncu range b$b rc=$?" | tee -a $O/summary.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 300 -c 8 -o $O/full_conv_tc_b1 -f \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/ncu_full1.log 2>&1
echo "ncu full conv b1 rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 300 -c 8 -o $O/full_conv_tc_b8 -f \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --chunks-per-gpu 8 > $O/ncu_full8.log 2>&1
echo "ncu full conv b8 rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_attn_tc|k_apply_norm" -s 100 -c 10 -o $O/full_other_b1 -f \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/ncu_full2.log 2>&1
echo "ncu full other b1 rc=$?" | tee -a $O/summary.txt
ls -la $O/*.ncu-rep 2>/dev/null
