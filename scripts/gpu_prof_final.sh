#!/bin/bash
# Final evidence: ncu launch lists with DRAM traffic (B=1, B=8) and --set full captures of the dominant kernels.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
K='regex:k_(conv|attn|gn_|splitk|linear|temb|pack|ddim|apply|tc_|qkv)'
for b in 1 8; do
  # skip the three warm-up steps (plan launches + pack_in + ddim_step each), capture the timed step
  n=$(python -c "import torch;from moditalker_b200 import BASE_UNET_CONFIG as C,DiffusionWrapper as W,UNetModel as U;from moditalker_b200.synth import synth_state_dict as S;m=W(U(**C));m.load_state_dict(S(C,0,'diffusion_model.'));m=m.cuda().eval();print(m.diffusion_model.plan_info($b)['launches'])" 2>/dev/null | tail -1)
  echo "launches per step (B=$b): $n" | tee -a gpurun_out/summary.txt
  timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
      -k "$K" -s $((3 * n)) -c $n --csv --log-file gpurun_out/launches_b${b}.csv \
      env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --chunks-per-gpu $b > gpurun_out/ncu_b${b}.log 2>&1
  echo "ncu list b$b rc=$?" | tee -a gpurun_out/summary.txt
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 300 -c 6 -o gpurun_out/full_conv_tc_b1 -f \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1
echo "ncu full conv b1 rc=$?" | tee -a gpurun_out/summary.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 340 -c 6 -o gpurun_out/full_conv_tc_b8 -f \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --chunks-per-gpu 8 > gpurun_out/ncu_full8.log 2>&1
echo "ncu full conv b8 rc=$?" | tee -a gpurun_out/summary.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_attn_tc|k_tc_splitk_reduce_apply|k_apply_norm" -s 120 -c 8 -o gpurun_out/full_other_b1 -f \
    env MTV_NO_GRAPH=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
echo "ncu full other b1 rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out/*.ncu-rep
