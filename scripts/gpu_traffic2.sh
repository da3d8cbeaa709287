#!/bin/bash
set -u
O=gpurun_out/${1:-t2}; mkdir -p $O
python __graft_entry__.py > $O/build.log 2>&1
timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
for b in 1 8; do
  env MTV_NO_GRAPH=1 timeout 600 ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
      --csv --log-file $O/range_b${b}.csv python scripts/step_traffic.py $b > $O/range_b${b}.log 2>&1
  echo "range b$b rc=$?"; grep -E "dram__bytes|gpu__time" $O/range_b${b}.csv | cut -d, -f10-
done
AB_BATCHES="1 8" bash scripts/gpu_ab_env.sh ${1:-t2}_ab - 2>&1 | grep -v sampling
python -c "
import torch
from moditalker_b200 import BASE_UNET_CONFIG as C, DiffusionWrapper as W, UNetModel as U
from moditalker_b200.synth import synth_state_dict as S
m=W(U(**C)); m.load_state_dict(S(C,0,'diffusion_model.')); m=m.cuda().eval()
for b in (1,8): print('plan', b, m.diffusion_model.plan_info(b))
"
