#!/bin/bash
# compute-sanitizer over smoke() (tiny config: one forward + a 4-step DDIM sample through the C ABI; exercises split-KV attention with
# the DSMEM merge) and over one base.yaml B=1 forward (exercises the in-kernel split-K reduction / tickets / slot sums)
set -u
O=gpurun_out/${1:-san}; mkdir -p $O
python __graft_entry__.py > $O/build.log 2>&1
cat > /tmp/san_base.py <<PY
import sys; sys.path.insert(0, ".")
import torch
from moditalker_b200 import BASE_UNET_CONFIG as C, DiffusionWrapper as W, UNetModel as U
from moditalker_b200.synth import synth_inputs, synth_state_dict
m = W(U(**C)); m.load_state_dict(synth_state_dict(C, 0, "diffusion_model."), strict=True); m = m.cuda().eval()
x, c, ic, t = [v.cuda() for v in synth_inputs(1, seed=5)]
with torch.no_grad():
    a = m(x, c, ic, t); b = m(x, c, ic, t)
torch.cuda.synchronize()
print("base forward ok, finite:", bool(torch.isfinite(a).all()), "repeatable:", bool(torch.equal(a, b)))
PY
for tool in memcheck racecheck; do
  timeout ${SAN_TIMEOUT:-400} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" $O/sanitizer_${tool}_smoke.log | head -4
done
MTV_NO_GRAPH=1 timeout ${SAN_TIMEOUT:-400} compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 python /tmp/san_base.py > $O/sanitizer_memcheck_base.log 2>&1
echo "memcheck base rc=$?" | tee -a $O/summary.txt
grep -E "ERROR SUMMARY|base forward ok|Error" $O/sanitizer_memcheck_base.log | head -6
