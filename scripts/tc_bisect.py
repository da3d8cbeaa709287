"""GPU-side diagnostic: which tensor-core op class departs from the fp32 CUDA-core path?
Runs the forward with MTV_TC_MASK restricted to one class at a time and prints the error of
every stage against the all-CUDA-core run.  (bits: 0-2 conv3x3 at level 0/1/2, 3 = 1x1 GEMMs,
4 = conv3x3 with fused skip 1x1, 5 = allow split-K, 6 = tcgen05 attention, 7 = tensor-core tiles at the
32-token level, 8 = GroupNorm statistics fused into the tap-GEMM epilogue / apply prologue, 9 = launch fusions:
qkv split in the GEMM epilogue, attention writes the proj operand, shared skip-operand apply)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from moditalker_b200 import BASE_UNET_CONFIG, TINY_UNET_CONFIG, DiffusionWrapper, UNetModel, build_arch
from moditalker_b200.arch import tokens_at
from moditalker_b200.synth import synth_inputs, synth_state_dict


def stages(cfg):
    arch = build_arch(**cfg)
    out = [("in0", arch.model_channels, tokens_at(0))]
    for i, st in enumerate(arch.input_stages[1:], start=1):
        out.append((f"in{i}", st.joint.channels, tokens_at(st.level_out)))
    out.append(("mid", arch.middle.joint.channels, tokens_at(arch.middle.level_out)))
    for i, st in enumerate(arch.output_stages):
        out.append((f"out{i}", st.joint.channels, tokens_at(st.level_out)))
    return out


def make(cfg, kernel_path, mask=None):
    if mask is not None:
        os.environ["MTV_TC_MASK"] = hex(mask)
    m = DiffusionWrapper(UNetModel(**cfg))
    m.diffusion_model.kernel_path = kernel_path
    m.load_state_dict(synth_state_dict(cfg, 0, "diffusion_model."), strict=True)
    return m.to("cuda:0").eval()


def main():
    cfg = TINY_UNET_CONFIG if "tiny" in sys.argv else BASE_UNET_CONFIG
    B = 2
    x, c, ic, t = [v.cuda() for v in synth_inputs(B, seed=5, t=[100, 900])]
    ref = make(cfg, 1)
    with torch.no_grad():
        e_ref = ref(x, c, ic, t)
    taps_ref = {k: ref.diffusion_model.debug_read(k, B, C, L) for k, C, L in stages(cfg)}
    cases = (("conv L0", 1), ("conv L1", 2), ("conv L2", 4), ("gemm 1x1", 8), ("conv+skip", 16), ("conv L2 + splitK", 4 | 32))
    if "new" in sys.argv:   # the classes verified so far as a base, plus one newer feature at a time
        cases = (("verified 0x1ff, no PDL", 0x1ff), ("verified 0x1ff + PDL", 0x1ff), ("+fused launches, no PDL", 0x3ff),
                 ("all (0x3ff + PDL)", 0x3ff))
    for name, mask in cases:
        os.environ["MTV_PDL"] = "0" if "no PDL" in name else "31"
        try:
            m = make(cfg, 0, mask)
            with torch.no_grad():
                for _ in range(3):          # eager, graph capture, graph replay
                    e = m(x, c, ic, t)
            torch.cuda.synchronize()
            errs = []
            for k, C, L in stages(cfg):
                a = m.diffusion_model.debug_read(k, B, C, L)
                errs.append((k, float((a - taps_ref[k]).norm() / taps_ref[k].norm())))
            worst = max(errs, key=lambda kv: kv[1])
            first = next((kv for kv in errs if kv[1] > 1e-4), None)
            print(f"[{name:18s}] eps err {float((e - e_ref).norm() / e_ref.norm()):.2e}  worst {worst[0]} {worst[1]:.2e}  first>1e-4: {first}", flush=True)
            del m
        except Exception as ex:  # noqa
            print(f"[{name:18s}] FAILED: {ex}", flush=True)


if __name__ == "__main__":
    main()
