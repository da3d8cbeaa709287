"""Chain kernel vs stand-alone launches on the same GPU: outputs must be bit-identical (same arithmetic, only the
kernel boundaries differ).  Usage: python scripts/chain_check.py [cfg] [B ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from moditalker_b200 import DiffusionWrapper, UNetModel  # noqa: E402
from moditalker_b200.synth import synth_inputs, synth_state_dict  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import config_by_name  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "base"
Bs = [int(a) for a in sys.argv[2:]] or [1, 3, 8]
cfg = config_by_name(cfg_name)
sd = synth_state_dict(cfg, 0, "diffusion_model.")


def make(mask):
    if mask is None:
        os.environ.pop("MTV_TC_MASK", None)
    else:
        os.environ["MTV_TC_MASK"] = mask
    m = DiffusionWrapper(UNetModel(**cfg))
    m.load_state_dict(sd, strict=True)
    return m.to("cuda:0").eval()


chain, plain = make("0x1fff"), make("0xfff")
ok = True
for B in Bs:
    x, cond, ic, t = synth_inputs(B, seed=123 + B, t=[(37 * (i + 1)) % 1000 for i in range(B)])
    outs = []
    for m in (chain, plain):
        with torch.no_grad():
            for _ in range(3):          # eager, capture, replay
                o = m(x.cuda(), cond.cuda(), ic.cuda(), t.cuda())
        torch.cuda.synchronize()
        outs.append(o.cpu())
    same = torch.equal(outs[0], outs[1])
    err = float((outs[0].double() - outs[1].double()).norm() / outs[1].double().norm())
    info_c, info_p = chain.diffusion_model.plan_info(B), plain.diffusion_model.plan_info(B)
    print(f"{cfg_name} B={B}: chain launches {info_c['launches']} vs {info_p['launches']}; bit-identical={same} rel-L2={err:.2e} "
          f"finite={bool(torch.isfinite(outs[0]).all())}", flush=True)
    ok = ok and (same or err < 1e-6)
sys.exit(0 if ok else 1)
