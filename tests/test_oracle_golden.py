"""Pins the CPU oracle (oracle/unet_oracle.py) against outputs of the reference's
OWN modules (fixtures written by oracle/make_golden.py in the build container).
fp32 vs fp32 on the same CPU ops: expected agreement ~1e-6; bound 2e-5."""
import pytest
import torch

from conftest import assert_close, config_by_name, load_golden
from moditalker_b200.synth import synth_inputs, synth_state_dict
from oracle.make_golden import TAP_CSTRIDE, TAP_LSTRIDE
from oracle.unet_oracle import Oracle, ddim_sample, ddim_time_pairs, schedule

TIGHT = dict(tol_l2=2e-5, tol_max=2e-5)


def _oracle_for(g, dtype=torch.float32):
    cfg = config_by_name(g["config"])
    return cfg, Oracle(cfg, synth_state_dict(cfg, int(g["wseed"])), dtype=dtype)


@pytest.mark.parametrize("name", ["unet_tiny_b2", "unet_base_b1"])
def test_forward_matches_reference(name):
    g = load_golden(name)
    cfg, orc = _oracle_for(g)
    x, cond, ic, t = synth_inputs(int(g["B"]), int(g["iseed"]), int(g["ic_len"]), [int(v) for v in g["t"]])
    taps = [k[4:] for k in g if k.startswith("tap_")]
    orc.capture = set(taps)
    eps = orc.forward(x, cond, ic, t)
    assert_close(eps, torch.from_numpy(g["eps"]), f"{name} eps", **TIGHT)
    for k in taps:
        got = orc.taps[k][:, ::TAP_CSTRIDE, ::TAP_LSTRIDE]
        assert_close(got, torch.from_numpy(g["tap_" + k]), f"{name} tap {k}", **TIGHT)


def test_forward_fp64_oracle_is_consistent():
    """fp64 restatement vs the reference's fp32 output: the reference's own rounding
    error (~6e-7) bounds the difference."""
    g = load_golden("unet_tiny_b2")
    cfg, orc = _oracle_for(g, torch.float64)
    x, cond, ic, t = synth_inputs(int(g["B"]), int(g["iseed"]), int(g["ic_len"]), [int(v) for v in g["t"]])
    assert_close(orc.forward(x, cond, ic, t), torch.from_numpy(g["eps"]), "fp64 oracle vs reference fp32", **TIGHT)


def _draw_noise(g, shape):
    """Replays the reference's global-generator draws: one randn(shape) (or the fixed
    q_sample noise after manual_seed(1004)), then one randn_like per non-final step."""
    S, noised, ratio = int(g["S"]), bool(g["noised"]), float(g["ratio"])
    pairs = ddim_time_pairs(1000, S)
    if noised:
        pairs = pairs[int(len(pairs) * (1 - ratio)):]
    torch.manual_seed(int(g["nseed"]))
    if noised and bool(g["fix_noise"]):
        torch.manual_seed(1004)
    out = [torch.randn(shape)]
    for _, tn in pairs:
        if tn >= 0:
            out.append(torch.randn(shape))
    return out


@pytest.mark.parametrize("name", ["ddim_tiny_s10", "ddim_tiny_noised"])
def test_ddim_matches_reference(name):
    g = load_golden(name)
    cfg, orc = _oracle_for(g)
    B = int(g["B"])
    x, cond, ic, _ = synth_inputs(B, int(g["iseed"]), int(g["ic_len"]), 0)
    noises = _draw_noise(g, (B, 4, 2048))
    if bool(g["noised"]):
        z = ddim_sample(orc, cond, ic, noises, int(g["S"]), x_start=torch.tanh(x), ratio=float(g["ratio"]))
    else:
        z = ddim_sample(orc, cond, ic, noises, int(g["S"]))
    assert_close(z, torch.from_numpy(g["z"]), f"{name} final latent", tol_l2=1e-4, tol_max=1e-4)


def test_schedule_buffers():
    s = schedule()
    ac = s["alphas_cumprod"]
    assert ac.shape == (1000,) and ac.dtype == torch.float32
    assert abs(float(ac[0]) - (1 - 0.0015)) < 1e-7
    assert torch.all(ac[1:] < ac[:-1])
    assert ddim_time_pairs(1000, 100)[0] == (999, 989) and ddim_time_pairs(1000, 100)[-1] == (9, -1)
    assert len(ddim_time_pairs(1000, 50)) == 50
