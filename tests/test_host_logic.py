"""Host-side logic that needs no GPU: DDPM schedule / step scalars against the
oracle's restatement of losses/ddpm.py, and the chunk sharding + all-gather over a
2-rank gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moditalker_b200 import DDPM, TINY_UNET_CONFIG, DiffusionWrapper, UNetModel, chunk_partition, sample_chunks_sharded
from oracle.unet_oracle import ddim_time_pairs, ddim_update, schedule


def _ddpm(S=100):
    return DDPM(DiffusionWrapper(UNetModel(**TINY_UNET_CONFIG)), channels=4, image_size=32, sampling_timesteps=S, w=0.0)


def test_schedule_buffers_match_oracle():
    d, s = _ddpm(), schedule()
    for k, v in s.items():
        assert torch.equal(getattr(d, k), v), k
    assert d.image_size == 2048 and d.channels == 4 and d.is_ddim_sampling and d.ddim_sampling_eta == 1.0
    assert d.num_timesteps == 1000


@pytest.mark.parametrize("S", [50, 100])
def test_time_pairs_and_scalars(S):
    d, s = _ddpm(S), schedule()
    pairs = d.time_pairs()
    assert pairs == ddim_time_pairs(1000, S) and len(pairs) == S and pairs[-1][1] == -1
    g = torch.Generator().manual_seed(0)
    img, eps, nz = (torch.randn(64, generator=g) for _ in range(3))
    for time, tn in (pairs[0], pairs[S // 2], pairs[-2], pairs[-1]):
        sr, srm1, san, c, sigma = d.step_scalars(time, tn)
        f = lambda v: torch.tensor(v, dtype=torch.float32)
        x0 = (f(sr) * img - f(srm1) * eps).clamp(-1, 1)
        mine = x0 if tn < 0 else x0 * f(san) + f(c) * eps + f(sigma) * nz
        assert torch.equal(mine, ddim_update(img, eps, nz, s, time, tn))   # bit-identical host scalars


def test_step_scalars_follow_the_live_schedule_buffers():
    """the fused step takes host scalars; they must track the registered buffers the reference indexes
    (losses/ddpm.py:390-394), also after load_state_dict / in-place edits."""
    d = _ddpm(50)
    t, tn = d.time_pairs()[10]
    before = d.step_scalars(t, tn)
    other = DDPM(DiffusionWrapper(UNetModel(**TINY_UNET_CONFIG)), channels=4, image_size=32, sampling_timesteps=50, w=0.0,
                 linear_start=0.0008, linear_end=0.012)
    d.load_state_dict(other.state_dict())                       # a checkpoint trained with another beta schedule
    after = d.step_scalars(t, tn)
    assert after == other.step_scalars(t, tn) and after != before
    with torch.no_grad():
        d.alphas_cumprod.mul_(0.5)                              # in-place edit of a live buffer
    assert d.step_scalars(t, tn) != after


def test_sampler_rejects_foreign_model_and_ddpm_mode():
    d = DDPM(torch.nn.Identity(), channels=4, sampling_timesteps=10)
    with pytest.raises(TypeError):
        d._unet()
    with pytest.raises(NotImplementedError):
        _ddpm(1000).sample(batch_size=1, cond=torch.zeros(1, 8, 2048), image_cond=torch.zeros(1, 4, 1024))


def test_chunk_partition():
    assert chunk_partition(9, 4, 0) == [0, 4, 8] and chunk_partition(9, 4, 3) == [3, 7]
    allc = sorted(sum((chunk_partition(9, 4, r) for r in range(4)), []))
    assert allc == list(range(9))
    assert chunk_partition(2, 4, 3) == []


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_chunks, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        cond = torch.randn(n_chunks, 8, 2048, generator=g)
        ic = torch.randn(n_chunks, 4, 1024, generator=g)
        # stand-in sampler with a per-chunk deterministic result, so the gathered
        # tensor is checkable without a GPU: z = f(cond, image_cond) per chunk
        fn = lambda c, i, ns: c[:, :4] * 2.0 + torch.nn.functional.pad(i, (0, 1024)) + (0 if ns is None else ns)
        z = sample_chunks_sharded(fn, cond, ic)
        want = fn(cond, ic, None)
        ret[rank] = bool(torch.equal(z, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_chunks", [4, 5, 1])
def test_sharded_sampling_gloo_world2(n_chunks):
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_chunks, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_bench_reference_arm_prints_one_json_line():
    """bench.py's contract: exactly ONE JSON line on stdout (library chatter goes to stderr), reference arm keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "tiny", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "chunk-steps/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under moditalker_b200/ may import it (statically checked over every module)."""
    import ast
    import glob
    import os

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "moditalker_b200")
    files = glob.glob(os.path.join(root, "**", "*.py"), recursive=True)
    assert files
    for f in files:
        for node in ast.walk(ast.parse(open(f).read())):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), f"{f} imports the oracle"
