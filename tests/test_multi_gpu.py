"""Chunk-sharded sampling on 2 GPUs (NCCL): the gathered latents must equal the single-GPU
result for the same per-chunk noise — sharding is invisible in the output."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _chunk_noise_fn(chunk_ids):
    """noise_fn for DDPM: every draw is a stack of per-chunk tensors seeded by (chunk id, draw index),
    so a chunk sees the same noise whatever rank / local batch it lands in."""
    state = {"n": 0}

    def fn(kind, shape, device):
        k = state["n"]; state["n"] += 1
        outs = []
        for c in chunk_ids:
            g = torch.Generator().manual_seed(1000003 * int(c) + k)
            outs.append(torch.randn(tuple(shape[1:]), generator=g))
        return torch.stack(outs).to(device)
    return fn


def _sample(rank_dev, cond, ic, ids, steps=6):
    from moditalker_b200 import DDPM, TINY_UNET_CONFIG, DiffusionWrapper, UNetModel
    from moditalker_b200.synth import synth_state_dict
    cfg = TINY_UNET_CONFIG
    m = DiffusionWrapper(UNetModel(**cfg))
    m.load_state_dict(synth_state_dict(cfg, 0, "diffusion_model."), strict=True)
    m = m.to(rank_dev).eval()
    d = DDPM(m, channels=4, image_size=32, sampling_timesteps=steps, w=0.0).to(rank_dev)
    d.noise_fn = _chunk_noise_fn(ids)
    return d.sample(batch_size=cond.shape[0], cond=cond, image_cond=ic)


def _worker(rank, world, port, n_chunks, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from moditalker_b200 import chunk_partition, sample_chunks_sharded
        from moditalker_b200.synth import synth_inputs
        _, cond, ic, _ = synth_inputs(n_chunks, seed=9)
        cond, ic = cond.to(dev), ic.to(dev)
        mine = chunk_partition(n_chunks, world, rank)
        z = sample_chunks_sharded(lambda c, i, ns: _sample(dev, c, i, mine), cond, ic)
        torch.cuda.synchronize()
        if rank == 0:
            ref = _sample(dev, cond, ic, list(range(n_chunks)))
            err = float((z - ref).norm() / ref.norm())
            ret["err"] = err
            ret["shape"] = tuple(z.shape)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("n_chunks", [4, 3])
def test_sharded_sampling_nccl_world2(n_chunks):
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n_chunks, ret), nprocs=2, join=True)
    assert ret["shape"] == (n_chunks, 4, 2048)
    assert ret["err"] < 1e-4, f"sharded vs single-GPU latents differ: rel-L2 {ret['err']:.3e}"
