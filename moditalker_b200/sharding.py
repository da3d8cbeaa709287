"""Chunk-level multi-GPU sampling.

Each 16-frame chunk's DDIM loop touches only its own (cond, image_cond, noise)
(sample.py:305-384) and nothing inside the UNet mixes samples, so chunks shard
over ranks with NO collective inside the loop: rank r samples chunks
{c : c mod W == r} as one local batch, and a single all-gather of the final
latents [n_local, 4, 2048] (NCCL over NVLink on GPUs, gloo in the CPU tests)
reassembles them in chunk order.  The reference has no multi-GPU inference path
(sample.py:185-203 pins rank 0), so this function has no reference counterpart.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.distributed as dist


def chunk_partition(n_chunks: int, world: int, rank: int) -> List[int]:
    """Round-robin: chunk c -> rank c % world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return list(range(rank, n_chunks, world))


def sample_chunks_sharded(
    sample_fn: Callable[[torch.Tensor, torch.Tensor, Optional[torch.Tensor]], torch.Tensor],
    cond: torch.Tensor,
    image_cond: torch.Tensor,
    noised_start: Optional[torch.Tensor] = None,
    group=None,
) -> torch.Tensor:
    """Sample all chunks of ``cond`` [N,8,2048] / ``image_cond`` [N,4,*] across the
    ranks of ``group`` and return the full [N,4,2048] latent on every rank.

    ``sample_fn(cond_local, image_cond_local, noised_start_local) -> z_local`` is the
    per-rank sampler, e.g. ``lambda c, ic, ns: ddpm.sample(batch_size=c.shape[0],
    cond=c, image_cond=ic, noised_start=ns, ratio_=0.25)``.  Ranks with fewer
    chunks pad their contribution so one fixed-size all-gather suffices.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = cond.shape[0]
    mine = chunk_partition(n, world, rank)
    per = (n + world - 1) // world
    z_local = None
    if mine:
        idx = torch.tensor(mine, device=cond.device)
        ns = noised_start.index_select(0, idx) if noised_start is not None else None
        z_local = sample_fn(cond.index_select(0, idx), image_cond.index_select(0, idx), ns)
    if world == 1:
        return z_local
    shape_tail = tuple(z_local.shape[1:]) if z_local is not None else (4, 2048)
    dtype = z_local.dtype if z_local is not None else torch.float32
    send = torch.zeros((per,) + shape_tail, device=cond.device, dtype=dtype)
    if mine:
        send[: len(mine)] = z_local
    recv = torch.empty((world * per,) + shape_tail, device=cond.device, dtype=dtype)
    dist.all_gather_into_tensor(recv, send, group=group)
    out = torch.empty((n,) + shape_tail, device=cond.device, dtype=dtype)
    for r in range(world):
        ids = chunk_partition(n, world, r)
        if ids:
            out[torch.tensor(ids, device=cond.device)] = recv[r * per: r * per + len(ids)]
    return out
