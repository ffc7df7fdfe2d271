"""Data-parallel plumbing of the hot path: images shard naturally over ranks -- no op on
the path mixes batch elements -- so the only collective is the all-gather of the decoded images.

One process per GPU (``torch.distributed``, NCCL on B200 / gloo in the CPU tests).  RNG draws are generated for the
GLOBAL batch and sliced, so results do not depend on the number of ranks.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous split of ``n`` images over ``world`` ranks (first ``n % world`` ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, world: int | None = None, rank: int | None = None) -> torch.Tensor:
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(t.shape[0], world, rank)
    return t[lo:hi]


def global_noise(batch: int, latent_hw: tuple[int, int], seed: int, device="cpu"):
    """(posterior noise, diffusion noise) for the global batch, identical on every rank (autoencoder.py:152,
    unifie.py:87 draw them per call; here they are pre-drawn so sharding is invisible)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    h, w = latent_hw
    return (torch.randn(batch, 4, h, w, generator=g).to(device), torch.randn(batch, 4, h, w, generator=g).to(device))


def gather_restored(local: torch.Tensor, global_batch: int) -> torch.Tensor:
    """All-gather the decoded ``[b_local, 3, H, W]`` tensors into ``[global_batch, 3, H, W]`` on every rank."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(global_batch, world, r) for r in range(world)]
    maxb = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < maxb:
        pad = torch.cat([local, local.new_zeros((maxb - local.shape[0],) + tuple(local.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous())
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


def restore_sharded(model, images: torch.Tensor, task: str, seed: int = 1234) -> torch.Tensor:
    """``model.forward`` on this rank's shard of ``images`` (global batch, same tensor on every rank) followed by the
    all-gather; the result equals the single-process forward on the whole batch."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n = images.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    h, w = images.shape[-2:]
    hh, ww = max(h, 512) if min(h, w) >= 512 else round(h * 512 / min(h, w)), max(w, 512) if min(h, w) >= 512 else round(w * 512 / min(h, w))
    hh, ww = hh + (64 - hh % 64) % 64, ww + (64 - ww % 64) % 64
    n_post, n_diff = global_noise(n, (hh // 8, ww // 8), seed, images.device)
    out = model(images[lo:hi], task, noise=(n_post[lo:hi], n_diff[lo:hi]))
    return gather_restored(out, n)
