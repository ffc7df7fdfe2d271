"""world_size-2 gloo test of the data-parallel plumbing (unirestore_b200/dist.py): shard -> per-rank forward ->
all-gather equals the single-process result, independent of the number of ranks (CPU stand-in model)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unirestore_b200 import dist as ud


class _StandIn(torch.nn.Module):
    """Per-image op with the forward signature of DiffUIE (uses the injected noise so slicing errors show up)."""

    def forward(self, images, task, noise=None):
        n_post, n_diff = noise
        s = (n_post.mean(dim=(1, 2, 3)) + 2 * n_diff.mean(dim=(1, 2, 3))).view(-1, 1, 1, 1)
        return images * 0.5 + s + (1.0 if task == "ir" else 0.0)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        images = torch.rand(n, 3, 512, 512, generator=g)
        out = ud.restore_sharded(_StandIn(), images, "ir", seed=99)
        q.put((rank, out[:, :, :2, :2].clone()))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_batch():
    for n in (1, 2, 7, 8, 32):
        for world in (1, 2, 3, 8):
            spans = [ud.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


@pytest.mark.parametrize("n", [4, 5])
def test_two_rank_gather_equals_single_process(n):
    g = torch.Generator().manual_seed(5)
    images = torch.rand(n, 3, 512, 512, generator=g)
    noise = ud.global_noise(n, (64, 64), 99)
    want = _StandIn()(images, "ir", noise=noise)[:, :, :2, :2]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    for r in range(2):
        assert torch.allclose(got[r], want, atol=1e-6), r
