"""On-GPU multi-rank check: N NCCL ranks running ``unirestore_b200.dist.restore_sharded`` -- each
rank restores its shard of the batch, one all-gather of the decoded images -- reproduce the single-GPU forward on the
whole batch.  Skipped with fewer than 2 devices (run it with ``gpurun --gpus 2``)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
CFG = (dict(type="CFRM"), dict(type="scedit", num_inference_steps=2), dict(type="TFA", prompt_len=1, task=["ir", "seg"]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    from unirestore_b200 import dist as urdist
    from unirestore_b200.diffuie import DiffUIE
    from unirestore_b200.init_utils import deterministic_init_
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    m = deterministic_init_(DiffUIE(*CFG)).eval().requires_grad_(False).to(dev)
    m.use_cuda_graph = True                                            # the benchmarked configuration of the path
    img = torch.rand(n, 3, 512, 512, generator=torch.Generator().manual_seed(42)).to(dev)
    y = urdist.restore_sharded(m, img, "ir", seed=1234)
    y = urdist.restore_sharded(m, img, "ir", seed=1234)                # (second call: graph replay)
    torch.save(y.cpu(), os.path.join(out_dir, "rank%d.pt" % rank))
    if rank == 0:                                                      # single-GPU forward on the whole batch
        n_post, n_diff = urdist.global_noise(n, (64, 64), 1234, dev)
        m.use_cuda_graph = False
        torch.save(m(img, "ir", noise=(n_post, n_diff)).cpu(), os.path.join(out_dir, "single.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [4, 5])
def test_nccl_ranks_equal_single_gpu(tmp_path, n):
    from tests.util import assert_close
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 CUDA devices")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    single = torch.load(tmp_path / "single.pt")
    outs = [torch.load(tmp_path / ("rank%d.pt" % r)) for r in range(world)]
    for r in range(1, world):
        assert torch.equal(outs[r], outs[0]), "ranks hold different gathered tensors"
    # Not bit-exact by design: the per-rank batch changes GEMM tile shapes / split-K factors (summation order).
    assert_close(outs[0], single, 5e-3, "%d NCCL ranks (restore_sharded, %d images) vs single-GPU forward" % (world, n))
