"""GPU, world_size 2, NCCL: the sharded SSIM mean with the CUDA kernel underneath (skipped with fewer than 2 GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import b200wave
        from b200wave.sharding import shard_batch, sharded_ssim
        gen = torch.Generator().manual_seed(11)
        a = torch.rand((n_batch, 1, 96, 80), generator=gen)
        b = (a + 0.1 * torch.randn(a.shape, generator=gen)).clamp_(0, 1)
        crit = b200wave.SSIM()
        # single-GPU result over the whole batch, on this rank's device
        fa = a.to(dev).requires_grad_(True)
        full = crit(fa, b.to(dev))
        full.backward()
        # sharded: value identical on every rank; gradients under DDP's averaging convention
        la = shard_batch(a).to(dev).requires_grad_(True)
        lb = shard_batch(b).to(dev)
        val = sharded_ssim(crit, la, lb, reducer="mean", total_count=a.numel())
        val.backward()
        assert abs(float(val) - float(full)) < 1e-6, (float(val), float(full))
        lo = rank * (n_batch // world) + min(rank, n_batch % world)
        want = fa.grad[lo:lo + la.shape[0]] * world     # averaged over ranks later: each shard carries world x its share
        err = float((la.grad - want).abs().max() / want.abs().max())
        assert err < 1e-5, err
        np.save(os.path.join(out_dir, "nccl%d.npy" % rank), np.array([float(val)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_batch", [6, 7])
def test_sharded_ssim_on_two_gpus(tmp_path, n_batch):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_batch, str(tmp_path)), nprocs=2, join=True)
    v0 = np.load(tmp_path / "nccl0.npy")
    v1 = np.load(tmp_path / "nccl1.npy")
    assert v0 == v1
