"""CPU, world_size 2, gloo: the N>1 host logic (batch sharding + the scalar SSIM mean exchange).
The per-shard arithmetic is stood in for by the oracle here (no GPU); on the GPU box the same helpers
wrap the CUDA kernels (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200wave.sharding import global_mean, shard_batch, shard_range
from oracle import ssim_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        a = rng.random((n_batch, 2, 20, 24))
        b = np.clip(a + 0.1 * rng.standard_normal(a.shape), 0, 1)
        ta, tb = torch.tensor(a), torch.tensor(b)
        la, lb = shard_batch(ta), shard_batch(tb)
        lo, hi = shard_range(n_batch, rank, world)
        assert la.shape[0] == hi - lo
        # stand-in for the fused kernel on this rank's shard
        local = torch.tensor(ssim_oracle.ssim(la.numpy(), lb.numpy()), requires_grad=True)
        g = global_mean(local, la.numel())
        g.backward()
        full = ssim_oracle.ssim(a, b)
        assert abs(float(g) - full) < 1e-12, (float(g), full)
        # d global / d local mean = share of the elements held by this rank
        assert abs(float(local.grad) - (hi - lo) / n_batch) < 1e-12
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([float(g)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_batch", [4, 5])
def test_sharded_ssim_mean_world2(tmp_path, n_batch):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_batch, str(tmp_path)), nprocs=2, join=True)
    v0 = np.load(tmp_path / "rank0.npy")
    v1 = np.load(tmp_path / "rank1.npy")
    assert v0 == v1


def test_global_mean_single_process_is_identity():
    t = torch.tensor(0.25, requires_grad=True)
    assert global_mean(t, 10) is t
