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

from b200wave.sharding import average_gradients, global_mean, shard_batch, shard_range
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
        g = global_mean(local, la.numel(), reducer="sum")
        g.backward()
        full = ssim_oracle.ssim(a, b)
        assert abs(float(g) - full) < 1e-12, (float(g), full)
        # summed over ranks: d global / d local mean = share of the elements held by this rank
        assert abs(float(local.grad) - (hi - lo) / n_batch) < 1e-12
        # DDP convention (gradients averaged over ranks), count known on the host: no read-back
        local2 = torch.tensor(ssim_oracle.ssim(la.numpy(), lb.numpy()), requires_grad=True)
        g2 = global_mean(local2, la.numel(), reducer="mean", total_count=ta.numel())
        g2.backward()
        assert abs(float(g2) - full) < 1e-12
        assert abs(float(local2.grad) - world * (hi - lo) / n_batch) < 1e-12
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([float(g)]))
    finally:
        dist.destroy_process_group()


def _ddp_worker(rank, world, port, n_batch, out_dir):
    """A shared parameter, a mean-type loss over a sharded batch, gradients averaged over ranks as DDP does: the
    reduced gradient must equal the unsharded one -- also for ragged shards."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(3)
        x = torch.rand((n_batch, 1, 6, 7), generator=gen, dtype=torch.float64)
        w0 = torch.rand((1, 1, 6, 7), generator=gen, dtype=torch.float64)

        def loss_of(w, batch):      # any per-pixel map followed by .mean(), like ssim_map.mean() (ssim.py:35)
            return torch.sigmoid(w * batch).mean()

        w_full = w0.clone().requires_grad_(True)
        loss_of(w_full, x).backward()
        w = w0.clone().requires_grad_(True)
        xs = shard_batch(x)
        total = global_mean(loss_of(w, xs), xs.numel(), reducer="mean", total_count=x.numel())
        total.backward()
        grad = w.grad.clone()
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)
        grad /= world               # DDP's gradient averaging
        assert torch.allclose(grad, w_full.grad, rtol=1e-12, atol=1e-15), (grad - w_full.grad).abs().max()
        assert abs(float(total) - float(loss_of(w0, x))) < 1e-12
        np.save(os.path.join(out_dir, "ddp%d.npy" % rank), np.array([1.0]))
    finally:
        dist.destroy_process_group()


def _avg_worker(rank, world, port, out_dir):
    """average_gradients: one flat all-reduce equals the per-parameter mean over ranks; parameters that no rank used
    (grad None) are skipped."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(5)
        params = [torch.nn.Parameter(torch.rand(s, generator=gen, dtype=torch.float64)) for s in ((3, 4), (7,), (2, 2, 2))]
        unused = torch.nn.Parameter(torch.zeros(5, dtype=torch.float64))
        local = [torch.full_like(p, float(rank + 1)) * (i + 1) for i, p in enumerate(params)]
        for p, g in zip(params, local):
            p.grad = g.clone()
        average_gradients(params + [unused])
        mean_rank = sum(range(1, world + 1)) / world
        for i, p in enumerate(params):
            assert torch.allclose(p.grad, torch.full_like(p, mean_rank * (i + 1)), rtol=0, atol=1e-15)
        assert unused.grad is None
        np.save(os.path.join(out_dir, "avg%d.npy" % rank), np.array([1.0]))
    finally:
        dist.destroy_process_group()


def test_average_gradients_flat_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_avg_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "avg0.npy").exists() and (tmp_path / "avg1.npy").exists()
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.ones(3)
    average_gradients([p])                      # no process group: a no-op
    assert torch.equal(p.grad, torch.ones(3))


@pytest.mark.parametrize("n_batch", [4, 5])
def test_sharded_mean_gradients_match_unsharded_under_ddp_averaging(tmp_path, n_batch):
    port = _free_port()
    mp.spawn(_ddp_worker, args=(2, port, n_batch, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ddp0.npy").exists() and (tmp_path / "ddp1.npy").exists()


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
