"""Pinned-memory copy bandwidth of the box: H2D alone, D2H alone, both at once (two streams).

    python tools/pcie_probe.py                                    one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/pcie_probe.py                                       N ranks copying AT THE SAME TIME

Under torchrun every rank copies 64 MiB buffers between its own pinned host memory and its own GPU inside a
barrier-bracketed region (the same structure as bench.py's e2e leg: one cudaMemcpyAsync per buffer per direction);
rank 0 prints one JSON line with the per-rank minimum and the whole-box sum per direction.  That sum is the bound of
`e2e` at N GPUs (`e2e.frac_of_host_bound` in the bench line uses the committed profiles/r02_pcie_probe.json).
"""
import json
import os

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = 64 * 1024 * 1024 // 4
h1, h2 = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
d1, d2 = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=20):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if world > 1:   # the slowest rank bounds a synchronous step
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    per_dir = reps * n * 4 / (ms * 1e-3) / 1e9     # GB/s per direction per rank, at the slowest rank's time
    return per_dir


run(True, True, 3)
up, down, both = run(True, False), run(False, True), run(True, True)
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_copy": n * 4,
                      "h2d_gbs_per_rank": round(up, 2), "d2h_gbs_per_rank": round(down, 2),
                      "duplex_gbs_per_rank_per_direction": round(both, 2),
                      "h2d_gbs_box": round(up * world, 2), "d2h_gbs_box": round(down * world, 2),
                      "duplex_gbs_box_per_direction": round(both * world, 2),
                      "note": "all ranks copy at once; time = max over ranks"}))
if world > 1:
    dist.destroy_process_group()
