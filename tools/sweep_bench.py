"""cfg5 (BASELINE.json configs[4]): throughput sweep DWT J = 1..5 x db1..db8 on 1024 x 1024 and 2048 x 2048
single-channel images, at 1 / 2 / 4 / 8 B200 (weak scaling: every rank runs the same per-GPU batch, no data-path
collective) next to the unmodified reference on the host cores.  BENCH INFRASTRUCTURE (the CPU leg imports oracle/_ref).

    python tools/sweep_bench.py [--modes symmetric] [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sweep_bench.py

Per point: one multi-level forward transform and one inverse, each timed alone (CUDA graph of back-to-back launches
over rotating inputs larger than L2, CUDA events, max over ranks), algorithmic bytes = input + all coefficient tensors
(SURVEY.md 8d), fraction of the measured HBM peak.  Rank 0 prints one JSON line per point and a final summary line.
The CPU leg (rank 0, N = 1 only, --cpu) times DWTForward + DWTInverse of the unmodified reference on ONE image per
point with all host threads.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

PEAK = 6538.9
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
        PEAK = float(json.load(fh).get("hbm_gbs", PEAK))
except (OSError, ValueError):
    pass


def pass_bytes(n, h, w, L, J, mode):
    tot = h * w
    for _ in range(J):
        h = (h + 1) // 2 if mode == "periodization" else (h + L - 1) // 2
        w = (w + 1) // 2 if mode == "periodization" else (w + L - 1) // 2
        tot += 3 * h * w
    return 4 * n * (tot + h * w)


def timeit(fn, nsets, reps):
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    keep = []
    with torch.cuda.graph(g):
        for i in range(reps):
            keep.append(fn(i))
            if len(keep) > nsets:
                keep.pop(0)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist.is_initialized():
        dist.barrier()
        torch.cuda.synchronize()
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if dist.is_initialized():
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    del g, keep
    return ms / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="symmetric")
    ap.add_argument("--quick", action="store_true", help="db1, db4, db8 and J = 1, 3, 5 only")
    ap.add_argument("--cpu", action="store_true", help="time the unmodified reference (oracle/_ref) on the host cores too")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import b200wave

    ref = None
    if args.cpu and rank == 0 and world == 1:
        from oracle import ref_runner
        if ref_runner.available():
            ref, _ = ref_runner.load()
            torch.set_num_threads(os.cpu_count() or 1)

    waves = ["db1", "db4", "db8"] if args.quick else ["db%d" % k for k in range(1, 9)]
    levels = [1, 3, 5] if args.quick else [1, 2, 3, 4, 5]
    shapes = [(64, 1024, 1024), (16, 2048, 2048)]
    rows = []
    for mode in args.modes.split(","):
        for (n, h, w) in shapes:
            nsets = max(2, int(2 * 126e6 * 1.05 / (4 * n * h * w)) + 1)
            xs = [torch.rand(n, 1, h, w, device="cuda") for _ in range(nsets)]
            for wave in waves:
                for J in levels:
                    xfm = b200wave.DWTForward(J=J, wave=wave, mode=mode).cuda()
                    ifm = b200wave.DWTInverse(wave=wave, mode=mode).cuda()
                    L = xfm.h0_col.numel()
                    with torch.no_grad():
                        cs = [xfm(x) for x in xs]
                        ta = timeit(lambda i: xfm(xs[i % nsets]), nsets, 12)
                        ts = timeit(lambda i: ifm(cs[i % nsets]), nsets, 12)
                    by = pass_bytes(n, h, w, L, J, mode)
                    row = {"wave": wave, "J": J, "mode": mode, "per_gpu_batch": [n, 1, h, w], "n_gpus": world,
                           "dwt_us": ta * 1e6, "idwt_us": ts * 1e6,
                           "dwt_Mpix_s": world * n * h * w / ta / 1e6, "idwt_Mpix_s": world * n * h * w / ts / 1e6,
                           "dwt_hbm_frac": by / ta / 1e9 / PEAK, "idwt_hbm_frac": by / ts / 1e9 / PEAK}
                    if ref is not None:
                        x1 = torch.rand(1, 1, h, w)
                        rx, ri = ref.DWTForward(J=J, wave=wave, mode=mode), ref.DWTInverse(wave=wave, mode=mode)
                        with torch.no_grad():
                            c1 = rx(x1)
                            ri(c1)
                            td, ti = [], []
                            for _ in range(3):
                                t0 = time.perf_counter()
                                c1 = rx(x1)
                                t1 = time.perf_counter()
                                ri(c1)
                                t2 = time.perf_counter()
                                td.append(t1 - t0)
                                ti.append(t2 - t1)
                        row["cpu_reference"] = {"dwt_Mpix_s": h * w / sorted(td)[1] / 1e6, "idwt_Mpix_s": h * w / sorted(ti)[1] / 1e6,
                                                "cores": torch.get_num_threads(), "sample": "1 image, median of 3"}
                    rows.append(row)
                    if rank == 0:
                        print(json.dumps(row), flush=True)
                    del cs
            del xs
            torch.cuda.empty_cache()
    if rank == 0:
        fr = [r["dwt_hbm_frac"] for r in rows] + [r["idwt_hbm_frac"] for r in rows]
        print(json.dumps({"summary": "cfg5 sweep", "n_gpus": world, "points": len(rows), "hbm_peak_gbs": PEAK,
                          "hbm_frac_min": min(fr), "hbm_frac_median": sorted(fr)[len(fr) // 2], "hbm_frac_max": max(fr),
                          "Mpix_s_median_dwt": sorted(r["dwt_Mpix_s"] for r in rows)[len(rows) // 2]}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
