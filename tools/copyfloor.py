"""What a plain device copy of the same algorithmic bytes costs in the harness the kernels are timed in
(CUDA graph of back-to-back launches over rotating buffers > L2): the practical floor for small passes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
from kbench import timeit, PEAK  # noqa: E402

for label, n in [("64x304x304 (cfg2 level: 24 MB in, 24 MB out)", 64 * 304 * 304),
                 ("64x154x154", 64 * 154 * 154), ("64x1024x1024", 64 * 1024 * 1024), ("8x304x304", 8 * 304 * 304)]:
    nsets = max(2, int(2 * 126e6 * 1.05 / (4 * n)) + 1)
    src = [torch.rand(n, device="cuda") for _ in range(nsets)]
    dst = [torch.empty(n, device="cuda") for _ in range(nsets)]
    t = timeit(lambda i: dst[i % nsets].copy_(src[i % nsets]), nsets)
    print("copy %-48s %7.1f us  %6.0f GB/s (%4.1f%% of %.0f)" % (label, t * 1e6, 8 * n / t / 1e9, 8 * n / t / 1e9 / PEAK * 100, PEAK), flush=True)
