"""Tiny driver for ncu: launches each hot kernel a few times on BASELINE-shaped inputs.

    python tools/prof_run.py [dwt|dwtbig|ssim|all] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import b200wave  # noqa: E402
from b200wave import lowlevel, ops  # noqa: E402
from b200wave.ssim import _win_taps  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
torch.manual_seed(0)
xfm = b200wave.DWTForward(J=3, wave="db3", mode="symmetric").to(dev)
ifm = b200wave.DWTInverse(wave="db3", mode="symmetric").to(dev)
if what in ("dwt", "all"):
    x = torch.rand(64, 1, 304, 304, device=dev)
    for _ in range(reps):
        yl, yh = xfm(x)
        rec = ifm((yl, yh))
if what in ("dwt", "dwtbig", "all"):
    xfm1 = b200wave.DWTForward(J=1, wave="db3", mode="symmetric").to(dev)
    big = torch.rand(64, 1, 1024, 1024, device=dev)
    for _ in range(reps):
        yl, yh = xfm1(big)
        rec = ifm((yl, yh))
if what in ("ssim", "all"):
    a = torch.rand(64, 1, 400, 400, device=dev)
    b = (a + 0.1 * torch.randn_like(a)).clamp_(0, 1)
    win = _win_taps(11)
    g = torch.ones((), device=dev)
    for _ in range(reps):
        val, maps = ops.ssim_fwd(a, b, win, True, 3)
        d1, _ = ops.ssim_bwd(a, b, maps, g, win, True, False)
torch.cuda.synchronize()
print("ok")
