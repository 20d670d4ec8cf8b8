"""Tiny driver for ncu: one multi-level DWT + IDWT on a BASELINE-shaped batch.

    python tools/prof_dwt.py [n] [size] [wave] [mode] [J]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import b200wave  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
s = int(sys.argv[2]) if len(sys.argv) > 2 else 304
wave = sys.argv[3] if len(sys.argv) > 3 else "db3"
mode = sys.argv[4] if len(sys.argv) > 4 else "symmetric"
J = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = "cuda"
xfm = b200wave.DWTForward(J=J, wave=wave, mode=mode).to(dev)
ifm = b200wave.DWTInverse(wave=wave, mode=mode).to(dev)
x = torch.rand(n, 1, s, s, device=dev)
with torch.no_grad():
    for _ in range(2):
        yl, yh = xfm(x)
        rec = ifm((yl, yh))
torch.cuda.synchronize()
print("ok", float((rec[..., :s, :s] - x).abs().max()))
