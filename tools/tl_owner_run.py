"""Run one multi-level DWT eagerly (timeline build) so that the last launch leaves its per-CTA timeline file."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import b200wave  # noqa: E402

n, h, w, wave, mode, J = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5], int(sys.argv[6])
xfm = b200wave.DWTForward(J=J, wave=wave, mode=mode).cuda()
xs = [torch.rand(n, 1, h, w, device="cuda") for _ in range(12)]
ifm = b200wave.DWTInverse(wave=wave, mode=mode).cuda()
with torch.no_grad():
    for x in xs:
        yl, yh = xfm(x)
        ifm((yl, yh))
torch.cuda.synchronize()
