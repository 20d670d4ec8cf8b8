"""One multi-level transform pair a few times (target of ncu captures): python tools/one_dwt.py [n h w wave mode J reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import b200wave  # noqa: E402

a = sys.argv[1:]
n, h, w, wave, mode, J, reps = (int(a[0]), int(a[1]), int(a[2]), a[3], a[4], int(a[5]), int(a[6])) if len(a) >= 7 else (64, 304, 304, "db3", "symmetric", 3, 4)
xfm = b200wave.DWTForward(J=J, wave=wave, mode=mode).cuda()
ifm = b200wave.DWTInverse(wave=wave, mode=mode).cuda()
xs = [torch.rand(n, 1, h, w, device="cuda") for _ in range(reps)]
with torch.no_grad():
    for x in xs:
        c = xfm(x)
        y = ifm(c)
torch.cuda.synchronize()
print("ok", float((y - xs[-1]).abs().max()))
