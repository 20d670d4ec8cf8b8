import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, b200wave
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n, s = (64, 304) if which == "cfg2" else (64, 1024)
xfm = b200wave.DWTForward(J=3, wave="db3", mode="symmetric").to(dev)
ifm = b200wave.DWTInverse(wave="db3", mode="symmetric").to(dev)
x = torch.rand(n, 1, s, s, device=dev)
with torch.no_grad():
    for _ in range(2):
        yl, yh = xfm(x)
        rec = ifm((yl, yh))
torch.cuda.synchronize()
print("ok")
