"""Debug helper: one inverse transform on a given shape against the oracle (python tools/try_idwt.py n h w wave mode J)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pywt_standin"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import b200wave  # noqa: E402
from b200wave import _cabi  # noqa: E402
from oracle import dwt_oracle  # noqa: E402

a = sys.argv[1:]
n, h, w, wave, mode, J = int(a[0]), int(a[1]), int(a[2]), a[3], a[4], int(a[5])
rng = np.random.default_rng(0)
x = rng.standard_normal((n, 1, h, w)).astype(np.float32)
xfm = b200wave.DWTForward(J=J, wave=wave, mode=mode).cuda()
ifm = b200wave.DWTInverse(wave=wave, mode=mode).cuda()
hc = (xfm.h0_col.flatten().cpu().numpy().astype(np.float64), xfm.h1_col.flatten().cpu().numpy().astype(np.float64))
gc = (ifm.g0_col.flatten().cpu().numpy().astype(np.float64), ifm.g1_col.flatten().cpu().numpy().astype(np.float64))
oyl, oyh = dwt_oracle.dwt_forward(x.astype(np.float64), J, hc, hc, mode)
orec = dwt_oracle.dwt_inverse(oyl, oyh, gc, gc, mode)
yl = torch.tensor(oyl, dtype=torch.float32, device="cuda")
yh = [torch.tensor(t, dtype=torch.float32, device="cuda") for t in oyh]
with torch.no_grad():
    rec = ifm((yl, yh))
torch.cuda.synchronize()
err = np.abs(rec.cpu().numpy() - orec).max() / np.abs(orec).max()
print("idwt %s: kernels %s  rel err %.2e" % (" ".join(a), _cabi.recent_kernels(2), err))
