"""Per-CTA phase times of the TMA owner kernels (debug build-free: B200W_TMA_TIMELINE=<file> makes the launcher dump
32 clock stamps per CTA; slot 0 = %globaltimer at start, slot 1 = clock at start, slot 2 = set-up done, slots 3.. =
after each level / patch phase).

    B200W_TMA_TIMELINE=/tmp/tl.bin python tools/tma_timeline.py [n h w wave mode J]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import b200wave  # noqa: E402

n, h, w, wave, mode, J = 64, 304, 304, "db3", "symmetric", 3
if len(sys.argv) > 6:
    n, h, w, wave, mode, J = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5], int(sys.argv[6])
which = os.environ.get("TL_WHICH", "dwt")
path = os.environ.get("B200W_TMA_TIMELINE" if which == "dwt" else "B200W_TMA_TIMELINE_SFB")
assert path, "set B200W_TMA_TIMELINE (TL_WHICH=dwt) or B200W_TMA_TIMELINE_SFB (TL_WHICH=idwt)"
xfm = b200wave.DWTForward(J=J, wave=wave, mode=mode).cuda()
ifm = b200wave.DWTInverse(wave=wave, mode=mode).cuda()
x = torch.rand(n, 1, h, w, device="cuda")
with torch.no_grad():
    for _ in range(3):
        c = xfm(x)
        if which != "dwt":
            y = ifm(c)
torch.cuda.synchronize()
t = np.fromfile(path, dtype=np.uint64).reshape(-1, 64).astype(np.int64)
clk = 1.9   # GHz, close enough for phase shares
g0 = t[:, 0] - t[:, 0].min()
print("%s %dx%dx%d %s %s J=%d: %d CTAs; start spread (globaltimer) %.2f us" % (which, n, h, w, wave, mode, J, len(t), g0.max() / 1e3))
names = ["set-up"]
if which == "dwt":
    for j in range(1, J):
        names += ["level/pos %d" % (j - 1), "patch %d" % j]
    names += ["level/pos %d" % (J - 1)]
else:
    names += ["(mark)"] + ["level/pos %d" % j for j in range(J)]
prev = t[:, 1]
for i, nm in enumerate(names):
    cur = t[:, 2 + i]
    d = (cur - prev) / clk / 1e3
    print("  %-14s median %6.2f us  min %6.2f  max %6.2f" % (nm, np.median(d), d.min(), d.max()))
    prev = cur
tot = (prev - t[:, 1]) / clk / 1e3
print("  %-14s median %6.2f us  min %6.2f  max %6.2f" % ("total", np.median(tot), tot.min(), tot.max()))

ref = t[:, 2]
if which != "dwt":
    sys.exit(0)
for lbl, base in (("stream 0: patch(0): enter, landed, rows done, released", 8), ("stream 0: stage patched + released", 24), ("stream 0: consumers start stage", 16), ("stream 0: service start, issue(0) done, issue(1) done", 48),
                  ("stream 0: stage k free again (service warp)", 32), ("stream 0: stage k+D issued", 40)):
    vals = []
    for k in range(8):
        col = t[:, base + k]
        ok = col > 0
        if ok.any():
            vals.append("%5.2f" % float(np.median((col[ok] - ref[ok]) / clk / 1e3)))
    print("  %-44s %s  (us after set-up)" % (lbl, " ".join(vals)))

col = t[:, 56:61]
if (col > 0).all():
    d = np.median(col - col[:, :1], axis=0)
    print("  issue(stage 1) of stream 0, clocks after entry: fence %d, regular-check %d, expect_tx %d, tiles issued %d" % tuple(d[1:]))
