"""Decode the per-CTA timeline a B200W_TIMELINE build writes (debug tool)."""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 16)
t0 = a[:, 0][a[:, 0] > 0].min()
lvl = (a[:, 3] >> np.uint64(48)).astype(int)
plane = ((a[:, 3] >> np.uint64(24)) & np.uint64(0xffffff)).astype(int)
border = (a[:, 4] >> np.uint64(32)).astype(int)
GHZ = 1.965
ck = lambda k: (a[:, k].astype(np.int64) - a[:, 15].astype(np.int64)) / (GHZ * 1e3)   # us since CTA start (SM clock)
st = (a[:, 0].astype(np.int64) - int(t0)) / 1e3
fin = (a[:, 5].astype(np.int64) - int(t0)) / 1e3
rd = st + ck(1)
en = st + ck(2)
print("items", len(a), "span us", fin.max())
for L in sorted(set(lvl)):
    for b in (0, 1):
        m = (lvl == L) & (border == b)
        if not m.any():
            continue
        print("level %d %s: n=%4d start %6.1f..%6.1f  ready %6.1f..%6.1f  end %6.1f..%6.1f  fin ..%6.1f | wait med %.1f  work med %.1f max %.1f  signal med %.1f" % (
            L, "border" if b else "ring  ", m.sum(), st[m].min(), st[m].max(), rd[m].min(), rd[m].max(), en[m].min(), en[m].max(), fin[m].max(),
            np.median(rd[m] - st[m]), np.median(en[m] - rd[m]), (en[m] - rd[m]).max(), np.median(fin[m] - en[m])))
# last planes: per-plane chain
for pl in sorted(set(plane))[-2:]:
    for L in sorted(set(lvl)):
        m = (lvl == L) & (plane == pl)
        print("  plane %d level %d: start %.1f ready %.1f..%.1f end %.1f..%.1f fin %.1f" % (pl, L, st[m].min(), rd[m].min(), rd[m].max(), en[m].min(), en[m].max(), fin[m].max()))

for L in (0, 1, 2):
    m = (lvl == L) & (border == 0)
    if not m.any():
        continue
    rel = lambda k: np.median(ck(k)[m] - ck(1)[m])
    print("level-%d ring CTA, warp 0, us after ready: prologue issued %.2f | pair0 ready %.2f done %.2f | pair1 ready %.2f done %.2f | pair2 %.2f %.2f | pair3 %.2f %.2f | end %.2f" % (
        L, rel(14), rel(6), rel(7), rel(8), rel(9), rel(10), rel(11), rel(12), rel(13), rel(2)))
    pl = plane[m].max()
    mm = m & (plane == pl)
    rel = lambda k: np.median(ck(k)[mm] - ck(1)[mm])
    print("   last plane only:                       prologue issued %.2f | pair0 ready %.2f done %.2f | pair1 ready %.2f done %.2f | pair2 %.2f %.2f | pair3 %.2f %.2f | end %.2f" % (
        rel(14), rel(6), rel(7), rel(8), rel(9), rel(10), rel(11), rel(12), rel(13), rel(2)))
