"""Decode the per-CTA timeline of the owner kernel written by a B200W_TIMELINE build (debug tool).

    B200W_TIMELINE=1 python -m b200wave._build; B200W_TIMELINE_FILE=/tmp/tl.bin python tools/onecase.py ...
    python tools/timeline_owner.py /tmp/tl.bin
"""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 16)
a = a[1:]            # row 0 also receives the stream-kernel marks of every CTA (item index 0)
a = a[a[:, 0] > 0]
GHZ = 1.965
t0 = a[:, 0].min()
start = (a[:, 0].astype(np.int64) - int(t0)) / 1e3
end = (a[:, 14].astype(np.int64) - int(t0)) / 1e3
print("CTAs %d  start %.2f..%.2f us  end %.2f..%.2f us (median %.2f)" % (len(a), start.min(), start.max(), end.min(), end.max(), np.median(end)))
ck = lambda k: (a[:, k].astype(np.int64) - a[:, 15].astype(np.int64)) / (GHZ * 1e3)
prev = ck(13)
print('maps (all levels) %.2f us' % np.median(prev))
for j in range(4):
    if 3 + 3 * j > 12 or not (np.median(a[:, 1 + 3 * j]) > 0):
        break
    m, i, b = ck(1 + 3 * j), ck(2 + 3 * j), ck(3 + 3 * j)
    print("level +%d: maps %.2f us | interior %.2f (max %.2f) | border %.2f (max %.2f) | level end at %.2f (max %.2f)" % (
        j, np.median(m - prev), np.median(i - m), (i - m).max(), np.median(b - i), (b - i).max(), np.median(b), b.max()))
    prev = b
