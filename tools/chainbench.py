"""Time whole multi-level transforms (one launch each) over rotating inputs: cfg2 and sweep shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import b200wave  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kbench import timeit, PEAK  # noqa: E402

dev = "cuda"


def pass_bytes(n, h, w, L, J, mode):
    tot = h * w
    for _ in range(J):
        h = (h + 1) // 2 if mode == "periodization" else (h + L - 1) // 2
        w = (w + 1) // 2 if mode == "periodization" else (w + L - 1) // 2
        tot += 3 * h * w
    return 4 * n * (tot + h * w)


def case(n, h, w, wave, mode, J):
    xfm = b200wave.DWTForward(J=J, wave=wave, mode=mode).to(dev)
    ifm = b200wave.DWTInverse(wave=wave, mode=mode).to(dev)
    L = xfm.h0_col.numel()
    nsets = max(2, int(2 * 126e6 * 1.05 / (4 * n * h * w)) + 1)
    xs = [torch.rand(n, 1, h, w, device=dev) for _ in range(nsets)]
    with torch.no_grad():
        cs = [xfm(x) for x in xs]
        ta = timeit(lambda i: xfm(xs[i % nsets]), nsets)
        ts = timeit(lambda i: ifm(cs[i % nsets]), nsets)
    by = pass_bytes(n, h, w, L, J, mode)
    from b200wave import _cabi
    ks = sorted(set(_cabi.recent_kernels(8)))
    print("%-5s %-13s J=%d %4dx%4dx%4d  DWT %7.1f us %5.0f GB/s (%4.1f%%)   IDWT %7.1f us %5.0f GB/s (%4.1f%%)" % (
        wave, mode, J, n, h, w, ta * 1e6, by / ta / 1e9, by / ta / 1e9 / PEAK * 100, ts * 1e6, by / ts / 1e9,
        by / ts / 1e9 / PEAK * 100), " ".join(k.replace("_kernel", "") for k in ks), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "cfg2":
    case(64, 304, 304, "db3", "symmetric", 3)
    case(64, 304, 304, "db3", "symmetric", 2)
    case(64, 154, 154, "db3", "symmetric", 2)
    case(64, 304, 304, "haar", "zero", 3)
    case(256, 128, 128, "db2", "reflect", 3)
elif __name__ == "__main__":
    case(64, 304, 304, "db3", "symmetric", 1)
    case(64, 304, 304, "db3", "symmetric", 2)
    case(64, 304, 304, "db3", "symmetric", 3)
    case(64, 154, 154, "db3", "symmetric", 1)
    case(64, 154, 154, "db3", "symmetric", 2)
    case(8, 304, 304, "haar", "zero", 3)
    case(64, 1024, 1024, "db3", "symmetric", 1)
    case(64, 1024, 1024, "db3", "symmetric", 3)
    case(64, 1024, 1024, "haar", "zero", 3)
    case(16, 2048, 2048, "db4", "zero", 5)
    case(64, 1024, 1024, "db8", "symmetric", 3)
