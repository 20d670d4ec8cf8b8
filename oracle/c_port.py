"""ctypes wrapper of ``oracle/_build/libref_port.so`` (the C/OpenMP port of the reference's CPU path).

TEST INFRASTRUCTURE ONLY: used by tests (cross-check against the numpy oracle) and by ``bench.py`` for the
``cpu_baseline`` / ``--impl reference`` legs.  Never imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libref_port.so")
_lib = None
_fp = ctypes.POINTER(ctypes.c_float)


def build():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        i = ctypes.c_int
        lib.ref_num_threads.restype = i
        lib.ref_afb2d.argtypes = [_fp, i, i, i, _fp, _fp, i, _fp, _fp, i, i, _fp, _fp]
        lib.ref_sfb2d.argtypes = [_fp, _fp, i, i, i, _fp, _fp, i, _fp, _fp, i, i, _fp, i, i]
        lib.ref_ssim.argtypes = [_fp, _fp, i, i, i, i, _fp, i, i, _fp, _fp, _fp, _fp]
        _lib = lib
    return _lib


def num_threads():
    return int(load().ref_num_threads())


def use_all_cores():
    """Let the port use every core this process may run on (torchrun sets OMP_NUM_THREADS=1 for its workers, which
    would otherwise turn the all-cores CPU baseline into a single-thread one).  Returns the thread count."""
    import os
    lib = load()
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib.ref_set_num_threads.argtypes = [ctypes.c_int]
    lib.ref_set_num_threads.restype = None
    lib.ref_set_num_threads(int(n))
    return num_threads()


def _p(a):
    return None if a is None else a.ctypes.data_as(_fp)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


_MODE = {"zero": 0, "symmetric": 1, "per": 2, "periodization": 2, "reflect": 4, "periodic": 6}


def _coeff_len(n, l, mode):
    return (n + 1) // 2 if mode == 2 else (n + l - 1) // 2


def afb2d_level(x, hW, hH, mode):
    """x (N,C,H,W) fp32; hW/hH = (lo, hi) correlation taps along W / H.  Returns low, highs."""
    m = _MODE[mode] if isinstance(mode, str) else mode
    x = _f32(x)
    N, C, H, W = x.shape
    wl, wh, hl, hh = (_f32(t) for t in (hW[0], hW[1], hH[0], hH[1]))
    Ho, Wo = _coeff_len(H, hl.size, m), _coeff_len(W, wl.size, m)
    low = np.empty((N, C, Ho, Wo), np.float32)
    highs = np.empty((N, C, 3, Ho, Wo), np.float32)
    load().ref_afb2d(_p(x), N * C, H, W, _p(wl), _p(wh), wl.size, _p(hl), _p(hh), hl.size, m, _p(low), _p(highs))
    return low, highs


def sfb2d_level(low, highs, gW, gH, mode, out_hw=None):
    m = _MODE[mode] if isinstance(mode, str) else mode
    low = _f32(low)
    highs = None if highs is None else _f32(highs)
    N, C, h, w = low.shape
    wl, wh, hl, hh = (_f32(t) for t in (gW[0], gW[1], gH[0], gH[1]))
    full = (2 * h if m == 2 else 2 * h - hl.size + 2, 2 * w if m == 2 else 2 * w - wl.size + 2)
    oh, ow = full if out_hw is None else out_hw
    y = np.empty((N, C, oh, ow), np.float32)
    load().ref_sfb2d(_p(low), _p(highs), N * C, h, w, _p(wl), _p(wh), wl.size, _p(hl), _p(hh), hl.size, m, _p(y),
                     oh, ow)
    return y


def dwt_forward(x, J, h_col, h_row, mode):
    """DWTForward: the ``*_col`` buffers run along W, ``*_row`` along H (transform2d.py:70-71)."""
    yh, ll = [], x
    for _ in range(J):
        ll, high = afb2d_level(ll, h_col, h_row, mode)
        yh.append(high)
    return ll, yh


def dwt_inverse(yl, yh, g_col, g_row, mode):
    ll = yl
    for h in yh[::-1]:
        if h is not None:
            if ll.shape[-2] > h.shape[-2]:
                ll = ll[..., :-1, :]
            if ll.shape[-1] > h.shape[-1]:
                ll = ll[..., :-1]
        ll = sfb2d_level(ll, h, g_col, g_row, mode)
    return ll


def dwt_roundtrip_fwd_bwd(x, grec, J, h_col, h_row, g_col, g_row, mode):
    """The cfg2 step on the CPU: DWT fwd, IDWT fwd, then the reference's two custom backward chains
    (SFB2D.backward = analysis with the synthesis taps; AFB2D.backward = synthesis with the analysis taps)."""
    yl, yh = dwt_forward(x, J, h_col, h_row, mode)
    rec = dwt_inverse(yl, yh, g_col, g_row, mode)
    m = _MODE[mode]
    dy = grec
    dhs = []
    for j in range(J):
        dlow, dhigh = afb2d_level(dy, g_col, g_row, mode)
        dhs.append(dhigh)
        if j + 1 < J:
            tgt = yh[j + 1].shape[-2:]
            full = tuple(2 * t if m == 2 else 2 * t - len(g_col[0]) + 2 for t in tgt)
            if full != dlow.shape[-2:]:
                pad = np.zeros(dlow.shape[:2] + full, np.float32)
                pad[..., :dlow.shape[-2], :dlow.shape[-1]] = dlow
                dlow = pad
            dy = dlow
    shapes = [x.shape[-2:]] + [h.shape[-2:] for h in yh[:-1]]
    d = dlow
    for j in reversed(range(J)):
        d = sfb2d_level(d, dhs[j], h_col, h_row, mode, out_hw=tuple(shapes[j]))
    return yl, yh, rec, d


def ssim(img1, img2, win2d, size_average=True, grad_out=None, want_d1=False, want_d2=False):
    a, b = _f32(img1), _f32(img2)
    N, C, H, W = a.shape
    w = _f32(win2d)
    out = np.empty((1 if size_average else N,), np.float32)
    d1 = np.empty_like(a) if want_d1 else None
    d2 = np.empty_like(a) if want_d2 else None
    g = None if grad_out is None else _f32(np.atleast_1d(grad_out))
    load().ref_ssim(_p(a), _p(b), N, C, H, W, _p(w), w.shape[0], int(size_average), _p(out), _p(g), _p(d1), _p(d2))
    return (out[0] if size_average else out), d1, d2
