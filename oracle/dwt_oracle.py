"""CPU oracle for the 2-D DWT hot path -- TEST INFRASTRUCTURE ONLY.

A plain-numpy restatement (float64 by default) of the reference's analysis /
synthesis filter banks.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module; the product package never does and fails loudly when its CUDA library
is missing.

Parity pin: this oracle is checked against outputs of the *unmodified*
reference executed in the build container (``tests/golden/make_golden.py``
imports ``/root/reference/pytorch_wavelets`` with the ``pywt`` stand-in and
writes ``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` replays them.

Each function cites the reference lines it follows
(paths relative to ``/root/reference/pytorch_wavelets/pytorch_wavelets``).
Filters are passed exactly as the reference's module buffers hold them:
analysis taps already time-reversed (``prep_filt_afb1d`` ``dwt/lowlevel.py:970-971``),
synthesis taps as-is (``prep_filt_sfb1d`` ``dwt/lowlevel.py:918-922``).
"""
import numpy as np

# mode_to_int / int_to_mode, dwt/lowlevel.py:274-309
MODE_CODES = {"zero": 0, "symmetric": 1, "per": 2, "periodization": 2,
              "constant": 3, "reflect": 4, "replicate": 5, "periodic": 6}
INT_TO_MODE = {0: "zero", 1: "symmetric", 2: "periodization", 3: "constant",
               4: "reflect", 5: "replicate", 6: "periodic"}


def mode_to_int(mode):
    if mode not in MODE_CODES:
        raise ValueError("Unkown pad type: {}".format(mode))
    return MODE_CODES[mode]


def int_to_mode(mode):
    if mode not in INT_TO_MODE:
        raise ValueError("Unkown pad type: {}".format(mode))
    return INT_TO_MODE[mode]


def dwt_coeff_len(n, l, mode):
    """pywt.dwt_coeff_len as called at dwt/lowlevel.py:153."""
    if mode in ("per", "periodization"):
        return (n + 1) // 2
    return (n + l - 1) // 2


def reflect_index(idx, minx, maxx):
    """pytorch_wavelets/utils.py:146-163 (``reflect``) on integer arrays."""
    x = np.asarray(idx, dtype=np.float64)
    rng = maxx - minx
    rng2 = 2 * rng
    mod = np.fmod(x - minx, rng2)
    normed = np.where(mod < 0, mod + rng2, mod)
    out = np.where(normed >= rng, rng2 - normed, normed) + minx
    return out.astype(np.int64)


def _pad_axis(x, before, after, mode, axis):
    """mypad, dwt/lowlevel.py:28-88, restricted to one axis."""
    n = x.shape[axis]
    if mode == "symmetric":
        xe = reflect_index(np.arange(-before, n + after), -0.5, n - 0.5)
        return np.take(x, xe, axis=axis)
    if mode == "periodic":
        xe = np.pad(np.arange(n), (before, after), mode="wrap")
        return np.take(x, xe, axis=axis)
    if mode == "reflect":
        if before >= n or after >= n:
            raise RuntimeError("reflect padding must be smaller than the dimension")
        xe = np.pad(np.arange(n), (before, after), mode="reflect")
        return np.take(x, xe, axis=axis)
    if mode == "zero":
        pads = [(0, 0)] * x.ndim
        pads[axis] = (before, after)
        return np.pad(x, pads)
    raise ValueError("Unkown pad type: {}".format(mode))


def _corr_stride2(xp, w, axis, nout):
    """y[k] = sum_j w[j] xp[2k+j] along ``axis`` (what F.conv2d with stride 2 does)."""
    xp = np.moveaxis(xp, axis, -1)
    y = np.zeros(xp.shape[:-1] + (nout,), dtype=xp.dtype)
    for j, wj in enumerate(w):
        y += wj * xp[..., j:j + 2 * nout - 1:2]
    return np.moveaxis(y, -1, axis)


def afb1d(x, h0, h1, mode="zero", axis=-1):
    """1-D analysis bank, dwt/lowlevel.py:91-172.  ``h0``/``h1`` are the
    correlation taps (already reversed).  Returns (lo, hi)."""
    h0 = np.asarray(h0, dtype=x.dtype).ravel()
    h1 = np.asarray(h1, dtype=x.dtype).ravel()
    L = h0.size
    assert h1.size == L
    L2 = L // 2
    axis = axis % x.ndim
    N = x.shape[axis]
    if mode in ("per", "periodization"):
        if N % 2 == 1:                                           # :135-140
            last = np.take(x, [N - 1], axis=axis)
            x = np.concatenate([x, last], axis=axis)
            N += 1
        x = np.roll(x, -L2, axis=axis)                           # :141
        pads = [(0, 0)] * x.ndim
        pads[axis] = (L - 1, L - 1)                              # :142-143
        xp = np.pad(x, pads)
        nfull = (N + 2 * (L - 1) - L) // 2 + 1
        outs = []
        for h in (h0, h1):
            y = _corr_stride2(xp, h, axis, nfull)
            y = np.moveaxis(y, axis, -1).copy()
            N2 = N // 2
            y[..., :L2] = y[..., :L2] + y[..., N2:N2 + L2]       # :146-150
            outs.append(np.moveaxis(y[..., :N2], -1, axis))
        return outs[0], outs[1]
    outsize = dwt_coeff_len(N, L, mode)                          # :153
    p = 2 * (outsize - 1) - N + L                                # :154
    if mode == "zero":                                           # :155-164
        xp = _pad_axis(x, p // 2, p // 2 + (p % 2), "zero", axis)
    elif mode in ("symmetric", "reflect", "periodic"):           # :165-168
        xp = _pad_axis(x, p // 2, (p + 1) // 2, mode, axis)
    else:
        raise ValueError("Unkown pad type: {}".format(mode))
    return (_corr_stride2(xp, h0, axis, outsize),
            _corr_stride2(xp, h1, axis, outsize))


def sfb1d(lo, hi, g0, g1, mode="zero", axis=-1):
    """1-D synthesis bank, dwt/lowlevel.py:226-271."""
    g0 = np.asarray(g0, dtype=lo.dtype).ravel()
    g1 = np.asarray(g1, dtype=lo.dtype).ravel()
    L = g0.size
    assert g1.size == L
    axis = axis % lo.ndim
    lo_m = np.moveaxis(lo, axis, -1)
    hi_m = np.moveaxis(hi, axis, -1)
    M = lo_m.shape[-1]
    N = 2 * M
    full = np.zeros(lo_m.shape[:-1] + (2 * (M - 1) + L,), dtype=lo.dtype)
    for j in range(L):                                           # conv_transpose2d, stride 2
        full[..., j:j + 2 * M - 1:2] += lo_m * g0[j] + hi_m * g1[j]
    if mode in ("per", "periodization"):                         # :252-261
        y = full.copy()
        if L > 2:
            y[..., :L - 2] = y[..., :L - 2] + y[..., N:N + L - 2]
        y = y[..., :N]
        y = np.roll(y, 1 - L // 2, axis=-1)
    elif mode in ("zero", "symmetric", "reflect", "periodic"):   # :263-267
        y = full[..., L - 2: full.shape[-1] - (L - 2)]
    else:
        raise ValueError("Unkown pad type: {}".format(mode))
    return np.moveaxis(y, -1, axis)


def afb2d_level(x, hW_lo, hW_hi, hH_lo, hH_hi, mode):
    """AFB2D.forward, dwt/lowlevel.py:336-347.  ``hW_*`` filter along W (dim 3,
    the ``h*_row`` parameters of AFB2D.forward), ``hH_*`` along H (dim 2).
    Returns low (N,C,h,w) and highs (N,C,3,h,w) with band order LH, HL, HH
    = (W-lo,H-hi), (W-hi,H-lo), (W-hi,H-hi)."""
    if isinstance(mode, int):
        mode = int_to_mode(mode)
    lo_w, hi_w = afb1d(x, hW_lo, hW_hi, mode, axis=3)
    ll, lh = afb1d(lo_w, hH_lo, hH_hi, mode, axis=2)
    hl, hh = afb1d(hi_w, hH_lo, hH_hi, mode, axis=2)
    highs = np.stack([lh, hl, hh], axis=2)
    return np.ascontiguousarray(ll), np.ascontiguousarray(highs)


def sfb2d_level(low, highs, gW_lo, gW_hi, gH_lo, gH_hi, mode):
    """SFB2D.forward, dwt/lowlevel.py:671-680."""
    if isinstance(mode, int):
        mode = int_to_mode(mode)
    lh, hl, hh = highs[:, :, 0], highs[:, :, 1], highs[:, :, 2]
    lo = sfb1d(low, lh, gH_lo, gH_hi, mode, axis=2)
    hi = sfb1d(hl, hh, gH_lo, gH_hi, mode, axis=2)
    return np.ascontiguousarray(sfb1d(lo, hi, gW_lo, gW_hi, mode, axis=3))


def afb2d_backward(dlow, dhighs, hW_lo, hW_hi, hH_lo, hH_hi, mode, in_hw):
    """AFB2D.backward, dwt/lowlevel.py:349-365: synthesis with the saved
    (reversed) analysis taps, then crop to the forward input's H, W."""
    dx = sfb2d_level(dlow, dhighs, hW_lo, hW_hi, hH_lo, hH_hi, mode)
    H, W = in_hw
    return np.ascontiguousarray(dx[:, :, :H, :W])


def sfb2d_backward(dy, gW_lo, gW_hi, gH_lo, gH_hi, mode):
    """SFB2D.backward, dwt/lowlevel.py:682-694: analysis of dy with the
    un-reversed synthesis taps used as correlation kernels."""
    return afb2d_level(dy, gW_lo, gW_hi, gH_lo, gH_hi, mode)


def prep_afb(dec_lo, dec_hi):
    """prep_filt_afb1d, dwt/lowlevel.py:956-975 (time reversal)."""
    return (np.asarray(dec_lo, dtype=np.float64)[::-1].copy(),
            np.asarray(dec_hi, dtype=np.float64)[::-1].copy())


def dwt_forward(x, J, h_col, h_row, mode):
    """DWTForward.forward, dwt/transform2d.py:44-74.  ``h_col``=(lo,hi) are the
    module's ``h*_col`` buffers (prepped) which -- because of the argument
    order at transform2d.py:70-71 -- filter along W; ``h_row`` filter along H."""
    yh = []
    ll = x
    for _ in range(J):
        ll, high = afb2d_level(ll, h_col[0], h_col[1], h_row[0], h_row[1], mode)
        yh.append(high)
    return ll, yh


def dwt_inverse(yl, yh, g_col, g_row, mode):
    """DWTInverse.forward, dwt/transform2d.py:111-148."""
    ll = yl
    for h in yh[::-1]:
        if h is None:                                            # :137-139
            h = np.zeros(ll.shape[:2] + (3,) + ll.shape[-2:], dtype=ll.dtype)
        if ll.shape[-2] > h.shape[-2]:                           # :141-145
            ll = ll[..., :-1, :]
        if ll.shape[-1] > h.shape[-1]:
            ll = ll[..., :-1]
        ll = sfb2d_level(ll, h, g_col[0], g_col[1], g_row[0], g_row[1], mode)
    return ll
