"""Minimal stand-in for PyWavelets -- TEST INFRASTRUCTURE ONLY.

PyWavelets (``requirements.txt:3`` of the vendored pytorch_wavelets, unpinned
``PyWavelets>=1.0.0``) is not installed in this image and there is no network.
The reference imports it at module top (``dwt/lowlevel.py:6``,
``dwt/transform2d.py:2``) but only uses

* ``pywt.Wavelet(name).{dec_lo,dec_hi,rec_lo,rec_hi}`` (``transform2d.py:22-26,91-95``)
* ``isinstance(w, pywt.Wavelet)``
* ``pywt.dwt_coeff_len(N, L, mode)`` (``lowlevel.py:153``)

This module provides exactly those, restating PyWavelets' published rules:
``dbN`` = minimum-phase Daubechies scaling filter obtained by spectral
factorisation (sum = sqrt(2)), ``dec_lo = rec_lo[::-1]``,
``dec_hi[k] = (-1)^(k+1) rec_lo[k]``, ``rec_hi = dec_hi[::-1]``; ``haar == db1``;
``dwt_coeff_len = ceil(N/2)`` for periodization else ``floor((N+L-1)/2)``.
``bior2.4`` (used by ``tests/test_dwt.py:37``) is the CDF 5/9-tap spline pair in
PyWavelets' zero-padded 10-tap layout.

It is used (a) to import the unmodified reference in the build container when
generating golden vectors (``tests/golden/make_golden.py``) and (b) by the
oracle.  The product package never imports it.
"""
import math

import numpy as np

__all__ = ["Wavelet", "dwt_coeff_len", "wavelist"]


def _daubechies_rec_lo(p):
    """rec_lo of dbp (length 2p) by spectral factorisation, float64."""
    if p == 1:
        return np.array([math.sqrt(0.5), math.sqrt(0.5)])
    # P(y) = sum_{k<p} C(p-1+k, k) y^k ; roots in y
    coeffs = [math.comb(p - 1 + k, k) for k in range(p)]
    yroots = np.roots(coeffs[::-1])
    zroots = []
    for y in yroots:
        # y = (2 - z - 1/z)/4  ->  z^2 - (2-4y) z + 1 = 0 ; keep |z| < 1
        b = 2.0 - 4.0 * y
        disc = np.sqrt(b * b - 4.0 + 0j)
        z1, z2 = (b + disc) / 2.0, (b - disc) / 2.0
        zroots.append(z1 if abs(z1) < 1 else z2)
    poly = np.array([1.0 + 0j])
    for _ in range(p):
        poly = np.convolve(poly, [1.0, 1.0])
    for z in zroots:
        poly = np.convolve(poly, [1.0, -z])
    h = np.real(poly)
    h = h * (math.sqrt(2.0) / h.sum())
    return h


_S2 = math.sqrt(2.0)
_BIOR24_DEC_LO = np.array([0.0, 3.0, -6.0, -16.0, 38.0, 90.0, 38.0, -16.0, -6.0, 3.0]) * _S2 / 128.0
_BIOR24_REC_LO = np.array([0.0, 0.0, 0.0, 1.0, 2.0, 1.0, 0.0, 0.0, 0.0, 0.0]) * _S2 / 4.0


def _qmf_bior(dec_lo, rec_lo):
    """PyWavelets biorthogonal convention: dec_hi[k] = (-1)^(k+1) rec_lo[k],
    rec_hi[k] = (-1)^k dec_lo[k]."""
    n = len(dec_lo)
    sgn = np.array([(-1.0) ** k for k in range(n)])
    dec_hi = -sgn * rec_lo
    rec_hi = sgn * dec_lo
    return dec_hi, rec_hi


class Wavelet(object):
    def __init__(self, name):
        if isinstance(name, Wavelet):
            name = name.name
        self.name = name
        key = name.lower()
        if key == "haar":
            key = "db1"
        if key.startswith("db") and key[2:].isdigit():
            p = int(key[2:])
            if not 1 <= p <= 12:
                raise ValueError("stand-in supports db1..db12, got %r" % name)
            rec_lo = _daubechies_rec_lo(p)
            dec_lo = rec_lo[::-1].copy()
            sgn = np.array([(-1.0) ** (k + 1) for k in range(2 * p)])
            dec_hi = sgn * rec_lo
            rec_hi = dec_hi[::-1].copy()
        elif key == "bior2.4":
            dec_lo, rec_lo = _BIOR24_DEC_LO.copy(), _BIOR24_REC_LO.copy()
            dec_hi, rec_hi = _qmf_bior(dec_lo, rec_lo)
        else:
            raise ValueError("Unknown wavelet name %r for the pywt stand-in" % name)
        self.dec_lo = [float(v) for v in dec_lo]
        self.dec_hi = [float(v) for v in dec_hi]
        self.rec_lo = [float(v) for v in rec_lo]
        self.rec_hi = [float(v) for v in rec_hi]
        self.dec_len = len(self.dec_lo)
        self.rec_len = len(self.rec_lo)

    @property
    def filter_bank(self):
        return (self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi)


def dwt_coeff_len(data_len, filter_len, mode="symmetric"):
    if isinstance(filter_len, Wavelet):
        filter_len = filter_len.dec_len
    if mode in ("per", "periodization"):
        return (data_len + 1) // 2
    return (data_len + filter_len - 1) // 2


def wavelist():
    return ["haar"] + ["db%d" % p for p in range(1, 13)] + ["bior2.4"]
