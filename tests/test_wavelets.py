"""CPU: built-in tap tables of the product vs the independent float64 spectral factorisation of the
oracle's pywt stand-in, plus filter-bank identities."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pywt_standin"))
import pywt as standin  # noqa: E402

import b200wave  # noqa: E402
from b200wave import wavelets  # noqa: E402
from oracle import dwt_oracle  # noqa: E402


@pytest.mark.parametrize("name", ["haar"] + ["db%d" % p for p in range(1, 11)] + ["bior2.4"])
def test_tables_match_standin(name):
    a, b = wavelets.Wavelet(name), standin.Wavelet(name)
    for attr in ("dec_lo", "dec_hi", "rec_lo", "rec_hi"):
        assert np.allclose(getattr(a, attr), getattr(b, attr), rtol=0, atol=2e-11), attr


@pytest.mark.parametrize("p", range(1, 21))
def test_daubechies_orthonormal(p):
    h = np.array(wavelets.Wavelet("db%d" % p).rec_lo)
    assert len(h) == 2 * p
    assert abs(h.sum() - np.sqrt(2)) < 1e-13
    for m in range(p):
        acc = np.dot(h[:2 * p - 2 * m], h[2 * m:])
        assert abs(acc - (1.0 if m == 0 else 0.0)) < 1e-12


def test_pywavelets_anchor_values():
    # PyWavelets literals quoted in SURVEY.md 8c (its tables are accurate to ~2e-13)
    db2 = [0.48296291314469025, 0.836516303737469, 0.22414386804185735, -0.12940952255092145]
    assert np.allclose(wavelets.Wavelet("db2").rec_lo, db2, atol=1e-12, rtol=0)
    db4 = [0.23037781330885523, 0.7148465705525415, 0.6308807679295904]
    assert np.allclose(wavelets.Wavelet("db4").rec_lo[:3], db4, atol=1e-12, rtol=0)
    s = np.sqrt(0.5)
    w = wavelets.Wavelet("haar")
    assert np.allclose(w.dec_lo, [s, s]) and np.allclose(w.dec_hi, [-s, s]) and np.allclose(w.rec_hi, [s, -s])


@pytest.mark.parametrize("name", ["db1", "db2", "db5", "db8", "db12", "bior2.4"])
def test_perfect_reconstruction_periodization(name):
    w = wavelets.Wavelet(name)
    h = dwt_oracle.prep_afb(w.dec_lo, w.dec_hi)
    g = (np.array(w.rec_lo), np.array(w.rec_hi))
    x = np.random.default_rng(3).standard_normal((1, 2, 64, 48))
    yl, yh = dwt_oracle.dwt_forward(x, 2, h, h, "periodization")
    rec = dwt_oracle.dwt_inverse(yl, yh, g, g, "periodization")
    assert np.abs(rec - x).max() < 1e-10


def test_unknown_wavelet_and_duck_typing():
    with pytest.raises(ValueError):
        wavelets.Wavelet("sym4")

    class Fake(object):
        dec_lo, dec_hi, rec_lo, rec_hi = [1.0, 1.0], [-1.0, 1.0], [1.0, 1.0], [1.0, -1.0]
    m = b200wave.DWTForward(J=1, wave=Fake(), mode="zero")
    assert m.h0_col.flatten().tolist() == [1.0, 1.0] and m.h1_col.flatten().tolist() == [1.0, -1.0]
    assert wavelets.dwt_coeff_len(9, 6, "symmetric") == 7 and wavelets.dwt_coeff_len(9, 6, "periodization") == 5
