"""Filter taps for ``DWTForward`` / ``DWTInverse``.

The reference obtains its taps from PyWavelets (``pywt.Wavelet(wave)``,
``dwt/transform2d.py:22-26, 91-95``).  PyWavelets is an optional dependency
here: when it is importable it is used for every wavelet name, exactly like the
reference; when it is not, the built-in tables below cover ``haar`` and
``db1``..``db20`` (literal, generated to 20 digits by
``tools/gen_wavelet_tables.py``) and ``bior2.4`` (the one non-Daubechies wavelet
the reference's tests use, ``tests/test_dwt.py:37``).

Conventions (PyWavelets'): ``dec_lo = rec_lo[::-1]``,
``dec_hi[k] = (-1)^(k+1) rec_lo[k]``, ``rec_hi = dec_hi[::-1]`` for the
orthogonal families; for biorthogonal ones ``dec_hi[k] = (-1)^(k+1) rec_lo[k]``
and ``rec_hi[k] = (-1)^k dec_lo[k]``.
"""
import math

from ._wavelet_tables import DB_REC_LO

try:  # optional: identical behaviour to the reference when PyWavelets is present
    import pywt as _pywt
    if not hasattr(_pywt, "Wavelet"):
        _pywt = None
except Exception:  # pragma: no cover - not installed in the build image
    _pywt = None

_S2 = math.sqrt(2.0)
_BIOR = {
    # name: (dec_lo, rec_lo)  -- CDF spline pair, PyWavelets' zero-padded layout
    "bior2.4": ([c * _S2 / 128.0 for c in (0.0, 3.0, -6.0, -16.0, 38.0, 90.0, 38.0, -16.0, -6.0, 3.0)],
                [c * _S2 / 4.0 for c in (0.0, 0.0, 0.0, 1.0, 2.0, 1.0, 0.0, 0.0, 0.0, 0.0)]),
}


class Wavelet(object):
    """Duck-type of ``pywt.Wavelet``: ``dec_lo, dec_hi, rec_lo, rec_hi``,
    ``dec_len``, ``rec_len``, ``filter_bank``, ``name``."""

    def __init__(self, name):
        if isinstance(name, Wavelet):
            name = name.name
        if not isinstance(name, str):
            raise TypeError("wavelet name must be a string, got %r" % (type(name),))
        self.name = name
        key = name.lower()
        if key == "haar":
            key = "db1"
        if key.startswith("db") and key[2:].isdigit() and int(key[2:]) in DB_REC_LO:
            rec_lo = [float(c) for c in DB_REC_LO[int(key[2:])]]
            n = len(rec_lo)
            dec_lo = rec_lo[::-1]
            dec_hi = [((-1.0) ** (k + 1)) * rec_lo[k] for k in range(n)]
            rec_hi = dec_hi[::-1]
        elif key in _BIOR:
            dec_lo, rec_lo = [list(map(float, f)) for f in _BIOR[key]]
            n = len(dec_lo)
            dec_hi = [((-1.0) ** (k + 1)) * rec_lo[k] for k in range(n)]
            rec_hi = [((-1.0) ** k) * dec_lo[k] for k in range(n)]
        else:
            raise ValueError(
                "Unknown wavelet name '%s': built-in tables cover %s; install PyWavelets for the "
                "other families or pass the filters as a tuple of arrays" % (name, ", ".join(wavelist())))
        self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi = dec_lo, dec_hi, rec_lo, rec_hi
        self.dec_len = len(dec_lo)
        self.rec_len = len(rec_lo)

    @property
    def filter_bank(self):
        return (self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi)

    def __repr__(self):
        return "Wavelet(%r)" % self.name


def wavelist():
    return ["haar"] + ["db%d" % p for p in sorted(DB_REC_LO)] + sorted(_BIOR)


def is_wavelet(obj):
    """True for our ``Wavelet``, a real ``pywt.Wavelet`` or anything exposing the four filters."""
    return all(hasattr(obj, a) for a in ("dec_lo", "dec_hi", "rec_lo", "rec_hi"))


def as_wavelet(wave):
    """``pywt.Wavelet(wave)`` when PyWavelets is installed, else the built-in table."""
    if is_wavelet(wave):
        return wave
    if _pywt is not None:
        return _pywt.Wavelet(wave)
    return Wavelet(wave)


def dwt_coeff_len(data_len, filter_len, mode="symmetric"):
    """``pywt.dwt_coeff_len`` as used at ``dwt/lowlevel.py:153``."""
    if is_wavelet(filter_len):
        filter_len = len(filter_len.dec_lo)
    if mode in ("per", "periodization"):
        return (data_len + 1) // 2
    return (data_len + filter_len - 1) // 2
