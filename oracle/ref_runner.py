"""Loader of the UNMODIFIED reference staged under ``oracle/_ref`` (``make -C oracle ref``; git-ignored, travels with
the gpurun snapshot).  TEST / BENCH INFRASTRUCTURE ONLY: the product (the ``b200wave`` package) never imports this.

The reference imports PyWavelets at module top (``pytorch_wavelets/dwt/lowlevel.py:6``, ``dwt/transform2d.py:2``);
PyWavelets is not installed in this image, so ``oracle/pywt_standin`` (taps + ``dwt_coeff_len`` only) is put on the
path first -- the same arrangement the golden-vector generators use (``tests/golden/make_golden.py``).
"""
import importlib
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "ssim.py")) and \
        os.path.isdir(os.path.join(REF_DIR, "pytorch_wavelets"))


def load():
    """Returns (pytorch_wavelets module, ssim module) of the staged reference; raises if it is not there."""
    if not available():
        raise RuntimeError("the reference is not staged: run `make -C oracle ref` where /root/reference exists")
    for p in (os.path.join(HERE, "pywt_standin"), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pw = importlib.import_module("pytorch_wavelets")
        ss = importlib.import_module("ssim")
    if not os.path.abspath(pw.__file__).startswith(REF_DIR) or not os.path.abspath(ss.__file__).startswith(REF_DIR):
        raise RuntimeError("another pytorch_wavelets / ssim module shadows the staged reference")
    return pw, ss
