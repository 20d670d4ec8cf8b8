"""CPU: the C-ABI shared library builds for sm_100a, loads without a GPU, exports every symbol that
include/b200wave.h declares, and its host-only entry points / argument validation behave."""
import ctypes
import os
import re
import subprocess

import pytest

import b200wave  # noqa: F401
from b200wave import _build, _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200wave.h")


@pytest.fixture(scope="module")
def lib():
    return _cabi.load()


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200w_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_functions()
    assert names == sorted(_cabi.SYMBOLS), (names, sorted(_cabi.SYMBOLS))


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(_cabi.library_path())
    for name in declared_functions():
        assert hasattr(raw, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (b200w_\w+)", out))
    assert exported == set(declared_functions())


def test_library_is_sm100a_only():
    cuobjdump = os.path.join(os.path.dirname(_build.find_nvcc()), "cuobjdump")
    out = subprocess.run([cuobjdump, "-lelf", _cabi.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_host_only_entry_points(lib):
    assert lib.b200w_abi_version() == _cabi.ABI_VERSION
    assert lib.b200w_dwt_coeff_len(304, 6, 1) == 154
    assert lib.b200w_dwt_coeff_len(127, 6, 2) == 64
    assert lib.b200w_idwt_len(66, 6, 1) == 128 and lib.b200w_idwt_len(64, 6, 2) == 128
    assert lib.b200w_dwt_coeff_len(10, 4, 3) == _cabi.ERR_BAD_MODE      # 'constant' is rejected downstream
    assert lib.b200w_dwt_coeff_len(10, 4, 5) == _cabi.ERR_BAD_MODE      # 'replicate' too
    assert lib.b200w_status_string(_cabi.ERR_BAD_MODE) == b"Unkown pad type"
    assert lib.b200w_ssim_workspace_bytes(2, 3, 400, 400) > 0
    assert lib.b200w_ssim_workspace_bytes(0, 3, 400, 400) == 0


def test_argument_validation_happens_before_any_launch(lib):
    """These calls must return an error code without touching the (absent) GPU."""
    taps, n = _cabi.taps_array([0.5, 0.5])
    bogus = ctypes.c_void_p(256)
    args = (bogus, 64, 8, 1, 8, 8, taps, taps, n, taps, taps, n)
    assert lib.b200w_afb2d_f32(*args, 3, bogus, bogus, None) == _cabi.ERR_BAD_MODE
    assert lib.b200w_afb2d_f32(None, 64, 8, 1, 8, 8, taps, taps, n, taps, taps, n, 0, bogus, bogus, None) \
        == _cabi.ERR_NULL_POINTER
    assert lib.b200w_afb2d_f32(bogus, 64, 8, 1, 8, 8, taps, taps, 0, taps, taps, n, 0, bogus, bogus, None) \
        == _cabi.ERR_BAD_TAPS
    assert lib.b200w_afb2d_f32(bogus, 64, 8, 0, 8, 8, taps, taps, n, taps, taps, n, 0, bogus, bogus, None) \
        == _cabi.ERR_BAD_SHAPE
    t6, n6 = _cabi.taps_array([0.1] * 6)
    # reflect padding needs pad < dimension (torch F.pad rule): 4-sample axis, 6 taps -> pad 4
    assert lib.b200w_afb2d_f32(bogus, 16, 4, 1, 4, 4, t6, t6, n6, t6, t6, n6, 4, bogus, bogus, None) \
        == _cabi.ERR_REFLECT_PAD
    assert lib.b200w_afb2d_f32(bogus, 16, 4, 1, 4, 4, t6, t6, n6, t6, t6, n6, 2, bogus, bogus, None) \
        == _cabi.ERR_PER_TOO_SHORT
    # synthesis: out larger than the natural output
    assert lib.b200w_sfb2d_f32(bogus, 16, 4, None, 1, 4, 4, taps, taps, n, taps, taps, n, 0, bogus, 9, 8, None) \
        == _cabi.ERR_BAD_SHAPE
    w11, nw = _cabi.taps_array([1.0 / 11] * 11)
    w4, nw4 = _cabi.taps_array([0.25] * 4)
    assert lib.b200w_ssim_fwd_f32(bogus, bogus, 1, 1, 8, 8, w4, nw4, 1, 0, None, bogus, bogus, 1024, None) \
        == _cabi.ERR_BAD_WINDOW
    assert lib.b200w_ssim_fwd_f32(bogus, bogus, 1, 1, 8, 8, w11, nw, 1, 0, None, bogus, None, 0, None) \
        == _cabi.ERR_WORKSPACE
    with pytest.raises(ValueError, match="Unkown pad type: constant"):
        _cabi.check(_cabi.ERR_BAD_MODE, "constant")
    with pytest.raises(_cabi.B200WaveError):
        _cabi.check(_cabi.ERR_BAD_SHAPE)
