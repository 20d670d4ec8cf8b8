"""CPU: host-side mirror of the reference interface (names, buffers, mode codes, errors)."""
import numpy as np
import pytest
import torch

import b200wave
from b200wave import lowlevel, ops
from b200wave.sharding import shard_range


def test_mode_codes_roundtrip():
    # pw/dwt/lowlevel.py:274-309
    codes = {"zero": 0, "symmetric": 1, "periodization": 2, "constant": 3, "reflect": 4, "replicate": 5,
             "periodic": 6}
    for name, code in codes.items():
        assert lowlevel.mode_to_int(name) == code
        assert lowlevel.int_to_mode(code) == name
    assert lowlevel.mode_to_int("per") == 2
    with pytest.raises(ValueError, match="Unkown pad type: bogus"):
        lowlevel.mode_to_int("bogus")
    with pytest.raises(ValueError, match="Unkown pad type: 9"):
        lowlevel.int_to_mode(9)


def test_module_buffers_match_reference_layout():
    xfm = b200wave.DWTForward(J=3, wave="db3", mode="symmetric")
    ifm = b200wave.DWTInverse(wave="db3", mode="symmetric")
    assert sorted(dict(xfm.named_buffers())) == ["h0_col", "h0_row", "h1_col", "h1_row"]
    assert sorted(dict(ifm.named_buffers())) == ["g0_col", "g0_row", "g1_col", "g1_row"]
    assert xfm.h0_col.shape == (1, 1, 6, 1) and xfm.h1_row.shape == (1, 1, 1, 6)
    assert ifm.g0_col.shape == (1, 1, 6, 1) and ifm.g1_row.shape == (1, 1, 1, 6)
    assert xfm.h0_col.dtype == torch.float32
    w = b200wave.Wavelet("db3")
    # analysis buffers are time reversed, synthesis buffers are not (lowlevel.py:918-922, 970-971)
    assert np.allclose(xfm.h0_col.flatten().numpy(), np.array(w.dec_lo[::-1], dtype=np.float32))
    assert np.allclose(ifm.g1_col.flatten().numpy(), np.array(w.rec_hi, dtype=np.float32))
    assert xfm.J == 3 and xfm.mode == "symmetric" and ifm.mode == "symmetric"
    d = b200wave.DWTForward()
    assert d.J == 1 and d.mode == "zero" and d.h0_col.numel() == 2
    assert b200wave.DWT is b200wave.DWTForward and b200wave.IDWT2D is b200wave.DWTInverse


def test_wave_tuple_forms():
    a, b = b200wave.Wavelet("haar"), b200wave.Wavelet("db2")
    m = b200wave.DWTForward(wave=(a.dec_lo, a.dec_hi, b.dec_lo, b.dec_hi))
    assert m.h0_col.shape == (1, 1, 2, 1) and m.h0_row.shape == (1, 1, 1, 4)
    m2 = b200wave.DWTForward(wave=(np.array(b.dec_lo), np.array(b.dec_hi)))
    assert torch.equal(m2.h0_col.flatten(), m2.h0_row.flatten())
    with pytest.raises(ValueError):
        b200wave.DWTForward(wave=(a.dec_lo,) * 3)


def test_default_dtype_double_buffers():
    old = torch.get_default_dtype()
    try:
        torch.set_default_dtype(torch.float64)
        assert b200wave.DWTForward(wave="db2").h0_col.dtype == torch.float64
    finally:
        torch.set_default_dtype(old)


def test_host_taps_cache():
    t = torch.tensor([1.0, 2.0, 3.0]).reshape(1, 1, 3, 1)
    a = lowlevel.host_taps(t)
    assert a == (1.0, 2.0, 3.0) and lowlevel.host_taps(t) is a
    t.mul_(2)                                   # in-place change bumps the version counter
    assert lowlevel.host_taps(t) == (2.0, 4.0, 6.0)
    assert lowlevel._as_taps([1.0, 2.0], True) == (2.0, 1.0) and lowlevel._as_taps([1.0, 2.0], False) == (1.0, 2.0)


def test_shape_arithmetic():
    # cfg2 db3/symmetric 304 -> 154 -> 79 -> 42 ; db8 1024 -> 519 -> 267 -> 141 -> 78 -> 46
    n, sizes = 304, []
    for _ in range(3):
        n = ops.coeff_len(n, 6, 1)
        sizes.append(n)
    assert sizes == [154, 79, 42]
    n, sizes = 1024, []
    for _ in range(5):
        n = ops.coeff_len(n, 16, 0)
        sizes.append(n)
    assert sizes == [519, 267, 141, 78, 46]
    assert ops.coeff_len(127, 6, 2) == 64 and ops.idwt_len(64, 6, 2) == 128 and ops.idwt_len(66, 6, 1) == 128


def test_no_cpu_fallback():
    xfm = b200wave.DWTForward(J=1, wave="haar", mode="reflect")
    with pytest.raises(RuntimeError, match="CUDA-only"):
        xfm(torch.zeros(1, 1, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        b200wave.DWTInverse()((torch.zeros(1, 1, 4, 4), [torch.zeros(1, 1, 3, 4, 4)]))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        b200wave.SSIM()(torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16))
    with pytest.raises(ValueError, match="Unkown pad type"):
        b200wave.DWTForward(mode="nonsense")(torch.zeros(1, 1, 8, 8))


def test_fake_tensor_shapes():
    x = torch.empty(4, 2, 37, 50, device="meta")
    w = [0.1] * 6
    low, highs = torch.ops.b200wave.afb2d(x, w, w, w, w, 1)
    assert low.shape == (4, 2, 21, 27) and highs.shape == (4, 2, 3, 21, 27)
    y = torch.ops.b200wave.sfb2d(low, highs, w, w, w, w, 1, -1, -1)
    assert y.shape == (4, 2, 38, 50)
    val, maps = torch.ops.b200wave.ssim_fwd(x, x, [0.2] * 5, False, 3)
    assert val.shape == (4,) and maps.shape == (3, 4, 2, 37, 50)


def test_ssim_window_matches_reference_values():
    # SURVEY 8a-a9 taps
    import importlib
    mod = importlib.import_module("b200wave.ssim")   # (b200wave.ssim the attribute is the function)
    g = mod.gaussian(11, 1.5)
    ref = [0.0010283801, 0.0075987582, 0.036000773, 0.10936069, 0.21300553, 0.26601171]
    assert np.allclose(g[:6].numpy(), ref, rtol=2e-7)
    w = mod.create_window(11, 3)
    assert w.shape == (3, 1, 11, 11) and abs(float(w[0, 0].sum()) - 0.99999988) < 2e-7
    m = mod.SSIM()
    assert m.window_size == 11 and m.size_average is True and m.channel == 1


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_compat_aliases_resolve_reference_imports():
    import sys
    import b200wave.compat
    saved = {k: sys.modules.get(k) for k in ("pytorch_wavelets", "pytorch_wavelets.dwt", "ssim",
                                              "pytorch_wavelets.dwt.lowlevel", "pytorch_wavelets.dwt.transform2d")}
    try:
        b200wave.compat.install(force=True)
        from pytorch_wavelets import DWTForward, DWTInverse, IDWT
        import pytorch_wavelets.dwt.lowlevel as ll
        import ssim as ssim_alias
        assert DWTForward is b200wave.DWTForward and IDWT is DWTInverse
        assert ll.AFB2D is lowlevel.AFB2D and ll.mode_to_int("reflect") == 4
        assert ssim_alias.SSIM is b200wave.SSIM and callable(ssim_alias.ssim)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_ssim_window_sizes_outside_the_kernel_range_are_refused_at_construction():
    """the fused kernel holds an odd window <= 11; SSIM(window_size=8) / 15 raise a ValueError naming the
    supported range instead of failing later inside the C ABI."""
    import b200wave
    for w in (8, 15, 0, 12):
        with pytest.raises(ValueError, match="odd window sizes 1..11"):
            b200wave.SSIM(window_size=w)
    assert b200wave.SSIM(window_size=5).window.shape == (1, 1, 5, 5)


def test_backwards_are_marked_once_differentiable():
    """the hand-written backwards call kernels without autograd formulas, so they are declared
    once-differentiable (a double backward raises instead of silently returning no gradient)."""
    import inspect
    from b200wave import ops, losses
    for fn in (ops.DWT2Function, ops.IDWT2Function, losses._TV, losses._PhaseCos):
        assert "once_differentiable" in inspect.getsource(fn)


def test_tools_and_entry_points_parse():
    """bench.py, __graft_entry__.py and every script under tools/ are syntactically valid (they only run on the GPU box)."""
    import ast
    import glob
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "tools", "*.py")) + [os.path.join(root, "bench.py"),
                                                              os.path.join(root, "__graft_entry__.py")]
    assert len(files) > 10
    for f in files:
        with open(f) as fh:
            ast.parse(fh.read(), filename=f)
