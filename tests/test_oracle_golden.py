"""CPU: the numpy oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from oracle import dwt_oracle, freq_oracle, fsd_oracle, phase_oracle, ssim_oracle, swt_oracle, tv_oracle
from helpers import (load_dwt_cases, load_freq_cases, load_fsd_cases, load_ssim_cases, load_tv_cases, case_filters,
                     rel_err)

DWT_CASES = load_dwt_cases()
SSIM_CASES = load_ssim_cases()


@pytest.mark.parametrize("case", DWT_CASES, ids=[c["id"] for c in DWT_CASES])
def test_dwt_forward_inverse(case):
    h_col, h_row, g_col, g_row = case_filters(case)
    J, mode = case["J"], case["mode"]
    yl, yh = dwt_oracle.dwt_forward(case["x"], J, h_col, h_row, mode)
    assert rel_err(yl, case["yl"]) < 1e-12
    for j in range(J):
        assert yh[j].shape == case["yh%d" % j].shape
        assert rel_err(yh[j], case["yh%d" % j]) < 1e-12
    rec = dwt_oracle.dwt_inverse(case["yl"], [case["yh%d" % j] for j in range(J)], g_col, g_row, mode)
    assert rel_err(rec, case["recon"]) < 1e-12


@pytest.mark.parametrize("case", DWT_CASES, ids=[c["id"] for c in DWT_CASES])
def test_dwt_backward_chains(case):
    """AFB2D.backward / SFB2D.backward restatements chained over J levels == reference autograd."""
    h_col, h_row, g_col, g_row = case_filters(case)
    J, mode = case["J"], case["mode"]
    # forward input sizes per level
    shapes = [case["x"].shape[-2:]]
    for j in range(J - 1):
        shapes.append(case["yh%d" % j].shape[-2:])
    # d(yl, yh)/dx : coarsest level first
    d = case["gyl"]
    for j in reversed(range(J)):
        d = dwt_oracle.afb2d_backward(d, case["gyh%d" % j], h_col[0], h_col[1], h_row[0], h_row[1], mode, shapes[j])
    assert rel_err(d, case["dx"]) < 1e-12
    # d recon / d(yl, yh): finest level first; the 'unpad' crop back-propagates as zero padding
    dy = case["grec"]
    for j in range(J):
        dlow, dhigh = dwt_oracle.sfb2d_backward(dy, g_col[0], g_col[1], g_row[0], g_row[1], mode)
        assert rel_err(dhigh, case["dch%d" % j]) < 1e-12
        if j + 1 < J:
            tgt = case["yh%d" % (j + 1)].shape[-2:]   # size of the ll that level j+1 reconstructed
            full = (dwt_oracle_len(tgt[0], len(g_row[0]), mode), dwt_oracle_len(tgt[1], len(g_col[0]), mode))
            pad = np.zeros(dlow.shape[:2] + full, dtype=dlow.dtype)
            pad[..., :dlow.shape[-2], :dlow.shape[-1]] = dlow
            dy = pad
        else:
            assert rel_err(dlow, case["dcl"]) < 1e-12


def dwt_oracle_len(m, L, mode):
    return 2 * m if mode in ("per", "periodization") else 2 * m - L + 2


@pytest.mark.parametrize("case", SSIM_CASES, ids=[c["id"] for c in SSIM_CASES])
def test_ssim_value_and_grads(case):
    val = ssim_oracle.ssim(case["img1"], case["img2"], 11, case["size_average"])
    assert rel_err(np.atleast_1d(val), np.atleast_1d(case["val"])) < 1e-12
    d1, d2 = ssim_oracle.ssim_backward(case["img1"], case["img2"], case["gout"], 11, case["size_average"])
    if np.abs(case["d1"]).max() > 1e-12:
        assert rel_err(d1, case["d1"]) < 1e-10
        assert rel_err(d2, case["d2"]) < 1e-10
    else:  # ssim(x, x): gradient is zero up to rounding
        assert np.abs(d1).max() < 1e-12


def test_known_answers():
    # docs/dwt.rst:64-75: db3, zero, J=3 on 64x64 -> yl 12x12, yh 34, 19, 12
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pywt_standin"))
    import pywt
    w = pywt.Wavelet("db3")
    h = dwt_oracle.prep_afb(w.dec_lo, w.dec_hi)
    x = np.random.default_rng(0).standard_normal((1, 1, 64, 64))
    yl, yh = dwt_oracle.dwt_forward(x, 3, h, h, "zero")
    assert yl.shape[-2:] == (12, 12)
    assert [t.shape[-1] for t in yh] == [34, 19, 12]
    # haar on a constant image: LL = 2c, details 0
    w = pywt.Wavelet("haar")
    h = dwt_oracle.prep_afb(w.dec_lo, w.dec_hi)
    yl, yh = dwt_oracle.dwt_forward(np.full((1, 1, 8, 8), 3.0), 1, h, h, "zero")
    assert np.allclose(yl, 6.0) and np.allclose(yh[0], 0.0)
    # ssim(x, x) == 1
    a = np.random.default_rng(1).random((1, 2, 20, 20))
    assert abs(ssim_oracle.ssim(a, a) - 1.0) < 1e-12


FREQ_CASES = load_freq_cases()


@pytest.mark.parametrize("case", FREQ_CASES, ids=[c["id"] for c in FREQ_CASES])
def test_freq_split_oracle_vs_reference(case):
    """utils.high_pass / utils.low_pass of the unmodified reference (fp32 FFT) vs the float64 restatement."""
    fn = freq_oracle.high_pass if case["highpass"] else freq_oracle.low_pass
    out = fn(case["x"], case["radius"])
    assert out.shape == case["y"].shape
    assert rel_err(out, case["y"]) < 3e-6        # the reference itself is fp32
    assert (out >= 0).all() if case["highpass"] else (out <= 0).all()      # low_pass returns -|.| (utils.py:117)


def test_freq_split_backward_is_the_gradient():
    """oracle backward (self-adjoint filter, sgn through abs) == finite differences of the oracle forward."""
    rng = np.random.default_rng(3)
    x = rng.random((12, 10))
    g = rng.standard_normal((12, 10))
    for hp, sign in [(True, 1.0), (False, -1.0)]:
        d = freq_oracle.split_backward(x, g, 3, hp, sign)
        eps = 1e-6
        for (i, j) in [(0, 0), (5, 7), (11, 9)]:
            xp, xm = x.copy(), x.copy()
            xp[i, j] += eps
            xm[i, j] -= eps
            fd = (g * (freq_oracle.split(xp, 3, hp, sign) - freq_oracle.split(xm, 3, hp, sign))).sum() / (2 * eps)
            assert abs(fd - d[i, j]) < 1e-6 * max(1.0, abs(fd))


FSD_CASES = load_fsd_cases()


@pytest.mark.parametrize("case", FSD_CASES, ids=[c["id"] for c in FSD_CASES])
def test_filter_wavelet_oracle_vs_reference(case):
    """fsd_oracle vs outputs and input gradients of FS_DiscriminatorA/B.filter_wavelet (model.py:166-179, 222-235)."""
    got = fsd_oracle.filter_wavelet(case["x"], case["cs"], case["norm"], case["variant"])
    assert len(got) == len(case["y"])
    for a, b in zip(got, case["y"]):
        assert a.shape == b.shape and rel_err(a, b) < 1e-12
    dx = fsd_oracle.filter_wavelet_backward(case["g"], case["x"].shape, case["cs"], case["norm"], case["variant"])
    assert rel_err(dx, case["dx"]) < 1e-12


def test_filter_wavelet_host_logic_without_gpu():
    """Format validation happens on the host; a CPU tensor is refused (no CPU fallback)."""
    import torch
    from b200wave import fsd
    with pytest.raises(NotImplementedError, match="not recognized"):
        fsd.WaveletFilter(cs="avg")
    filt = fsd.WaveletFilter(cs="cat", variant="B")
    assert sorted(k for k in filt.state_dict()) == ["h0_col", "h0_row", "h1_col", "h1_row"]   # DWTForward's names
    with pytest.raises(RuntimeError, match="CUDA-only"):
        filt(torch.rand(1, 1, 8, 8))


TV_CASES = load_tv_cases()


@pytest.mark.parametrize("case", TV_CASES, ids=[c["id"] for c in TV_CASES])
def test_tv_oracle_vs_reference(case):
    """tv_oracle vs the value and input gradient of the reference's TVLoss (model.py:17-33)."""
    assert abs(tv_oracle.tv_loss(case["x"], case["weight"]) - case["loss"]) <= 1e-12 * abs(case["loss"])
    assert rel_err(tv_oracle.tv_loss_backward(case["x"], 1.0, case["weight"]), case["dx"]) < 1e-12


def test_tv_loss_host_logic_without_gpu():
    import torch
    from b200wave.losses import TVLoss
    crit = TVLoss(TVLoss_weight=0.5)
    assert crit.TVLoss_weight == 0.5 and crit._tensor_size(torch.zeros(2, 3, 4, 5)) == 60
    with pytest.raises(RuntimeError, match="CUDA-only"):
        crit(torch.rand(1, 1, 8, 8))


PHASE_CASES = __import__("helpers").load_phase_cases()


@pytest.mark.parametrize("case", PHASE_CASES, ids=[c["id"] for c in PHASE_CASES])
def test_phase_oracle_vs_reference(case):
    """phase_oracle vs the value and both input gradients of the reference's phase_consistency_loss (model.py:36-58)."""
    assert abs(phase_oracle.phase_consistency_loss(case["x"], case["y"]) - case["loss"]) <= 1e-11
    assert rel_err(phase_oracle.phase_consistency_grad(case["x"], case["y"], 0), case["dx"]) < 1e-9
    assert rel_err(phase_oracle.phase_consistency_grad(case["x"], case["y"], 1), case["dy"]) < 1e-9


def test_phase_loss_host_logic_without_gpu():
    """No CPU fallback: CPU tensors are refused loudly, shapes are checked before any device work."""
    import torch
    import b200wave
    crit = b200wave.phase_consistency_loss()
    with pytest.raises(RuntimeError, match="CUDA-only"):
        crit(torch.rand(1, 1, 8, 8), torch.rand(1, 1, 8, 8))


DWT1D_CASES = __import__("helpers").load_dwt1d_cases()


@pytest.mark.parametrize("case", DWT1D_CASES, ids=[c["id"] for c in DWT1D_CASES])
def test_dwt1d_oracle_vs_reference(case):
    """dwt_oracle.afb1d / sfb1d chained as DWT1DForward / DWT1DInverse do (transform1d.py:37-62, 93-115) and as their
    backward passes do (lowlevel.py:407-424, 732-743) vs the unmodified reference's outputs and gradients."""
    J, mode = case["J"], case["mode"]
    h0, h1, g0, g1 = case["h0"], case["h1"], case["g0"], case["g1"]
    lo = case["x"]
    lens = []
    for j in range(J):
        lens.append(lo.shape[-1])
        lo, hi = dwt_oracle.afb1d(lo, h0, h1, mode)
        assert rel_err(hi, case["yh%d" % j]) < 1e-12
    assert rel_err(lo, case["yl"]) < 1e-12
    # AFB1D.backward chain: synthesis with the analysis taps, cropped to each level's input length
    d = case["gl"]
    for j in range(J - 1, -1, -1):
        d = dwt_oracle.sfb1d(d, case["gh%d" % j], h0, h1, mode)[..., :lens[j]]
    assert rel_err(d, case["dx"]) < 1e-12
    # inverse + SFB1D.backward chain
    r = case["yl"]
    unpads = []
    for j in range(J - 1, -1, -1):
        hi = case["yh%d" % j]
        unpads.append(r.shape[-1] > hi.shape[-1])
        if unpads[-1]:
            r = r[..., :-1]
        r = dwt_oracle.sfb1d(r, hi, g0, g1, mode)
    assert rel_err(r, case["rec"]) < 1e-12
    g = case["gy"]
    for j in range(J):
        g, dhi = dwt_oracle.afb1d(g, g0, g1, mode)
        assert rel_err(dhi, case["dch%d" % j]) < 1e-12
        if unpads[J - 1 - j]:                       # autograd through x0[..., :-1]: a zero is appended
            g = np.concatenate([g, np.zeros(g.shape[:-1] + (1,))], axis=-1)
    assert rel_err(g, case["dcl"]) < 1e-12


SWT_CASES = __import__("helpers").load_swt_cases()


@pytest.mark.parametrize("case", SWT_CASES, ids=[c["id"] for c in SWT_CASES])
def test_swt_oracle_vs_reference(case):
    """swt_oracle.afb2d_atrous and its adjoint vs the unmodified reference function and autograd through it
    (lowlevel.py:475-521)."""
    fl = [case[k] for k in ("h0_col", "h1_col", "h0_row", "h1_row")]
    assert rel_err(swt_oracle.afb2d_atrous(case["x"], *fl, case["mode"], case["dilation"]), case["y"]) < 1e-12
    assert rel_err(swt_oracle.afb2d_atrous_backward(case["gy"], *fl, case["mode"], case["dilation"]), case["dx"]) < 1e-12


def test_swt_host_logic_without_gpu():
    """What the reference's SWTForward does with its default mode (recorded in the golden file) is what ours does:
    ValueError("Unkown pad type: periodization") before any device work; CPU tensors are refused."""
    import torch
    import b200wave
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "swt_cases.npz"))
    assert str(z["default_mode_error"]).startswith("ValueError: Unkown pad type: periodization")
    with pytest.raises(ValueError, match="Unkown pad type: periodization"):
        b200wave.SWTForward()(torch.zeros(1, 1, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        b200wave.SWTForward(mode="zero")(torch.zeros(1, 1, 8, 8))
