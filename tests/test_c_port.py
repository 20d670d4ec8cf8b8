"""CPU: the C/OpenMP port (the CPU baseline bench.py times) against the numpy oracle and the golden vectors."""
import numpy as np
import pytest

from oracle import c_port, dwt_oracle, ssim_oracle
from helpers import case_filters, load_dwt_cases, load_ssim_cases, rel_err

DWT_CASES = load_dwt_cases()
SSIM_CASES = load_ssim_cases()


@pytest.mark.parametrize("case", DWT_CASES, ids=[c["id"] for c in DWT_CASES])
def test_port_matches_golden(case):
    hc, hr, gc, gr = case_filters(case)
    J, mode = case["J"], case["mode"]
    yl, yh, rec0, _ = c_port.dwt_roundtrip_fwd_bwd(case["x"], case["grec"], J, hc, hr, gc, gr, mode)
    assert rel_err(yl, case["yl"]) < 1e-5
    for j in range(J):
        assert rel_err(yh[j], case["yh%d" % j]) < 1e-5
    rec = c_port.dwt_inverse(case["yl"], [case["yh%d" % j] for j in range(J)], gc, gr, mode)
    assert rel_err(rec, case["recon"]) < 1e-5


def test_port_backward_chain_matches_oracle():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 1, 45, 52)).astype(np.float32)
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pywt_standin"))
    import pywt
    w = pywt.Wavelet("db3")
    h = dwt_oracle.prep_afb(w.dec_lo, w.dec_hi)
    g = (np.array(w.rec_lo), np.array(w.rec_hi))
    J, mode = 3, "symmetric"
    yl, yh = dwt_oracle.dwt_forward(x.astype(np.float64), J, h, h, mode)
    rec = dwt_oracle.dwt_inverse(yl, yh, g, g, mode)
    grec = rng.standard_normal(rec.shape).astype(np.float32)
    _, _, rec_c, dx_c = c_port.dwt_roundtrip_fwd_bwd(x, grec, J, h, h, g, g, mode)
    assert rel_err(rec_c, rec) < 1e-5
    # oracle chain
    dy = grec.astype(np.float64)
    dhs = []
    for j in range(J):
        dlow, dhigh = dwt_oracle.sfb2d_backward(dy, g[0], g[1], g[0], g[1], mode)
        dhs.append(dhigh)
        if j + 1 < J:
            tgt = yh[j + 1].shape[-2:]
            full = tuple(2 * t - 6 + 2 for t in tgt)
            pad = np.zeros(dlow.shape[:2] + full)
            pad[..., :dlow.shape[-2], :dlow.shape[-1]] = dlow
            dy = pad
    shapes = [x.shape[-2:]] + [t.shape[-2:] for t in yh[:-1]]
    d = dlow
    for j in reversed(range(J)):
        d = dwt_oracle.afb2d_backward(d, dhs[j], h[0], h[1], h[0], h[1], mode, shapes[j])
    assert rel_err(dx_c, d) < 1e-5


@pytest.mark.parametrize("case", SSIM_CASES, ids=[c["id"] for c in SSIM_CASES])
def test_port_ssim_matches_golden(case):
    w2 = ssim_oracle.window2d(11)
    val, d1, d2 = c_port.ssim(case["img1"], case["img2"], w2, case["size_average"], case["gout"], True, True)
    assert rel_err(np.atleast_1d(val), np.atleast_1d(case["val"])) < 1e-5
    if np.abs(case["d1"]).max() > 1e-12:
        assert rel_err(d1, case["d1"]) < 2e-5
        assert rel_err(d2, case["d2"]) < 2e-5
