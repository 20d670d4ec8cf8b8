"""CPU oracle for the Gaussian-window SSIM -- TEST INFRASTRUCTURE ONLY.

Plain-numpy restatement of ``/root/reference/ssim.py`` (float64 by default):
the dense 2-D 11x11 window (fp32 outer product of the fp32-normalised 1-D
Gaussian, ``ssim.py:7-15``), five zero-padded "same" correlations
(``ssim.py:18-27``), the SSIM map (``ssim.py:29-32``) and its mean
(``ssim.py:34-37``).  The backward is the closed form of what autograd does
through ``_ssim``.

The reference has no test, fixture or known answer for SSIM, so this oracle is
pinned against outputs of the reference module itself executed in the build
container (``tests/golden/make_golden.py`` -> ``tests/golden/ssim_*.npz``) and
against torch autograd for the gradient.  Only tests, ``smoke()`` and the
``cpu_baseline`` leg of ``bench.py`` may import it.
"""
from math import exp

import numpy as np

C1 = 0.01 ** 2   # ssim.py:29
C2 = 0.03 ** 2   # ssim.py:30


def gaussian(window_size=11, sigma=1.5):
    """ssim.py:7-9 -- python-double exp, stored as fp32, normalised in fp32
    (torch's CPU sum of 11 floats is exactly rounded: accumulate in double)."""
    g = np.array([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2))
                  for x in range(window_size)], dtype=np.float32)
    return (g / np.float32(g.astype(np.float64).sum())).astype(np.float32)


def window2d(window_size=11):
    """ssim.py:11-15 -- fp32 outer product."""
    g = gaussian(window_size, 1.5)
    return np.outer(g, g).astype(np.float32)


def blur_same(x, win2d):
    """F.conv2d(x, window, padding=ws//2, groups=C), ssim.py:18 -- zero padded."""
    ws = win2d.shape[0]
    p = ws // 2
    H, W = x.shape[-2:]
    xp = np.pad(x, [(0, 0)] * (x.ndim - 2) + [(p, p), (p, p)])
    out = np.zeros_like(x)
    for i in range(ws):
        for j in range(ws):
            out += win2d[i, j] * xp[..., i:i + H, j:j + W]
    return out


def ssim_terms(img1, img2, window_size=11):
    w = window2d(window_size).astype(img1.dtype)
    mu1 = blur_same(img1, w)
    mu2 = blur_same(img2, w)
    s11 = blur_same(img1 * img1, w) - mu1 * mu1
    s22 = blur_same(img2 * img2, w) - mu2 * mu2
    s12 = blur_same(img1 * img2, w) - mu1 * mu2
    A1 = 2 * mu1 * mu2 + C1
    A2 = 2 * s12 + C2
    B1 = mu1 * mu1 + mu2 * mu2 + C1
    B2 = s11 + s22 + C2
    return w, mu1, mu2, A1, A2, B1, B2


def ssim_map(img1, img2, window_size=11):
    _, _, _, A1, A2, B1, B2 = ssim_terms(img1, img2, window_size)
    return (A1 * A2) / (B1 * B2)


def ssim(img1, img2, window_size=11, size_average=True):
    """_ssim, ssim.py:17-37."""
    m = ssim_map(img1, img2, window_size)
    if size_average:
        return m.mean()
    return m.mean(axis=(1, 2, 3))


def ssim_backward(img1, img2, grad_out=1.0, window_size=11, size_average=True):
    """Gradients (d/dimg1, d/dimg2) of ``ssim`` times ``grad_out`` (scalar, or
    shape (N,) when size_average=False)."""
    w, mu1, mu2, A1, A2, B1, B2 = ssim_terms(img1, img2, window_size)
    N, C, H, W = img1.shape
    S = (A1 * A2) / (B1 * B2)
    if size_average:
        g = np.full((N, 1, 1, 1), float(grad_out) / (N * C * H * W), dtype=img1.dtype)
    else:
        g = (np.asarray(grad_out, dtype=img1.dtype).reshape(N, 1, 1, 1)) / (C * H * W)
    dS_dmu1 = 2 * mu2 * (A2 - A1) / (B1 * B2) - 2 * mu1 * S * (1 / B1 - 1 / B2)
    dS_dmu2 = 2 * mu1 * (A2 - A1) / (B1 * B2) - 2 * mu2 * S * (1 / B1 - 1 / B2)
    dS_dE = -S / B2                 # w.r.t. E[x1^2] and E[x2^2]
    dS_dE12 = 2 * A1 / (B1 * B2)
    # the window is symmetric and the blur zero padded => self-adjoint
    d1 = blur_same(g * dS_dmu1, w) + 2 * img1 * blur_same(g * dS_dE, w) + img2 * blur_same(g * dS_dE12, w)
    d2 = blur_same(g * dS_dmu2, w) + 2 * img2 * blur_same(g * dS_dE, w) + img1 * blur_same(g * dS_dE12, w)
    return d1, d2
