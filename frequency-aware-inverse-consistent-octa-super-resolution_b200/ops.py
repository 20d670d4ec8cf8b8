"""``torch.library`` custom ops ``b200wave::*`` -- the thin layer between the
Python modules and the C ABI.

Each op takes device tensors plus host scalars / tap lists, allocates its
outputs with torch (the library never allocates) and launches on torch's
current stream, so the ops are asynchronous and CUDA-graph capturable.
CUDA-only: there is deliberately no CPU implementation.

Autograd follows the reference's hand-written backward passes, not the
mathematical adjoint:

* ``afb2d`` backward  = ``sfb2d`` with the same (reversed) analysis taps, cropped
  to the input size                       (pw/dwt/lowlevel.py:349-365)
* ``sfb2d`` backward  = ``afb2d`` of ``dy`` with the synthesis taps as
  correlation kernels                     (pw/dwt/lowlevel.py:682-694)
* ``ssim_fwd`` backward = ``ssim_bwd`` (closed form of autograd through ssim.py:17-37)
"""
import ctypes

import torch

from . import _cabi

_LIB = torch.library.Library("b200wave", "DEF")

_LIB.define("afb2d(Tensor x, float[] w_lo, float[] w_hi, float[] h_lo, float[] h_hi, int mode) -> (Tensor, Tensor)")
_LIB.define("sfb2d(Tensor low, Tensor? highs, float[] w_lo, float[] w_hi, float[] h_lo, float[] h_hi, int mode, "
            "int out_h, int out_w) -> Tensor")
_LIB.define("ssim_fwd(Tensor img1, Tensor img2, float[] win, bool size_average, int n_maps) -> (Tensor, Tensor)")
_LIB.define("ssim_bwd(Tensor img1, Tensor img2, Tensor maps, Tensor grad_out, float[] win, bool size_average, "
            "bool need_d2) -> (Tensor, Tensor)")

_INT_TO_MODE = {0: "zero", 1: "symmetric", 2: "periodization", 3: "constant", 4: "reflect", 5: "replicate",
                6: "periodic"}


def _mode_name(mode):
    return _INT_TO_MODE.get(int(mode), mode)


def coeff_len(n, l, mode):
    """pywt.dwt_coeff_len as used at pw/dwt/lowlevel.py:153 (host arithmetic only)."""
    return (n + 1) // 2 if mode == 2 else (n + l - 1) // 2


def idwt_len(m, l, mode):
    return 2 * m if mode == 2 else 2 * m - l + 2


def _check_mode(mode):
    if mode not in (0, 1, 2, 4, 6):
        raise ValueError("Unkown pad type: {}".format(_mode_name(mode)))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda_f32(t, name):
    if not t.is_cuda:
        raise RuntimeError("b200wave::%s is CUDA-only (sm_100a); there is no CPU fallback -- got a %s tensor"
                           % (name, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError("b200wave::%s: expected scalar type Float but found %s" % (name, t.dtype))


def _planes_view(t):
    """(N,C,H,W) tensor -> (tensor to keep alive, plane_stride, row_stride); copies only when the
    layout cannot be expressed as planes with unit column stride."""
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    ok = (sw == 1 or w == 1) and (n == 1 or c == 1 or sn == c * sc)
    if not ok:
        t = t.contiguous()
        sn, sc, sh, sw = t.stride()
    plane_stride = sc if c > 1 else (sn if n > 1 else h * sh)
    return t, plane_stride, sh


# ------------------------------------------------------------------------------------------- afb2d
def _afb2d_cuda(x, w_lo, w_hi, h_lo, h_hi, mode):
    _check_mode(mode)
    _require_cuda_f32(x, "afb2d")
    if x.dim() != 4:
        raise IndexError("b200wave::afb2d expects a 4-D (N, C, H, W) tensor, got %d-D" % x.dim())
    lib = _cabi.load()
    N, C, H, W = x.shape
    Lw, Lh = len(w_lo), len(h_lo)
    if len(w_hi) != Lw or len(h_hi) != Lh:
        raise RuntimeError("low- and high-pass filters must have the same length along an axis")
    Ho, Wo = coeff_len(H, Lh, mode), coeff_len(W, Lw, mode)
    low = torch.empty((N, C, Ho, Wo), device=x.device, dtype=torch.float32)
    highs = torch.empty((N, C, 3, Ho, Wo), device=x.device, dtype=torch.float32)
    if low.numel() == 0:
        return low, highs
    xk, ps, rs = _planes_view(x)
    a_wl, _ = _cabi.taps_array(w_lo)
    a_wh, _ = _cabi.taps_array(w_hi)
    a_hl, _ = _cabi.taps_array(h_lo)
    a_hh, _ = _cabi.taps_array(h_hi)
    with torch.cuda.device(x.device):
        rc = lib.b200w_afb2d_f32(xk.data_ptr(), ps, rs, N * C, H, W, a_wl, a_wh, Lw, a_hl, a_hh, Lh, int(mode),
                                 low.data_ptr(), highs.data_ptr(), _stream())
    _cabi.check(rc, _mode_name(mode))
    return low, highs


def _afb2d_fake(x, w_lo, w_hi, h_lo, h_hi, mode):
    N, C, H, W = x.shape
    Ho, Wo = coeff_len(H, len(h_lo), mode), coeff_len(W, len(w_lo), mode)
    return x.new_empty((N, C, Ho, Wo)), x.new_empty((N, C, 3, Ho, Wo))


# ------------------------------------------------------------------------------------------- sfb2d
def _sfb2d_cuda(low, highs, w_lo, w_hi, h_lo, h_hi, mode, out_h, out_w):
    _check_mode(mode)
    _require_cuda_f32(low, "sfb2d")
    if low.dim() != 4:
        raise IndexError("b200wave::sfb2d expects a 4-D (N, C, h, w) lowpass tensor, got %d-D" % low.dim())
    lib = _cabi.load()
    N, C, h, w = low.shape
    Lw, Lh = len(w_lo), len(h_lo)
    if len(w_hi) != Lw or len(h_hi) != Lh:
        raise RuntimeError("low- and high-pass filters must have the same length along an axis")
    if highs is not None:
        _require_cuda_f32(highs, "sfb2d")
        if tuple(highs.shape) != (N, C, 3, h, w):
            raise RuntimeError("b200wave::sfb2d: highs must have shape %s, got %s"
                               % ((N, C, 3, h, w), tuple(highs.shape)))
        highs = highs.contiguous()
    full_h, full_w = idwt_len(h, Lh, mode), idwt_len(w, Lw, mode)
    oh = full_h if out_h < 0 else out_h
    ow = full_w if out_w < 0 else out_w
    y = torch.empty((N, C, oh, ow), device=low.device, dtype=torch.float32)
    if y.numel() == 0:
        return y
    lk, ps, rs = _planes_view(low)
    a_wl, _ = _cabi.taps_array(w_lo)
    a_wh, _ = _cabi.taps_array(w_hi)
    a_hl, _ = _cabi.taps_array(h_lo)
    a_hh, _ = _cabi.taps_array(h_hi)
    with torch.cuda.device(low.device):
        rc = lib.b200w_sfb2d_f32(lk.data_ptr(), ps, rs, None if highs is None else highs.data_ptr(), N * C, h, w,
                                 a_wl, a_wh, Lw, a_hl, a_hh, Lh, int(mode), y.data_ptr(), oh, ow, _stream())
    _cabi.check(rc, _mode_name(mode))
    return y


def _sfb2d_fake(low, highs, w_lo, w_hi, h_lo, h_hi, mode, out_h, out_w):
    N, C, h, w = low.shape
    oh = idwt_len(h, len(h_lo), mode) if out_h < 0 else out_h
    ow = idwt_len(w, len(w_lo), mode) if out_w < 0 else out_w
    return low.new_empty((N, C, oh, ow))


# ------------------------------------------------------------------------------------------- autograd (DWT)
def _afb2d_setup(ctx, inputs, output):
    x, w_lo, w_hi, h_lo, h_hi, mode = inputs
    ctx.taps = (w_lo, w_hi, h_lo, h_hi)
    ctx.mode = mode
    ctx.in_hw = (x.shape[-2], x.shape[-1])   # AFB2D saves only the shape, not x (lowlevel.py:337-338)
    ctx.set_materialize_grads(False)


def _afb2d_backward(ctx, dlow, dhighs):
    dx = None
    if ctx.needs_input_grad[0]:
        if dlow is None and dhighs is None:
            return None, None, None, None, None, None
        if dlow is None:
            n, c, _, h, w = dhighs.shape
            dlow = dhighs.new_zeros((n, c, h, w))
        w_lo, w_hi, h_lo, h_hi = ctx.taps
        H, W = ctx.in_hw
        # synthesis with the saved analysis taps, cropped to the forward input (lowlevel.py:356-364)
        dx = torch.ops.b200wave.sfb2d(dlow, dhighs, w_lo, w_hi, h_lo, h_hi, ctx.mode, H, W)
    return dx, None, None, None, None, None


def _sfb2d_setup(ctx, inputs, output):
    low, highs, w_lo, w_hi, h_lo, h_hi, mode, out_h, out_w = inputs
    ctx.taps = (w_lo, w_hi, h_lo, h_hi)
    ctx.mode = mode
    ctx.has_highs = highs is not None
    ctx.cropped = tuple(output.shape[-2:]) != (idwt_len(low.shape[-2], len(h_lo), mode),
                                               idwt_len(low.shape[-1], len(w_lo), mode))


def _sfb2d_backward(ctx, dy):
    dlow = dhighs = None
    need_low = ctx.needs_input_grad[0]
    need_high = ctx.has_highs and ctx.needs_input_grad[1]
    if need_low or need_high:
        if ctx.cropped:
            raise RuntimeError("b200wave::sfb2d: backward through a cropped synthesis is not defined "
                               "(the crop only exists inside AFB2D.backward)")
        w_lo, w_hi, h_lo, h_hi = ctx.taps
        # analysis of dy with the un-reversed synthesis taps as correlators (lowlevel.py:687-693)
        dlow, dhighs = torch.ops.b200wave.afb2d(dy, w_lo, w_hi, h_lo, h_hi, ctx.mode)
        if not need_low:
            dlow = None
        if not need_high:
            dhighs = None
    return dlow, dhighs, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------- ssim
def _ssim_common(img1, img2, win, name):
    _require_cuda_f32(img1, name)
    _require_cuda_f32(img2, name)
    if img1.dim() != 4 or img1.shape != img2.shape:
        raise RuntimeError("b200wave::%s expects two (N, C, H, W) tensors of equal shape, got %s and %s"
                           % (name, tuple(img1.shape), tuple(img2.shape)))
    if len(win) % 2 == 0 or len(win) > _cabi.SSIM_MAX_WINDOW:
        raise RuntimeError("b200wave::%s: window size must be odd and <= %d, got %d"
                           % (name, _cabi.SSIM_MAX_WINDOW, len(win)))


def _ssim_fwd_cuda(img1, img2, win, size_average, n_maps):
    _ssim_common(img1, img2, win, "ssim_fwd")
    lib = _cabi.load()
    img1, img2 = img1.contiguous(), img2.contiguous()
    N, C, H, W = img1.shape
    out = torch.empty(() if size_average else (N,), device=img1.device, dtype=torch.float32)
    maps = torch.empty((n_maps, N, C, H, W), device=img1.device, dtype=torch.float32)
    ws_bytes = lib.b200w_ssim_workspace_bytes(N, C, H, W)
    work = torch.empty((max(ws_bytes, 4) // 4,), device=img1.device, dtype=torch.float32)
    a_win, nwin = _cabi.taps_array(win)
    with torch.cuda.device(img1.device):
        rc = lib.b200w_ssim_fwd_f32(img1.data_ptr(), img2.data_ptr(), N, C, H, W, a_win, nwin, int(size_average),
                                    int(n_maps), maps.data_ptr() if n_maps else None, out.data_ptr(),
                                    work.data_ptr(), ws_bytes, _stream())
    _cabi.check(rc)
    return out, maps


def _ssim_fwd_fake(img1, img2, win, size_average, n_maps):
    N, C, H, W = img1.shape
    return img1.new_empty(() if size_average else (N,)), img1.new_empty((n_maps, N, C, H, W))


def _ssim_bwd_cuda(img1, img2, maps, grad_out, win, size_average, need_d2):
    _ssim_common(img1, img2, win, "ssim_bwd")
    lib = _cabi.load()
    img1, img2, maps = img1.contiguous(), img2.contiguous(), maps.contiguous()
    N, C, H, W = img1.shape
    n_maps = maps.shape[0]
    if n_maps not in (3, 4) or (need_d2 and n_maps != 4):
        raise RuntimeError("b200wave::ssim_bwd: forward saved %d derivative maps, not enough for this backward"
                           % n_maps)
    g = grad_out.to(device=img1.device, dtype=torch.float32).contiguous()
    if g.numel() != (1 if size_average else N):
        raise RuntimeError("b200wave::ssim_bwd: grad_out has %d elements" % g.numel())
    d1 = torch.empty_like(img1)
    d2 = torch.empty_like(img2) if need_d2 else torch.empty((0,), device=img1.device, dtype=torch.float32)
    a_win, nwin = _cabi.taps_array(win)
    with torch.cuda.device(img1.device):
        rc = lib.b200w_ssim_bwd_f32(img1.data_ptr(), img2.data_ptr(), maps.data_ptr(), n_maps, g.data_ptr(),
                                    N, C, H, W, a_win, nwin, int(size_average), d1.data_ptr(),
                                    d2.data_ptr() if need_d2 else None, _stream())
    _cabi.check(rc)
    return d1, d2


def _ssim_bwd_fake(img1, img2, maps, grad_out, win, size_average, need_d2):
    return torch.empty_like(img1), (torch.empty_like(img2) if need_d2 else img1.new_empty((0,)))


def _ssim_setup(ctx, inputs, output):
    img1, img2, win, size_average, n_maps = inputs
    _, maps = output
    ctx.save_for_backward(img1, img2, maps)
    ctx.win = win
    ctx.size_average = size_average
    ctx.n_maps = n_maps
    ctx.set_materialize_grads(False)


def _ssim_backward(ctx, dval, dmaps):
    need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    if not (need1 or need2) or dval is None:
        return None, None, None, None, None
    if ctx.n_maps < 3 or (need2 and ctx.n_maps < 4):
        raise RuntimeError("b200wave::ssim_fwd was called with n_maps=%d, which does not support the requested "
                           "gradient (use 3 for d/dimg1, 4 for both)" % ctx.n_maps)
    img1, img2, maps = ctx.saved_tensors
    d1, d2 = torch.ops.b200wave.ssim_bwd(img1, img2, maps, dval, ctx.win, ctx.size_average, bool(need2))
    return (d1 if need1 else None), (d2 if need2 else None), None, None, None


_LIB.impl("afb2d", _afb2d_cuda, "CUDA")
_LIB.impl("sfb2d", _sfb2d_cuda, "CUDA")
_LIB.impl("ssim_fwd", _ssim_fwd_cuda, "CUDA")
_LIB.impl("ssim_bwd", _ssim_bwd_cuda, "CUDA")


def _cpu_refuse(name):
    def impl(*args, **kwargs):
        raise RuntimeError("b200wave::%s is CUDA-only (sm_100a): there is no CPU fallback. Move the tensors to "
                           "a B200 (`.cuda()`)." % name)
    return impl


for _name in ("afb2d", "sfb2d", "ssim_fwd", "ssim_bwd"):
    _LIB.impl(_name, _cpu_refuse(_name), "CPU")

torch.library.register_fake("b200wave::afb2d", _afb2d_fake, lib=_LIB)
torch.library.register_fake("b200wave::sfb2d", _sfb2d_fake, lib=_LIB)
torch.library.register_fake("b200wave::ssim_fwd", _ssim_fwd_fake, lib=_LIB)
torch.library.register_fake("b200wave::ssim_bwd", _ssim_bwd_fake, lib=_LIB)
torch.library.register_autograd("b200wave::afb2d", _afb2d_backward, setup_context=_afb2d_setup, lib=_LIB)
torch.library.register_autograd("b200wave::sfb2d", _sfb2d_backward, setup_context=_sfb2d_setup, lib=_LIB)
torch.library.register_autograd("b200wave::ssim_fwd", _ssim_backward, setup_context=_ssim_setup, lib=_LIB)

afb2d = torch.ops.b200wave.afb2d
sfb2d = torch.ops.b200wave.sfb2d
ssim_fwd = torch.ops.b200wave.ssim_fwd
ssim_bwd = torch.ops.b200wave.ssim_bwd
