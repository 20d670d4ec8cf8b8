"""Reduction-style losses of the training step that sit beside the wavelet / SSIM path.

``TVLoss`` -- drop-in for ``model.py:17-33`` (constructed at ``train.py:98``): same constructor argument, same scalar
``TVLoss_weight * 2 * (h_tv / count_h + w_tv / count_w) / batch_size``.  The reference slices ``x`` four times,
subtracts, squares and reduces (about ten elementwise passes and as many again in autograd's backward); here the two
sums come out of one pass over ``x`` (``b200w_tv_fwd_f32``) and the gradient out of one more (``b200w_tv_bwd_f32``).
CUDA-only, like the rest of the package.
"""
import ctypes

import torch
import torch.nn as nn

from . import _cabi

_LIB = torch.library.Library("b200wave_losses", "DEF")
_LIB.define("tv_sums(Tensor x) -> Tensor")
_LIB.define("tv_grad(Tensor x, Tensor grad_out, float ch, float cw) -> Tensor")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(x, name):
    if not x.is_cuda:
        raise RuntimeError("b200wave_losses::%s is CUDA-only (sm_100a); there is no CPU fallback -- got a %s tensor"
                           % (name, x.device))
    if x.dtype != torch.float32:
        raise RuntimeError("b200wave_losses::%s: expected scalar type Float but found %s" % (name, x.dtype))
    if x.dim() != 4:
        raise IndexError("b200wave_losses::%s expects a 4-D (N, C, H, W) tensor, got %d-D" % (name, x.dim()))


def _tv_sums_cuda(x):
    """(2,) tensor: sum of squared vertical / horizontal forward differences (model.py:28-29)."""
    _check(x, "tv_sums")
    lib = _cabi.load()
    xc = x.contiguous()
    n, c, h, w = xc.shape
    out = torch.zeros(2, device=x.device, dtype=torch.float32)
    if xc.numel() == 0:
        return out
    ws = torch.empty(int(lib.b200w_tv_workspace_bytes(n * c, h)), device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        rc = lib.b200w_tv_fwd_f32(xc.data_ptr(), n * c, h, w, ws.data_ptr(), ws.numel(), out.data_ptr(), _stream())
    _cabi.check(rc, "tv")
    return out


def _tv_grad_cuda(x, grad_out, ch, cw):
    _check(x, "tv_grad")
    lib = _cabi.load()
    xc = x.contiguous()
    n, c, h, w = xc.shape
    dx = torch.empty_like(xc)
    if xc.numel() == 0:
        return dx
    g = grad_out.reshape(-1)[:1].to(torch.float32).contiguous()
    with torch.cuda.device(x.device):
        rc = lib.b200w_tv_bwd_f32(xc.data_ptr(), g.data_ptr(), float(ch), float(cw), n * c, h, w, dx.data_ptr(),
                                  _stream())
    _cabi.check(rc, "tv")
    return dx


def _cpu_refuse(name):
    def impl(*args, **kwargs):
        raise RuntimeError("b200wave_losses::%s is CUDA-only (sm_100a): there is no CPU fallback. Move the tensors "
                           "to a B200 (`.cuda()`)." % name)
    return impl


_LIB.impl("tv_sums", _tv_sums_cuda, "CUDA")
_LIB.impl("tv_grad", _tv_grad_cuda, "CUDA")
_LIB.impl("tv_sums", _cpu_refuse("tv_sums"), "CPU")
_LIB.impl("tv_grad", _cpu_refuse("tv_grad"), "CPU")
torch.library.register_fake("b200wave_losses::tv_sums", lambda x: x.new_empty((2,)), lib=_LIB)
torch.library.register_fake("b200wave_losses::tv_grad", lambda x, g, ch, cw: torch.empty_like(x), lib=_LIB)


class _TV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        n, c, h, w = x.shape
        count_h = c * (h - 1) * w          # model.py:26-27 via _tensor_size
        count_w = c * h * (w - 1)
        sums = torch.ops.b200wave_losses.tv_sums(x)
        ctx.save_for_backward(x)
        ctx.scales = (weight, n, count_h, count_w)
        # the same arithmetic as model.py:30 (a zero count divides by zero exactly as the reference does)
        return weight * 2 * (sums[0] / count_h + sums[1] / count_w) / n

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        weight, n, count_h, count_w = ctx.scales
        ch = weight * 2.0 / (n * count_h) if count_h else float("nan")
        cw = weight * 2.0 / (n * count_w) if count_w else float("nan")
        return torch.ops.b200wave_losses.tv_grad(x, g, ch, cw), None


class TVLoss(nn.Module):
    """``model.TVLoss`` (model.py:17-33)."""

    def __init__(self, TVLoss_weight=1):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x):
        return _TV.apply(x, self.TVLoss_weight)

    def _tensor_size(self, t):
        return t.size()[1] * t.size()[2] * t.size()[3]
