"""The stage-2 loop's patch losses with the reference's names and signatures (models/image_losses.py:13-48, 97-158):
`PyramidL2Loss(use_cuda=True)(pred_img, trgt_img)` and `ssim_loss_fn(X, Y, mask=None, ...)`, on [B, C, H, W] images
(the reference passes B = 1, C = 3 views of the renderer's [H, W, 3] buffer, render_surface.py:594-598 -- read in place
through their strides).  Both are terminal losses, so the CUDA kernels (csrc/losses.cu) produce the value and the gradient
w.r.t. the first argument in one pass; backward scales that gradient by the upstream scalar.  Gradients flow to `pred_img`
/ `X` only (the reference's targets are data).  No host synchronisation: usable inside GraphedStage2Step's capture."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _strides3(t: torch.Tensor, b: int):
    """(pointer of image b, int64[3] element strides {channel, row, column}) of a [B, C, H, W] fp32 CUDA tensor."""
    s = t.stride()
    return t.data_ptr() + 4 * b * s[0], (C.c_int64 * 3)(s[1], s[2], s[3])


def _check(pred: torch.Tensor, other: torch.Tensor, what: str):
    if pred.dim() != 4 or pred.shape != other.shape:
        raise ValueError(f"{what}: inputs must be [B, C, H, W] tensors of the same shape")
    if not (pred.is_cuda and other.is_cuda):
        raise RuntimeError(f"{what}: iron_b200 runs on CUDA tensors only (there is no CPU path)")
    if pred.dtype != torch.float32 or other.dtype != torch.float32:
        raise ValueError(f"{what}: fp32 images expected (the reference's dtype)")


def _grad_like(t: torch.Tensor) -> torch.Tensor:
    """Gradient buffer with the memory layout of `t` when that is dense (a permuted view of a contiguous buffer)."""
    try:
        return torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, device=t.device) if t.is_non_overlapping_and_dense() \
            else torch.empty_like(t, memory_format=torch.contiguous_format)
    except Exception:
        return torch.empty_like(t, memory_format=torch.contiguous_format)


class _PyramidL2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, trgt):
        lib = _lib.load()
        B, Cc, H, W = pred.shape
        need = ctx.needs_input_grad[0]
        loss = torch.zeros((), dtype=torch.float32, device=pred.device)
        grad = _grad_like(pred) if need else None
        with torch.cuda.device(pred.device):
            nb = int(lib.ironb_patch_loss_workspace_bytes(Cc, H, W))
            if nb < 0:
                raise ValueError("PyramidL2Loss: at most 4 channels are supported")
            ws = torch.empty(nb, dtype=torch.uint8, device=pred.device)
            for b in range(B):
                pp, ps = _strides3(pred, b)
                tp, ts = _strides3(trgt, b)
                gp, gs = _strides3(grad, b) if need else (None, None)
                _lib.check(lib.ironb_pyramid_l2(pp, ps, tp, ts, Cc, H, W, _lib.ptr(loss), gp, gs, _lib.ptr(ws), ws.numel(),
                                                _lib.stream()), "pyramid_l2")
        ctx.grad = grad
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        return (ctx.grad * gout if ctx.grad is not None else None), None


class PyramidL2Loss(torch.nn.Module):
    """models/image_losses.py:13-48.  `use_cuda` is accepted for signature parity; the kernels always run on the inputs' GPU."""

    def __init__(self, use_cuda=True):
        super().__init__()

    def forward(self, pred_img, trgt_img):
        _check(pred_img, trgt_img, "PyramidL2Loss")
        return _PyramidL2.apply(pred_img, trgt_img.detach())


class _SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, Y, mask, data_range, win_size, win_sigma, K1, K2):
        lib = _lib.load()
        B, Cc, H, W = X.shape
        need = ctx.needs_input_grad[0]
        grad = _grad_like(X) if need else None
        losses = torch.empty(B, dtype=torch.float32, device=X.device)
        m8 = None
        if mask is not None:
            m8 = mask.reshape(B, H, W).to(torch.uint8).contiguous()
        with torch.cuda.device(X.device):
            nb = int(lib.ironb_patch_loss_workspace_bytes(Cc, H, W))
            if nb < 0:
                raise ValueError("ssim_loss_fn: at most 4 channels are supported")
            for b in range(B):
                ws = torch.empty(nb, dtype=torch.uint8, device=X.device)
                xp, xs = _strides3(X, b)
                yp, ys = _strides3(Y, b)
                gp, gs = _strides3(grad, b) if need else (None, None)
                mp = m8.data_ptr() + b * H * W if m8 is not None else None
                _lib.check(lib.ironb_ssim_loss(xp, xs, yp, ys, mp, Cc, H, W, float(data_range), int(win_size), float(win_sigma),
                                               float(K1), float(K2), losses.data_ptr() + 4 * b, gp, gs, _lib.ptr(ws), ws.numel(),
                                               _lib.stream()), "ssim_loss")
        ctx.grad, ctx.B = grad, B
        if B == 1:
            return losses.reshape(())
        raise NotImplementedError("ssim_loss_fn: batches > 1 with a joint mean are not built (the reference passes one patch)")

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        return (ctx.grad * gout if ctx.grad is not None else None), None, None, None, None, None, None, None


def ssim_loss_fn(X, Y, mask=None, data_range=1.0, win_size=11, win_sigma=1.5, K=(0.01, 0.03)):
    """models/image_losses.py:97-158: 1 - mean SSIM; with `mask` ([B, 1, H, W] bool) the map is padded back with 1.0 and averaged
    over the mask eroded by the 11 x 11 window (kornia.morphology.erosion, restated and pinned against OpenCV)."""
    if not X.shape == Y.shape:
        raise ValueError("Input images should have the same dimensions.")
    if not X.type() == Y.type():
        raise ValueError("Input images should have the same dtype.")
    if len(X.shape) != 4:
        raise ValueError(f"Input images should be 4-d tensors, but got {X.shape}")
    if not (win_size % 2 == 1):
        raise ValueError("Window size should be odd.")
    _check(X, Y, "ssim_loss_fn")
    return _SSIM.apply(X, Y.detach(), mask, data_range, win_size, win_sigma, K[0], K[1])
