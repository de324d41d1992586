"""Autograd wrappers of the fused glue kernels (csrc/glue.cu): the element-wise chains the reference issues as separate ATen
ops between the big kernels of a shading step -- one launch forward, one backward each.  CUDA fp32 only (no CPU path)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _c(t):
    return None if t is None else _lib.f32c(t)


def _cuda_f32(*ts):
    for t in ts:
        if t is not None and not (t.is_cuda and t.dtype == torch.float32):
            raise RuntimeError("iron_b200: fused glue ops need fp32 CUDA tensors (there is no CPU path)")


class _UnitDist(torch.autograd.Function):
    """n = g / (|g| + 1e-10), dist = |x - o| (render_surface.py:135-146)."""

    @staticmethod
    def forward(ctx, g, x, o):
        g, x, o = _c(g), _c(x), _c(o)
        M = g.shape[0]
        n = torch.empty_like(g)
        dist = torch.empty(M, 1, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().ironb_unit_dist_fwd(_lib.ptr(g), _lib.ptr(x), _lib.ptr(o), M, _lib.ptr(n), _lib.ptr(dist),
                                                       _lib.stream()), "unit_dist_fwd")
        ctx.save_for_backward(g, x, o)
        ctx.set_materialize_grads(False)
        return n, dist

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dn, ddist):
        g, x, o = ctx.saved_tensors
        M = g.shape[0]
        dg = torch.empty_like(g) if ctx.needs_input_grad[0] else None
        dx = torch.empty_like(x) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().ironb_unit_dist_bwd(_lib.ptr(g), _lib.ptr(x), _lib.ptr(o), _lib.ptr(_c(dn)), _lib.ptr(_c(ddist)), M,
                                                       _lib.ptr(dg), _lib.ptr(dx), _lib.stream()), "unit_dist_bwd")
        return dg, dx, None


def unit_normal_and_distance(grads, points, ray_o):
    _cuda_f32(grads, points, ray_o)
    return _UnitDist.apply(grads, points, ray_o)


class _Reparam(torch.autograd.Function):
    """models/raytracer.py:17-24: the value is the points themselves; the gradient w.r.t. the SDF values is -v/(g.v) . d x."""

    @staticmethod
    def forward(ctx, points, grads, dirs, f):
        ctx.save_for_backward(_c(grads), _c(dirs))
        return points.view_as(points)          # the value IS the points (f - f.detach() == 0); no kernel

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dx):
        g, v = ctx.saved_tensors
        M = g.shape[0]
        df = torch.empty(M, 1, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().ironb_reparam_bwd(_lib.ptr(g), _lib.ptr(v), _lib.ptr(_c(dx)), M, _lib.ptr(df), _lib.stream()),
                       "reparam_bwd")
        return None, None, None, df


def reparam(points, grads, dirs, f):
    _cuda_f32(points, grads, dirs, f)
    return _Reparam.apply(points, grads, dirs, f)


class _MatPost(torch.autograd.Function):
    """models/rendering_func.py:5-16 after the three networks: abs / channel mean / + 0.01."""

    @staticmethod
    def forward(ctx, a, b, c, is_metal):
        a, b, c = _c(a), _c(b), _c(c)
        M = a.shape[0]
        kd, ks, al = torch.empty_like(a), torch.empty_like(b), torch.empty_like(c)
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().ironb_matpost_fwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(c), M, int(bool(is_metal)), _lib.ptr(kd),
                                                     _lib.ptr(ks), _lib.ptr(al), _lib.stream()), "matpost_fwd")
        ctx.save_for_backward(a, b, c)
        ctx.is_metal = int(bool(is_metal))
        ctx.set_materialize_grads(False)
        return kd, ks, al

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dkd, dks, dal):
        a, b, c = ctx.saved_tensors
        M = a.shape[0]
        da, db, dc = torch.empty_like(a), torch.empty_like(b), torch.empty_like(c)
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().ironb_matpost_bwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(_c(dkd)), _lib.ptr(_c(dks)),
                                                     _lib.ptr(_c(dal)), M, ctx.is_metal, _lib.ptr(da), _lib.ptr(db), _lib.ptr(dc),
                                                     _lib.stream()), "matpost_bwd")
        return da, db, dc, None


def material_post(a, b, c, is_metal=False):
    _cuda_f32(a, b, c)
    return _MatPost.apply(a, b, c, is_metal)


class _EikSum(torch.autograd.Function):
    """sum_m w_m (|g_m| - 1)^2 (render_surface.py:580-583, 601-603); w: [M] weights or None."""

    @staticmethod
    def forward(ctx, g, w):
        g, w = _c(g), _c(w)
        out = torch.zeros((), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().ironb_eik_sum_fwd(_lib.ptr(g), _lib.ptr(w), g.shape[0], _lib.ptr(out), _lib.stream()), "eik_sum_fwd")
        ctx.save_for_backward(g, w) if w is not None else ctx.save_for_backward(g)
        ctx.has_w = w is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, up):
        saved = ctx.saved_tensors
        g, w = saved[0], (saved[1] if ctx.has_w else None)
        dg = torch.empty_like(g)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().ironb_eik_sum_bwd(_lib.ptr(g), _lib.ptr(w), _lib.ptr(_c(up)), g.shape[0], _lib.ptr(dg),
                                                     _lib.stream()), "eik_sum_bwd")
        return dg, None


def eikonal_sum(g, w=None):
    _cuda_f32(g, w)
    return _EikSum.apply(g.reshape(-1, 3), None if w is None else w.reshape(-1))


class _RoughRange(torch.autograd.Function):
    """render_surface.py:609-613 without the host-side emptiness test: weight * mean over the selected pixels, 0 if none."""

    @staticmethod
    def forward(ctx, r, w, value, weight):
        r, w = _c(r), _c(w)
        acc = torch.zeros(2, dtype=torch.float32, device=r.device)
        loss = torch.empty((), dtype=torch.float32, device=r.device)
        with torch.cuda.device(r.device):
            _lib.check(_lib.load().ironb_roughrange_fwd(_lib.ptr(r), _lib.ptr(w), r.numel(), float(value), float(weight), _lib.ptr(acc),
                                                        _lib.ptr(loss), _lib.stream()), "roughrange_fwd")
        ctx.save_for_backward(r, w, acc)
        ctx.vw = (float(value), float(weight))
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, up):
        r, w, acc = ctx.saved_tensors
        dr = torch.empty_like(r)
        with torch.cuda.device(r.device):
            _lib.check(_lib.load().ironb_roughrange_bwd(_lib.ptr(r), _lib.ptr(w), _lib.ptr(acc), _lib.ptr(_c(up)), r.numel(), ctx.vw[0],
                                                        ctx.vw[1], _lib.ptr(dr), _lib.stream()), "roughrange_bwd")
        return dr, None, None, None


def roughrange(roughness, mask_f, value=0.5, weight=0.1):
    _cuda_f32(roughness, mask_f)
    return _RoughRange.apply(roughness.reshape(-1), mask_f.reshape(-1), value, weight)


def _mask_rows_launch(srcs, w):
    M = w.numel()
    outs = [torch.empty_like(s) for s in srcs]
    n = len(srcs)
    widths = (C.c_int * n)(*[(s.numel() // M) if M else 1 for s in srcs])
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().ironb_mask_rows(_lib.ptr_array(srcs), _lib.ptr_array(outs), widths, n, _lib.ptr(w), M, _lib.stream()),
                   "mask_rows")
    return outs


class _MaskRows(torch.autograd.Function):
    """y_t[m] = x_t[m] * w[m] for up to 8 row tensors in one launch (dense shading zeroes the non-hit pixels)."""

    @staticmethod
    def forward(ctx, w, *xs):
        w = _c(w)
        xs = [_c(x) for x in xs]
        ctx.save_for_backward(w)
        ctx.set_materialize_grads(False)
        return tuple(_mask_rows_launch(xs, w))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gs):
        (w,) = ctx.saved_tensors
        live = [(i, _c(g)) for i, g in enumerate(gs) if g is not None and ctx.needs_input_grad[i + 1]]
        out = [None] * len(gs)
        if live:
            for (i, _), r in zip(live, _mask_rows_launch([g for _, g in live], w)):
                out[i] = r
        return (None, *out)


def mask_rows(mask_f, tensors):
    """tensors: list of [M] / [M, k] fp32 CUDA tensors (at most 8 per launch); returns them multiplied by mask_f[m]."""
    _cuda_f32(mask_f, *tensors)
    outs = []
    for s in range(0, len(tensors), 8):
        outs += list(_MaskRows.apply(mask_f.reshape(-1), *tensors[s:s + 8]))
    return outs
