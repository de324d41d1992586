"""SDFNetwork / RenderingNetwork with the reference's constructor, methods and state-dict keys
(models/fields.py:9-137, 141-239), evaluated by the sm_100a kernels of libiron_b200.so.

The reference differentiates these modules with autograd (including the double backward through
`get_all(..., is_training=True)` and `gradient()`); here the same derivatives come from closed-form CUDA
kernels wrapped in `torch.autograd.Function`s that live inside the modules, so the training loop above
them does not change.  Gradients are produced for every parameter and for the inputs the reference's
stage-2 step differentiates (normals / features / points of the material nets); d/dx of the SDF outputs
is not produced (the reference never uses it: the query points are detached tracer outputs).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .embedder import get_embedder


class _WNLinear(nn.Module):
    """Parameter holder with the keys of nn.utils.weight_norm(nn.Linear): weight_g [out,1], weight_v [out,in], bias."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, weight_norm: bool):
        super().__init__()
        self.weight_normed = weight_norm
        self.in_features = weight.shape[1]
        self.out_features = weight.shape[0]
        self.bias = nn.Parameter(bias.detach().clone())
        if weight_norm:
            self.weight_g = nn.Parameter(weight.detach().norm(dim=1, keepdim=True))
            self.weight_v = nn.Parameter(weight.detach().clone())
        else:
            self.weight = nn.Parameter(weight.detach().clone())

    def tensors(self) -> List[nn.Parameter]:
        return [self.weight_g, self.weight_v, self.bias] if self.weight_normed else [self.weight, self.bias]


class _FoldedMLP(nn.Module):
    """Shared machinery: packed folded weights (cached per parameter version) and their unfold."""

    layout: _lib.MlpLayout

    def _lins(self) -> List[_WNLinear]:
        return [getattr(self, f"lin{l}") for l in range(self.layout.n_lin)]

    def _param_list(self) -> List[nn.Parameter]:
        out: List[nn.Parameter] = []
        for lin in self._lins():
            out += lin.tensors()
        return out

    def _check_device(self, t: torch.Tensor) -> None:
        p = self.lin0.bias
        if not p.is_cuda:
            raise RuntimeError("iron_b200 modules run on CUDA only (no CPU path): call .cuda() first")
        if t.device != p.device:
            raise RuntimeError(f"input on {t.device}, parameters on {p.device}")
        if p.dtype != torch.float32:
            raise RuntimeError("iron_b200 modules are fp32 (the reference's dtype)")

    def _v_g_b(self):
        v, g, b = [], [], []
        for lin in self._lins():
            if lin.weight_normed:
                v.append(lin.weight_v.detach()); g.append(lin.weight_g.detach())
            else:
                v.append(lin.weight.detach()); g.append(None)
            b.append(lin.bias.detach())
        return v, g, b

    def folded(self) -> torch.Tensor:
        """Packed effective weights W = v * g/||v|| (+ transposes, biases), refolded when a parameter changed."""
        params = self._param_list()
        key = tuple((p.data_ptr(), p._version) for p in params)
        # under CUDA-graph capture the fold must be part of the graph (the parameters change between replays): the first
        # call of a capture always refolds, later calls of the same capture reuse it (GraphedStage2Step resets the flag)
        capturing = torch.cuda.is_current_stream_capturing()
        if (getattr(self, "_fold_key", None) == key and self._fold_buf is not None
                and (not capturing or getattr(self, "_fold_captured", False))):
            return self._fold_buf
        dev = params[0].device
        v, g, b = self._v_g_b()
        for t in v + b + [x for x in g if x is not None]:
            if not t.is_contiguous():
                raise RuntimeError("iron_b200: parameters must be contiguous")
        buf = getattr(self, "_fold_buf", None)
        if buf is None or buf.device != dev:
            buf = torch.zeros(int(self.layout.packed_total_floats), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ironb_mlp_fold(C.byref(self.layout), _lib.ptr_array(v), _lib.ptr_array(g),
                                                  _lib.ptr_array(b), _lib.ptr(buf), _lib.stream()), "mlp_fold")
        self._fold_buf = buf
        self._fold_key = key
        if capturing:
            self._fold_captured = True
        return buf

    def unfold_grads(self, dpacked: torch.Tensor) -> List[torch.Tensor]:
        """dL/d(packed W, b) -> gradients in _param_list() order."""
        v, g, _ = self._v_g_b()
        dv = [torch.empty_like(t) for t in v]
        dg = [torch.empty_like(t) if t is not None else None for t in g]
        db = [torch.empty_like(lin.bias) for lin in self._lins()]
        with torch.cuda.device(dpacked.device):
            _lib.check(_lib.load().ironb_mlp_fold_bwd(C.byref(self.layout), _lib.ptr_array(v), _lib.ptr_array(g),
                                                      _lib.ptr(dpacked), _lib.ptr_array(dv), _lib.ptr_array(dg),
                                                      _lib.ptr_array(db), _lib.stream()), "mlp_fold_bwd")
        out: List[torch.Tensor] = []
        for l, lin in enumerate(self._lins()):
            out += [dg[l], dv[l], db[l]] if lin.weight_normed else [dv[l], db[l]]
        return out


# ------------------------------------------------------------------------------------------------ SDF
class _SDFEval(torch.autograd.Function):
    """(y, feature, grad) = get_all(x); backward = the closed-form double backward w.r.t. the parameters."""

    @staticmethod
    def forward(ctx, net: "SDFNetwork", x: torch.Tensor, want_yf: bool, want_grad: bool, *params):
        lib = _lib.load()
        ctx.set_materialize_grads(False)
        M = x.shape[0]
        dev = x.device
        packed = net.folded()
        save = any(ctx.needs_input_grad[4:])
        ctx.simt = bool(os.environ.get("IRONB_EIK_SIMT")) and not want_yf      # diagnostic: FFMA GEMMs for gradient-only calls
        prev_mode = lib.ironb_set_gemm_mode(0) if ctx.simt else None
        lay = net.layout
        y = torch.empty(M, 1, dtype=torch.float32, device=dev) if want_yf else None
        feat = torch.empty(M, lay.d_out - 1, dtype=torch.float32, device=dev) if want_yf else None
        grad = torch.empty(M, 3, dtype=torch.float32, device=dev) if want_grad else None
        with torch.cuda.device(dev):
            nbytes = lib.ironb_sdf_getall_workspace_bytes(C.byref(lay), M, int(want_grad), int(save))
            ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
            _lib.check(lib.ironb_sdf_getall_fwd(C.byref(lay), _lib.ptr(packed), _lib.ptr(x), M, _lib.ptr(y), _lib.ptr(feat),
                                                _lib.ptr(grad), int(save), _lib.ptr(ws), ws.numel(), _lib.stream()),
                       "sdf_getall_fwd")
        if prev_mode is not None:
            lib.ironb_set_gemm_mode(prev_mode)
        ctx.net, ctx.M, ctx.x = net, M, x
        ctx.ws = ws if save else None
        ctx.packed = packed if save else None
        ctx.want = (want_yf, want_grad)
        e = lambda: torch.empty(0, dtype=torch.float32, device=dev)
        outs = (y if want_yf else e(), feat if want_yf else e(), grad if want_grad else e())
        ctx.mark_non_differentiable(*[o for o, w in zip(outs, (want_yf, want_yf, want_grad)) if not w])
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy, gfeat, ggrad):
        net, M = ctx.net, ctx.M
        lib = _lib.load()
        lay = net.layout
        if ctx.ws is None:
            raise RuntimeError("iron_b200: SDF forward state was not saved (no parameter required grad)")
        want_yf, want_grad = ctx.want
        gy = _lib.f32c(gy).reshape(M) if (want_yf and gy is not None) else None
        gfeat = _lib.f32c(gfeat) if (want_yf and gfeat is not None) else None
        ggrad = _lib.f32c(ggrad) if (want_grad and ggrad is not None) else None
        dev = ctx.x.device
        dpacked = torch.zeros(int(lay.packed_floats), dtype=torch.float32, device=dev)
        prev_mode = lib.ironb_set_gemm_mode(0) if ctx.simt else None
        with torch.cuda.device(dev):
            _lib.check(lib.ironb_sdf_getall_bwd(C.byref(lay), _lib.ptr(ctx.packed), _lib.ptr(ctx.x), M, _lib.ptr(gy),
                                                _lib.ptr(gfeat), _lib.ptr(ggrad), _lib.ptr(ctx.ws), ctx.ws.numel(),
                                                _lib.ptr(dpacked), _lib.stream(), _lib.wgrad_stream()), "sdf_getall_bwd")
        if prev_mode is not None:
            lib.ironb_set_gemm_mode(prev_mode)
        grads = net.unfold_grads(dpacked)
        ctx.ws = None
        return (None, None, None, None, *grads)


class SDFNetwork(_FoldedMLP):
    """models/fields.py:9-137.  Same ctor kwargs, same state-dict keys (lin{l}.weight_g / weight_v / bias)."""

    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1,
                 geometric_init=True, weight_norm=True, inside_outside=False):
        super().__init__()
        dims = [d_in] + [d_hidden for _ in range(n_layers)] + [d_out]
        self.embed_fn_fine = None
        if multires > 0:
            embed_fn, input_ch = get_embedder(multires, input_dims=d_in)
            self.embed_fn_fine = embed_fn
            dims[0] = input_ch
        self.num_layers = len(dims)
        self.skip_in = tuple(skip_in)
        self.scale = scale
        self.multires = multires
        if len(self.skip_in) > 1:
            raise NotImplementedError("iron_b200.SDFNetwork supports at most one skip layer")
        # same RNG consumption as the reference ctor: one nn.Linear per layer, then the init overrides (:47-73)
        for l in range(0, self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if geometric_init:
                if l == self.num_layers - 2:
                    if not inside_outside:
                        torch.nn.init.normal_(lin.weight, mean=np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                        torch.nn.init.constant_(lin.bias, -bias)
                    else:
                        torch.nn.init.normal_(lin.weight, mean=-np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                        torch.nn.init.constant_(lin.bias, bias)
                elif multires > 0 and l == 0:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.constant_(lin.weight[:, 3:], 0.0)
                    torch.nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
                elif multires > 0 and l in self.skip_in:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    torch.nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
                else:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            setattr(self, "lin" + str(l), _WNLinear(lin.weight.data, lin.bias.data, weight_norm))
        self.layout = _lib.MlpLayout()
        skip = self.skip_in[0] if self.skip_in else -1
        _lib.check(_lib.load().ironb_sdf_layout(d_in, d_out, d_hidden, n_layers, skip, multires, float(scale), 100.0,
                                                C.byref(self.layout)), "sdf_layout")

    # -- evaluation -----------------------------------------------------------------------------------
    def _eval(self, x: torch.Tensor, want_yf: bool, want_grad: bool, params=None):
        self._check_device(x)
        sh = list(x.shape[:-1])
        xf = _lib.f32c(x.detach().reshape(-1, 3))
        y, feat, grad = _SDFEval.apply(self, xf, want_yf, want_grad, *(params if params is not None else self._param_list()))
        y = y.reshape(sh + [1]) if want_yf else None
        feat = feat.reshape(sh + [feat.shape[-1]]) if want_yf else None
        grad = grad.reshape(sh + [3]) if want_grad else None
        return y, feat, grad

    def forward(self, inputs):
        y, feat, _ = self._eval(inputs, True, False)
        return torch.cat([y, feat], dim=-1)

    def sdf(self, x):
        return self.forward(x)[..., :1]

    def sdf_hidden_appearance(self, x):
        return self.forward(x)

    def gradient(self, x):
        """d sdf / d x  [..., 3]; differentiable w.r.t. the parameters (create_graph=True in the reference, :106-118)."""
        return self._eval(x, False, True)[2]

    def gradient_aliased(self, x):
        """gradient(x), differentiated w.r.t. ALIAS leaves of the parameters (same storage, separate autograd identity).

        A stage-2 step uses the SDF parameters twice (hit points and eikonal samples), so autograd ends the backward pass with
        one `p.grad += g` launch per parameter tensor (27 launches on the step's critical path).  With the second use
        differentiated w.r.t. aliases, both gradient sets arrive separately and `add_alias_grads` merges them with one
        multi-tensor launch.  Returns (gradient, aliases); call `add_alias_grads(aliases)` after `backward()`."""
        aliases = [p.detach().requires_grad_(True) for p in self._param_list()]
        return self._eval(x, False, True, params=aliases)[2], aliases

    @torch.no_grad()
    def add_alias_grads(self, aliases):
        ps = self._param_list()
        dst, src = [], []
        for p, a in zip(ps, aliases):
            if a.grad is None:
                continue
            if p.grad is None:
                p.grad = a.grad
            else:
                dst.append(p.grad)
                src.append(a.grad)
        if dst:
            torch._foreach_add_(dst, src)

    def get_all(self, x, is_training=True):
        """(sdf [...,1], feature [...,d_out-1], d sdf/d x [...,3]), detached when not training (:120-137)."""
        if not is_training:
            with torch.no_grad():
                return self._eval(x, True, True)
        with torch.enable_grad():
            return self._eval(x, True, True)


# ------------------------------------------------------------------------------------------------ material nets
_MODES = {"idr": 0, "no_view_dir": 1, "no_normal": 2, "points_only": 3}


class _MatEval(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net: "RenderingNetwork", points, normals, view_dirs, feats, *params):
        lib = _lib.load()
        M = points.shape[0]
        dev = points.device
        packed = net.folded()
        lay, cfg = net.layout, net.cfg
        out = torch.empty(M, lay.d_out, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nbytes = lib.ironb_matnet_workspace_bytes(C.byref(lay), M)
            ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
            _lib.check(lib.ironb_matnet_fwd(C.byref(lay), C.byref(cfg), _lib.ptr(packed), _lib.ptr(points),
                                            _lib.ptr(normals), _lib.ptr(view_dirs), _lib.ptr(feats), M, _lib.ptr(out),
                                            _lib.ptr(ws), ws.numel(), _lib.stream()), "matnet_fwd")
        ctx.net, ctx.M = net, M
        need = any(ctx.needs_input_grad[1:])
        ctx.ws = ws if need else None
        ctx.packed = packed if need else None
        ctx.out = out
        ctx.have = (normals is not None, view_dirs is not None)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        net, M = ctx.net, ctx.M
        lib = _lib.load()
        lay, cfg = net.layout, net.cfg
        dev = gout.device
        gout = _lib.f32c(gout)
        ng = ctx.needs_input_grad
        d_points = torch.empty(M, 3, dtype=torch.float32, device=dev) if ng[1] else None
        d_normals = torch.empty(M, 3, dtype=torch.float32, device=dev) if (ng[2] and ctx.have[0]) else None
        d_view = torch.empty(M, 3, dtype=torch.float32, device=dev) if (ng[3] and ctx.have[1]) else None
        d_feats = torch.empty(M, cfg.d_feature, dtype=torch.float32, device=dev) if ng[4] else None
        dpacked = torch.zeros(int(lay.packed_floats), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ironb_matnet_bwd(C.byref(lay), C.byref(cfg), _lib.ptr(ctx.packed), M, _lib.ptr(ctx.out),
                                            _lib.ptr(gout), _lib.ptr(ctx.ws), ctx.ws.numel(), _lib.ptr(dpacked),
                                            _lib.ptr(d_points), _lib.ptr(d_normals), _lib.ptr(d_view), _lib.ptr(d_feats),
                                            _lib.stream(), _lib.wgrad_stream()), "matnet_bwd")
        grads = net.unfold_grads(dpacked) if any(ng[5:]) else [None] * (len(ng) - 5)
        ctx.ws = None
        return (None, d_points, d_normals, d_view, d_feats, *grads)


class RenderingNetwork(_FoldedMLP):
    """models/fields.py:141-239.  skip_in: () (every 'ggx' / 'comp2' network) or ONE layer (the stage-1 colour network:
    n_layers = 8, skip_in = [4])."""

    def __init__(self, d_feature, mode, d_in, d_out, d_hidden, n_layers, weight_norm=True, multires=0,
                 multires_view=0, squeeze_out=True, squeeze_out_scale=1.0, output_bias=0.0, output_scale=1.0,
                 skip_in=()):
        super().__init__()
        skip_in = tuple(skip_in)
        if len(skip_in) > 1 or any(not (1 <= int(l) <= n_layers) for l in skip_in):
            raise NotImplementedError("iron_b200.RenderingNetwork: at most one skip layer, in [1, n_layers]")
        if mode not in _MODES:
            raise ValueError(f"unknown mode {mode!r}")
        self.mode = mode
        self.squeeze_out = squeeze_out
        dims = [d_in + d_feature] + [d_hidden for _ in range(n_layers)] + [d_out]
        self.embed_fn = None
        if multires > 0:
            embed_fn, input_ch = get_embedder(multires)
            self.embed_fn = embed_fn
            dims[0] += input_ch - 3
        self.embedview_fn = None
        if multires_view > 0:
            embedview_fn, input_ch = get_embedder(multires_view)
            self.embedview_fn = embedview_fn
            dims[0] += input_ch - 3
        self.num_layers = len(dims)
        self.skip_in = skip_in
        in0 = dims[0]
        for l in range(0, self.num_layers - 1):                 # :179-181: the skip layer reads cat(h, input)
            if l in self.skip_in:
                dims[l] += in0
        for l in range(0, self.num_layers - 1):                 # :183-187 (same nn.Linear shapes = same RNG consumption)
            out_dim = dims[l + 1] - in0 if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            setattr(self, "lin" + str(l), _WNLinear(lin.weight.data, lin.bias.data, weight_norm))
        self.output_bias = output_bias
        self.output_scale = output_scale
        self.squeeze_out_scale = squeeze_out_scale
        uses_view = mode in ("idr", "no_normal")
        self.cfg = _lib.MatnetCfg(_MODES[mode], max(multires, 0), max(multires_view, 0) if uses_view else 0, d_feature,
                                  int(bool(squeeze_out)), float(output_bias), float(output_scale),
                                  float(squeeze_out_scale))
        lib = _lib.load()
        kernel_in = lib.ironb_matnet_in_dim(C.byref(self.cfg))
        if kernel_in != in0:
            raise ValueError(f"RenderingNetwork: d_in={d_in} is inconsistent with mode {mode!r} "
                             f"(kernel input width {kernel_in}, ctor width {in0})")
        self.layout = _lib.MlpLayout()
        _lib.check(lib.ironb_matnet_layout_skip(in0, d_out, d_hidden, n_layers, int(skip_in[0]) if skip_in else -1,
                                                C.byref(self.layout)), "matnet_layout")

    def forward(self, points, normals, view_dirs, feature_vectors):
        self._check_device(points)
        sh = list(points.shape[:-1])
        flat = lambda t, w: None if t is None else _lib.f32c(t.reshape(-1, w))
        uses_n = self.mode in ("idr", "no_view_dir")
        uses_v = self.mode in ("idr", "no_normal")
        out = _MatEval.apply(self, flat(points, 3), flat(normals, 3) if uses_n else None,
                             flat(view_dirs, 3) if uses_v else None,
                             flat(feature_vectors, feature_vectors.shape[-1]), *self._param_list())
        return out.reshape(sh + [out.shape[-1]])
