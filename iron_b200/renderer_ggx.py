"""GGXColocatedRenderer (models/renderer_ggx.py:61-146) and CompositeRenderer (:520-858, the "comp2" model of the fork's
working driver) with the reference's interfaces, evaluated by the fused forward / backward kernels ironb_ggx_* and
ironb_composite_*."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib

_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ggx_tables.npz")


class _GGXShade(torch.autograd.Function):
    @staticmethod
    def forward(ctx, light, dist, normal, viewdir, kd, ks, alpha, trans, diff_trans):
        ctx.set_materialize_grads(False)
        M = dist.shape[0]
        dev = dist.device
        outs = [torch.empty(M, 3, dtype=torch.float32, device=dev) for _ in range(3)]
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ironb_ggx_fwd(_lib.ptr(light), _lib.ptr(dist), _lib.ptr(normal), _lib.ptr(viewdir),
                                                 _lib.ptr(kd), _lib.ptr(ks), _lib.ptr(alpha), _lib.ptr(trans),
                                                 _lib.ptr(diff_trans), M, _lib.ptr(outs[0]), _lib.ptr(outs[1]),
                                                 _lib.ptr(outs[2]), _lib.stream()), "ggx_fwd")
        ctx.save_for_backward(light, dist, normal, viewdir, kd, ks, alpha, trans, diff_trans)
        return tuple(outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_diff, g_spec, g_rgb):
        light, dist, normal, viewdir, kd, ks, alpha, trans, diff_trans = ctx.saved_tensors
        M = dist.shape[0]
        dev = dist.device
        g = [None if t is None else _lib.f32c(t) for t in (g_diff, g_spec, g_rgb)]
        d_light = torch.zeros((), dtype=torch.float32, device=dev)
        d_dist = torch.empty(M, 1, dtype=torch.float32, device=dev)
        d_alpha = torch.empty(M, 1, dtype=torch.float32, device=dev)
        d_normal, d_kd, d_ks = (torch.empty(M, 3, dtype=torch.float32, device=dev) for _ in range(3))
        d_view = torch.empty(M, 3, dtype=torch.float32, device=dev) if ctx.needs_input_grad[3] else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ironb_ggx_bwd(_lib.ptr(light), _lib.ptr(dist), _lib.ptr(normal), _lib.ptr(viewdir),
                                                 _lib.ptr(kd), _lib.ptr(ks), _lib.ptr(alpha), _lib.ptr(trans),
                                                 _lib.ptr(diff_trans), M, _lib.ptr(g[0]), _lib.ptr(g[1]), _lib.ptr(g[2]),
                                                 _lib.ptr(d_light), _lib.ptr(d_dist), _lib.ptr(d_normal), _lib.ptr(d_view),
                                                 _lib.ptr(d_kd), _lib.ptr(d_ks), _lib.ptr(d_alpha), _lib.stream()),
                       "ggx_bwd")
        return d_light, d_dist, d_normal, d_view, d_kd, d_ks, d_alpha, None, None


class GGXColocatedRenderer(nn.Module):
    def __init__(self, use_cuda=False):
        super().__init__()
        t = np.load(_TABLES)
        self.MTS_TRANS = torch.from_numpy(t["ext_rtrans"].astype(np.float32))        # 5000 entries, external IOR
        self.MTS_DIFF_TRANS = torch.from_numpy(t["int_diff_rtrans"].astype(np.float32))  # 50 entries, internal IOR
        self.num_theta_samples = 100
        self.num_alpha_samples = 50
        if use_cuda:
            self.MTS_TRANS = self.MTS_TRANS.cuda()
            self.MTS_DIFF_TRANS = self.MTS_DIFF_TRANS.cuda()

    def _tables(self, dev):
        if self.MTS_TRANS.device != dev:   # the reference moves them on every call (:84-85)
            self.MTS_TRANS = self.MTS_TRANS.to(dev)
            self.MTS_DIFF_TRANS = self.MTS_DIFF_TRANS.to(dev)
        return self.MTS_TRANS, self.MTS_DIFF_TRANS

    def forward(self, light, distance, normal, viewdir, params={}):
        """light: scalar tensor; distance [...,1]; normal, viewdir [...,3]; params: diffuse_albedo [...,3],
        specular_albedo [...,3], specular_roughness [...,1].  Returns diffuse_rgb / specular_rgb / rgb [...,3]."""
        if not normal.is_cuda:
            raise RuntimeError("iron_b200.GGXColocatedRenderer runs on CUDA only (no CPU path)")
        dev = normal.device
        sh = list(normal.shape[:-1])
        trans, diff_trans = self._tables(dev)
        if not torch.is_tensor(light):
            light = torch.tensor(float(light), dtype=torch.float32, device=dev)
        light = _lib.f32c(light.reshape(()))
        full = lambda t, w: _lib.f32c(t.expand(sh + [w]).reshape(-1, w))
        outs = _GGXShade.apply(light, full(distance, 1), full(normal, 3), full(viewdir, 3),
                               full(params["diffuse_albedo"], 3), full(params["specular_albedo"], 3),
                               full(params["specular_roughness"], 1), trans, diff_trans)
        names = ("diffuse_rgb", "specular_rgb", "rgb")
        return {k: o.reshape(sh + [3]) for k, o in zip(names, outs)}


class _CompositeShade(torch.autograd.Function):
    @staticmethod
    def forward(ctx, light, dist, normal, viewdir, kd, ks, alpha, meta, mk, deta, trans, diff_trans):
        ctx.set_materialize_grads(False)
        M = dist.shape[0]
        dev = dist.device
        outs = [torch.empty(M, 3, dtype=torch.float32, device=dev) for _ in range(4)]     # rgb, specular, metallic, dielectric
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ironb_composite_fwd(*[_lib.ptr(t) for t in (light, dist, normal, viewdir, kd, ks, alpha, meta, mk,
                                                                               deta, trans, diff_trans)], M,
                                                       *[_lib.ptr(o) for o in outs], _lib.stream()), "composite_fwd")
        ctx.save_for_backward(light, dist, normal, viewdir, kd, ks, alpha, meta, mk, deta, trans, diff_trans)
        return tuple(outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_rgb, g_spec, g_met, g_diel):
        saved = ctx.saved_tensors
        dist = saved[1]
        M, dev = dist.shape[0], dist.device
        g = [None if t is None else _lib.f32c(t) for t in (g_rgb, g_spec, g_met, g_diel)]
        d_light = torch.zeros((), dtype=torch.float32, device=dev)
        d1 = lambda: torch.empty(M, 1, dtype=torch.float32, device=dev)
        d3 = lambda: torch.empty(M, 3, dtype=torch.float32, device=dev)
        d_dist, d_normal, d_kd, d_ks, d_alpha, d_meta, d_mk, d_deta = d1(), d3(), d3(), d3(), d1(), d1(), d1(), d1()
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ironb_composite_bwd(*[_lib.ptr(t) for t in saved], M, *[_lib.ptr(t) for t in g],
                                                       _lib.ptr(d_light), _lib.ptr(d_dist), _lib.ptr(d_normal), _lib.ptr(d_kd),
                                                       _lib.ptr(d_ks), _lib.ptr(d_alpha), _lib.ptr(d_meta), _lib.ptr(d_mk),
                                                       _lib.ptr(d_deta), _lib.stream()), "composite_bwd")
        return d_light, d_dist, d_normal, None, d_kd, d_ks, d_alpha, d_meta, d_mk, d_deta, None, None


class CompositeRenderer(GGXColocatedRenderer):
    """models/renderer_ggx.py:520-858 (`forward` only: the rough-plastic diffuse lobe + exact-Fresnel conductor lobe +
    dielectric microfacet lobe of the "comp2" renderer, model_bed.py:227-298).  The reference's ctor also globs spectral IOR
    files from ./resource/ior (absent from its own tree; `forward` never reads them): not loaded here.

    Kept quirks: D is evaluated with eta in the roughness slot (:803); the metallic / dielectric mixing weights are
    overwritten (:828-830) and therefore receive no gradient; rgb is accumulated in place into the diffuse tensor
    (:846-851), so the returned "diffuse_rgb" is the SAME tensor as "rgb"."""

    def forward(self, light, distance, normal, viewdir, params={}, use_env_light=False):
        if use_env_light:
            raise NotImplementedError("iron_b200.CompositeRenderer: use_env_light=True is not built (no driver of the reference sets it)")
        if not normal.is_cuda:
            raise RuntimeError("iron_b200.CompositeRenderer runs on CUDA only (no CPU path)")
        dev = normal.device
        sh = list(normal.shape[:-1])
        trans, diff_trans = self._tables(dev)
        if not torch.is_tensor(light):
            light = torch.tensor(float(light), dtype=torch.float32, device=dev)
        light = _lib.f32c(light.reshape(()))
        full = lambda t, w: _lib.f32c(t.expand(sh + [w]).reshape(-1, w))
        rgb, spec, met, diel = _CompositeShade.apply(
            light, full(distance, 1), full(normal, 3), full(viewdir, 3), full(params["diffuse_albedo"], 3),
            full(params["specular_albedo"], 3), full(params["specular_roughness"], 1), full(params["metallic_eta"], 1),
            full(params["metallic_k"], 1), full(params["dielectric_eta"], 1), trans, diff_trans)
        rgb = rgb.reshape(sh + [3])
        return {"diffuse_rgb": rgb, "specular_rgb": spec.reshape(sh + [3]), "metallic_rgb": met.reshape(sh + [3]),
                "dielectric_rgb": diel.reshape(sh + [3]), "rgb": rgb}
