"""Stage-1 NeuS volume renderer (second half): `NeuSRenderer` of models/renderer.py:128-453 with the
reference's constructor, methods and return dictionaries, on this library's kernels.

Per training batch the reference evaluates, for every ray: 64 + 3 x 16 SDF values for the hierarchical sampling (`up_sample`,
`cat_z_vals`, no gradients), then at the 128 section midpoints the SDF network, its gradient (with the double backward) and
the colour network, then ~45 elementwise / scan launches of alpha compositing (and as many again under autograd).  Here:

  * the SDF values / features / gradients come from ONE `SDFNetwork.get_all` call (`ironb_sdf_getall_fwd / _bwd`: the reference
    calls `sdf_network(pts)` and `sdf_network.gradient(pts)`, i.e. runs the network twice);
  * the colour network is `RenderingNetwork` with its skip layer (`ironb_matnet_fwd / _bwd`);
  * everything from `true_cos` to `color` / `weights` / `gradient_error` and its backward is `ironb_neus_composite_fwd / _bwd`
    (csrc/neus.cu): one launch each way;
  * the background NeRF (`n_outside > 0`) is a plain-`nn.Linear` MLP in the reference and stays one here (`NeRF` below: cuBLAS
    fp32 GEMMs through ATen -- it is outside the SDF path this library accelerates);
  * the hierarchical sampling (`up_sample`, `sample_pdf`, `cat_z_vals`) is the reference's sequence of small tensor ops on
    [rays, <= 128] arrays under `no_grad`, with the SDF evaluations on the tensor-core forward.

Random numbers: `render` draws `torch.rand` exactly where the reference does (:378, :384) on the rays' device.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .embedder import get_embedder


class SingleVarianceNetwork(nn.Module):
    """models/fields.py:415-421."""

    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones([len(x), 1], device=self.variance.device) * torch.exp(self.variance * 10.0)


class NeRF(nn.Module):
    """The background model, models/fields.py:241-322 (use_viewdirs=True): plain Linear layers, ReLU, one skip; same module
    names and state-dict keys (`pts_linears.{i}`, `views_linears.0`, `feature_linear`, `alpha_linear`, `rgb_linear`)."""

    def __init__(self, D=8, W=256, d_in=3, d_in_view=3, multires=0, multires_view=0, output_ch=4, skips=(4,), use_viewdirs=False):
        super().__init__()
        self.D, self.W, self.d_in, self.d_in_view = D, W, d_in, d_in_view
        self.input_ch, self.input_ch_view = 3, 3
        self.embed_fn, self.embed_fn_view = None, None
        if multires > 0:
            self.embed_fn, self.input_ch = get_embedder(multires, input_dims=d_in)
        if multires_view > 0:
            self.embed_fn_view, self.input_ch_view = get_embedder(multires_view, input_dims=d_in_view)
        self.skips = list(skips)
        self.use_viewdirs = use_viewdirs
        self.pts_linears = nn.ModuleList(
            [nn.Linear(self.input_ch, W)]
            + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + self.input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(self.input_ch_view + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)

    def forward(self, input_pts, input_views):
        if self.embed_fn is not None:
            input_pts = self.embed_fn(input_pts)
        if self.embed_fn_view is not None:
            input_views = self.embed_fn_view(input_views)
        h = input_pts
        for i in range(len(self.pts_linears)):
            h = F.relu(self.pts_linears[i](h))
            if i in self.skips:
                h = torch.cat([input_pts, h], -1)
        assert self.use_viewdirs, "NeRF: use_viewdirs=False has no forward in the reference either (fields.py:321)"
        alpha = self.alpha_linear(h)
        h = torch.cat([self.feature_linear(h), input_views], -1)
        for i in range(len(self.views_linears)):
            h = F.relu(self.views_linears[i](h))
        return alpha, self.rgb_linear(h)


def sample_pdf(bins, weights, n_samples, det=False):
    """models/renderer.py:43-73 (inverse-CDF sampling, from NeRF)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if det:
        u = torch.linspace(0.0 + 0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples, device=bins.device)
        u = u.expand(list(cdf.shape[:-1]) + [n_samples])
    else:
        u = torch.rand(list(cdf.shape[:-1]) + [n_samples], device=bins.device)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bins_b, bins_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bins_b + t * (bins_a - bins_b)


class _NeusComposite(torch.autograd.Function):
    """(color, weights, cdf, inside_sphere, gradient_error) of render_core from per-section inputs; csrc/neus.cu."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, mid_z, dists, sdf, grad, color, inv_s, bg_alpha, bg_color, bg_rgb, anneal):
        lib = _lib.load()
        N, n = mid_z.shape
        n_tot = bg_alpha.shape[1] if bg_alpha is not None else n
        dev = mid_z.device
        f = _lib.f32c
        args = [f(rays_o), f(rays_d), f(mid_z), f(dists), f(sdf.reshape(N, n)), f(grad.reshape(N, n, 3)), f(color.reshape(N, n, 3)),
                f(inv_s.reshape(1)), None if bg_alpha is None else f(bg_alpha), None if bg_color is None else f(bg_color),
                None if bg_rgb is None else f(bg_rgb.reshape(3))]
        out_color = torch.empty(N, 3, dtype=torch.float32, device=dev)
        weights = torch.empty(N, n_tot, dtype=torch.float32, device=dev)
        cdf = torch.empty(N, n, dtype=torch.float32, device=dev)
        inside = torch.empty(N, n, dtype=torch.float32, device=dev)
        acc = torch.empty(2, dtype=torch.float32, device=dev)
        gerr = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ironb_neus_composite_fwd(*[_lib.ptr(a) for a in args], N, n, n_tot, float(anneal), _lib.ptr(out_color),
                                                    _lib.ptr(weights), _lib.ptr(cdf), _lib.ptr(inside), _lib.ptr(acc),
                                                    _lib.ptr(gerr), _lib.stream()), "neus_composite_fwd")
        ctx.args, ctx.dims, ctx.anneal = args, (N, n, n_tot), float(anneal)
        ctx.save_for_backward(weights, acc)
        ctx.shapes = (sdf.shape, grad.shape, color.shape, inv_s.shape)
        ctx.mark_non_differentiable(cdf, inside)
        return out_color, weights, cdf, inside, gerr

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_color, d_weights, _dcdf, _dinside, d_gerr):
        lib = _lib.load()
        weights, acc = ctx.saved_tensors
        N, n, n_tot = ctx.dims
        dev = weights.device
        args = ctx.args
        has_bg = args[8] is not None
        d_sdf = torch.empty(N, n, dtype=torch.float32, device=dev)
        d_grad = torch.empty(N, n, 3, dtype=torch.float32, device=dev)
        d_colors = torch.empty(N, n, 3, dtype=torch.float32, device=dev)
        d_inv_s = torch.empty(1, dtype=torch.float32, device=dev)
        d_bga = torch.empty(N, n_tot, dtype=torch.float32, device=dev) if has_bg else None
        d_bgc = torch.empty(N, n_tot, 3, dtype=torch.float32, device=dev) if has_bg else None
        g = lambda t: None if t is None else _lib.f32c(t)
        with torch.cuda.device(dev):
            _lib.check(lib.ironb_neus_composite_bwd(*[_lib.ptr(a) for a in args], N, n, n_tot, ctx.anneal, _lib.ptr(weights),
                                                    _lib.ptr(acc), _lib.ptr(g(d_color)), _lib.ptr(g(d_weights)),
                                                    _lib.ptr(None if d_gerr is None else g(d_gerr).reshape(1)), _lib.ptr(d_sdf),
                                                    _lib.ptr(d_grad), _lib.ptr(d_colors), _lib.ptr(d_inv_s), _lib.ptr(d_bga),
                                                    _lib.ptr(d_bgc), _lib.stream()), "neus_composite_bwd")
        s_sdf, s_grad, s_color, s_inv = ctx.shapes
        return (None, None, None, None, d_sdf.reshape(s_sdf), d_grad.reshape(s_grad), d_colors.reshape(s_color),
                d_inv_s.reshape(s_inv), d_bga, d_bgc, None, None)


class NeuSRenderer:
    """models/renderer.py:128-453: same constructor, `render(rays_o, rays_d, near, far, perturb_overwrite=-1,
    background_rgb=None, cos_anneal_ratio=0.0)` and return keys."""

    def __init__(self, nerf, sdf_network, deviation_network, color_network, n_samples, n_importance, n_outside, up_sample_steps,
                 perturb):
        self.nerf = nerf
        self.sdf_network = sdf_network
        self.deviation_network = deviation_network
        self.color_network = color_network
        self.n_samples = n_samples
        self.n_importance = n_importance
        self.n_outside = n_outside
        self.up_sample_steps = up_sample_steps
        self.perturb = perturb
        self.rand_fn = None          # tests inject the uniform numbers the reference drew; None = torch.rand on the rays' device

    # ---- background model (:140-190) ------------------------------------------------------------------------------
    def render_core_outside(self, rays_o, rays_d, z_vals, sample_dist, nerf, background_rgb=None):
        batch_size, n_samples = z_vals.shape
        dists = z_vals[..., 1:] - z_vals[..., :-1]
        dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)
        mid_z_vals = z_vals + dists * 0.5
        pts = rays_o[:, None, :] + rays_d[:, None, :] * mid_z_vals[..., :, None]
        dis_to_center = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).clip(1.0, 1e10)
        pts = torch.cat([pts / dis_to_center, 1.0 / dis_to_center], dim=-1)
        dirs = rays_d[:, None, :].expand(batch_size, n_samples, 3)
        pts = pts.reshape(-1, 3 + int(self.n_outside > 0))
        dirs = dirs.reshape(-1, 3)
        density, sampled_color = nerf(pts, dirs)
        alpha = 1.0 - torch.exp(-F.softplus(density.reshape(batch_size, n_samples)) * dists)
        alpha = alpha.reshape(batch_size, n_samples)
        weights = alpha * torch.cumprod(torch.cat([torch.ones([batch_size, 1], device=alpha.device), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
        sampled_color = sampled_color.reshape(batch_size, n_samples, 3)
        color = (weights[:, :, None] * sampled_color).sum(dim=1)
        if background_rgb is not None:
            color = color + background_rgb * (1.0 - weights.sum(dim=-1, keepdim=True))
        return {"color": color, "sampled_color": sampled_color, "alpha": alpha, "weights": weights}

    # ---- hierarchical sampling (:192-246), no gradients -----------------------------------------------------------------
    def up_sample(self, rays_o, rays_d, z_vals, sdf, n_importance, inv_s):
        batch_size, n_samples = z_vals.shape
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
        radius = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=False)
        inside_sphere = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
        sdf = sdf.reshape(batch_size, n_samples)
        prev_sdf, next_sdf = sdf[:, :-1], sdf[:, 1:]
        prev_z_vals, next_z_vals = z_vals[:, :-1], z_vals[:, 1:]
        mid_sdf = (prev_sdf + next_sdf) * 0.5
        cos_val = (next_sdf - prev_sdf) / (next_z_vals - prev_z_vals + 1e-5)
        prev_cos_val = torch.cat([torch.zeros([batch_size, 1], device=z_vals.device), cos_val[:, :-1]], dim=-1)
        cos_val = torch.minimum(prev_cos_val, cos_val)
        cos_val = cos_val.clip(-1e3, 0.0) * inside_sphere
        dist = next_z_vals - prev_z_vals
        prev_esti_sdf = mid_sdf - cos_val * dist * 0.5
        next_esti_sdf = mid_sdf + cos_val * dist * 0.5
        prev_cdf = torch.sigmoid(prev_esti_sdf * inv_s)
        next_cdf = torch.sigmoid(next_esti_sdf * inv_s)
        alpha = (prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)
        weights = alpha * torch.cumprod(torch.cat([torch.ones([batch_size, 1], device=z_vals.device), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
        return sample_pdf(z_vals, weights, n_importance, det=True).detach()

    def cat_z_vals(self, rays_o, rays_d, z_vals, new_z_vals, sdf, last=False):
        batch_size, n_samples = z_vals.shape
        _, n_importance = new_z_vals.shape
        pts = rays_o[:, None, :] + rays_d[:, None, :] * new_z_vals[..., :, None]
        z_vals = torch.cat([z_vals, new_z_vals], dim=-1)
        z_vals, index = torch.sort(z_vals, dim=-1)
        if not last:
            new_sdf = self.sdf_network.sdf(pts.reshape(-1, 3)).reshape(batch_size, n_importance)
            sdf = torch.cat([sdf, new_sdf], dim=-1)
            sdf = torch.gather(sdf, -1, index)
        return z_vals, sdf

    # ---- the differentiable core (:248-351) ---------------------------------------------------------------------------------
    def render_core(self, rays_o, rays_d, z_vals, sample_dist, sdf_network, deviation_network, color_network,
                    background_alpha=None, background_sampled_color=None, background_rgb=None, cos_anneal_ratio=0.0):
        batch_size, n_samples = z_vals.shape
        dists = z_vals[..., 1:] - z_vals[..., :-1]
        dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)
        mid_z_vals = z_vals + dists * 0.5
        pts = rays_o[:, None, :] + rays_d[:, None, :] * mid_z_vals[..., :, None]
        dirs = rays_d[:, None, :].expand(pts.shape)
        pts = pts.reshape(-1, 3)
        dirs = dirs.reshape(-1, 3)
        # sdf_network(pts) + sdf_network.gradient(pts) of the reference (:270-274) in one evaluation
        sdf, feature_vector, gradients = sdf_network.get_all(pts, is_training=torch.is_grad_enabled())
        sampled_color = color_network(pts, gradients, dirs, feature_vector).reshape(batch_size, n_samples, 3)
        inv_s = deviation_network(torch.zeros([1, 3], device=pts.device))[:, :1].clip(1e-6, 1e6)     # single parameter
        color, weights, cdf, inside_sphere, gradient_error = _NeusComposite.apply(
            rays_o, rays_d, mid_z_vals, dists, sdf, gradients, sampled_color, inv_s, background_alpha, background_sampled_color,
            background_rgb, float(cos_anneal_ratio))
        return {
            "color": color,
            "sdf": sdf,
            "dists": dists,
            "gradients": gradients.reshape(batch_size, n_samples, 3),
            "s_val": 1.0 / inv_s.expand(batch_size * n_samples, 1),
            "mid_z_vals": mid_z_vals,
            "weights": weights,
            "cdf": cdf,
            "gradient_error": gradient_error,
            "inside_sphere": inside_sphere,
        }

    def render(self, rays_o, rays_d, near, far, perturb_overwrite=-1, background_rgb=None, cos_anneal_ratio=0.0):
        dev = rays_o.device
        batch_size = len(rays_o)
        sample_dist = 2.0 / self.n_samples
        z_vals = torch.linspace(0.0, 1.0, self.n_samples, device=dev)
        z_vals = near + (far - near) * z_vals[None, :]
        z_vals_outside = None
        if self.n_outside > 0:
            z_vals_outside = torch.linspace(1e-3, 1.0 - 1.0 / (self.n_outside + 1.0), self.n_outside, device=dev)
        n_samples = self.n_samples
        perturb = self.perturb
        if perturb_overwrite >= 0:
            perturb = perturb_overwrite
        rand = self.rand_fn if self.rand_fn is not None else (lambda shape: torch.rand(shape, device=dev))
        if perturb > 0:
            t_rand = rand([batch_size, 1]) - 0.5
            z_vals = z_vals + t_rand * 2.0 / self.n_samples
            if self.n_outside > 0:
                mids = 0.5 * (z_vals_outside[..., 1:] + z_vals_outside[..., :-1])
                upper = torch.cat([mids, z_vals_outside[..., -1:]], -1)
                lower = torch.cat([z_vals_outside[..., :1], mids], -1)
                t_rand = rand([batch_size, z_vals_outside.shape[-1]])
                z_vals_outside = lower[None, :] + (upper - lower)[None, :] * t_rand
        if self.n_outside > 0:
            z_vals_outside = far / torch.flip(z_vals_outside, dims=[-1]) + 1.0 / self.n_samples
        background_alpha = None
        background_sampled_color = None
        if self.n_importance > 0:
            with torch.no_grad():
                pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
                sdf = self.sdf_network.sdf(pts.reshape(-1, 3)).reshape(batch_size, self.n_samples)
                for i in range(self.up_sample_steps):
                    new_z_vals = self.up_sample(rays_o, rays_d, z_vals, sdf, self.n_importance // self.up_sample_steps, 64 * 2 ** i)
                    z_vals, sdf = self.cat_z_vals(rays_o, rays_d, z_vals, new_z_vals, sdf, last=(i + 1 == self.up_sample_steps))
            n_samples = self.n_samples + self.n_importance
        if self.n_outside > 0:
            z_vals_feed = torch.cat([z_vals, z_vals_outside], dim=-1)
            z_vals_feed, _ = torch.sort(z_vals_feed, dim=-1)
            ret_outside = self.render_core_outside(rays_o, rays_d, z_vals_feed, sample_dist, self.nerf)
            background_sampled_color = ret_outside["sampled_color"]
            background_alpha = ret_outside["alpha"]
        ret_fine = self.render_core(rays_o, rays_d, z_vals, sample_dist, self.sdf_network, self.deviation_network,
                                    self.color_network, background_rgb=background_rgb, background_alpha=background_alpha,
                                    background_sampled_color=background_sampled_color, cos_anneal_ratio=cos_anneal_ratio)
        weights = ret_fine["weights"]
        # s_val over the sections the reference averages over: (n_samples) of them, all equal to 1 / inv_s
        s_val = ret_fine["s_val"].reshape(batch_size, n_samples).mean(dim=-1, keepdim=True)
        return {
            "color_fine": ret_fine["color"],
            "s_val": s_val,
            "cdf_fine": ret_fine["cdf"],
            "weight_sum": weights.sum(dim=-1, keepdim=True),
            "weight_max": torch.max(weights, dim=-1, keepdim=True)[0],
            "gradients": ret_fine["gradients"],
            "weights": weights,
            "gradient_error": ret_fine["gradient_error"],
            "inside_sphere": ret_fine["inside_sphere"],
        }
