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


class _TCLinear(torch.autograd.Function):
    """y = act(x W^T + b) for a plain nn.Linear on the tcgen05 GEMMs (3xTF32, fp32-grade): `ironb_linear_fwd`; backward =
    ReLU mask (`ironb_relu_mask`), input gradient `ironb_gemm_nt(dy, W^T)`, weight + bias gradient `ironb_linear_wgrad`."""

    @staticmethod
    def forward(ctx, x, W, b, relu):
        lib = _lib.load()
        x, Wc = _lib.f32c(x), _lib.f32c(W)
        M, K = x.shape
        N = Wc.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.ironb_linear_fwd(_lib.ptr(x), K, _lib.ptr(Wc), K, _lib.ptr(None if b is None else _lib.f32c(b)), M, N, K,
                                            int(relu), _lib.ptr(y), N, _lib.stream()), "linear_fwd")
        ctx.relu = bool(relu)
        ctx.save_for_backward(x, Wc, y if relu else None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        lib = _lib.load()
        x, W, y = ctx.saved_tensors
        M, K = x.shape
        N = W.shape[0]
        dev = x.device
        dy = _lib.f32c(dy)
        with torch.cuda.device(dev):
            if ctx.relu:
                dym = torch.empty_like(dy)
                _lib.check(lib.ironb_relu_mask(_lib.ptr(dy), _lib.ptr(y), M * N, _lib.ptr(dym), _lib.stream()), "relu_mask")
            else:
                dym = dy
            dx = None
            if ctx.needs_input_grad[0]:
                Wt = W.t().contiguous()
                dx = torch.empty(M, K, dtype=torch.float32, device=dev)
                _lib.check(lib.ironb_gemm_nt(_lib.ptr(dym), N, _lib.ptr(Wt), N, M, K, N, _lib.ptr(dx), K, _gemm_nt_mode(lib),
                                             _lib.stream()), "linear dgrad")
            dW = torch.zeros(N, K, dtype=torch.float32, device=dev)
            db = torch.zeros(N, dtype=torch.float32, device=dev) if ctx.needs_input_grad[2] else None
            scratch = torch.empty(max(int(lib.ironb_gemm_tn_scratch_bytes(M, N, K)), 16), dtype=torch.uint8, device=dev)
            _lib.check(lib.ironb_linear_wgrad(_lib.ptr(dym), N, _lib.ptr(x), K, M, N, K, _lib.ptr(dW), K, _lib.ptr(db),
                                              _lib.ptr(scratch), _lib.stream()), "linear_wgrad")
        return dx, dW, db, None


def _gemm_nt_mode(lib):
    """ironb_gemm_nt's mode argument for the library's current GEMM arithmetic: 0 = FFMA, 2 = 3xTF32 on tcgen05."""
    cur = lib.ironb_set_gemm_mode(2)
    lib.ironb_set_gemm_mode(cur)
    return 0 if cur == 0 else 2


def _linear(layer, h, relu):
    """A plain nn.Linear (+ ReLU): on this library's GEMMs when the shapes allow (fan-in and fan-out multiples of 4), else ATen
    (the 1- and 3-wide heads and the 283-wide view layer of the NeRF: < 3 % of its FLOPs)."""
    W = layer.weight
    if h.is_cuda and W.shape[0] % 4 == 0 and W.shape[1] % 4 == 0 and h.dim() == 2:
        return _TCLinear.apply(h, W, layer.bias, relu)
    out = layer(h)
    return F.relu(out) if relu else out


class NeRF(nn.Module):
    """The background model, models/fields.py:241-322 (use_viewdirs=True): plain Linear layers, ReLU, one skip; same module
    names and state-dict keys (`pts_linears.{i}`, `views_linears.0`, `feature_linear`, `alpha_linear`, `rgb_linear`).  The
    256-wide layers run on the tensor cores (`_TCLinear`)."""

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
            h = _linear(self.pts_linears[i], h, True)
            if i in self.skips:
                h = torch.cat([input_pts, h], -1)
        assert self.use_viewdirs, "NeRF: use_viewdirs=False has no forward in the reference either (fields.py:321)"
        alpha = _linear(self.alpha_linear, h, False)
        h = torch.cat([_linear(self.feature_linear, h, False), input_views], -1)
        for i in range(len(self.views_linears)):
            h = _linear(self.views_linears[i], h, True)
        return alpha, _linear(self.rgb_linear, h, False)


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


def _cf(t):
    return _lib.f32c(t)


def ray_sections(rays_o, rays_d, z_vals, sample_dist, outside=False, want_dirs=True):
    """(dists [N,n], mid_z [N,n], pts [N*n, 3 or 4], dirs [N*n,3] or None) of a ray batch in one launch (`ironb_neus_sections`):
    section lengths (last = sample_dist), midpoints, the midpoints' positions -- for the background model the inverted-sphere
    points (x / max(|x|,1), 1 / max(|x|,1)) -- and the per-point view directions (models/renderer.py:254-262, :145-160)."""
    N, n = z_vals.shape
    dev = z_vals.device
    z = _cf(z_vals)
    dists = torch.empty(N, n, dtype=torch.float32, device=dev)
    mid = torch.empty(N, n, dtype=torch.float32, device=dev)
    pts = torch.empty(N * n, 4 if outside else 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(N * n, 3, dtype=torch.float32, device=dev) if want_dirs else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().ironb_neus_sections(_lib.ptr(_cf(rays_o)), _lib.ptr(_cf(rays_d)), _lib.ptr(z), N, n, float(sample_dist),
                                                   int(outside), _lib.ptr(dists), _lib.ptr(mid), _lib.ptr(pts), _lib.ptr(dirs),
                                                   _lib.stream()), "neus_sections")
    return dists, mid, pts, dirs


class _NeusComposite(torch.autograd.Function):
    """(color, weights, cdf, inside_sphere, gradient_error) of render_core from per-section inputs; csrc/neus.cu.  The background
    enters as the NeRF's raw density + section lengths + colour; its alpha is formed inside the kernel."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, mid_z, dists, sdf, grad, color, inv_s, bg_density, bg_dists, bg_color, bg_rgb, anneal):
        lib = _lib.load()
        N, n = mid_z.shape
        n_tot = bg_density.shape[1] if bg_density is not None else n
        dev = mid_z.device
        opt = lambda t: None if t is None else _cf(t)
        args = [_cf(rays_o), _cf(rays_d), _cf(mid_z), _cf(dists), _cf(sdf.reshape(N, n)), _cf(grad.reshape(N, n, 3)),
                _cf(color.reshape(N, n, 3)), _cf(inv_s.reshape(1)), opt(bg_density), opt(bg_dists), opt(bg_color),
                None if bg_rgb is None else _cf(bg_rgb.reshape(3))]
        new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        out_color, weights, cdf, inside, acc, gerr = new(N, 3), new(N, n_tot), new(N, n), new(N, n), new(2), new(())
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
        new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        d_sdf, d_grad, d_colors, d_inv_s = new(N, n), new(N, n, 3), new(N, n, 3), new(1)
        d_bgd = new(N, n_tot) if has_bg else None
        d_bgc = new(N, n_tot, 3) if has_bg else None
        opt = lambda t: None if t is None else _cf(t)
        with torch.cuda.device(dev):
            _lib.check(lib.ironb_neus_composite_bwd(*[_lib.ptr(a) for a in args], N, n, n_tot, ctx.anneal, _lib.ptr(weights),
                                                    _lib.ptr(acc), _lib.ptr(opt(d_color)), _lib.ptr(opt(d_weights)),
                                                    _lib.ptr(None if d_gerr is None else _cf(d_gerr).reshape(1)), _lib.ptr(d_sdf),
                                                    _lib.ptr(d_grad), _lib.ptr(d_colors), _lib.ptr(d_inv_s), _lib.ptr(d_bgd),
                                                    _lib.ptr(d_bgc), _lib.stream()), "neus_composite_bwd")
        s_sdf, s_grad, s_color, s_inv = ctx.shapes
        return (None, None, None, None, d_sdf.reshape(s_sdf), d_grad.reshape(s_grad), d_colors.reshape(s_color),
                d_inv_s.reshape(s_inv), d_bgd, None, d_bgc, None, None)


class NeuSRenderer:
    """models/renderer.py:128-453: same constructor, `render(rays_o, rays_d, near, far, perturb_overwrite=-1,
    background_rgb=None, cos_anneal_ratio=0.0)` and return keys; `up_sample`, `cat_z_vals`, `render_core`,
    `render_core_outside` keep their signatures."""

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
    def _outside(self, rays_o, rays_d, z_vals, sample_dist, nerf):
        """Raw density [N,n], section lengths [N,n] and colour [N,n,3] of the background model on the given sections."""
        N, n = z_vals.shape
        dists, _, pts, dirs = ray_sections(rays_o, rays_d, z_vals, sample_dist, outside=self.n_outside > 0)
        density, sampled_color = nerf(pts, dirs)
        return density.reshape(N, n), dists, sampled_color.reshape(N, n, 3)

    def render_core_outside(self, rays_o, rays_d, z_vals, sample_dist, nerf, background_rgb=None):
        density, dists, sampled_color = self._outside(rays_o, rays_d, z_vals, sample_dist, nerf)
        alpha = 1.0 - torch.exp(-F.softplus(density) * dists)
        trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
        weights = alpha * trans
        color = (weights[:, :, None] * sampled_color).sum(dim=1)
        if background_rgb is not None:
            color = color + background_rgb * (1.0 - weights.sum(dim=-1, keepdim=True))
        return {"color": color, "sampled_color": sampled_color, "alpha": alpha, "weights": weights}

    # ---- hierarchical sampling (:192-246), no gradients -----------------------------------------------------------------
    def _up_sample(self, rays_o, rays_d, z_vals, sdf, n_importance, inv_s, want_pts):
        N, n = z_vals.shape
        dev = z_vals.device
        new_z = torch.empty(N, n_importance, dtype=torch.float32, device=dev)
        pts = torch.empty(N * n_importance, 3, dtype=torch.float32, device=dev) if want_pts else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ironb_neus_upsample(_lib.ptr(_cf(rays_o)), _lib.ptr(_cf(rays_d)), _lib.ptr(_cf(z_vals)),
                                                       _lib.ptr(_cf(sdf.reshape(N, n))), N, n, int(n_importance), float(inv_s),
                                                       _lib.ptr(new_z), _lib.ptr(pts), _lib.stream()), "neus_upsample")
        return new_z, pts

    def up_sample(self, rays_o, rays_d, z_vals, sdf, n_importance, inv_s):
        """n_importance new samples per ray at the quantiles of the interval weights under the fixed sharpness inv_s: one
        launch (`ironb_neus_upsample`) for the reference's alpha / cumprod / sample_pdf(det=True) sequence."""
        return self._up_sample(rays_o, rays_d, z_vals, sdf, n_importance, inv_s, False)[0]

    def cat_z_vals(self, rays_o, rays_d, z_vals, new_z_vals, sdf, last=False, new_pts=None):
        """Sorted union of the two ascending sample lists (a per-ray merge, `ironb_neus_merge`); unless `last`, the SDF is
        evaluated at the new samples and travels with them."""
        N, n = z_vals.shape
        m = new_z_vals.shape[1]
        dev = z_vals.device
        new_sdf = None
        if not last:
            if new_pts is None:
                new_pts = (rays_o[:, None, :] + rays_d[:, None, :] * new_z_vals[..., :, None]).reshape(-1, 3)
            new_sdf = _cf(self.sdf_network.sdf(new_pts).reshape(N, m))
        out_z = torch.empty(N, n + m, dtype=torch.float32, device=dev)
        out_sdf = torch.empty(N, n + m, dtype=torch.float32, device=dev) if not last else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ironb_neus_merge(_lib.ptr(_cf(z_vals)), _lib.ptr(None if last else _cf(sdf.reshape(N, n))), n,
                                                    _lib.ptr(_cf(new_z_vals)), _lib.ptr(new_sdf), m, N, _lib.ptr(out_z),
                                                    _lib.ptr(out_sdf), _lib.stream()), "neus_merge")
        return out_z, (sdf if last else out_sdf)

    # ---- the differentiable core (:248-351) ---------------------------------------------------------------------------------
    def _core(self, rays_o, rays_d, z_vals, sample_dist, sdf_network, deviation_network, color_network, bg, background_rgb,
              cos_anneal_ratio):
        N, n = z_vals.shape
        dists, mid_z_vals, pts, dirs = ray_sections(rays_o, rays_d, z_vals, sample_dist)
        # sdf_network(pts) + sdf_network.gradient(pts) of the reference (:270-274) in one evaluation
        sdf, feature_vector, gradients = sdf_network.get_all(pts, is_training=torch.is_grad_enabled())
        sampled_color = color_network(pts, gradients, dirs, feature_vector)
        inv_s = deviation_network(torch.zeros([1, 3], device=pts.device))[:, :1].clip(1e-6, 1e6)     # single parameter
        bg_density, bg_dists, bg_color = bg if bg is not None else (None, None, None)
        color, weights, cdf, inside_sphere, gradient_error = _NeusComposite.apply(
            rays_o, rays_d, mid_z_vals, dists, sdf, gradients, sampled_color, inv_s, bg_density, bg_dists, bg_color,
            background_rgb, float(cos_anneal_ratio))
        return {"color": color, "sdf": sdf, "dists": dists, "gradients": gradients.reshape(N, n, 3),
                "s_val": 1.0 / inv_s.expand(N * n, 1), "mid_z_vals": mid_z_vals, "weights": weights, "cdf": cdf,
                "gradient_error": gradient_error, "inside_sphere": inside_sphere}

    def render_core(self, rays_o, rays_d, z_vals, sample_dist, sdf_network, deviation_network, color_network,
                    background_alpha=None, background_sampled_color=None, background_rgb=None, cos_anneal_ratio=0.0):
        """The reference's signature: `background_alpha` is an ALPHA in [0, 1) (what render_core_outside returns).  The kernel
        takes a density and a section length, so the alpha is passed as the density that reproduces it with length 1:
        alpha = 1 - exp(-softplus(x)), x = log(expm1(-log1p(-alpha)))."""
        bg = None
        if background_alpha is not None:
            tau = -torch.log1p(-background_alpha.clamp(max=1.0 - 1e-7))                  # optical depth
            density = torch.where(tau > 20.0, tau, torch.log(torch.expm1(tau.clamp(min=1e-30))))
            bg = (density, torch.ones_like(density), background_sampled_color)
        return self._core(rays_o, rays_d, z_vals, sample_dist, sdf_network, deviation_network, color_network, bg, background_rgb,
                          cos_anneal_ratio)

    def render(self, rays_o, rays_d, near, far, perturb_overwrite=-1, background_rgb=None, cos_anneal_ratio=0.0):
        dev = rays_o.device
        N = len(rays_o)
        ns, n_out = self.n_samples, self.n_outside
        sample_dist = 2.0 / ns                                  # the region of interest is the unit sphere
        rand = self.rand_fn if self.rand_fn is not None else (lambda shape: torch.rand(shape, device=dev))
        perturb = perturb_overwrite if perturb_overwrite >= 0 else self.perturb
        # ---- coarse samples: uniform in [near, far], jittered by one draw per ray (:356-379)
        z_vals = near + (far - near) * torch.linspace(0.0, 1.0, ns, device=dev)[None, :]
        if perturb > 0:
            z_vals = z_vals + (rand([N, 1]) - 0.5) * (2.0 / ns)
        # ---- outside samples: uniform in inverse depth beyond `far`, stratified jitter (:361-389)
        z_out = None
        if n_out > 0:
            z_out = torch.linspace(1e-3, 1.0 - 1.0 / (n_out + 1.0), n_out, device=dev)
            if perturb > 0:
                mids = 0.5 * (z_out[1:] + z_out[:-1])
                lower, upper = torch.cat([z_out[:1], mids]), torch.cat([mids, z_out[-1:]])
                z_out = lower[None, :] + (upper - lower)[None, :] * rand([N, n_out])
            z_out = far / torch.flip(z_out, dims=[-1]) + 1.0 / ns
        # ---- hierarchical sampling: up_sample_steps rounds of n_importance / steps new samples at sharpness 64 * 2^i (:394-419)
        n = ns
        if self.n_importance > 0:
            with torch.no_grad():
                # the SDF AT the coarse samples (not at section midpoints: :397-398)
                sdf = self.sdf_network.sdf((rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]).reshape(-1, 3)).reshape(N, ns)
                per = self.n_importance // self.up_sample_steps
                for i in range(self.up_sample_steps):
                    last = i + 1 == self.up_sample_steps
                    new_z, new_pts = self._up_sample(rays_o, rays_d, z_vals, sdf, per, 64 * 2 ** i, not last)
                    z_vals, sdf = self.cat_z_vals(rays_o, rays_d, z_vals, new_z, sdf, last=last, new_pts=new_pts)
            n = ns + self.n_importance
        # ---- background model on the union of all samples (:422-428), then the core
        bg = None
        if n_out > 0:
            z_feed, _ = torch.sort(torch.cat([z_vals, z_out], dim=-1), dim=-1)
            bg = self._outside(rays_o, rays_d, z_feed, sample_dist, self.nerf)
        fine = self._core(rays_o, rays_d, z_vals, sample_dist, self.sdf_network, self.deviation_network, self.color_network, bg,
                          background_rgb, cos_anneal_ratio)
        weights = fine["weights"]
        return {
            "color_fine": fine["color"],
            "s_val": fine["s_val"].reshape(N, n).mean(dim=-1, keepdim=True),
            "cdf_fine": fine["cdf"],
            "weight_sum": weights.sum(dim=-1, keepdim=True),
            "weight_max": torch.max(weights, dim=-1, keepdim=True)[0],
            "gradients": fine["gradients"],
            "weights": weights,
            "gradient_error": fine["gradient_error"],
            "inside_sphere": fine["inside_sphere"],
        }
