"""CPU oracle for the IRON surface-rendering hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch-CPU fp32 and as stateless functions over
explicit parameter dictionaries, the algorithm of the reference's hot path
(SURVEY.md section 8a).  It is NOT part of the product: only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product (`iron_b200/`) never does.

Parity status: PINNED.  `oracle/make_golden.py` imports the real reference
modules from /root/reference in the build container, runs them on seeded
inputs and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks
every function here against those files (masks exactly, floats to ~1e-6).

Every function cites the reference lines it follows (paths relative to
/root/reference).  Arithmetic is fp32; the oracle uses torch.autograd for
derivatives exactly where the reference does, and additionally carries the
closed-form input-gradient / double-backward of Appendix A so the
CUDA kernels' formulas can be checked on CPU.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

SQRT2 = float(np.sqrt(2))

# --------------------------------------------------------------------------
# positional encoding                     models/embedder.py:11-36, 39-54
# --------------------------------------------------------------------------


def posenc(x: Tensor, n_freqs: int) -> Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)].

    models/embedder.py:15-36 with include_input=True, log_sampling=True
    (freq_bands = 2 ** linspace(0, L-1, L), :23).
    """
    if n_freqs <= 0:
        return x
    bands = 2.0 ** torch.linspace(0.0, n_freqs - 1, n_freqs)
    out = [x]
    for f in bands:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, dim=-1)


def posenc_dim(n_freqs: int, d: int = 3) -> int:
    return d * (1 + 2 * n_freqs) if n_freqs > 0 else d


# --------------------------------------------------------------------------
# weight-normalised linear layers         models/fields.py:75-76
# --------------------------------------------------------------------------


def wn_weight(p: Params, l: int) -> Tensor:
    """Effective weight of old-style nn.utils.weight_norm(dim=0): W = v * (g / ||v||_row).
    torch._weight_norm is the primitive nn.utils.weight_norm itself calls; using it (rather than
    spelling the formula) keeps the oracle bit-identical to the reference -- a 1-ulp change of W
    already moves ~0.5 % of silhouette rays by ~5e-5 inside the convergence band."""
    return torch._weight_norm(p[f"lin{l}.weight_v"], p[f"lin{l}.weight_g"], 0)


def wn_linear(p: Params, l: int, x: Tensor) -> Tensor:
    return torch.nn.functional.linear(x, wn_weight(p, l), p[f"lin{l}.bias"])


def n_lin(p: Params) -> int:
    n = 0
    while f"lin{n}.weight_v" in p:
        n += 1
    return n


# --------------------------------------------------------------------------
# parameter construction (mirrors the RNG consumption of the reference ctors)
# --------------------------------------------------------------------------


def _fresh_linear(n_in: int, n_out: int) -> Tuple[Tensor, Tensor]:
    lin = torch.nn.Linear(n_in, n_out)
    return lin.weight.data, lin.bias.data


def make_sdf_params(
    d_hidden: int = 256,
    n_layers: int = 8,
    d_out: int = 257,
    skip_in: Sequence[int] = (4,),
    multires: int = 6,
    bias: float = 0.5,
    d_in: int = 3,
) -> Params:
    """IDR geometric init, models/fields.py:26-78 (inside_outside=False).

    Uses torch's global RNG in the same order as the reference constructor
    (one nn.Linear per layer, then the init overrides), so the same
    torch.manual_seed gives the same weights.
    """
    e = posenc_dim(multires, d_in)
    dims = [e] + [d_hidden] * n_layers + [d_out]
    L = len(dims)
    p: Params = {}
    for l in range(L - 1):
        out_dim = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
        w, b = _fresh_linear(dims[l], out_dim)
        if l == L - 2:  # fields.py:48-55
            torch.nn.init.normal_(w, mean=np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
            torch.nn.init.constant_(b, -bias)
        elif multires > 0 and l == 0:  # :63-66
            torch.nn.init.constant_(b, 0.0)
            torch.nn.init.constant_(w[:, 3:], 0.0)
            torch.nn.init.normal_(w[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
        elif multires > 0 and l in skip_in:  # :67-70
            torch.nn.init.constant_(b, 0.0)
            torch.nn.init.normal_(w, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            torch.nn.init.constant_(w[:, -(dims[0] - 3):], 0.0)
        else:  # :71-73
            torch.nn.init.constant_(b, 0.0)
            torch.nn.init.normal_(w, 0.0, np.sqrt(2) / np.sqrt(out_dim))
        # weight_norm(dim=0): g = ||w||_row, v = w     (:75-76)
        p[f"lin{l}.bias"] = b.clone()
        p[f"lin{l}.weight_g"] = w.norm(dim=1, keepdim=True).clone()
        p[f"lin{l}.weight_v"] = w.clone()
    return p


# kwargs of the three material nets of the 'ggx' config, models/network_conf.py:72-120
MATERIAL_NETS = {
    "diffuse_albedo_network": dict(d_in=9, d_out=3, multires=0, multires_view=4, mode="idr",
                                   squeeze_out=True, output_bias=0.0, output_scale=1.0),
    "specular_albedo_network": dict(d_in=6, d_out=3, multires=6, multires_view=-1, mode="no_view_dir",
                                    squeeze_out=False, output_bias=0.4, output_scale=0.1),
    "specular_roughness_network": dict(d_in=6, d_out=1, multires=6, multires_view=-1, mode="no_view_dir",
                                       squeeze_out=False, output_bias=0.1, output_scale=0.1),
}


def material_in_dim(cfg: dict, d_feature: int = 256) -> int:
    """models/fields.py:163-175."""
    d = cfg["d_in"] + d_feature
    if cfg["multires"] > 0:
        d += posenc_dim(cfg["multires"]) - 3
    if cfg["multires_view"] > 0:
        d += posenc_dim(cfg["multires_view"]) - 3
    return d


def make_material_params(cfg: dict, d_feature: int = 256, d_hidden: int = 256, n_layers: int = 4) -> Params:
    """Default nn.Linear init + weight_norm, models/fields.py:163-195 (cfg["skip_in"]: layers that read cat(h, input))."""
    dims = [material_in_dim(cfg, d_feature)] + [d_hidden] * n_layers + [cfg["d_out"]]
    skip_in = tuple(cfg.get("skip_in", ()))
    in0 = dims[0]
    for l in range(len(dims) - 1):          # :179-181
        if l in skip_in:
            dims[l] += in0
    p: Params = {}
    for l in range(len(dims) - 1):          # :183-187
        w, b = _fresh_linear(dims[l], dims[l + 1] - in0 if (l + 1) in skip_in else dims[l + 1])
        p[f"lin{l}.bias"] = b.clone()
        p[f"lin{l}.weight_g"] = w.norm(dim=1, keepdim=True).clone()
        p[f"lin{l}.weight_v"] = w.clone()
    return p


def make_material_dict(d_feature: int = 256) -> Dict[str, Params]:
    """The three nets in the order the reference dict literal builds them
    (models/network_conf.py:50-120: color_network first -- its RNG draw is
    reproduced and discarded -- then diffuse, specular x2 (duplicate key), roughness)."""
    color_cfg = dict(d_in=9, d_out=3, multires=0, multires_view=4)
    make_material_params(color_cfg, d_feature)  # color_network: consumes RNG, unused by 'ggx' render_fn
    out = {"diffuse_albedo_network": make_material_params(MATERIAL_NETS["diffuse_albedo_network"], d_feature)}
    make_material_params(MATERIAL_NETS["specular_albedo_network"], d_feature)  # first (overwritten) duplicate
    out["specular_albedo_network"] = make_material_params(MATERIAL_NETS["specular_albedo_network"], d_feature)
    out["specular_roughness_network"] = make_material_params(MATERIAL_NETS["specular_roughness_network"], d_feature)
    return out


# --------------------------------------------------------------------------
# SDF MLP                                   models/fields.py:82-137
# --------------------------------------------------------------------------


def sdf_forward(p: Params, x: Tensor, scale: float = 1.0, multires: int = 6,
                skip_in: Sequence[int] = (4,), beta: float = 100.0) -> Tensor:
    """models/fields.py:82-98.  Returns [..., d_out]; column 0 is the SDF."""
    L = n_lin(p)
    e = posenc(x * scale, multires)
    h = e
    for l in range(L):
        if l in skip_in:
            h = torch.cat([h, e], dim=-1) / SQRT2
        h = wn_linear(p, l, h)
        if l < L - 1:
            h = torch.nn.functional.softplus(h, beta=beta)
    return torch.cat([h[..., :1] / scale, h[..., 1:]], dim=-1)


def sdf_gradient(p: Params, x: Tensor, **kw) -> Tensor:
    """models/fields.py:106-118 (graph kept)."""
    x.requires_grad_(True)
    y = sdf_forward(p, x, **kw)[..., :1]
    return torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True)[0]


def sdf_get_all(p: Params, x: Tensor, is_training: bool = True, **kw) -> Tuple[Tensor, Tensor, Tensor]:
    """models/fields.py:120-137."""
    with torch.enable_grad():
        x.requires_grad_(True)
        out = sdf_forward(p, x, **kw)
        y, feat = out[..., :1], out[..., 1:]
        g = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=is_training,
                                retain_graph=is_training)[0]
    if not is_training:
        return y.detach(), feat.detach(), g.detach()
    return y, feat, g


# ---- closed form: what the CUDA kernels compute ----


def _sp(z, beta):
    return torch.nn.functional.softplus(z, beta=beta)


def _sp1(z, beta):  # sp'(z) = sigmoid(beta z), 1 beyond the threshold (beta*z > 20)
    return torch.where(z * beta > 20.0, torch.ones_like(z), torch.sigmoid(beta * z))


def _sp2(z, beta):  # sp''(z) = beta * s * (1 - s), 0 beyond the threshold
    s = torch.sigmoid(beta * z)
    return torch.where(z * beta > 20.0, torch.zeros_like(z), beta * s * (1.0 - s))


def posenc_jt(xs: Tensor, pvec: Tensor, n_freqs: int) -> Tensor:
    """J_e(x')^T p for the encoding e(x') (Appendix A 'Je^T p')."""
    out = pvec[..., 0:3].clone()
    for k in range(n_freqs):
        f = 2.0 ** k
        ps = pvec[..., 3 + 6 * k: 6 + 6 * k]
        pc = pvec[..., 6 + 6 * k: 9 + 6 * k]
        out = out + f * (torch.cos(xs * f) * ps - torch.sin(xs * f) * pc)
    return out


def posenc_j(xs: Tensor, nbar: Tensor, n_freqs: int) -> Tensor:
    """J_e(x') nbar : 3-vector -> E-vector."""
    out = [nbar]
    for k in range(n_freqs):
        f = 2.0 ** k
        out.append(f * torch.cos(xs * f) * nbar)
        out.append(-f * torch.sin(xs * f) * nbar)
    return torch.cat(out, dim=-1)


def sdf_get_all_closed_form(p: Params, x: Tensor, scale: float = 1.0, multires: int = 6,
                            skip_in: Sequence[int] = (4,), beta: float = 100.0):
    """Forward + analytic d sdf / d x without autograd.  Returns (y, feat, n, saved)."""
    L = n_lin(p)
    W = [wn_weight(p, l) for l in range(L)]
    b = [p[f"lin{l}.bias"] for l in range(L)]
    xs = x * scale
    e = posenc(xs, multires)
    E = e.shape[-1]
    u, z = [], []
    a = e
    for l in range(L):
        ul = torch.cat([a, e], dim=-1) / SQRT2 if l in skip_in else a
        zl = ul @ W[l].t() + b[l]
        u.append(ul)
        z.append(zl)
        a = _sp(zl, beta) if l < L - 1 else zl
    y = z[-1][..., :1] / scale
    feat = z[-1][..., 1:]
    # reverse chain for q = d y / d (layer input)
    q = (W[L - 1][0:1, :] / scale).expand(x.shape[0], -1)  # q_{L-1}: grad wrt u_{L-1}
    pacc = torch.zeros(x.shape[0], E, dtype=x.dtype)
    r: List[Optional[Tensor]] = [None] * L
    qa: List[Optional[Tensor]] = [None] * L
    for l in range(L - 2, -1, -1):
        if (l + 1) in skip_in:
            nh = u[l + 1].shape[-1] - E
            qal = q[..., :nh] / SQRT2
            pacc = pacc + q[..., nh:] / SQRT2
        else:
            qal = q
        qa[l] = qal
        r[l] = _sp1(z[l], beta) * qal
        q = r[l] @ W[l]
    pacc = pacc + q
    n = scale * posenc_jt(xs, pacc, multires)
    saved = dict(W=W, u=u, z=z, r=r, qa=qa, xs=xs, E=E)
    return y, feat, n, saved


def sdf_get_all_backward_closed_form(p: Params, saved, ybar: Tensor, fbar: Tensor, nbar: Tensor,
                                     scale: float = 1.0, multires: int = 6,
                                     skip_in: Sequence[int] = (4,), beta: float = 100.0) -> Params:
    """Gradients of <ybar,y> + <fbar,feat> + <nbar,n> w.r.t. weight_g / weight_v / bias
    (Appendix A parts A, B and the weight-norm chain)."""
    L = n_lin(p)
    W, u, z, r, qa, xs, E = (saved[k] for k in ("W", "u", "z", "r", "qa", "xs", "E"))
    dW = [torch.zeros_like(w) for w in W]
    db = [torch.zeros_like(p[f"lin{l}.bias"]) for l in range(L)]
    # ---- part B: through n ----
    pbar = scale * posenc_j(xs, nbar, multires)  # E-vector
    qbar = pbar
    zbarB: List[Optional[Tensor]] = [None] * L
    for l in range(L - 1):
        rbar = qbar @ W[l].t()
        dW[l] += r[l].t() @ qbar
        zbarB[l] = _sp2(z[l], beta) * qa[l] * rbar
        qabar = _sp1(z[l], beta) * rbar
        if (l + 1) in skip_in:
            qbar = torch.cat([qabar / SQRT2, pbar / SQRT2], dim=-1)
        else:
            qbar = qabar
    dW[L - 1][0, :] += qbar.sum(0) / scale
    # ---- part A: ordinary backprop ----
    delta = torch.cat([ybar / scale, fbar], dim=-1)
    for l in range(L - 1, -1, -1):
        dW[l] += delta.t() @ u[l]
        db[l] += delta.sum(0)
        if l == 0:
            break
        ubar = delta @ W[l]
        if l in skip_in:
            abar = ubar[..., : u[l].shape[-1] - E] / SQRT2
        else:
            abar = ubar
        delta = _sp1(z[l - 1], beta) * abar + zbarB[l - 1]
    # ---- weight norm ----
    grads: Params = {}
    for l in range(L):
        v = p[f"lin{l}.weight_v"]
        g = p[f"lin{l}.weight_g"]
        nv = v.norm(dim=1, keepdim=True)
        vh = v / nv
        dg = (dW[l] * vh).sum(dim=1, keepdim=True)
        grads[f"lin{l}.weight_g"] = dg
        grads[f"lin{l}.weight_v"] = (g / nv) * (dW[l] - dg * vh)
        grads[f"lin{l}.bias"] = db[l]
    return grads


# --------------------------------------------------------------------------
# material MLPs                   models/fields.py:203-239, rendering_func.py:5-16
# --------------------------------------------------------------------------


def material_forward(p: Params, cfg: dict, points: Tensor, normals: Tensor,
                     view_dirs: Optional[Tensor], feats: Tensor) -> Tensor:
    """models/fields.py:203-239 (modes 'idr' and 'no_view_dir'; cfg["skip_in"]: x = cat(x, input) / sqrt 2 before layer l, :226-227)."""
    pts = posenc(points, cfg["multires"]) if cfg["multires"] > 0 else points
    if cfg["mode"] == "idr":
        vd = posenc(view_dirs, cfg["multires_view"]) if cfg["multires_view"] > 0 else view_dirs
        h = torch.cat([pts, vd, normals, feats], dim=-1)
    elif cfg["mode"] == "no_view_dir":
        h = torch.cat([pts, normals, feats], dim=-1)
    else:
        raise ValueError(cfg["mode"])
    L = n_lin(p)
    h0 = h
    for l in range(L):
        if l in tuple(cfg.get("skip_in", ())):
            h = torch.cat([h, h0], dim=-1) / np.sqrt(2)
        h = wn_linear(p, l, h)
        if l < L - 1:
            h = torch.relu(h)
    h = cfg.get("output_scale", 1.0) * (h + cfg.get("output_bias", 0.0))
    if cfg.get("squeeze_out", True):
        h = torch.sigmoid(h)  # squeeze_out_scale = 1.0
    return h


def comp2_material_cfgs() -> Dict[str, Tuple[str, dict]]:
    """Output head -> (network name, RenderingNetwork kwargs) of the 'comp2' dictionary, models/network_conf.py:318-478, as
    get_material_comp queries them (model_bed.py:115-154 / models/rendering_func.py:19-48)."""
    idr = dict(mode="idr", multires=0, multires_view=4, squeeze_out=True, output_bias=0.0, output_scale=1.0)
    nv = lambda bias: dict(mode="no_view_dir", multires=6, multires_view=0, squeeze_out=False, output_bias=bias, output_scale=1.0)
    return {"diffuse_albedo": ("diffuse_albedo_network", idr), "specular_albedo": ("specular_albedo_network", nv(0.0)),
            "specular_roughness": ("specular_roughness_network", nv(0.1)), "metallic": ("metallic_network", nv(0.1)),
            "dielectric": ("dielectric_network", nv(0.1)), "metallic_eta": ("metallic_eta_network", nv(0.1)),
            "metallic_k": ("metallic_k_network", nv(0.1)), "dielectric_eta": ("dielectric_eta_network", nv(0.1))}


def get_materials(nets: Dict[str, Params], points: Tensor, normals: Tensor, feats: Tensor,
                  is_metal: bool = False) -> Dict[str, Tensor]:
    """models/rendering_func.py:5-16."""
    kd = material_forward(nets["diffuse_albedo_network"], MATERIAL_NETS["diffuse_albedo_network"],
                          points, normals, -normals, feats).abs()
    ks = material_forward(nets["specular_albedo_network"], MATERIAL_NETS["specular_albedo_network"],
                          points, normals, None, feats).abs()
    if not is_metal:
        ks = torch.mean(ks, dim=-1, keepdim=True).expand_as(ks)
    al = material_forward(nets["specular_roughness_network"], MATERIAL_NETS["specular_roughness_network"],
                          points, normals, None, feats).abs() + 0.01
    return {"diffuse_albedo": kd, "specular_albedo": ks, "specular_roughness": al}


# --------------------------------------------------------------------------
# GGX colocated shading                     models/renderer_ggx.py:12-16, 82-146
# --------------------------------------------------------------------------

_TABLES = None
_TABLES_DEV: Dict[str, Tuple[Tensor, Tensor]] = {}


def ggx_tables(device=None) -> Tuple[Tensor, Tensor]:
    """MTS_TRANS[5000], MTS_DIFF_TRANS[50] (models/renderer_ggx.py:65-74); the repo ships the
    same numbers as iron_b200/data/ggx_tables.npz (float32).  `device`: where the caller's tensors live (the
    reference's use_cuda flag, :75-78); bench.py's cuda-eager leg runs this file on the GPU."""
    global _TABLES
    if device is not None and torch.device(device).type != "cpu":
        key = str(torch.device(device))
        if key not in _TABLES_DEV:
            _TABLES_DEV[key] = tuple(t.to(device) for t in ggx_tables())
        return _TABLES_DEV[key]
    if _TABLES is None:
        import os
        f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "iron_b200", "data",
                                 "ggx_tables.npz"))
        _TABLES = (torch.from_numpy(f["ext_rtrans"].copy()), torch.from_numpy(f["int_diff_rtrans"].copy()))
    return _TABLES


def ggx_shade(light: Tensor, distance: Tensor, normal: Tensor, viewdir: Tensor,
              kd: Tensor, ks: Tensor, alpha: Tensor) -> Dict[str, Tensor]:
    """models/renderer_ggx.py:82-146.  Table lookups are floor-indexed (no gradient)."""
    trans, diff_trans = ggx_tables(normal.device)
    L = light / (distance * distance + 1e-10)
    c = torch.sum(viewdir * normal, dim=-1, keepdim=True)
    c = torch.clamp(c, min=0.00001, max=0.99999)
    eta = 1.48958738
    inv_eta2 = 1.0 / (eta * eta)
    alpha = torch.clamp(alpha, min=0.0001)
    c2 = c * c
    root = c2 + (1.0 - c2) / (alpha * alpha + 1e-10)
    D = 1.0 / (np.pi * alpha * alpha * root * root + 1e-10)
    F = 0.03867
    # smithG1, :12-16
    sin_t = torch.sqrt(1.0 - c * c)
    tan_t = sin_t / (c + 1e-10)
    rt = alpha * tan_t
    G1 = 2.0 / (1.0 + torch.hypot(rt, torch.ones_like(rt)))
    G = G1 ** 2
    spec = L * ks * F * D * G / (4.0 * c + 1e-10)
    wc = c ** 0.25
    wa = ((alpha - 0) / (4 - 0)) ** 0.25
    tx = torch.floor(wc * 100).long()
    ty = torch.floor(wa * 50).long()
    idx = torch.clamp(ty * 100 + tx, min=0, max=4999)
    T12 = torch.clamp(trans[idx], min=0.0, max=1.0)
    idy = torch.clamp(ty, min=0, max=49)
    Fdr = torch.clamp(1.0 - diff_trans[idy], min=0.0, max=1.0)
    diff = L * (kd / (1.0 - Fdr + 1e-10) / np.pi) * c * T12 * T12 * inv_eta2
    return {"diffuse_rgb": diff, "specular_rgb": spec, "rgb": diff + spec}


def composite_shade(light: Tensor, distance: Tensor, normal: Tensor, viewdir: Tensor, params: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """CompositeRenderer.forward (use_env_light=False), models/renderer_ggx.py:781-858, with its helpers: calc_D_specular
    :767-771 (called with eta in the roughness slot, :803), calc_G_specular / smithG1 :12-16, 773-776,
    fresnel_conductor_exact :592-606, the MODULE-level fresnel_dielectric :398-416 (called by dielectric_reflection :613),
    diffuse_reflection_ggx :669-697.  rgb is accumulated in place into the diffuse tensor (:846-851): "diffuse_rgb" is rgb."""
    trans, diff_trans = ggx_tables(normal.device)
    alpha = torch.clamp(params["specular_roughness"], min=0.00001)
    deta = torch.clamp(params["dielectric_eta"], min=1.000001, max=1.999999)
    meta = torch.clamp(params["metallic_eta"], min=0.099999, max=4.999999)
    mk = torch.clamp(params["metallic_k"], min=0.099999, max=9.999999)
    ks = torch.clamp(params["specular_albedo"], min=0.00001)
    kd = torch.clamp(params["diffuse_albedo"], min=0.00001)
    eta = 1.48958738
    c = torch.clamp(torch.sum(viewdir * normal, dim=-1, keepdim=True), min=0.00001, max=0.99999)
    c2 = c * c
    root = c2 + (1.0 - c2) / (eta * eta + 1e-10)
    D = 1.0 / (np.pi * eta * eta * root * root + 1e-10)
    tan_t = torch.sqrt(1.0 - c * c) / (c + 1e-10)
    rt = alpha * tan_t
    G1 = 2.0 / (1.0 + torch.hypot(rt, torch.ones_like(rt)))
    G = G1 * G1
    # conductor
    s2 = 1 - c2
    t1 = meta * meta - mk * mk - s2
    a2pb2 = torch.sqrt(t1 * t1 + 4 * mk * mk * meta * meta)
    a = torch.sqrt(0.5 * (a2pb2 + t1))
    term1 = a2pb2 + c2
    term2 = 2 * a * c
    Rs2 = (term1 - term2) / (term1 + term2)
    term3 = a2pb2 * c2 + s2 * s2
    term4 = term2 * s2
    Fm = 0.5 * (Rs2 * (term3 - term4) / (term3 + term4) + Rs2)
    # dielectric (c > 0: scale = 1 / eta)
    scale = 1.0 / deta
    ct = torch.sqrt(1 - (1 - c ** 2) * (scale ** 2))
    Rs = (c - deta * ct) / (c + deta * ct)
    Rp = (deta * c - ct) / (deta * c + ct)
    Fd = 0.5 * (Rs * Rs + Rp * Rp)
    L = light / (distance * distance + 1e-10)
    metallic_rgb = ks * Fm * L
    dielectric_rgb = ks * Fd * D * G / (4.0 * torch.abs(c)) * L
    spec = dielectric_rgb + metallic_rgb
    a4 = torch.clamp(alpha, min=0.0001)
    wc = c ** 0.25
    wa = ((a4 - 0) / (4 - 0)) ** 0.25
    tx = torch.floor(wc * 100).long()
    ty = torch.floor(wa * 50).long()
    T12 = torch.clamp(trans[torch.clamp(ty * 100 + tx, min=0, max=4999)], min=0.0, max=1.0)
    Fdr = torch.clamp(1.0 - diff_trans[torch.clamp(ty, min=0, max=49)], min=0.0, max=1.0)
    diffuse = L * (kd / (1.0 - Fdr + 1e-10) / np.pi) * c * T12 * T12 * (1.0 / (eta * eta))
    rgb = diffuse + spec
    return {"diffuse_rgb": rgb, "specular_rgb": spec, "metallic_rgb": metallic_rgb, "dielectric_rgb": dielectric_rgb, "rgb": rgb}


# --------------------------------------------------------------------------
# camera + unit-sphere clip                 models/raytracer.py:223-237, 240-364
# --------------------------------------------------------------------------

FIXTURE_K = [811.9282694049824, 0.0, 256.0, 0.0, 0.0, 811.9282694049824, 256.0, 0.0,
             0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0]
FIXTURE_W2C = [0.998867339183008, 0.0, -0.04758191582374219, 1.5416074755814572e-17,
               -0.013163727354886733, -0.9609695958324571, -0.27634064516051604, 1.1553154537250536e-16,
               -0.045724774418075535, 0.27665400030652737, -0.959881143224937, 2.0,
               0.0, 0.0, 0.0, 1.0]


class OCamera:
    """Pinhole camera, models/raytracer.py:240-364 (rays, crop, resize only)."""

    def __init__(self, W: int, H: int, K: Tensor, W2C: Tensor):
        self.W, self.H, self.K, self.W2C = W, H, K, W2C
        self.K_inv = torch.inverse(K)
        self.C2W = torch.inverse(W2C)

    @staticmethod
    def fixture() -> "OCamera":
        """tests/data_singleview/cam_dict_norm.json["12.png"] (512x512, f=811.93, distance 2)."""
        K = torch.tensor(FIXTURE_K, dtype=torch.float64).reshape(4, 4).float()
        W2C = torch.tensor(FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float()
        return OCamera(512, 512, K, W2C)

    def pixel_uv(self) -> Tensor:  # :300-303
        u, v = np.meshgrid(np.arange(self.W), np.arange(self.H))
        return torch.from_numpy(np.stack((u, v), axis=-1).astype(np.float32)).to(self.K.device) + 0.5

    def rays(self, uv: Tensor):  # :254-286
        sh = list(uv.shape[:-1])
        uv1 = torch.cat((uv.reshape(-1, 2), torch.ones(uv.numel() // 2, 1)), dim=-1)
        d = torch.matmul(torch.matmul(uv1, self.K_inv[:3, :3].t()), self.C2W[:3, :3].t()).reshape(sh + [3])
        dn = d.norm(dim=-1)
        d = d / dn.unsqueeze(-1)
        o = self.C2W[:3, 3].unsqueeze(0).expand(uv1.shape[0], -1).reshape(sh + [3])
        return o, d, dn

    def crop(self, w: int, h: int, ul: Tuple[int, int]) -> "OCamera":  # :327-351 (explicit ul_corner)
        K = self.K.clone()
        K[0, 2] -= ul[0]
        K[1, 2] -= ul[1]
        return OCamera(w, h, K, self.W2C.clone())

    def resize(self, factor: float) -> "OCamera":  # :353-358
        h, w = int(self.H * factor), int(self.W * factor)
        K = self.K.clone()
        K[0, :3] *= w / self.W
        K[1, :3] *= h / self.H
        return OCamera(w, h, K, self.W2C.clone())


def intersect_sphere(o: Tensor, d: Tensor, r: float = 1.0):
    """models/raytracer.py:223-237."""
    d1 = -torch.sum(d * o, dim=-1) / torch.sum(d * d, dim=-1)
    pmid = o + d1.unsqueeze(-1) * d
    tmp = r * r - torch.sum(pmid * pmid, dim=-1)
    hit = tmp > 0.0
    d2 = torch.sqrt(torch.clamp(tmp, min=0.0)) / torch.norm(d, dim=-1)
    return hit, torch.clamp(d1 - d2, min=0.0), d1 + d2


# --------------------------------------------------------------------------
# tracer                                     models/raytracer.py:45-220
# --------------------------------------------------------------------------


class TraceStats:
    def __init__(self):
        self.evals_sphere = 0
        self.evals_sampler = 0
        self.evals_bisect = 0
        self.k_max = 0
        self.n_sampler_rays = 0
        self.n_root_rays = 0

    @property
    def evals(self):
        return self.evals_sphere + self.evals_sampler + self.evals_bisect


def sphere_trace(sdf: Callable[[Tensor], Tensor], o, d, t_min, t_max, work, thr=5e-5, iters=16,
                 stats: Optional[TraceStats] = None):
    """models/raytracer.py:105-140."""
    unf = work.clone()
    t = t_min.clone()
    x = o + d * t.unsqueeze(-1)
    f = sdf(x)
    if stats:
        stats.evals_sphere += x.shape[0]
    k = 0
    while True:
        unf = unf & (f.abs() > thr) & (t < t_max)
        if k == iters or unf.sum() == 0:
            break
        k += 1
        step = f[unf]
        t[unf] += step
        x[unf] += d[unf] * step.unsqueeze(-1)
        f[unf] = sdf(x[unf])
        if stats:
            stats.evals_sphere += int(unf.sum())
    conv = work & ~unf & (f.abs() <= thr) & (t < t_max)
    return conv, unf, x, f, t


def bisect(sdf, f_lo, f_hi, t_lo, t_hi, o, d, thr=5e-5, stats: Optional[TraceStats] = None):
    """models/raytracer.py:199-220: every ray of the call is halved until no ray is working."""
    work = (f_lo > 0) & (f_hi < 0)
    t_mid = (t_lo + t_hi) / 2.0
    k = 0
    while work.any():
        f_mid = sdf(o + d * t_mid.unsqueeze(-1))
        if stats:
            stats.evals_bisect += t_mid.shape[0]
        pos = f_mid > 0
        neg = f_mid <= 0
        t_lo = torch.where(pos, t_mid, t_lo)
        t_hi = torch.where(neg, t_mid, t_hi)
        t_mid = (t_lo + t_hi) / 2.0
        work = work & ((t_hi - t_lo) > 2 * thr)
        k += 1
    x = o + d * t_mid.unsqueeze(-1)
    f = sdf(x)
    if stats:
        stats.evals_bisect += t_mid.shape[0]
        stats.k_max = max(stats.k_max, k)
    return x, t_mid, f


def dense_sample(sdf, o, d, t_min, t_max, thr=5e-5, n_steps=128, max_num_pts=200000,
                 stats: Optional[TraceStats] = None):
    """models/raytracer.py:142-197."""
    lin = torch.linspace(0, 1, steps=n_steps).float().view(1, n_steps)
    ts = t_min.unsqueeze(-1) + lin * (t_max.unsqueeze(-1) - t_min.unsqueeze(-1))
    pts = o.unsqueeze(-2) + d.unsqueeze(-2) * ts.unsqueeze(-1)
    vals = torch.cat([sdf(c) for c in torch.split(pts.reshape(-1, 3), max_num_pts, dim=0)], dim=0)
    vals = vals.reshape(-1, n_steps)
    if stats:
        stats.evals_sampler += vals.numel()
        stats.n_sampler_rays += vals.shape[0]
    out_x = torch.zeros_like(d)
    out_f = torch.zeros_like(t_min)
    out_t = torch.zeros_like(t_min)
    score = torch.sign(vals) * torch.arange(n_steps, 0, -1).float().reshape(1, n_steps)
    mval, midx = torch.min(score, dim=-1)
    root = (mval < 0.0) & (midx >= 1)
    if root.sum() > 0:
        j = midx[root].unsqueeze(-1)
        t_lo = torch.gather(ts[root], -1, j - 1).squeeze(-1)
        f_lo = torch.gather(vals[root], -1, j - 1).squeeze(-1)
        t_hi = torch.gather(ts[root], -1, j).squeeze(-1)
        f_hi = torch.gather(vals[root], -1, j).squeeze(-1)
        if stats:
            stats.n_root_rays += int(root.sum())
        px, pt, pf = bisect(sdf, f_lo, f_hi, t_lo, t_hi, o[root], d[root], thr, stats)
        out_x[root] = px
        out_f[root] = pf
        out_t[root] = pt
    return root, out_x, out_f, out_t


@torch.no_grad()
def trace_rays(sdf, o, d, t_min, t_max, work, thr=5e-5, iters=16, n_steps=128, max_num_pts=200000,
               stats: Optional[TraceStats] = None) -> Dict[str, Tensor]:
    """RayTracer.forward, models/raytracer.py:45-86."""
    conv, unf, x, f, t = sphere_trace(sdf, o, d, t_min, t_max, work, thr, iters, stats)
    if unf.sum() > 0:
        outside = (f[unf] > 0.0).float()
        s_min = outside * t[unf] + (1.0 - outside) * t_min[unf]
        s_max = outside * t_max[unf] + (1.0 - outside) * t[unf]
        s_conv, s_x, s_f, s_t = dense_sample(sdf, o[unf], d[unf], s_min, s_max, thr, n_steps, max_num_pts, stats)
        conv[unf] = s_conv
        x[unf] = s_x
        f[unf] = s_f
        t[unf] = s_t
    return {"convergent_mask": conv, "points": x, "sdf": f, "distance": t}


@torch.no_grad()
def trace_pixels(p: Params, cam: OCamera, uv: Tensor, max_num_rays: int = 200000,
                 stats: Optional[TraceStats] = None, **sdf_kw) -> Dict[str, Tensor]:
    """raytrace_pixels, models/raytracer.py:367-409 (mask=None); one tracer call per ray chunk."""
    sh = list(uv.shape[:-1])
    o, d, dn = cam.rays(uv)
    sdf = lambda x: sdf_forward(p, x, **sdf_kw)[..., 0]
    parts: Dict[str, List[Tensor]] = {}
    for oc, dc, dnc in zip(torch.split(o.reshape(-1, 3), max_num_rays), torch.split(d.reshape(-1, 3), max_num_rays),
                           torch.split(dn.reshape(-1), max_num_rays)):
        hit, t0, t1 = intersect_sphere(oc, dc, 1.0)
        res = trace_rays(sdf, oc, dc, t0, t1, hit, stats=stats)
        res["depth"] = res["distance"] / dnc
        for k, v in res.items():
            parts.setdefault(k, []).append(v)
    out = {}
    for k, v in parts.items():
        v = torch.cat(v, dim=0).reshape(sh + [-1])
        out[k] = v[..., 0] if v.shape[-1] == 1 else v
    out.update({"uv": uv, "ray_o": o, "ray_d": d, "ray_d_norm": dn})
    return out


def morph_closing3(depth: Tensor) -> Tensor:
    """kornia.morphology.closing(depth[None,None], ones(3,3)) restated: flat 3x3 dilation then erosion with the
    'geodesic' border (borders never win), models/raytracer.py:555-557.  kornia is not installable in the build
    container, so this row is restated from kornia's documentation and pinned against an independent implementation:
    cv2.morphologyEx(MORPH_CLOSE), bit-exact on tests/golden/morph_cv2.npz (oracle/make_golden_cv2.py)."""
    x = depth[None, None]
    dil = torch.nn.functional.max_pool2d(x, 3, stride=1, padding=1)            # implicit -inf padding
    ero = -torch.nn.functional.max_pool2d(-dil, 3, stride=1, padding=1)
    return ero[0, 0]


def sobel_magnitude(depth: Tensor) -> Tensor:
    """kornia.filters.sobel(depth[None,None]) restated (normalized=True, eps=1e-6): 3x3 Sobel derivatives / 8 on a
    replicate-padded image, magnitude sqrt(gx^2 + gy^2 + eps).  models/raytracer.py:569.  kornia is not installable in the
    build container; restated from its documentation and pinned against cv2.Sobel (scale 1/8, BORDER_REPLICATE) within
    5e-7 on tests/golden/morph_cv2.npz (oracle/make_golden_cv2.py)."""
    x = torch.nn.functional.pad(depth[None, None], (1, 1, 1, 1), mode="replicate")
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]) / 8.0
    gx = torch.nn.functional.conv2d(x, kx[None, None])
    gy = torch.nn.functional.conv2d(x, kx.t()[None, None])
    return torch.sqrt(gx * gx + gy * gy + 1e-6)[0, 0]


def fill_holes(res: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """raytrace_camera's hole filling, models/raytracer.py:552-564 (depth already masked by the hit mask)."""
    res = dict(res)
    res["depth"] = res["depth"] * res["convergent_mask"].float()
    depth = morph_closing3(res["depth"])
    new_mask = depth > 1e-2
    upd = new_mask & (~res["convergent_mask"])
    if upd.any():
        res["depth"] = torch.where(upd, depth, res["depth"])
        res["convergent_mask"] = new_mask
        res["distance"] = res["depth"] * res["ray_d_norm"]
        res["points"] = res["ray_o"] + res["ray_d"] * res["distance"].unsqueeze(-1)
    return res


# --------------------------------------------------------------------------
# shading of hit points            models/raytracer.py:17-24, 593-662; render_surface.py:117-156
# --------------------------------------------------------------------------


def reparam_points(x, grads, dirs, f):
    """IDR implicit differentiation, models/raytracer.py:17-24."""
    dot = torch.clamp((grads * dirs).sum(dim=-1, keepdim=True), min=1e-4)
    return x - dirs / dot * (f - f.detach())


def shade_hits(sdf_p: Params, nets: Dict[str, Params], light: Tensor, res: Dict[str, Tensor],
               is_training: bool = True, max_num_pts: int = 320000, **sdf_kw) -> Dict[str, Tensor]:
    """render_normal_and_color (models/raytracer.py:593-662) with the 'ggx' render_fn
    (render_surface.py:117-156) inlined.  Returns dense image-shaped buffers."""
    sh = list(res["convergent_mask"].shape)
    outs: Dict[str, List[Tensor]] = {}
    P = res["points"].reshape(-1, 3)
    D = res["ray_d"].reshape(-1, 3)
    O = res["ray_o"].reshape(-1, 3)
    Mk = res["convergent_mask"].reshape(-1)
    for pc, dc, oc, mk in zip(torch.split(P, max_num_pts), torch.split(D, max_num_pts),
                              torch.split(O, max_num_pts), torch.split(Mk, max_num_pts)):
        n = mk.shape[0]
        buf = {k: torch.zeros(n, 3) for k in ("color", "diffuse_color", "specular_color", "diffuse_albedo",
                                              "specular_albedo", "normal")}
        buf["specular_roughness"] = torch.zeros(n)
        if mk.any():
            x, dd, oo = pc[mk], dc[mk], oc[mk]
            f, feat, g = sdf_get_all(sdf_p, x, is_training=is_training, **sdf_kw)
            if is_training:
                x = reparam_points(x, g.detach(), -dd.detach(), f)
            with torch.set_grad_enabled(is_training):
                nrm = g / (g.norm(dim=-1, keepdim=True) + 1e-10)
                mats = get_materials(nets, x, nrm, feat)
                sh_out = ggx_shade(light, (x - oo).norm(dim=-1, keepdim=True), nrm, -dd,
                                   mats["diffuse_albedo"], mats["specular_albedo"], mats["specular_roughness"])
                buf["color"][mk] = sh_out["rgb"]
                buf["diffuse_color"][mk] = sh_out["diffuse_rgb"]
                buf["specular_color"][mk] = sh_out["specular_rgb"]
                buf["diffuse_albedo"][mk] = mats["diffuse_albedo"]
                buf["specular_albedo"][mk] = mats["specular_albedo"]
                buf["specular_roughness"][mk] = mats["specular_roughness"].squeeze(-1)
                buf["normal"][mk] = nrm
        for k, v in buf.items():
            outs.setdefault(k, []).append(v)
    out = {}
    for k, v in outs.items():
        v = torch.cat(v, dim=0)
        out[k] = v.reshape(sh + [3]) if v.dim() == 2 else v.reshape(sh)
    return out


# --------------------------------------------------------------------------
# edge sampling (row f-1)         models/raytracer.py:412-539, 566-585, 665-775
# --------------------------------------------------------------------------


def project(cam: OCamera, pts: Tensor) -> Tensor:
    """Camera.project, models/raytracer.py:305-325."""
    p = torch.cat([pts.reshape(-1, 3), torch.ones(pts.numel() // 3, 1)], dim=1)
    uv = torch.matmul(torch.matmul(p, cam.W2C.t()), cam.K.t())
    return (uv[:, :2] / uv[:, 2:3]).reshape(list(pts.shape[:-1]) + [2])


def first_unique(x: Tensor):
    """`unique` of models/raytracer.py:412-419: sorted unique values and the index of each value's FIRST occurrence."""
    uniq, inv = torch.unique(x, return_inverse=True)
    first = torch.full((uniq.numel(),), x.numel(), dtype=torch.long).scatter_reduce(0, inv, torch.arange(x.numel()), "amin")
    return uniq, first


@torch.no_grad()
def locate_edge_points(sdf_p: Params, cam: OCamera, start: Tensor, mask: Tensor, max_step=16, step_size=1e-3,
                       dot_threshold=5e-2, **sdf_kw) -> Dict[str, Tensor]:
    """models/raytracer.py:422-539: walk along the surface towards the silhouette (|n.v| <= dot_threshold)."""
    H, W = cam.H, cam.W
    finish = start.clone()
    found = mask.clone()
    if mask.any():
        cur = start[mask].clone().reshape(-1, 3)
        fnd = torch.zeros(cur.shape[0], dtype=torch.bool)
        nf = ~fnd
        o = cam.C2W[:3, 3]
        i = 0
        while True:
            v = o[None] - cur[nf]
            v = v / (v.norm(dim=-1, keepdim=True) + 1e-10)
            f, _, n = sdf_get_all(sdf_p, cur[nf].reshape(-1, 3), is_training=False, **sdf_kw)
            n = n / (n.norm(dim=-1, keepdim=True) + 1e-10)
            dot = (n * v).sum(dim=-1)
            still = dot.abs() > dot_threshold
            fnd[nf] = ~still
            nf_new = ~fnd
            if i >= max_step or nf_new.sum() == 0:
                nf = nf_new
                break
            walk = n - v / dot.unsqueeze(-1)
            walk = walk / (walk.norm(dim=-1, keepdim=True) + 1e-10)
            walk = walk - f * n
            cur[nf_new] += (step_size * walk)[still]
            nf = nf_new
            i += 1
        finish[mask] = cur
        found[mask] = fnd
    pts = finish[found]
    edge_mask = torch.zeros(H, W, dtype=torch.bool)
    edge_uv = torch.zeros(pts.shape[0], 2)
    pix = torch.zeros(0, dtype=torch.long)
    if found.any():
        edge_uv = project(cam, pts)
        pix = torch.floor(edge_uv).long()
        pix = pix[:, 1] * W + pix[:, 0]
        ok = (pix < H * W) & (pix >= 0)
        pix, pts, edge_uv = pix[ok], pts[ok], edge_uv[ok]
        if ok.any():
            pix, first = first_unique(pix)
            pts, edge_uv = pts[first], edge_uv[first]
            edge_mask.view(-1)[pix] = True
    return {"edge_mask": edge_mask, "edge_points": pts, "edge_uv": edge_uv, "edge_pixel_idx": pix}


@torch.no_grad()
def trace_camera(sdf_p: Params, cam: OCamera, max_num_rays=200000, do_fill_holes=False, detect_edges=False,
                 stats: Optional[TraceStats] = None, **sdf_kw) -> Dict[str, Tensor]:
    """raytrace_camera, models/raytracer.py:542-590."""
    res = trace_pixels(sdf_p, cam, cam.pixel_uv(), max_num_rays=max_num_rays, stats=stats, **sdf_kw)
    res["depth"] = res["depth"] * res["convergent_mask"].float()
    if do_fill_holes:
        res = fill_holes(res)
    if detect_edges:
        gnorm = sobel_magnitude(res["depth"])
        emask = (gnorm > 1e-2) & res["convergent_mask"]
        res.update(locate_edge_points(sdf_p, cam, res["points"], emask, max_step=16, step_size=1e-3, dot_threshold=5e-2, **sdf_kw))
        res["convergent_mask"] = res["convergent_mask"] & ~res["edge_mask"]
    return res


def render_edge_pixels(sdf_p: Params, nets: Dict[str, Params], light: Tensor, cam: OCamera, res: Dict[str, Tensor],
                       is_training: bool, **sdf_kw) -> None:
    """models/raytracer.py:665-727: blend the colours of two sub-pixel rays on either side of the located edge."""
    pts, uv, pix = res["edge_points"], res["edge_uv"], res["edge_pixel_idx"]
    centre = torch.floor(uv) + 0.5
    f, _, g = sdf_get_all(sdf_p, pts, is_training=is_training, **sdf_kw)
    n = g.detach() / (g.detach().norm(dim=-1, keepdim=True) + 1e-10)
    if is_training:
        pts = reparam_points(pts, g.detach(), n, f)
        uv = project(cam, pts)
    n2 = torch.matmul(n, cam.W2C[:3, :3].t())[:, :2]
    n2 = n2 / (n2.norm(dim=-1, keepdim=True) + 1e-10)
    radius = 0.707
    pos_uv = centre - radius * n2
    neg_uv = centre + radius * n2
    dot2 = torch.sum((uv - centre) * n2, dim=-1)
    alpha = 2 * torch.arccos(torch.clamp(dot2 / radius, min=0.0, max=1.0))
    w_pos = 1.0 - (alpha - torch.sin(alpha)) / (2.0 * np.pi)
    sides = []
    for suv in (pos_uv, neg_uv):
        r = trace_pixels(sdf_p, cam, suv, **sdf_kw)
        r.update(shade_hits(sdf_p, nets, light, r, is_training=is_training, **sdf_kw))
        sides.append(r)
    edge_color = sides[0]["color"] * w_pos.unsqueeze(-1) + sides[1]["color"] * (1.0 - w_pos.unsqueeze(-1))
    color = res["color"].reshape(-1, 3).clone()
    color[pix] = edge_color
    res["color"] = color.reshape(res["color"].shape)
    normal = res["normal"].reshape(-1, 3).clone()
    normal[pix] = g
    res["normal"] = normal.reshape(res["normal"].shape)
    res["edge_pos_neg_normal"] = torch.cat([sides[0]["normal"][sides[0]["convergent_mask"]],
                                            sides[1]["normal"][sides[1]["convergent_mask"]]], dim=0)
    res["uv"] = res["uv"].clone()
    res["uv"].view(-1, 2)[pix] = uv.detach()
    res["points"] = res["points"].clone()
    res["points"].view(-1, 3)[pix] = pts.detach()


def render_camera(sdf_p: Params, nets: Dict[str, Params], light: Tensor, cam: OCamera, do_fill_holes=False,
                  handle_edges=True, is_training=False, stats: Optional[TraceStats] = None, **sdf_kw) -> Dict[str, Tensor]:
    """models/raytracer.py:778-814."""
    res = trace_camera(sdf_p, cam, max_num_rays=50000, do_fill_holes=do_fill_holes, detect_edges=handle_edges, stats=stats, **sdf_kw)
    res.update(shade_hits(sdf_p, nets, light, res, is_training=is_training, **sdf_kw))
    if handle_edges and res["edge_mask"].sum() > 0:
        render_edge_pixels(sdf_p, nets, light, cam, res, is_training, **sdf_kw)
    return res


# --------------------------------------------------------------------------
# one stage-2 step (the unit bench.py counts)       render_surface.py:533-653, BASELINE.md section 3
# --------------------------------------------------------------------------


def stage2_step(sdf_p: Params, nets: Dict[str, Params], light: Tensor, cam: OCamera, target: Tensor,
                eik_points: Tensor, eik_weight: float = 0.1, stats: Optional[TraceStats] = None,
                max_num_rays: int = 50000, do_fill_holes: bool = False, handle_edges: bool = False,
                image_loss: str = "l2", ssim_weight: float = 1.0, roughrange_weight: float = 0.1, **sdf_kw):
    """render_camera(is_training) -> L2 image loss + eik_weight * eikonal(random pts + hit normals [+ edge normals])
    -> backward.  image_loss="reference" is the loss the reference trains with (row f-3: PyramidL2 + SSIM + roughness range,
    render_surface.py:594-613); "l2" is the plain L2 of round 1's goldens.  With do_fill_holes / handle_edges the
    step is the drivers' default configuration (render_surface.py:541-549, 566-567, 601-607).

    The `normal` buffer holds the *normalised* normal (render_surface.py:146), so the hit-normal eikonal
    term is ~0 exactly as in the reference (render_surface.py:601-603).
    Gradients are left in `.grad` of every tensor in sdf_p / nets / light that requires grad.
    """
    res = render_camera(sdf_p, nets, light, cam, do_fill_holes=do_fill_holes, handle_edges=handle_edges, is_training=True,
                        stats=stats, **sdf_kw)
    mask = res["convergent_mask"]
    if handle_edges:
        mask = mask | res["edge_mask"]
    eg = sdf_gradient(sdf_p, eik_points, **sdf_kw).view(-1, 3)
    eik_cnt = eg.shape[0]
    eik = ((eg.norm(dim=-1) - 1) ** 2).sum()
    img = torch.zeros(())
    if mask.any():
        if image_loss == "reference":            # render_surface.py:594-599, 609-613 (ssim_weight 1.0, roughrange_weight 0.1)
            pred_img = res["color"].permute(2, 0, 1).unsqueeze(0)
            gt_img = target.permute(2, 0, 1).unsqueeze(0)
            img = pyramid_l2_loss(pred_img, gt_img) + ssim_weight * ssim_loss(pred_img, gt_img, mask.unsqueeze(0).unsqueeze(0))
            img = img + roughrange_loss(res["specular_roughness"], mask, roughrange_weight)
        else:
            img = ((res["color"] - target) ** 2).sum() / float(mask.numel())
        hn = res["normal"][mask]
        eik_cnt += hn.shape[0]
        eik = eik + ((hn.norm(dim=-1) - 1) ** 2).sum()
        if "edge_pos_neg_normal" in res:
            en = res["edge_pos_neg_normal"]
            eik_cnt += en.shape[0]
            eik = eik + ((en.norm(dim=-1) - 1) ** 2).sum()
    loss = img + eik / eik_cnt * eik_weight
    loss.backward()
    return loss.detach(), res


# --------------------------------------------------------------------------
# patch losses (row f-3)                   models/image_losses.py:13-158, render_surface.py:594-613
# --------------------------------------------------------------------------

_GAUSS7 = None


def gauss7() -> Tensor:
    """The 7x7 kernel of PyramidL2Loss (models/image_losses.py:17-20): scipy.ndimage.gaussian_filter(dirac 7x7, sigma=1)
    = the outer product of the 1-D sigma-1 Gaussian truncated at 4 sigma (radius 4) and normalised, applied to a 7-wide
    dirac with scipy's default 'reflect' border.  Restated without scipy: w[d], d = -4..4; the 7-sample response at offset
    i from the centre collects w[i] plus the taps reflected at the array ends (half-sample symmetric)."""
    global _GAUSS7
    if _GAUSS7 is None:
        r = 4
        x = np.arange(-r, r + 1, dtype=np.float64)
        w = np.exp(-0.5 * x * x)
        w /= w.sum()
        out = np.zeros(7, dtype=np.float64)
        for j in range(7):                       # output sample j, dirac at index 3
            for d in range(-r, r + 1):
                i = j + d                        # input index read by tap d, reflected into [0, 6] (d c b a | a b c d)
                while i < 0 or i > 6:
                    i = -i - 1 if i < 0 else 2 * 7 - 1 - i
                if i == 3:
                    out[j] += w[d + r]
        g1 = out.astype(np.float64)
        _GAUSS7 = torch.from_numpy(np.outer(g1, g1).astype(np.float32))
    return _GAUSS7


def pyramid_l2_loss(pred: Tensor, trgt: Tensor) -> Tensor:
    """PyramidL2Loss.forward, models/image_losses.py:29-48.  pred, trgt: [B, 3, H, W]."""
    C = pred.shape[1]
    f = torch.zeros(C, C, 7, 7, dtype=pred.dtype, device=pred.device)
    g = gauss7().to(pred.device, pred.dtype)
    for c in range(C):
        f[c, c] = g
    h, w = pred.shape[-2:]
    d = pred - trgt
    loss = d.pow(2).sum() / (h * w)
    for k in range(1, 5):
        d = torch.nn.functional.avg_pool2d(torch.nn.functional.conv2d(d, f, padding=3), 2)
        loss = loss + d.pow(2).sum() / ((h / 2.0 ** k) * (w / 2.0 ** k))
    return loss


def erode_mask(mask: Tensor, k: int) -> Tensor:
    """kornia.morphology.erosion(mask.float(), ones(k, k)) > 0.5 restated (models/image_losses.py:154): flat k x k erosion
    with kornia's default 'geodesic' border (pixels outside the image never lose).  kornia is not installable here; pinned
    against cv2.erode (default border) in tests/golden/morph_cv2.npz."""
    x = mask.float()
    return (-torch.nn.functional.max_pool2d(-x, k, stride=1, padding=k // 2)) > 0.5


def ssim_loss(X: Tensor, Y: Tensor, mask: Optional[Tensor] = None, data_range: float = 1.0, win_size: int = 11,
              win_sigma: float = 1.5, K=(0.01, 0.03)) -> Tensor:
    """ssim_loss_fn, models/image_losses.py:97-158.  X, Y: [B, C, H, W]; mask: [B, 1, H, W] bool.  Valid (unpadded)
    separable Gaussian filtering; with a mask the map is padded back with 1.0 and averaged over the ERODED mask."""
    coords = torch.arange(win_size, dtype=torch.float, device=X.device) - win_size // 2
    g = torch.exp(-(coords ** 2) / (2 * win_sigma ** 2))
    g = g / g.sum()
    C = X.shape[1]
    wy = g.reshape(1, 1, win_size, 1).repeat(C, 1, 1, 1)
    wx = g.reshape(1, 1, 1, win_size).repeat(C, 1, 1, 1)

    def filt(t):
        out = t
        if t.shape[2] >= win_size:
            out = torch.nn.functional.conv2d(out, wy, groups=C)
        if t.shape[3] >= win_size:
            out = torch.nn.functional.conv2d(out, wx, groups=C)
        return out

    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = filt(X), filt(Y)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = filt(X * X) - mu1_sq
    s2 = filt(Y * Y) - mu2_sq
    s12 = filt(X * Y) - mu12
    cs = (2 * s12 + C2) / (s1 + s2 + C2)
    m = ((2 * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs
    m = m.mean(dim=1, keepdim=True)
    if mask is not None:
        r = win_size // 2
        m = torch.nn.functional.pad(m, (r, r, r, r), mode="constant", value=1.0)
        m = m[erode_mask(mask, win_size)]
    return 1.0 - m.mean()


def roughrange_loss(roughness: Tensor, mask: Tensor, weight: float = 0.1, value: float = 0.5) -> Tensor:
    """render_surface.py:609-613: mean excess of the hit pixels' roughness over 0.5, times the weight (0 if none)."""
    r = roughness[mask]
    r = r[r > value]
    if r.numel() > 0:
        return (r - value).mean() * weight
    return torch.zeros((), dtype=roughness.dtype, device=roughness.device)


# ---------------------------------------------------------------------------------------------- optimiser
def adam_step(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """One Adam update of one tensor, restated from torch/optim/adam.py::_single_tensor_adam (amsgrad=False, maximize=False),
    the optimiser every network of the reference's stage-2 loop uses (render_surface.py:112-113, 651-653;
    models/network_conf.py:707-716; third-party: PyTorch, pinned torch==1.11 in create_env.sh, same update rule in 2.x).
    numpy float32 arrays in, (p, m, v) out; t = 1 for the first step.  Pinned against torch.optim.Adam in
    tests/test_oracle_golden.py."""
    import numpy as np
    p, g, m, v = (np.asarray(a, dtype=np.float32) for a in (p, g, m, v))
    if weight_decay != 0.0:
        g = g + np.float32(weight_decay) * p
    m = m + np.float32(1.0 - beta1) * (g - m)                      # exp_avg.lerp_(grad, 1 - beta1)
    v = np.float32(beta2) * v + np.float32(1.0 - beta2) * g * g     # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1.0 - beta1 ** t                                        # Python doubles, like the reference
    bc2 = 1.0 - beta2 ** t
    step_size = np.float32(lr / bc1)
    denom = np.sqrt(v) / np.float32(bc2 ** 0.5) + np.float32(eps)
    p = p - step_size * (m / denom)
    return p.astype(np.float32), m.astype(np.float32), v.astype(np.float32)


# --------------------------------------------------------------------------
# stage-1 NeuS volume renderer                 models/renderer.py:43-73, 128-453
# (second half).  Stateless restatement: parameter dicts in, tensors out.
# --------------------------------------------------------------------------
NEUS_COLOR_CFG = dict(d_in=9, d_out=3, multires=10, multires_view=4, mode="idr", squeeze_out=True, output_bias=0.0,
                      output_scale=1.0, skip_in=(4,))          # confs/*_iron.conf rendering_network


def sample_pdf(bins: Tensor, weights: Tensor, n_samples: int) -> Tensor:
    """models/renderer.py:43-73 with det=True (the only mode up_sample uses, :231)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0.0 + 0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples, dtype=bins.dtype)
    u = u.expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bins_b, bins_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return bins_b + (u - cdf_b) / denom * (bins_a - bins_b)


def nerf_forward(p: Params, pts: Tensor, views: Tensor, multires: int = 10, multires_view: int = 4, skips=(4,)):
    """models/fields.py:241-322 (use_viewdirs=True): the background model; keys = the module's state dict."""
    x = posenc(pts, multires) if multires > 0 else pts
    v = posenc(views, multires_view) if multires_view > 0 else views
    h = x
    i = 0
    while f"pts_linears.{i}.weight" in p:
        h = torch.relu(h @ p[f"pts_linears.{i}.weight"].t() + p[f"pts_linears.{i}.bias"])
        if i in skips:
            h = torch.cat([x, h], -1)
        i += 1
    alpha = h @ p["alpha_linear.weight"].t() + p["alpha_linear.bias"]
    feat = h @ p["feature_linear.weight"].t() + p["feature_linear.bias"]
    h = torch.relu(torch.cat([feat, v], -1) @ p["views_linears.0.weight"].t() + p["views_linears.0.bias"])
    return alpha, h @ p["rgb_linear.weight"].t() + p["rgb_linear.bias"]


def _cumprod_weights(alpha: Tensor) -> Tensor:
    ones = torch.ones([alpha.shape[0], 1], dtype=alpha.dtype)
    return alpha * torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-7], -1), -1)[:, :-1]


def neus_up_sample(rays_o, rays_d, z_vals, sdf, n_importance: int, inv_s: float) -> Tensor:
    """:192-232."""
    B, n = z_vals.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
    radius = torch.linalg.norm(pts, ord=2, dim=-1)
    inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    sdf = sdf.reshape(B, n)
    prev_sdf, next_sdf = sdf[:, :-1], sdf[:, 1:]
    prev_z, next_z = z_vals[:, :-1], z_vals[:, 1:]
    mid_sdf = (prev_sdf + next_sdf) * 0.5
    cos_val = (next_sdf - prev_sdf) / (next_z - prev_z + 1e-5)
    prev_cos = torch.cat([torch.zeros([B, 1], dtype=z_vals.dtype), cos_val[:, :-1]], dim=-1)
    cos_val = torch.min(torch.stack([prev_cos, cos_val], dim=-1), dim=-1)[0]
    cos_val = cos_val.clip(-1e3, 0.0) * inside
    dist = next_z - prev_z
    prev_cdf = torch.sigmoid((mid_sdf - cos_val * dist * 0.5) * inv_s)
    next_cdf = torch.sigmoid((mid_sdf + cos_val * dist * 0.5) * inv_s)
    alpha = (prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)
    return sample_pdf(z_vals, _cumprod_weights(alpha), n_importance).detach()


def neus_cat_z_vals(sdf_fn, rays_o, rays_d, z_vals, new_z_vals, sdf, last: bool):
    """:234-246."""
    B, n = z_vals.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * new_z_vals[..., :, None]
    z_vals = torch.cat([z_vals, new_z_vals], dim=-1)
    z_vals, index = torch.sort(z_vals, dim=-1)
    if not last:
        new_sdf = sdf_fn(pts.reshape(-1, 3)).reshape(B, new_z_vals.shape[1])
        sdf = torch.gather(torch.cat([sdf, new_sdf], dim=-1), -1, index)
    return z_vals, sdf


def neus_render_core_outside(nerf_p: Params, rays_o, rays_d, z_vals, sample_dist: float):
    """:140-190 (n_outside > 0: 4-D inverted-sphere points)."""
    B, n = z_vals.shape
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)
    mid = z_vals + dists * 0.5
    pts = rays_o[:, None, :] + rays_d[:, None, :] * mid[..., :, None]
    dc = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).clip(1.0, 1e10)
    pts = torch.cat([pts / dc, 1.0 / dc], dim=-1).reshape(-1, 4)
    dirs = rays_d[:, None, :].expand(B, n, 3).reshape(-1, 3)
    density, color = nerf_forward(nerf_p, pts, dirs)
    alpha = 1.0 - torch.exp(-torch.nn.functional.softplus(density.reshape(B, n)) * dists)
    return alpha, color.reshape(B, n, 3)


def neus_render_core(sdf_p: Params, color_p: Params, variance: Tensor, rays_o, rays_d, z_vals, sample_dist: float,
                     background_alpha=None, background_sampled_color=None, background_rgb=None, cos_anneal_ratio: float = 0.0,
                     color_cfg: dict = NEUS_COLOR_CFG, sdf_kw: Optional[dict] = None):
    """:248-351."""
    sdf_kw = sdf_kw or {}
    B, n = z_vals.shape
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)
    mid = z_vals + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid[..., :, None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(B, n, 3).reshape(-1, 3)
    out = sdf_forward(sdf_p, pts, **sdf_kw)
    sdf, feat = out[:, :1], out[:, 1:]
    gradients = sdf_gradient(sdf_p, pts, **sdf_kw)
    sampled_color = material_forward(color_p, color_cfg, pts, gradients, dirs, feat).reshape(B, n, 3)
    inv_s = torch.exp(variance * 10.0).reshape(1, 1).clip(1e-6, 1e6).expand(B * n, 1)
    true_cos = (dirs * gradients).sum(-1, keepdim=True)
    iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + torch.relu(-true_cos) * cos_anneal_ratio)
    est_next = sdf + iter_cos * dists.reshape(-1, 1) * 0.5
    est_prev = sdf - iter_cos * dists.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s)
    next_cdf = torch.sigmoid(est_next * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).reshape(B, n).clip(0.0, 1.0)
    pts_norm = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).reshape(B, n)
    inside = (pts_norm < 1.0).float().detach()
    relax = (pts_norm < 1.2).float().detach()
    if background_alpha is not None:
        alpha = alpha * inside + background_alpha[:, :n] * (1.0 - inside)
        alpha = torch.cat([alpha, background_alpha[:, n:]], dim=-1)
        sampled_color = sampled_color * inside[:, :, None] + background_sampled_color[:, :n] * (1.0 - inside)[:, :, None]
        sampled_color = torch.cat([sampled_color, background_sampled_color[:, n:]], dim=1)
    weights = _cumprod_weights(alpha)
    weights_sum = weights.sum(dim=-1, keepdim=True)
    color = (sampled_color * weights[:, :, None]).sum(dim=1)
    if background_rgb is not None:
        color = color + background_rgb * (1.0 - weights_sum)
    gerr = (torch.linalg.norm(gradients.reshape(B, n, 3), ord=2, dim=-1) - 1.0) ** 2
    gerr = (relax * gerr).sum() / (relax.sum() + 1e-5)
    return {"color": color, "sdf": sdf, "dists": dists, "gradients": gradients.reshape(B, n, 3), "s_val": 1.0 / inv_s,
            "mid_z_vals": mid, "weights": weights, "cdf": prev_cdf.reshape(B, n), "gradient_error": gerr, "inside_sphere": inside}


def neus_render(sdf_p: Params, color_p: Params, variance: Tensor, nerf_p: Optional[Params], rays_o, rays_d, near, far,
                n_samples: int = 64, n_importance: int = 64, n_outside: int = 0, up_sample_steps: int = 4,
                t_rand: Optional[Tensor] = None, t_rand_outside: Optional[Tensor] = None, background_rgb=None,
                cos_anneal_ratio: float = 0.0, color_cfg: dict = NEUS_COLOR_CFG, sdf_kw: Optional[dict] = None):
    """:353-453.  t_rand [B,1] / t_rand_outside [B,n_outside]: the uniform numbers the reference draws at :378 / :384 when
    perturb > 0 (None = perturb off)."""
    sdf_kw = sdf_kw or {}
    B = len(rays_o)
    sample_dist = 2.0 / n_samples
    z_vals = near + (far - near) * torch.linspace(0.0, 1.0, n_samples)[None, :]
    z_out = None
    if n_outside > 0:
        z_out = torch.linspace(1e-3, 1.0 - 1.0 / (n_outside + 1.0), n_outside)
    if t_rand is not None:
        z_vals = z_vals + (t_rand - 0.5) * 2.0 / n_samples
        if n_outside > 0:
            mids = 0.5 * (z_out[..., 1:] + z_out[..., :-1])
            upper = torch.cat([mids, z_out[..., -1:]], -1)
            lower = torch.cat([z_out[..., :1], mids], -1)
            z_out = lower[None, :] + (upper - lower)[None, :] * t_rand_outside
    if n_outside > 0:
        z_out = far / torch.flip(z_out, dims=[-1]) + 1.0 / n_samples
    sdf_fn = lambda x: sdf_forward(sdf_p, x, **sdf_kw)[:, :1]
    n = n_samples
    if n_importance > 0:
        with torch.no_grad():
            pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
            sdf = sdf_fn(pts.reshape(-1, 3)).reshape(B, n_samples)
            for i in range(up_sample_steps):
                new_z = neus_up_sample(rays_o, rays_d, z_vals, sdf, n_importance // up_sample_steps, 64 * 2 ** i)
                z_vals, sdf = neus_cat_z_vals(sdf_fn, rays_o, rays_d, z_vals, new_z, sdf, last=(i + 1 == up_sample_steps))
        n = n_samples + n_importance
    bg_alpha = bg_color = None
    if n_outside > 0:
        z_feed, _ = torch.sort(torch.cat([z_vals, z_out], dim=-1), dim=-1)
        bg_alpha, bg_color = neus_render_core_outside(nerf_p, rays_o, rays_d, z_feed, sample_dist)
    ret = neus_render_core(sdf_p, color_p, variance, rays_o, rays_d, z_vals, sample_dist, bg_alpha, bg_color, background_rgb,
                           cos_anneal_ratio, color_cfg, sdf_kw)
    w = ret["weights"]
    return {"color_fine": ret["color"], "s_val": ret["s_val"].reshape(B, n).mean(dim=-1, keepdim=True), "cdf_fine": ret["cdf"],
            "weight_sum": w.sum(dim=-1, keepdim=True), "weight_max": torch.max(w, dim=-1, keepdim=True)[0],
            "gradients": ret["gradients"], "weights": w, "gradient_error": ret["gradient_error"],
            "inside_sphere": ret["inside_sphere"], "z_vals": z_vals, "z_vals_outside": z_out}
