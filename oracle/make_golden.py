"""Generate tests/golden/*.npz by running the REAL reference modules (CPU).

Runs only in the build container (needs /root/reference).  The GPU box never
executes this: it consumes the committed .npz files.  Usage:

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

`turtle`, `kornia`, `icecream` are stubbed in sys.modules (the reference imports
them at module scope, models/raytracer.py:5,9,12, but the hot path never calls them
with fill_holes=False / handle_edges=False).
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")


def import_reference():
    for name in ("turtle", "kornia", "icecream"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "icecream":
                m.ic = lambda *a, **k: None
            if name == "turtle":
                m.update = lambda *a, **k: None
            if name == "kornia":
                # kornia cannot be installed here.  The two calls the hot path's "next" rows make are restated from
                # kornia's documentation (oracle/iron_oracle.py: morph_closing3, sobel_magnitude) -- those two functions are
                # pinned independently against OpenCV (oracle/make_golden_cv2.py); everything around them (edge walk, sub-pixel blending, hole update) is the reference's.
                sys.path.insert(0, os.path.join(HERE, ".."))
                from oracle import iron_oracle as _O
                m.morphology = types.SimpleNamespace(closing=lambda x, kernel: _O.morph_closing3(x[0, 0])[None, None])
                m.filters = types.SimpleNamespace(sobel=lambda x: _O.sobel_magnitude(x[0, 0])[None, None])
            sys.modules[name] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import models.fields as fields
    import models.raytracer as raytracer
    import models.renderer_ggx as renderer_ggx
    import models.rendering_func as rendering_func
    import models.network_conf as network_conf
    return fields, raytracer, renderer_ggx, rendering_func, network_conf


def sd_np(module, prefix=""):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def build_ggx_nets(fields):
    """RenderingNetwork kwargs of models/network_conf.py:50-120, built on CPU in dict-literal order."""
    R = fields.RenderingNetwork
    d = {}
    d["color_network"] = R(d_in=9, d_out=3, d_feature=256, d_hidden=256, n_layers=4, multires_view=4, mode="idr",
                           squeeze_out=True)
    d["diffuse_albedo_network"] = R(d_in=9, d_out=3, d_feature=256, d_hidden=256, n_layers=4, multires_view=4,
                                    mode="idr", squeeze_out=True)
    for _ in range(2):  # duplicate key in the reference's dict literal: built twice, second wins
        d["specular_albedo_network"] = R(d_in=6, d_out=3, d_feature=256, d_hidden=256, n_layers=4, multires=6,
                                         multires_view=-1, mode="no_view_dir", squeeze_out=False, output_bias=0.4,
                                         output_scale=0.1)
    d["specular_roughness_network"] = R(d_in=6, d_out=1, d_feature=256, d_hidden=256, n_layers=4, multires=6,
                                        multires_view=-1, mode="no_view_dir", squeeze_out=False, output_bias=0.1,
                                        output_scale=0.1)
    return d


def make_render_fn(renderer, get_materials):
    """The 'ggx' render_fn (render_surface.py:117-156) -- the script itself cannot be imported
    (it parses argv and trains at import time), so its glue is re-assembled around the reference's
    get_materials / GGXColocatedRenderer calls."""

    def render_fn(interior_mask, nets, ray_o, ray_d, points, normals, features):
        sh = list(interior_mask.shape)
        z3 = lambda: torch.zeros(sh + [3], dtype=torch.float32)
        out = {k: z3() for k in ("color", "diffuse_color", "specular_color", "diffuse_albedo", "specular_albedo",
                                 "normal")}
        out["specular_roughness"] = torch.zeros(sh, dtype=torch.float32)
        if interior_mask.any():
            normals = normals / (normals.norm(dim=-1, keepdim=True) + 1e-10)
            params = get_materials(network_dict=nets, points=points, normals=normals, features=features)
            res = renderer(nets["point_light_network"](), (points - ray_o).norm(dim=-1, keepdim=True), normals,
                           -ray_d, params=params)
            out["color"][interior_mask] = res["rgb"]
            out["diffuse_color"][interior_mask] = res["diffuse_rgb"]
            out["specular_color"][interior_mask] = res["specular_rgb"]
            out["diffuse_albedo"][interior_mask] = params["diffuse_albedo"]
            out["specular_albedo"][interior_mask] = params["specular_albedo"]
            out["specular_roughness"][interior_mask] = params["specular_roughness"].squeeze(-1)
            out["normal"][interior_mask] = normals
        return out

    return render_fn


def perturb(sdf_net, sigma, seed=1):
    """'shape' variant: Gaussian noise on lin1..lin7.weight_v."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for l in range(1, 8):
            v = getattr(sdf_net, f"lin{l}").weight_v
            v.add_(torch.randn(v.shape, generator=g) * sigma)


def fixture_camera(raytracer):
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle.iron_oracle import FIXTURE_K, FIXTURE_W2C
    K = torch.tensor(FIXTURE_K, dtype=torch.float64).reshape(4, 4).float()
    W2C = torch.tensor(FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float()
    return raytracer.Camera(512, 512, K, W2C)


def main():
    os.makedirs(OUT, exist_ok=True)
    fields, raytracer, renderer_ggx, rendering_func, network_conf = import_reference()
    torch.set_num_threads(8)

    # ---------------- 1. GGX forward / backward ----------------
    g = torch.Generator().manual_seed(0)
    M = 384
    c = torch.rand(M, 1, generator=g)
    # normals / viewdirs with a controlled cosine, plus edge rows (grazing, back-facing, head-on)
    v = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    t = torch.nn.functional.normalize(torch.cross(v, torch.randn(M, 3, generator=g), dim=-1), dim=-1)
    n = c * v + torch.sqrt(1 - c * c) * t
    n[0] = v[0]                      # cos = 1  -> clamp max
    n[1] = -v[1]                     # cos = -1 -> clamp min
    n[2] = t[2]                      # cos = 0
    alpha = torch.rand(M, 1, generator=g) * 0.99 + 0.01
    alpha[3] = 5e-5                  # below the alpha clamp
    alpha[4] = 4.5                   # beyond the table's alpha range
    kd = torch.rand(M, 3, generator=g)
    ks = torch.rand(M, 3, generator=g)
    dist = torch.rand(M, 1, generator=g) + 1.5
    light = torch.tensor(32.0)
    wout = torch.randn(3, M, 3, generator=g)
    leaves = [x.clone().requires_grad_(True) for x in (light, dist, n, kd, ks, alpha)]
    rend = renderer_ggx.GGXColocatedRenderer(use_cuda=False)
    res = rend(leaves[0], leaves[1], leaves[2], v, {"diffuse_albedo": leaves[3], "specular_albedo": leaves[4],
                                                    "specular_roughness": leaves[5]})
    loss = (res["diffuse_rgb"] * wout[0]).sum() + (res["specular_rgb"] * wout[1]).sum() + (res["rgb"] * wout[2]).sum()
    grads = torch.autograd.grad(loss, leaves)
    np.savez(os.path.join(OUT, "ggx.npz"), light=light.numpy(), dist=dist.numpy(), normal=n.numpy(), viewdir=v.numpy(),
             kd=kd.numpy(), ks=ks.numpy(), alpha=alpha.numpy(), wout=wout.numpy(),
             diffuse_rgb=res["diffuse_rgb"].detach().numpy(), specular_rgb=res["specular_rgb"].detach().numpy(),
             rgb=res["rgb"].detach().numpy(), g_light=grads[0].numpy(), g_dist=grads[1].numpy(),
             g_normal=grads[2].numpy(), g_kd=grads[3].numpy(), g_ks=grads[4].numpy(), g_alpha=grads[5].numpy())
    print("ggx.npz")

    # ---------------- 2. small SDF net: forward, get_all, double backward ----------------
    torch.manual_seed(3)
    small = fields.SDFNetwork(d_in=3, d_out=17, d_hidden=64, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                              geometric_init=True, weight_norm=True)
    perturb(small, 0.05, seed=5)
    with torch.no_grad():  # make biases non-trivial so db paths are exercised
        gg = torch.Generator().manual_seed(6)
        for l in range(9):
            getattr(small, f"lin{l}").bias.add_(torch.randn(getattr(small, f"lin{l}").bias.shape, generator=gg) * 0.05)
    x = (torch.rand(41, 3, generator=g) * 2 - 1) * 0.8
    y, feat, grad = small.get_all(x.clone(), is_training=True)
    up = [torch.randn(y.shape, generator=g), torch.randn(feat.shape, generator=g), torch.randn(grad.shape, generator=g)]
    loss = (y * up[0]).sum() + (feat * up[1]).sum() + (grad * up[2]).sum()
    pg = torch.autograd.grad(loss, list(small.parameters()))
    out = sd_np(small, "w.")
    for (k, _), gr in zip(small.named_parameters(), pg):
        out["g." + k] = gr.numpy()
    out.update(x=x.numpy(), y=y.detach().numpy(), feat=feat.detach().numpy(), grad=grad.detach().numpy(),
               up_y=up[0].numpy(), up_feat=up[1].numpy(), up_grad=up[2].numpy(),
               fwd=small(x).detach().numpy())
    np.savez(os.path.join(OUT, "sdf_small.npz"), **out)
    print("sdf_small.npz")

    # ---------------- 3. seeded H=256 / H=512 nets: init + forward pins ----------------
    out = {}
    for H in (256, 512):
        torch.manual_seed(0)
        net = fields.SDFNetwork(d_in=3, d_out=257, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5,
                                scale=1.0, geometric_init=True, weight_norm=True)
        xs = (torch.rand(64, 3, generator=torch.Generator().manual_seed(7)) * 2 - 1)
        with torch.no_grad():
            fw = net(xs)
        _, _, gr = net.get_all(xs.clone(), is_training=False)
        out[f"h{H}.x"] = xs.numpy()
        out[f"h{H}.fwd"] = fw.numpy()
        out[f"h{H}.grad"] = gr.numpy()
        for k, vv in net.state_dict().items():
            out[f"h{H}.sum.{k}"] = np.array([vv.double().sum().item(), vv.double().abs().sum().item()])
    np.savez(os.path.join(OUT, "sdf_seeded.npz"), **out)
    print("sdf_seeded.npz")

    # ---------------- 4. material nets ----------------
    torch.manual_seed(0)
    nets = build_ggx_nets(fields)
    Mh = 53
    pts = (torch.rand(Mh, 3, generator=g) - 0.5)
    nrm = torch.nn.functional.normalize(torch.randn(Mh, 3, generator=g), dim=-1)
    fts = torch.randn(Mh, 256, generator=g) * 0.3
    leaves = [t_.clone().requires_grad_(True) for t_ in (pts, nrm, fts)]
    mats = rendering_func.get_materials(nets, leaves[0], leaves[1], leaves[2])
    ups = {k: torch.randn(vv.shape, generator=g) for k, vv in mats.items()}
    loss = sum((mats[k] * ups[k]).sum() for k in mats)
    plist = []
    for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network"):
        plist += [(nm + "." + k, p_) for k, p_ in nets[nm].named_parameters()]
    gr = torch.autograd.grad(loss, leaves + [p_ for _, p_ in plist])
    out = dict(points=pts.numpy(), normals=nrm.numpy(), feats=fts.numpy(), g_points=gr[0].numpy(),
               g_normals=gr[1].numpy(), g_feats=gr[2].numpy())
    for k in mats:
        out["out." + k] = mats[k].detach().numpy()
        out["up." + k] = ups[k].numpy()
    for (k, _), gg_ in zip(plist, gr[3:]):
        out["gsum." + k] = np.array([gg_.double().sum().item(), gg_.double().abs().sum().item(),
                                     gg_.double().pow(2).sum().sqrt().item()])
        if gg_.numel() <= 1024 or k.endswith("lin4.weight_v"):
            out["g." + k] = gg_.numpy()
    for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network"):
        for k, vv in nets[nm].state_dict().items():
            out[f"sum.{nm}.{k}"] = np.array([vv.double().sum().item(), vv.double().abs().sum().item()])
    np.savez(os.path.join(OUT, "materials.npz"), **out)
    print("materials.npz")

    # ---------------- 5. tracer: 64x64 crop (BASELINE config 2 geometry), H=256 seed 0 + noise 0.005 ----------------
    cam512 = fixture_camera(raytracer)
    torch.manual_seed(0)
    sdf_net = fields.SDFNetwork(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5,
                                scale=1.0, geometric_init=True, weight_norm=True)
    perturb(sdf_net, 0.005, seed=1)
    rt = raytracer.RayTracer()
    out = {}
    # (a) centre crop: mostly sphere-tracing hits; (b) silhouette crop: sampler + bisection + misses
    for tag, ul in (("centre", (224, 224)), ("edge", (430, 224))):
        cam, _, _ = cam512.crop_region(64, 64, ul_corner=ul)
        res = raytracer.raytrace_pixels(sdf_net, rt, cam.get_uv(), cam, max_num_rays=200000)
        for k in ("convergent_mask", "points", "sdf", "distance", "depth", "ray_o", "ray_d", "ray_d_norm"):
            out[f"{tag}.{k}"] = res[k].numpy()
        out[f"{tag}.ul"] = np.array(ul)
        print(tag, "hits", int(res["convergent_mask"].sum()), "of", res["convergent_mask"].numel())
    # (c) a coarse whole-object view (resize 1/8 -> 64x64): rays that miss the unit sphere are absent in this
    #     camera, so add a synthetic ray bundle that does miss it
    cam8, _ = cam512.resize(0.125)
    res = raytracer.raytrace_pixels(sdf_net, rt, cam8.get_uv(), cam8, max_num_rays=1500)  # 3 tracer calls: k_max per call
    for k in ("convergent_mask", "points", "sdf", "distance", "depth"):
        out[f"coarse.{k}"] = res[k].numpy()
    print("coarse hits", int(res["convergent_mask"].sum()))
    o = torch.tensor([[0.0, 0.0, -3.0]]).expand(256, 3).contiguous()
    d = torch.nn.functional.normalize(torch.cat([torch.randn(256, 2, generator=g) * 0.25, torch.ones(256, 1)], -1), dim=-1)
    hit, t0, t1 = raytracer.intersect_sphere(o, d, r=1.0)
    res = rt(lambda q: sdf_net(q)[..., 0], o, d, t0, t1, hit)
    out.update({"bundle.ray_o": o.numpy(), "bundle.ray_d": d.numpy(), "bundle.hit": hit.numpy(), "bundle.t0": t0.numpy(),
                "bundle.t1": t1.numpy()})
    for k in ("convergent_mask", "points", "sdf", "distance"):
        out[f"bundle.{k}"] = res[k].numpy()
    print("bundle: sphere hits", int(hit.sum()), "surface hits", int(res["convergent_mask"].sum()))
    np.savez_compressed(os.path.join(OUT, "trace_h256.npz"), **out)
    print("trace_h256.npz")

    # ---------------- 6. full stage-2 step on a 32x32 crop (render_camera is_training=True + loss + backward) -------
    nets["point_light_network"] = network_conf.PointLightNetwork()
    nets["point_light_network"].set_light(32.0)
    rend = renderer_ggx.GGXColocatedRenderer(use_cuda=False)
    render_fn = make_render_fn(rend, rendering_func.get_materials)
    cam, _, _ = cam512.crop_region(32, 32, ul_corner=(448, 230))
    results = raytracer.render_camera(cam, sdf_net, rt, nets, render_fn, fill_holes=False, handle_edges=False,
                                      is_training=True)
    mask = results["convergent_mask"]
    tgt = torch.rand(32, 32, 3, generator=torch.Generator().manual_seed(11)) * 0.5
    eik_pts = torch.empty(32 * 32 // 2, 3).uniform_(-1.0, 1.0, generator=torch.Generator().manual_seed(12))
    eg = sdf_net.gradient(eik_pts.clone()).view(-1, 3)
    eik_cnt = eg.shape[0]
    eik = ((eg.norm(dim=-1) - 1) ** 2).sum()
    img = ((results["color"] - tgt) ** 2).sum() / float(mask.numel())
    hn = results["normal"][mask]
    eik_cnt += hn.shape[0]
    eik = eik + ((hn.norm(dim=-1) - 1) ** 2).sum()
    loss = img + eik / eik_cnt * 0.1
    loss.backward()
    out = dict(ul=np.array((448, 230)), target=tgt.numpy(), eik_points=eik_pts.numpy(), loss=loss.detach().numpy(),
               mask=mask.numpy())
    for k in ("color", "diffuse_color", "specular_color", "diffuse_albedo", "specular_albedo", "specular_roughness",
              "normal", "points", "distance"):
        out["res." + k] = results[k].detach().numpy()
    allp = [("sdf." + k, p_) for k, p_ in sdf_net.named_parameters()]
    for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network", "point_light_network"):
        allp += [(nm + "." + k, p_) for k, p_ in nets[nm].named_parameters()]
    for k, p_ in allp:
        gr_ = p_.grad
        out["gsum." + k] = np.array([gr_.double().sum().item(), gr_.double().abs().sum().item(),
                                     gr_.double().pow(2).sum().sqrt().item()])
        if gr_.numel() <= 1024 or k.endswith("lin8.weight_v") or k.endswith("lin0.weight_v"):
            out["g." + k] = gr_.numpy()
    np.savez_compressed(os.path.join(OUT, "step_h256.npz"), **out)
    print("step_h256.npz  hits", int(mask.sum()), "loss", float(loss))

    # ---------------- 7. stage-2 step WITH hole filling and edge sampling (the drivers' default), 32x32 silhouette crop -----
    for p_ in [p for _, p in allp]:
        p_.grad = None
    results = raytracer.render_camera(cam, sdf_net, rt, nets, render_fn, fill_holes=True, handle_edges=True, is_training=True)
    mask = results["convergent_mask"] | results["edge_mask"]
    eg = sdf_net.gradient(eik_pts.clone()).view(-1, 3)
    eik_cnt = eg.shape[0]
    eik = ((eg.norm(dim=-1) - 1) ** 2).sum()
    img = ((results["color"] - tgt) ** 2).sum() / float(mask.numel())
    hn = results["normal"][mask]
    eik_cnt += hn.shape[0]
    eik = eik + ((hn.norm(dim=-1) - 1) ** 2).sum()
    en = results["edge_pos_neg_normal"]
    eik_cnt += en.shape[0]
    eik = eik + ((en.norm(dim=-1) - 1) ** 2).sum()
    loss = img + eik / eik_cnt * 0.1
    loss.backward()
    out = dict(ul=np.array((448, 230)), target=tgt.numpy(), eik_points=eik_pts.numpy(), loss=loss.detach().numpy(),
               mask=results["convergent_mask"].numpy(), edge_mask=results["edge_mask"].numpy(),
               edge_pixel_idx=results["edge_pixel_idx"].numpy(), edge_uv=results["edge_uv"].detach().numpy(),
               edge_points=results["edge_points"].detach().numpy(), n_edge_normals=np.array(en.shape[0]))
    for k in ("color", "normal", "points", "distance", "depth", "uv"):
        out["res." + k] = results[k].detach().numpy()
    for k, p_ in allp:
        gr_ = p_.grad
        out["gsum." + k] = np.array([gr_.double().sum().item(), gr_.double().abs().sum().item(),
                                     gr_.double().pow(2).sum().sqrt().item()])
        if gr_.numel() <= 1024 or k.endswith("lin8.weight_v") or k.endswith("lin0.weight_v"):
            out["g." + k] = gr_.numpy()
    np.savez_compressed(os.path.join(OUT, "step_edges_h256.npz"), **out)
    print("step_edges_h256.npz  hits", int(results["convergent_mask"].sum()), "edge pixels", int(results["edge_mask"].sum()),
          "loss", float(loss))


if __name__ == "__main__":
    main()
