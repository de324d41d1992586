"""Golden vectors of the stage-1 NeuS volume renderer from the REAL reference (models/renderer.py NeuSRenderer with the
reference's SDFNetwork / RenderingNetwork(skip_in=[4]) / SingleVarianceNetwork / NeRF), CPU, build container only:

    python oracle/make_golden_neus.py        # rewrites tests/golden/neus.npz

Two cases on small networks (hidden width 64, so the file stays small; layer structure = confs/*_iron.conf):
  a: n_outside = 0, perturb off, no fixed background, cos_anneal_ratio 0
  b: n_outside = 8 (background NeRF), perturb on (the uniform numbers the reference drew are stored), white background,
     cos_anneal_ratio 0.3
For each: the render() outputs and the gradients of the stage-1 loss (render_volume.py:262-283: L1 colour + 0.1 eikonal +
0.1 BCE mask) w.r.t. every parameter.  `mcubes` / `icecream` are stubbed (imported at module scope, unused here)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.make_golden import import_reference, OUT  # noqa: E402


def build(fields, n_outside):
    torch.manual_seed(0)
    sdf = fields.SDFNetwork(d_in=3, d_out=65, d_hidden=64, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                            geometric_init=True, weight_norm=True)
    color = fields.RenderingNetwork(d_feature=64, mode="idr", d_in=9, d_out=3, d_hidden=64, n_layers=8, skip_in=[4],
                                    weight_norm=True, multires=10, multires_view=4, squeeze_out=True)
    dev = fields.SingleVarianceNetwork(0.3)
    nerf = fields.NeRF(D=8, W=64, d_in=4, d_in_view=3, multires=10, multires_view=4, output_ch=4, skips=[4], use_viewdirs=True)
    return sdf, color, dev, nerf


def rays(B, seed):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(B, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * 2.0                      # cameras 2.0 from the origin
    tgt = (torch.rand(B, 3, generator=g) - 0.5) * 1.2              # aim at / around the r ~ 0.5 object
    d = tgt - o
    d = d / d.norm(dim=-1, keepdim=True)
    mid = -(o * d).sum(-1, keepdim=True) / (d * d).sum(-1, keepdim=True)     # dataset.near_far_from_sphere
    return o, d, mid - 1.0, mid + 1.0, torch.rand(B, 3, generator=g), (torch.rand(B, 1, generator=g) > 0.4).float()


def loss_fn(out, target, mask):
    color_loss = torch.nn.functional.l1_loss(out["color_fine"] - target, torch.zeros_like(target), reduction="sum") / target.shape[0]
    mask_loss = torch.nn.functional.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    return color_loss + 0.1 * out["gradient_error"] + 0.1 * mask_loss


def main():
    fields, *_ = import_reference()
    for name in ("mcubes",):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    import models.renderer as R
    out = {}
    B = 24
    for case, (n_outside, perturb, bg, anneal) in {"a": (0, 0.0, None, 0.0), "b": (8, 1.0, torch.ones(1, 3), 0.3)}.items():
        sdf, color, dev, nerf = build(fields, n_outside)
        o, d, near, far, target, mask = rays(B, 5 + ord(case))
        ren = R.NeuSRenderer(nerf, sdf, dev, color, n_samples=16, n_importance=16, n_outside=n_outside, up_sample_steps=4,
                             perturb=perturb)
        torch.manual_seed(77)
        t_rand = torch.rand([B, 1])                                 # what render() draws at :378 ...
        t_rand_out = torch.rand([B, max(n_outside, 1)])             # ... and at :384 (only when n_outside > 0)
        torch.manual_seed(77)
        res = ren.render(o, d, near, far, background_rgb=bg, cos_anneal_ratio=anneal)
        loss = loss_fn(res, target, mask)
        mods = {"sdf": sdf, "color": color, "dev": dev, "nerf": nerf}
        params = [(f"{m}.{k}", p) for m, mod in mods.items() for k, p in mod.named_parameters()]
        grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
        for m, mod in mods.items():
            for k, v in mod.state_dict().items():
                out[f"{case}.w.{m}.{k}"] = v.detach().numpy().copy()
        for (k, _), g in zip(params, grads):
            if g is not None:
                out[f"{case}.g.{k}"] = g.detach().numpy().copy()
        for k in ("color_fine", "s_val", "cdf_fine", "weight_sum", "weight_max", "gradients", "weights", "gradient_error",
                  "inside_sphere"):
            out[f"{case}.{k}"] = res[k].detach().numpy().copy()
        out[f"{case}.loss"] = np.float32(loss.item())
        for k, v in dict(o=o, d=d, near=near, far=far, target=target, mask=mask, t_rand=t_rand, t_rand_out=t_rand_out).items():
            out[f"{case}.{k}"] = v.numpy().copy()
        out[f"{case}.cfg"] = np.array([n_outside, perturb, 0.0 if bg is None else 1.0, anneal], dtype=np.float64)
        print(case, "loss", float(loss), "weight_sum mean", float(res["weight_sum"].mean()), "gradient_error",
              float(res["gradient_error"]), "unused grads:", [k for (k, _), g in zip(params, grads) if g is None])
    np.savez_compressed(os.path.join(OUT, "neus.npz"), **out)
    print("wrote", os.path.join(OUT, "neus.npz"), os.path.getsize(os.path.join(OUT, "neus.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
