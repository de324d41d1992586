"""Stage-1 NeuS volume renderer: iron_b200.NeuSRenderer -- SDFNetwork.get_all, RenderingNetwork with its skip
layer, the fused compositing kernels (csrc/neus.cu), the background NeRF -- against vectors of the REAL reference renderer
(oracle/make_golden_neus.py -> tests/golden/neus.npz) and, at a larger size, against the oracle restatement."""
import numpy as np
import pytest
import torch

from oracle import iron_oracle as O
from util import T, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def stage1_loss(out, target, mask):
    """render_volume.py:262-283 with the weights make_golden_neus.py used."""
    color_loss = (out["color_fine"] - target).abs().sum() / target.shape[0]
    mask_loss = torch.nn.functional.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    return color_loss + 0.1 * out["gradient_error"] + 0.1 * mask_loss


def build(H=64, d_out=65):
    import iron_b200 as ib
    sdf = ib.SDFNetwork(d_in=3, d_out=d_out, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True)
    color = ib.RenderingNetwork(d_feature=d_out - 1, mode="idr", d_in=9, d_out=3, d_hidden=H, n_layers=8, skip_in=[4],
                                weight_norm=True, multires=10, multires_view=4, squeeze_out=True)
    dev = ib.SingleVarianceNetwork(0.3)
    nerf = ib.NeRF(D=8, W=H, d_in=4, d_in_view=3, multires=10, multires_view=4, output_ch=4, skips=[4], use_viewdirs=True)
    return sdf, color, dev, nerf


@pytest.mark.parametrize("case", ["a", "b"])
def test_neus_render_golden(golden, case, gemm_mode):
    import iron_b200 as ib
    g = golden("neus")
    pre = case + "."
    mods = dict(zip(("sdf", "color", "dev", "nerf"), build()))
    for m, mod in mods.items():
        sd = {k[len(pre + "w." + m + "."):]: T(v) for k, v in g.items() if k.startswith(pre + "w." + m + ".")}
        mod.load_state_dict(sd)
        mod.to(DEV)
    n_outside, perturb, has_bg, anneal = g[pre + "cfg"]
    ren = ib.NeuSRenderer(mods["nerf"], mods["sdf"], mods["dev"], mods["color"], n_samples=16, n_importance=16,
                          n_outside=int(n_outside), up_sample_steps=4, perturb=float(perturb))
    draws = [T(g[pre + "t_rand"]).to(DEV), T(g[pre + "t_rand_out"]).to(DEV)]
    ren.rand_fn = lambda shape: draws.pop(0).reshape(shape)
    dv = lambda k: T(g[pre + k]).to(DEV)
    out = ren.render(dv("o"), dv("d"), dv("near"), dv("far"), background_rgb=torch.ones(1, 3, device=DEV) if has_bg else None,
                     cos_anneal_ratio=float(anneal))
    # fp32-grade GEMMs: 1e-4 on the rendered quantities (BASELINE: RGB 1e-3)
    for k, tol in (("color_fine", 2e-4), ("weights", 2e-4), ("weight_sum", 2e-4), ("weight_max", 2e-4), ("cdf_fine", 2e-4),
                   ("gradients", 5e-4), ("s_val", 1e-6)):
        e = np.abs(out[k].detach().cpu().numpy() - g[pre + k])
        assert e.max() <= tol * max(1.0, np.abs(g[pre + k]).max()), (k, e.max())
    assert (out["inside_sphere"].cpu().numpy() == g[pre + "inside_sphere"]).mean() >= 0.999
    assert abs(float(out["gradient_error"]) - float(g[pre + "gradient_error"])) <= 1e-4 * float(g[pre + "gradient_error"])
    loss = stage1_loss(out, dv("target"), dv("mask"))
    assert abs(float(loss.detach()) - float(g[pre + "loss"])) <= 2e-4 * abs(float(g[pre + "loss"]))
    loss.backward()
    worst, n = 0.0, 0
    for m, mod in mods.items():
        # Per-tensor relative L2 plus an absolute term at 2e-6 of the network's largest gradient (the first colour layers'
        # gradients are 1e-4 of the last layer's here and are sums of cancelling terms: their error is set by the scale of the
        # terms, not by their own norm).  Exact-fp32 FFMA products prove the algorithm against the REAL reference at 2e-4
        # (measured 3e-5); the tensor-core gradient products (3xTF32, truncating TMEM accumulation) are held to 5e-3 on this
        # 64-wide toy network with its 768-row sums (measured 2.5e-3 on one tensor, <= 5e-4 on the rest), the all-3xTF32
        # diagnostic mode to 2e-2; the production width is checked at 2e-3 below (measured 7e-6).
        rel = {"ffma": 2e-4, "tcgen05": 5e-3, "tcgen05-tf32": 2e-2}[gemm_mode]
        scale = max([float(np.linalg.norm(v)) for k, v in g.items() if k.startswith(f"{pre}g.{m}.")] + [0.0])
        for k, p in mod.named_parameters():
            key = f"{pre}g.{m}.{k}"
            if key not in g:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, key
                continue
            assert p.grad is not None, key
            ref = g[key]
            err = float(np.linalg.norm(p.grad.cpu().numpy().astype(np.float64) - ref))
            assert err <= rel * float(np.linalg.norm(ref)) + 2e-6 * scale, (key, err, float(np.linalg.norm(ref)), scale)
            if np.linalg.norm(ref) > 1e-3 * scale:
                worst = max(worst, rel_l2(p.grad.cpu().numpy(), ref))
            n += 1
    print(f"[neus {case} / {gemm_mode}] loss {float(loss.detach()):.6f} / {float(g[pre + 'loss']):.6f}, {n} gradient tensors, worst rel-L2 {worst:.2e}")


def test_neus_render_h256_vs_oracle():
    """The stage-1 configuration of confs/*_iron.conf (hidden width 256, 64 + 64 samples, 32 outside samples, 4 up-sample
    steps) on 64 rays against the oracle restatement with the same weights and the same uniform numbers: (1) the whole
    render(); (2) the hierarchical sampler's section positions; (3) render_core on IDENTICAL sections (the oracle's), where the
    tolerances can be tight because no sample position moves."""
    import iron_b200 as ib
    torch.manual_seed(3)
    sdf, color, dev, nerf = build(H=256, d_out=257)
    p = lambda mod: {k: v.detach().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    sdf_p, color_p, nerf_p = p(sdf), p(color), p(nerf)
    var = dev.variance.detach().clone().requires_grad_(True)
    for mod in (sdf, color, dev, nerf):
        mod.to(DEV)
    B = 64
    gen = torch.Generator().manual_seed(9)
    o = torch.randn(B, 3, generator=gen)
    o = o / o.norm(dim=-1, keepdim=True) * 2.0
    d = (torch.rand(B, 3, generator=gen) - 0.5) * 1.2 - o
    d = d / d.norm(dim=-1, keepdim=True)
    mid = -(o * d).sum(-1, keepdim=True)
    near, far = mid - 1.0, mid + 1.0
    target, mask = torch.rand(B, 3, generator=gen), (torch.rand(B, 1, generator=gen) > 0.4).float()
    t_rand, t_out = torch.rand(B, 1, generator=gen), torch.rand(B, 32, generator=gen)
    bg = torch.ones(1, 3)
    ren = ib.NeuSRenderer(nerf, sdf, dev, color, n_samples=64, n_importance=64, n_outside=32, up_sample_steps=4, perturb=1.0)
    ref = O.neus_render(sdf_p, color_p, var, nerf_p, o, d, near, far, n_samples=64, n_importance=64, n_outside=32,
                        up_sample_steps=4, t_rand=t_rand, t_rand_outside=t_out, background_rgb=bg, cos_anneal_ratio=0.5)
    # (1) whole render
    draws = [t_rand.to(DEV), t_out.to(DEV)]
    ren.rand_fn = lambda shape: draws.pop(0).reshape(shape)
    with torch.no_grad():
        out = ren.render(o.to(DEV), d.to(DEV), near.to(DEV), far.to(DEV), background_rgb=bg.to(DEV), cos_anneal_ratio=0.5)
    assert out["weights"].shape == (B, 160) and out["gradients"].shape == (B, 128, 3)
    e_col = (out["color_fine"].cpu() - ref["color_fine"].detach()).abs().max().item()
    e_ws = (out["weight_sum"].cpu() - ref["weight_sum"].detach()).abs().max().item()
    assert e_col <= 1e-3 and e_ws <= 1e-3, (e_col, e_ws)              # BASELINE: RGB 1e-3
    # (2) + (3): the oracle's sections through this library's render_core
    z = ref["z_vals"].detach()
    z_out = ref["z_vals_outside"].detach()
    z_feed, _ = torch.sort(torch.cat([z, z_out], dim=-1), dim=-1)
    outside = ren.render_core_outside(o.to(DEV), d.to(DEV), z_feed.to(DEV), 2.0 / 64, nerf)
    core = ren.render_core(o.to(DEV), d.to(DEV), z.to(DEV), 2.0 / 64, sdf, dev, color, background_alpha=outside["alpha"],
                           background_sampled_color=outside["sampled_color"], background_rgb=bg.to(DEV), cos_anneal_ratio=0.5)
    res = {"color_fine": core["color"], "weight_sum": core["weights"].sum(dim=-1, keepdim=True),
           "gradient_error": core["gradient_error"]}
    for k, tol in (("color", 1e-4), ("weights", 1e-4)):
        e = (core[k].detach().cpu() - ref["color_fine" if k == "color" else k].detach()).abs().max().item()
        assert e <= tol, (k, e)
    loss = stage1_loss(res, target.to(DEV), mask.to(DEV))
    loss.backward()
    rloss = stage1_loss(ref, target, mask)
    rloss.backward()
    assert abs(float(loss.detach()) - float(rloss.detach())) <= 1e-4 * abs(float(rloss.detach()))
    worst = 0.0
    for mod, pr in ((sdf, sdf_p), (color, color_p), (nerf, nerf_p)):
        scale = max(float(v.grad.norm()) for v in pr.values())
        for k, q in mod.named_parameters():
            err = float((q.grad.cpu().double() - pr[k].grad.double()).norm())
            assert err <= 2e-3 * float(pr[k].grad.norm()) + 2e-6 * scale, (k, err, float(pr[k].grad.norm()), scale)
            if float(pr[k].grad.norm()) > 1e-3 * scale:
                worst = max(worst, err / float(pr[k].grad.norm()))
    assert abs(float(dev.variance.grad) - float(var.grad)) <= 2e-3 * abs(float(var.grad))
    print(f"neus H=256: render() colour err {e_col:.2e}, weight_sum err {e_ws:.2e}; render_core on the oracle's sections: loss "
          f"{float(loss.detach()):.6f} / {float(rloss.detach()):.6f}, worst gradient rel-L2 {worst:.2e}")


def test_graphed_neus_step_matches_eager():
    """GraphedNeusStep (the stage-1 iteration as one CUDA-graph replay) against the eager step: same loss, same gradients,
    also after new rays are copied in.  perturb = 0 so both see the same sections."""
    import iron_b200 as ib
    torch.manual_seed(5)
    sdf, color, dev, nerf = build(H=128, d_out=129)
    for mod in (sdf, color, dev, nerf):
        mod.to(DEV)
    B = 96
    ren = ib.NeuSRenderer(nerf, sdf, dev, color, n_samples=32, n_importance=32, n_outside=16, up_sample_steps=4, perturb=0.0)
    bg = torch.ones(1, 3, device=DEV)
    gs = ib.GraphedNeusStep(ren, B, stage1_loss, background_rgb=bg, cos_anneal_ratio=0.25)
    params = [(f"{i}.{k}", p) for i, m in enumerate((sdf, color, dev, nerf)) for k, p in m.named_parameters()]
    for seed in (1, 2):
        gen = torch.Generator().manual_seed(seed)
        o = torch.randn(B, 3, generator=gen)
        o = o / o.norm(dim=-1, keepdim=True) * 2.0
        d = (torch.rand(B, 3, generator=gen) - 0.5) * 1.2 - o
        d = d / d.norm(dim=-1, keepdim=True)
        mid = -(o * d).sum(-1, keepdim=True)
        batch = [o, d, mid - 1.0, mid + 1.0, torch.rand(B, 3, generator=gen), (torch.rand(B, 1, generator=gen) > 0.4).float()]
        lg = float(gs.step(*[t.pin_memory() for t in batch]))
        torch.cuda.synchronize()
        gg = {k: p.grad.clone() for k, p in params}
        for _, p in params:
            p.grad = None
        bd = [t.to(DEV) for t in batch]
        out = ren.render(bd[0], bd[1], bd[2], bd[3], background_rgb=bg, cos_anneal_ratio=0.25)
        le = stage1_loss(out, bd[4], bd[5])
        le.backward()
        assert abs(lg - float(le.detach())) <= 1e-6 * abs(float(le.detach())), (lg, float(le.detach()))
        worst = max(rel_l2(gg[k].cpu().numpy(), p.grad.cpu().numpy()) for k, p in params if float(p.grad.abs().max()) > 0)
        assert worst <= 1e-4, worst                      # atomic accumulation order of the split-K weight gradients
    gs.close()
