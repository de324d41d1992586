"""Pins oracle/iron_oracle.py against vectors produced by the real reference modules
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import iron_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a).copy())


def close(a, b, atol, rtol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    assert (err <= tol).all(), f"max err {err.max():.3e} (tol {atol:.1e}+{rtol:.1e}*|b|), worst at {np.argmax(err - tol)}"


def sdf_params_from(g, prefix):
    return {k[len(prefix):]: T(v) for k, v in g.items() if k.startswith(prefix)}


def test_ggx_forward_backward(golden):
    g = golden("ggx")
    leaves = [T(g[k]).requires_grad_(True) for k in ("light", "dist", "normal", "kd", "ks", "alpha")]
    out = O.ggx_shade(leaves[0], leaves[1], leaves[2], T(g["viewdir"]), leaves[3], leaves[4], leaves[5])
    for k in ("diffuse_rgb", "specular_rgb", "rgb"):
        close(out[k].detach().numpy(), g[k], 0.0, 1e-6)
    w = T(g["wout"])
    loss = (out["diffuse_rgb"] * w[0]).sum() + (out["specular_rgb"] * w[1]).sum() + (out["rgb"] * w[2]).sum()
    gr = torch.autograd.grad(loss, leaves)
    for got, k in zip(gr, ("g_light", "g_dist", "g_normal", "g_kd", "g_ks", "g_alpha")):
        close(got.numpy(), g[k], 1e-6 * np.abs(g[k]).max(), 1e-5)


def test_sdf_small_forward_getall(golden):
    g = golden("sdf_small")
    p = sdf_params_from(g, "w.")
    x = T(g["x"])
    close(O.sdf_forward(p, x).numpy(), g["fwd"], 1e-6, 1e-6)
    y, f, n = O.sdf_get_all(p, x.clone(), is_training=False)
    close(y.numpy(), g["y"], 1e-6)
    close(f.numpy(), g["feat"], 1e-6, 1e-6)
    close(n.numpy(), g["grad"], 1e-5, 1e-5)


def test_sdf_small_double_backward_autograd(golden):
    g = golden("sdf_small")
    p = {k: v.requires_grad_(True) for k, v in sdf_params_from(g, "w.").items()}
    y, f, n = O.sdf_get_all(p, T(g["x"]), is_training=True)
    loss = (y * T(g["up_y"])).sum() + (f * T(g["up_feat"])).sum() + (n * T(g["up_grad"])).sum()
    names = sorted(p)
    gr = torch.autograd.grad(loss, [p[k] for k in names])
    for k, got in zip(names, gr):
        ref = g["g." + k]
        close(got.numpy(), ref, 2e-5 * max(1.0, np.abs(ref).max()), 1e-4)


def test_sdf_small_closed_form_matches_reference(golden):
    """Appendix-A closed form (what the CUDA kernels implement) vs the reference's autograd, in fp64."""
    g = golden("sdf_small")
    p = {k: v.double() for k, v in sdf_params_from(g, "w.").items()}
    x = T(g["x"]).double()
    y, f, n, saved = O.sdf_get_all_closed_form(p, x)
    close(y.numpy(), g["y"], 2e-6)
    close(f.numpy(), g["feat"], 2e-6, 1e-6)
    close(n.numpy(), g["grad"], 1e-5, 1e-5)
    grads = O.sdf_get_all_backward_closed_form(p, saved, T(g["up_y"]).double(), T(g["up_feat"]).double(),
                                               T(g["up_grad"]).double())
    for k, got in grads.items():
        ref = g["g." + k]
        close(got.numpy(), ref, 3e-5 * max(1.0, np.abs(ref).max()), 2e-4)
    # and against the oracle's own autograd in fp64: must agree to rounding
    pa = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ya, fa, na = O.sdf_get_all(pa, x.clone(), is_training=True)
    loss = (ya * T(g["up_y"])).sum() + (fa * T(g["up_feat"])).sum() + (na * T(g["up_grad"])).sum()
    names = sorted(pa)
    for k, ga in zip(names, torch.autograd.grad(loss, [pa[k] for k in names])):
        close(grads[k].numpy(), ga.numpy(), 1e-9 * max(1.0, ga.abs().max().item()), 1e-9)


@pytest.mark.parametrize("H", [256, 512])
def test_seeded_init_and_forward(golden, H):
    g = golden("sdf_seeded")
    torch.manual_seed(0)
    p = O.make_sdf_params(d_hidden=H)
    for k, v in p.items():
        s = g[f"h{H}.sum.{k}"]
        assert abs(v.double().sum().item() - s[0]) <= 1e-9 * max(1, abs(s[0])), k
        assert abs(v.double().abs().sum().item() - s[1]) <= 1e-9 * max(1, abs(s[1])), k
    x = T(g[f"h{H}.x"])
    close(O.sdf_forward(p, x).numpy(), g[f"h{H}.fwd"], 2e-6, 1e-6)
    _, _, n = O.sdf_get_all(p, x.clone(), is_training=False)
    close(n.numpy(), g[f"h{H}.grad"], 5e-6, 1e-6)


def test_materials(golden):
    g = golden("materials")
    torch.manual_seed(0)
    nets = O.make_material_dict()
    for nm, p in nets.items():
        for k, v in p.items():
            s = g[f"sum.{nm}.{k}"]
            assert abs(v.double().sum().item() - s[0]) <= 1e-9 * max(1, abs(s[0])), (nm, k)
            v.requires_grad_(True)
    leaves = [T(g[k]).requires_grad_(True) for k in ("points", "normals", "feats")]
    mats = O.get_materials(nets, *leaves)
    for k in mats:
        close(mats[k].detach().numpy(), g["out." + k], 1e-6, 1e-6)
    loss = sum((mats[k] * T(g["up." + k])).sum() for k in mats)
    plist = [(f"{nm}.{k}", v) for nm, p in nets.items() for k, v in p.items()]
    gr = torch.autograd.grad(loss, leaves + [v for _, v in plist])
    for got, k in zip(gr[:3], ("g_points", "g_normals", "g_feats")):
        close(got.numpy(), g[k], 1e-6 * max(1.0, np.abs(g[k]).max()), 1e-5)
    for (k, _), got in zip(plist, gr[3:]):
        s = g["gsum." + k]
        assert abs(got.double().pow(2).sum().sqrt().item() - s[2]) <= 1e-5 * max(1e-6, s[2]), k
        if "g." + k in g:
            close(got.numpy(), g["g." + k], 1e-6 * max(1.0, np.abs(g["g." + k]).max()), 1e-5)


def _trace_net():
    torch.manual_seed(0)
    p = O.make_sdf_params(d_hidden=256)
    gen = torch.Generator().manual_seed(1)
    for l in range(1, 8):
        v = p[f"lin{l}.weight_v"]
        v.add_(torch.randn(v.shape, generator=gen) * 0.005)
    return p


def _cmp_trace(res, g, tag, depth=True):
    m_ref = g[f"{tag}.convergent_mask"].astype(bool)
    m = res["convergent_mask"].numpy().astype(bool)
    assert (m == m_ref).all(), f"{tag}: {(m != m_ref).sum()} mask mismatches of {m.size}"
    for k in ("points", "sdf", "distance") + (("depth",) if depth else ()):
        a, b = res[k].numpy()[m], g[f"{tag}.{k}"][m_ref]
        close(a, b, 2e-6)


@pytest.mark.parametrize("tag", ["centre", "edge"])
def test_trace_crops(golden, tag):
    g = golden("trace_h256")
    p = _trace_net()
    ul = tuple(int(v) for v in g[f"{tag}.ul"])
    cam = O.OCamera.fixture().crop(64, 64, ul)
    o, d, dn = cam.rays(cam.pixel_uv())
    close(d.numpy(), g[f"{tag}.ray_d"], 1e-7)
    close(dn.numpy(), g[f"{tag}.ray_d_norm"], 1e-6)
    close(o.numpy(), g[f"{tag}.ray_o"], 1e-7)
    st = O.TraceStats()
    res = O.trace_pixels(p, cam, cam.pixel_uv(), stats=st)
    _cmp_trace(res, g, tag)
    assert st.evals > 4096


def test_trace_coarse_chunked_and_bundle(golden):
    g = golden("trace_h256")
    p = _trace_net()
    cam = O.OCamera.fixture().resize(0.125)
    res = O.trace_pixels(p, cam, cam.pixel_uv(), max_num_rays=1500)
    _cmp_trace(res, g, "coarse")
    o, d = T(g["bundle.ray_o"]), T(g["bundle.ray_d"])
    hit, t0, t1 = O.intersect_sphere(o, d, 1.0)
    assert (hit.numpy() == g["bundle.hit"]).all()
    close(t0.numpy(), g["bundle.t0"], 1e-7)
    close(t1.numpy(), g["bundle.t1"], 1e-7)
    res = O.trace_rays(lambda q: O.sdf_forward(p, q)[..., 0], o, d, t0, t1, hit)
    _cmp_trace(res, g, "bundle", depth=False)


def test_full_step(golden):
    g = golden("step_h256")
    p = _trace_net()
    torch.manual_seed(0)
    nets = O.make_material_dict()
    light = torch.tensor(32.0, requires_grad=True)
    for d in [p] + list(nets.values()):
        for v in d.values():
            v.requires_grad_(True)
    cam = O.OCamera.fixture().crop(32, 32, tuple(int(v) for v in g["ul"]))
    loss, res = O.stage2_step(p, nets, light, cam, T(g["target"]), T(g["eik_points"]))
    assert (res["convergent_mask"].numpy() == g["mask"]).all()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for k in ("color", "diffuse_color", "specular_color", "diffuse_albedo", "specular_albedo", "specular_roughness",
              "normal"):
        close(res[k].detach().numpy(), g["res." + k], 2e-6, 1e-5)
    allp = [("sdf." + k, v) for k, v in p.items()]
    for nm, d in nets.items():
        allp += [(f"{nm}.{k}", v) for k, v in d.items()]
    allp.append(("point_light_network.light", light))
    for k, v in allp:
        s = g["gsum." + k]
        nrm = v.grad.double().pow(2).sum().sqrt().item()
        assert abs(nrm - s[2]) <= 1e-3 * max(s[2], 1e-12), (k, nrm, s[2])
        if "g." + k in g:
            ref = g["g." + k]
            close(v.grad.numpy(), ref, 1e-4 * np.abs(ref).max(), 1e-3)


def test_step_with_edges_golden(golden):
    """The drivers' default step (fill_holes + edge sampling): oracle vs the real reference's render_camera.  The two kornia
    calls are shared restatements (pinned against OpenCV below); the edge walk, uniqueness, sub-pixel blending and their gradients are pinned."""
    import torch
    g = golden("step_edges_h256")
    torch.manual_seed(0)
    mats = O.make_material_dict()
    torch.manual_seed(0)
    sdf = O.make_sdf_params(d_hidden=256)
    gen = torch.Generator().manual_seed(1)
    for l in range(1, 8):
        v = sdf[f"lin{l}.weight_v"]
        v.add_(torch.randn(v.shape, generator=gen) * 0.005)
    for d in [sdf] + list(mats.values()):
        for v in d.values():
            v.requires_grad_(True)
    light = torch.tensor(32.0, requires_grad=True)
    cam = O.OCamera.fixture().crop(32, 32, tuple(int(v) for v in g["ul"]))
    loss, res = O.stage2_step(sdf, mats, light, cam, T(g["target"]), T(g["eik_points"]), do_fill_holes=True, handle_edges=True)
    assert (res["convergent_mask"].numpy() == g["mask"]).all()
    assert (res["edge_mask"].numpy() == g["edge_mask"]).all()
    assert (res["edge_pixel_idx"].numpy() == g["edge_pixel_idx"]).all()
    close(res["edge_uv"].detach().numpy(), g["edge_uv"], 2e-4)
    close(res["color"].detach().numpy(), g["res.color"], 2e-5, 1e-4)
    close(res["normal"].detach().numpy(), g["res.normal"], 2e-5, 1e-4)
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert res["edge_pos_neg_normal"].shape[0] == int(g["n_edge_normals"])
    for k, v in sdf.items():
        ref = g["gsum.sdf." + k]
        got = v.grad.double().pow(2).sum().sqrt().item()
        assert abs(got - ref[2]) <= 2e-3 * ref[2] + 1e-9, (k, got, ref[2])
    assert abs(light.grad.item() - g["g.point_light_network.light"]) <= 1e-4 * abs(g["g.point_light_network.light"])


def test_oracle_adam_matches_torch_optim_adam():
    """The optimiser restatement against the reference's own optimiser (torch.optim.Adam on CPU), ragged tensors, 6 steps,
    with and without weight decay."""
    import torch
    from oracle import iron_oracle as O
    for wd in (0.0, 0.01):
        gen = torch.Generator().manual_seed(3)
        ps = [torch.randn(n, generator=gen) for n in (1, 7, 1000, 257 * 33)]
        ref = [p.clone().requires_grad_(True) for p in ps]
        opt = torch.optim.Adam(ref, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
        mine = [(p.numpy().copy(), np.zeros(p.numel(), np.float32), np.zeros(p.numel(), np.float32)) for p in ps]
        for t in range(1, 7):
            grads = [torch.randn(p.shape, generator=gen) * (10.0 ** (t - 4)) for p in ps]
            for r, g in zip(ref, grads):
                r.grad = g.clone()
            opt.step()
            mine = [O.adam_step(p, g.numpy(), m, v, t, 1e-2, 0.9, 0.999, 1e-8, wd) for (p, m, v), g in zip(mine, grads)]
            for r, (p, m, v) in zip(ref, mine):
                assert np.allclose(p, r.detach().numpy(), rtol=2e-6, atol=1e-7), (wd, t, np.abs(p - r.detach().numpy()).max())


def test_closing_and_sobel_restatements_match_opencv(golden):
    """The oracle's restatement of kornia.morphology.closing / kornia.filters.sobel (models/raytracer.py:557, 569) against an
    independent implementation: OpenCV (oracle/make_golden_cv2.py).  Closing bit-exact, Sobel magnitude within 5e-7."""
    g = golden("morph_cv2")
    assert int(g["n"]) >= 8
    for i in range(int(g["n"])):
        d = torch.from_numpy(g[f"depth{i}"])
        assert np.array_equal(O.morph_closing3(d).numpy(), g[f"closing{i}"]), f"closing, image {i}"
        assert np.abs(O.sobel_magnitude(d).numpy() - g[f"sobel{i}"]).max() <= 5e-7, f"sobel, image {i}"


def test_erosion_restatement_matches_opencv(golden):
    """kornia.morphology.erosion(mask, ones(11, 11)) of ssim_loss_fn (models/image_losses.py:154) restated by
    oracle.erode_mask, against cv2.erode (default border: outside pixels never lose) -- bit-exact."""
    g = golden("morph_cv2")
    for i in range(int(g["n_masks"])):
        m = torch.from_numpy(g[f"mask{i}"])[None, None]
        assert np.array_equal(O.erode_mask(m, 11)[0, 0].numpy(), g[f"erode11_{i}"]), f"mask {i}"


def test_patch_losses_match_the_reference(golden):
    """PyramidL2Loss / ssim_loss_fn (models/image_losses.py:13-158) restated by the oracle, against values and gradients the
    real reference module produced (oracle/make_golden_losses.py), including odd sizes, a masked and an unmasked SSIM."""
    g = golden("losses")
    close(O.gauss7().numpy(), g["gauss7"], 1e-8)
    for c in [str(x) for x in g["cases"]]:
        pred = T(g[f"{c}.pred"]).requires_grad_(True)
        gt = T(g[f"{c}.gt"])
        mask = T(g[f"{c}.mask"]) if f"{c}.mask" in g else None
        lp = O.pyramid_l2_loss(pred, gt)
        gp, = torch.autograd.grad(lp, pred)
        assert abs(float(lp) - float(g[f"{c}.pyr"])) <= 2e-6 * abs(float(g[f"{c}.pyr"])), c
        close(gp.numpy(), g[f"{c}.pyr_grad"], 1e-8, 1e-5)
        ls = O.ssim_loss(pred, gt, mask)
        gs, = torch.autograd.grad(ls, pred)
        assert abs(float(ls) - float(g[f"{c}.ssim"])) <= 2e-6, c
        close(gs.numpy(), g[f"{c}.ssim_grad"], 2e-8, 1e-4)


def test_full_step_with_the_reference_loss(golden):
    """stage2_step(image_loss="reference"): PyramidL2 + SSIM + roughness range + eikonal, the loss the reference trains with
    (render_surface.py:594-613), against a step of the real reference modules (oracle/make_golden_losses.py)."""
    g = golden("step_refloss_h256")
    p = _trace_net()
    torch.manual_seed(0)
    nets = O.make_material_dict()
    with torch.no_grad():
        nets["specular_roughness_network"]["lin4.bias"].add_(float(g["rough_bias_shift"]))
    light = torch.tensor(32.0, requires_grad=True)
    for d in [p] + list(nets.values()):
        for v in d.values():
            v.requires_grad_(True)
    cam = O.OCamera.fixture().crop(32, 32, tuple(int(v) for v in g["ul"]))
    loss, res = O.stage2_step(p, nets, light, cam, T(g["target"]), T(g["eik_points"]), image_loss="reference")
    assert (res["convergent_mask"].numpy() == g["mask"]).all()
    assert int(g["n_rough"]) > 0
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    allp = [("sdf." + k, v) for k, v in p.items()]
    for nm, d in nets.items():
        allp += [(f"{nm}.{k}", v) for k, v in d.items()]
    allp.append(("point_light_network.light", light))
    for k, v in allp:
        s = g["gsum." + k]
        nrm = v.grad.double().pow(2).sum().sqrt().item()
        assert abs(nrm - s[2]) <= 1e-3 * max(s[2], 1e-12), (k, nrm, s[2])
        if "g." + k in g:
            ref = g["g." + k]
            close(v.grad.numpy(), ref, 1e-4 * np.abs(ref).max(), 1e-3)


def _composite_inputs(g, requires_grad=True):
    names = ["diffuse_albedo", "specular_albedo", "specular_roughness", "metallic", "dielectric", "metallic_eta", "metallic_k",
             "dielectric_eta"]
    P = {k: T(g["p." + k]).requires_grad_(requires_grad) for k in names}
    return T(g["light"]).requires_grad_(requires_grad), T(g["dist"]).requires_grad_(requires_grad), \
        T(g["normal"]).requires_grad_(requires_grad), T(g["viewdir"]), P


def test_composite_renderer_forward_backward(golden):
    """oracle.composite_shade against the real CompositeRenderer.forward (models/renderer_ggx.py:781-858), values of all five
    outputs and gradients w.r.t. every differentiable input, including rows that sit on the clamps."""
    g = golden("composite")
    light, dist, n, v, P = _composite_inputs(g)
    out = O.composite_shade(light, dist, n, v, P)
    for k in ("rgb", "diffuse_rgb", "specular_rgb", "metallic_rgb", "dielectric_rgb"):
        close(out[k].detach().numpy(), g["out." + k], 1e-7, 2e-6)
    loss = sum((out[k] * T(g["up." + k])).sum() for k in ("rgb", "specular_rgb", "metallic_rgb", "dielectric_rgb", "diffuse_rgb"))
    names = ["diffuse_albedo", "specular_albedo", "specular_roughness", "metallic_eta", "metallic_k", "dielectric_eta"]
    grads = torch.autograd.grad(loss, [light, dist, n] + [P[k] for k in names])
    for k, gr in zip(["light", "dist", "normal"] + names, grads):
        ref = g["g." + k]
        close(gr.numpy(), ref, 1e-6 * max(np.abs(ref).max(), 1e-12), 2e-5)


# ---------------------------------------------------------------------------------------------------------------- stage-1 NeuS
def _neus_loss(out, target, mask):
    """render_volume.py:262-283: L1 colour + 0.1 eikonal + 0.1 BCE mask (the weights make_golden_neus.py used)."""
    color_loss = (out["color_fine"] - target).abs().sum() / target.shape[0]
    mask_loss = torch.nn.functional.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    return color_loss + 0.1 * out["gradient_error"] + 0.1 * mask_loss


@pytest.mark.parametrize("case", ["a", "b"])
def test_neus_render_matches_the_reference_renderer(golden, case):
    """oracle.neus_render (restated models/renderer.py:128-453, incl. the colour network's skip layer and the background NeRF)
    against the REAL NeuSRenderer: outputs and the stage-1 loss gradients of every parameter."""
    g = golden("neus")
    pre = case + "."
    grp = lambda m: {k[len(pre + "w." + m + "."):]: T(v).requires_grad_(True) for k, v in g.items() if k.startswith(pre + "w." + m + ".")}
    sdf_p, color_p, nerf_p = grp("sdf"), grp("color"), grp("nerf")
    var = grp("dev")["variance"]
    n_outside, perturb, has_bg, anneal = g[pre + "cfg"]
    kw = dict(n_samples=16, n_importance=16, n_outside=int(n_outside), up_sample_steps=4, cos_anneal_ratio=float(anneal),
              background_rgb=torch.ones(1, 3) if has_bg else None)
    if perturb > 0:
        kw.update(t_rand=T(g[pre + "t_rand"]), t_rand_outside=T(g[pre + "t_rand_out"]))
    out = O.neus_render(sdf_p, color_p, var, nerf_p if n_outside > 0 else None, T(g[pre + "o"]), T(g[pre + "d"]),
                        T(g[pre + "near"]), T(g[pre + "far"]), **kw)
    for k in ("color_fine", "weights", "weight_sum", "weight_max", "cdf_fine", "gradients", "s_val", "inside_sphere"):
        close(out[k].detach().numpy(), g[pre + k], 2e-6, 1e-5)
    close(out["gradient_error"].detach().numpy(), g[pre + "gradient_error"], 1e-7, 1e-5)
    loss = _neus_loss(out, T(g[pre + "target"]), T(g[pre + "mask"]))
    assert abs(float(loss) - float(g[pre + "loss"])) <= 1e-5 * abs(float(g[pre + "loss"]))
    named = [("sdf." + k, v) for k, v in sdf_p.items()] + [("color." + k, v) for k, v in color_p.items()] + [("dev.variance", var)]
    if n_outside > 0:
        named += [("nerf." + k, v) for k, v in nerf_p.items()]
    grads = torch.autograd.grad(loss, [v for _, v in named], allow_unused=True)
    checked = 0
    for (k, _), got in zip(named, grads):
        if pre + "g." + k not in g:
            continue
        ref = g[pre + "g." + k]
        close(got.numpy(), ref, 2e-5 * max(np.abs(ref).max(), 1e-6), 1e-4)
        checked += 1
    assert checked >= 50
