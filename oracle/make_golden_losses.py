"""Golden vectors for the patch losses from the REAL reference module models/image_losses.py (PyramidL2Loss,
ssim_loss_fn), run on CPU in the build container.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_losses.py      -> tests/golden/losses.npz

kornia (not installable here) is stubbed: the one call on this path, kornia.morphology.erosion(mask, ones(11, 11))
(models/image_losses.py:154), is restated by oracle.erode_mask and pinned independently against cv2.erode
(oracle/make_golden_cv2.py)."""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import iron_oracle as O   # noqa: E402


def import_losses():
    for name in ("kornia", "icecream"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "icecream":
                m.ic = lambda *a, **k: None
            else:
                m.morphology = types.SimpleNamespace(
                    erosion=lambda x, kernel: O.erode_mask(x > 0.5, int(kernel.shape[-1])).float())
            sys.modules[name] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import models.image_losses as L
    return L


def blob_mask(h, w, rng, holes=0.01):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    m = ((xx - 0.55 * w) ** 2 / (0.42 * w) ** 2 + (yy - 0.5 * h) ** 2 / (0.46 * h) ** 2) < 1.0
    m &= ~(rng.random((h, w)) < holes)
    return m


def main():
    L = import_losses()
    pyr = L.PyramidL2Loss(use_cuda=False)
    rng = np.random.default_rng(3)
    data = {"gauss7": pyr.f[0, 0].numpy().copy()}
    cases = [("a", 64, 64, "blob"), ("b", 37, 53, "blob"), ("c", 128, 128, "full"), ("d", 48, 40, "none"), ("e", 16, 16, "full")]
    for name, h, w, mk in cases:
        g = torch.Generator().manual_seed(h * 1000 + w)
        gt = torch.rand(1, 3, h, w, generator=g) * 0.8
        pred = (gt + 0.15 * torch.randn(1, 3, h, w, generator=g)).clamp(0, 2).requires_grad_(True)
        if mk == "blob":
            mask = torch.from_numpy(blob_mask(h, w, rng))[None, None]
        elif mk == "full":
            mask = torch.ones(1, 1, h, w, dtype=torch.bool)
        else:
            mask = None
        if mask is not None:                      # rendered colours are zero outside the hit mask (render_surface.py:136-156)
            pred = (pred.detach() * mask.float()).requires_grad_(True)
        lp = pyr(pred, gt)
        gp, = torch.autograd.grad(lp, pred)
        ls = L.ssim_loss_fn(pred, gt, mask)
        gs, = torch.autograd.grad(ls, pred)
        data.update({f"{name}.pred": pred.detach().numpy(), f"{name}.gt": gt.numpy(),
                     f"{name}.pyr": np.float64(lp.item()), f"{name}.pyr_grad": gp.numpy(),
                     f"{name}.ssim": np.float64(ls.item()), f"{name}.ssim_grad": gs.numpy()})
        if mask is not None:
            data[f"{name}.mask"] = mask.numpy()
        print(name, (h, w), mk, "pyramid", float(lp), "ssim", float(ls))
    data["cases"] = np.array([c[0] for c in cases])
    path = os.path.join(HERE, "..", "tests", "golden", "losses.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")


def step_with_reference_loss(L):
    """A full stage-2 step of the REAL reference modules (render_camera is_training=True) with the loss the reference trains
    with (render_surface.py:594-613: PyramidL2 + 1.0 * SSIM + 0.1 * roughness range, + 0.1 * eikonal), on the 32x32
    silhouette crop of tests/golden/step_h256.npz (same weights, target and eikonal samples) -> step_refloss_h256.npz"""
    sys.path.insert(0, HERE)
    import make_golden as MG
    fields, raytracer, renderer_ggx, rendering_func, network_conf = MG.import_reference()
    torch.set_num_threads(8)
    torch.manual_seed(0)
    nets = MG.build_ggx_nets(fields)
    torch.manual_seed(0)
    sdf_net = fields.SDFNetwork(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5,
                                scale=1.0, geometric_init=True, weight_norm=True)
    MG.perturb(sdf_net, 0.005, seed=1)
    # make the roughness-range term live: shift the roughness head so that some hit pixels exceed 0.5 (:609-613)
    with torch.no_grad():
        nets["specular_roughness_network"].lin4.bias.add_(6.0)
    nets["point_light_network"] = network_conf.PointLightNetwork()
    nets["point_light_network"].set_light(32.0)
    rt = raytracer.RayTracer()
    render_fn = MG.make_render_fn(renderer_ggx.GGXColocatedRenderer(use_cuda=False), rendering_func.get_materials)
    ul = (448, 230)
    cam, _, _ = MG.fixture_camera(raytracer).crop_region(32, 32, ul_corner=ul)
    results = raytracer.render_camera(cam, sdf_net, rt, nets, render_fn, fill_holes=False, handle_edges=False, is_training=True)
    mask = results["convergent_mask"]
    tgt = torch.rand(32, 32, 3, generator=torch.Generator().manual_seed(11)) * 0.5
    eik_pts = torch.empty(32 * 32 // 2, 3).uniform_(-1.0, 1.0, generator=torch.Generator().manual_seed(12))
    eg = sdf_net.gradient(eik_pts.clone()).view(-1, 3)
    eik_cnt = eg.shape[0]
    eik = ((eg.norm(dim=-1) - 1) ** 2).sum()
    pred_img = results["color"].permute(2, 0, 1).unsqueeze(0)[:, :3]
    gt_img = tgt.permute(2, 0, 1).unsqueeze(0)
    l_pyr = L.PyramidL2Loss(use_cuda=False)(pred_img, gt_img)
    l_ssim = 1.0 * L.ssim_loss_fn(pred_img, gt_img, mask.unsqueeze(0).unsqueeze(0))
    hn = results["normal"][mask]
    eik_cnt += hn.shape[0]
    eik = eik + ((hn.norm(dim=-1) - 1) ** 2).sum()
    rough = results["specular_roughness"][mask]
    rough = rough[rough > 0.5]
    l_rough = (rough - 0.5).mean() * 0.1 if rough.numel() > 0 else torch.zeros(())
    loss = l_pyr + l_ssim + eik / eik_cnt * 0.1 + l_rough
    loss.backward()
    out = dict(ul=np.array(ul), target=tgt.numpy(), eik_points=eik_pts.numpy(), loss=loss.detach().numpy(), mask=mask.numpy(),
               l_pyr=l_pyr.detach().numpy(), l_ssim=l_ssim.detach().numpy(), l_rough=np.float32(float(l_rough)),
               n_rough=np.int64(rough.numel()), rough_bias_shift=np.float32(6.0))
    out["res.color"] = results["color"].detach().numpy()
    out["res.specular_roughness"] = results["specular_roughness"].detach().numpy()
    allp = [("sdf." + k, p_) for k, p_ in sdf_net.named_parameters()]
    for nm in ("diffuse_albedo_network", "specular_albedo_network", "specular_roughness_network", "point_light_network"):
        allp += [(nm + "." + k, p_) for k, p_ in nets[nm].named_parameters()]
    for k, p_ in allp:
        gr_ = p_.grad
        out["gsum." + k] = np.array([gr_.double().sum().item(), gr_.double().abs().sum().item(), gr_.double().pow(2).sum().sqrt().item()])
        if gr_.numel() <= 1024 or k.endswith("lin8.weight_v") or k.endswith("lin0.weight_v"):
            out["g." + k] = gr_.numpy()
    path = os.path.join(HERE, "..", "tests", "golden", "step_refloss_h256.npz")
    np.savez_compressed(path, **out)
    print("step_refloss_h256.npz hits", int(mask.sum()), "loss", float(loss), "pyr", float(l_pyr), "ssim", float(l_ssim),
          "rough", float(l_rough), "n_rough", int(rough.numel()), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    step_with_reference_loss(import_losses())
