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


if __name__ == "__main__":
    main()
