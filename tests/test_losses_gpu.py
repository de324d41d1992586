"""Patch losses: iron_b200.PyramidL2Loss / ssim_loss_fn (csrc/losses.cu) against values and gradients produced
by the REAL reference module models/image_losses.py (tests/golden/losses.npz, oracle/make_golden_losses.py) and, at the
training patch sizes, against the oracle's restatement on the same seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import iron_oracle as O
from util import T, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_patch_losses_against_reference_golden(golden):
    import iron_b200 as ib
    g = golden("losses")
    pyr = ib.PyramidL2Loss(use_cuda=True)
    for c in [str(x) for x in g["cases"]]:
        gt = T(g[f"{c}.gt"]).to(DEV)
        mask = T(g[f"{c}.mask"]).to(DEV) if f"{c}.mask" in g else None
        # (1) contiguous [1, 3, H, W]; (2) the layout the training loop passes: a permuted view of an [H, W, 3] buffer
        for layout in ("nchw", "hwc-view"):
            base = T(g[f"{c}.pred"]).to(DEV)
            if layout == "hwc-view":
                hwc = base[0].permute(1, 2, 0).contiguous().requires_grad_(True)
                pred = hwc.permute(2, 0, 1).unsqueeze(0)
                leaf = hwc
            else:
                pred = base.clone().requires_grad_(True)
                leaf = pred
            lp = pyr(pred, gt)
            ls = ib.ssim_loss_fn(pred, gt, mask)
            gp, = torch.autograd.grad(lp, leaf)
            gs, = torch.autograd.grad(ls, leaf)
            if layout == "hwc-view":
                gp, gs = gp.permute(2, 0, 1).unsqueeze(0), gs.permute(2, 0, 1).unsqueeze(0)
            assert abs(float(lp) - float(g[f"{c}.pyr"])) <= 5e-6 * abs(float(g[f"{c}.pyr"])), (c, layout, float(lp), float(g[f"{c}.pyr"]))
            assert_close(gp.cpu().numpy(), g[f"{c}.pyr_grad"], 1e-8, 2e-5, what=f"pyramid grad {c} {layout}")
            assert abs(float(ls) - float(g[f"{c}.ssim"])) <= 5e-6, (c, layout, float(ls), float(g[f"{c}.ssim"]))
            assert_close(gs.cpu().numpy(), g[f"{c}.ssim_grad"], 5e-8, 2e-4, what=f"ssim grad {c} {layout}")
        print(f"case {c}: pyramid {float(lp):.7f} / {float(g[f'{c}.pyr']):.7f}  ssim {float(ls):.7f} / {float(g[f'{c}.ssim']):.7f}")


@pytest.mark.parametrize("S", [128, 256])
def test_patch_losses_at_training_patch_sizes(S):
    """128 x 128 is the reference's default patch (render_surface.py:50), 256 x 256 the BASELINE configs[3] patch."""
    import iron_b200 as ib
    gen = torch.Generator().manual_seed(S)
    gt = torch.rand(1, 3, S, S, generator=gen) * 0.8
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    mask = (((xx - 0.5 * S) ** 2 + (yy - 0.45 * S) ** 2) < (0.4 * S) ** 2)[None, None]
    pred0 = ((gt + 0.1 * torch.randn(1, 3, S, S, generator=gen)) * mask.float())
    pc = pred0.clone().requires_grad_(True)
    lo = O.pyramid_l2_loss(pc, gt) + O.ssim_loss(pc, gt, mask)
    go, = torch.autograd.grad(lo, pc)
    pg = pred0.to(DEV).requires_grad_(True)
    lg = ib.PyramidL2Loss()(pg, gt.to(DEV)) + ib.ssim_loss_fn(pg, gt.to(DEV), mask.to(DEV))
    gg, = torch.autograd.grad(lg, pg)
    assert abs(float(lg) - float(lo)) <= 1e-5 * abs(float(lo)), (float(lg), float(lo))
    assert_close(gg.cpu().numpy(), go.numpy(), 5e-8, 2e-4, what="loss gradient")


def test_patch_losses_reject_what_is_not_built():
    import iron_b200 as ib
    x = torch.rand(1, 3, 8, 8, device=DEV)
    with pytest.raises(RuntimeError, match="16 x 16"):
        ib.PyramidL2Loss()(x, x)
    with pytest.raises(RuntimeError, match="11 x 11"):
        ib.ssim_loss_fn(x, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        ib.ssim_loss_fn(x.cpu(), x.cpu())
