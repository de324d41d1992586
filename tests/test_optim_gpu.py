"""FusedAdam (one ironb_adam_step launch for every tensor of every parameter group) against torch.optim.Adam, the optimiser
the reference's stage-2 loop steps per network (render_surface.py:112-113, 651-653; models/network_conf.py:707-716), and
against the oracle's numpy restatement."""
import numpy as np
import pytest
import torch

from oracle import iron_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SHAPES = [(1,), (3,), (1000,), (512, 512), (257, 512), (512, 40), (5,)]


def make(gen):
    return [torch.randn(s, generator=gen).to(DEV) for s in SHAPES]


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_adam_matches_torch_adam(wd):
    import iron_b200
    gen = torch.Generator().manual_seed(11)
    base = make(gen)
    a = [p.clone().requires_grad_(True) for p in base]
    b = [p.clone().requires_grad_(True) for p in base]
    groups = lambda ps: [{"params": ps[:3], "lr": 1e-5}, {"params": ps[3:6], "lr": 1e-4}, {"params": ps[6:], "lr": 1e-2}]
    ref = torch.optim.Adam(groups(a), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)   # the reference's lrs
    mine = iron_b200.FusedAdam(groups(b), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    ora = [(p.cpu().numpy().copy(), np.zeros(p.numel(), np.float32).reshape(p.shape), np.zeros(p.numel(), np.float32).reshape(p.shape))
           for p in base]
    lrs = [1e-5] * 3 + [1e-4] * 3 + [1e-2]
    for t in range(1, 21):
        grads = [torch.randn(p.shape, generator=gen).to(DEV) * (10.0 ** ((t % 7) - 4)) for p in base]
        skip = -1
        for i, (x, y, g) in enumerate(zip(a, b, grads)):
            x.grad = None if i == skip else g.clone()
            y.grad = None if i == skip else g.clone()
        ref.step()
        mine.step()
        if skip < 0 and t < 5:
            ora = [O.adam_step(p, g.cpu().numpy(), m, v, t, lr, 0.9, 0.999, 1e-8, wd) for (p, m, v), g, lr in zip(ora, grads, lrs)]
            for y, (p, m, v) in zip(b, ora):
                assert np.allclose(y.detach().cpu().numpy(), p, rtol=2e-6, atol=1e-7)
    torch.cuda.synchronize()
    assert mine.steps_taken == 20
    for i, (x, y) in enumerate(zip(a, b)):
        err = float((x.detach() - y.detach()).abs().max() / x.detach().abs().max().clamp_min(1e-30))
        assert err <= 2e-6, (i, err)
        # moments too (torch keeps them in optimizer.state)
        for k in ("exp_avg", "exp_avg_sq"):
            rx, ry = ref.state[x][k], mine.state[y][k]
            assert float((rx - ry).abs().max()) <= 2e-6 * float(rx.abs().max().clamp_min(1e-30)) + 1e-12, (i, k)


def test_fused_adam_refuses_cpu_and_amsgrad_like_misuse():
    import iron_b200
    with pytest.raises(RuntimeError, match="CUDA"):
        iron_b200.FusedAdam([torch.zeros(3, requires_grad=True)])
    with pytest.raises(ValueError):
        iron_b200.FusedAdam([{"params": [torch.zeros(3, device=DEV, requires_grad=True)], "betas": (0.5, 0.9)},
                             {"params": [torch.zeros(3, device=DEV, requires_grad=True)]}])


def test_fused_adam_skips_parameters_without_gradient():
    """grad is None -> parameter and moments untouched (torch.optim.Adam's behaviour).  The step count is per call, not per
    parameter: a parameter that skipped steps uses the call's count for its bias correction (documented difference)."""
    import iron_b200
    a = torch.randn(100, device=DEV).requires_grad_(True)
    b = torch.randn(50, device=DEV).requires_grad_(True)
    opt = iron_b200.FusedAdam([a, b], lr=1e-2)
    a0, b0 = a.detach().clone(), b.detach().clone()
    a.grad = torch.randn(100, device=DEV)
    opt.step()
    torch.cuda.synchronize()
    assert not torch.equal(a.detach(), a0) and torch.equal(b.detach(), b0)
    assert float(opt.state[b]["exp_avg"].abs().max()) == 0.0
