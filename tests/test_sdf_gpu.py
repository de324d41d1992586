"""SDFNetwork: CUDA forward / get_all / double backward vs the golden vectors made by the reference module,
and vs the oracle's autograd on seeded H=256 / H=512 networks."""
import numpy as np
import pytest
import torch

from oracle import iron_oracle as O
from util import T, TOL_GRAD_REL, assert_close, oracle_params, perturb, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def vt(mode, base):
    """value tolerance: tight for the fp32 FFMA GEMM, 10x for tcgen05 3xTF32 (truncating tensor-core accumulation)"""
    return base * (1.0 if mode == "ffma" else 10.0)


def small_net(golden):
    import iron_b200
    g = golden("sdf_small")
    net = iron_b200.SDFNetwork(d_in=3, d_out=17, d_hidden=64, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                               geometric_init=True, weight_norm=True)
    net.load_state_dict({k[2:]: T(v) for k, v in g.items() if k.startswith("w.")})
    return net.to(DEV), g


def test_small_forward_and_get_all(golden, gemm_mode):
    net, g = small_net(golden)
    x = T(g["x"]).to(DEV)
    with torch.no_grad():
        fwd = net(x)
    assert_close(fwd.cpu().numpy(), g["fwd"], vt(gemm_mode, 2e-6), vt(gemm_mode, 2e-6), what="forward")
    y, f, n = net.get_all(x.clone(), is_training=False)
    assert not y.requires_grad and not n.requires_grad
    assert_close(y.cpu().numpy(), g["y"], vt(gemm_mode, 2e-6), what="sdf")
    assert_close(f.cpu().numpy(), g["feat"], vt(gemm_mode, 2e-6), vt(gemm_mode, 2e-6), what="feature")
    assert_close(n.cpu().numpy(), g["grad"], vt(gemm_mode, 2e-5), vt(gemm_mode, 2e-5), what="gradient")
    assert_close(net.sdf(x).detach().cpu().numpy(), g["y"], vt(gemm_mode, 2e-6), what="sdf()")
    assert_close(net.gradient(x).detach().cpu().numpy(), g["grad"], vt(gemm_mode, 2e-5), vt(gemm_mode, 2e-5), what="gradient()")


def test_small_double_backward(golden, gemm_mode):
    """loss = <up_y,y> + <up_f,feat> + <up_n,grad>: parameter gradients (through the closed-form double backward)."""
    net, g = small_net(golden)
    y, f, n = net.get_all(T(g["x"]).to(DEV), is_training=True)
    loss = (y * T(g["up_y"]).to(DEV)).sum() + (f * T(g["up_feat"]).to(DEV)).sum() + (n * T(g["up_grad"]).to(DEV)).sum()
    loss.backward()
    for k, p in net.named_parameters():
        ref = g["g." + k]
        got = p.grad.cpu().numpy()
        # parameter gradients within 1e-3 relative (BASELINE); stated as relative L2 per tensor + elementwise bound
        assert rel_l2(got, ref) < TOL_GRAD_REL, (k, rel_l2(got, ref))
        assert_close(got, ref, 1e-4 * max(1.0, np.abs(ref).max()), 1e-3, what=k)


def test_small_partial_upstreams(golden, gemm_mode):
    """Each output alone (eikonal-only = gradient(); sdf-only; feature-only) against the oracle's autograd."""
    net, g = small_net(golden)
    p = {k: v.requires_grad_(True) for k, v in oracle_params(net).items()}
    x = T(g["x"])
    names = sorted(p)
    up_n = T(g["up_grad"])
    ref = torch.autograd.grad((O.sdf_gradient(p, x.clone()) * up_n).sum(), [p[k] for k in names], allow_unused=True)
    net.zero_grad()
    (net.gradient(x.to(DEV)) * up_n.to(DEV)).sum().backward()
    for k, r in zip(names, ref):
        got = dict(net.named_parameters())[k].grad
        r = torch.zeros_like(p[k]) if r is None else r
        assert_close(got.cpu().numpy(), r.numpy(), 1e-4 * max(1.0, float(r.abs().max())), 1e-3, what="eik " + k)
    out = O.sdf_forward(p, x)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    ref = torch.autograd.grad((out * w).sum(), [p[k] for k in names])
    net.zero_grad()
    (net(x.to(DEV)) * w.to(DEV)).sum().backward()
    for k, r in zip(names, ref):
        got = dict(net.named_parameters())[k].grad
        assert_close(got.cpu().numpy(), r.numpy(), 1e-4 * max(1.0, float(r.abs().max())), 1e-3, what="fwd " + k)


@pytest.mark.parametrize("H", [256, 512])
def test_seeded_forward_and_gradient(golden, H, gemm_mode):
    import iron_b200
    g = golden("sdf_seeded")
    torch.manual_seed(0)
    net = iron_b200.SDFNetwork(d_in=3, d_out=257, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                               geometric_init=True, weight_norm=True).to(DEV)
    x = T(g[f"h{H}.x"]).to(DEV)
    with torch.no_grad():
        fwd = net(x)
    assert_close(fwd.cpu().numpy(), g[f"h{H}.fwd"], vt(gemm_mode, 5e-6), vt(gemm_mode, 5e-6), what=f"forward H={H}")
    _, _, n = net.get_all(x, is_training=False)
    assert_close(n.cpu().numpy(), g[f"h{H}.grad"], vt(gemm_mode, 5e-5), vt(gemm_mode, 5e-5), what=f"gradient H={H}")


@pytest.mark.parametrize("H,M", [(256, 1000), (512, 300), (256, 0)])
def test_seeded_double_backward_vs_oracle(H, M, gemm_mode):
    import iron_b200
    torch.manual_seed(0)
    net = iron_b200.SDFNetwork(d_in=3, d_out=257, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                               geometric_init=True, weight_norm=True)
    perturb(net, 0.005 if H == 256 else 0.002)
    p = {k: v.requires_grad_(True) for k, v in oracle_params(net).items()}
    net = net.to(DEV)
    gen = torch.Generator().manual_seed(9)
    x = (torch.rand(M, 3, generator=gen) * 2 - 1) * 0.7
    ups = [torch.randn(M, 1, generator=gen), torch.randn(M, 256, generator=gen) * 0.1, torch.randn(M, 3, generator=gen)]
    y, f, n = net.get_all(x.to(DEV), is_training=True)
    assert y.shape == (M, 1) and f.shape == (M, 256) and n.shape == (M, 3)
    if M == 0:
        return
    (y * ups[0].to(DEV)).sum().add((f * ups[1].to(DEV)).sum()).add((n * ups[2].to(DEV)).sum()).backward()
    yo, fo, no = O.sdf_get_all(p, x.clone(), is_training=True)
    assert_close(y.detach().cpu().numpy(), yo.detach().numpy(), vt(gemm_mode, 5e-6), what="sdf")
    assert_close(n.detach().cpu().numpy(), no.detach().numpy(), vt(gemm_mode, 5e-5), vt(gemm_mode, 5e-5), what="grad")
    names = sorted(p)
    ref = torch.autograd.grad((yo * ups[0]).sum() + (fo * ups[1]).sum() + (no * ups[2]).sum(), [p[k] for k in names])
    for k, r in zip(names, ref):
        got = dict(net.named_parameters())[k].grad.cpu().numpy()
        assert rel_l2(got, r.numpy()) < TOL_GRAD_REL, (k, rel_l2(got, r.numpy()))
