"""Calibration of the truncation de-bias of the default tracer's MLP (csrc/mlp_h16.cu: mlp16_debias): error of ONE sdf
evaluation per point against an fp64 evaluation of the same folded weights, for a sweep of the de-bias factor g, on the
seed-0 geometric init (H = 256 / 512), a perturbed 'shape' variant and a scaled-activation variant.  Prints only.

    python tests/probe_debias.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import iron_b200 as ib  # noqa: E402
from iron_b200 import _lib  # noqa: E402
from util import perturb  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


def mknet(H, sigma=0.0):
    torch.manual_seed(0)
    net = ib.SDFNetwork(d_in=3, d_out=257, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True)
    if sigma > 0:
        perturb(net, sigma, seed=1)
    return net.to(dev)


def sdf64(net, x):
    h = x.double()
    pe = [h]
    for k in range(6):
        pe += [torch.sin(h * 2.0 ** k), torch.cos(h * 2.0 ** k)]
    pe = torch.cat(pe, -1)
    h = pe
    for l in range(9):
        lin = getattr(net, f"lin{l}")
        v, g, b = lin.weight_v.double(), lin.weight_g.double(), lin.bias.double()
        W = g * v / v.norm(dim=1, keepdim=True)
        if l == 4:
            h = torch.cat([h, pe], -1) / np.sqrt(2)
        h = h @ W.t() + b
        if l < 8:
            h = torch.nn.functional.softplus(h, beta=100)
    return h[:, 0]


xs = (torch.rand(16384, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(5)) - 0.5) * 1.6
d = torch.zeros_like(xs); d[:, 2] = 1.0
z = torch.zeros(xs.shape[0], device=dev)
wm = torch.zeros(xs.shape[0], dtype=torch.bool, device=dev)
rt = ib.RayTracer()
for name, H, sigma in (("H=256 seed0", 256, 0.0), ("H=256 perturbed", 256, 0.005), ("H=512 seed0", 512, 0.0),
                       ("H=512 perturbed", 512, 0.002)):
    net = mknet(H, sigma)
    ref = sdf64(net, xs)
    lib.ironb_set_trace_mode(0)
    e = rt(net, xs, d, z, z + 1.0, wm)["sdf"].double() - ref
    print(f"{name:16s} ffma            : signed mean {e.mean():+.2e}  mean|err| {e.abs().mean():.2e}  max {e.abs().max():.2e}")
    lib.ironb_set_trace_mode(2)
    for g in (0.0, 0.2, 0.25, 0.3, 0.5, 1.0):
        lib.ironb_set_mlp_debias(g)
        e = rt(net, xs, d, z, z + 1.0, wm)["sdf"].double() - ref
        print(f"{name:16s} fp16x2 g={g:4.2f}   : signed mean {e.mean():+.2e}  mean|err| {e.abs().mean():.2e}  max {e.abs().max():.2e}")
lib.ironb_set_mlp_debias(0.25)
