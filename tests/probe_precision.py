"""Prints (does not assert) the forward / gradient error of the CUDA SDF network against the reference golden
vectors for both GEMM arithmetics.  Run on the GPU box: python tests/probe_precision.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iron_b200  # noqa: E402
from iron_b200 import _lib  # noqa: E402

g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "sdf_seeded.npz")))
lib = _lib.load()
for H in (256, 512):
    torch.manual_seed(0)
    net = iron_b200.SDFNetwork(d_in=3, d_out=257, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                               geometric_init=True, weight_norm=True).cuda()
    x = torch.from_numpy(g[f"h{H}.x"]).cuda()
    for mode, name in ((0, "ffma"), (1, "tcgen05 3xTF32")):
        lib.ironb_set_gemm_mode(mode)
        y, f, n = net.get_all(x, is_training=False)
        fwd = torch.cat([y, f], -1).cpu().numpy()
        e = np.abs(fwd - g[f"h{H}.fwd"])
        en = np.abs(n.cpu().numpy() - g[f"h{H}.grad"])
        print(f"H={H} {name:16s} sdf max/mean err {e[:, 0].max():.2e}/{e[:, 0].mean():.2e}  feature max/mean "
              f"{e[:, 1:].max():.2e}/{e[:, 1:].mean():.2e}  grad max/mean {en.max():.2e}/{en.mean():.2e}")
