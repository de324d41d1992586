"""Runs only the tracer of the bench workload (configs[1] rays, H=512 by default) a few times: the target of ncu
captures of the tracer kernels.   python tests/profile_trace.py [--hidden 512] [--patch 64] [--reps 3] [--mode 1]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iron_b200 as ib  # noqa: E402
from iron_b200 import _lib  # noqa: E402
from oracle import iron_oracle as O  # noqa: E402  (camera constants only)

ap = argparse.ArgumentParser()
ap.add_argument("--hidden", type=int, default=512)
ap.add_argument("--patch", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--mode", type=int, default=1)
ap.add_argument("--ul", type=int, nargs=2, default=None)
a = ap.parse_args()
dev = torch.device("cuda:0")
_lib.load().ironb_set_trace_mode(a.mode)
torch.manual_seed(0)
net = ib.SDFNetwork(d_in=3, d_out=257, d_hidden=a.hidden, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                    geometric_init=True, weight_norm=True).to(dev)
K = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float().to(dev)
W2C = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float().to(dev)
c = 256 - a.patch // 2
cam, _, _ = ib.Camera(512, 512, K, W2C).crop_region(a.patch, a.patch, ul_corner=tuple(a.ul) if a.ul else (c, c))
rt = ib.RayTracer()
rt.collect_stats = True
uv = cam.get_uv()
for i in range(a.reps + 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = ib.raytrace_pixels(net, rt, uv, cam)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"trace {i}: {dt * 1e3:.3f} ms, hits {int(res['convergent_mask'].sum())}/{uv.shape[0] * uv.shape[1]}")
import ctypes
lib = _lib.load()
if os.environ.get("IRONB_MLP_DBG"):
    lib.ironb_debug_mlp_timeline(None, 0)      # allocate the stamp buffer
    res = ib.raytrace_pixels(net, rt, uv, cam)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    lib.ironb_debug_mlp_timeline(buf, 64)
    v = list(buf)
    t0 = v[0]
    print("fused MLP timeline of the LAST launch (cycles from layer-0 start; cols: wait-full-start, first-full, mma-issued, acc-ready, epi-done, barrier-exit)")
    for l in range(8):
        print(l, [x - t0 for x in v[l * 8:l * 8 + 6]])
st = rt.last_stats.cpu().tolist()
print("stats (summed over calls):", st)
