"""Prints (does not assert) accuracy and speed of the three tracer MLP arithmetics on the GPU box:

  * sdf error of ONE MLP evaluation per point (RayTracer with work_mask = False finalises every ray after the first
    evaluation) against the reference's golden forward values (tests/golden/sdf_seeded.npz) and against an fp64
    evaluation of the same folded weights;
  * tracer time, hit mask and distances at the bench workload (configs[1]) and at the 65,536-ray crop.

    python tests/probe_mlp_h16.py [--big]          (IRONB_MLP_NHH=2 selects the two-accumulator diagnostic variant)
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iron_b200 as ib  # noqa: E402
from iron_b200 import _lib  # noqa: E402
from oracle import iron_oracle as O  # noqa: E402  (camera constants only)

ap = argparse.ArgumentParser()
ap.add_argument("--big", action="store_true")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--timeline-only", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = _lib.load()
MODES = ((0, "ffma"), (2, "fp16x2"))
print("IRONB_MLP_NHH =", os.environ.get("IRONB_MLP_NHH", "1"))


def mknet(H):
    torch.manual_seed(0)
    return ib.SDFNetwork(d_in=3, d_out=257, d_hidden=H, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                         geometric_init=True, weight_norm=True).to(dev)


def sdf64(net, x):
    """fp64 evaluation of the sdf row with the module's effective (weight-normalised) weights."""
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


g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "sdf_seeded.npz")))
for H in (() if a.timeline_only else (256, 512)):
    net = mknet(H)
    x = torch.from_numpy(g[f"h{H}.x"]).to(dev)
    xs = torch.cat([x, (torch.rand(8192, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(5)) - 0.5) * 1.6])
    ref64 = sdf64(net, xs)
    rt = ib.RayTracer()
    d = torch.zeros_like(xs); d[:, 2] = 1.0
    z = torch.zeros(xs.shape[0], device=dev)
    wm = torch.zeros(xs.shape[0], dtype=torch.bool, device=dev)
    for mode, name in MODES:
        lib.ironb_set_trace_mode(mode)
        res = rt(net, xs, d, z, z + 1.0, wm)
        torch.cuda.synchronize()
        e64 = (res["sdf"].double() - ref64)
        eg = np.abs(res["sdf"][: x.shape[0]].cpu().numpy() - g[f"h{H}.fwd"][:, 0])
        print(f"H={H} {name:7s} one eval, {xs.shape[0]} pts: |err| vs fp64 max {e64.abs().max():.2e} mean {e64.abs().mean():.2e} "
              f"signed mean {e64.mean():+.2e}; vs golden(ref fp32) max {eg.max():.2e} mean {eg.mean():.2e}")

K = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float().to(dev)
W2C = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float().to(dev)
cam512 = ib.Camera(512, 512, K, W2C)
cases = [(512, 64, (224, 224)), (256, 64, (224, 224)), (512, 64, (330, 236))]
if a.big:
    cases += [(512, 256, (128, 128)), (256, 256, (128, 128))]
if a.timeline_only:
    cases = []
for H, S, ul in cases:
    net = mknet(H)
    cam, _, _ = cam512.crop_region(S, S, ul_corner=ul)
    uv = cam.get_uv()
    out = {}
    for mode, name in MODES:
        lib.ironb_set_trace_mode(mode)
        rt = ib.RayTracer()
        rt.collect_stats = True
        ts = []
        for i in range(a.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            rt.last_stats = None
            e0.record()
            res = ib.raytrace_pixels(net, rt, uv, cam, max_num_rays=200000)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[mode] = res
        st = rt.last_stats.cpu().tolist()
        evals = st[0] + st[1] + st[2]
        tf = evals * 2.0 * (7 * H * H + H) / (min(ts[1:]) * 1e-3) / 1e12
        print(f"H={H} {S}x{S}@{ul} {name:7s}: best {min(ts[1:]):8.3f} ms  median {sorted(ts[1:])[len(ts[1:]) // 2]:8.3f} ms  "
              f"evals {evals}  {tf:6.1f} TFLOP/s  hits {int(res['convergent_mask'].sum())}")
    m0 = out[0]["convergent_mask"]
    for mode, name in MODES[1:]:
        m1 = out[mode]["convergent_mask"]
        both = m0 & m1
        dd = (out[0]["distance"] - out[mode]["distance"])[both].abs()
        print(f"    {name} vs ffma: mask differs on {int((m0 != m1).sum())} of {m0.numel()} rays; |d dist| max {float(dd.max()) if dd.numel() else 0:.2e} "
              f"within 1e-4: {float((dd <= 1e-4).float().mean()) if dd.numel() else 1:.5f}")

if os.environ.get("IRONB_MLP_DBG"):
    import ctypes
    lib.ironb_set_trace_mode(2)
    net = mknet(512)
    cam, _, _ = cam512.crop_region(64, 64, ul_corner=(224, 224))
    lib.ironb_debug_mlp_timeline(None, 0)
    # first launch of a trace (4096 rows): stop after the first evaluation so the stamps are not overwritten
    rt = ib.RayTracer(sphere_tracing_iters=0)
    o, dd_, _ = cam.get_rays(cam.get_uv())
    hit, tmin, tmax = ib.intersect_sphere(o, dd_, 1.0)
    rt(net, o.reshape(-1, 3), dd_.reshape(-1, 3), tmin.reshape(-1), tmax.reshape(-1), torch.zeros_like(hit.reshape(-1)))
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    lib.ironb_debug_mlp_timeline(buf, 64)
    v = list(buf)
    t0 = v[0]
    print("mlp_h16 timeline, cluster 0 / rank 0 / first tile (cycles from layer-0 start): wait-full-start, first-full, mma-issued, "
          "acc-ready, epi-done, half-0 staged (store warp), half-0 ready-signalled, half-1 ready-signalled")
    for l in range(8):
        print(l, [x - t0 if x else 0 for x in v[l * 8:l * 8 + 8]])
