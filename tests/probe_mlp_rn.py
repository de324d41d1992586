"""Prints the in-kernel timeline (IRONB debug stamps) and the tracer time of both cluster shapes of mlp_h16_kernel."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iron_b200 as ib
from iron_b200 import _lib
from oracle import iron_oracle as O
lib = _lib.load()
lib.ironb_debug_mlp_resident_clusters.restype = ctypes.c_int
lib.ironb_debug_mlp_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda:0")
for H in (512, 256):
    print(f"H={H}: resident clusters  RN=64: {lib.ironb_debug_mlp_resident_clusters(64, H)}   RN=128: {lib.ironb_debug_mlp_resident_clusters(128, H)}")
torch.manual_seed(0)
net = ib.SDFNetwork(d_in=3, d_out=257, d_hidden=512, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0, geometric_init=True, weight_norm=True).to(dev)
K = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float().to(dev)
W2C = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float().to(dev)
cam, _, _ = ib.Camera(512, 512, K, W2C).crop_region(64, 64, ul_corner=(224, 224))
o, d, _ = cam.get_rays(cam.get_uv())
hit, tmin, tmax = ib.intersect_sphere(o, d, 1.0)
args = (o.reshape(-1, 3), d.reshape(-1, 3), tmin.reshape(-1), tmax.reshape(-1))
lib.ironb_debug_mlp_timeline(None, 0)
for rn in (128, 64):
    lib.ironb_set_mlp_rn(rn)
    for nrows in (4096, 2048, 512):
        rt = ib.RayTracer(sphere_tracing_iters=0)
        a = [t[:nrows].contiguous() for t in args]
        wm = torch.zeros(nrows, dtype=torch.bool, device=dev)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rt(net, *a, wm); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        buf = (ctypes.c_longlong * 64)()
        lib.ironb_debug_mlp_timeline(buf, 64)
        v = list(buf)
        t0 = v[0]
        print(f"RN={rn} rows={nrows}: one-evaluation trace call best {min(ts)*1e3:.1f} us")
        for l in (0, 1, 2, 6, 7):
            print("   layer", l, [x - t0 if x else 0 for x in v[l * 8:l * 8 + 8]])
    rt = ib.RayTracer()
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rt(net, *args, hit.reshape(-1)); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"RN={rn}: full trace of 4096 rays best {min(ts):.3f} ms")
