"""How much of a stage-2 step is host-side issue time?  Prints issue time (no sync) vs device time per step."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iron_b200 as ib
from oracle import iron_oracle as O
dev = torch.device("cuda:0")
H, S = 512, 64
torch.manual_seed(0); nets = ib.init_rendering_network_dict("ggx")
torch.manual_seed(0)
sdf = ib.SDFNetwork(3, 257, H, 8, skip_in=[4], multires=6).to(dev)
nets["point_light_network"].set_light(32.0)
K = torch.tensor(O.FIXTURE_K, dtype=torch.float64).reshape(4, 4).float().to(dev)
W2C = torch.tensor(O.FIXTURE_W2C, dtype=torch.float64).reshape(4, 4).float().to(dev)
cam = ib.Camera(512, 512, K, W2C).crop_region(S, S, ul_corner=(224, 224))[0]
rend = ib.GGXColocatedRenderer(use_cuda=True); rf = ib.make_render_fn(rend); rt = ib.RayTracer()
target = torch.rand(S, S, 3, device=dev) * 0.5
eik = torch.empty(S * S // 2, 3, device=dev).uniform_(-1, 1)
params = list(sdf.parameters()) + [p for k in nets for p in nets[k].parameters()]
def step():
    for p in params: p.grad = None
    return ib.stage2_step(sdf, nets, rt, rf, cam, target, eik)
for _ in range(3): step()
torch.cuda.synchronize()
iss, tot = [], []
for _ in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    iss.append(t1 - t0); tot.append(t2 - t0)
print(f"issue {1e3*sum(iss)/10:.2f} ms/step, total {1e3*sum(tot)/10:.2f} ms/step")
# phases
def timed(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return r, 1e3 * (t1 - t0), 1e3 * (t2 - t0)
res, a, b = timed(lambda: ib.raytrace_camera(cam, sdf, rt, max_num_rays=50000)); print(f"trace: issue {a:.2f} total {b:.2f}")
def shade():
    r = dict(res); ib.render_normal_and_color(r, sdf, nets, rf, is_training=True); return r
r2, a, b = timed(shade); print(f"shade fwd: issue {a:.2f} total {b:.2f}")
def eikf(): return sdf.gradient(eik)
eg, a, b = timed(eikf); print(f"eik fwd: issue {a:.2f} total {b:.2f}")
def bwd():
    loss = ((r2["color"] - target) ** 2).sum() + ((eg.norm(dim=-1) - 1) ** 2).sum()
    loss.backward()
_, a, b = timed(bwd); print(f"backward: issue {a:.2f} total {b:.2f}")
import cProfile, pstats, io
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
