"""Times SDFNetwork.get_all forward / double backward and one material net (CUDA events, warm) per GEMM mode and batch size.
Prints only.        python tests/probe_getall.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iron_b200 as ib
from iron_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)
nets = ib.init_rendering_network_dict("ggx")
torch.manual_seed(0)
sdf = ib.SDFNetwork(d_in=3, d_out=257, d_hidden=512, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0, geometric_init=True, weight_norm=True).to(dev)
mat = nets["specular_albedo_network"]


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for M in (4096, 65536):
    x = (torch.rand(M, 3, device=dev) - 0.5)
    nrm = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
    feats = torch.randn(M, 256, device=dev) * 0.3
    gy, gf, gn = torch.randn(M, 1, device=dev), torch.randn(M, 256, device=dev), torch.randn(M, 3, device=dev)
    for mode, name in ((1, "3xTF32"), (2, "fp16x2 fwd + 3xTF32 bwd")):
        lib.ironb_set_gemm_mode(mode)
        state = {}

        def fwd():
            state["out"] = sdf.get_all(x, is_training=True)

        def bwd():
            y, f, n = sdf.get_all(x, is_training=True)
            torch.autograd.grad([y, f, n], list(sdf.parameters()), [gy, gf, gn])

        def mfwd():
            state["m"] = mat(x, nrm, None, feats)

        def mbwd():
            o = mat(x, nrm, None, feats)
            torch.autograd.grad(o, list(mat.parameters()), torch.ones_like(o))

        print("  ..", M, name, flush=True); tf = timeit(fwd); print("  fwd ok", flush=True); tb = timeit(bwd); print("  bwd ok", flush=True); mf = timeit(mfwd); print("  mfwd ok", flush=True); mb = timeit(mbwd)
        fl = 2.0 * (7 * 512 * 512 + 257 * 512) * M
        print(f"M={M:6d} {name:24s}: get_all fwd (F+Q) {tf:8.3f} ms = {2 * fl / tf / 1e9:6.1f} TFLOP/s | fwd+bwd {tb:8.3f} ms (bwd {tb - tf:8.3f} = "
              f"{4 * fl / (tb - tf) / 1e9:6.1f} TFLOP/s) | material net fwd {mf:7.3f} ms, fwd+bwd {mb:7.3f} ms")
lib.ironb_set_gemm_mode(2)
