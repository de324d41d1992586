"""Times the per-layer tcgen05 GEMM (ironb_gemm_nt, modes 1 / 2) and the weight-gradient GEMM at the shapes of the step,
and prints the in-kernel clock64 timeline of CTA (0,0).   python tests/probe_gemm.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iron_b200 import _lib  # noqa: E402

dev = "cuda:0"
lib = _lib.load()
lib.ironb_debug_mlp_timeline(None, 0)


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts), float(np.median(ts))


for M, N, K in ((4096, 512, 512), (2048, 512, 512), (12288, 512, 512), (4096, 256, 256)):
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev) / np.sqrt(K)
    C = torch.empty(M, N, device=dev)
    for mode in (1, 2):
        f = lambda: _lib.check(lib.ironb_gemm_nt(_lib.ptr(A), K, _lib.ptr(B), K, M, N, K, _lib.ptr(C), N, mode, _lib.stream()), "g")
        best, med = timeit(f)
        f()
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 512)()
        lib.ironb_debug_mlp_timeline(buf, 512)
        v = list(buf)[256:265]
        print(f"gemm_nt mode {mode} {M}x{N}x{K}: best {best:.1f} us median {med:.1f} us  ({2.0 * M * N * K / best / 1e6:.1f} TFLOP/s) "
              f"timeline [start, setup, first-full, first-conv, mma-issued, split-done, acc-ready, staged, end]: {[x - v[0] for x in v]}")
    ref = torch.matmul(A, B.t())
    tb, tm = timeit(lambda: torch.matmul(A, B.t(), out=ref))
    print(f"   torch.matmul fp32 (cuBLAS): best {tb:.1f} us")
for M, Nd, Kd in ((4096, 512, 512), (2048, 512, 512)):
    A = torch.randn(M, Nd, device=dev) / np.sqrt(M)
    B = torch.randn(M, Kd, device=dev)
    C = torch.zeros(Nd, Kd, device=dev)
    scratch = torch.empty(int(lib.ironb_gemm_tn_scratch_bytes(M, Nd, Kd)), dtype=torch.uint8, device=dev)
    f = lambda: _lib.check(lib.ironb_gemm_tn(_lib.ptr(A), Nd, _lib.ptr(B), Kd, M, Nd, Kd, _lib.ptr(C), Kd, 1, _lib.ptr(scratch), _lib.stream()), "t")
    best, med = timeit(f)
    print(f"gemm_tn (wgrad: 2 transposes + split-K GEMM) {M}x{Nd}x{Kd}: best {best:.1f} us median {med:.1f} us")
