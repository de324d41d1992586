"""The two tile GEMMs behind the differentiable part of the path -- fp32 FFMA tiles and the tcgen05 3xTF32 kernel
(TMA -> in-smem hi/lo split -> tcgen05.mma.kind::tf32 x3 -> TMEM -> epilogue) -- against an fp64 product."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run(mode, A, B):
    from iron_b200 import _lib
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    C = torch.full((M, N), float("nan"), device=DEV)
    _lib.check(lib.ironb_gemm_nt(_lib.ptr(A), A.stride(0), _lib.ptr(B), B.stride(0), M, N, K, _lib.ptr(C), N, mode,
                                 _lib.stream()), "gemm_nt")
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 64), (41, 64, 40), (300, 264, 256), (1000, 256, 296),
                                   (4096, 512, 512), (130, 8, 304), (5000, 512, 40)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gemm_nt(M, N, K, mode):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    B = (torch.randn(N, K, generator=g) / np.sqrt(K)).to(DEV)
    ref = A.double() @ B.double().t()
    C = run(mode, A, B)
    assert torch.isfinite(C).all(), f"mode {mode}: non-finite / unwritten outputs"
    scale = (A.double().abs() @ B.double().abs().t())            # sum_k |a||b|: the natural error scale
    rel = ((C.double() - ref).abs() / scale).max().item()
    # fp32 FFMA: ~1e-7; 3xTF32 split with fp32 accumulation: a few 1e-7 (products exact, lo*lo term dropped)
    print(f"mode {mode} M={M} N={N} K={K}: max |err| / sum|a||b| = {rel:.2e}")
    assert rel < (5e-7 if mode == 0 else 2e-6), rel


def split_h(x):
    hi = x.half()
    lo = ((x - hi.float()) * 2048.0).half()
    return hi.contiguous(), lo.contiguous()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (41, 64, 40), (300, 264, 256), (1000, 256, 296), (4096, 512, 512),
                                   (130, 8, 304), (5000, 512, 40), (5000, 40, 512)])
def test_gemm_nt_h16(M, N, K):
    """The pre-split fp16x2 GEMM (forward-type products of get_all / the material nets) against an fp64 product."""
    from iron_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    B = (torch.randn(N, K, generator=g) / np.sqrt(K)).to(DEV)
    Ah, Al = split_h(A)
    Bh, Bl = split_h(B)
    C = torch.full((M, N), float("nan"), device=DEV)
    _lib.check(lib.ironb_gemm_nt_h16(Ah.data_ptr(), Al.data_ptr(), K, Bh.data_ptr(), Bl.data_ptr(), K, M, N, K, _lib.ptr(C), N,
                                     _lib.stream()), "gemm_nt_h16")
    torch.cuda.synchronize()
    assert torch.isfinite(C).all()
    ref = A.double() @ B.double().t()
    scale = (A.double().abs() @ B.double().abs().t())
    rel = ((C.double() - ref).abs() / scale).max().item()
    print(f"h16 M={M} N={N} K={K}: max |err| / sum|a||b| = {rel:.2e}")
    assert rel < 2e-6, rel


def test_gemm_modes_agree_on_structured_input():
    """identity-like B picks single A entries: any swizzle / descriptor mistake shows up as misplaced columns."""
    M, N, K = 256, 128, 64
    A = torch.arange(M * K, dtype=torch.float32).reshape(M, K).to(DEV) / 7.0
    B = torch.zeros(N, K)
    for n in range(N):
        B[n, (n * 5) % K] = 1.0 + n / 64.0
    B = B.to(DEV)
    ref = (A.double() @ B.double().t())
    for mode in (0, 1, 2):
        C = run(mode, A, B)
        err = (C.double() - ref).abs().max().item()
        assert err <= 1e-3 * ref.abs().max().item() * 1e-3, (mode, err)


@pytest.mark.parametrize("M,Nd,Kd", [(128, 128, 128), (256, 128, 128), (4096, 512, 512), (1000, 264, 512), (53, 8, 296), (2048, 512, 40),
                                     (70000, 256, 256)])
@pytest.mark.parametrize("mode", [0, 1])
def test_gemm_tn(M, Nd, Kd, mode):
    """C += A^T B (weight gradients): FFMA tiles, or transpose + split-K tcgen05 GEMM with an atomically adding epilogue."""
    from iron_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + Nd + Kd)
    A = (torch.randn(M, Nd, generator=g) / np.sqrt(M)).to(DEV)
    B = torch.randn(M, Kd, generator=g).to(DEV)
    C0 = torch.randn(Nd, Kd, generator=g).to(DEV)
    C = C0.clone()
    scratch = torch.empty(int(lib.ironb_gemm_tn_scratch_bytes(M, Nd, Kd)), dtype=torch.uint8, device=DEV)
    _lib.check(lib.ironb_gemm_tn(_lib.ptr(A), Nd, _lib.ptr(B), Kd, M, Nd, Kd, _lib.ptr(C), Kd, mode, _lib.ptr(scratch),
                                 _lib.stream()), "gemm_tn")
    torch.cuda.synchronize()
    ref = C0.double() + A.double().t() @ B.double()
    scale = A.double().abs().t() @ B.double().abs() + C0.double().abs()
    rel = ((C.double() - ref).abs() / scale).max().item()
    print(f"tn mode {mode} M={M} Nd={Nd} Kd={Kd}: max |err| / scale = {rel:.2e}")
    assert torch.isfinite(C).all() and rel < 2e-6, rel


def test_gemm_tc_two_streams_bitwise():
    """Regression test of round 1's multi-stream fault (profiles/r2j_multistream_fault_rootcause.md): two streams of 3xTF32 GEMM
    grids next to a copy stream, every output compared BITWISE with a quiet single-stream run of the same (deterministic)
    kernel.  With the old splitter protocol (IRONB_SPLIT_STRICT=0) this shape died with 'unspecified launch failure' within
    ~9,000 launches (3 s); the fixed kernel ran 190 k launches clean.  Here: ~4 s of it."""
    import time
    from iron_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    M, N, K = 16384, 512, 2048
    A = [torch.randn(M, K, generator=g).to(DEV) for _ in range(2)]
    B = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    hog_src = torch.empty(64 << 20, dtype=torch.float32, device=DEV).normal_()
    hog_dst = torch.empty_like(hog_src)
    streams = [torch.cuda.Stream() for _ in range(3)]

    def gemm(a, c, st):
        _lib.check(lib.ironb_gemm_nt(_lib.ptr(a), K, _lib.ptr(B), K, M, N, K, _lib.ptr(c), N, 2, st.cuda_stream), "gemm_nt")

    ref = [torch.empty(M, N, device=DEV) for _ in range(2)]
    for i in range(2):
        gemm(A[i], ref[i], torch.cuda.current_stream())
    torch.cuda.synchronize()
    out = [torch.empty(M, N, device=DEV) for _ in range(2)]
    bad = torch.zeros(2, dtype=torch.int64, device=DEV)
    launches, t0 = 0, time.time()
    while time.time() - t0 < 4.0:
        for _ in range(10):
            with torch.cuda.stream(streams[2]):
                hog_dst.copy_(hog_src, non_blocking=True)
            for i in range(2):
                with torch.cuda.stream(streams[i]):
                    for _ in range(4):
                        gemm(A[i], out[i], streams[i])
                        bad[i] += (out[i] != ref[i]).any().to(torch.int64)
            launches += 8
        torch.cuda.synchronize()
    assert bad.tolist() == [0, 0], f"{bad.tolist()} mismatching launches of {launches}"
    print(f"two-stream 3xTF32 GEMM: {launches} launches bit-identical to the quiet run")
