// tcgen05 tile GEMM on PRE-SPLIT fp16x2 operands:  C[m][n] = epi( sum_k A[m][k] * B[n][k] ),  fp32-grade accuracy.
//
// Every fp32 operand x lives in memory as hi = fp16_rn(x) and lo = fp16_rn((x - hi) * 2^11) (two K-major half arrays), written
// by whoever produced it: the weight-norm fold for weights, the producing GEMM's epilogue for activations.  The kernel then
// has nothing to convert: TMA (SWIZZLE_128B, 64-half boxes) -> three tcgen05.mma.kind::f16 per 16-wide k-step (hi*hi into one
// TMEM accumulator; lo*hi + hi*lo, scaled by 2^11, into a second) -> the epilogue folds the two with one FMA, applies the
// truncation de-bias of the hi*hi chain (mlp_h16.cu: mlp16_debias) and runs the same transposed, coalesced epilogue functors as
// the 3xTF32 kernel (gemm_tc.cuh).
//
// Why (round 2): the 3xTF32 kernel splits fp32 tiles inside shared memory and sits at the shared-memory bandwidth of that
// split (18k cycles of mainloop for 4096 x 512 x 512, 39 % of its shared-memory wavefronts bank conflicts,
// profiles/r1t_gemm_nt_tc_ncu.md).  Pre-split fp16 operands are the same 4 bytes per element, need no conversion, and
// kind::f16 issues at twice the tf32 rate: the mainloop drops to the SM's L2 ingest time (512 KiB per CTA at ~64 B/clk = 8k
// cycles).  Used for the forward-type products (get_all forward, the input-gradient chain, material-net forward), whose
// operands are O(1); gradient operands keep the 3xTF32 kernel (they need fp32's exponent range).
#pragma once
#include <cuda_fp16.h>

#include "gemm_tc.cuh"

namespace ironb {
float mlp16_debias();   // mlp_h16.cu

namespace h16 {
using namespace tc;

constexpr int HBK = 64;                              // K elements (halfs) per stage: one 128-byte swizzled row
constexpr int HTILE = 128 * 128;                     // bytes of one operand box (128 rows x 128 B)
constexpr int HSTAGE = 4 * HTILE;                    // [A_hi][B_hi][A_lo][B_lo]
constexpr int HSTAGES = 3;
constexpr int HSMEM = HSTAGES * HSTAGE + 1024 + 256;
constexpr int HTHREADS = 320;                        // warp 0 TMA, warp 1 MMA (+ TMEM alloc), warps 2-9 epilogue
constexpr uint32_t HTMEM_COLS = 256;                 // hi*hi and cross-term accumulators, 128 columns each
constexpr float H_LO_INV = 1.f / 2048.f;
// kind::f16: fp16 A and B, fp32 accumulate, both K-major, M = 128, N = 128
constexpr uint32_t HIDESC = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void h_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(HIDESC), "r"(acc)
      : "memory");
}

// hi / lo halfs of four consecutive fp32 values -> two 8-byte stores (element offset `o` is a multiple of 4)
__device__ __forceinline__ void store_split4(__half* __restrict__ hi, __half* __restrict__ lo, int64_t o, const float (&v)[4]) {
  __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  __half2 l01 = __floats2half2_rn((v[0] - f01.x) * 2048.f, (v[1] - f01.y) * 2048.f);
  __half2 l23 = __floats2half2_rn((v[2] - f23.x) * 2048.f, (v[3] - f23.y) * 2048.f);
  uint2 ph, pl;
  ph.x = *reinterpret_cast<uint32_t*>(&h01); ph.y = *reinterpret_cast<uint32_t*>(&h23);
  pl.x = *reinterpret_cast<uint32_t*>(&l01); pl.y = *reinterpret_cast<uint32_t*>(&l23);
  *reinterpret_cast<uint2*>(hi + o) = ph;
  *reinterpret_cast<uint2*>(lo + o) = pl;
}

template <class Epi>
__global__ void __launch_bounds__(HTHREADS, 1)
gemm_h16_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, int M, int N, int K, Epi epi,
                float gam) {
  if ((int)blockIdx.y * BM >= M) return;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + HSTAGES * HSTAGE;                      // full[3] empty[3] acc[1] | tmem ptr
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (HSTAGES + s); };
  const uint32_t acc_bar = bars + 8u * (2 * HSTAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * HSTAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + HSTAGES * HSTAGE + 8 * (2 * HSTAGES + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nk = (K + HBK - 1) / HBK;
  if (nk <= 0) return;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
    for (int s = 0; s < HSTAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(HTMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    const bool leader = elect_one();
    for (int it = 0; it < nk; ++it) {
      const int s = it % HSTAGES;
      const uint32_t ph = (it / HSTAGES) & 1;
      mbar_wait(empty(s), ph ^ 1, 21, it);
      const uint32_t st = base + s * HSTAGE;
      if (leader) {
        mbar_arrive_expect_tx(full(s), HSTAGE);
        tma_load_2d(st, &mapAh, it * HBK, m0, full(s));
        tma_load_2d(st + HTILE, &mapBh, it * HBK, n0, full(s));
        tma_load_2d(st + 2 * HTILE, &mapAl, it * HBK, m0, full(s));
        tma_load_2d(st + 3 * HTILE, &mapBl, it * HBK, n0, full(s));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const bool leader = elect_one();
    for (int it = 0; it < nk; ++it) {
      const int s = it % HSTAGES;
      const uint32_t ph = (it / HSTAGES) & 1;
      mbar_wait(full(s), ph, 22, it);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t st = base + s * HSTAGE;
      const uint64_t a_hi = make_desc(st), b_hi = make_desc(st + HTILE);
      const uint64_t a_lo = make_desc(st + 2 * HTILE), b_lo = make_desc(st + 3 * HTILE);
#pragma unroll
      for (int kk = 0; kk < HBK / 16; ++kk) {
        const uint64_t adv = (uint64_t)(kk * 2);          // 16 halfs = 32 B = 2 x 16 B along the swizzled row
        const int g = it * (HBK / 16) + kk;
        if (leader) {
          h_mma_f16(tmem, a_hi + adv, b_hi + adv, g >= 1 ? 1u : 0u);
          h_mma_f16(tmem + 128u, a_lo + adv, b_hi + adv, g >= 1 ? 1u : 0u);
          h_mma_f16(tmem + 128u, a_hi + adv, b_lo + adv, 1u);
        }
      }
      if (leader) tc_commit(empty(s));
    }
    if (leader) tc_commit(acc_bar);
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..9) =================
    mbar_wait(acc_bar, 0, 24, nk);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr int TLD = BN + 4;
    float* tile = reinterpret_cast<float*>(base_ptr);
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = half * 64 + cc * 32;
      uint32_t r[32], r1[32];
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      tmem_ld32(taddr, r);
      tmem_ld32(taddr + 128u, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float hh = __uint_as_float(r[g * 4 + j]);
          v[j] = fmaf(__uint_as_float(r1[g * 4 + j]), H_LO_INV, fmaf(hh, gam, hh));
        }
        *reinterpret_cast<float4*>(tile + row * TLD + c0 + g * 4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int w8 = warp - 2;
#pragma unroll 2
    for (int rr = 0; rr < BM / 8; ++rr) {
      const int trow = w8 * (BM / 8) + rr;
      const int m = m0 + trow;
      const int n = n0 + lane * 4;
      if (m < M && n < N) {
        const float4 v4 = *reinterpret_cast<const float4*>(tile + trow * TLD + lane * 4);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        epi(m, n, v);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(HTMEM_COLS) : "memory");
  }
}

// A: M x K as two half arrays (hi, lo) with row pitch lda halfs; B: N x K likewise (ldb).  Pitches multiples of 8 halfs.
template <class Epi>
int launch_gemm_h16(const __half* Ah, const __half* Al, int lda, const __half* Bh, const __half* Bl, int ldb, int M, int N, int K,
                    const Epi& epi, cudaStream_t st, const char* what) {
  if (M <= 0 || N <= 0) return IRONB_OK;
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if ((rc = make_map_h(&mAh, Ah, M, K, lda))) return rc;
  if ((rc = make_map_h(&mAl, Al, M, K, lda))) return rc;
  if ((rc = make_map_h(&mBh, Bh, N, K, ldb))) return rc;
  if ((rc = make_map_h(&mBl, Bl, N, K, ldb))) return rc;
  auto kern = gemm_h16_kernel<Epi>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HSMEM);
    if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  dim3 grid((unsigned)ceil_div64(N, BN), (unsigned)ceil_div64(M, BM), 1);
  const float gam = mlp16_debias() * (float)((K + 15) / 16) * 5.9604645e-8f;     // g * n_steps * 2^-24
  kern<<<grid, HTHREADS, HSMEM, st>>>(mAh, mAl, mBh, mBl, M, N, K, epi, gam);
  IRONB_CHECK_LAUNCH(what);
  return IRONB_OK;
}

}  // namespace h16

// 2 = fp16x2 pre-split operands for the forward-type products + 3xTF32 for gradient operands (default); 1 = 3xTF32
// everywhere; 0 = fp32 FFMA tiles.  (ironb_set_gemm_mode / IRONB_GEMM)
int gemm_mode();
}  // namespace ironb
