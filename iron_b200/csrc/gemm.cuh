// fp32 SIMT tile GEMMs with fused epilogues -- the exact-arithmetic (FFMA, fp32 accumulate) workhorse
// of the differentiable part of the path (get_all forward / double backward, material MLPs).
//
//   gemm_nt :  C[m][n] = epi( sum_k A[m][k] * B[n][k] )          A: [M][lda], B: [N][ldb], both K-major
//   gemm_tn :  C[n][k] += sum_m A[m][n] * B[m][k]                weight gradients, split over m, fp32 atomics
//   colsum  :  out[n]  += scale * sum_m A[m][n]                  bias gradients
//
// Every operand is padded by the packed layout (leading dimensions and K multiples of 8, zero filled),
// so all global accesses are 128-bit; M is arbitrary.  Thread block = 256 threads as 16 x 16, each
// thread owns (BM/16) x (BN/16) accumulators in groups of 4 contiguous rows / columns, operands are
// staged K-major in shared memory with register prefetch of the next K-slab.
#pragma once
#include "common.cuh"

namespace ironb {

constexpr int GEMM_BK = 16;

template <int BM, int BN, class Epi>
__global__ void __launch_bounds__(256) gemm_nt_kernel(const float* __restrict__ A, int lda,
                                                      const float* __restrict__ B, int ldb, int M, int N, int K,
                                                      Epi epi) {
  static_assert(BM % 64 == 0 && BN % 64 == 0, "tile must be a multiple of 64");
  constexpr int GM = BM / 64, GN = BN / 64;          // groups of 4 rows / cols per thread
  constexpr int LA = BM / 64, LB = BN / 64;          // float4 loads per thread per slab
  __shared__ float As[2][GEMM_BK][BM + 4];
  __shared__ float Bs[2][GEMM_BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;     // loader: 64 rows x 4 float4 per pass

  float acc[GM * 4][GN * 4];
#pragma unroll
  for (int i = 0; i < GM * 4; ++i)
#pragma unroll
    for (int j = 0; j < GN * 4; ++j) acc[i][j] = 0.f;

  float4 ra[LA], rb[LB];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int m = m0 + lrow + i * 64, k = k0 + lk;
      ra[i] = (m < M && k < K) ? __ldg(reinterpret_cast<const float4*>(A + (int64_t)m * lda + k))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int n = n0 + lrow + i * 64, k = k0 + lk;
      rb[i] = (n < N && k < K) ? __ldg(reinterpret_cast<const float4*>(B + (int64_t)n * ldb + k))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int r = lrow + i * 64;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int r = lrow + i * 64;
      Bs[buf][lk + 0][r] = rb[i].x; Bs[buf][lk + 1][r] = rb[i].y; Bs[buf][lk + 2][r] = rb[i].z; Bs[buf][lk + 3][r] = rb[i].w;
    }
  };

  const int nslab = (K + GEMM_BK - 1) / GEMM_BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int s = 0; s < nslab; ++s) {
    const int buf = s & 1;
    if (s + 1 < nslab) gload((s + 1) * GEMM_BK);
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; ++kk) {
      float a[GM * 4], b[GN * 4];
#pragma unroll
      for (int g = 0; g < GM; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][g * 64 + ty * 4]);
        a[g * 4] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < GN; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * 64 + tx * 4]);
        b[g * 4] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < GM * 4; ++i)
#pragma unroll
        for (int j = 0; j < GN * 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (s + 1 < nslab) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int gi = 0; gi < GM; ++gi)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m = m0 + gi * 64 + ty * 4 + i;
      if (m >= M) continue;
#pragma unroll
      for (int gj = 0; gj < GN; ++gj) {
        int n = n0 + gj * 64 + tx * 4;
        if (n >= N) continue;
        float v[4] = {acc[gi * 4 + i][gj * 4], acc[gi * 4 + i][gj * 4 + 1], acc[gi * 4 + i][gj * 4 + 2],
                      acc[gi * 4 + i][gj * 4 + 3]};
        epi(m, n, v);
      }
    }
}

// N and K must be multiples of 4 (they are padded dims); lda/ldb multiples of 4; A/B 16-byte aligned.
template <class Epi>
int launch_gemm_nt(const float* A, int lda, const float* B, int ldb, int M, int N, int K, const Epi& epi,
                   cudaStream_t st, const char* what) {
  if (M <= 0 || N <= 0) return IRONB_OK;
  if ((lda & 3) || (ldb & 3) || (N & 3) || (K & 3)) {
    set_error("%s: gemm_nt needs padded dims (lda=%d ldb=%d N=%d K=%d)", what, lda, ldb, N, K);
    return IRONB_EINVAL;
  }
  int64_t big = ceil_div64(M, 128) * ceil_div64(N, 128);
  if (big >= 2 * (int64_t)num_sms()) {
    dim3 grid((unsigned)ceil_div64(N, 128), (unsigned)ceil_div64(M, 128));
    gemm_nt_kernel<128, 128, Epi><<<grid, 256, 0, st>>>(A, lda, B, ldb, M, N, K, epi);
  } else {
    dim3 grid((unsigned)ceil_div64(N, 64), (unsigned)ceil_div64(M, 64));
    gemm_nt_kernel<64, 64, Epi><<<grid, 256, 0, st>>>(A, lda, B, ldb, M, N, K, epi);
  }
  IRONB_CHECK_LAUNCH(what);
  return IRONB_OK;
}

// ---- weight gradient:  C[n][k] += sum_m A[m][n] * B[m][k] ------------------------------------------
// grid = (ceil(Kd/BN2), ceil(Nd/BN1), splits); each z-slice reduces rows [z*chunk, (z+1)*chunk).
template <int BT>
__global__ void __launch_bounds__(256) gemm_tn_kernel(const float* __restrict__ A, int lda,
                                                      const float* __restrict__ B, int ldb, int M, int Nd, int Kd,
                                                      int chunk, float* __restrict__ C, int ldc) {
  constexpr int G = BT / 64;
  constexpr int PER = BT / 64;   // float4 loads per thread per operand per slab (16 x BT floats)
  __shared__ float As[2][GEMM_BK][BT + 4];
  __shared__ float Bs[2][GEMM_BK][BT + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.y * BT, k0 = blockIdx.x * BT;
  const int mbeg = blockIdx.z * chunk;
  const int mend = min(M, mbeg + chunk);
  if (mbeg >= mend) return;
  // loader mapping: BT/4 float4 per row, 16 rows
  constexpr int F4_PER_ROW = BT / 4;

  float acc[G * 4][G * 4];
#pragma unroll
  for (int i = 0; i < G * 4; ++i)
#pragma unroll
    for (int j = 0; j < G * 4; ++j) acc[i][j] = 0.f;

  float4 ra[PER], rb[PER];
  auto gload = [&](int mb) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int f = tid + i * 256;
      int d = f / F4_PER_ROW, c = (f % F4_PER_ROW) * 4;
      int m = mb + d;
      ra[i] = (m < mend && n0 + c < Nd) ? __ldg(reinterpret_cast<const float4*>(A + (int64_t)m * lda + n0 + c))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = (m < mend && k0 + c < Kd) ? __ldg(reinterpret_cast<const float4*>(B + (int64_t)m * ldb + k0 + c))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int f = tid + i * 256;
      int d = f / F4_PER_ROW, c = (f % F4_PER_ROW) * 4;
      *reinterpret_cast<float4*>(&As[buf][d][c]) = ra[i];
      *reinterpret_cast<float4*>(&Bs[buf][d][c]) = rb[i];
    }
  };

  const int nslab = (mend - mbeg + GEMM_BK - 1) / GEMM_BK;
  gload(mbeg);
  sstore(0);
  __syncthreads();
  for (int s = 0; s < nslab; ++s) {
    const int buf = s & 1;
    if (s + 1 < nslab) gload(mbeg + (s + 1) * GEMM_BK);
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; ++kk) {
      float a[G * 4], b[G * 4];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][g * 64 + ty * 4]);
        a[g * 4] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
        float4 w = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * 64 + tx * 4]);
        b[g * 4] = w.x; b[g * 4 + 1] = w.y; b[g * 4 + 2] = w.z; b[g * 4 + 3] = w.w;
      }
#pragma unroll
      for (int i = 0; i < G * 4; ++i)
#pragma unroll
        for (int j = 0; j < G * 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (s + 1 < nslab) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int gi = 0; gi < G; ++gi)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n = n0 + gi * 64 + ty * 4 + i;
      if (n >= Nd) continue;
#pragma unroll
      for (int gj = 0; gj < G; ++gj)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int k = k0 + gj * 64 + tx * 4 + j;
          if (k < Kd) atomicAdd(C + (int64_t)n * ldc + k, acc[gi * 4 + i][gj * 4 + j]);
        }
    }
}

inline int launch_gemm_tn(const float* A, int lda, const float* B, int ldb, int M, int Nd, int Kd, float* C,
                          int ldc, cudaStream_t st, const char* what) {
  if (M <= 0 || Nd <= 0 || Kd <= 0) return IRONB_OK;
  if ((lda & 3) || (ldb & 3) || (Nd & 3) || (Kd & 3)) {
    set_error("%s: gemm_tn needs padded dims", what);
    return IRONB_EINVAL;
  }
  int64_t tiles128 = ceil_div64(Nd, 128) * ceil_div64(Kd, 128);
  const int sms = num_sms();
  if (tiles128 >= 8) {
    int64_t splits = (2 * sms + tiles128 - 1) / tiles128;
    int64_t maxs = ceil_div64(M, 64);
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    int chunk = (int)(ceil_div64(ceil_div64(M, splits), GEMM_BK) * GEMM_BK);
    splits = ceil_div64(M, chunk);
    dim3 grid((unsigned)ceil_div64(Kd, 128), (unsigned)ceil_div64(Nd, 128), (unsigned)splits);
    gemm_tn_kernel<128><<<grid, 256, 0, st>>>(A, lda, B, ldb, M, Nd, Kd, chunk, C, ldc);
  } else {
    int64_t tiles = ceil_div64(Nd, 64) * ceil_div64(Kd, 64);
    int64_t splits = (2 * sms + tiles - 1) / tiles;
    int64_t maxs = ceil_div64(M, 64);
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    int chunk = (int)(ceil_div64(ceil_div64(M, splits), GEMM_BK) * GEMM_BK);
    splits = ceil_div64(M, chunk);
    dim3 grid((unsigned)ceil_div64(Kd, 64), (unsigned)ceil_div64(Nd, 64), (unsigned)splits);
    gemm_tn_kernel<64><<<grid, 256, 0, st>>>(A, lda, B, ldb, M, Nd, Kd, chunk, C, ldc);
  }
  IRONB_CHECK_LAUNCH(what);
  return IRONB_OK;
}

// ---- column sums: out[n] += scale * sum_m A[m][n],  n < ncols --------------------------------------
static __global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int lda, int M, int ncols,
                                                     int chunk, float scale, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + x;
  const int mbeg = blockIdx.y * chunk, mend = min(M, mbeg + chunk);
  float s = 0.f;
  if (n < ncols)
    for (int m = mbeg + y; m < mend; m += 8) s += A[(int64_t)m * lda + n];
  red[y][x] = s;
  __syncthreads();
  if (y == 0 && n < ncols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][x];
    atomicAdd(out + n, t * scale);
  }
}

inline int launch_colsum(const float* A, int lda, int M, int ncols, float scale, float* out, cudaStream_t st,
                         const char* what) {
  if (M <= 0 || ncols <= 0) return IRONB_OK;
  int bx = (ncols + 31) / 32;
  int64_t splits = (2 * num_sms() + bx - 1) / bx;
  int64_t maxs = ceil_div64(M, 64);
  if (splits > maxs) splits = maxs;
  if (splits < 1) splits = 1;
  int chunk = (int)ceil_div64(M, splits);
  splits = ceil_div64(M, chunk);
  dim3 grid(bx, (unsigned)splits);
  colsum_kernel<<<grid, 256, 0, st>>>(A, lda, M, ncols, chunk, scale, out);
  IRONB_CHECK_LAUNCH(what);
  return IRONB_OK;
}

}  // namespace ironb
