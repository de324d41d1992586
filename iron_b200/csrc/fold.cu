// Weight-norm fold / unfold.  W = v * (g / ||v||_row) is what nn.utils.weight_norm(dim=0) feeds every
// nn.Linear of the reference (models/fields.py:75-76, 194-195).  The fold writes the effective weight
// both as W [out_pad][in_pad] and transposed W^T [in_pad][out_pad] into the packed buffer (pads stay
// zero), so every GEMM of the path is a K-major x K-major ("NT") product.  The unfold is the chain rule
//   dg = sum_k dW * v/||v||,   dv = (g/||v||) * (dW - dg * v/||v||).
#include <cuda_fp16.h>

#include "common.cuh"

namespace ironb {
namespace {

struct FoldArgs {
  const float* v[IRONB_MAX_LIN];
  const float* g[IRONB_MAX_LIN];
  const float* b[IRONB_MAX_LIN];
  float* dv[IRONB_MAX_LIN];
  float* dg[IRONB_MAX_LIN];
  float* db[IRONB_MAX_LIN];
  int row_start[IRONB_MAX_LIN + 1];   // prefix sum of out_dim: global row id -> layer
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// one warp per output row of one layer
__global__ void __launch_bounds__(256) fold_kernel(ironb_mlp_layout L, FoldArgs A, float* __restrict__ packed) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= A.row_start[L.n_lin]) return;
  int l = 0;
  while (row >= A.row_start[l + 1]) ++l;
  int n = row - A.row_start[l];
  int K = L.in_dim[l], Kp = L.in_pad[l], Np = L.out_pad[l];
  const float* v = A.v[l] + (int64_t)n * K;
  float sc = 1.f;
  if (A.g[l] != nullptr) {
    float ss = 0.f;
    for (int k = lane; k < K; k += 32) { float x = v[k]; ss += x * x; }
    ss = warp_sum(ss);
    sc = __fdiv_rn(A.g[l][n], sqrtf(ss));
  }
  float* W = packed + L.off_w[l] + (int64_t)n * Kp;
  float* WT = packed + L.off_wt[l] + n;
  // fp16x2-split copies for the tensor cores: hi = fp16(w), lo = fp16((w - hi) * 2^11); the lo block follows the hi block
  __half* Whi = L.off_h16[l] > 0 ? reinterpret_cast<__half*>(packed + L.off_h16[l]) + (int64_t)n * Kp : nullptr;
  __half* Wlo = Whi ? Whi + (int64_t)Np * Kp : nullptr;
  for (int k = lane; k < K; k += 32) {
    float w = v[k] * sc;
    W[k] = w;
    WT[(int64_t)k * Np] = w;
    const __half h = __float2half_rn(w);
    const __half lo = __float2half_rn((w - __half2float(h)) * 2048.f);
    if (Whi) { Whi[k] = h; Wlo[k] = lo; }
  }
  if (lane == 0) packed[L.off_b[l] + n] = A.b[l] ? A.b[l][n] : 0.f;
}

// fp16x2-split copy of W_l^T for every layer of an SDF net: consecutive threads write consecutive halfs of a W^T row (the fold
// kernel's warp-per-row layout would scatter 2-byte stores with a pitch of out_pad halfs); reads the fp32 W^T written above.
__global__ void __launch_bounds__(256) fold_wt_h16_kernel(ironb_mlp_layout L, float* __restrict__ packed) {
  const int l = blockIdx.y;
  if (L.off_h16t[l] <= 0) return;
  const int64_t n_el = (int64_t)L.in_pad[l] * L.out_pad[l];
  const float* __restrict__ WT = packed + L.off_wt[l];
  __half* __restrict__ hi = reinterpret_cast<__half*>(packed + L.off_h16t[l]);
  __half* __restrict__ lo = hi + n_el;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 2; i < n_el; i += (int64_t)gridDim.x * blockDim.x * 2) {
    const float2 w = *reinterpret_cast<const float2*>(WT + i);
    const __half2 h = __floats2half2_rn(w.x, w.y);
    const float2 hf = __half22float2(h);
    *reinterpret_cast<__half2*>(hi + i) = h;
    *reinterpret_cast<__half2*>(lo + i) = __floats2half2_rn((w.x - hf.x) * 2048.f, (w.y - hf.y) * 2048.f);
  }
}

__global__ void __launch_bounds__(256) fold_bwd_kernel(ironb_mlp_layout L, FoldArgs A, const float* __restrict__ dpacked) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= A.row_start[L.n_lin]) return;
  int l = 0;
  while (row >= A.row_start[l + 1]) ++l;
  int n = row - A.row_start[l];
  int K = L.in_dim[l], Kp = L.in_pad[l];
  const float* v = A.v[l] + (int64_t)n * K;
  const float* dW = dpacked + L.off_w[l] + (int64_t)n * Kp;
  float* dv = A.dv[l] + (int64_t)n * K;
  if (A.g[l] != nullptr) {
    float ss = 0.f, dot = 0.f;
    for (int k = lane; k < K; k += 32) { float x = v[k]; ss += x * x; dot += dW[k] * x; }
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    float nv = sqrtf(ss);
    float dg = dot / nv;                       // sum dW * v/||v||
    float gs = A.g[l][n] / nv;
    for (int k = lane; k < K; k += 32) dv[k] = gs * (dW[k] - dg * (v[k] / nv));
    if (lane == 0 && A.dg[l] != nullptr) A.dg[l][n] = dg;
  } else {
    for (int k = lane; k < K; k += 32) dv[k] = dW[k];
  }
  if (lane == 0 && A.db[l] != nullptr) A.db[l][n] = dpacked[L.off_b[l] + n];
}

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int ironb_mlp_fold(const ironb_mlp_layout* lay, const float* const* v, const float* const* g,
                              const float* const* b, float* packed, void* stream) {
  IRONB_REQUIRE(lay && v && packed, "mlp_fold: null argument");
  FoldArgs A;
  memset(&A, 0, sizeof(A));
  int rows = 0;
  for (int l = 0; l < lay->n_lin; ++l) {
    IRONB_REQUIRE(v[l] != nullptr, "mlp_fold: null weight_v for layer %d", l);
    A.v[l] = v[l];
    A.g[l] = g ? g[l] : nullptr;
    A.b[l] = b ? b[l] : nullptr;
    A.row_start[l] = rows;
    rows += lay->out_dim[l];
  }
  A.row_start[lay->n_lin] = rows;
  int blocks = (rows + 7) / 8;
  fold_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*lay, A, packed);
  IRONB_CHECK_LAUNCH("fold_kernel");
  if (lay->kind == 0) {
    fold_wt_h16_kernel<<<dim3(64, (unsigned)lay->n_lin), 256, 0, as_stream(stream)>>>(*lay, packed);
    IRONB_CHECK_LAUNCH("fold_wt_h16_kernel");
  }
  return IRONB_OK;
}

extern "C" int ironb_mlp_fold_bwd(const ironb_mlp_layout* lay, const float* const* v, const float* const* g,
                                  const float* dpacked, float* const* dv, float* const* dg, float* const* db,
                                  void* stream) {
  IRONB_REQUIRE(lay && v && dpacked && dv, "mlp_fold_bwd: null argument");
  FoldArgs A;
  memset(&A, 0, sizeof(A));
  int rows = 0;
  for (int l = 0; l < lay->n_lin; ++l) {
    IRONB_REQUIRE(v[l] != nullptr && dv[l] != nullptr, "mlp_fold_bwd: null pointer for layer %d", l);
    A.v[l] = v[l];
    A.g[l] = g ? g[l] : nullptr;
    A.dv[l] = dv[l];
    A.dg[l] = dg ? dg[l] : nullptr;
    A.db[l] = db ? db[l] : nullptr;
    A.row_start[l] = rows;
    rows += lay->out_dim[l];
  }
  A.row_start[lay->n_lin] = rows;
  int blocks = (rows + 7) / 8;
  fold_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*lay, A, dpacked);
  IRONB_CHECK_LAUNCH("fold_bwd_kernel");
  return IRONB_OK;
}
