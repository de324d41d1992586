// Small fused elementwise / reduction kernels for the glue between the big kernels of a stage-2 step.  In the reference these
// are chains of ATen element-wise ops (normalise, norm, clamp, abs, mean, masked sums ...); one step issued ~180 such launches,
// ~3 us each and mostly on the step's critical path.  Each function here is one launch forward and one backward.
//
//   ironb_unit_dist_*      n = g / (|g| + 1e-10),  dist = |x - o|                       render_surface.py:135-146
//   ironb_reparam_bwd      d f = -sum_c v_c / clamp(g.v, 1e-4) * d x_c                  models/raytracer.py:17-24 (forward == x)
//   ironb_matpost_*        kd = |a|;  ks = mean_c |b_c| (or |b| if is_metal);  alpha = |c| + 0.01    models/rendering_func.py:5-16
//   ironb_eik_sum_*        sum_m w_m (|g_m| - 1)^2                                       render_surface.py:580-583, 601-603
//   ironb_roughrange_*     mean over {m: w_m and r_m > 0.5} of (r_m - 0.5)              render_surface.py:609-613
//   ironb_mask_rows_*      y_m = x_m * w_m for up to 8 row tensors of widths 1 / 3      (dense shading: zero the non-hit pixels)
#include "common.cuh"

namespace ironb {
namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) unit_dist_fwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                            const float* __restrict__ o, int64_t M, float* __restrict__ n,
                                                            float* __restrict__ dist) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float gx = g[m * 3], gy = g[m * 3 + 1], gz = g[m * 3 + 2];
  const float nr = sqrtf(gx * gx + gy * gy + gz * gz) + 1e-10f;
  n[m * 3] = gx / nr; n[m * 3 + 1] = gy / nr; n[m * 3 + 2] = gz / nr;
  const float dx = x[m * 3] - o[m * 3], dy = x[m * 3 + 1] - o[m * 3 + 1], dz = x[m * 3 + 2] - o[m * 3 + 2];
  dist[m] = sqrtf(dx * dx + dy * dy + dz * dz);
}

// d g = (d n - n (n . d n) * |g| / (|g| + eps)) / (|g| + eps)   (autograd of g / (|g| + eps); d|g|/dg = g/|g|, 0 at g = 0)
// d x = d dist * (x - o) / |x - o|   (0 at x == o, like torch.norm's backward)
__global__ void __launch_bounds__(256) unit_dist_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                            const float* __restrict__ o, const float* __restrict__ dn,
                                                            const float* __restrict__ ddist, int64_t M, float* __restrict__ dg,
                                                            float* __restrict__ dx) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  if (dg != nullptr) {
    const float gv[3] = {g[m * 3], g[m * 3 + 1], g[m * 3 + 2]};
    const float nrm = sqrtf(gv[0] * gv[0] + gv[1] * gv[1] + gv[2] * gv[2]);
    const float den = nrm + 1e-10f;
    float up[3] = {0.f, 0.f, 0.f};
    if (dn != nullptr) { up[0] = dn[m * 3]; up[1] = dn[m * 3 + 1]; up[2] = dn[m * 3 + 2]; }
    // n_c = g_c / den;  d n_c / d g_k = delta_ck / den - g_c (g_k / nrm) / den^2
    const float dotg = gv[0] * up[0] + gv[1] * up[1] + gv[2] * up[2];
    const float s = nrm > 0.f ? dotg / (nrm * den * den) : 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) dg[m * 3 + k] = up[k] / den - gv[k] * s;
  }
  if (dx != nullptr) {
    const float d[3] = {x[m * 3] - o[m * 3], x[m * 3 + 1] - o[m * 3 + 1], x[m * 3 + 2] - o[m * 3 + 2]};
    const float nr = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float u = (ddist != nullptr && nr > 0.f) ? ddist[m] / nr : 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) dx[m * 3 + k] = d[k] * u;
  }
}

__global__ void __launch_bounds__(256) reparam_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                                          const float* __restrict__ dx, int64_t M, float* __restrict__ df) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float vx = v[m * 3], vy = v[m * 3 + 1], vz = v[m * 3 + 2];
  const float dot = fmaxf(g[m * 3] * vx + g[m * 3 + 1] * vy + g[m * 3 + 2] * vz, 1e-4f);
  df[m] = -((vx / dot) * dx[m * 3] + (vy / dot) * dx[m * 3 + 1] + (vz / dot) * dx[m * 3 + 2]);
}

__global__ void __launch_bounds__(256) matpost_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          const float* __restrict__ c, int64_t M, int is_metal,
                                                          float* __restrict__ kd, float* __restrict__ ks, float* __restrict__ al) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) kd[m * 3 + k] = fabsf(a[m * 3 + k]);
  const float b0 = fabsf(b[m * 3]), b1 = fabsf(b[m * 3 + 1]), b2 = fabsf(b[m * 3 + 2]);
  if (is_metal) { ks[m * 3] = b0; ks[m * 3 + 1] = b1; ks[m * 3 + 2] = b2; }
  else { const float mean = (b0 + b1 + b2) / 3.f; ks[m * 3] = mean; ks[m * 3 + 1] = mean; ks[m * 3 + 2] = mean; }   // torch.mean: sum / 3
  al[m] = fabsf(c[m]) + 0.01f;
}

__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }   // torch.abs backward: sign(x)

__global__ void __launch_bounds__(256) matpost_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          const float* __restrict__ c, const float* __restrict__ dkd,
                                                          const float* __restrict__ dks, const float* __restrict__ dal, int64_t M,
                                                          int is_metal, float* __restrict__ da, float* __restrict__ db,
                                                          float* __restrict__ dc) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) da[m * 3 + k] = dkd ? dkd[m * 3 + k] * sgn(a[m * 3 + k]) : 0.f;
  float u[3] = {0.f, 0.f, 0.f};
  if (dks) { u[0] = dks[m * 3]; u[1] = dks[m * 3 + 1]; u[2] = dks[m * 3 + 2]; }
  if (!is_metal) { const float s = (u[0] + u[1] + u[2]) / 3.f; u[0] = u[1] = u[2] = s; }
#pragma unroll
  for (int k = 0; k < 3; ++k) db[m * 3 + k] = u[k] * sgn(b[m * 3 + k]);
  dc[m] = dal ? dal[m] * sgn(c[m]) : 0.f;
}

// out[0] += sum_m w_m (|g_m| - 1)^2     (w == NULL: all ones)
__global__ void __launch_bounds__(256) eik_sum_fwd_kernel(const float* __restrict__ g, const float* __restrict__ w, int64_t M,
                                                          float* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const float gx = g[m * 3], gy = g[m * 3 + 1], gz = g[m * 3 + 2];
    const float e = sqrtf(gx * gx + gy * gy + gz * gz) - 1.f;
    acc += (w ? w[m] : 1.f) * e * e;
  }
  acc = warp_sum_f(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0 && t != 0.f) atomicAdd(out, t);
  }
}

// d g_m = up * w_m * 2 (|g| - 1) g / |g|     (0 at g = 0)
__global__ void __launch_bounds__(256) eik_sum_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                          const float* __restrict__ up, int64_t M, float* __restrict__ dg) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float gx = g[m * 3], gy = g[m * 3 + 1], gz = g[m * 3 + 2];
  const float nr = sqrtf(gx * gx + gy * gy + gz * gz);
  const float s = nr > 0.f ? (*up) * (w ? w[m] : 1.f) * 2.f * (nr - 1.f) / nr : 0.f;
  dg[m * 3] = gx * s; dg[m * 3 + 1] = gy * s; dg[m * 3 + 2] = gz * s;
}

// acc[0] += sum over selected of (r - 0.5), acc[1] += count;  selected = w_m != 0 and r_m > value
__global__ void __launch_bounds__(256) roughrange_fwd_kernel(const float* __restrict__ r, const float* __restrict__ w, int64_t M,
                                                             float value, float* __restrict__ acc) {
  __shared__ float red[2][8];
  float s = 0.f, c = 0.f;
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    if (w[m] != 0.f && r[m] > value) { s += r[m] - value; c += 1.f; }
  }
  s = warp_sum_f(s); c = warp_sum_f(c);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x < 8) {
    float a = red[0][threadIdx.x], b = red[1][threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffu, a, o); b += __shfl_xor_sync(0xffu, b, o); }
    if (threadIdx.x == 0 && b != 0.f) { atomicAdd(acc, a); atomicAdd(acc + 1, b); }
  }
}
// loss = weight * acc[0] / acc[1] (0 if none selected);  d r_m = up * weight / count on the selected rows
__global__ void __launch_bounds__(256) roughrange_bwd_kernel(const float* __restrict__ r, const float* __restrict__ w,
                                                             const float* __restrict__ acc, const float* __restrict__ up,
                                                             int64_t M, float value, float weight, float* __restrict__ dr) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float cnt = acc[1];
  dr[m] = (cnt > 0.f && w[m] != 0.f && r[m] > value) ? (*up) * weight / cnt : 0.f;
}
__global__ void roughrange_finish_kernel(const float* __restrict__ acc, float weight, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *loss = acc[1] > 0.f ? weight * acc[0] / acc[1] : 0.f;
}

struct MaskArgs {
  const float* src[8];
  float* dst[8];
  int width[8];
  int n;
};
__global__ void __launch_bounds__(256) mask_rows_kernel(MaskArgs A, const float* __restrict__ w, int64_t M) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float f = w[m];
  for (int t = 0; t < A.n; ++t) {
    const int wd = A.width[t];
    for (int k = 0; k < wd; ++k) A.dst[t][m * wd + k] = A.src[t][m * wd + k] * f;
  }
}

// Gradient bucket of the data-parallel step: tensor t (src[t], numel = off[n + t], src may be null = no gradient) is
// copied to flat[off[t] ..): blockIdx.y = tensor, grid-stride over its elements (float4 when both sides are 16-byte aligned).
__global__ void __launch_bounds__(256) pack_tensors_kernel(const float* const* __restrict__ src, const int64_t* __restrict__ off,
                                                           float* __restrict__ flat, float scale) {
  const int t = blockIdx.y;
  const float* __restrict__ s = src[t];
  const int64_t o = off[t], n = off[gridDim.y + t];
  float* __restrict__ d = flat + o;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (s == nullptr) {
    for (int64_t i = tid; i < n; i += nth) d[i] = 0.f;
    return;
  }
  if ((((uintptr_t)s | (uintptr_t)d) & 15u) == 0) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nth) {
      float4 v = reinterpret_cast<const float4*>(s)[i];
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      reinterpret_cast<float4*>(d)[i] = v;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nth) d[i] = s[i] * scale;
  } else {
    for (int64_t i = tid; i < n; i += nth) d[i] = s[i] * scale;
  }
}

inline unsigned blocks_for(int64_t M) { return (unsigned)ceil_div64(M, 256); }
// reductions: one block up to 4,096 rows (a fixed summation order: the training patch reproduces bit for bit), else atomics
inline unsigned red_blocks(int64_t M) { int64_t b = ceil_div64(M, 4096); return (unsigned)(b < 1 ? 1 : (b > 296 ? 296 : b)); }

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int ironb_unit_dist_fwd(const float* g, const float* x, const float* o, int64_t M, float* n, float* dist, void* stream) {
  if (M <= 0) return IRONB_OK;
  IRONB_REQUIRE(g && x && o && n && dist, "unit_dist_fwd: null pointer");
  unit_dist_fwd_kernel<<<blocks_for(M), 256, 0, as_stream(stream)>>>(g, x, o, M, n, dist);
  IRONB_CHECK_LAUNCH("unit_dist_fwd_kernel");
  return IRONB_OK;
}
extern "C" int ironb_unit_dist_bwd(const float* g, const float* x, const float* o, const float* dn, const float* ddist, int64_t M,
                                   float* dg, float* dx, void* stream) {
  if (M <= 0) return IRONB_OK;
  IRONB_REQUIRE(g && x && o, "unit_dist_bwd: null pointer");
  unit_dist_bwd_kernel<<<blocks_for(M), 256, 0, as_stream(stream)>>>(g, x, o, dn, ddist, M, dg, dx);
  IRONB_CHECK_LAUNCH("unit_dist_bwd_kernel");
  return IRONB_OK;
}
extern "C" int ironb_reparam_bwd(const float* g, const float* v, const float* dx, int64_t M, float* df, void* stream) {
  if (M <= 0) return IRONB_OK;
  IRONB_REQUIRE(g && v && dx && df, "reparam_bwd: null pointer");
  reparam_bwd_kernel<<<blocks_for(M), 256, 0, as_stream(stream)>>>(g, v, dx, M, df);
  IRONB_CHECK_LAUNCH("reparam_bwd_kernel");
  return IRONB_OK;
}
extern "C" int ironb_matpost_fwd(const float* a, const float* b, const float* c, int64_t M, int is_metal, float* kd, float* ks,
                                 float* alpha, void* stream) {
  if (M <= 0) return IRONB_OK;
  IRONB_REQUIRE(a && b && c && kd && ks && alpha, "matpost_fwd: null pointer");
  matpost_fwd_kernel<<<blocks_for(M), 256, 0, as_stream(stream)>>>(a, b, c, M, is_metal, kd, ks, alpha);
  IRONB_CHECK_LAUNCH("matpost_fwd_kernel");
  return IRONB_OK;
}
extern "C" int ironb_matpost_bwd(const float* a, const float* b, const float* c, const float* dkd, const float* dks,
                                 const float* dalpha, int64_t M, int is_metal, float* da, float* db, float* dc, void* stream) {
  if (M <= 0) return IRONB_OK;
  IRONB_REQUIRE(a && b && c && da && db && dc, "matpost_bwd: null pointer");
  matpost_bwd_kernel<<<blocks_for(M), 256, 0, as_stream(stream)>>>(a, b, c, dkd, dks, dalpha, M, is_metal, da, db, dc);
  IRONB_CHECK_LAUNCH("matpost_bwd_kernel");
  return IRONB_OK;
}
/* out[0] += sum_m w_m (|g_m| - 1)^2; the caller zeroes out. */
extern "C" int ironb_eik_sum_fwd(const float* g, const float* w, int64_t M, float* out, void* stream) {
  if (M <= 0) return IRONB_OK;
  IRONB_REQUIRE(g && out, "eik_sum_fwd: null pointer");
  eik_sum_fwd_kernel<<<red_blocks(M), 256, 0, as_stream(stream)>>>(g, w, M, out);
  IRONB_CHECK_LAUNCH("eik_sum_fwd_kernel");
  return IRONB_OK;
}
extern "C" int ironb_eik_sum_bwd(const float* g, const float* w, const float* up, int64_t M, float* dg, void* stream) {
  if (M <= 0) return IRONB_OK;
  IRONB_REQUIRE(g && up && dg, "eik_sum_bwd: null pointer");
  eik_sum_bwd_kernel<<<blocks_for(M), 256, 0, as_stream(stream)>>>(g, w, up, M, dg);
  IRONB_CHECK_LAUNCH("eik_sum_bwd_kernel");
  return IRONB_OK;
}
/* acc: 2 floats, zeroed by the caller; loss written. */
extern "C" int ironb_roughrange_fwd(const float* r, const float* w, int64_t M, float value, float weight, float* acc, float* loss,
                                    void* stream) {
  IRONB_REQUIRE(r && w && acc && loss, "roughrange_fwd: null pointer");
  if (M > 0) {
    roughrange_fwd_kernel<<<red_blocks(M), 256, 0, as_stream(stream)>>>(r, w, M, value, acc);
    IRONB_CHECK_LAUNCH("roughrange_fwd_kernel");
  }
  roughrange_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(acc, weight, loss);
  IRONB_CHECK_LAUNCH("roughrange_finish_kernel");
  return IRONB_OK;
}
extern "C" int ironb_roughrange_bwd(const float* r, const float* w, const float* acc, const float* up, int64_t M, float value,
                                    float weight, float* dr, void* stream) {
  if (M <= 0) return IRONB_OK;
  IRONB_REQUIRE(r && w && acc && up && dr, "roughrange_bwd: null pointer");
  roughrange_bwd_kernel<<<blocks_for(M), 256, 0, as_stream(stream)>>>(r, w, acc, up, M, value, weight, dr);
  IRONB_CHECK_LAUNCH("roughrange_bwd_kernel");
  return IRONB_OK;
}
/* dst[t][m][:] = src[t][m][:] * w[m] for n <= 8 row tensors of width[t] (1 or 3 ... any small width). */
extern "C" int ironb_mask_rows(const float* const* src, float* const* dst, const int* width, int n, const float* w, int64_t M,
                               void* stream) {
  if (M <= 0 || n <= 0) return IRONB_OK;
  IRONB_REQUIRE(src && dst && width && w && n <= 8, "mask_rows: bad arguments");
  MaskArgs A;
  memset(&A, 0, sizeof(A));
  A.n = n;
  for (int t = 0; t < n; ++t) { A.src[t] = src[t]; A.dst[t] = dst[t]; A.width[t] = width[t]; }
  mask_rows_kernel<<<blocks_for(M), 256, 0, as_stream(stream)>>>(A, w, M);
  IRONB_CHECK_LAUNCH("mask_rows_kernel");
  return IRONB_OK;
}

/* flat[off[t] .. off[t] + off[n + t]) = scale * src[t][:] for t < n (src[t] == NULL: zeros).  src (n pointers) and off (n element
 * offsets, then n element counts) live in DEVICE memory, so the launch can sit in a CUDA graph and the table can be refreshed by a copy.  The
 * gradient bucket of the data-parallel step: one launch instead of cat + div + copy-back around the all-reduce. */
extern "C" int ironb_pack_tensors(const float* const* src_dev, const int64_t* off_dev, int n, int64_t max_numel, float* flat,
                                  float scale, void* stream) {
  if (n <= 0) return IRONB_OK;
  IRONB_REQUIRE(src_dev && off_dev && flat && n <= 65535, "pack_tensors: bad arguments");
  int64_t bx = ceil_div64(max_numel, 256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 128) bx = 128;
  pack_tensors_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, as_stream(stream)>>>(src_dev, off_dev, flat, scale);
  IRONB_CHECK_LAUNCH("pack_tensors_kernel");
  return IRONB_OK;
}
