// Stage-1 NeuS volume renderer, the per-ray part of NeuSRenderer.render_core (models/renderer.py:248-351): from the SDF values,
// SDF gradients and colours the MLP kernels produced at the n section midpoints of every ray (plus, with a background model,
// the NeRF's alpha / colour at n_tot >= n sections) to the composited colour, the weights, the eikonal term -- and back.
//
//   true_cos = d . g;  iter_cos = -(relu(0.5 - 0.5 true_cos)(1 - a) + relu(-true_cos) a)              (:277-284, a = cos_anneal_ratio)
//   e_prev / e_next = sdf -/+ iter_cos dist / 2;  pc = sigmoid(e_prev s), nc = sigmoid(e_next s)       (:287-291)
//   alpha = clip((pc - nc + 1e-5) / (pc + 1e-5), 0, 1)                                                   (:293-296)
//   inside = |x| < 1, relax = |x| < 1.2; with a background: alpha = alpha inside + bg_alpha (1 - inside), likewise the colour,
//   then the n_tot - n outside sections are appended                                                      (:298-311)
//   w_i = alpha_i prod_{j<i} (1 - alpha_j + 1e-7);  colour = sum w_i c_i [+ bg_rgb (1 - sum w)]        (:313-318)
//   gradient_error = sum relax (|g| - 1)^2 / (sum relax + 1e-5)                                          (:321-322)
//
// One thread per ray walks its sections in order (the transmittance is a running product, like torch.cumprod): the work per
// ray is ~n_tot x 60 flops, against ~n x 4 MFLOP of MLP evaluations for the same ray, so this kernel only has to be one
// launch instead of the ~45 elementwise / scan launches (and as many again under autograd) of the reference.  The backward
// recomputes the forward quantities from the same inputs (nothing but the two eikonal sums is saved) and walks the sections
// in reverse with the suffix sum  S_i = sum_{k>i} dL/dw_k w_k:   dL/dalpha_i = dL/dw_i T_i - S_i / (1 - alpha_i + 1e-7).
#include "common.cuh"

namespace ironb {
namespace {

struct NeusArgs {
  const float* ray_o;      // [N][3]
  const float* ray_d;      // [N][3]
  const float* mid_z;      // [N][n]    section midpoints
  const float* dists;      // [N][n]
  const float* sdf;        // [N][n]
  const float* grad;       // [N][n][3]
  const float* color;      // [N][n][3]
  const float* inv_s;      // [1]
  const float* bg_alpha;   // [N][n_tot] or null
  const float* bg_color;   // [N][n_tot][3] or null
  const float* bg_rgb;     // [3] or null
  int64_t N;
  int n, n_tot;
  float anneal;
};

struct Sec {          // forward quantities of one section
  float alpha, inside, relax, gn, q, pc, nc, e_prev, e_next, tc, u, v;
};

__device__ __forceinline__ float sigm(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

__device__ __forceinline__ Sec section(const NeusArgs& A, int64_t r, int i, const float o[3], const float d[3], float s) {
  Sec c;
  const int64_t k = r * A.n + i;
  const float g0 = A.grad[k * 3], g1 = A.grad[k * 3 + 1], g2 = A.grad[k * 3 + 2];
  const float f = A.sdf[k], dist = A.dists[k], z = A.mid_z[k];
  c.tc = d[0] * g0 + d[1] * g1 + d[2] * g2;
  c.u = -c.tc * 0.5f + 0.5f;
  c.v = -c.tc;
  const float ic = -(fmaxf(c.u, 0.f) * (1.f - A.anneal) + fmaxf(c.v, 0.f) * A.anneal);
  c.e_next = f + ic * dist * 0.5f;
  c.e_prev = f - ic * dist * 0.5f;
  c.pc = sigm(c.e_prev * s);
  c.nc = sigm(c.e_next * s);
  c.q = __fdiv_rn(c.pc - c.nc + 1e-5f, c.pc + 1e-5f);
  c.alpha = fminf(fmaxf(c.q, 0.f), 1.f);
  const float x0 = o[0] + d[0] * z, x1 = o[1] + d[1] * z, x2 = o[2] + d[2] * z;
  const float pn = sqrtf(x0 * x0 + x1 * x1 + x2 * x2);
  c.inside = pn < 1.0f ? 1.f : 0.f;
  c.relax = pn < 1.2f ? 1.f : 0.f;
  c.gn = sqrtf(g0 * g0 + g1 * g1 + g2 * g2);
  return c;
}

// alpha and colour of section i after the background mix (i < n) / of outside section i (i >= n)
__device__ __forceinline__ void mixed(const NeusArgs& A, int64_t r, int i, const Sec* c, float& alpha, float col[3]) {
  if (i < A.n) {
    const int64_t k = r * A.n + i;
    alpha = c->alpha;
    col[0] = A.color[k * 3]; col[1] = A.color[k * 3 + 1]; col[2] = A.color[k * 3 + 2];
    if (A.bg_alpha != nullptr) {
      const int64_t kb = r * A.n_tot + i;
      const float in = c->inside, out = 1.f - in;
      alpha = alpha * in + A.bg_alpha[kb] * out;
#pragma unroll
      for (int j = 0; j < 3; ++j) col[j] = col[j] * in + A.bg_color[kb * 3 + j] * out;
    }
  } else {
    const int64_t kb = r * A.n_tot + i;
    alpha = A.bg_alpha[kb];
    col[0] = A.bg_color[kb * 3]; col[1] = A.bg_color[kb * 3 + 1]; col[2] = A.bg_color[kb * 3 + 2];
  }
}

// acc[0] += sum relax (|g| - 1)^2, acc[1] += sum relax
__global__ void __launch_bounds__(128) neus_fwd_kernel(NeusArgs A, float* __restrict__ out_color, float* __restrict__ weights,
                                                       float* __restrict__ cdf, float* __restrict__ inside,
                                                       float* __restrict__ acc) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float e_num = 0.f, e_cnt = 0.f;
  if (r < A.N) {
    const float o[3] = {A.ray_o[r * 3], A.ray_o[r * 3 + 1], A.ray_o[r * 3 + 2]};
    const float d[3] = {A.ray_d[r * 3], A.ray_d[r * 3 + 1], A.ray_d[r * 3 + 2]};
    const float s = *A.inv_s;
    float T = 1.f, wsum = 0.f, col[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < A.n_tot; ++i) {
      Sec c;
      if (i < A.n) {
        c = section(A, r, i, o, d, s);
        cdf[r * A.n + i] = c.pc;
        inside[r * A.n + i] = c.inside;
        e_num += c.relax * (c.gn - 1.f) * (c.gn - 1.f);
        e_cnt += c.relax;
      }
      float a, ci[3];
      mixed(A, r, i, &c, a, ci);
      const float w = a * T;
      weights[r * A.n_tot + i] = w;
      wsum += w;
      col[0] += w * ci[0]; col[1] += w * ci[1]; col[2] += w * ci[2];
      T *= (1.f - a + 1e-7f);
    }
    if (A.bg_rgb != nullptr) {
#pragma unroll
      for (int j = 0; j < 3; ++j) col[j] += A.bg_rgb[j] * (1.f - wsum);
    }
    out_color[r * 3] = col[0]; out_color[r * 3 + 1] = col[1]; out_color[r * 3 + 2] = col[2];
  }
  // block reduction of the two eikonal sums (warp shuffles, then one atomic pair per block)
  __shared__ float sh[2][4];
  for (int o = 16; o > 0; o >>= 1) { e_num += __shfl_xor_sync(0xffffffffu, e_num, o); e_cnt += __shfl_xor_sync(0xffffffffu, e_cnt, o); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = e_num; sh[1][threadIdx.x >> 5] = e_cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; }
    if (b != 0.f) { atomicAdd(acc, a); atomicAdd(acc + 1, b); }
  }
}

__global__ void neus_finish_kernel(const float* __restrict__ acc, float* __restrict__ gerr) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *gerr = __fdiv_rn(acc[0], acc[1] + 1e-5f);
}

// d_color [N][3], d_weights [N][n_tot] or null, d_gerr [1] or null  ->  d_sdf [N][n], d_grad [N][n][3], d_colors [N][n][3],
// d_inv_s [1] (accumulated: zero it first), d_bg_alpha [N][n_tot], d_bg_color [N][n_tot][3] (null without a background)
__global__ void __launch_bounds__(128) neus_bwd_kernel(NeusArgs A, const float* __restrict__ weights, const float* __restrict__ acc,
                                                       const float* __restrict__ d_color, const float* __restrict__ d_weights,
                                                       const float* __restrict__ d_gerr, float* __restrict__ d_sdf,
                                                       float* __restrict__ d_grad, float* __restrict__ d_colors,
                                                       float* __restrict__ d_inv_s, float* __restrict__ d_bg_alpha,
                                                       float* __restrict__ d_bg_color) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float ds_acc = 0.f;
  if (r < A.N) {
    const float o[3] = {A.ray_o[r * 3], A.ray_o[r * 3 + 1], A.ray_o[r * 3 + 2]};
    const float d[3] = {A.ray_d[r * 3], A.ray_d[r * 3 + 1], A.ray_d[r * 3 + 2]};
    const float s = *A.inv_s;
    const float dc[3] = {d_color ? d_color[r * 3] : 0.f, d_color ? d_color[r * 3 + 1] : 0.f, d_color ? d_color[r * 3 + 2] : 0.f};
    float dbg = 0.f;                                     // d colour . bg_rgb: every weight also lowers the background share
    if (A.bg_rgb != nullptr) dbg = dc[0] * A.bg_rgb[0] + dc[1] * A.bg_rgb[1] + dc[2] * A.bg_rgb[2];
    const float eik_scale = d_gerr ? __fdiv_rn(*d_gerr, acc[1] + 1e-5f) : 0.f;
    float S = 0.f;                                       // sum_{k>i} dL/dw_k w_k
    for (int i = A.n_tot - 1; i >= 0; --i) {
      Sec c;
      if (i < A.n) c = section(A, r, i, o, d, s);
      float a, ci[3];
      mixed(A, r, i, &c, a, ci);
      const float w = weights[r * A.n_tot + i];
      const float dW = dc[0] * ci[0] + dc[1] * ci[1] + dc[2] * ci[2] - dbg + (d_weights ? d_weights[r * A.n_tot + i] : 0.f);
      // T_i = w / alpha when alpha > 0; recompute it exactly from the stored weight where possible, else it multiplies 0 anyway
      const float om = 1.f - a + 1e-7f;
      const float T = a > 0.f ? __fdiv_rn(w, a) : 0.f;
      float d_alpha = dW * T - __fdiv_rn(S, om);
      if (!(a > 0.f)) {
        // alpha == 0: w == 0 and T is not recoverable from w; dL/dalpha_i = dW_i T_i needs T_i: walk it forward (rare)
        float Tf = 1.f;
        for (int j = 0; j < i; ++j) {
          Sec cj;
          if (j < A.n) cj = section(A, r, j, o, d, s);
          float aj, cjj[3];
          mixed(A, r, j, &cj, aj, cjj);
          Tf *= (1.f - aj + 1e-7f);
        }
        d_alpha = dW * Tf - __fdiv_rn(S, om);
      }
      S += dW * w;
      const float dcol[3] = {w * dc[0], w * dc[1], w * dc[2]};
      float d_alpha_f = d_alpha;
      if (i < A.n) {
        const int64_t k = r * A.n + i;
        float in = 1.f;
        if (A.bg_alpha != nullptr) {
          in = c.inside;
          const int64_t kb = r * A.n_tot + i;
          d_bg_alpha[kb] = d_alpha * (1.f - in);
#pragma unroll
          for (int j = 0; j < 3; ++j) d_bg_color[kb * 3 + j] = dcol[j] * (1.f - in);
          d_alpha_f = d_alpha * in;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) d_colors[k * 3 + j] = dcol[j] * in;
        // alpha = clip(q, 0, 1)
        const float dq = (c.q >= 0.f && c.q <= 1.f) ? d_alpha_f : 0.f;
        const float den = c.pc + 1e-5f;
        const float dp = __fdiv_rn(dq, den);
        const float dcc = -dq * __fdiv_rn(c.pc - c.nc + 1e-5f, den * den);
        const float dpc = dp + dcc, dnc = -dp;
        const float dzp = dpc * c.pc * (1.f - c.pc);      // d / d (e_prev s)
        const float dzn = dnc * c.nc * (1.f - c.nc);      // d / d (e_next s)
        ds_acc += c.e_prev * dzp + c.e_next * dzn;
        const float dep = s * dzp, den_ = s * dzn;
        d_sdf[k] = dep + den_;
        const float dic = (den_ - dep) * A.dists[k] * 0.5f;
        const float dtc = (0.5f * (1.f - A.anneal) * (c.u > 0.f ? 1.f : 0.f) + A.anneal * (c.v > 0.f ? 1.f : 0.f)) * dic;
        const float g0 = A.grad[k * 3], g1 = A.grad[k * 3 + 1], g2 = A.grad[k * 3 + 2];
        const float ek = (c.gn > 0.f) ? eik_scale * c.relax * 2.f * __fdiv_rn(c.gn - 1.f, c.gn) : 0.f;
        d_grad[k * 3] = dtc * d[0] + ek * g0;
        d_grad[k * 3 + 1] = dtc * d[1] + ek * g1;
        d_grad[k * 3 + 2] = dtc * d[2] + ek * g2;
      } else {
        const int64_t kb = r * A.n_tot + i;
        d_bg_alpha[kb] = d_alpha;
#pragma unroll
        for (int j = 0; j < 3; ++j) d_bg_color[kb * 3 + j] = dcol[j];
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, o);
  if ((threadIdx.x & 31) == 0 && ds_acc != 0.f) atomicAdd(d_inv_s, ds_acc);
}

}  // namespace
}  // namespace ironb

using namespace ironb;

static int fill_args(NeusArgs& A, const float* ray_o, const float* ray_d, const float* mid_z, const float* dists, const float* sdf,
                     const float* grad, const float* color, const float* inv_s, const float* bg_alpha, const float* bg_color,
                     const float* bg_rgb, int64_t N, int n, int n_tot, float anneal) {
  IRONB_REQUIRE(ray_o && ray_d && mid_z && dists && sdf && grad && color && inv_s, "neus: null pointer");
  IRONB_REQUIRE(N >= 0 && n >= 1 && n_tot >= n, "neus: bad sizes");
  IRONB_REQUIRE((bg_alpha != nullptr) == (bg_color != nullptr), "neus: background alpha and colour come together");
  IRONB_REQUIRE(n_tot == n || bg_alpha != nullptr, "neus: outside sections need the background arrays");
  A.ray_o = ray_o; A.ray_d = ray_d; A.mid_z = mid_z; A.dists = dists; A.sdf = sdf; A.grad = grad; A.color = color;
  A.inv_s = inv_s; A.bg_alpha = bg_alpha; A.bg_color = bg_color; A.bg_rgb = bg_rgb;
  A.N = N; A.n = n; A.n_tot = n_tot; A.anneal = anneal;
  return IRONB_OK;
}

extern "C" int ironb_neus_composite_fwd(const float* ray_o, const float* ray_d, const float* mid_z, const float* dists,
                                        const float* sdf, const float* grad, const float* color, const float* inv_s,
                                        const float* bg_alpha, const float* bg_color, const float* bg_rgb, int64_t N, int n,
                                        int n_tot, float cos_anneal_ratio, float* out_color, float* weights, float* cdf,
                                        float* inside, float* acc, float* gradient_error, void* stream) {
  NeusArgs A;
  int rc = fill_args(A, ray_o, ray_d, mid_z, dists, sdf, grad, color, inv_s, bg_alpha, bg_color, bg_rgb, N, n, n_tot, cos_anneal_ratio);
  if (rc) return rc;
  IRONB_REQUIRE(out_color && weights && cdf && inside && acc && gradient_error, "neus_composite_fwd: null output");
  cudaStream_t st = as_stream(stream);
  IRONB_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(float), st));
  if (N > 0) {
    neus_fwd_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, st>>>(A, out_color, weights, cdf, inside, acc);
    IRONB_CHECK_LAUNCH("neus_fwd_kernel");
  }
  neus_finish_kernel<<<1, 32, 0, st>>>(acc, gradient_error);
  IRONB_CHECK_LAUNCH("neus_finish_kernel");
  return IRONB_OK;
}

extern "C" int ironb_neus_composite_bwd(const float* ray_o, const float* ray_d, const float* mid_z, const float* dists,
                                        const float* sdf, const float* grad, const float* color, const float* inv_s,
                                        const float* bg_alpha, const float* bg_color, const float* bg_rgb, int64_t N, int n,
                                        int n_tot, float cos_anneal_ratio, const float* weights, const float* acc,
                                        const float* d_color, const float* d_weights, const float* d_gradient_error,
                                        float* d_sdf, float* d_grad, float* d_colors, float* d_inv_s, float* d_bg_alpha,
                                        float* d_bg_color, void* stream) {
  NeusArgs A;
  int rc = fill_args(A, ray_o, ray_d, mid_z, dists, sdf, grad, color, inv_s, bg_alpha, bg_color, bg_rgb, N, n, n_tot, cos_anneal_ratio);
  if (rc) return rc;
  IRONB_REQUIRE(weights && acc && d_sdf && d_grad && d_colors && d_inv_s, "neus_composite_bwd: null pointer");
  IRONB_REQUIRE(bg_alpha == nullptr || (d_bg_alpha && d_bg_color), "neus_composite_bwd: background gradients missing");
  cudaStream_t st = as_stream(stream);
  IRONB_CUDA(cudaMemsetAsync(d_inv_s, 0, sizeof(float), st));
  if (N > 0) {
    neus_bwd_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, st>>>(A, weights, acc, d_color, d_weights, d_gradient_error, d_sdf,
                                                                 d_grad, d_colors, d_inv_s, d_bg_alpha, d_bg_color);
    IRONB_CHECK_LAUNCH("neus_bwd_kernel");
  }
  return IRONB_OK;
}
