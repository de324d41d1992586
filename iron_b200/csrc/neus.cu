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
  const float* bg_alpha;   // [N][n_tot] or null: the background NeRF's raw DENSITY; alpha = 1 - exp(-softplus(density) dist)
  const float* bg_dists;   // [N][n_tot] section lengths of the background sections (:147-149)
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
// F.softplus (beta 1, threshold 20) and the background alpha of render_core_outside (:167)
__device__ __forceinline__ float softplus1(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float bg_alpha_of(const NeusArgs& A, int64_t kb) {
  return 1.f - expf(-softplus1(A.bg_alpha[kb]) * A.bg_dists[kb]);
}
// d alpha / d density = dist exp(-softplus dist) sigmoid(density)
__device__ __forceinline__ float bg_dalpha_ddensity(const NeusArgs& A, int64_t kb) {
  const float x = A.bg_alpha[kb], dist = A.bg_dists[kb];
  return dist * expf(-softplus1(x) * dist) * (x > 20.f ? 1.f : sigm(x));
}

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
      alpha = alpha * in + bg_alpha_of(A, kb) * out;
#pragma unroll
      for (int j = 0; j < 3; ++j) col[j] = col[j] * in + A.bg_color[kb * 3 + j] * out;
    }
  } else {
    const int64_t kb = r * A.n_tot + i;
    alpha = bg_alpha_of(A, kb);
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
// d_inv_s [1] (accumulated: zero it first), d_bg_alpha [N][n_tot] (= d loss / d DENSITY), d_bg_color [N][n_tot][3] (null without a
// background)
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
          d_bg_alpha[kb] = d_alpha * (1.f - in) * bg_dalpha_ddensity(A, kb);
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
        d_bg_alpha[kb] = d_alpha * bg_dalpha_ddensity(A, kb);
#pragma unroll
        for (int j = 0; j < 3; ++j) d_bg_color[kb * 3 + j] = dcol[j];
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, o);
  if ((threadIdx.x & 31) == 0 && ds_acc != 0.f) atomicAdd(d_inv_s, ds_acc);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Sections of a ray batch (:254-262 / :145-160): dists_i = z_{i+1} - z_i (last: sample_dist), mid = z + dists / 2, the section
// midpoints x = o + d mid and the per-point view directions.  outside != 0: the background model's inverted-sphere points
// (x / max(|x|, 1), 1 / max(|x|, 1)), 4 wide (:153-154).
__global__ void __launch_bounds__(256) neus_sections_kernel(const float* __restrict__ ray_o, const float* __restrict__ ray_d,
                                                            const float* __restrict__ z, int64_t N, int n, float sample_dist,
                                                            int outside, float* __restrict__ dists, float* __restrict__ mid,
                                                            float* __restrict__ pts, float* __restrict__ dirs) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= N * n) return;
  const int64_t r = k / n;
  const int i = (int)(k - r * n);
  const float zi = z[k];
  const float dist = (i + 1 < n) ? z[k + 1] - zi : sample_dist;
  const float m = zi + dist * 0.5f;
  dists[k] = dist;
  mid[k] = m;
  const float d0 = ray_d[r * 3], d1 = ray_d[r * 3 + 1], d2 = ray_d[r * 3 + 2];
  float x0 = ray_o[r * 3] + d0 * m, x1 = ray_o[r * 3 + 1] + d1 * m, x2 = ray_o[r * 3 + 2] + d2 * m;
  if (dirs != nullptr) { dirs[k * 3] = d0; dirs[k * 3 + 1] = d1; dirs[k * 3 + 2] = d2; }
  if (outside) {
    const float c = fminf(fmaxf(sqrtf(x0 * x0 + x1 * x1 + x2 * x2), 1.0f), 1e10f);
    pts[k * 4] = __fdiv_rn(x0, c); pts[k * 4 + 1] = __fdiv_rn(x1, c); pts[k * 4 + 2] = __fdiv_rn(x2, c);
    pts[k * 4 + 3] = __fdiv_rn(1.0f, c);
  } else {
    pts[k * 3] = x0; pts[k * 3 + 1] = x1; pts[k * 3 + 2] = x2;
  }
}

// One hierarchical-sampling step for one ray per thread: NeuSRenderer.up_sample (:192-232) + sample_pdf(det=True) (:43-73).
// From the n current samples (z ascending, SDF values) the interval weights under a FIXED sharpness inv_s, their normalised
// CDF, and m new samples at the CDF's quantiles u_j = (j + 0.5) / m; also the new samples' points o + d z for the SDF
// evaluation that follows.  n <= NEUS_MAX_N (the CDF lives in local memory).
constexpr int NEUS_MAX_N = 256;
__global__ void __launch_bounds__(128) neus_upsample_kernel(const float* __restrict__ ray_o, const float* __restrict__ ray_d,
                                                            const float* __restrict__ z, const float* __restrict__ sdf, int64_t N,
                                                            int n, int m, float inv_s, float* __restrict__ new_z,
                                                            float* __restrict__ new_pts) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  const float o[3] = {ray_o[r * 3], ray_o[r * 3 + 1], ray_o[r * 3 + 2]};
  const float d[3] = {ray_d[r * 3], ray_d[r * 3 + 1], ray_d[r * 3 + 2]};
  const float* zr = z + r * n;
  const float* fr = sdf + r * n;
  float cdf[NEUS_MAX_N];
  auto radius = [&](float t) {
    const float x0 = o[0] + d[0] * t, x1 = o[1] + d[1] * t, x2 = o[2] + d[2] * t;
    return sqrtf(x0 * x0 + x1 * x1 + x2 * x2);
  };
  float T = 1.f, sum = 0.f, prev_cos = 0.f, rad_prev = radius(zr[0]);
  for (int i = 0; i + 1 < n; ++i) {
    const float rad_next = radius(zr[i + 1]);
    const float inside = (rad_prev < 1.0f || rad_next < 1.0f) ? 1.f : 0.f;
    const float dist = zr[i + 1] - zr[i];
    const float mid_sdf = (fr[i] + fr[i + 1]) * 0.5f;
    const float cosv = __fdiv_rn(fr[i + 1] - fr[i], dist + 1e-5f);
    // min(cos of the previous interval, cos): biased towards the steeper slope, robust where the SDF has a kink (:203-221)
    const float c = fminf(fmaxf(fminf(prev_cos, cosv), -1e3f), 0.f) * inside;
    prev_cos = cosv;
    rad_prev = rad_next;
    const float pc = sigm((mid_sdf - c * dist * 0.5f) * inv_s);
    const float nc = sigm((mid_sdf + c * dist * 0.5f) * inv_s);
    const float alpha = __fdiv_rn(pc - nc + 1e-5f, pc + 1e-5f);
    const float w = alpha * T + 1e-5f;                 // sample_pdf: weights + 1e-5
    T *= (1.f - alpha + 1e-7f);
    cdf[i + 1] = w;
    sum += w;
  }
  cdf[0] = 0.f;
  float run = 0.f;
  for (int i = 1; i < n; ++i) { run += __fdiv_rn(cdf[i], sum); cdf[i] = run; }
  // inverse CDF at ascending quantiles: the search index only moves forward
  int idx = 0;                                          // first index with cdf[idx] > u (searchsorted right=True), n if none
  const float u0 = 0.5f / (float)m, step = m > 1 ? (1.0f - 1.0f / (float)m) / (float)(m - 1) : 0.f;
  for (int j = 0; j < m; ++j) {
    const float u = u0 + step * (float)j;
    while (idx < n && !(cdf[idx] > u)) ++idx;
    const int below = idx > 0 ? idx - 1 : 0;
    const int above = idx < n - 1 ? idx : n - 1;
    float den = cdf[above] - cdf[below];
    if (den < 1e-5f) den = 1.f;
    const float t = __fdiv_rn(u - cdf[below], den);
    const float zn = zr[below] + t * (zr[above] - zr[below]);
    new_z[r * m + j] = zn;
    if (new_pts != nullptr) {
      const int64_t k = (r * m + j) * 3;
      new_pts[k] = o[0] + d[0] * zn; new_pts[k + 1] = o[1] + d[1] * zn; new_pts[k + 2] = o[2] + d[2] * zn;
    }
  }
}

// cat_z_vals (:234-246) without the sort: both sample lists of a ray are ascending, so the sorted concatenation is a two-pointer
// merge (the old sample first on ties); the SDF values travel with their samples when given.
__global__ void __launch_bounds__(128) neus_merge_kernel(const float* __restrict__ z, const float* __restrict__ sdf, int n,
                                                         const float* __restrict__ zn, const float* __restrict__ sdfn, int m,
                                                         int64_t N, float* __restrict__ out_z, float* __restrict__ out_sdf) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  const float* a = z + r * n;
  const float* b = zn + r * m;
  float* oz = out_z + r * (n + m);
  int i = 0, j = 0;
  for (int k = 0; k < n + m; ++k) {
    const bool take_old = j >= m || (i < n && a[i] <= b[j]);
    oz[k] = take_old ? a[i] : b[j];
    if (out_sdf != nullptr) out_sdf[r * (n + m) + k] = take_old ? sdf[r * n + i] : sdfn[r * m + j];
    if (take_old) ++i; else ++j;
  }
}

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int ironb_neus_sections(const float* ray_o, const float* ray_d, const float* z, int64_t N, int n, float sample_dist,
                                   int outside, float* dists, float* mid, float* pts, float* dirs, void* stream) {
  IRONB_REQUIRE(ray_o && ray_d && z && dists && mid && pts, "neus_sections: null pointer");
  IRONB_REQUIRE(N >= 0 && n >= 1, "neus_sections: bad sizes");
  if (N == 0) return IRONB_OK;
  neus_sections_kernel<<<(unsigned)ceil_div64(N * n, 256), 256, 0, as_stream(stream)>>>(ray_o, ray_d, z, N, n, sample_dist, outside,
                                                                                    dists, mid, pts, dirs);
  IRONB_CHECK_LAUNCH("neus_sections_kernel");
  return IRONB_OK;
}

extern "C" int ironb_neus_upsample(const float* ray_o, const float* ray_d, const float* z, const float* sdf, int64_t N, int n, int m,
                                   float inv_s, float* new_z, float* new_pts, void* stream) {
  IRONB_REQUIRE(ray_o && ray_d && z && sdf && new_z, "neus_upsample: null pointer");
  IRONB_REQUIRE(N >= 0 && n >= 2 && n <= NEUS_MAX_N && m >= 1, "neus_upsample: 2 <= n <= %d samples per ray, m >= 1", NEUS_MAX_N);
  if (N == 0) return IRONB_OK;
  neus_upsample_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, as_stream(stream)>>>(ray_o, ray_d, z, sdf, N, n, m, inv_s, new_z, new_pts);
  IRONB_CHECK_LAUNCH("neus_upsample_kernel");
  return IRONB_OK;
}

extern "C" int ironb_neus_merge(const float* z, const float* sdf, int n, const float* new_z, const float* new_sdf, int m, int64_t N,
                                float* out_z, float* out_sdf, void* stream) {
  IRONB_REQUIRE(z && new_z && out_z && n >= 1 && m >= 1 && N >= 0, "neus_merge: bad arguments");
  IRONB_REQUIRE(out_sdf == nullptr || (sdf && new_sdf), "neus_merge: SDF values of both lists are needed to merge them");
  if (N == 0) return IRONB_OK;
  neus_merge_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, as_stream(stream)>>>(z, sdf, n, new_z, new_sdf, m, N, out_z, out_sdf);
  IRONB_CHECK_LAUNCH("neus_merge_kernel");
  return IRONB_OK;
}

static int fill_args(NeusArgs& A, const float* ray_o, const float* ray_d, const float* mid_z, const float* dists, const float* sdf,
                     const float* grad, const float* color, const float* inv_s, const float* bg_alpha, const float* bg_dists,
                     const float* bg_color, const float* bg_rgb, int64_t N, int n, int n_tot, float anneal) {
  IRONB_REQUIRE(ray_o && ray_d && mid_z && dists && sdf && grad && color && inv_s, "neus: null pointer");
  IRONB_REQUIRE(N >= 0 && n >= 1 && n_tot >= n, "neus: bad sizes");
  IRONB_REQUIRE((bg_alpha != nullptr) == (bg_color != nullptr) && (bg_alpha != nullptr) == (bg_dists != nullptr),
                "neus: background density, section lengths and colour come together");
  IRONB_REQUIRE(n_tot == n || bg_alpha != nullptr, "neus: outside sections need the background arrays");
  A.ray_o = ray_o; A.ray_d = ray_d; A.mid_z = mid_z; A.dists = dists; A.sdf = sdf; A.grad = grad; A.color = color;
  A.inv_s = inv_s; A.bg_alpha = bg_alpha; A.bg_dists = bg_dists; A.bg_color = bg_color; A.bg_rgb = bg_rgb;
  A.N = N; A.n = n; A.n_tot = n_tot; A.anneal = anneal;
  return IRONB_OK;
}

extern "C" int ironb_neus_composite_fwd(const float* ray_o, const float* ray_d, const float* mid_z, const float* dists,
                                        const float* sdf, const float* grad, const float* color, const float* inv_s,
                                        const float* bg_density, const float* bg_dists, const float* bg_color,
                                        const float* bg_rgb, int64_t N, int n,
                                        int n_tot, float cos_anneal_ratio, float* out_color, float* weights, float* cdf,
                                        float* inside, float* acc, float* gradient_error, void* stream) {
  NeusArgs A;
  int rc = fill_args(A, ray_o, ray_d, mid_z, dists, sdf, grad, color, inv_s, bg_density, bg_dists, bg_color, bg_rgb, N, n, n_tot,
                     cos_anneal_ratio);
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
                                        const float* bg_density, const float* bg_dists, const float* bg_color,
                                        const float* bg_rgb, int64_t N, int n,
                                        int n_tot, float cos_anneal_ratio, const float* weights, const float* acc,
                                        const float* d_color, const float* d_weights, const float* d_gradient_error,
                                        float* d_sdf, float* d_grad, float* d_colors, float* d_inv_s, float* d_bg_alpha,
                                        float* d_bg_color, void* stream) {
  NeusArgs A;
  int rc = fill_args(A, ray_o, ray_d, mid_z, dists, sdf, grad, color, inv_s, bg_density, bg_dists, bg_color, bg_rgb, N, n, n_tot,
                     cos_anneal_ratio);
  if (rc) return rc;
  IRONB_REQUIRE(weights && acc && d_sdf && d_grad && d_colors && d_inv_s, "neus_composite_bwd: null pointer");
  IRONB_REQUIRE(bg_density == nullptr || (d_bg_alpha && d_bg_color), "neus_composite_bwd: background gradients missing");
  cudaStream_t st = as_stream(stream);
  IRONB_CUDA(cudaMemsetAsync(d_inv_s, 0, sizeof(float), st));
  if (N > 0) {
    neus_bwd_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, st>>>(A, weights, acc, d_color, d_weights, d_gradient_error, d_sdf,
                                                                 d_grad, d_colors, d_inv_s, d_bg_alpha, d_bg_color);
    IRONB_CHECK_LAUNCH("neus_bwd_kernel");
  }
  return IRONB_OK;
}
