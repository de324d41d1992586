// Colocated-flash GGX shading, forward and analytic backward.
// Follows GGXColocatedRenderer.forward / smithG1 (models/renderer_ggx.py:12-16, 82-146) term by term;
// the backward is the closed form of what autograd produces for it: the two
// floor-indexed table reads carry no gradient, the clamps pass gradients on the closed interval.
//
// HBM-bound elementwise kernels: 4 points per thread, every [M,3] stream moved as 3x 128-bit
// accesses per thread (12 floats = 4 points), [M] streams as one 128-bit access.
// Algorithmic traffic: 92 B/point forward, 112..136 B/point backward.
#include "common.cuh"

namespace ironb {
namespace {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kF = 0.03867f;                                   // renderer_ggx.py:110
constexpr float kInvEta2 = 1.0f / (1.48958738f * 1.48958738f);   // :99-100

struct GgxIn {
  float dist, alpha;
  float n[3], v[3], kd[3], ks[3];
};

struct GgxCommon {   // everything both directions need
  float L, d2, c_raw, c, a, c2, a2e, root, D, sin_t, cpe, tan_t, rt, h, G1, G, q, S, Kc;
};

// FAST = false: IEEE divisions (the forward pass: results within ~1e-6 of the reference).  FAST = true: the backward
// pass recomputes the same quantities with __fdividef (2 ulp, one MUFU.RCP + multiply instead of a ~10-instruction
// sequence with a slow-path call): gradients are asserted to 1e-4 relative, and the kernel drops under 128 registers.
template <bool FAST>
__device__ __forceinline__ float qdiv(float a, float b) { return FAST ? __fdividef(a, b) : a / b; }

template <bool FAST = false>
__device__ __forceinline__ GgxCommon ggx_common(const GgxIn& p, float light, const float* __restrict__ trans,
                                                const float* __restrict__ diff_trans) {
  GgxCommon s;
  s.d2 = p.dist * p.dist + 1e-10f;
  s.L = qdiv<FAST>(light, s.d2);                                            // :88
  s.c_raw = p.v[0] * p.n[0] + p.v[1] * p.n[1] + p.v[2] * p.n[2];
  s.c = fminf(fmaxf(s.c_raw, 0.00001f), 0.99999f);               // :90-91
  s.a = fmaxf(p.alpha, 0.0001f);                                 // :103
  s.c2 = s.c * s.c;
  s.a2e = s.a * s.a + 1e-10f;
  s.root = s.c2 + qdiv<FAST>(1.0f - s.c2, s.a2e);                         // :107
  s.D = qdiv<FAST>(1.0f, kPi * s.a * s.a * s.root * s.root + 1e-10f);     // :108
  s.sin_t = sqrtf(1.0f - s.c * s.c);                             // smithG1 :12-16
  s.cpe = s.c + 1e-10f;
  s.tan_t = qdiv<FAST>(s.sin_t, s.cpe);
  s.rt = s.a * s.tan_t;
  s.h = hypotf(s.rt, 1.0f);
  s.G1 = qdiv<FAST>(2.0f, 1.0f + s.h);
  s.G = s.G1 * s.G1;                                             // :111
  s.q = 4.0f * s.c + 1e-10f;
  s.S = qdiv<FAST>(kF * s.D * s.G, s.q);                                    // :113-115 without light / albedo
  // table lookups, floor indexed                                  :121-137
  float wc = sqrtf(sqrtf(s.c));                                  // c ** 0.25
  float wa = sqrtf(sqrtf(s.a * 0.25f));                          // ((alpha - 0)/(4 - 0)) ** 0.25
  int tx = (int)floorf(wc * 100.0f);
  int ty = (int)floorf(wa * 50.0f);
  int idx = min(max(ty * 100 + tx, 0), 4999);
  float T12 = fminf(fmaxf(__ldg(trans + idx), 0.0f), 1.0f);
  int idy = min(max(ty, 0), 49);
  float Fdr = fminf(fmaxf(1.0f - __ldg(diff_trans + idy), 0.0f), 1.0f);
  s.Kc = qdiv<FAST>(T12 * T12 * kInvEta2, kPi * (1.0f - Fdr + 1e-10f));   // :139-144 without light / albedo / cos
  return s;
}

// ---- 4-point vector access helpers -------------------------------------------------------------
__device__ __forceinline__ void load3x4(const float* __restrict__ base, int64_t g, float out[4][3]) {
  const float4* p = reinterpret_cast<const float4*>(base) + g * 3;
  float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  out[0][0] = a.x; out[0][1] = a.y; out[0][2] = a.z;
  out[1][0] = a.w; out[1][1] = b.x; out[1][2] = b.y;
  out[2][0] = b.z; out[2][1] = b.w; out[2][2] = c.x;
  out[3][0] = c.y; out[3][1] = c.z; out[3][2] = c.w;
}
__device__ __forceinline__ void store3x4(float* __restrict__ base, int64_t g, const float v[4][3]) {
  float4* p = reinterpret_cast<float4*>(base) + g * 3;
  p[0] = make_float4(v[0][0], v[0][1], v[0][2], v[1][0]);
  p[1] = make_float4(v[1][1], v[1][2], v[2][0], v[2][1]);
  p[2] = make_float4(v[2][2], v[3][0], v[3][1], v[3][2]);
}
__device__ __forceinline__ void load1x4(const float* __restrict__ base, int64_t g, float out[4]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(base) + g);
  out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
}
__device__ __forceinline__ void store1x4(float* __restrict__ base, int64_t g, const float v[4]) {
  reinterpret_cast<float4*>(base)[g] = make_float4(v[0], v[1], v[2], v[3]);
}

struct GgxPtrs {
  const float *light, *dist, *normal, *viewdir, *kd, *ks, *alpha, *trans, *diff_trans;
};

__device__ __forceinline__ void load_group(const GgxPtrs& P, int64_t g, int64_t M, bool full, GgxIn in[4]) {
  if (full) {
    float d[4], a[4], n[4][3], v[4][3], kd[4][3], ks[4][3];
    load1x4(P.dist, g, d); load1x4(P.alpha, g, a);
    load3x4(P.normal, g, n); load3x4(P.viewdir, g, v); load3x4(P.kd, g, kd); load3x4(P.ks, g, ks);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      in[i].dist = d[i]; in[i].alpha = a[i];
#pragma unroll
      for (int c = 0; c < 3; ++c) { in[i].n[c] = n[i][c]; in[i].v[c] = v[i][c]; in[i].kd[c] = kd[i][c]; in[i].ks[c] = ks[i][c]; }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int64_t m = g * 4 + i;
      bool ok = m < M;
      int64_t mm = ok ? m : (M - 1);
      in[i].dist = P.dist[mm]; in[i].alpha = P.alpha[mm];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        in[i].n[c] = P.normal[mm * 3 + c]; in[i].v[c] = P.viewdir[mm * 3 + c];
        in[i].kd[c] = P.kd[mm * 3 + c]; in[i].ks[c] = P.ks[mm * 3 + c];
      }
    }
  }
}

__device__ __forceinline__ void store3_group(float* base, int64_t g, int64_t M, bool full, const float v[4][3]) {
  if (base == nullptr) return;
  if (full) { store3x4(base, g, v); return; }
  for (int i = 0; i < 4; ++i) {
    int64_t m = g * 4 + i;
    if (m < M) { base[m * 3] = v[i][0]; base[m * 3 + 1] = v[i][1]; base[m * 3 + 2] = v[i][2]; }
  }
}
__device__ __forceinline__ void store1_group(float* base, int64_t g, int64_t M, bool full, const float v[4]) {
  if (base == nullptr) return;
  if (full) { store1x4(base, g, v); return; }
  for (int i = 0; i < 4; ++i) {
    int64_t m = g * 4 + i;
    if (m < M) base[m] = v[i];
  }
}

__global__ void __launch_bounds__(256) ggx_fwd_kernel(GgxPtrs P, int64_t M, float* __restrict__ o_diff,
                                                      float* __restrict__ o_spec, float* __restrict__ o_rgb) {
  const float light = __ldg(P.light);
  const int64_t groups = (M + 3) / 4;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    bool full = (g * 4 + 4 <= M);
    GgxIn in[4];
    load_group(P, g, M, full, in);
    float od[4][3], os[4][3], orgb[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      GgxCommon s = ggx_common(in[i], light, P.trans, P.diff_trans);
      float dsc = s.L * s.Kc * s.c, ssc = s.L * s.S;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        od[i][c] = in[i].kd[c] * dsc;
        os[i][c] = in[i].ks[c] * ssc;
        orgb[i][c] = od[i][c] + os[i][c];                        // :146
      }
    }
    store3_group(o_diff, g, M, full, od);
    store3_group(o_spec, g, M, full, os);
    store3_group(o_rgb, g, M, full, orgb);
  }
}

struct GgxGradPtrs {
  const float *g_diff, *g_spec, *g_rgb;
  float *d_light, *d_dist, *d_normal, *d_view, *d_kd, *d_ks, *d_alpha;
};

__device__ __forceinline__ void load_up(const float* base, int64_t g, int64_t M, bool full, float out[4][3]) {
  if (base == nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i][0] = out[i][1] = out[i][2] = 0.f;
    return;
  }
  if (full) { load3x4(base, g, out); return; }
  for (int i = 0; i < 4; ++i) {
    int64_t m = g * 4 + i;
    bool ok = m < M;
    for (int c = 0; c < 3; ++c) out[i][c] = ok ? base[m * 3 + c] : 0.f;
  }
}

__global__ void __launch_bounds__(256, 2) ggx_bwd_kernel(GgxPtrs P, GgxGradPtrs Gp, int64_t M) {
  const float light = __ldg(P.light);
  const int64_t groups = (M + 3) / 4;
  float light_acc = 0.f;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    bool full = (g * 4 + 4 <= M);
    GgxIn in[4];
    load_group(P, g, M, full, in);
    float ud[4][3], us[4][3], ur[4][3];
    load_up(Gp.g_diff, g, M, full, ud);
    load_up(Gp.g_spec, g, M, full, us);
    load_up(Gp.g_rgb, g, M, full, ur);
    float o_dist[4], o_alpha[4], o_n[4][3], o_v[4][3], o_kd[4][3], o_ks[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const GgxIn& p = in[i];
      GgxCommon s = ggx_common<true>(p, light, P.trans, P.diff_trans);
      float Ad = 0.f, As = 0.f;
      float dsc = s.L * s.Kc * s.c, ssc = s.L * s.S;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float gd = ud[i][c] + ur[i][c];
        float gs = us[i][c] + ur[i][c];
        o_kd[i][c] = gd * dsc;
        o_ks[i][c] = gs * ssc;
        Ad += gd * p.kd[c];
        As += gs * p.ks[c];
      }
      float dL = Ad * s.Kc * s.c + As * s.S;                     // d/dL of (diff + spec)
      bool valid = (g * 4 + i) < M;
      if (valid) light_acc += __fdividef(dL, s.d2);
      o_dist[i] = __fdividef(dL * light * (-2.0f * p.dist), s.d2 * s.d2);
      float dc = Ad * s.L * s.Kc;                                // diffuse cosine
      float dS = As * s.L;
      const float rq = __fdividef(1.0f, s.q);
      float dD = dS * kF * s.G * rq;
      float dG = dS * kF * s.D * rq;
      dc += dS * kF * s.D * s.G * (-4.0f) * rq * rq;
      float dden = -dD * s.D * s.D;                              // D = 1/den
      float da = dden * kPi * 2.0f * s.a * s.root * s.root;
      float droot = dden * kPi * s.a * s.a * 2.0f * s.root;
      const float ra2e = __fdividef(1.0f, s.a2e);
      float dc2 = droot * (1.0f - ra2e);
      float da2e = droot * (-(1.0f - s.c2) * ra2e * ra2e);
      da += da2e * 2.0f * s.a;
      dc += dc2 * 2.0f * s.c;
      float dG1 = dG * 2.0f * s.G1;
      float dh = -dG1 * s.G1 * s.G1 * 0.5f;                      // G1 = 2/(1+h)
      float drt = __fdividef(dh * s.rt, s.h);
      da += drt * s.tan_t;
      float dtan = drt * s.a;
      const float rcpe = __fdividef(1.0f, s.cpe);
      float dsin = dtan * rcpe;
      dc -= dtan * s.sin_t * rcpe * rcpe;
      dc += dsin * __fdividef(-s.c, s.sin_t);
      bool pass = (s.c_raw >= 0.00001f) && (s.c_raw <= 0.99999f);  // clamp backward
      float dcr = pass ? dc : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) { o_n[i][c] = dcr * p.v[c]; o_v[i][c] = dcr * p.n[c]; }
      o_alpha[i] = (p.alpha >= 0.0001f) ? da : 0.f;
    }
    store1_group(Gp.d_dist, g, M, full, o_dist);
    store1_group(Gp.d_alpha, g, M, full, o_alpha);
    store3_group(Gp.d_normal, g, M, full, o_n);
    store3_group(Gp.d_view, g, M, full, o_v);
    store3_group(Gp.d_kd, g, M, full, o_kd);
    store3_group(Gp.d_ks, g, M, full, o_ks);
  }
  // d light: warp shuffle -> one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) light_acc += __shfl_xor_sync(0xffffffffu, light_acc, o);
  if ((threadIdx.x & 31) == 0 && Gp.d_light != nullptr && light_acc != 0.f) atomicAdd(Gp.d_light, light_acc);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int ggx_grid(int64_t M) {
  int64_t groups = (M + 3) / 4;
  int64_t blocks = ceil_div64(groups, 256);
  int64_t cap = (int64_t)num_sms() * 8;   // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int ironb_ggx_fwd(const float* light, const float* dist, const float* normal, const float* viewdir,
                             const float* kd, const float* ks, const float* alpha, const float* trans,
                             const float* diff_trans, int64_t M, float* diffuse_rgb, float* specular_rgb,
                             float* rgb, void* stream) {
  IRONB_REQUIRE(M >= 0, "ggx_fwd: M < 0");
  if (M == 0) return IRONB_OK;
  IRONB_REQUIRE(light && dist && normal && viewdir && kd && ks && alpha && trans && diff_trans, "ggx_fwd: null input");
  IRONB_REQUIRE(aligned16(dist) && aligned16(normal) && aligned16(viewdir) && aligned16(kd) && aligned16(ks) &&
                    aligned16(alpha) && aligned16(diffuse_rgb) && aligned16(specular_rgb) && aligned16(rgb),
                "ggx_fwd: pointers must be 16-byte aligned");
  GgxPtrs P{light, dist, normal, viewdir, kd, ks, alpha, trans, diff_trans};
  ggx_fwd_kernel<<<ggx_grid(M), 256, 0, as_stream(stream)>>>(P, M, diffuse_rgb, specular_rgb, rgb);
  IRONB_CHECK_LAUNCH("ggx_fwd_kernel");
  return IRONB_OK;
}

extern "C" int ironb_ggx_bwd(const float* light, const float* dist, const float* normal, const float* viewdir,
                             const float* kd, const float* ks, const float* alpha, const float* trans,
                             const float* diff_trans, int64_t M, const float* g_diffuse, const float* g_specular,
                             const float* g_rgb, float* d_light, float* d_dist, float* d_normal, float* d_viewdir,
                             float* d_kd, float* d_ks, float* d_alpha, void* stream) {
  IRONB_REQUIRE(M >= 0, "ggx_bwd: M < 0");
  if (M == 0) return IRONB_OK;
  IRONB_REQUIRE(light && dist && normal && viewdir && kd && ks && alpha && trans && diff_trans, "ggx_bwd: null input");
  const void* ptrs[] = {dist, normal, viewdir, kd, ks, alpha, g_diffuse, g_specular, g_rgb,
                        d_dist, d_normal, d_viewdir, d_kd, d_ks, d_alpha};
  for (const void* p : ptrs) IRONB_REQUIRE(aligned16(p), "ggx_bwd: pointers must be 16-byte aligned");
  GgxPtrs P{light, dist, normal, viewdir, kd, ks, alpha, trans, diff_trans};
  GgxGradPtrs Gp{g_diffuse, g_specular, g_rgb, d_light, d_dist, d_normal, d_viewdir, d_kd, d_ks, d_alpha};
  ggx_bwd_kernel<<<ggx_grid(M), 256, 0, as_stream(stream)>>>(P, Gp, M);
  IRONB_CHECK_LAUNCH("ggx_bwd_kernel");
  return IRONB_OK;
}
