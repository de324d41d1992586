// CompositeRenderer.forward (models/renderer_ggx.py:781-858), the colocated-flash "comp2" shading model of the fork's working
// driver (model_bed.py:227-298): table-driven rough-plastic diffuse lobe + an exact-Fresnel conductor lobe + a dielectric
// microfacet lobe, forward and backward.
//
//   c    = clamp(v . n, 1e-5, 0.99999)                                         :798-799
//   L    = light / (dist^2 + 1e-10)                                            :814
//   D    = calc_D_specular(c, eta = 1.48958738)   -- the reference passes ETA where alpha belongs (:803); kept
//   G    = smithG1(c, alpha)^2,  alpha = clamp(specular_roughness, 1e-5)       :785, 802-804
//   Fm   = fresnel_conductor_exact(c, metallic_eta, metallic_k)                :806, 592-606 (clamps :787-788)
//   Fd   = fresnel_dielectric(c, c, dielectric_eta)  (the MODULE-level function, :398-416, called at :613)
//   metallic_rgb   = ks * Fm * L                                               :820-826
//   dielectric_rgb = ks * Fd * D * G / (4 |c|) * L                             :823-827
//   specular_rgb   = dielectric_rgb + metallic_rgb                             :830 (the metallic / dielectric weights of
//                                                                              :828 are overwritten: they get no gradient)
//   diffuse        = L * kd / (1 - Fdr + 1e-10) / pi * c * T12^2 / 1.48958738^2   (tables floor-indexed: no gradient)
//   rgb            = diffuse + specular_rgb, accumulated IN PLACE into the diffuse tensor (:846-851): the reference's
//                    "diffuse_rgb" output therefore equals "rgb"; the Python module returns the same tensor for both.
//
// The backward recomputes the three scalar lobes with forward-mode dual numbers over (c, alpha, metallic_eta, metallic_k,
// dielectric_eta), so the derivative code IS the forward code (no hand-derived Fresnel gradients to get wrong); the clamps
// pass gradients on the closed interval like torch.clamp.  HBM-bound elementwise kernels, one point per thread.
#include "common.cuh"

namespace ironb {
namespace {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kEta = 1.48958738f;
constexpr float kInvEta2 = 1.0f / (kEta * kEta);
constexpr int NV = 5;   // c, alpha, metallic_eta, metallic_k, dielectric_eta

template <int N>
struct Dual {
  float v;
  float d[N];
};
struct Plain { float v; };

template <int N> __device__ __forceinline__ Dual<N> mk_const(float x, Dual<N>*) { Dual<N> r; r.v = x; for (int i = 0; i < N; ++i) r.d[i] = 0.f; return r; }
__device__ __forceinline__ Plain mk_const(float x, Plain*) { return Plain{x}; }
template <int N> __device__ __forceinline__ Dual<N> mk_var(float x, int k, Dual<N>*) { Dual<N> r = mk_const(x, (Dual<N>*)nullptr); r.d[k] = 1.f; return r; }
__device__ __forceinline__ Plain mk_var(float x, int, Plain*) { return Plain{x}; }

template <int N> __device__ __forceinline__ Dual<N> operator+(Dual<N> a, Dual<N> b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(Dual<N> a, Dual<N> b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator*(Dual<N> a, Dual<N> b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator/(Dual<N> a, Dual<N> b) {
  Dual<N> r; const float ib = 1.f / b.v; r.v = a.v * ib;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator+(Dual<N> a, float b) { a.v += b; return a; }
template <int N> __device__ __forceinline__ Dual<N> operator*(Dual<N> a, float b) { a.v *= b; for (int i = 0; i < N; ++i) a.d[i] *= b; return a; }
template <int N> __device__ __forceinline__ Dual<N> operator*(float b, Dual<N> a) { return a * b; }
template <int N> __device__ __forceinline__ Dual<N> operator-(float b, Dual<N> a) { a.v = b - a.v; for (int i = 0; i < N; ++i) a.d[i] = -a.d[i]; return a; }
template <int N> __device__ __forceinline__ Dual<N> operator/(float b, Dual<N> a) { return mk_const(b, (Dual<N>*)nullptr) / a; }
template <int N> __device__ __forceinline__ Dual<N> dsqrt(Dual<N> a) { Dual<N> r; r.v = sqrtf(a.v); const float h = 0.5f / r.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * h; return r; }

__device__ __forceinline__ Plain operator+(Plain a, Plain b) { return Plain{a.v + b.v}; }
__device__ __forceinline__ Plain operator-(Plain a, Plain b) { return Plain{a.v - b.v}; }
__device__ __forceinline__ Plain operator*(Plain a, Plain b) { return Plain{a.v * b.v}; }
__device__ __forceinline__ Plain operator/(Plain a, Plain b) { return Plain{a.v / b.v}; }
__device__ __forceinline__ Plain operator+(Plain a, float b) { return Plain{a.v + b}; }
__device__ __forceinline__ Plain operator*(Plain a, float b) { return Plain{a.v * b}; }
__device__ __forceinline__ Plain operator*(float b, Plain a) { return Plain{a.v * b}; }
__device__ __forceinline__ Plain operator-(float b, Plain a) { return Plain{b - a.v}; }
__device__ __forceinline__ Plain operator/(float b, Plain a) { return Plain{b / a.v}; }
__device__ __forceinline__ Plain dsqrt(Plain a) { return Plain{sqrtf(a.v)}; }

// fresnel_conductor_exact, models/renderer_ggx.py:592-606
template <class T>
__device__ __forceinline__ T fresnel_conductor(T c, T eta, T k) {
  const T c2 = c * c;
  const T s2 = 1.f - c2;
  const T s4 = s2 * s2;
  const T t1 = eta * eta - k * k - s2;
  const T a2pb2 = dsqrt(t1 * t1 + 4.f * (k * k) * (eta * eta));
  const T a = dsqrt(0.5f * (a2pb2 + t1));
  const T term1 = a2pb2 + c2;
  const T term2 = 2.f * (a * c);
  const T Rs2 = (term1 - term2) / (term1 + term2);
  const T term3 = a2pb2 * c2 + s4;
  const T term4 = term2 * s2;
  const T Rp2 = Rs2 * ((term3 - term4) / (term3 + term4));
  return 0.5f * (Rp2 + Rs2);
}

// fresnel_dielectric(cosThetaI = c, cosThetaT = c, eta), models/renderer_ggx.py:398-416; c > 0 here, so scale = 1 / eta
template <class T>
__device__ __forceinline__ T fresnel_dielectric(T c, T eta) {
  const T scale = 1.f / eta;
  const T ct = dsqrt(1.f - (1.f - c * c) * (scale * scale));
  const T Rs = (c - eta * ct) / (c + eta * ct);
  const T Rp = (eta * c - ct) / (eta * c + ct);
  return 0.5f * (Rs * Rs + Rp * Rp);
}

// the dielectric lobe without albedo and light: Fd * D(c; "alpha" = eta) * smithG1(c, alpha)^2 / (4 |c|)
template <class T>
__device__ __forceinline__ T dielectric_lobe(T c, T alpha, T deta) {
  const T Fd = fresnel_dielectric(c, deta);
  const T c2 = c * c;
  const T root = c2 + (1.f - c2) * (1.f / (kEta * kEta + 1e-10f));              // calc_D_specular(c, eta), :767-771, 803
  const T D = 1.f / ((kPi * kEta * kEta) * (root * root) + 1e-10f);
  const T sin_t = dsqrt(1.f - c * c);                                             // smithG1, :12-16
  const T tan_t = sin_t / (c + 1e-10f);
  const T rt = alpha * tan_t;
  const T G1 = 2.f / (dsqrt(rt * rt + 1.f) + 1.f);                                // torch.hypot(root, 1)
  return Fd * D * (G1 * G1) / (4.f * c);
}

struct CompIn {
  const float *light, *dist, *normal, *viewdir, *kd, *ks, *alpha, *meta, *mk, *deta, *trans, *diff_trans;
};

struct Pt {
  float L, d2, c, c_pass, alpha, a_pass, meta, m_pass, mk, k_pass, deta, e_pass, Kc;
  float n[3], v[3], kd[3], ks[3], kd_pass[3], ks_pass[3];
};

__device__ __forceinline__ float clampf(float x, float lo, float hi, float& pass) {
  pass = (x >= lo && x <= hi) ? 1.f : 0.f;
  return fminf(fmaxf(x, lo), hi);
}

__device__ __forceinline__ Pt load_point(const CompIn& I, int64_t m) {
  Pt p;
  const float dist = I.dist[m];
  p.d2 = dist * dist + 1e-10f;
  p.L = *I.light / p.d2;                                                          // :814
  float cr = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    p.n[j] = I.normal[m * 3 + j]; p.v[j] = I.viewdir[m * 3 + j];
    cr += p.v[j] * p.n[j];
    float pass;
    p.kd[j] = clampf(I.kd[m * 3 + j], 0.00001f, INFINITY, pass); p.kd_pass[j] = pass;   // :790
    p.ks[j] = clampf(I.ks[m * 3 + j], 0.00001f, INFINITY, pass); p.ks_pass[j] = pass;   // :789
  }
  p.c = clampf(cr, 0.00001f, 0.99999f, p.c_pass);                                 // :798-799
  p.alpha = clampf(I.alpha[m], 0.00001f, INFINITY, p.a_pass);                     // :785
  p.deta = clampf(I.deta[m], 1.000001f, 1.999999f, p.e_pass);                     // :786
  p.meta = clampf(I.meta[m], 0.099999f, 4.999999f, p.m_pass);                     // :787
  p.mk = clampf(I.mk[m], 0.099999f, 9.999999f, p.k_pass);                         // :788
  // diffuse_reflection_ggx (:669-697): table lookups with alpha re-clamped at 1e-4, floor indexed
  const float a4 = fmaxf(p.alpha, 0.0001f);
  const float wc = sqrtf(sqrtf(p.c));
  const float wa = sqrtf(sqrtf(a4 * 0.25f));
  const int tx = (int)floorf(wc * 100.0f), ty = (int)floorf(wa * 50.0f);
  const int idx = min(max(ty * 100 + tx, 0), 4999);
  const float T12 = fminf(fmaxf(__ldg(I.trans + idx), 0.0f), 1.0f);
  const float Fdr = fminf(fmaxf(1.0f - __ldg(I.diff_trans + min(max(ty, 0), 49)), 0.0f), 1.0f);
  p.Kc = T12 * T12 * kInvEta2 / (kPi * (1.0f - Fdr + 1e-10f));
  return p;
}

__global__ void __launch_bounds__(256) comp_fwd_kernel(CompIn I, int64_t M, float* __restrict__ rgb, float* __restrict__ spec,
                                                       float* __restrict__ met, float* __restrict__ diel) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const Pt p = load_point(I, m);
  const float Fm = fresnel_conductor(Plain{p.c}, Plain{p.meta}, Plain{p.mk}).v;
  const float Sd = dielectric_lobe(Plain{p.c}, Plain{p.alpha}, Plain{p.deta}).v;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float mr = p.ks[j] * Fm * p.L;
    const float dr = p.ks[j] * Sd * p.L;
    const float df = p.L * p.kd[j] * p.Kc * p.c;
    met[m * 3 + j] = mr;
    diel[m * 3 + j] = dr;
    spec[m * 3 + j] = dr + mr;
    rgb[m * 3 + j] = df + (dr + mr);
  }
}

struct CompGrad {
  const float *g_rgb, *g_spec, *g_met, *g_diel;     // upstream (any may be NULL); g_rgb already includes the diffuse alias
  float *d_light, *d_dist, *d_normal, *d_kd, *d_ks, *d_alpha, *d_meta, *d_mk, *d_deta;
};

__global__ void __launch_bounds__(256) comp_bwd_kernel(CompIn I, int64_t M, CompGrad Gp) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float dL = 0.f;
  if (m < M) {
    const Pt p = load_point(I, m);
    typedef Dual<NV> D5;
    const D5 c = mk_var(p.c, 0, (D5*)nullptr), al = mk_var(p.alpha, 1, (D5*)nullptr), me = mk_var(p.meta, 2, (D5*)nullptr),
             k = mk_var(p.mk, 3, (D5*)nullptr), de = mk_var(p.deta, 4, (D5*)nullptr);
    const D5 Fm = fresnel_conductor(c, me, k);
    const D5 Sd = dielectric_lobe(c, al, de);
    float Am = 0.f, Ad = 0.f, Af = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float gr = Gp.g_rgb ? Gp.g_rgb[m * 3 + j] : 0.f, gs = Gp.g_spec ? Gp.g_spec[m * 3 + j] : 0.f;
      const float gm = gr + gs + (Gp.g_met ? Gp.g_met[m * 3 + j] : 0.f);
      const float gd = gr + gs + (Gp.g_diel ? Gp.g_diel[m * 3 + j] : 0.f);
      Gp.d_ks[m * 3 + j] = p.ks_pass[j] * p.L * (gm * Fm.v + gd * Sd.v);
      Gp.d_kd[m * 3 + j] = p.kd_pass[j] * gr * p.L * p.Kc * p.c;
      Am += gm * p.ks[j];
      Ad += gd * p.ks[j];
      Af += gr * p.kd[j];
    }
    dL = Am * Fm.v + Ad * Sd.v + Af * p.Kc * p.c;                 // d loss / d L
    Am *= p.L; Ad *= p.L; Af *= p.L * p.Kc;
    const float dc = p.c_pass * (Am * Fm.d[0] + Ad * Sd.d[0] + Af);
#pragma unroll
    for (int j = 0; j < 3; ++j) Gp.d_normal[m * 3 + j] = dc * p.v[j];
    Gp.d_alpha[m] = p.a_pass * Ad * Sd.d[1];
    Gp.d_meta[m] = p.m_pass * Am * Fm.d[2];
    Gp.d_mk[m] = p.k_pass * Am * Fm.d[3];
    Gp.d_deta[m] = p.e_pass * Ad * Sd.d[4];
    const float dist = I.dist[m];
    Gp.d_dist[m] = -dL * (*I.light) * 2.f * dist / (p.d2 * p.d2);
    dL = dL / p.d2;                                               // contribution to d loss / d light
  }
  // d light: warp shuffle -> one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dL += __shfl_xor_sync(0xffffffffu, dL, o);
  if ((threadIdx.x & 31) == 0 && Gp.d_light != nullptr && dL != 0.f) atomicAdd(Gp.d_light, dL);
}

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int ironb_composite_fwd(const float* light, const float* dist, const float* normal, const float* viewdir,
                                   const float* kd, const float* ks, const float* alpha, const float* metallic_eta,
                                   const float* metallic_k, const float* dielectric_eta, const float* trans,
                                   const float* diff_trans, int64_t M, float* rgb, float* specular_rgb, float* metallic_rgb,
                                   float* dielectric_rgb, void* stream) {
  IRONB_REQUIRE(M >= 0, "composite_fwd: M < 0");
  if (M == 0) return IRONB_OK;
  IRONB_REQUIRE(light && dist && normal && viewdir && kd && ks && alpha && metallic_eta && metallic_k && dielectric_eta && trans &&
                    diff_trans && rgb && specular_rgb && metallic_rgb && dielectric_rgb, "composite_fwd: null pointer");
  CompIn I{light, dist, normal, viewdir, kd, ks, alpha, metallic_eta, metallic_k, dielectric_eta, trans, diff_trans};
  comp_fwd_kernel<<<(unsigned)ceil_div64(M, 256), 256, 0, as_stream(stream)>>>(I, M, rgb, specular_rgb, metallic_rgb, dielectric_rgb);
  IRONB_CHECK_LAUNCH("comp_fwd_kernel");
  return IRONB_OK;
}

extern "C" int ironb_composite_bwd(const float* light, const float* dist, const float* normal, const float* viewdir,
                                   const float* kd, const float* ks, const float* alpha, const float* metallic_eta,
                                   const float* metallic_k, const float* dielectric_eta, const float* trans,
                                   const float* diff_trans, int64_t M, const float* g_rgb, const float* g_specular,
                                   const float* g_metallic, const float* g_dielectric, float* d_light, float* d_dist,
                                   float* d_normal, float* d_kd, float* d_ks, float* d_alpha, float* d_metallic_eta,
                                   float* d_metallic_k, float* d_dielectric_eta, void* stream) {
  IRONB_REQUIRE(M >= 0, "composite_bwd: M < 0");
  if (M == 0) return IRONB_OK;
  IRONB_REQUIRE(light && dist && normal && viewdir && kd && ks && alpha && metallic_eta && metallic_k && dielectric_eta && trans &&
                    diff_trans, "composite_bwd: null input");
  IRONB_REQUIRE(d_dist && d_normal && d_kd && d_ks && d_alpha && d_metallic_eta && d_metallic_k && d_dielectric_eta,
                "composite_bwd: null gradient output");
  CompIn I{light, dist, normal, viewdir, kd, ks, alpha, metallic_eta, metallic_k, dielectric_eta, trans, diff_trans};
  CompGrad Gp{g_rgb, g_specular, g_metallic, g_dielectric, d_light, d_dist, d_normal, d_kd, d_ks, d_alpha, d_metallic_eta,
              d_metallic_k, d_dielectric_eta};
  comp_bwd_kernel<<<(unsigned)ceil_div64(M, 256), 256, 0, as_stream(stream)>>>(I, M, Gp);
  IRONB_CHECK_LAUNCH("comp_bwd_kernel");
  return IRONB_OK;
}
