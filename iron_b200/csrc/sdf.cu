// SDFNetwork forward, analytic input gradient, and the double backward, layer at a time.
//
// Restates models/fields.py:82-137 (+ models/embedder.py:11-36) as explicit linear algebra (SURVEY.md
// appendix A):   a_0 = e(x*s);  u_l = a_l  (or cat(a_l, e)/sqrt2 at the skip layer);  z_l = W_l u_l + b_l;
// a_{l+1} = softplus_beta(z_l);  y = z_last[0]/s, feature = z_last[1:].
//   grad = dy/dx:  q_last = W_last[0,:]/s;  r_l = sp'(z_l) * qa_{l+1};  q_l = W_l^T r_l;  n = s * Je^T p.
//   backward of (y, feature, grad) w.r.t. W, b: part B (through grad, ascending l) then part A (ordinary
//   back-propagation, descending l), with  zbarB_l = sp''(z_l) qa_{l+1} rbar_l = beta (1 - sp'(z_l)) r_l rbar_l.
// Every product is one fp32 tile GEMM (gemm.cuh) with the elementwise work fused into its epilogue.
#include "gemm_h16.cuh"

namespace ironb {
namespace {

// ------------------------------------------------------------------ positional encoding
// e = [x', sin(2^0 x'), cos(2^0 x'), ..., sin(2^(L-1) x'), cos(2^(L-1) x')], x' = x*scale; pad columns = 0.
// ehi / elo (optional): the fp16x2-split copy of the row, the first layer's tensor-core operand (gemm_h16.cuh).
__global__ void __launch_bounds__(256) pe_kernel(const float* __restrict__ x, int64_t M, int multires, float scale,
                                                 int Epad, float* __restrict__ e, __half* __restrict__ ehi,
                                                 __half* __restrict__ elo) {
  int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  float xs[3] = {x[m * 3] * scale, x[m * 3 + 1] * scale, x[m * 3 + 2] * scale};
  float* o = e + m * Epad;
  o[0] = xs[0]; o[1] = xs[1]; o[2] = xs[2];
  int w = 3;
  float f = 1.f;
  for (int k = 0; k < multires; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      sincosf(xs[c] * f, &s, &co);     // accurate path: arguments reach 32 rad
      o[w + c] = s;
      o[w + 3 + c] = co;
    }
    w += 6;
    f *= 2.f;
  }
  for (; w < Epad; ++w) o[w] = 0.f;
  if (ehi != nullptr) {      // the row was just written by this thread: re-read it 4 floats at a time (Epad is a multiple of 8)
    for (int i = 0; i < Epad; i += 4) {
      const float4 v4 = *reinterpret_cast<const float4*>(o + i);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      h16::store_split4(ehi, elo, m * Epad + i, v);
    }
  }
}

// ------------------------------------------------------------------ epilogues
struct EpiFwdHidden {
  const float* bias;   // [Npad]
  float* Z;            // [M][ld] or null
  float* Unext;        // [M][ld]
  float* R;            // [M][ld] or null: r = sp'(z) * qrow[n]  (only for the layer before the last)
  const float* qrow;   // W_last[0,:] (divided by scale via qscale)
  const float* e;      // PE buffer (pre-skip layer only)
  int ld, n_true, pre_skip, Epad, E;
  float beta, qscale;
  __half *Uh, *Ul;     // fp16x2-split copy of Unext (next layer's tensor-core operand) or null
  __half *Rh, *Rl;     // the same for R
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[4]) const {
    float zz[4], uu[4], rr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + j;
      if (n < n_true) {
        float z = acc[j] + __ldg(bias + n);
        float a = softplus_beta_fast(z, beta);
        if (pre_skip) a = __fdiv_rn(a, IRONB_SQRT2F);
        zz[j] = z; uu[j] = a;
        rr[j] = R ? softplus_d1_fast(z, beta) * (__ldg(qrow + n) * qscale) : 0.f;
      } else {
        zz[j] = 0.f; rr[j] = 0.f;
        int c = n - n_true;
        uu[j] = (pre_skip && c < E) ? __fdiv_rn(__ldg(e + (int64_t)m * Epad + c), IRONB_SQRT2F) : 0.f;
      }
    }
    int64_t o = (int64_t)m * ld + n0;
    if (Z) *reinterpret_cast<float4*>(Z + o) = make_float4(zz[0], zz[1], zz[2], zz[3]);
    *reinterpret_cast<float4*>(Unext + o) = make_float4(uu[0], uu[1], uu[2], uu[3]);
    if (Uh) h16::store_split4(Uh, Ul, o, uu);
    if (R) {
      *reinterpret_cast<float4*>(R + o) = make_float4(rr[0], rr[1], rr[2], rr[3]);
      if (Rh) h16::store_split4(Rh, Rl, o, rr);
    }
  }
};

struct EpiFwdLast {
  const float* bias;
  float* y;      // [M] or null
  float* feat;   // [M][d_out-1] or null
  int d_out;
  float inv_scale;
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[4]) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + j;
      if (n >= d_out) continue;
      float z = acc[j] + __ldg(bias + n);
      if (n == 0) { if (y) y[m] = z * inv_scale; }
      else if (feat) feat[(int64_t)m * (d_out - 1) + (n - 1)] = z;
    }
  }
};

// reverse chain: acc = q_l[m][k] (gradient of y w.r.t. u_l).  For l >= 1 it becomes r_{l-1}; for l == 0 it
// is added into the PE-gradient accumulator P.
struct EpiQ {
  const float* Zprev;  // Z_{l-1} [M][ld]       (l >= 1)
  float* Rprev;        // R_{l-1} [M][ld]       (l >= 1)
  float* P;            // [M][Epad]
  int ld, n_true_prev, is_skip, is_first, Epad, E;
  float beta;
  __half *Rh, *Rl;     // fp16x2-split copy of Rprev (the next product's tensor-core operand) or null
  __device__ __forceinline__ void operator()(int m, int k0, const float (&acc)[4]) const {
    if (is_first) {
      float4* p = reinterpret_cast<float4*>(P + (int64_t)m * Epad + k0);
      float4 v = *p;
      v.x += acc[0]; v.y += acc[1]; v.z += acc[2]; v.w += acc[3];
      *p = v;
      return;
    }
    float rr[4];
    int64_t o = (int64_t)m * ld + k0;
    float4 z4 = *reinterpret_cast<const float4*>(Zprev + o);
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + j;
      if (k < n_true_prev) {
        float qa = is_skip ? __fdiv_rn(acc[j], IRONB_SQRT2F) : acc[j];
        rr[j] = softplus_d1_fast(zz[j], beta) * qa;
      } else {
        rr[j] = 0.f;
        int c = k - n_true_prev;
        if (is_skip && c < E) P[(int64_t)m * Epad + c] = __fdiv_rn(acc[j], IRONB_SQRT2F);
      }
    }
    *reinterpret_cast<float4*>(Rprev + o) = make_float4(rr[0], rr[1], rr[2], rr[3]);
    if (Rh) h16::store_split4(Rh, Rl, o, rr);
  }
};

// part B: acc = rbar_l[m][n].  Emits zbarB_l and qbar_{l+1}.  (r_l is left intact: the weight gradient r_l^T qbar_l of the
// same layer reads it from a second stream.)
struct EpiB {
  const float* Z;    // Z_l
  const float* R;    // r_l
  float* ZB;         // out: zbarB_l
  float* QBnext;     // [M][ld]
  const float* PB;   // pbar [M][Epad]
  int ld, n_true, pre_skip, Epad, E;
  float beta;
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[4]) const {
    int64_t o = (int64_t)m * ld + n0;
    float4 z4 = *reinterpret_cast<const float4*>(Z + o);
    float4 r4 = *reinterpret_cast<const float4*>(R + o);
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
    const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
    float zb[4], qb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + j;
      if (n < n_true) {
        float s1 = softplus_d1_fast(zz[j], beta);
        bool lin = zz[j] * beta > 20.f;
        zb[j] = lin ? 0.f : beta * (1.f - s1) * rr[j] * acc[j];
        float qa = s1 * acc[j];
        qb[j] = pre_skip ? __fdiv_rn(qa, IRONB_SQRT2F) : qa;
      } else {
        zb[j] = 0.f;
        int c = n - n_true;
        qb[j] = (pre_skip && c < E) ? __fdiv_rn(__ldg(PB + (int64_t)m * Epad + c), IRONB_SQRT2F) : 0.f;
      }
    }
    *reinterpret_cast<float4*>(ZB + o) = make_float4(zb[0], zb[1], zb[2], zb[3]);
    *reinterpret_cast<float4*>(QBnext + o) = make_float4(qb[0], qb[1], qb[2], qb[3]);
  }
};

// part A: acc = ubar_l[m][k];  delta_{l-1} = sp'(z_{l-1}) * abar + zbarB_{l-1}
struct EpiA {
  const float* Zprev;
  const float* ZBprev;  // or null
  float* Dprev;         // [M][ld]
  int ld, n_true_prev, is_skip;
  float beta;
  __device__ __forceinline__ void operator()(int m, int k0, const float (&acc)[4]) const {
    int64_t o = (int64_t)m * ld + k0;
    float4 z4 = *reinterpret_cast<const float4*>(Zprev + o);
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
    float zb[4] = {0.f, 0.f, 0.f, 0.f};
    if (ZBprev) {
      float4 b4 = *reinterpret_cast<const float4*>(ZBprev + o);
      zb[0] = b4.x; zb[1] = b4.y; zb[2] = b4.z; zb[3] = b4.w;
    }
    float dd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + j;
      if (k < n_true_prev) {
        float ab = is_skip ? __fdiv_rn(acc[j], IRONB_SQRT2F) : acc[j];
        dd[j] = softplus_d1_fast(zz[j], beta) * ab + zb[j];
      } else {
        dd[j] = 0.f;
      }
    }
    *reinterpret_cast<float4*>(Dprev + o) = make_float4(dd[0], dd[1], dd[2], dd[3]);
  }
};

// n = s * Je^T p    (Je^T p = p[0:3] + sum_k 2^k (cos_k * p_sin,k - sin_k * p_cos,k))
__global__ void __launch_bounds__(256) normal_kernel(const float* __restrict__ e, const float* __restrict__ P,
                                                     int64_t M, int multires, int Epad, float scale,
                                                     float* __restrict__ grad) {
  int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* em = e + m * Epad;
  const float* pm = P + m * Epad;
  float out[3] = {pm[0], pm[1], pm[2]};
  float f = 1.f;
  int w = 3;
  for (int k = 0; k < multires; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] += f * (em[w + 3 + c] * pm[w + c] - em[w + c] * pm[w + 3 + c]);
    w += 6;
    f *= 2.f;
  }
  grad[m * 3] = out[0] * scale; grad[m * 3 + 1] = out[1] * scale; grad[m * 3 + 2] = out[2] * scale;
}

// pbar = s * Je nbar  -> written into the P buffer (which doubles as qbar_0)
__global__ void __launch_bounds__(256) pbar_kernel(const float* __restrict__ e, const float* __restrict__ nbar,
                                                   int64_t M, int multires, int Epad, float scale,
                                                   float* __restrict__ P) {
  int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* em = e + m * Epad;
  float* pm = P + m * Epad;
  float nb[3] = {nbar[m * 3] * scale, nbar[m * 3 + 1] * scale, nbar[m * 3 + 2] * scale};
  pm[0] = nb[0]; pm[1] = nb[1]; pm[2] = nb[2];
  float f = 1.f;
  int w = 3;
  for (int k = 0; k < multires; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      pm[w + c] = f * em[w + 3 + c] * nb[c];
      pm[w + 3 + c] = -f * em[w + c] * nb[c];
    }
    w += 6;
    f *= 2.f;
  }
  for (; w < Epad; ++w) pm[w] = 0.f;
}

// delta_last = [ybar/s, fbar, 0...]
__global__ void __launch_bounds__(256) dlast_kernel(const float* __restrict__ ybar, const float* __restrict__ fbar,
                                                    int64_t M, int d_out, int ld, float inv_scale,
                                                    float* __restrict__ D) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M * ld) return;
  int64_t m = i / ld;
  int n = (int)(i - m * ld);
  float v = 0.f;
  if (n == 0) v = ybar ? ybar[m] * inv_scale : 0.f;
  else if (n < d_out) v = fbar ? fbar[m * (d_out - 1) + n - 1] : 0.f;
  D[i] = v;
}

struct SdfWs {
  float* e; float* P;
  float* U[IRONB_MAX_LIN];   // U[0] == e
  float* Z[IRONB_MAX_LIN];
  float* R[IRONB_MAX_LIN];
  float* QB[IRONB_MAX_LIN];  // qbar_l, l >= 1 (qbar_0 = P); one buffer per layer: the weight gradients read them later
  float* D[IRONB_MAX_LIN];   // delta_l
  float* ZB[IRONB_MAX_LIN];  // zbarB_l
  float* wg;                 // transposed-operand scratch of the tensor-core weight gradients
  __half *Uh[IRONB_MAX_LIN], *Ul[IRONB_MAX_LIN];   // fp16x2-split copies of U_l (gemm mode 2: forward-type tensor-core operands)
  __half *Rh[IRONB_MAX_LIN], *Rl[IRONB_MAX_LIN];   // ... and of R_l
  int64_t floats;
};

int max_pad(const ironb_mlp_layout* L) {
  int mx = 0;
  for (int l = 0; l < L->n_lin; ++l) { mx = max(mx, L->in_pad[l]); mx = max(mx, L->out_pad[l]); }
  return mx;
}

// Carve the workspace.  full == false: only the PE buffer and two ping-pong activation buffers.
SdfWs carve(const ironb_mlp_layout* L, int64_t M, bool full, float* base) {
  SdfWs w;
  memset(&w, 0, sizeof(w));
  int64_t off = 0;
  auto take = [&](int64_t cols) { float* p = base ? base + off : nullptr; off += (M * cols + 63) / 64 * 64; return p; };
  // hi and lo half arrays [M][cols] each = M * cols floats in total
  auto take_h = [&](int64_t cols, __half*& hi, __half*& lo) {
    float* p = take(cols);
    hi = reinterpret_cast<__half*>(p);
    lo = p ? hi + M * cols : nullptr;
  };
  const bool h16 = gemm_mode() == 2;
  const int last = L->n_lin - 1;
  const int mp = max_pad(L);
  w.e = take(L->in_pad[0]);
  w.U[0] = w.e;
  if (!full) {
    float* a = take(mp); float* b = take(mp);
    for (int l = 1; l <= last; ++l) w.U[l] = (l & 1) ? a : b;
    if (h16) {
      take_h(L->in_pad[0], w.Uh[0], w.Ul[0]);
      __half *ah, *al, *bh, *bl;
      take_h(mp, ah, al); take_h(mp, bh, bl);
      for (int l = 1; l <= last; ++l) { w.Uh[l] = (l & 1) ? ah : bh; w.Ul[l] = (l & 1) ? al : bl; }
    }
  } else {
    w.P = take(L->in_pad[0]);
    for (int l = 1; l <= last; ++l) w.U[l] = take(L->in_pad[l]);
    for (int l = 0; l < last; ++l) w.Z[l] = take(L->out_pad[l]);
    for (int l = 0; l < last; ++l) w.R[l] = take(L->out_pad[l]);
    w.QB[0] = w.P;
    for (int l = 1; l <= last; ++l) w.QB[l] = take(L->in_pad[l]);
    for (int l = 0; l <= last; ++l) w.D[l] = take(L->out_pad[l]);
    for (int l = 0; l < last; ++l) w.ZB[l] = take(L->out_pad[l]);
    w.wg = base ? base + off : nullptr;
    off += (wgrad_scratch_floats(M, mp) + 63) / 64 * 64;
    if (h16) {
      for (int l = 0; l <= last; ++l) take_h(L->in_pad[l], w.Uh[l], w.Ul[l]);
      for (int l = 0; l < last; ++l) take_h(L->out_pad[l], w.Rh[l], w.Rl[l]);
    }
  }
  w.floats = off;
  return w;
}

// fp16x2-split weights written by the fold: W_l (rows out_pad, K = in_pad) and W_l^T (rows in_pad, K = out_pad)
inline const __half* w_hi(const ironb_mlp_layout* L, const float* packed, int l) { return reinterpret_cast<const __half*>(packed + L->off_h16[l]); }
inline const __half* wt_hi(const ironb_mlp_layout* L, const float* packed, int l) { return reinterpret_cast<const __half*>(packed + L->off_h16t[l]); }
inline int64_t w_elems(const ironb_mlp_layout* L, int l) { return (int64_t)L->out_pad[l] * L->in_pad[l]; }

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int64_t ironb_sdf_getall_workspace_bytes(const ironb_mlp_layout* lay, int64_t M, int want_grad, int save) {
  if (!lay || M < 0) return -1;
  SdfWs w = carve(lay, M, want_grad || save, nullptr);
  return w.floats * (int64_t)sizeof(float);
}

extern "C" int ironb_sdf_getall_fwd(const ironb_mlp_layout* lay, const float* packed, const float* x, int64_t M,
                                    float* y, float* feat, float* grad, int save, void* ws, int64_t ws_bytes,
                                    void* stream) {
  IRONB_REQUIRE(lay && lay->kind == 0, "sdf_getall_fwd: layout is not an SDF layout");
  IRONB_REQUIRE(M >= 0 && M < (1ll << 31), "sdf_getall_fwd: M out of range");
  if (M == 0) return IRONB_OK;
  IRONB_REQUIRE(packed && x && ws, "sdf_getall_fwd: null pointer");
  const bool full = (grad != nullptr) || save;
  SdfWs w = carve(lay, M, full, reinterpret_cast<float*>(ws));
  IRONB_REQUIRE(ws_bytes >= w.floats * (int64_t)sizeof(float), "sdf_getall_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int last = lay->n_lin - 1;
  const int Epad = lay->in_pad[0], E = lay->pe_dim;
  const int mblocks = (int)ceil_div64(M, 256);

  const bool h16m = gemm_mode() == 2;     // forward-type products on pre-split fp16x2 operands (gemm_h16.cuh)
  pe_kernel<<<mblocks, 256, 0, st>>>(x, M, lay->multires, lay->scale, Epad, w.e, h16m ? w.Uh[0] : nullptr, h16m ? w.Ul[0] : nullptr);
  IRONB_CHECK_LAUNCH("pe_kernel");

  for (int l = 0; l < last; ++l) {
    EpiFwdHidden ep;
    ep.bias = packed + lay->off_b[l];
    ep.Z = full ? w.Z[l] : nullptr;
    ep.Unext = w.U[l + 1];
    ep.R = (grad != nullptr && l == last - 1) ? w.R[l] : nullptr;
    ep.qrow = packed + lay->off_w[last];     // row 0 of W_last
    ep.e = w.e;
    ep.ld = lay->out_pad[l];
    ep.n_true = lay->out_dim[l];
    ep.pre_skip = (l + 1 == lay->skip_layer);
    ep.Epad = Epad; ep.E = E;
    ep.beta = lay->beta;
    ep.qscale = 1.f / lay->scale;
    ep.Uh = h16m ? w.Uh[l + 1] : nullptr; ep.Ul = h16m ? w.Ul[l + 1] : nullptr;
    ep.Rh = (h16m && ep.R) ? w.Rh[l] : nullptr; ep.Rl = (h16m && ep.R) ? w.Rl[l] : nullptr;
    int rc = h16m ? h16::launch_gemm_h16(w.Uh[l], w.Ul[l], lay->in_pad[l], w_hi(lay, packed, l), w_hi(lay, packed, l) + w_elems(lay, l),
                                         lay->in_pad[l], (int)M, lay->out_pad[l], lay->in_pad[l], ep, st, "sdf fwd hidden gemm (h16)")
                  : launch_gemm_nt_auto(w.U[l], lay->in_pad[l], packed + lay->off_w[l], lay->in_pad[l], (int)M,
                                        lay->out_pad[l], lay->in_pad[l], ep, st, "sdf fwd hidden gemm");
    if (rc) return rc;
  }
  if (y != nullptr || feat != nullptr) {
    EpiFwdLast ep{packed + lay->off_b[last], y, feat, lay->d_out, 1.f / lay->scale};
    int rc = h16m ? h16::launch_gemm_h16(w.Uh[last], w.Ul[last], lay->in_pad[last], w_hi(lay, packed, last),
                                         w_hi(lay, packed, last) + w_elems(lay, last), lay->in_pad[last], (int)M,
                                         lay->out_pad[last], lay->in_pad[last], ep, st, "sdf fwd last gemm (h16)")
                  : launch_gemm_nt_auto(w.U[last], lay->in_pad[last], packed + lay->off_w[last], lay->in_pad[last], (int)M,
                                        lay->out_pad[last], lay->in_pad[last], ep, st, "sdf fwd last gemm");
    if (rc) return rc;
  }
  if (grad != nullptr) {
    IRONB_CUDA(cudaMemsetAsync(w.P, 0, (size_t)M * Epad * sizeof(float), st));
    for (int l = last - 1; l >= 0; --l) {
      EpiQ ep;
      ep.Zprev = l >= 1 ? w.Z[l - 1] : nullptr;
      ep.Rprev = l >= 1 ? w.R[l - 1] : nullptr;
      ep.P = w.P;
      ep.ld = lay->in_pad[l];
      ep.n_true_prev = l >= 1 ? lay->out_dim[l - 1] : 0;
      ep.is_skip = (l == lay->skip_layer);
      ep.is_first = (l == 0);
      ep.Epad = Epad; ep.E = E;
      ep.beta = lay->beta;
      ep.Rh = (h16m && l >= 1) ? w.Rh[l - 1] : nullptr; ep.Rl = (h16m && l >= 1) ? w.Rl[l - 1] : nullptr;
      int rc = h16m ? h16::launch_gemm_h16(w.Rh[l], w.Rl[l], lay->out_pad[l], wt_hi(lay, packed, l), wt_hi(lay, packed, l) + w_elems(lay, l),
                                           lay->out_pad[l], (int)M, lay->in_pad[l], lay->out_pad[l], ep, st, "sdf grad gemm (h16)")
                    : launch_gemm_nt_auto(w.R[l], lay->out_pad[l], packed + lay->off_wt[l], lay->out_pad[l], (int)M,
                                          lay->in_pad[l], lay->out_pad[l], ep, st, "sdf grad gemm");
      if (rc) return rc;
    }
    normal_kernel<<<mblocks, 256, 0, st>>>(w.e, w.P, M, lay->multires, Epad, lay->scale, grad);
    IRONB_CHECK_LAUNCH("normal_kernel");
  }
  return IRONB_OK;
}

extern "C" int ironb_sdf_getall_bwd(const ironb_mlp_layout* lay, const float* packed, const float* x, int64_t M,
                                    const float* ybar, const float* fbar, const float* nbar, void* ws,
                                    int64_t ws_bytes, float* dpacked, void* stream, void* wgrad_stream) {
  (void)x;
  IRONB_REQUIRE(lay && lay->kind == 0, "sdf_getall_bwd: layout is not an SDF layout");
  IRONB_REQUIRE(M >= 0 && M < (1ll << 31), "sdf_getall_bwd: M out of range");
  if (M == 0) return IRONB_OK;
  IRONB_REQUIRE(packed && ws && dpacked, "sdf_getall_bwd: null pointer");
  SdfWs w = carve(lay, M, true, reinterpret_cast<float*>(ws));
  IRONB_REQUIRE(ws_bytes >= w.floats * (int64_t)sizeof(float), "sdf_getall_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  // The two back-propagation chains (part B, part A: one GEMM per layer, each waiting for the previous) run on `st`; the 17
  // weight gradients (transpose + split-K GEMM each) only CONSUME what the chains produce, so they go to `wst` as soon as their
  // operands exist and leave the chains' critical path (round 1 interleaved them on one stream: 34 launches of the step's
  // longest dependency chain).  wgrad_stream == NULL keeps everything on `st`.
  cudaStream_t wst = wgrad_stream ? as_stream(wgrad_stream) : st;
  // the overlap pays only while one GEMM does not fill the GPU (a 16,384-row layer is already 512 CTAs = 3.5 waves)
  if (M > 16384) wst = st;
  int frc;
  const int last = lay->n_lin - 1;
  const int Epad = lay->in_pad[0], E = lay->pe_dim;
  const int mblocks = (int)ceil_div64(M, 256);
  const float inv_s = 1.f / lay->scale;
  const bool has_n = nbar != nullptr;
  const bool has_yf = (ybar != nullptr) || (fbar != nullptr);
  if (!has_n && !has_yf) return IRONB_OK;

  // ---- part B: through grad ----
  if (has_n) {
    pbar_kernel<<<mblocks, 256, 0, st>>>(w.e, nbar, M, lay->multires, Epad, lay->scale, w.P);
    IRONB_CHECK_LAUNCH("pbar_kernel");
    for (int l = 0; l < last; ++l) {
      const float* qb = w.QB[l];
      // dW_l += r_l^T qbar_l   (weight-gradient stream; qbar_l is complete on st here)
      if ((frc = fork_to(st, wst))) return frc;
      int rc = launch_wgrad_auto(w.R[l], lay->out_pad[l], qb, lay->in_pad[l], (int)M, lay->out_pad[l], lay->in_pad[l],
                                 dpacked + lay->off_w[l], lay->in_pad[l], w.wg, wst, "sdf bwd B wgrad");
      if (rc) return rc;
      EpiB ep;
      ep.Z = w.Z[l]; ep.R = w.R[l]; ep.ZB = w.ZB[l];
      ep.QBnext = w.QB[l + 1];
      ep.PB = w.P;
      ep.ld = lay->out_pad[l];
      ep.n_true = lay->out_dim[l];
      ep.pre_skip = (l + 1 == lay->skip_layer);
      ep.Epad = Epad; ep.E = E;
      ep.beta = lay->beta;
      rc = launch_gemm_nt_auto(qb, lay->in_pad[l], packed + lay->off_w[l], lay->in_pad[l], (int)M, lay->out_pad[l],
                          lay->in_pad[l], ep, st, "sdf bwd B gemm");
      if (rc) return rc;
    }
    // dW_last[0,:] += sum_m qbar_last / s
    if ((frc = fork_to(st, wst))) return frc;
    int rc = launch_colsum(w.QB[last], lay->in_pad[last], (int)M, lay->in_dim[last], inv_s, dpacked + lay->off_w[last], wst,
                           "sdf bwd B colsum");
    if (rc) return rc;
  }

  // ---- part A: ordinary back-propagation ----
  int lstart = last;
  const float* D = nullptr;
  if (has_yf) {
    int64_t tot = M * lay->out_pad[last];
    dlast_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(ybar, fbar, M, lay->d_out, lay->out_pad[last], inv_s,
                                                                w.D[last]);
    IRONB_CHECK_LAUNCH("dlast_kernel");
    D = w.D[last];
  } else {
    // delta_last == 0: start one layer down, where delta equals zbarB
    lstart = last - 1;
    D = w.ZB[lstart];
  }
  for (int l = lstart; l >= 0; --l) {
    // dW_l += delta_l^T u_l, and db_l += column sums of delta_l (taken from the transposed tiles of the same launch)
    if ((frc = fork_to(st, wst))) return frc;
    int rc = launch_wgrad_auto(D, lay->out_pad[l], w.U[l], lay->in_pad[l], (int)M, lay->out_pad[l], lay->in_pad[l],
                               dpacked + lay->off_w[l], lay->in_pad[l], w.wg, wst, "sdf bwd A wgrad",
                               dpacked + lay->off_b[l], lay->out_dim[l], 1.f);
    if (rc) return rc;
    if (l == 0) break;
    EpiA ep;
    ep.Zprev = w.Z[l - 1];
    ep.ZBprev = has_n ? w.ZB[l - 1] : nullptr;
    ep.Dprev = w.D[l - 1];
    ep.ld = lay->in_pad[l];
    ep.n_true_prev = lay->out_dim[l - 1];
    ep.is_skip = (l == lay->skip_layer);
    ep.beta = lay->beta;
    rc = launch_gemm_nt_auto(D, lay->out_pad[l], packed + lay->off_wt[l], lay->out_pad[l], (int)M, lay->in_pad[l],
                        lay->out_pad[l], ep, st, "sdf bwd A gemm");
    if (rc) return rc;
    D = w.D[l - 1];
  }
  if ((frc = fork_to(wst, st))) return frc;      // join: dpacked is complete on st
  return IRONB_OK;
}

namespace ironb { IRONB_DEFINE_HANG_SETTER(hang_set_sdf) }
