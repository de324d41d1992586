// RenderingNetwork (the three material MLPs of the 'ggx' configuration) forward and backward.
// models/fields.py:203-239: input = cat(PE(points), [PE(view_dirs)], [normals], features) -> (Linear, ReLU) x n_layers
// -> Linear -> output_scale * (x + output_bias) -> [squeeze_out_scale * sigmoid].  At most one skip layer s (:226-227: layer s
// reads cat(h, input) / sqrt 2; the stage-1 colour network has n_layers = 8, skip_in = [4]): the layer before it writes its H
// outputs, scaled, into the first H columns of a row of width in_pad[s]; skip_fill_kernel appends the scaled input.
// One fp32 tile GEMM per layer (gemm.cuh); ReLU / output transform fused into the epilogues, the input
// concatenation + positional encodings and their backward are two small elementwise kernels.
#include "gemm_h16.cuh"

namespace ironb {
namespace {

struct CatPlan {
  int ep, ev;           // encoded widths of points / view dirs (0 = block absent)
  int off_v, off_n, off_f;
  int has_v, has_n;
  int total;
};

__host__ __device__ inline CatPlan make_plan(const ironb_matnet_cfg& c) {
  CatPlan p;
  p.ep = c.multires > 0 ? 3 * (1 + 2 * c.multires) : 3;
  p.has_v = (c.mode == 0 || c.mode == 2);
  p.has_n = (c.mode == 0 || c.mode == 1);
  p.ev = p.has_v ? (c.multires_view > 0 ? 3 * (1 + 2 * c.multires_view) : 3) : 0;
  p.off_v = p.ep;
  p.off_n = p.off_v + p.ev;
  p.off_f = p.off_n + (p.has_n ? 3 : 0);
  p.total = p.off_f + c.d_feature;
  return p;
}

// value of column j (< 3*(1+2L)) of the encoding of a 3-vector
__device__ __forceinline__ float pe_col(const float* __restrict__ x3, int j) {
  if (j < 3) return x3[j];
  int k = (j - 3) / 6, r = (j - 3) % 6;
  float arg = x3[r % 3] * (float)(1 << k);
  return r < 3 ? sinf(arg) : cosf(arg);
}

__global__ void __launch_bounds__(256) assemble_kernel(ironb_matnet_cfg cfg, CatPlan pl, const float* __restrict__ pts,
                                                       const float* __restrict__ nrm, const float* __restrict__ view,
                                                       const float* __restrict__ feats, int64_t M, int ld,
                                                       float* __restrict__ U0, __half* __restrict__ U0h,
                                                       __half* __restrict__ U0l) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M * ld) return;
  int64_t m = i / ld;
  int j = (int)(i - m * ld);
  float v = 0.f;
  if (j < pl.off_v) v = pe_col(pts + m * 3, j);
  else if (j < pl.off_n) v = pe_col(view + m * 3, j - pl.off_v);
  else if (j < pl.off_f) v = nrm[m * 3 + (j - pl.off_n)];
  else if (j < pl.total) v = __ldg(feats + m * cfg.d_feature + (j - pl.off_f));
  U0[i] = v;
  if (U0h != nullptr) {        // fp16x2-split copy: the first layer's tensor-core operand (gemm_h16.cuh)
    const __half h = __float2half_rn(v);
    U0h[i] = h;
    U0l[i] = __float2half_rn((v - __half2float(h)) * 2048.f);
  }
}

// backward of one encoded 3-vector block: dx = de[0:3] + sum_k 2^k (cos_k de_sin,k - sin_k de_cos,k)
__device__ __forceinline__ void pe_bwd(const float* __restrict__ u, const float* __restrict__ du, int width, float out[3]) {
  out[0] = du[0]; out[1] = du[1]; out[2] = du[2];
  float f = 1.f;
  for (int w = 3; w + 6 <= width; w += 6) {
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] += f * (u[w + 3 + c] * du[w + c] - u[w + c] * du[w + 3 + c]);
    f *= 2.f;
  }
}

__global__ void __launch_bounds__(256) disassemble_kernel(ironb_matnet_cfg cfg, CatPlan pl, const float* __restrict__ U0,
                                                          const float* __restrict__ dcat, int64_t M, int ld,
                                                          float* __restrict__ d_pts, float* __restrict__ d_nrm,
                                                          float* __restrict__ d_view, float* __restrict__ d_feats) {
  // one warp per point: lanes stride the feature block, lane 0/1/2 handle the small blocks
  int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (m >= M) return;
  const float* u = U0 + m * ld;
  const float* du = dcat + m * ld;
  if (d_feats)
    for (int j = lane; j < cfg.d_feature; j += 32) d_feats[m * cfg.d_feature + j] = du[pl.off_f + j];
  if (lane == 0 && d_pts) {
    float o[3];
    pe_bwd(u, du, pl.ep, o);
    d_pts[m * 3] = o[0]; d_pts[m * 3 + 1] = o[1]; d_pts[m * 3 + 2] = o[2];
  }
  if (lane == 1 && d_view) {
    float o[3] = {0.f, 0.f, 0.f};
    if (pl.has_v) pe_bwd(u + pl.off_v, du + pl.off_v, pl.ev, o);
    d_view[m * 3] = o[0]; d_view[m * 3 + 1] = o[1]; d_view[m * 3 + 2] = o[2];
  }
  if (lane == 2 && d_nrm) {
    float o[3] = {0.f, 0.f, 0.f};
    if (pl.has_n) { o[0] = du[pl.off_n]; o[1] = du[pl.off_n + 1]; o[2] = du[pl.off_n + 2]; }
    d_nrm[m * 3] = o[0]; d_nrm[m * 3 + 1] = o[1]; d_nrm[m * 3 + 2] = o[2];
  }
}

struct EpiRelu {
  const float* bias;
  float* Unext;
  int ld, n_true;
  __half *Uh, *Ul;      // fp16x2-split copy of Unext or null
  float scale;          // 1, or 1/sqrt(2) for the layer that feeds the skip layer
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[4]) const {
    float u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + j;
      u[j] = n < n_true ? fmaxf(acc[j] + __ldg(bias + n), 0.f) * scale : 0.f;
    }
    *reinterpret_cast<float4*>(Unext + (int64_t)m * ld + n0) = make_float4(u[0], u[1], u[2], u[3]);
    if (Uh) h16::store_split4(Uh, Ul, (int64_t)m * ld + n0, u);
  }
};

struct EpiMatOut {
  const float* bias;
  float* out;   // [M][d_out]
  int d_out, squeeze;
  float out_bias, out_scale, squeeze_scale;
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[4]) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + j;
      if (n >= d_out) continue;
      float t = out_scale * (acc[j] + __ldg(bias + n) + out_bias);
      if (squeeze) t = squeeze_scale * __fdiv_rn(1.f, 1.f + expf(-t));
      out[(int64_t)m * d_out + n] = t;
    }
  }
};

// Columns [H, ld) of the skip layer's input row: the assembled network input / sqrt 2, zero padding after it.
__global__ void __launch_bounds__(256) skip_fill_kernel(const float* __restrict__ U0, int ld0, int in0, int64_t M, int H, int ld,
                                                        float* __restrict__ Us, __half* __restrict__ Ush,
                                                        __half* __restrict__ Usl) {
  const int w = ld - H;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M * w) return;
  const int64_t m = i / w;
  const int j = (int)(i - m * w);
  const float v = j < in0 ? U0[m * ld0 + j] * 0.70710678118654752f : 0.f;
  const int64_t o = m * ld + H + j;
  Us[o] = v;
  if (Ush != nullptr) {
    const __half h = __float2half_rn(v);
    Ush[o] = h;
    Usl[o] = __float2half_rn((v - __half2float(h)) * 2048.f);
  }
}

// The dgrad out of the skip layer: columns [0, H) are the gradient of the previous layer's (scaled) ReLU output, columns
// [H, ld) the gradient that reaches the network input through the skip connection (kept in S for the input dgrad).
struct EpiReluBwdSkip {
  const float* Uthis;   // U_s [M][ld]: relu(h) / sqrt 2 in the first H columns
  float* Dprev;         // D_{s-1} [M][ld]
  float* S;             // [M][ld0] or null
  int ld, H, ld0;
  __device__ __forceinline__ void operator()(int m, int k0, const float (&acc)[4]) const {
    const int64_t o = (int64_t)m * ld + k0;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (k0 < H) {
      const float4 u4 = *reinterpret_cast<const float4*>(Uthis + o);
      const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = uu[j] > 0.f ? acc[j] * 0.70710678118654752f : 0.f;
    } else if (S != nullptr && k0 - H < ld0) {
      *reinterpret_cast<float4*>(S + (int64_t)m * ld0 + (k0 - H)) =
          make_float4(acc[0] * 0.70710678118654752f, acc[1] * 0.70710678118654752f, acc[2] * 0.70710678118654752f,
                      acc[3] * 0.70710678118654752f);
    }
    *reinterpret_cast<float4*>(Dprev + o) = make_float4(d[0], d[1], d[2], d[3]);
  }
};

// delta_{l-1} = (u_l > 0) ? ubar : 0     (ReLU backward on the stored post-activation)
struct EpiReluBwd {
  const float* Uthis;   // U_l [M][ld]  (post-ReLU output of layer l-1)
  float* Dprev;
  int ld, n_true_prev;
  __device__ __forceinline__ void operator()(int m, int k0, const float (&acc)[4]) const {
    int64_t o = (int64_t)m * ld + k0;
    float4 u4 = *reinterpret_cast<const float4*>(Uthis + o);
    const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
    float d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = (k0 + j < n_true_prev && uu[j] > 0.f) ? acc[j] : 0.f;
    *reinterpret_cast<float4*>(Dprev + o) = make_float4(d[0], d[1], d[2], d[3]);
  }
};

struct EpiPlain {
  float* C;
  int ld;
  const float* add;     // [M][ld] or null: the skip connection's share of the input gradient
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[4]) const {
    const int64_t o = (int64_t)m * ld + n0;
    float4 v = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (add != nullptr) {
      const float4 a = *reinterpret_cast<const float4*>(add + o);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    *reinterpret_cast<float4*>(C + o) = v;
  }
};

__global__ void __launch_bounds__(256) mat_dlast_kernel(const float* __restrict__ out, const float* __restrict__ gout,
                                                        int64_t M, int d_out, int ld, int squeeze, float out_scale,
                                                        float squeeze_scale, float* __restrict__ D) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M * ld) return;
  int64_t m = i / ld;
  int n = (int)(i - m * ld);
  float v = 0.f;
  if (n < d_out) {
    float g = gout[m * d_out + n];
    if (squeeze) {
      float o = out[m * d_out + n];
      g = g * o * (1.f - o / squeeze_scale);
    }
    v = g * out_scale;
  }
  D[i] = v;
}

struct MatWs {
  float* U[IRONB_MAX_LIN];
  float* D[IRONB_MAX_LIN + 1];   // delta_l per layer (D[l] = gradient w.r.t. the output of layer l; D[n_lin] unused) + the input gradient
  float* wg;
  float* S;                      // skip connection's share of the input gradient [M][in_pad0] (skip nets only)
  __half *Uh[IRONB_MAX_LIN], *Ul[IRONB_MAX_LIN];   // fp16x2-split copies of U_l (gemm mode 2)
  int64_t floats;
};

MatWs carve_mat(const ironb_mlp_layout* L, int64_t M, float* base) {
  MatWs w;
  memset(&w, 0, sizeof(w));
  int64_t off = 0;
  auto take = [&](int64_t cols) { float* p = base ? base + off : nullptr; off += (M * cols + 63) / 64 * 64; return p; };
  int mp = 0;
  for (int l = 0; l < L->n_lin; ++l) { mp = max(mp, L->in_pad[l]); mp = max(mp, L->out_pad[l]); }
  for (int l = 0; l < L->n_lin; ++l) w.U[l] = take(L->in_pad[l]);
  for (int l = 0; l < L->n_lin; ++l) w.D[l] = take(L->out_pad[l]);
  w.D[L->n_lin] = take(L->in_pad[0]);                 // d loss / d (assembled input)
  if (L->skip_layer >= 1) w.S = take(L->in_pad[0]);
  w.wg = base ? base + off : nullptr;
  off += (wgrad_scratch_floats(M, mp) + 63) / 64 * 64;
  if (gemm_mode() == 2) {
    for (int l = 0; l < L->n_lin; ++l) {
      float* p = take(L->in_pad[l]);
      w.Uh[l] = reinterpret_cast<__half*>(p);
      w.Ul[l] = p ? w.Uh[l] + M * L->in_pad[l] : nullptr;
    }
  }
  w.floats = off;
  return w;
}

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int ironb_matnet_in_dim(const ironb_matnet_cfg* cfg) {
  if (!cfg) return -1;
  return make_plan(*cfg).total;
}

extern "C" int64_t ironb_matnet_workspace_bytes(const ironb_mlp_layout* lay, int64_t M) {
  if (!lay || M < 0) return -1;
  return carve_mat(lay, M, nullptr).floats * (int64_t)sizeof(float);
}

extern "C" int ironb_matnet_fwd(const ironb_mlp_layout* lay, const ironb_matnet_cfg* cfg, const float* packed,
                                const float* points, const float* normals, const float* view_dirs,
                                const float* feats, int64_t M, float* out, void* ws, int64_t ws_bytes,
                                void* stream) {
  IRONB_REQUIRE(lay && cfg && lay->kind == 1, "matnet_fwd: bad layout");
  IRONB_REQUIRE(M >= 0 && M < (1ll << 31), "matnet_fwd: M out of range");
  if (M == 0) return IRONB_OK;
  CatPlan pl = make_plan(*cfg);
  IRONB_REQUIRE(pl.total == lay->in_dim[0], "matnet_fwd: cfg input width %d != layout fan-in %d", pl.total, lay->in_dim[0]);
  IRONB_REQUIRE(packed && points && feats && out && ws, "matnet_fwd: null pointer");
  IRONB_REQUIRE(!pl.has_n || normals, "matnet_fwd: mode needs normals");
  IRONB_REQUIRE(!pl.has_v || view_dirs, "matnet_fwd: mode needs view_dirs");
  MatWs w = carve_mat(lay, M, reinterpret_cast<float*>(ws));
  IRONB_REQUIRE(ws_bytes >= w.floats * (int64_t)sizeof(float), "matnet_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int last = lay->n_lin - 1;
  int64_t tot = M * lay->in_pad[0];
  const bool h16m = gemm_mode() == 2;     // forward products on pre-split fp16x2 operands (gemm_h16.cuh)
  assemble_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(*cfg, pl, points, normals, view_dirs, feats, M,
                                                                 lay->in_pad[0], w.U[0], h16m ? w.Uh[0] : nullptr,
                                                                 h16m ? w.Ul[0] : nullptr);
  IRONB_CHECK_LAUNCH("assemble_kernel");
  auto whi = [&](int l) { return reinterpret_cast<const __half*>(packed + lay->off_h16[l]); };
  auto wlo = [&](int l) { return whi(l) + (int64_t)lay->out_pad[l] * lay->in_pad[l]; };
  const int skip = lay->skip_layer;
  for (int l = 0; l < last; ++l) {
    const bool feeds_skip = (l + 1 == skip);          // writes H scaled columns of the wider skip-layer row
    const int N = feeds_skip ? round_up(lay->out_dim[l], 8) : lay->out_pad[l];
    EpiRelu ep{packed + lay->off_b[l], w.U[l + 1], lay->out_pad[l], lay->out_dim[l], h16m ? w.Uh[l + 1] : nullptr,
               h16m ? w.Ul[l + 1] : nullptr, feeds_skip ? 0.70710678118654752f : 1.f};
    int rc = h16m ? h16::launch_gemm_h16(w.Uh[l], w.Ul[l], lay->in_pad[l], whi(l), wlo(l), lay->in_pad[l], (int)M, N,
                                         lay->in_pad[l], ep, st, "matnet fwd gemm (h16)")
                  : launch_gemm_nt_auto(w.U[l], lay->in_pad[l], packed + lay->off_w[l], lay->in_pad[l], (int)M, N,
                                        lay->in_pad[l], ep, st, "matnet fwd gemm");
    if (rc) return rc;
    if (feeds_skip) {
      const int H = round_up(lay->out_dim[l], 8), ld = lay->in_pad[skip];
      skip_fill_kernel<<<(unsigned)ceil_div64(M * (ld - H), 256), 256, 0, st>>>(w.U[0], lay->in_pad[0], lay->in_dim[0], M, H, ld,
                                                                               w.U[skip], h16m ? w.Uh[skip] : nullptr,
                                                                               h16m ? w.Ul[skip] : nullptr);
      IRONB_CHECK_LAUNCH("skip_fill_kernel");
    }
  }
  EpiMatOut ep{packed + lay->off_b[last], out, lay->d_out, cfg->squeeze, cfg->out_bias, cfg->out_scale,
               cfg->squeeze_scale};
  return h16m ? h16::launch_gemm_h16(w.Uh[last], w.Ul[last], lay->in_pad[last], whi(last), wlo(last), lay->in_pad[last], (int)M,
                                     lay->out_pad[last], lay->in_pad[last], ep, st, "matnet fwd out gemm (h16)")
              : launch_gemm_nt_auto(w.U[last], lay->in_pad[last], packed + lay->off_w[last], lay->in_pad[last], (int)M,
                                    lay->out_pad[last], lay->in_pad[last], ep, st, "matnet fwd out gemm");
}

extern "C" int ironb_matnet_bwd(const ironb_mlp_layout* lay, const ironb_matnet_cfg* cfg, const float* packed,
                                int64_t M, const float* out, const float* gout, void* ws, int64_t ws_bytes,
                                float* dpacked, float* d_points, float* d_normals, float* d_view, float* d_feats,
                                void* stream, void* wgrad_stream) {
  IRONB_REQUIRE(lay && cfg && lay->kind == 1, "matnet_bwd: bad layout");
  IRONB_REQUIRE(M >= 0 && M < (1ll << 31), "matnet_bwd: M out of range");
  if (M == 0) return IRONB_OK;
  IRONB_REQUIRE(packed && out && gout && ws && dpacked, "matnet_bwd: null pointer");
  CatPlan pl = make_plan(*cfg);
  MatWs w = carve_mat(lay, M, reinterpret_cast<float*>(ws));
  IRONB_REQUIRE(ws_bytes >= w.floats * (int64_t)sizeof(float), "matnet_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  cudaStream_t wst = wgrad_stream ? as_stream(wgrad_stream) : st;    // weight gradients off the dgrad chain (see sdf.cu)
  // the overlap pays only while one GEMM does not fill the GPU (a 16,384-row layer is already 512 CTAs = 3.5 waves)
  if (M > 16384) wst = st;
  int frc;
  const int last = lay->n_lin - 1;
  int64_t tot = M * lay->out_pad[last];
  mat_dlast_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(out, gout, M, lay->d_out, lay->out_pad[last],
                                                                  cfg->squeeze, cfg->out_scale, cfg->squeeze_scale,
                                                                  w.D[last]);
  IRONB_CHECK_LAUNCH("mat_dlast_kernel");
  const bool need_in = d_points || d_normals || d_view || d_feats;
  const int skip = lay->skip_layer;
  for (int l = last; l >= 0; --l) {
    const float* D = w.D[l];
    // true output width (rounded to 8): the layer feeding a skip layer has out_pad = the skip row width, its gradient columns
    // beyond H are zero
    const int No = round_up(lay->out_dim[l], 8);
    if ((frc = fork_to(st, wst))) return frc;
    int rc = launch_wgrad_auto(D, lay->out_pad[l], w.U[l], lay->in_pad[l], (int)M, No, lay->in_pad[l],
                               dpacked + lay->off_w[l], lay->in_pad[l], w.wg, wst, "matnet wgrad (+ bias grad)",
                               dpacked + lay->off_b[l], lay->out_dim[l], 1.f);
    if (rc) return rc;
    if (l > 0 && l == skip) {
      EpiReluBwdSkip ep{w.U[l], w.D[l - 1], need_in ? w.S : nullptr, lay->in_pad[l], round_up(lay->out_dim[l - 1], 8), lay->in_pad[0]};
      rc = launch_gemm_nt_auto(D, lay->out_pad[l], packed + lay->off_wt[l], lay->out_pad[l], (int)M, lay->in_pad[l], No, ep, st,
                               "matnet dgrad (skip layer)");
      if (rc) return rc;
    } else if (l > 0) {
      EpiReluBwd ep{w.U[l], w.D[l - 1], lay->in_pad[l], lay->out_dim[l - 1]};
      rc = launch_gemm_nt_auto(D, lay->out_pad[l], packed + lay->off_wt[l], lay->out_pad[l], (int)M, lay->in_pad[l],
                          No, ep, st, "matnet dgrad");
      if (rc) return rc;
    } else if (need_in) {
      float* Dn = w.D[lay->n_lin];
      EpiPlain ep{Dn, lay->in_pad[0], skip >= 1 ? w.S : nullptr};
      rc = launch_gemm_nt_auto(D, lay->out_pad[0], packed + lay->off_wt[0], lay->out_pad[0], (int)M, lay->in_pad[0],
                          No, ep, st, "matnet input dgrad");
      if (rc) return rc;
      int64_t thr = M * 32;
      disassemble_kernel<<<(unsigned)ceil_div64(thr, 256), 256, 0, st>>>(*cfg, pl, w.U[0], Dn, M, lay->in_pad[0],
                                                                        d_points, d_normals, d_view, d_feats);
      IRONB_CHECK_LAUNCH("disassemble_kernel");
    }
  }
  if ((frc = fork_to(wst, st))) return frc;      // join: dpacked is complete on st
  return IRONB_OK;
}

namespace ironb { IRONB_DEFINE_HANG_SETTER(hang_set_matnet) }
