// Persistent ray-queue tracer: RayTracer.forward of the reference (models/raytracer.py:45-220) as four
// persistent kernels with no host round trip.
//
//   phase 0  sphere tracing      (:105-140)  one slot = one ray, <= iters+1 SDF evaluations
//   phase 1  dense sampler       (:142-197)  one 32-slot group = one unfinished ray, evaluated 32 samples at a
//                                            time in order until the first negative sample (the reference
//                                            evaluates all n_steps and takes the first negative: same result)
//   phase 2  bisection, pass A   (:199-214)  every root ray halves until ITS interval is <= 2*thr, records k_i
//   phase 3  bisection, pass B   (:204-219)  every root ray runs the remaining k_max - k_i halvings (the
//                                            reference loop runs all rays while ANY ray works) + the final eval
//
// A CTA owns TM ray slots.  Each round it evaluates the whole SDF MLP for its TM points with the
// activations resident in shared memory (k-major [H][TM]) and the folded weights streamed from L2 with
// cp.async double buffering, then advances every slot's state machine; finished slots are refilled from a
// global atomic queue (warp ballot + popc gives each empty slot its ray), so lanes never idle on
// divergent march lengths.  Kernel boundaries are the only grid-wide syncs (they carry the sampler /
// root lists and k_max).  MLP arithmetic: fp32 FFMA with fp32 accumulation.
#include "common.cuh"

namespace ironb {
namespace {

constexpr int KC = 8;          // weight rows (k) per cp.async stage
constexpr int TPB = 256;

struct TraceNet {
  const float* wt[IRONB_MAX_LIN];   // W_l^T [in_pad][H]
  const float* b[IRONB_MAX_LIN];
  int in_pad[IRONB_MAX_LIN];
  int n_true[IRONB_MAX_LIN];
  const float* w_last;              // row 0 of W_last [H]
  const float* b_last;
  int n_hidden;                     // number of layers evaluated as tiles (n_lin - 1)
  int skip_layer, E, Epad, multires;
  float scale, beta;
};

struct TraceArgs {
  const float *ray_o, *ray_d, *min_dis, *max_dis;
  const uint8_t* work_mask;
  int N;
  float thr, two_thr;
  int iters, n_steps;
  const float* linspace;
  uint8_t* conv;
  float *points, *sdf, *dist;
  unsigned long long* stats;   // may be null
  // workspace
  int* counters;               // [0] next ray  [1] n_unf  [2] next unf  [3] n_root  [4] next root A  [5] k_max  [6] next root B
  int* unf_list;
  int* root_ray;
  float* root_lo;
  float* root_hi;
  int* root_k;                 // -1: never worked
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

template <int NJ, int TM>
struct Smem {
  static constexpr int H = NJ * 128;
  static constexpr int TMS = TM + 4;
  static constexpr int EPAD_MAX = 64;
  static constexpr size_t act_floats = (size_t)H * TMS;
  static constexpr size_t e_floats = (size_t)EPAD_MAX * TMS;
  static constexpr size_t w_floats = (size_t)2 * KC * H;
  static constexpr size_t red_floats = (size_t)TPB;
  static constexpr size_t bytes = (act_floats + e_floats + w_floats + red_floats) * sizeof(float);
};

// Evaluate the SDF for the TM points whose encodings sit in ebuf; result f[TM].
template <int NJ, int TM>
__device__ __forceinline__ void mlp_eval(const TraceNet& net, float* __restrict__ act, const float* __restrict__ ebuf,
                                         float* __restrict__ wbuf, float* __restrict__ red, float* __restrict__ fout) {
  constexpr int H = NJ * 128;
  constexpr int TMS = TM + 4;
  constexpr int RM = TM / 8;            // rows per thread (8 warps = 8 row groups)
  const int tid = threadIdx.x, lane = tid & 31, tr = tid >> 5;

  for (int l = 0; l < net.n_hidden; ++l) {
    const float* in = (l == 0) ? ebuf : act;
    const int K = net.in_pad[l];
    const float* __restrict__ WT = net.wt[l];
    float acc[RM][NJ * 4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < NJ * 4; ++j) acc[i][j] = 0.f;

    const int nchunk = K / KC;
    auto issue = [&](int c) {
      const float* src = WT + (size_t)c * KC * H;
      float* dst = wbuf + (size_t)(c & 1) * KC * H;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {   // KC*H/4 float4 = NJ*256
        int f = tid + i * TPB;
        cp_async16(dst + f * 4, src + f * 4);
      }
      cp_async_commit();
    };
    issue(0);
    for (int c = 0; c < nchunk; ++c) {
      cp_async_wait_all();
      __syncthreads();
      if (c + 1 < nchunk) issue(c + 1);
      const float* wb = wbuf + (size_t)(c & 1) * KC * H;
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const float* arow = in + (size_t)(c * KC + kk) * TMS + tr * RM;
        float a[RM];
#pragma unroll
        for (int q = 0; q < RM / 4; ++q) {
          float4 v = *reinterpret_cast<const float4*>(arow + q * 4);
          a[q * 4] = v.x; a[q * 4 + 1] = v.y; a[q * 4 + 2] = v.z; a[q * 4 + 3] = v.w;
        }
        float w[NJ * 4];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          float4 v = *reinterpret_cast<const float4*>(wb + kk * H + j * 128 + lane * 4);
          w[j * 4] = v.x; w[j * 4 + 1] = v.y; w[j * 4 + 2] = v.z; w[j * 4 + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < NJ * 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
    }
    __syncthreads();   // everyone is done reading `in` (== act for l >= 1) and wbuf
    // epilogue: bias + softplus (+ skip concat), written k-major for the next layer
    const bool pre_skip = (l + 1 == net.skip_layer);
    const int n_true = net.n_true[l];
    const float* __restrict__ bias = net.b[l];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int n = j * 128 + lane * 4 + c;
        float v[RM];
        if (n < n_true) {
          const float bn = __ldg(bias + n);
#pragma unroll
          for (int i = 0; i < RM; ++i) {
            float a = softplus_beta(acc[i][j * 4 + c] + bn, net.beta);
            v[i] = pre_skip ? __fdiv_rn(a, IRONB_SQRT2F) : a;
          }
        } else {
          const int ce = n - n_true;
#pragma unroll
          for (int i = 0; i < RM; ++i)
            v[i] = (pre_skip && ce < net.E) ? __fdiv_rn(ebuf[(size_t)ce * TMS + tr * RM + i], IRONB_SQRT2F) : 0.f;
        }
        float* dst = act + (size_t)n * TMS + tr * RM;
#pragma unroll
        for (int q = 0; q < RM / 4; ++q)
          *reinterpret_cast<float4*>(dst + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      }
    __syncthreads();
  }
  // last layer, row 0 only: f = (w_last . a + b_last) / scale
  {
    constexpr int PARTS = TPB / TM;
    const int row = tid % TM, part = tid / TM;
    float s = 0.f;
    for (int k = part; k < H; k += PARTS) s = fmaf(act[(size_t)k * TMS + row], __ldg(net.w_last + k), s);
    red[tid] = s;
    __syncthreads();
    if (tid < TM) {
      float t = 0.f;
#pragma unroll
      for (int p = 0; p < PARTS; ++p) t += red[p * TM + tid];
      fout[tid] = __fdiv_rn(t + __ldg(net.b_last), net.scale);
    }
    __syncthreads();
  }
}

// positional encoding of the TM slot positions into ebuf (k-major)
template <int TM>
__device__ __forceinline__ void encode(const TraceNet& net, const float* __restrict__ sx /*[3][TM]*/,
                                       float* __restrict__ ebuf) {
  constexpr int TMS = TM + 4;
  const int tid = threadIdx.x;
  for (int i = tid; i < TM * 3; i += TPB) {
    int c = i / TM, s = i % TM;
    ebuf[(size_t)c * TMS + s] = sx[c * TM + s] * net.scale;
  }
  const int tot = TM * 3 * net.multires;
  for (int i = tid; i < tot; i += TPB) {
    int s = i % TM, r = i / TM;
    int k = r / 3, c = r % 3;
    float arg = (sx[c * TM + s] * net.scale) * (float)(1 << k);
    float sn, cs;
    sincosf(arg, &sn, &cs);
    ebuf[(size_t)(3 + 6 * k + c) * TMS + s] = sn;
    ebuf[(size_t)(6 + 6 * k + c) * TMS + s] = cs;
  }
}

template <int TM>
struct Slots {
  int ray[TM];          // >= 0: active (ray id or root index)
  float o[3][TM], d[3][TM], x[3][TM];
  float a[TM], b[TM], c[TM];   // phase 0: t, t_max, -   phase 2/3: lo, hi, mid
  int k[TM];
  int flags[TM];        // phase 0: bit0 work, bit1 unfinished
  float f[TM];
  float ts[TM];
  // sampler groups
  int g_ray[TM / 32];
  int g_chunk[TM / 32];
  float g_smin[TM / 32], g_smax[TM / 32], g_prev_t[TM / 32], g_prev_f[TM / 32];
  float g_o[3][TM / 32], g_d[3][TM / 32];
  int exhausted;
  int kmax;
};

template <int NJ, int TM, int MODE>
__global__ void __launch_bounds__(TPB, 1) trace_kernel(TraceNet net, TraceArgs A) {
  using SM = Smem<NJ, TM>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* act = reinterpret_cast<float*>(smem_raw);
  float* ebuf = act + SM::act_floats;
  float* wbuf = ebuf + SM::e_floats;
  float* red = wbuf + SM::w_floats;
  __shared__ Slots<TM> S;
  constexpr int TMS = TM + 4;
  constexpr int G = TM / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // one-time init
  for (int i = tid; i < TM; i += TPB) { S.ray[i] = -1; S.x[0][i] = S.x[1][i] = S.x[2][i] = 0.f; }
  if (tid < G) S.g_ray[tid] = -1;
  if (tid == 0) {
    S.exhausted = 0;
    S.kmax = (MODE == 3) ? A.counters[5] : 0;
    if (MODE == 3 && blockIdx.x == 0 && A.stats) {   // lists are final here: publish their sizes
      atomicAdd(A.stats + 3, (unsigned long long)A.counters[1]);
      atomicAdd(A.stats + 4, (unsigned long long)A.counters[3]);
      atomicMax(A.stats + 5, (unsigned long long)A.counters[5]);
    }
  }
  for (int i = tid; i < SM::EPAD_MAX * TMS; i += TPB) ebuf[i] = 0.f;
  __syncthreads();

  const int n_items = (MODE == 0) ? A.N : (MODE == 1 ? A.counters[1] : A.counters[3]);
  int* next_ctr = A.counters + (MODE == 0 ? 0 : (MODE == 1 ? 2 : (MODE == 2 ? 4 : 6)));

  while (true) {
    // ------------------------------------------------------------ refill
    if (MODE == 1) {
      if (tid < G && S.g_ray[tid] < 0 && !S.exhausted) {
        int i = atomicAdd(next_ctr, 1);
        if (i < n_items) {
          int r = A.unf_list[i];
          float t = A.dist[r], f = A.sdf[r];
          bool outside = f > 0.f;                               // raytracer.py:59-65
          S.g_ray[tid] = r;
          S.g_chunk[tid] = 0;
          S.g_smin[tid] = outside ? t : A.min_dis[r];
          S.g_smax[tid] = outside ? A.max_dis[r] : t;
          S.g_prev_t[tid] = 0.f; S.g_prev_f[tid] = 0.f;
          for (int c = 0; c < 3; ++c) { S.g_o[c][tid] = A.ray_o[(size_t)r * 3 + c]; S.g_d[c][tid] = A.ray_d[(size_t)r * 3 + c]; }
        } else {
          S.exhausted = 1;
        }
      }
      __syncthreads();
      // sample positions of the current 32-sample chunk                  :144-150
      if (tid < TM) {
        int g = tid >> 5;
        if (S.g_ray[g] >= 0) {
          int j = S.g_chunk[g] * 32 + (tid & 31);
          int jc = min(j, A.n_steps - 1);
          float span = __fsub_rn(S.g_smax[g], S.g_smin[g]);
          float ts = __fadd_rn(S.g_smin[g], __fmul_rn(__ldg(A.linspace + jc), span));
          S.ts[tid] = ts;
#pragma unroll
          for (int c = 0; c < 3; ++c) S.x[c][tid] = __fadd_rn(S.g_o[c][g], __fmul_rn(S.g_d[c][g], ts));
          S.ray[tid] = (j < A.n_steps) ? S.g_ray[g] : -1;
        } else {
          S.ray[tid] = -1;
        }
      }
    } else {
      if (warp == 0) {
        for (int base = 0; base < TM; base += 32) {
          int s = base + lane;
          bool empty = S.ray[s] < 0;
          unsigned m = __ballot_sync(0xffffffffu, empty);
          int n = __popc(m);
          int start = 0;
          if (lane == 0 && n > 0 && !S.exhausted) {
            start = atomicAdd(next_ctr, n);
            if (start + n >= n_items) S.exhausted = 1;
          } else if (lane == 0) {
            start = n_items;
          }
          start = __shfl_sync(0xffffffffu, start, 0);
          int mine = start + __popc(m & ((1u << lane) - 1u));
          if (empty && mine < n_items) {
            if (MODE == 0) {
              int r = mine;
              float t = A.min_dis[r];
              bool work = A.work_mask[r] != 0;
              S.ray[s] = r;
              S.a[s] = t; S.b[s] = A.max_dis[r];
              S.k[s] = 0;
              S.flags[s] = work ? 3 : 0;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float o = A.ray_o[(size_t)r * 3 + c], d = A.ray_d[(size_t)r * 3 + c];
                S.o[c][s] = o; S.d[c][s] = d;
                S.x[c][s] = __fadd_rn(o, __fmul_rn(d, t));       // :109
              }
            } else {
              int r = A.root_ray[mine];
              float lo = A.root_lo[mine], hi = A.root_hi[mine];
              int k = A.root_k[mine];
              bool take = (MODE == 2) ? (k == 0) : true;        // pass A: only rays that start working
              if (MODE == 2 && !take) {
                A.root_k[mine] = 0;                              // never worked: k_i = 0
              } else {
                float mid = __fmul_rn(__fadd_rn(lo, hi), 0.5f);  // :203
                S.ray[s] = mine;
                S.a[s] = lo; S.b[s] = hi; S.c[s] = mid;
                S.k[s] = (MODE == 2) ? 0 : (S.kmax - max(k, 0));
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  float o = A.ray_o[(size_t)r * 3 + c], d = A.ray_d[(size_t)r * 3 + c];
                  S.o[c][s] = o; S.d[c][s] = d;
                  S.x[c][s] = __fadd_rn(o, __fmul_rn(d, mid));   // :205
                }
              }
            }
          }
        }
      }
    }
    __syncthreads();
    const int exhausted = S.exhausted;   // read before the counting barrier: the next refill may set it
    const int active = __syncthreads_count(tid < TM && S.ray[tid] >= 0);
    if (active == 0) {
      if (exhausted) break;
      continue;   // (pass A may have skipped a whole batch of non-working roots)
    }
    if (tid == 0 && A.stats) {
      atomicAdd(A.stats + (MODE == 0 ? 0 : (MODE == 1 ? 1 : 2)), (unsigned long long)active);
      atomicAdd(A.stats + 6, 1ull);
    }

    // ------------------------------------------------------------ evaluate the SDF on all slots
    encode<TM>(net, &S.x[0][0], ebuf);
    __syncthreads();
    mlp_eval<NJ, TM>(net, act, ebuf, wbuf, red, S.f);

    // ------------------------------------------------------------ advance the state machines
    if (MODE == 0) {
      if (tid < TM && S.ray[tid] >= 0) {
        const int s = tid, r = S.ray[s];
        float f = S.f[s], t = S.a[s], tmax = S.b[s];
        bool work = S.flags[s] & 1, unf = (S.flags[s] & 2) != 0;
        unf = unf && (fabsf(f) > A.thr) && (t < tmax);                    // :113-116
        if (S.k[s] == A.iters || !unf) {                                   // :117-119
          bool conv = work && !unf && (fabsf(f) <= A.thr) && (t < tmax);   // :133-138
          A.points[(size_t)r * 3] = S.x[0][s]; A.points[(size_t)r * 3 + 1] = S.x[1][s]; A.points[(size_t)r * 3 + 2] = S.x[2][s];
          A.sdf[r] = f;
          A.dist[r] = t;
          A.conv[r] = conv ? 1 : 0;
          if (unf) A.unf_list[atomicAdd(A.counters + 1, 1)] = r;          // -> dense sampler  :55-67
          S.ray[s] = -1;
        } else {
          S.k[s] += 1;
          S.a[s] = __fadd_rn(t, f);                                        // :123-125
#pragma unroll
          for (int c = 0; c < 3; ++c) S.x[c][s] = __fadd_rn(S.x[c][s], __fmul_rn(S.d[c][s], f));
          S.flags[s] = (work ? 1 : 0) | 2;
        }
      }
    } else if (MODE == 1) {
      if (warp < G && S.g_ray[warp] >= 0) {
        const int g = warp, r = S.g_ray[g];
        const int chunk = S.g_chunk[g];
        const int j = chunk * 32 + lane;
        float val = S.f[g * 32 + lane], ts = S.ts[g * 32 + lane];
        bool neg = (j < A.n_steps) && (val < 0.f);                         // first negative sample  :162-166
        unsigned m = __ballot_sync(0xffffffffu, neg);
        float pv = __shfl_up_sync(0xffffffffu, val, 1), pt = __shfl_up_sync(0xffffffffu, ts, 1);
        if (lane == 0) { pv = S.g_prev_f[g]; pt = S.g_prev_t[g]; }
        const bool last_chunk = (chunk + 1) * 32 >= A.n_steps;
        if (m != 0u) {
          int jj = __ffs(m) - 1;
          if (lane == jj) {
            if (j >= 1) {                                                  // :167
              int idx = atomicAdd(A.counters + 3, 1);
              A.root_ray[idx] = r;
              A.root_lo[idx] = pt; A.root_hi[idx] = ts;
              A.root_k[idx] = ((pv > 0.f) && (val < 0.f)) ? 0 : -1;       // rootfind work mask  :201
            } else {
              A.points[(size_t)r * 3] = 0.f; A.points[(size_t)r * 3 + 1] = 0.f; A.points[(size_t)r * 3 + 2] = 0.f;
              A.sdf[r] = 0.f; A.dist[r] = 0.f; A.conv[r] = 0;              // :158-160
            }
            S.g_ray[g] = -1;
          }
        } else if (last_chunk) {
          if (lane == 0) {
            A.points[(size_t)r * 3] = 0.f; A.points[(size_t)r * 3 + 1] = 0.f; A.points[(size_t)r * 3 + 2] = 0.f;
            A.sdf[r] = 0.f; A.dist[r] = 0.f; A.conv[r] = 0;
            S.g_ray[g] = -1;
          }
        } else if (lane == 31) {
          S.g_prev_f[g] = val; S.g_prev_t[g] = ts;
          S.g_chunk[g] = chunk + 1;
        }
      }
    } else {
      if (tid < TM && S.ray[tid] >= 0) {
        const int s = tid, idx = S.ray[s];
        float f = S.f[s], lo = S.a[s], hi = S.b[s], mid = S.c[s];
        bool final_eval = (MODE == 3) && (S.k[s] <= 0);
        if (final_eval) {
          int r = A.root_ray[idx];
          A.points[(size_t)r * 3] = S.x[0][s]; A.points[(size_t)r * 3 + 1] = S.x[1][s]; A.points[(size_t)r * 3 + 2] = S.x[2][s];
          A.sdf[r] = f;                                                    // :216-219
          A.dist[r] = mid;
          A.conv[r] = 1;                                                   // :75
          S.ray[s] = -1;
        } else {
          if (f > 0.f) lo = mid; else hi = mid;                            // :207-212
          mid = __fmul_rn(__fadd_rn(lo, hi), 0.5f);                        // :213
          S.a[s] = lo; S.b[s] = hi; S.c[s] = mid;
#pragma unroll
          for (int c = 0; c < 3; ++c) S.x[c][s] = __fadd_rn(S.o[c][s], __fmul_rn(S.d[c][s], mid));
          if (MODE == 2) {
            S.k[s] += 1;
            bool work = __fsub_rn(hi, lo) > A.two_thr;                     // :214
            if (!work) {
              A.root_lo[idx] = lo; A.root_hi[idx] = hi; A.root_k[idx] = S.k[s];
              atomicMax(A.counters + 5, S.k[s]);
              S.ray[s] = -1;
            }
          } else {
            S.k[s] -= 1;
          }
        }
      }
    }
    __syncthreads();
  }
}

template <int NJ, int TM, int MODE>
int launch_phase(const TraceNet& net, const TraceArgs& A, int64_t max_items, cudaStream_t st) {
  using SM = Smem<NJ, TM>;
  auto kern = trace_kernel<NJ, TM, MODE>;
  static bool configured = false;
  static int occ = 1;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::bytes);
    if (e != cudaSuccess) { set_error("trace: cudaFuncSetAttribute(%zu B smem): %s", SM::bytes, cudaGetErrorString(e)); return (int)e; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TPB, SM::bytes) != cudaSuccess || occ < 1) occ = 1;
    configured = true;
  }
  int64_t per_cta = (MODE == 1) ? (TM / 32) : TM;
  int64_t want = ceil_div64(max_items, per_cta);
  int64_t cap = (int64_t)num_sms() * occ;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  kern<<<grid, TPB, SM::bytes, st>>>(net, A);
  IRONB_CHECK_LAUNCH("trace_kernel");
  return IRONB_OK;
}

template <int NJ, int TM>
int run_trace(const TraceNet& net, const TraceArgs& A, cudaStream_t st) {
  int rc;
  if ((rc = launch_phase<NJ, TM, 0>(net, A, A.N, st))) return rc;
  if ((rc = launch_phase<NJ, TM, 1>(net, A, A.N, st))) return rc;
  if ((rc = launch_phase<NJ, TM, 2>(net, A, A.N, st))) return rc;
  if ((rc = launch_phase<NJ, TM, 3>(net, A, A.N, st))) return rc;
  return IRONB_OK;
}

struct TraceWs {
  int* counters; int* unf_list; int* root_ray; float* root_lo; float* root_hi; int* root_k;
  int64_t bytes;
};
TraceWs carve_trace(int64_t N, unsigned char* base) {
  TraceWs w;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { unsigned char* p = base ? base + off : nullptr; off += (bytes + 255) / 256 * 256; return p; };
  w.counters = reinterpret_cast<int*>(take(64 * sizeof(int)));
  w.unf_list = reinterpret_cast<int*>(take(N * 4));
  w.root_ray = reinterpret_cast<int*>(take(N * 4));
  w.root_lo = reinterpret_cast<float*>(take(N * 4));
  w.root_hi = reinterpret_cast<float*>(take(N * 4));
  w.root_k = reinterpret_cast<int*>(take(N * 4));
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int64_t ironb_trace_workspace_bytes(const ironb_mlp_layout* lay, int64_t N) {
  if (N < 0 || !lay) return -1;
  int64_t fused = carve_trace(N, nullptr).bytes;
  int64_t batched = (trace_mode() >= 1 && trace_mlp_fused_supported(lay)) ? trace_batched_workspace_bytes(lay, N) : 0;
  return fused > batched ? fused : batched;
}

extern "C" int ironb_trace(const ironb_mlp_layout* lay, const float* packed, const float* ray_o,
                           const float* ray_d, const float* min_dis, const float* max_dis,
                           const uint8_t* work_mask, int64_t N, float sdf_threshold, int sphere_tracing_iters,
                           int n_steps, const float* linspace, uint8_t* conv, float* points, float* sdf,
                           float* dist, int64_t* stats, void* ws, int64_t ws_bytes, void* stream) {
  IRONB_REQUIRE(lay && lay->kind == 0, "trace: layout is not an SDF layout");
  IRONB_REQUIRE(N >= 0 && N < (1ll << 30), "trace: N out of range");
  if (N == 0) return IRONB_OK;
  IRONB_REQUIRE(packed && ray_o && ray_d && min_dis && max_dis && work_mask && linspace && conv && points && sdf && dist && ws,
                "trace: null pointer");
  IRONB_REQUIRE(n_steps >= 2 && sphere_tracing_iters >= 0, "trace: bad tracer hyper-parameters");
  const int H = lay->d_hidden;
  if (!(H == 128 || H == 256 || H == 384 || H == 512)) {
    set_error("trace: fused tracer is built for d_hidden in {128,256,384,512}, got %d", H);
    return IRONB_ENOSUP;
  }
  if (lay->in_pad[0] > 64) { set_error("trace: encoding width %d > 64", lay->in_pad[0]); return IRONB_ENOSUP; }
  if (trace_mode() >= 1 && trace_mlp_fused_supported(lay))
    return trace_batched(lay, packed, ray_o, ray_d, min_dis, max_dis, work_mask, N, sdf_threshold, sphere_tracing_iters,
                         n_steps, linspace, conv, points, sdf, dist, stats, ws, ws_bytes, as_stream(stream));
  const int last = lay->n_lin - 1;
  TraceNet net;
  memset(&net, 0, sizeof(net));
  for (int l = 0; l < last; ++l) {
    if (lay->out_pad[l] != H || (l > 0 && lay->in_pad[l] != H)) {
      set_error("trace: layer %d is not %d wide", l, H);
      return IRONB_ENOSUP;
    }
    net.wt[l] = packed + lay->off_wt[l];
    net.b[l] = packed + lay->off_b[l];
    net.in_pad[l] = lay->in_pad[l];
    net.n_true[l] = lay->out_dim[l];
  }
  net.w_last = packed + lay->off_w[last];
  net.b_last = packed + lay->off_b[last];
  net.n_hidden = last;
  net.skip_layer = lay->skip_layer;
  net.E = lay->pe_dim;
  net.Epad = lay->in_pad[0];
  net.multires = lay->multires;
  net.scale = lay->scale;
  net.beta = lay->beta;

  TraceWs w = carve_trace(N, reinterpret_cast<unsigned char*>(ws));
  IRONB_REQUIRE(ws_bytes >= w.bytes, "trace: workspace too small");
  cudaStream_t st = as_stream(stream);
  IRONB_CUDA(cudaMemsetAsync(w.counters, 0, 64 * sizeof(int), st));

  TraceArgs A;
  A.ray_o = ray_o; A.ray_d = ray_d; A.min_dis = min_dis; A.max_dis = max_dis; A.work_mask = work_mask;
  A.N = (int)N;
  A.thr = sdf_threshold;
  A.two_thr = 2.0f * sdf_threshold;
  A.iters = sphere_tracing_iters; A.n_steps = n_steps; A.linspace = linspace;
  A.conv = conv; A.points = points; A.sdf = sdf; A.dist = dist;
  A.stats = reinterpret_cast<unsigned long long*>(stats);
  A.counters = w.counters; A.unf_list = w.unf_list; A.root_ray = w.root_ray; A.root_lo = w.root_lo;
  A.root_hi = w.root_hi; A.root_k = w.root_k;

  // small batches: 32-slot CTAs so the queue spreads over more SMs
  const bool small = N < (int64_t)64 * 2 * num_sms();
  int rc = IRONB_ENOSUP;
  switch (H / 128) {
    case 1: rc = small ? run_trace<1, 32>(net, A, st) : run_trace<1, 64>(net, A, st); break;
    case 2: rc = small ? run_trace<2, 32>(net, A, st) : run_trace<2, 64>(net, A, st); break;
    case 3: rc = small ? run_trace<3, 32>(net, A, st) : run_trace<3, 64>(net, A, st); break;
    case 4: rc = small ? run_trace<4, 32>(net, A, st) : run_trace<4, 64>(net, A, st); break;
  }
  return rc;
}
