// Batched tensor-core tracer: RayTracer.forward (models/raytracer.py:45-220) as rounds of
//   [ 8 hidden-layer GEMMs + 1 sdf-row GEMM on tcgen05 (gemm_tc.cuh) ]  +  [ one state-machine kernel ]
// over device-side compacted lists of live work items, with no host round trip: every kernel of a round reads
// its row count from device memory and whole CTAs beyond it leave at once, so the host launches the worst-case
// schedule (sphere-tracing iters + sampler chunks + bisection bound) and the device decides what runs.
//
//   sphere tracing   item = ray.       Survivors of a round are re-compacted (atomic slot) and their next
//                                      encoded point e(x) is written straight into the next round's A operand.
//   dense sampler    item = sample.    32 consecutive samples per unfinished ray and round, in order, until the
//                                      first negative one (the reference evaluates all 128 and takes the first
//                                      negative: same root, fewer evaluations).  One warp per ray: ballot + ffs.
//   bisection        item = root ray.  Exactly the reference loop: every root of the call halves while ANY root
//                                      still works (the flag is an atomicMax'd row count), then one final eval.
//
// The fused persistent FFMA tracer (trace.cu) remains as the exact-fp32 implementation; this one trades
// ~4e-6 of sdf accuracy (truncating tensor-core accumulation, see gemm_tc.cuh) for tensor-core throughput.
#include <stdlib.h>

#include <atomic>

#include <cuda_fp16.h>

#include "gemm_tc.cuh"

namespace ironb {

bool trace_mlp_fused_supported(const ironb_mlp_layout* lay);
// the fused MLP evaluation (mlp_h16.cu): operands are fp16 hi and fp16 (x - hi) * 2^11
int mlp16_pick_rn(const ironb_mlp_layout* lay, int64_t rays);
int launch_trace_mlp_h16(const ironb_mlp_layout* lay, const float* packed, const CUtensorMap* mE, const CUtensorMap* mU,
                         const CUtensorMap* mW, const void* Ehi, const void* Elo, void* const* Uhi, void* const* Ulo,
                         float* Fpart, int rows_cap, int cap, const int* m_dev, int m_mul, int rn, cudaStream_t st);

namespace {

struct BArgs {
  // inputs
  const float *ray_o, *ray_d, *min_dis, *max_dis;
  const uint8_t* work_mask;
  const float* linspace;
  int N, iters, n_steps, bound;
  float thr, two_thr;
  // network bits the elementwise kernels need
  int multires, Epad, H;
  float scale;
  // outputs
  uint8_t* conv;
  float *points, *sdf, *dist;
  unsigned long long* stats;
  // workspace
  int* c;                       // counters, see indices below
  float* t; int* k; int* flags; float* x;            // per ray
  int* list[2];                 // sphere-tracing item -> ray
  int* unf_list; float *smin, *smax, *prev_f, *prev_t;   // per unfinished ray
  int* glist[2];                // sampler group -> unfinished-ray index
  int* root_ray; float *root_lo, *root_hi, *root_mid; int* root_work;
  float* Ehi; float* Elo;       // [CAP][Epad] encoded points, already split (the MLP's A operand): fp16 hi and
                                // fp16 (x - hi) * 2^11 (mlp_h16.cu)
  const float* Fpart;           // [nparts][cap] partial sums of the sdf row, written by the fused MLP kernel
  const float* b_last;
  int nparts, cap;
  int cap_groups;
};
// counter indices
constexpr int C_ST = 0;      // 0..2 rotating row counts (sphere tracing)
constexpr int C_NUNF = 3, C_NROOT = 4, C_KMAX = 5;
constexpr int C_SG = 8;      // 8..10 rotating group counts (sampler)
constexpr int C_BR = 16;     // 16.. row count of bisection round j (see the bisection kernels)
constexpr int NCOUNTERS = 128;

__device__ __forceinline__ void put_split(__half* __restrict__ ehi, __half* __restrict__ elo, int i, float v) {
  const __half h = __float2half_rn(v);
  ehi[i] = h;
  elo[i] = __float2half_rn((v - __half2float(h)) * 2048.f);
}
struct BArgs;
__device__ __forceinline__ void write_pe(const BArgs& A, size_t row, const float x[3], int part);

// sdf of work item i: either the sdf-row GEMM's output, or the fused MLP's partial sums added in a fixed order
__device__ __forceinline__ float read_f(const BArgs& A, size_t i) {
  float s = 0.f;
  for (int p = 0; p < A.nparts; ++p) s += A.Fpart[(size_t)p * A.cap + i];
  return __fdiv_rn(s + __ldg(A.b_last), A.scale);
}

// encoded point of work item `row`, written as the split A operand of the MLP's first layer
// part < 0: the whole encoding by one thread; part in [0, 8): thread `part` of the 8 that share a work item writes
// frequency `part` (part == multires: the raw coordinates + padding), so the dependent sincos chain per thread is 3 long
__device__ __forceinline__ void write_pe(const BArgs& A, size_t row, const float x[3], int part = -1) {
  __half* __restrict__ ehi = reinterpret_cast<__half*>(A.Ehi) + row * A.Epad;     // row pitch: Epad halfs
  __half* __restrict__ elo = reinterpret_cast<__half*>(A.Elo) + row * A.Epad;
  const float xs[3] = {x[0] * A.scale, x[1] * A.scale, x[2] * A.scale};
  if (part < 0 || part == A.multires) {                 // the raw coordinates and the zero padding
    put_split(ehi, elo, 0, xs[0]); put_split(ehi, elo, 1, xs[1]); put_split(ehi, elo, 2, xs[2]);
    for (int w = 3 + 6 * A.multires; w < A.Epad; ++w) put_split(ehi, elo, w, 0.f);
  }
  const int k0 = part < 0 ? 0 : part, k1 = part < 0 ? A.multires : min(part + 1, A.multires);
  for (int k = k0; k < k1; ++k) {                       // frequency 2^k: sin block then cos block (embedder.py:25-33)
    const float f = (float)(1 << k);
    const int w = 3 + 6 * k;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sn, co;
      sincosf(xs[c] * f, &sn, &co);
      put_split(ehi, elo, w + c, sn);
      put_split(ehi, elo, w + 3 + c, co);
    }
  }
}
__device__ __forceinline__ void ray_point(const BArgs& A, int r, float t, float x[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(A.ray_o[(size_t)r * 3 + c], __fmul_rn(A.ray_d[(size_t)r * 3 + c], t));
}
__device__ __forceinline__ void write_miss(const BArgs& A, int r) {   // raytracer.py:158-160
  A.points[(size_t)r * 3] = 0.f; A.points[(size_t)r * 3 + 1] = 0.f; A.points[(size_t)r * 3 + 2] = 0.f;
  A.sdf[r] = 0.f; A.dist[r] = 0.f; A.conv[r] = 0;
}

// ---------------------------------------------------------------- sphere tracing (:105-140)
// The per-ray kernels run 8 threads per work item (PE_T): all 8 compute the (cheap) state update redundantly, thread 0 of
// the group owns the state writes and the atomics, and each thread writes one frequency band of the next encoded point.
constexpr int PE_T = 8;

__global__ void __launch_bounds__(256) st_init_kernel(BArgs A) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / PE_T, part = tid % PE_T;
  if (tid == 0) { A.c[C_ST] = A.N; A.c[C_ST + 1] = 0; A.c[C_ST + 2] = 0; }
  if (i >= A.N) return;
  float t = A.min_dis[i], x[3];
  ray_point(A, i, t, x);                                           // :109
  if (part == 0) {
    A.t[i] = t; A.k[i] = 0; A.flags[i] = A.work_mask[i] ? 3 : 0;
    A.x[(size_t)i * 3] = x[0]; A.x[(size_t)i * 3 + 1] = x[1]; A.x[(size_t)i * 3 + 2] = x[2];
    A.list[0][i] = i;
  }
  write_pe(A, (size_t)i, x, part);
}

__global__ void __launch_bounds__(256) st_update_kernel(BArgs A, int round) {
  const int cur = C_ST + round % 3, nxt = C_ST + (round + 1) % 3, clr = C_ST + (round + 2) % 3;
  const int rows = A.c[cur];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / PE_T, part = tid % PE_T;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);              // the 8 lanes of this work item
  if (tid == 0) {
    A.c[clr] = 0;
    if (A.stats && rows > 0) { atomicAdd(A.stats + 0, (unsigned long long)rows); atomicAdd(A.stats + 6, (unsigned long long)((rows + 127) / 128)); }
  }
  if (i >= rows) return;
  const int r = A.list[round & 1][i];
  const float f = read_f(A, i);
  float t = A.t[r];
  const float tmax = A.max_dis[r];
  const int fl = A.flags[r];
  const bool work = fl & 1;
  bool unf = (fl & 2) != 0;
  unf = unf && (fabsf(f) > A.thr) && (t < tmax);                   // :113-116
  float x[3] = {A.x[(size_t)r * 3], A.x[(size_t)r * 3 + 1], A.x[(size_t)r * 3 + 2]};
  const int k = A.k[r];
  __syncwarp(gmask);                                               // every lane of the item has read the state
  if (k == A.iters || !unf) {                                      // :117-119
    if (part != 0) return;
    const bool conv = work && !unf && (fabsf(f) <= A.thr) && (t < tmax);   // :133-138
    A.points[(size_t)r * 3] = x[0]; A.points[(size_t)r * 3 + 1] = x[1]; A.points[(size_t)r * 3 + 2] = x[2];
    A.sdf[r] = f; A.dist[r] = t; A.conv[r] = conv ? 1 : 0;
    if (unf) A.unf_list[atomicAdd(A.c + C_NUNF, 1)] = r;          // -> dense sampler (:55-67)
  } else {
    t = __fadd_rn(t, f);                                           // :123-125
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(x[c], __fmul_rn(A.ray_d[(size_t)r * 3 + c], f));
    int slot = 0;
    if (part == 0) {
      A.t[r] = t; A.k[r] = k + 1; A.flags[r] = (work ? 1 : 0) | 2;
      A.x[(size_t)r * 3] = x[0]; A.x[(size_t)r * 3 + 1] = x[1]; A.x[(size_t)r * 3 + 2] = x[2];
      slot = atomicAdd(A.c + nxt, 1);
      A.list[(round + 1) & 1][slot] = r;
    }
    slot = __shfl_sync(gmask, slot, threadIdx.x & 24);
    write_pe(A, (size_t)slot, x, part);
  }
}

// ---------------------------------------------------------------- dense sampler (:142-197), one warp per ray
__device__ __forceinline__ float sample_t(const BArgs& A, float smin, float smax, int j) {
  const int jc = min(j, A.n_steps - 1);
  return __fadd_rn(smin, __fmul_rn(__ldg(A.linspace + jc), __fsub_rn(smax, smin)));   // :144-147
}

__global__ void __launch_bounds__(256) smp_prepare_kernel(BArgs A, int pass) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_unf = A.c[C_NUNF];
  const int first = pass * A.cap_groups;
  const int groups = max(0, min(A.cap_groups, n_unf - first));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    A.c[C_SG] = groups; A.c[C_SG + 1] = 0; A.c[C_SG + 2] = 0;
    if (A.stats && pass == 0) atomicAdd(A.stats + 3, (unsigned long long)n_unf);
  }
  if (g >= groups) return;
  const int u = first + g;
  const int r = A.unf_list[u];
  const float t = A.dist[r], f = A.sdf[r];
  const bool outside = f > 0.f;                                    // :59-65
  const float smin = outside ? t : A.min_dis[r];
  const float smax = outside ? A.max_dis[r] : t;
  if (lane == 0) { A.smin[u] = smin; A.smax[u] = smax; A.prev_f[u] = 0.f; A.prev_t[u] = 0.f; A.glist[0][g] = u; }
  float x[3];
  ray_point(A, r, sample_t(A, smin, smax, lane), x);
  write_pe(A, (size_t)g * 32 + lane, x);
}

__global__ void __launch_bounds__(256) smp_update_kernel(BArgs A, int chunk) {
  const int cur = C_SG + chunk % 3, nxt = C_SG + (chunk + 1) % 3, clr = C_SG + (chunk + 2) % 3;
  const int groups = A.c[cur];
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    A.c[clr] = 0;
    if (A.stats && groups > 0) { atomicAdd(A.stats + 1, (unsigned long long)groups * 32ull); atomicAdd(A.stats + 6, (unsigned long long)((groups * 32 + 127) / 128)); }
  }
  if (g >= groups) return;
  const int u = A.glist[chunk & 1][g];
  const int r = A.unf_list[u];
  const float smin = A.smin[u], smax = A.smax[u];
  const int j = chunk * 32 + lane;
  const float val = read_f(A, (size_t)g * 32 + lane);
  const float ts = sample_t(A, smin, smax, j);
  const bool neg = (j < A.n_steps) && (val < 0.f);                 // first negative sample (:162-166)
  const unsigned m = __ballot_sync(0xffffffffu, neg);
  float pv = __shfl_up_sync(0xffffffffu, val, 1), pt = __shfl_up_sync(0xffffffffu, ts, 1);
  if (lane == 0) { pv = A.prev_f[u]; pt = A.prev_t[u]; }
  const bool last_chunk = (chunk + 1) * 32 >= A.n_steps;
  if (m != 0u) {
    const int jj = __ffs(m) - 1;
    if (lane == jj) {
      if (j >= 1) {                                                // :167
        const int idx = atomicAdd(A.c + C_NROOT, 1);
        A.root_ray[idx] = r; A.root_lo[idx] = pt; A.root_hi[idx] = ts;
        A.root_work[idx] = ((pv > 0.f) && (val < 0.f)) ? 1 : 0;   // rootfind work mask (:201)
      } else {
        write_miss(A, r);
      }
    }
  } else if (last_chunk) {
    if (lane == 0) write_miss(A, r);
  } else {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(A.c + nxt, 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (lane == 31) { A.prev_f[u] = val; A.prev_t[u] = ts; }
    if (lane == 0) A.glist[(chunk + 1) & 1][slot] = u;
    float x[3];
    ray_point(A, r, sample_t(A, smin, smax, j + 32), x);
    write_pe(A, (size_t)slot * 32 + lane, x);
  }
}

// ---------------------------------------------------------------- bisection (:199-220), two iterations per MLP round
// The reference halves EVERY root of the call while any root still works (one batch-coupled loop, k_max iterations), then
// evaluates the final midpoints once more.  Its iterations are strictly sequential -- but the midpoint of iteration k+1 is
// one of only two values, (lo + mid)/2 or (mid + hi)/2, both computable before f(mid) is known.  A round therefore evaluates
// THREE points per root (mid and both candidate next midpoints, formed by the reference's own expression, so the chosen one is
// bit-identical to what the reference would compute) and advances the loop by two iterations; the "does any root still work"
// test of the reference sits between the two iterations and is a device-wide flag between two tiny kernels.  The final
// evaluation is whichever already-evaluated point the loop stops on.  k_max iterations cost ceil((k_max + 1) / 2) MLP rounds
// instead of k_max + 1 (7 + 1 -> 4 at the training patch), at 1.5x the (few) bisection evaluations.
//   c[C_BR + j]  rows of round j (n_root, or 0 = the loop and the final evaluation are done)
//   c[C_BA + j]  1 if some root works when round j starts (i.e. iteration 2j + 1 takes place)
//   c[C_BF + j]  1 if some root still works after iteration 2j + 1 (i.e. iteration 2j + 2 takes place)
constexpr int BIS_MAX_ROUNDS = 24, C_BA = C_BR + BIS_MAX_ROUNDS, C_BF = C_BA + BIS_MAX_ROUNDS;   // 16.., 40.., 64..87

__device__ __forceinline__ void bis_write_candidates(const BArgs& A, int i, int n_root, int r, float lo, float hi, float mid, int part) {
  const float cand[3] = {mid, __fmul_rn(__fadd_rn(lo, mid), 0.5f), __fmul_rn(__fadd_rn(mid, hi), 0.5f)};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float x[3];
    ray_point(A, r, cand[c], x);                                    // :205
    write_pe(A, (size_t)c * n_root + i, x, part);
  }
}
__device__ __forceinline__ void bis_finish(const BArgs& A, int i, float mid, float f) {   // :216-219, :75
  const int r = A.root_ray[i];
  float x[3];
  ray_point(A, r, mid, x);
  A.points[(size_t)r * 3] = x[0]; A.points[(size_t)r * 3 + 1] = x[1]; A.points[(size_t)r * 3 + 2] = x[2];
  A.sdf[r] = f; A.dist[r] = mid; A.conv[r] = 1;
}

__global__ void __launch_bounds__(256) bis_prepare_kernel(BArgs A) {
  const int n_root = A.c[C_NROOT];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / PE_T, part = tid % PE_T;
  if (tid == 0) {
    if (A.stats) atomicAdd(A.stats + 4, (unsigned long long)n_root);
    A.c[C_BR] = n_root;                                               // round 0 always runs: it holds the final evaluation at least
  }
  if (i >= n_root) return;
  const float lo = A.root_lo[i], hi = A.root_hi[i];
  const float mid = __fmul_rn(__fadd_rn(lo, hi), 0.5f);                // :203
  if (part == 0) {
    A.root_mid[i] = mid;
    if (A.root_work[i]) atomicMax(A.c + C_BA, 1);                     // while work.any()  (:204)
  }
  bis_write_candidates(A, i, n_root, A.root_ray[i], lo, hi, mid, part);
}

// first half of a round: only the device-wide "does any root still work after iteration 2j + 1" flag
__global__ void __launch_bounds__(256) bis_flag_kernel(BArgs A, int round) {
  const int rows = A.c[C_BR + round];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows || !A.c[C_BA + round] || !A.root_work[i]) return;
  const float f1 = read_f(A, (size_t)i);
  float lo = A.root_lo[i], hi = A.root_hi[i];
  const float mid = A.root_mid[i];
  if (f1 > 0.f) lo = mid; else hi = mid;
  if (__fsub_rn(hi, lo) > A.two_thr) atomicMax(A.c + C_BF + round, 1);
}

__global__ void __launch_bounds__(256) bis_update_kernel(BArgs A, int round, int is_last) {
  const int rows = A.c[C_BR + round];                                 // n_root, or 0 when everything is done
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / PE_T, part = tid % PE_T;
  // is_last: the host's schedule ends here (it covers the halvings the reference can need by construction); finish regardless
  const int active = A.c[C_BA + round], more = is_last ? 0 : A.c[C_BF + round];
  if (tid == 0 && rows > 0) {
    A.c[C_KMAX] = active ? (more ? 2 * round + 2 : 2 * round + 1) : 2 * round;     // iterations the reference's loop has run
    if (A.stats) { atomicAdd(A.stats + 2, (unsigned long long)rows * 3ull); atomicAdd(A.stats + 6, (unsigned long long)((rows * 3 + 127) / 128)); }
  }
  if (i >= rows) return;
  const float f1 = read_f(A, (size_t)i);
  float lo = A.root_lo[i], hi = A.root_hi[i], mid = A.root_mid[i];
  int work = A.root_work[i];
  __syncwarp(0xffu << (threadIdx.x & 24));                          // every lane of the item has read the state
  if (!active) {                       // the loop ended before this round: f1 is the final evaluation (:216-219)
    if (part == 0) bis_finish(A, i, mid, f1);
    return;
  }
  // iteration 2j + 1: every root of the call (:207-214)
  const bool up = f1 > 0.f;
  if (up) lo = mid; else hi = mid;
  const float f2 = read_f(A, (size_t)(up ? 2 : 1) * rows + i);        // f at the new midpoint, evaluated speculatively
  mid = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
  work = work && (__fsub_rn(hi, lo) > A.two_thr);
  if (!more) {                         // no root works any more: the loop ended after this iteration, f2 is the final evaluation
    if (part == 0) bis_finish(A, i, mid, f2);
    return;
  }
  // iteration 2j + 2
  if (f2 > 0.f) lo = mid; else hi = mid;
  mid = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
  work = work && (__fsub_rn(hi, lo) > A.two_thr);
  if (part == 0) {
    A.root_lo[i] = lo; A.root_hi[i] = hi; A.root_mid[i] = mid; A.root_work[i] = work ? 1 : 0;
    if (work) atomicMax(A.c + C_BA + round + 1, 1);
    if (i == 0) A.c[C_BR + round + 1] = rows;                         // the next round holds more iterations or the final evaluation
  }
  bis_write_candidates(A, i, rows, A.root_ray[i], lo, hi, mid, part);
}

__global__ void __launch_bounds__(32) bis_stats_kernel(BArgs A) {
  if (threadIdx.x == 0 && A.stats) atomicMax(A.stats + 5, (unsigned long long)A.c[C_KMAX]);
}

struct BWs {
  int* c; float* t; int* k; int* flags; float* x; int* list[2]; int* unf_list; float *smin, *smax, *prev_f, *prev_t;
  int* glist[2]; int* root_ray; float *root_lo, *root_hi, *root_mid; int* root_work;
  float *Ehi, *Elo, *Uhi[2], *Ulo[2], *Fpart;
  int64_t cap; int64_t bytes;
};

BWs carve_b(const ironb_mlp_layout* lay, int64_t N, unsigned char* base) {
  BWs w;
  memset(&w, 0, sizeof(w));
  int64_t off = 0;
  auto take = [&](int64_t bytes) { unsigned char* p = base ? base + off : nullptr; off += (bytes + 255) / 256 * 256; return p; };
  // sampler items per pass.  Small calls (the training patch) take every ray in ONE pass: a second pass is 9 launches that
  // find nothing to do whenever at most half of the rays reach the sampler, and 32 N rows are only ~1 GB of workspace there.
  // Large calls keep 16 N (50 % unfinished rays fit in one pass): the workspace grows with 8.5 KB per row at H = 512.
  int64_t cap = (N <= 8192 ? 32 : 16) * N;
  if (cap < 4096) cap = 4096;
  if (cap > (1 << 21)) cap = (1 << 21);
  if (cap < N) cap = N;
  cap = (cap + 127) / 128 * 128;
  w.cap = cap;
  const int H = lay->d_hidden, Epad = lay->in_pad[0], last = lay->n_lin - 1;
  w.c = (int*)take(NCOUNTERS * 4);
  w.t = (float*)take(N * 4); w.k = (int*)take(N * 4); w.flags = (int*)take(N * 4); w.x = (float*)take(N * 12);
  w.list[0] = (int*)take(N * 4); w.list[1] = (int*)take(N * 4);
  w.unf_list = (int*)take(N * 4);
  w.smin = (float*)take(N * 4); w.smax = (float*)take(N * 4); w.prev_f = (float*)take(N * 4); w.prev_t = (float*)take(N * 4);
  w.glist[0] = (int*)take(cap / 32 * 4); w.glist[1] = (int*)take(cap / 32 * 4);
  w.root_ray = (int*)take(N * 4); w.root_lo = (float*)take(N * 4); w.root_hi = (float*)take(N * 4);
  w.root_mid = (float*)take(N * 4); w.root_work = (int*)take(N * 4);
  w.Ehi = (float*)take(cap * Epad * 4); w.Elo = (float*)take(cap * Epad * 4);
  w.Fpart = (float*)take(cap * 4 * 32);          // up to 4 x 8 partial sums of the sdf row per work item
  for (int b = 0; b < 2; ++b) { w.Uhi[b] = (float*)take(cap * (int64_t)H * 4); w.Ulo[b] = (float*)take(cap * (int64_t)H * 4); }
  w.bytes = off;
  return w;
}

static std::atomic<int> g_trace_mode{-1};

}  // namespace

int trace_mode() {
  int m = g_trace_mode.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = getenv("IRONB_TRACE");
    // IRONB_TRACE=fused: fp32 FFMA tracer; default: batched tcgen05 tracer, fp16x2-split operands
    m = (e && (e[0] == 'f' || e[0] == 'F' || e[0] == '0')) ? 0 : 2;
    g_trace_mode.store(m, std::memory_order_relaxed);
  }
  return m;
}
int set_trace_mode(int mode) {
  int prev = trace_mode();
  g_trace_mode.store(mode <= 0 ? 0 : 2, std::memory_order_relaxed);   // 1 (the retired 3xTF32 tracer) means 2
  return prev;
}

int64_t trace_batched_workspace_bytes(const ironb_mlp_layout* lay, int64_t N) { return carve_b(lay, N, nullptr).bytes; }

int trace_batched(const ironb_mlp_layout* lay, const float* packed, const float* ray_o, const float* ray_d,
                  const float* min_dis, const float* max_dis, const uint8_t* work_mask, int64_t N, float thr, int iters,
                  int n_steps, const float* linspace, uint8_t* conv, float* points, float* sdf, float* dist,
                  int64_t* stats, void* ws, int64_t ws_bytes, cudaStream_t st) {
  const int H = lay->d_hidden, Epad = lay->in_pad[0], last = lay->n_lin - 1;
  if (!trace_mlp_fused_supported(lay)) { set_error("trace: batched tcgen05 tracer needs d_hidden in {128,256,512}"); return IRONB_ENOSUP; }
  for (int l = 0; l < last; ++l)
    if (lay->out_pad[l] != H || (l > 0 && lay->in_pad[l] != H)) { set_error("trace: layer %d is not %d wide", l, H); return IRONB_ENOSUP; }
  if (lay->multires >= PE_T) { set_error("trace: batched tracer supports multires <= %d", PE_T - 1); return IRONB_ENOSUP; }
  BWs w = carve_b(lay, N, reinterpret_cast<unsigned char*>(ws));
  if (ws_bytes < w.bytes) { set_error("trace: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)w.bytes); return IRONB_EINVAL; }
  IRONB_CUDA(cudaMemsetAsync(w.c, 0, NCOUNTERS * 4, st));

  BArgs A;
  memset(&A, 0, sizeof(A));
  A.ray_o = ray_o; A.ray_d = ray_d; A.min_dis = min_dis; A.max_dis = max_dis; A.work_mask = work_mask; A.linspace = linspace;
  A.N = (int)N; A.iters = iters; A.n_steps = n_steps;
  A.thr = thr; A.two_thr = 2.0f * thr;
  // halvings until an interval of at most 2r/(n_steps-1) (r = 1) is <= 2*thr, plus slack; the loop is data driven below it
  int bound = 1;
  for (double len = 2.0 / (n_steps - 1); len > 2.0 * thr && bound < 40; len *= 0.5) ++bound;
  bound += 2;
  if (bound > 2 * (BIS_MAX_ROUNDS - 2)) bound = 2 * (BIS_MAX_ROUNDS - 2);
  A.bound = bound;
  A.multires = lay->multires; A.Epad = Epad; A.H = H; A.scale = lay->scale;
  A.conv = conv; A.points = points; A.sdf = sdf; A.dist = dist;
  A.stats = reinterpret_cast<unsigned long long*>(stats);
  A.c = w.c; A.t = w.t; A.k = w.k; A.flags = w.flags; A.x = w.x; A.list[0] = w.list[0]; A.list[1] = w.list[1];
  A.unf_list = w.unf_list; A.smin = w.smin; A.smax = w.smax; A.prev_f = w.prev_f; A.prev_t = w.prev_t;
  A.glist[0] = w.glist[0]; A.glist[1] = w.glist[1];
  A.root_ray = w.root_ray; A.root_lo = w.root_lo; A.root_hi = w.root_hi; A.root_mid = w.root_mid; A.root_work = w.root_work;
  A.Ehi = w.Ehi; A.Elo = w.Elo;
  A.cap_groups = (int)(w.cap / 32);
  int rn = mlp16_pick_rn(lay, N);                 // features per CTA of the MLP kernel: 64 (latency shape) or 128
  int C = H / rn;
  A.Fpart = w.Fpart; A.b_last = packed + lay->off_b[last]; A.nparts = 4 * C; A.cap = (int)w.cap;

  // The fp16x2-split weight copies come with the packed buffer (ironb_mlp_fold writes them once per parameter update;
  // round 1 re-split all hidden layers on every tracer call: 8 launches).  Tensor maps: operands are fixed for the call.
  int rc;
  for (int l = 0; l < last; ++l)
    if (lay->off_h16[l] <= 0) { set_error("trace: the layout carries no fp16 weight copies (layer %d)", l); return IRONB_EINVAL; }
  CUtensorMap mE[2], mU[4], mW[2 * IRONB_MAX_LIN];
  auto mk = [&](CUtensorMap* m, const void* p, int rows, int K, int ld, int box_rows = 128) -> int {
    return tc::make_map_h(m, p, rows, K, ld, box_rows);
  };
  if ((rc = mk(&mE[0], w.Ehi, (int)w.cap, Epad, Epad))) return rc;
  if ((rc = mk(&mE[1], w.Elo, (int)w.cap, Epad, Epad))) return rc;
  for (int b = 0; b < 2; ++b) {
    if ((rc = mk(&mU[b * 2], w.Uhi[b], (int)w.cap, H, H))) return rc;
    if ((rc = mk(&mU[b * 2 + 1], w.Ulo[b], (int)w.cap, H, H))) return rc;
  }
  for (int l = 0; l < last; ++l) {
    const __half* whi = reinterpret_cast<const __half*>(packed + lay->off_h16[l]);
    const __half* wlo = whi + (int64_t)lay->out_pad[l] * lay->in_pad[l];
    if ((rc = mk(&mW[l * 2], whi, lay->out_pad[l], lay->in_pad[l], lay->in_pad[l], rn))) return rc;
    if ((rc = mk(&mW[l * 2 + 1], wlo, lay->out_pad[l], lay->in_pad[l], lay->in_pad[l], rn))) return rc;
  }
  // one MLP evaluation of the first (*m_dev x m_mul) rows: all hidden layers + the sdf row in one cluster launch
  // (mlp_h16.cu: fp16x2 split, two tiles in flight)
  auto mlp = [&](int rows_cap, const int* m_dev, int m_mul) -> int {
    void* uh[2] = {w.Uhi[0], w.Uhi[1]};
    void* ul[2] = {w.Ulo[0], w.Ulo[1]};
    return launch_trace_mlp_h16(lay, packed, mE, mU, mW, w.Ehi, w.Elo, uh, ul, w.Fpart, rows_cap, (int)w.cap, m_dev, m_mul, rn, st);
  };
  const int nb = (int)ceil_div64(N * PE_T, 256);    // st_* / bis_* kernels: PE_T threads per work item
  const int nb1 = (int)ceil_div64(N, 256);

  // ---- sphere tracing: iters + 1 evaluations
  st_init_kernel<<<nb, 256, 0, st>>>(A);
  IRONB_CHECK_LAUNCH("st_init_kernel");
  for (int r = 0; r <= iters; ++r) {
    if ((rc = mlp((int)N, w.c + C_ST + r % 3, 1))) return rc;
    st_update_kernel<<<nb, 256, 0, st>>>(A, r);
    IRONB_CHECK_LAUNCH("st_update_kernel");
  }
  // ---- dense sampler: passes of cap/32 rays, chunks of 32 samples
  const int chunks = (n_steps + 31) / 32;
  const int64_t passes = ceil_div64(N * 32, w.cap);
  const int gb = (int)ceil_div64(w.cap / 32 * 32, 256);
  for (int64_t p = 0; p < passes; ++p) {
    smp_prepare_kernel<<<gb, 256, 0, st>>>(A, (int)p);
    IRONB_CHECK_LAUNCH("smp_prepare_kernel");
    for (int ch = 0; ch < chunks; ++ch) {
      if ((rc = mlp((int)w.cap, w.c + C_SG + ch % 3, 32))) return rc;
      smp_update_kernel<<<gb, 256, 0, st>>>(A, ch);
      IRONB_CHECK_LAUNCH("smp_update_kernel");
    }
  }
  // ---- bisection: two iterations of the reference loop per round, the final evaluation rides on the last round
  bis_prepare_kernel<<<nb, 256, 0, st>>>(A);
  IRONB_CHECK_LAUNCH("bis_prepare_kernel");
  const int rounds = (bound + 1) / 2 + 1 < BIS_MAX_ROUNDS - 1 ? (bound + 1) / 2 + 1 : BIS_MAX_ROUNDS - 1;
  for (int j = 0; j < rounds; ++j) {
    if ((rc = mlp((int)(3 * N), w.c + C_BR + j, 3))) return rc;
    bis_flag_kernel<<<nb1, 256, 0, st>>>(A, j);
    IRONB_CHECK_LAUNCH("bis_flag_kernel");
    bis_update_kernel<<<nb, 256, 0, st>>>(A, j, j == rounds - 1 ? 1 : 0);
    IRONB_CHECK_LAUNCH("bis_update_kernel");
  }
  bis_stats_kernel<<<1, 32, 0, st>>>(A);
  IRONB_CHECK_LAUNCH("bis_stats_kernel");
  return IRONB_OK;
}

}  // namespace ironb

extern "C" int ironb_set_trace_mode(int mode) { return ironb::set_trace_mode(mode); }
