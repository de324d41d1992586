// Fused SDF-MLP evaluation for the batched tracer: ALL hidden layers of one 128-row tile in ONE launch.
//
// A cluster of C = H/128 CTAs owns a tile of 128 work items (rows).  CTA rank r computes output features
// [128 r, 128 r + 128) of every layer with the tcgen05 3xTF32 pipeline of gemm_tc.cuh (TMA -> in-smem hi/lo split
// -> 3 x tcgen05.mma.kind::tf32 -> split TMEM accumulators).  A layer's epilogue (bias + softplus [+ skip concat])
// writes its 128 x 128 activation block to a global ping-pong buffer (L2 resident), fences it to the async proxy,
// and a cluster barrier makes the whole 128 x H row block visible before any CTA of the cluster TMA-loads it as
// the next layer's A operand.  Nothing else synchronises: no kernel boundary, no TMEM re-allocation, no barrier
// re-initialisation between layers, and the mbarrier pipeline state simply carries on.
//
// The last hidden layer never stores its activations: its epilogue dots them with row 0 of the output layer (the sdf
// row) and writes per-(rank, column-half) partial sums; the tracer's state-machine kernel adds the 2C partials in a
// fixed order (deterministic) and applies bias / scale.
#include "gemm_tc.cuh"

namespace ironb {
namespace mlp {

using namespace tc;

constexpr int MAXL = IRONB_MAX_LIN;
constexpr int NT = 320;                     // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..9 split + epilogue
constexpr int NSPLIT = 256;
constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;

struct Maps {
  CUtensorMap e;          // encoded points  [cap][Epad]
  CUtensorMap u[2];       // activation ping-pong [cap][H]
  CUtensorMap w[MAXL];    // hidden-layer weights W_l [H][K_l]
};

struct Args {
  const float* bias[MAXL];
  const float* w_last;    // row 0 of the output layer [H]
  const float* E;         // encoded points (skip concat reads them)
  float* U[2];
  float* Fpart;           // [2C][cap] partial sdf sums
  int n_true[MAXL];
  int kpad[MAXL];
  int n_hidden;           // layers evaluated here (n_lin - 1)
  int skip_layer, Epad, Edim, H;
  float beta, inv_beta;
  const int* m_dev;
  int m_mul, rows_cap, cap, write_hi;
};

__device__ __forceinline__ float softplus_fast(float z, float beta, float inv_beta) {
  // max(z,0) + log1p(exp(-|beta z|))/beta: the log term is <= ln2 and is scaled by 1/beta, so ex2/lg2.approx
  // contribute ~1e-9 absolute; for beta*z > 20 the term vanishes in fp32, which is torch's threshold branch.
  const float e = __expf(-fabsf(z * beta));
  return fmaxf(z, 0.f) + __logf(1.f + e) * inv_beta;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NT, 1) mlp_fused_kernel(const __grid_constant__ Maps maps, const Args a) {
  int M = a.rows_cap;
  if (a.m_dev != nullptr) {
    const int md = *a.m_dev * a.m_mul;
    if (md < M) M = md;
  }
  const int m0 = blockIdx.y * BM;
  if (m0 >= M) return;                                   // uniform over the cluster (same blockIdx.y)
  const int rank = blockIdx.x, n0 = rank * BN;

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto conv = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  const uint32_t acc_bar = bars + 8u * (3 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), NSPLIT);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot_ptr;

  int it_base = 0;   // pipeline iterations consumed by earlier layers (every role advances it identically)
  for (int l = 0; l < a.n_hidden; ++l) {
    const int K = a.kpad[l];
    const int nk = (K + BK - 1) / BK;
    if (warp == 0) {
      // ================= TMA producer =================
      if (lane == 0) {
        const CUtensorMap* mapA = (l == 0) ? &maps.e : &maps.u[(l - 1) & 1];
        const CUtensorMap* mapB = &maps.w[l];
        for (int i = 0; i < nk; ++i) {
          const int it = it_base + i, s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty(s), ph ^ 1);
          mbar_arrive_expect_tx(full(s), HI_BYTES);
          const uint32_t st = base + s * STAGE_BYTES;
          tma_load_2d(st, mapA, i * BK, m0, full(s));
          tma_load_2d(st + TILE_BYTES, mapB, i * BK, n0, full(s));
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ================= MMA issuer =================
      if (lane == 0) {
        for (int i = 0; i < nk; ++i) {
          const int it = it_base + i, s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(conv(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = base + s * STAGE_BYTES;
          const uint64_t a_hi = make_desc(st), b_hi = make_desc(st + TILE_BYTES);
          const uint64_t a_lo = make_desc(st + 2 * TILE_BYTES), b_lo = make_desc(st + 3 * TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);
            const int g = i * (BK / 8) + kk;
            tc_mma_tf32(tmem + ((g & 1) ? 128u : 0u), a_hi + adv, b_hi + adv, IDESC, g >= 2 ? 1u : 0u);
            tc_mma_tf32(tmem + 256u, a_lo + adv, b_hi + adv, IDESC, g >= 1 ? 1u : 0u);
            tc_mma_tf32(tmem + 256u, a_hi + adv, b_lo + adv, IDESC, 1u);
          }
          tc_commit(empty(s));
        }
        tc_commit(acc_bar);
      }
      __syncwarp();
    } else {
      // ================= split (all 8 warps), then epilogue =================
      const int t = threadIdx.x - 64;   // 0..255
      for (int i = 0; i < nk; ++i) {
        const int it = it_base + i, s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(full(s), ph);
        float4* hi = reinterpret_cast<float4*>(base_ptr + s * STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(base_ptr + s * STAGE_BYTES + HI_BYTES);
        split_stage(hi, lo, t, NSPLIT, HI_BYTES / 16 / NSPLIT, a.write_hi);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(conv(s));
      }
      mbar_wait(acc_bar, (uint32_t)(l & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3, half = (warp - 2) >> 2;
      const int m = m0 + q * 32 + lane;
      const bool last = (l == a.n_hidden - 1);
      const bool pre_skip = (l + 1 == a.skip_layer);
      const int n_true = a.n_true[l];
      const float* __restrict__ bias = a.bias[l];
      float* __restrict__ Un = a.U[l & 1];
      float dot = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        uint32_t r[32], r1[32], r2[32];
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + 128u, r1);
        tmem_ld32(taddr + 256u, r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < M) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = g * 4 + j;
              const int n = n0 + c0 + i;
              const float z = __fadd_rn(__fadd_rn(__uint_as_float(r[i]), __uint_as_float(r1[i])), __uint_as_float(r2[i]));
              if (n < n_true) {
                float act = softplus_fast(z + __ldg(bias + n), a.beta, a.inv_beta);
                u[j] = pre_skip ? __fdiv_rn(act, IRONB_SQRT2F) : act;
              } else {
                const int ce = n - n_true;
                u[j] = (pre_skip && ce < a.Edim) ? __fdiv_rn(__ldg(a.E + (size_t)m * a.Epad + ce), IRONB_SQRT2F) : 0.f;
              }
              if (last) dot = fmaf(u[j], __ldg(a.w_last + n), dot);
            }
            if (!last) *reinterpret_cast<float4*>(Un + (size_t)m * a.H + n0 + c0 + g * 4) = make_float4(u[0], u[1], u[2], u[3]);
          }
        }
      }
      if (last) {
        if (m < M) a.Fpart[(size_t)(rank * 2 + half) * a.cap + m] = dot;
      } else {
        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy global stores -> visible to the TMA loads of the next layer
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    it_base += nk;
    if (l + 1 < a.n_hidden) {
      cluster_sync_all();   // the whole 128 x H row block of this layer is in L2 and TMEM has been drained
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace mlp

// Host launcher used by trace_batched.cu.  Returns IRONB_ENOSUP when the shape does not fit (caller falls back to
// the per-layer GEMM path).
int launch_trace_mlp_fused(const ironb_mlp_layout* lay, const float* packed, const CUtensorMap& mE, const CUtensorMap mU[2],
                           const CUtensorMap* mW, const float* E, float* const U[2], float* Fpart, int rows_cap, int cap,
                           const int* m_dev, int m_mul, cudaStream_t st) {
  using namespace mlp;
  const int H = lay->d_hidden, last = lay->n_lin - 1;
  const int C = H / 128;
  if (H % 128 != 0 || !(C == 1 || C == 2 || C == 4) || last > MAXL) return IRONB_ENOSUP;
  static Maps maps;   // ~1.5 KiB, copied into the launch's parameter space
  maps.e = mE; maps.u[0] = mU[0]; maps.u[1] = mU[1];
  Args a;
  memset(&a, 0, sizeof(a));
  for (int l = 0; l < last; ++l) {
    maps.w[l] = mW[l];
    a.bias[l] = packed + lay->off_b[l];
    a.n_true[l] = lay->out_dim[l];
    a.kpad[l] = lay->in_pad[l];
  }
  a.w_last = packed + lay->off_w[last];
  a.E = E; a.U[0] = U[0]; a.U[1] = U[1]; a.Fpart = Fpart;
  a.n_hidden = last; a.skip_layer = lay->skip_layer; a.Epad = lay->in_pad[0]; a.Edim = lay->pe_dim; a.H = H;
  a.beta = lay->beta; a.inv_beta = 1.0f / lay->beta;
  a.m_dev = m_dev; a.m_mul = m_mul; a.rows_cap = rows_cap; a.cap = cap;
  a.write_hi = tc::split_writes_hi() ? 1 : 0;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) { set_error("mlp_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)C, (unsigned)ceil_div64(rows_cap, BM), 1);
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_fused_kernel, maps, a);
  note_launch();
  if (e != cudaSuccess) { set_error("mlp_fused launch: %s", cudaGetErrorString(e)); return (int)e; }
  return IRONB_OK;
}

}  // namespace ironb
