// Fused SDF-MLP evaluation for the batched tracer: ALL hidden layers of a 128-row tile in ONE persistent launch.
//
// A cluster of C = H/128 CTAs owns tiles of 128 work items (rows) and walks them persistently.  CTA rank r computes
// output features [128 r, 128 r + 128) of every layer on the tensor cores with the 3xTF32 split:
//
//   * operands live in global memory / L2 ALREADY split into tf32-exact hi and lo parts (weights: split once per
//     trace call; activations: split by the epilogue that produces them; encoded points: split by the state-machine
//     kernel that writes them), so the mainloop is the plain two-role Blackwell pipeline
//         TMA (4 SWIZZLE_128B boxes per stage: A_hi, B_hi, A_lo, B_lo)  ->  mbarrier  ->  tcgen05.mma.kind::tf32 x 3
//     with no shared-memory pass between them;
//   * accumulation is spread over three TMEM accumulators (hi*hi on even / odd k-steps, the lo cross terms) because the
//     tensor core accumulates with truncation (see gemm_tc.cuh); the epilogue adds them in fp32 round-to-nearest;
//   * the epilogue (8 warps: tcgen05.ld, bias + softplus [+ skip concat]) writes the layer's 128 x 128 activation block
//     as hi / lo to a global ping-pong buffer, fences it to the async proxy, and a cluster barrier makes the whole
//     128 x H row block visible before any CTA of the cluster TMA-loads it as the next layer's A operand.
//
// The last hidden layer never stores its activations: its epilogue dots them with row 0 of the output layer (the sdf row)
// and writes per-(rank, 32-column block) partial sums; the tracer's state-machine kernel adds the 4C partials in a fixed
// order (deterministic) and applies bias / scale.
#include "gemm_tc.cuh"

namespace ironb {
namespace mlp {

using namespace tc;

constexpr int MAXH = 8;                     // hidden layers the fused kernel supports
constexpr int NT = 576;                     // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..17 epilogue (4 per scheduler)
constexpr int NEPI = 512;
constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256 + 2 * BN * 4;

struct Maps {
  CUtensorMap e[2];          // encoded points  [cap][Epad]           (hi, lo)
  CUtensorMap u[2][2];       // activation ping-pong [cap][H]          [buffer][hi, lo]
  CUtensorMap w[MAXH][2];    // hidden-layer weights W_l [H][K_l]      [layer][hi, lo]
};

struct Args {
  const float* bias[MAXH];
  const float* w_last;       // row 0 of the output layer [H]
  const float* Ehi; const float* Elo;
  float* Uhi[2]; float* Ulo[2];
  float* Fpart;              // [4C][cap] partial sdf sums
  int n_true[MAXH];
  int kpad[MAXH];
  int n_hidden, skip_layer, Epad, Edim, H;
  float beta, inv_beta;
  const int* m_dev;
  int m_mul, rows_cap, cap;
  long long* dbg;            // optional clock64 stamps of cluster 0 / rank 0 (IRONB_MLP_DBG)
};

__device__ __forceinline__ float softplus_fast(float z, float beta, float inv_beta) {
  // max(z,0) + log1p(exp(-|beta z|))/beta: the log term is <= ln2 and is scaled by 1/beta, so ex2/lg2.approx
  // contribute ~1e-9 absolute; for beta*z > 20 the term vanishes in fp32, which is torch's threshold branch.
  const float e = __expf(-fabsf(z * beta));
  return fmaxf(z, 0.f) + __logf(1.f + e) * inv_beta;
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
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
  if ((int)blockIdx.y * BM >= M) return;                 // uniform over the cluster (same blockIdx.y)
  const int ntiles = (M + BM - 1) / BM;
  const int rank = blockIdx.x, n0 = rank * BN;

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t acc_bar = bars + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sbias = reinterpret_cast<float*>(base_ptr + STAGES * STAGE_BYTES + 256);   // [128] bias of this CTA's columns
  float* swl = sbias + BN;                                                          // [128] sdf-row weights (last layer)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 1);
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

  int it_base = 0;       // pipeline iterations consumed so far (every role advances it identically)
  uint32_t acc_phase = 0;
  for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y) {
    const int m0 = tile * BM;
    for (int l = 0; l < a.n_hidden; ++l) {
      const int K = a.kpad[l];
      const int nk = (K + BK - 1) / BK;
      if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
          const CUtensorMap* mAh = (l == 0) ? &maps.e[0] : &maps.u[(l - 1) & 1][0];
          const CUtensorMap* mAl = (l == 0) ? &maps.e[1] : &maps.u[(l - 1) & 1][1];
          const CUtensorMap* mBh = &maps.w[l][0];
          const CUtensorMap* mBl = &maps.w[l][1];
          for (int i = 0; i < nk; ++i) {
            const int it = it_base + i, s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(empty(s), ph ^ 1);
            mbar_arrive_expect_tx(full(s), STAGE_BYTES);
            const uint32_t st = base + s * STAGE_BYTES;
            tma_load_2d(st, mAh, i * BK, m0, full(s));
            tma_load_2d(st + TILE_BYTES, mBh, i * BK, n0, full(s));
            tma_load_2d(st + 2 * TILE_BYTES, mAl, i * BK, m0, full(s));
            tma_load_2d(st + 3 * TILE_BYTES, mBl, i * BK, n0, full(s));
          }
        }
        __syncwarp();
      } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
          for (int i = 0; i < nk; ++i) {
            const int it = it_base + i, s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            if (a.dbg && i == 0 && blockIdx.x == 0 && blockIdx.y == 0 && tile == 0) a.dbg[l * 8 + 0] = clock64();
            mbar_wait(full(s), ph);
            if (a.dbg && i == 0 && blockIdx.x == 0 && blockIdx.y == 0 && tile == 0) a.dbg[l * 8 + 1] = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t st = base + s * STAGE_BYTES;
            const uint64_t a_hi = make_desc(st), b_hi = make_desc(st + TILE_BYTES);
            const uint64_t a_lo = make_desc(st + 2 * TILE_BYTES), b_lo = make_desc(st + 3 * TILE_BYTES);
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 2);
              const int g = i * (BK / 8) + kk;
              // hi*hi rotates over three accumulators (shorter truncating chains), the lo cross terms use the fourth
              tc_mma_tf32(tmem + 128u * (uint32_t)(g % 3), a_hi + adv, b_hi + adv, IDESC, g >= 3 ? 1u : 0u);
              tc_mma_tf32(tmem + 384u, a_lo + adv, b_hi + adv, IDESC, g >= 1 ? 1u : 0u);
              tc_mma_tf32(tmem + 384u, a_hi + adv, b_lo + adv, IDESC, 1u);
            }
            tc_commit(empty(s));
          }
          tc_commit(acc_bar);
          if (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tile == 0) a.dbg[l * 8 + 2] = clock64();
        }
        __syncwarp();
      } else {
        // ================= epilogue (16 warps: 4 TMEM lane quarters x 4 blocks of 32 columns) =================
        const int t = threadIdx.x - 64;
        const bool last = (l == a.n_hidden - 1);
        const bool pre_skip = (l + 1 == a.skip_layer);
        const int n_true = a.n_true[l];
        // stage this layer's bias (and the sdf row) for the CTA's 128 columns while the mainloop runs
        if (t < BN) {
          const int n = n0 + t;
          sbias[t] = (n < n_true) ? __ldg(a.bias[l] + n) : 0.f;
          if (last) swl[t] = __ldg(a.w_last + n);
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        mbar_wait(acc_bar, acc_phase);
        if (a.dbg && threadIdx.x == 64 && blockIdx.x == 0 && blockIdx.y == 0 && tile == 0) a.dbg[l * 8 + 3] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, blk = (warp - 2) >> 2;
        const int row = q * 32 + lane;          // row inside the tile == TMEM lane
        const int m = m0 + row;
        float dot = 0.f;
        // output staging in the (now idle) pipeline stages, in the SWIZZLE_128B box layout the TMA store expects
        unsigned char* stg_hi = base_ptr + blk * TILE_BYTES + row * 128;
        unsigned char* stg_lo = stg_hi + 4 * TILE_BYTES;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int c0 = blk * 32 + h * 16;     // 16 columns at a time keeps the 16-warp epilogue under 112 registers
          uint32_t r0[16], r1[16], r2[16], r3[16];
          const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
          tmem_ld16(taddr, r0);
          tmem_ld16(taddr + 128u, r1);
          tmem_ld16(taddr + 256u, r2);
          tmem_ld16(taddr + 384u, r3);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 b4 = *reinterpret_cast<const float4*>(sbias + c0 + g * 4);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
            float u[4], hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = g * 4 + j;
              const float z = __fadd_rn(__fadd_rn(__fadd_rn(__uint_as_float(r0[i]), __uint_as_float(r1[i])), __uint_as_float(r2[i])),
                                        __uint_as_float(r3[i]));
              u[j] = softplus_fast(z + bb[j], a.beta, a.inv_beta);
            }
            if (pre_skip) {   // cat(h, PE)/sqrt(2): warp-uniform branch, only the layer before the skip takes it
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int n = n0 + c0 + g * 4 + j;
                if (n < n_true) {
                  u[j] = __fdiv_rn(u[j], IRONB_SQRT2F);
                } else {
                  const int ce = n - n_true;
                  u[j] = 0.f;
                  if (ce < a.Edim && m < M) {
                    const size_t o = (size_t)m * a.Epad + ce;
                    u[j] = __fdiv_rn(__ldg(a.Ehi + o) + __ldg(a.Elo + o), IRONB_SQRT2F);
                  }
                }
              }
            }
            if (last) {
              const float4 w4 = *reinterpret_cast<const float4*>(swl + c0 + g * 4);
              dot = fmaf(u[0], w4.x, dot); dot = fmaf(u[1], w4.y, dot); dot = fmaf(u[2], w4.z, dot); dot = fmaf(u[3], w4.w, dot);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) split1(u[j], hi[j], lo[j]);
              const int off = (((h * 4 + g) ^ (row & 7)) << 4);   // 16-byte chunk of the 128-byte row, XOR-swizzled
              *reinterpret_cast<float4*>(stg_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<float4*>(stg_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
        if (last) {
          if (m < M) a.Fpart[(size_t)(rank * 4 + blk) * a.cap + m] = dot;
        } else {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged tile -> visible to the TMA store engine
          asm volatile("bar.sync 1, 512;" ::: "memory");
          if (t == 0) {
            const CUtensorMap* mh = &maps.u[l & 1][0];
            const CUtensorMap* ml = &maps.u[l & 1][1];
#pragma unroll
            for (int bk = 0; bk < 4; ++bk) {
              tma_store_2d(mh, base + bk * TILE_BYTES, n0 + bk * 32, m0);
              tma_store_2d(ml, base + (4 + bk) * TILE_BYTES, n0 + bk * 32, m0);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete (and smem free) before the barrier
          }
        }
        if (a.dbg && threadIdx.x == 64 && blockIdx.x == 0 && blockIdx.y == 0 && tile == 0) a.dbg[l * 8 + 4] = clock64();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      it_base += nk;
      acc_phase ^= 1u;
      // the layer's whole 128 x H row block is in L2 and this CTA's TMEM accumulators have been drained
      cluster_sync_all();
      if (a.dbg && threadIdx.x == 64 && blockIdx.x == 0 && blockIdx.y == 0 && tile == 0) a.dbg[l * 8 + 5] = clock64();
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

// W (fp32) -> tf32-exact hi / lo copies, once per trace call
__global__ void __launch_bounds__(256) split_array_kernel(const float* __restrict__ src, int64_t n, float* __restrict__ hi,
                                                          float* __restrict__ lo) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float h, l;
  split1(src[i], h, l);
  hi[i] = h;
  lo[i] = l;
}

}  // namespace mlp

long long* g_mlp_dbg = nullptr;   // shared with mlp_h16.cu

bool trace_mlp_fused_supported(const ironb_mlp_layout* lay) {
  const int H = lay->d_hidden, C = H / 128;
  return H % 128 == 0 && (C == 1 || C == 2 || C == 4) && lay->n_lin - 1 <= mlp::MAXH && lay->n_lin >= 2;
}

int split_weights(const float* src, int64_t n, float* hi, float* lo, cudaStream_t st) {
  mlp::split_array_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(src, n, hi, lo);
  IRONB_CHECK_LAUNCH("split_array_kernel");
  return IRONB_OK;
}

// maps: e[2], u[2][2], w[n_hidden][2] (hi, lo).  Pointers: the same buffers (epilogue stores / skip concat loads).
int launch_trace_mlp_fused(const ironb_mlp_layout* lay, const float* packed, const CUtensorMap* mE, const CUtensorMap* mU,
                           const CUtensorMap* mW, const float* Ehi, const float* Elo, float* const* Uhi, float* const* Ulo,
                           float* Fpart, int rows_cap, int cap, const int* m_dev, int m_mul, cudaStream_t st) {
  using namespace mlp;
  if (!trace_mlp_fused_supported(lay)) return IRONB_ENOSUP;
  const int H = lay->d_hidden, last = lay->n_lin - 1, C = H / 128;
  static Maps maps;   // ~2.8 KiB, copied into the launch's parameter space
  maps.e[0] = mE[0]; maps.e[1] = mE[1];
  for (int b = 0; b < 2; ++b) { maps.u[b][0] = mU[b * 2]; maps.u[b][1] = mU[b * 2 + 1]; }
  Args a;
  memset(&a, 0, sizeof(a));
  for (int l = 0; l < last; ++l) {
    maps.w[l][0] = mW[l * 2]; maps.w[l][1] = mW[l * 2 + 1];
    a.bias[l] = packed + lay->off_b[l];
    a.n_true[l] = lay->out_dim[l];
    a.kpad[l] = lay->in_pad[l];
  }
  a.w_last = packed + lay->off_w[last];
  a.Ehi = Ehi; a.Elo = Elo;
  a.Uhi[0] = Uhi[0]; a.Uhi[1] = Uhi[1]; a.Ulo[0] = Ulo[0]; a.Ulo[1] = Ulo[1];
  a.Fpart = Fpart;
  a.n_hidden = last; a.skip_layer = lay->skip_layer; a.Epad = lay->in_pad[0]; a.Edim = lay->pe_dim; a.H = H;
  a.beta = lay->beta; a.inv_beta = 1.0f / lay->beta;
  a.m_dev = m_dev; a.m_mul = m_mul; a.rows_cap = rows_cap; a.cap = cap;
  a.dbg = g_mlp_dbg;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) { set_error("mlp_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  int64_t tiles = ceil_div64(rows_cap, BM);
  int64_t resident = num_sms() / C;      // persistent: one cluster per C SMs, each walks its row tiles
  if (resident < 1) resident = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)C, (unsigned)(tiles < resident ? tiles : resident), 1);
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

// debugging aid: IRONB_MLP_DBG timeline of the fused MLP kernel (cluster 0, rank 0, first tile), 8 stamps per layer
extern "C" int ironb_debug_mlp_timeline(long long* host_out, int n) {
  if (ironb::g_mlp_dbg == nullptr) {
    if (cudaMalloc(&ironb::g_mlp_dbg, 64 * 8 * sizeof(long long)) != cudaSuccess) return -1;
    cudaMemset(ironb::g_mlp_dbg, 0, 64 * 8 * sizeof(long long));
    return 0;
  }
  if (host_out && n > 0) cudaMemcpy(host_out, ironb::g_mlp_dbg, (size_t)(n < 512 ? n : 512) * sizeof(long long), cudaMemcpyDeviceToHost);
  return 1;
}
