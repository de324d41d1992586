// Fused SDF-MLP evaluation for the batched tracer, fp16x2-split arithmetic, two row tiles in flight per cluster.
//
// Decomposition: a cluster of C = H/128 CTAs owns 128-row tiles, CTA rank r computes output features
// [128 r, 128 r + 128) of every hidden layer, activations are exchanged through L2.  Against round 1's 3xTF32 predecessor
// (mlp_tc.cu, retired in round 2) two changes roughly triple the throughput:
//
//   * ARITHMETIC.  Every fp32 operand x is split into hi = fp16_rn(x) and lo = fp16_rn((x - hi) * 2^11): two 11-bit
//     mantissas, the same 22 significant bits as the 3xTF32 split, but as 16-bit operands.  Three
//     tcgen05.mma.kind::f16 per 16-wide k-step compute hi*hi (exact products, fp32 accumulate) and lo*hi + hi*lo (kept in
//     their own accumulator, scaled by 2^11; the epilogue folds them in with one fp32 FMA).  kind::f16 runs at twice the
//     kind::tf32 rate and the operands are half the bytes, so both the tensor pipe and the L2 -> shared-memory stream
//     (which bounds the mainloop) cost half of what 3xTF32 costs.  The scaling keeps lo out of the fp16 subnormal range;
//     operands must stay below 65504 in magnitude (softplus activations and weight-normalised rows are O(1)).
//   * SCHEDULE.  With one hi*hi and one cross-term accumulator a tile needs 256 of the 512 TMEM columns, so a cluster
//     keeps TWO row tiles in flight: while the 16 epilogue warps drain layer l of tile A (tcgen05.ld, bias + softplus,
//     split, swizzled staging, TMA store), the MMA warp already runs layer l of tile B.  Nothing in the kernel is a
//     CTA-wide or cluster-wide barrier any more; every hand-off is an mbarrier:
//         full / empty [3]   TMA loads  <-> MMA                       (64 KiB stages: A_hi, B_hi, A_lo, B_lo)
//         acc_full [slot]    MMA -> epilogue                          (tcgen05.commit)
//         tmem_free [slot]   epilogue -> MMA                          (accumulators drained)
//         stg_full[half] / stg_empty   epilogue <-> the two store warps  (32 KiB staging: one 64-column half, hi + lo box)
//         ready [slot][half] store warps of ALL CTAs of the cluster -> TMA loader (remote arrive; acquire at cluster
//                            scope): one 64-column half (= one k-block of the next layer) of every CTA's 128 x 128
//                            activation block is in L2.  Measured alternatives: a release.cluster arrive per peer costs a
//                            GPU-scope membar each (4.7k cycles for 4 peers) -> ONE fence + relaxed arrives; plain
//                            st.global from the epilogue warps + per-warp fences was slower still (16 membars per half).
//     The loader prefetches the weight (B) boxes of a layer's first stages before it waits for `ready`, so only the
//     activation (A) boxes sit on the layer-to-layer critical path.
//
// nhh = 2 (diagnostic): hi*hi alternates between two accumulators (shorter truncating chains, see gemm_tc.cuh); a tile
// then needs 384 columns and the cluster runs one tile at a time.
#include <cuda_fp16.h>

#include "gemm_tc.cuh"

namespace ironb {
namespace mlp16 {

using namespace tc;

constexpr int MAXH = 8;
constexpr int RM = 128;                       // rows per tile
constexpr int BKH = 64;                       // K elements per stage: one 128-byte SWIZZLE_128B row of halfs
constexpr int TILE = 128 * 128;               // bytes of one 128-row operand box (128 rows x 128 B)
constexpr int STG = 2 * TILE;
constexpr int NWARP_EPI = 16;
constexpr int NCTRL = 4;                      // warp 0 TMA loads, warp 1 MMA + TMEM alloc, warps 2 / 3 TMA stores of half 0 / 1
constexpr int NT = (NCTRL + NWARP_EPI) * 32;  // warps 4.. epilogue
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f, RSQRT2 = 0.70710678118654752f;

// Per-variant geometry.  RN = output features per CTA: 128 (clusters of H/128 CTAs: the throughput shape, least L2 -> SM
// traffic per row) or 64 (clusters of H/64 CTAs: the LATENCY shape).  One 128-row tile's layer costs a CTA
// 96 x (RN/128) x 64 cycles of MMA issue plus an epilogue over RN columns; with RN = 64 both halve, so the layer-to-layer
// chain of a tile -- which is all that matters when a launch holds no more tiles than the GPU has clusters, i.e. for the
// 4,096-ray training patch and for every late tracer round -- drops from ~15k to ~10k cycles, while two tiles in flight per
// cluster (2 x 2 x 64 TMEM columns) keep the tensor pipe of each SM busy during the other tile's hand-off.
template <int RN>
struct Geo {
  static constexpr int TILE_B = RN * 128;                     // bytes of one weight box (RN rows x 128 B)
  static constexpr int STAGE = 2 * TILE + 2 * TILE_B;         // [A_hi][B_hi][A_lo][B_lo]
  static constexpr int NSTAGE = (RN == 128) ? 3 : 4;          // 192 KiB of operand ring either way
  static constexpr int HP = RN / 64;                          // 64-column halves (= k-blocks of the next layer) per CTA
  static constexpr int OFF_BH = TILE, OFF_AL = TILE + TILE_B, OFF_BL = 2 * TILE + TILE_B;
  static constexpr int SMEM = NSTAGE * STAGE + STG + 1024 + 256;
  // kind::f16: fp16 A and B (format 0), fp32 accumulate, both K-major, M = 128, N = RN
  static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(RN >> 3) << 17) | ((uint32_t)(RM >> 4) << 24);
};

struct Maps {
  CUtensorMap e[2];          // encoded points  [cap][Epad]   (hi, lo) fp16
  CUtensorMap u[2][2];       // activation ping-pong [cap][H]  [buffer][hi, lo] fp16
  CUtensorMap w[MAXH][2];    // hidden-layer weights W_l [H][K_l] [layer][hi, lo] fp16, (64 x RN) boxes
};

struct Args {
  const float* bias[MAXH];
  const float* w_last;       // row 0 of the output layer [H], fp32
  const __half* Ehi; const __half* Elo;
  __half* Uhi[2]; __half* Ulo[2];   // activation ping-pong buffers [cap][H] (the maps' memory)
  float* Fpart;              // [4C][cap] partial sdf sums
  int n_true[MAXH];
  int kpad[MAXH];
  float gam[MAXH];           // truncation de-bias of the hi*hi accumulator, per layer (see mlp16_debias)
  int n_hidden, skip_layer, Epad, Edim, H, C;
  float nbl2e, ln2_ib;       // -beta * log2(e),  ln(2) / beta
  const int* m_dev;
  int m_mul, rows_cap, cap, nhh, preb;
  long long* dbg;
};

// softplus_beta(z) = max(z, 0) + log1p(exp(-|beta z|)) / beta with the two MUFU approximations used directly:
// t = 2^(-|z| beta log2 e), L = log2(1 + t) in [0, 1], result = max(z, 0) + L ln2 / beta.  The log term is <= ln2 / beta, so
// the 2^-22 relative error of ex2 / lg2 is ~1e-9 absolute; for beta z > 20 the term vanishes in fp32 (torch's threshold).
__device__ __forceinline__ float softplus_fast(float z, float nbl2e, float ln2_ib) {
  float t, L;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(z) * nbl2e));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(L) : "f"(1.f + t));
  return fmaf(L, ln2_ib, fmaxf(z, 0.f));
}
__device__ __forceinline__ void split_h(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * LO_SCALE);
}
__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// wait on an mbarrier other CTAs of the cluster arrive on: acquire at cluster scope
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, uint32_t tag = 0, uint32_t extra = 0) {
  uint32_t ok;
#ifdef IRONB_DEBUG_HANG
  unsigned long long polls = 0;
#endif
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
#ifdef IRONB_DEBUG_HANG
    if (!ok && ++polls > (1ull << 26)) hang_report(tag, parity, extra);
#endif
  } while (!ok);
  (void)tag; (void)extra;
}
// Remote arrive WITHOUT its own release: every release.cluster arrive costs a GPU-scope membar (~1k cycles, measured
// 4 of them back to back), so the caller issues ONE fence.acq_rel.cluster and then these relaxed arrives.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t local_bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(local_bar),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct EpiCtx {
  int q, blk, lane, row, n0, rank, M;
  uint32_t tmem, stg_full0, stg_empty;
  unsigned char *srow_hi, *srow_lo;
};

// One (tile, layer) unit of the epilogue for one thread: row = TMEM lane, 16 columns in each of the two 64-column halves.
// PRE_SKIP: the layer before the skip connection (cat(h, PE) / sqrt 2 fills the columns past n_true); LAST: the last hidden
// layer is not stored, it is dotted with the sdf row of the output layer.
template <int RN, bool PRE_SKIP, bool LAST>
__device__ __forceinline__ void epi_unit(const Args& a, const EpiCtx& c, int l, int s, int m0, float mybias, float mywl,
                                         uint32_t tfree, uint32_t& se_n) {
  constexpr int HP = Geo<RN>::HP;
  const int m = m0 + c.row;
  const int n_true = a.n_true[l];
  const float gam = a.gam[l];
  float dot = 0.f;
#pragma unroll
  for (int h = 0; h < HP; ++h) {
    const int c0 = h * 64 + c.blk * 16;                   // column inside the CTA's RN
    uint32_t r0[16], rl[16];
    const uint32_t taddr = c.tmem + ((uint32_t)(c.q * 32) << 16) + (uint32_t)s * (uint32_t)(RN * (a.nhh + 1)) + (uint32_t)c0;
    tmem_ld16(taddr, r0);
    if (a.nhh == 2) {
      uint32_t r1[16];
      tmem_ld16(taddr + (uint32_t)RN, r1);
      tmem_ld16(taddr + 2u * (uint32_t)RN, rl);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) r0[i] = __float_as_uint(__fadd_rn(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
    } else {
      tmem_ld16(taddr + (uint32_t)RN, rl);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    if (h == HP - 1) {                                    // this warp has drained its part of the slot's accumulators
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (c.lane == 0) mbar_arrive(tfree);
    }
    uint32_t ph[8], pl[8];                                // 16 outputs as packed halfs: hi and scaled lo
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float u[2];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int i = g * 2 + jj;
        const float b = __shfl_sync(0xffffffffu, mybias, h * 16 + i);
        const float hh = __uint_as_float(r0[i]);
        const float z = fmaf(__uint_as_float(rl[i]), LO_INV, fmaf(hh, gam, hh)) + b;
        u[jj] = softplus_fast(z, a.nbl2e, a.ln2_ib);
      }
      if (PRE_SKIP) {   // cat(h, PE)/sqrt(2)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int n = c.n0 + c0 + g * 2 + jj;
          if (n < n_true) {
            u[jj] *= RSQRT2;                             // the accumulators already carry ~1e-6: no need for the exact division
          } else {
            const int ce = n - n_true;
            u[jj] = 0.f;
            if (ce < a.Edim && m < c.M) {
              const size_t o = (size_t)m * a.Epad + ce;
              u[jj] = fmaf(__half2float(a.Elo[o]), LO_INV, __half2float(a.Ehi[o])) * RSQRT2;
            }
          }
        }
      }
      if (LAST) {
        dot = fmaf(u[0], __shfl_sync(0xffffffffu, mywl, h * 16 + g * 2), dot);
        dot = fmaf(u[1], __shfl_sync(0xffffffffu, mywl, h * 16 + g * 2 + 1), dot);
      } else {
        const __half2 hi2 = __floats2half2_rn(u[0], u[1]);
        const float2 hf = __half22float2(hi2);
        const __half2 lo2 = __floats2half2_rn((u[0] - hf.x) * LO_SCALE, (u[1] - hf.y) * LO_SCALE);
        ph[g] = *reinterpret_cast<const uint32_t*>(&hi2);
        pl[g] = *reinterpret_cast<const uint32_t*>(&lo2);
      }
    }
    if (!LAST) {
      mbar_wait(c.stg_empty, (se_n & 1u) ^ 1u, 37, (uint32_t)(l * 4 + s * 2 + h));           // the previous half's TMA store has read the staging buffer
      ++se_n;
#pragma unroll
      for (int g = 0; g < 2; ++g) {                       // two 16-byte chunks (8 halfs each) per operand
        const int off = (((c.blk * 2 + g) ^ (c.row & 7)) << 4);
        *reinterpret_cast<uint4*>(c.srow_hi + off) = make_uint4(ph[g * 4], ph[g * 4 + 1], ph[g * 4 + 2], ph[g * 4 + 3]);
        *reinterpret_cast<uint4*>(c.srow_lo + off) = make_uint4(pl[g * 4], pl[g * 4 + 1], pl[g * 4 + 2], pl[g * 4 + 3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged rows -> visible to the TMA store engine
      __syncwarp();
      if (c.lane == 0) mbar_arrive(c.stg_full0 + 8u * (uint32_t)h);
    }
  }
  if (LAST && m < c.M) a.Fpart[(size_t)(c.rank * 4 + c.blk) * a.cap + m] = dot;
}

// The unit sequence every role walks: tile pairs (j = 0: t0, j = 1: t0 + G) of this cluster, layer by layer.  A pair keeps
// each tile on its own TMEM slot (s = j).  A cluster's last, unpaired tile alternates the two slots BETWEEN LAYERS
// (s = l & 1): the next layer's MMAs may then start on the first half of the activations while the epilogue still drains
// the other slot.  sp is the slot the previous layer of the same tile used (whose `ready` barriers the loader waits on).
#define MLP16_FOR_UNITS                                                     \
  for (int t0 = cid; t0 < ntiles; t0 += nslots * G)                         \
    for (int l = 0; l < a.n_hidden; ++l)                                    \
      for (int j = 0; j < nslots; ++j)                                      \
        if (t0 + j * G < ntiles)
#define MLP16_UNIT_SLOTS                                                    \
  const bool two = (nslots == 2) && (t0 + G < ntiles);                      \
  const int s = two ? j : ((nslots == 2) ? (l & 1) : 0);                    \
  const int sp = two ? j : ((nslots == 2) ? ((l + 1) & 1) : 0);             \
  (void)sp;

template <int RN>
__global__ void __launch_bounds__(NT, 1) mlp_h16_kernel(const __grid_constant__ Maps maps, const Args a) {
  typedef Geo<RN> GEO;
  constexpr int STAGE = GEO::STAGE, NSTAGE = GEO::NSTAGE, HP = GEO::HP;
  int M = a.rows_cap;
  if (a.m_dev != nullptr) {
    const int md = *a.m_dev * a.m_mul;
    if (md < M) M = md;
  }
  const int G = gridDim.y, cid = blockIdx.y;
  if (cid * RM >= M) return;                              // uniform over the cluster (same blockIdx.y)
  const int ntiles = (M + RM - 1) / RM;
  const int rank = blockIdx.x, n0 = rank * RN;
  const int nslots = (a.nhh == 1) ? 2 : 1;
  const uint32_t ncol_slot = (uint32_t)(RN * (a.nhh + 1));

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - raw);
  const uint32_t stg = base + NSTAGE * STAGE;
  unsigned char* stg_ptr = base_ptr + NSTAGE * STAGE;
  const uint32_t bars = stg + STG;
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (NSTAGE + i); };
  auto acc_full = [&](int s) { return bars + 8u * (2 * NSTAGE + s); };
  auto tmem_free = [&](int s) { return bars + 8u * (2 * NSTAGE + 2 + s); };
  auto ready = [&](int s, int h) { return bars + 8u * (2 * NSTAGE + 4 + s * 2 + h); };   // (slot, 64-column half)
  auto stg_full = [&](int h) { return bars + 8u * (2 * NSTAGE + 8 + h); };
  const uint32_t stg_empty = bars + 8u * (2 * NSTAGE + 10);
  const uint32_t tmem_slot = bars + 8u * (2 * NSTAGE + 11);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(stg_ptr + STG + 8 * (2 * NSTAGE + 11));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(full(i), 1);
      mbar_init(empty(i), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(tmem_free(s), NWARP_EPI);
      mbar_init(ready(s, 0), (uint32_t)a.C);
      mbar_init(ready(s, 1), (uint32_t)a.C);
    }
    mbar_init(stg_full(0), NWARP_EPI);
    mbar_init(stg_full(1), NWARP_EPI);
    mbar_init(stg_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                     // barriers initialised in every CTA before any remote arrive
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot_ptr;
  const bool stamp = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;

  if (warp == 0) {
    // ================= TMA loads (warp-uniform loop, elected lane issues) =================
    const bool leader = elect_one();
    uint32_t it = 0, rdy_bits = 0u;                     // phase parity per (slot, half) ready barrier
    MLP16_FOR_UNITS {
      MLP16_UNIT_SLOTS
      const int m0 = (t0 + j * G) * RM;
      const int nk = (a.kpad[l] + BKH - 1) / BKH;
      const CUtensorMap* mAh = (l == 0) ? &maps.e[0] : &maps.u[(l - 1) & 1][0];
      const CUtensorMap* mAl = (l == 0) ? &maps.e[1] : &maps.u[(l - 1) & 1][1];
      const CUtensorMap* mBh = &maps.w[l][0];
      const CUtensorMap* mBl = &maps.w[l][1];
      // k-blocks in the order the previous layer's epilogues release them: every CTA's first 64-column half
      // (k-blocks 0, 2, 4, ..), then every CTA's second half.  Layer 0 reads the encoded points: no dependency.
      const int C = a.C;
      int ib = 0;                                         // stages whose barrier is armed and whose weight boxes are issued
      for (int ia = 0; ia < nk; ++ia) {
        const bool wait_pt = (l > 0) && (ia % C == 0) && (ia / C < HP);
        const int ahead = wait_pt ? (ia + a.preb < nk ? ia + a.preb : nk) : ia + 1;
        for (; ib < ahead; ++ib) {                        // weights do not depend on the previous layer: run ahead
          const uint32_t sidx = (it + ib) % NSTAGE, ph = ((it + ib) / NSTAGE) & 1u;
          const int kb = (l == 0) ? ib : HP * (ib % C) + (ib / C);
          mbar_wait(empty(sidx), ph ^ 1u, 31, (uint32_t)(l * 64 + j * 32 + ib));
          const uint32_t st = base + sidx * STAGE;
          if (leader) {
            mbar_arrive_expect_tx(full(sidx), STAGE);
            tma_load_2d(st + GEO::OFF_BH, mBh, kb * BKH, n0, full(sidx));
            tma_load_2d(st + GEO::OFF_BL, mBl, kb * BKH, n0, full(sidx));
          }
        }
        if (wait_pt) {
          const int h = ia / C;
          mbar_wait_cluster(ready(sp, h), (rdy_bits >> (sp * 2 + h)) & 1u, 32, (uint32_t)(l * 4 + sp * 2 + h));
          rdy_bits ^= 1u << (sp * 2 + h);
        }
        const uint32_t sidx = (it + ia) % NSTAGE;
        const int kb = (l == 0) ? ia : HP * (ia % C) + (ia / C);
        const uint32_t st = base + sidx * STAGE;
        if (leader) {
          tma_load_2d(st, mAh, kb * BKH, m0, full(sidx));
          tma_load_2d(st + GEO::OFF_AL, mAl, kb * BKH, m0, full(sidx));
        }
      }
      it += nk;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (warp-uniform loop, elected lane issues) =================
    const bool leader = elect_one();
    uint32_t it = 0, free_n[2] = {0u, 0u};
    const uint32_t nhh = (uint32_t)a.nhh;
    MLP16_FOR_UNITS {
      MLP16_UNIT_SLOTS
      const int K = a.kpad[l];
      const int nk = (K + BKH - 1) / BKH, ksteps = (K + 15) / 16;
      mbar_wait(tmem_free(s), (free_n[s] & 1u) ^ 1u, 33, (uint32_t)(l * 2 + j));
      ++free_n[s];
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem + (uint32_t)s * ncol_slot;
      const uint32_t acc_lo = acc + (uint32_t)RN * nhh;
      for (int i = 0; i < nk; ++i) {
        const uint32_t sidx = (it + i) % NSTAGE, ph = ((it + i) / NSTAGE) & 1u;
        if (stamp && leader && i == 0 && t0 == 0 && j == 0) a.dbg[l * 8 + 0] = clock64();
        mbar_wait(full(sidx), ph, 34, (uint32_t)(l * 64 + j * 32 + i));
        if (stamp && leader && i == 0 && t0 == 0 && j == 0) a.dbg[l * 8 + 1] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = base + sidx * STAGE;
        const uint64_t a_hi = make_desc(st), b_hi = make_desc(st + GEO::OFF_BH);
        const uint64_t a_lo = make_desc(st + GEO::OFF_AL), b_lo = make_desc(st + GEO::OFF_BL);
        const int kmax = ksteps - i * (BKH / 16);         // k-steps of this stage that hold data (layer 0: K = 40)
#pragma unroll
        for (int kk = 0; kk < BKH / 16; ++kk) {
          const uint32_t g = (uint32_t)(i * (BKH / 16) + kk);
          const uint64_t adv = (uint64_t)(kk * 2);        // 16 halfs = 32 B = 2 x 16 B along the swizzled row
          const uint32_t acc_hh = acc + ((nhh == 2u && (g & 1u)) ? (uint32_t)RN : 0u);
          if (leader && kk < kmax) {
            tc_mma_f16(acc_hh, a_hi + adv, b_hi + adv, GEO::IDESC, g >= nhh ? 1u : 0u);
            tc_mma_f16(acc_lo, a_lo + adv, b_hi + adv, GEO::IDESC, g >= 1u ? 1u : 0u);
            tc_mma_f16(acc_lo, a_hi + adv, b_lo + adv, GEO::IDESC, 1u);
          }
        }
        if (leader) tc_commit(empty(sidx));
      }
      if (leader) tc_commit(acc_full(s));
      if (stamp && leader && t0 == 0 && j == 0) a.dbg[l * 8 + 2] = clock64();
      it += nk;
    }
    __syncwarp();
  } else if (warp < NCTRL) {
    // ================= TMA stores + cluster-wide "half ready": warp 2 owns the first 64-column half of every unit, warp 3
    // the second, so waiting for one half's completion never delays the other's issue (bulk groups are per thread)
    const int h = warp - 2;
    const bool leader = elect_one();
    uint32_t sf_n = 0;
    if (h < HP)
    MLP16_FOR_UNITS {
      if (l == a.n_hidden - 1) continue;
      MLP16_UNIT_SLOTS
      const int m0 = (t0 + j * G) * RM;
      const CUtensorMap* mh = &maps.u[l & 1][0];
      const CUtensorMap* ml = &maps.u[l & 1][1];
      mbar_wait(stg_full(h), sf_n & 1u, 35, (uint32_t)(l * 4 + j * 2 + h));
      ++sf_n;
      if (leader) {
        if (stamp && t0 == 0 && j == 0 && h == 0) a.dbg[l * 8 + 5] = clock64();
        tma_store_2d(mh, stg, n0 + h * 64, m0);
        tma_store_2d(ml, stg + TILE, n0 + h * 64, m0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // staging may be overwritten
        mbar_arrive(stg_empty);
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");          // this half (one k-block of the next layer) is in L2
        asm volatile("fence.acq_rel.cluster;" ::: "memory");               // ONE release for all peers ...
        for (int r = 0; r < a.C; ++r) mbar_arrive_remote_relaxed(ready(s, h), (uint32_t)r);   // ... then relaxed arrives
        if (stamp && t0 == 0 && j == 0) a.dbg[l * 8 + 6 + h] = clock64();
      }
      __syncwarp();
    }
    __syncwarp();
  } else {
    // ================= epilogue: 16 warps = 4 TMEM lane quarters x 4 blocks of 16 columns, two 64-column halves =================
    EpiCtx c;
    c.q = warp & 3; c.blk = (warp - NCTRL) >> 2; c.lane = lane;
    c.row = c.q * 32 + lane;                              // row inside the tile == TMEM lane
    c.tmem = tmem; c.n0 = n0; c.rank = rank; c.M = M;
    c.srow_hi = stg_ptr + c.row * 128; c.srow_lo = c.srow_hi + TILE;
    c.stg_full0 = stg_full(0); c.stg_empty = stg_empty;
    uint32_t acc_n[2] = {0u, 0u}, se_n = 0;
    MLP16_FOR_UNITS {
      MLP16_UNIT_SLOTS
      const int m0 = (t0 + j * G) * RM;
      const bool last = (l == a.n_hidden - 1);
      const bool pre_skip = (l + 1 == a.skip_layer);
      // bias (and, for the last layer, the sdf-row weights) of this warp's 2 x 16 columns: one value per lane, fetched
      // BEFORE the accumulator wait so the global-load latency hides behind the mainloop; broadcast by shuffle below
      const int mycol = n0 + (lane >> 4) * 64 + c.blk * 16 + (lane & 15);
      const bool mine = (lane >> 4) < HP;                   // RN = 64: lanes 16..31 hold no column
      const float mybias = mine ? __ldg(a.bias[l] + mycol) : 0.f;
      const float mywl = (last && mine) ? __ldg(a.w_last + mycol) : 0.f;
      mbar_wait(acc_full(s), acc_n[s] & 1u, 36, (uint32_t)(l * 2 + j));
      ++acc_n[s];
      const bool st = stamp && threadIdx.x == NCTRL * 32 && t0 == 0 && j == 0;
      if (st) a.dbg[l * 8 + 3] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tfree = tmem_free(s);
      if (last) epi_unit<RN, false, true>(a, c, l, s, m0, mybias, mywl, tfree, se_n);
      else if (pre_skip) epi_unit<RN, true, false>(a, c, l, s, m0, mybias, mywl, tfree, se_n);
      else epi_unit<RN, false, false>(a, c, l, s, m0, mybias, mywl, tfree, se_n);
      if (st) a.dbg[l * 8 + 4] = clock64();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                     // no CTA leaves while a peer may still arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace mlp16

long long* g_mlp_dbg = nullptr;   // optional clock64 stamp buffer (ironb_debug_mlp_timeline), shared with gemm_tc.cuh

bool trace_mlp_fused_supported(const ironb_mlp_layout* lay) {
  const int H = lay->d_hidden, C = H / 128;
  return H % 128 == 0 && (C == 1 || C == 2 || C == 4) && lay->n_lin - 1 <= mlp16::MAXH && lay->n_lin >= 2;
}

// The tensor core adds into its TMEM accumulator with TRUNCATION (measured: tests/probe_precision.py, probe_mlp_h16.py --
// the error of one evaluation is a pure bias, signed mean == -mean |err|, proportional to the chain length: -0.99e-6 at
// H = 256, -2.12e-6 at H = 512), i.e. every one of the K/16 accumulation steps of the hi*hi chain drops on average half
// an ulp of the running sum, always towards zero.  A round-to-nearest accumulator would drop +-half an ulp with zero
// mean, so the expected loss is added back: acc *= 1 + g * n_steps * 2^-24.  Calibrated on the box
// (tests/probe_debias.py, profiles/r2_mlp_h16_debias.md): g = 0.25 takes the signed mean error of one evaluation from
// -2.10e-6 to -1.4e-8 (H = 512) and from -0.99e-6 to -1.2e-7 (H = 256), on the seed-0 init and on perturbed weights alike,
// and the mean |error| from 2.1e-6 to 8.9e-8 -- the level of the exact-fp32 FFMA tracer (9.1e-8): the loss was almost
// entirely the deterministic bias.  IRONB_MLP_DEBIAS / ironb_set_mlp_debias override it, 0 switches it off.  The
// correction is ~5e-7 relative at H = 512, so a network whose statistics differ from the calibration stays fp32-grade.
static float g_debias = -1.f;
float mlp16_debias() {
  if (g_debias < 0.f) {
    const char* e = getenv("IRONB_MLP_DEBIAS");
    g_debias = e ? (float)atof(e) : 0.25f;
    if (!(g_debias >= 0.f && g_debias <= 4.f)) g_debias = 0.f;
  }
  return g_debias;
}

float mlp16_set_debias(float g) {
  const float prev = mlp16_debias();
  if (g >= 0.f && g <= 4.f) g_debias = g;
  return prev;
}

static int g_nhh = -1;
int mlp16_nhh() {
  if (g_nhh < 0) {
    const char* e = getenv("IRONB_MLP_NHH");
    g_nhh = (e && e[0] == '2') ? 2 : 1;
  }
  return g_nhh;
}

// Co-resident clusters of C CTAs of variant RN on this device (GPC packing decides, not SMs / C); < 0: the shape cannot be
// scheduled.  First use also raises the kernel's dynamic shared-memory limit.
template <int RN>
static int resident_clusters(int C) {
  using namespace mlp16;
  static int resident[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (C < 1 || C > 8) return -1;
  if (resident[C] == 0) {
    int nc = -1;
    if (cudaFuncSetAttribute(mlp_h16_kernel<RN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<RN>::SMEM) == cudaSuccess) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cfg.blockDim = dim3(NT, 1, 1);
      cfg.dynamicSmemBytes = Geo<RN>::SMEM;
      cfg.gridDim = dim3((unsigned)C, (unsigned)(num_sms() / C), 1);
      if (cudaOccupancyMaxActiveClusters(&nc, mlp_h16_kernel<RN>, &cfg) != cudaSuccess || nc < 1) nc = -1;
    }
    (void)cudaGetLastError();
    const char* ov = getenv("IRONB_MLP_CLUSTERS");
    if (nc > 0 && ov && atoi(ov) > 0) nc = atoi(ov);
    resident[C] = nc;
  }
  return resident[C];
}

// Narrow (RN = 64, clusters of H/64 CTAs) or wide (RN = 128, clusters of H/128 CTAs) variant for a tracer call over `rays`
// rays: the narrow one wins while a launch cannot fill the GPU's clusters with two tiles each, i.e. for small patches.
// IRONB_MLP_RN=64 / 128 forces one (if schedulable).
static int g_rn_force = -1;
int mlp16_pick_rn(const ironb_mlp_layout* lay, int64_t rays) {
  if (g_rn_force < 0) {
    const char* e = getenv("IRONB_MLP_RN");
    g_rn_force = (e && atoi(e) == 64) ? 64 : (e && atoi(e) == 128) ? 128 : 0;
  }
  const int H = lay->d_hidden;
  const bool narrow_ok = H % 64 == 0 && H / 64 <= 8 && resident_clusters<64>(H / 64) > 0;
  const bool wide_ok = H % 128 == 0 && resident_clusters<128>(H / 128) > 0;
  if (g_rn_force == 64 && narrow_ok) return 64;
  if (g_rn_force == 128 && wide_ok) return 128;
  // Measured (tests/probe_mlp_rn.py, profiles/r2_mlp_h16_rn.md): a tile's layer is bound by the SM's L2 ingest (~64 B/clk:
  // 512 KiB of operands per CTA and layer at RN = 128, 384 KiB at RN = 64), not by the tensor pipe, so the narrow shape
  // only pays while every tile has a cluster of its own (H = 512: 15 resident 8-CTA clusters -> at most 1,920 rays;
  // H = 256: 33 4-CTA clusters); beyond that its 1.5x operand traffic per row loses (3.6 vs 2.4 ms at 4,096 rays, H = 512).
  if (narrow_ok && (!wide_ok || rays <= (int64_t)mlp16::RM * resident_clusters<64>(H / 64))) return 64;
  return 128;
}

int mlp16_set_rn(int rn) {
  const int prev = g_rn_force < 0 ? 0 : g_rn_force;
  g_rn_force = (rn == 64 || rn == 128) ? rn : 0;
  return prev;
}

template <int RN>
static int launch_variant(const mlp16::Maps& maps, mlp16::Args& a, int C, int rows_cap, cudaStream_t st) {
  using namespace mlp16;
  typedef Geo<RN> GEO;
  { static int preb = -1;
    if (preb < 0) { const char* e = getenv("IRONB_MLP_PREB"); preb = (e && atoi(e) >= 1 && atoi(e) <= GEO::NSTAGE) ? atoi(e) : GEO::NSTAGE; }
    a.preb = preb; }
  const int res = resident_clusters<RN>(C);
  if (res < 0) return IRONB_ENOSUP;                         // this cluster shape cannot be scheduled here
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.dynamicSmemBytes = GEO::SMEM;
  cfg.stream = st;
  const int64_t tiles = ceil_div64(rows_cap, RM);
  cfg.gridDim = dim3((unsigned)C, (unsigned)(tiles < res ? tiles : res), 1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_h16_kernel<RN>, maps, a);
  note_launch();
  if (e != cudaSuccess) { set_error("mlp_h16<%d> launch: %s", RN, cudaGetErrorString(e)); return (int)e; }
  return IRONB_OK;
}

// maps: e[2], u[2][2] with (64 x 128) boxes; w[n_hidden][2] (hi, lo) with (64 x rn) boxes: the caller builds them for the
// variant `rn` it got from mlp16_pick_rn.  Returns IRONB_ENOSUP if that variant's cluster shape cannot be scheduled.
int launch_trace_mlp_h16(const ironb_mlp_layout* lay, const float* packed, const CUtensorMap* mE, const CUtensorMap* mU,
                         const CUtensorMap* mW, const void* Ehi, const void* Elo, void* const* Uhi, void* const* Ulo,
                         float* Fpart, int rows_cap, int cap, const int* m_dev, int m_mul, int rn, cudaStream_t st) {
  using namespace mlp16;
  if (!trace_mlp_fused_supported(lay)) return IRONB_ENOSUP;
  if (rn != 64 && rn != 128) return IRONB_EINVAL;
  const int H = lay->d_hidden, last = lay->n_lin - 1, C = H / rn;
  if (C < 1 || C > 8) return IRONB_ENOSUP;
  static Maps maps;
  maps.e[0] = mE[0]; maps.e[1] = mE[1];
  for (int b = 0; b < 2; ++b) { maps.u[b][0] = mU[b * 2]; maps.u[b][1] = mU[b * 2 + 1]; }
  Args a;
  memset(&a, 0, sizeof(a));
  for (int l = 0; l < last; ++l) {
    maps.w[l][0] = mW[l * 2]; maps.w[l][1] = mW[l * 2 + 1];
    a.bias[l] = packed + lay->off_b[l];
    a.n_true[l] = lay->out_dim[l];
    a.kpad[l] = lay->in_pad[l];
    if ((lay->in_pad[l] + 15) / 16 < 2) return IRONB_ENOSUP;
  }
  a.w_last = packed + lay->off_w[last];
  a.Ehi = reinterpret_cast<const __half*>(Ehi); a.Elo = reinterpret_cast<const __half*>(Elo);
  for (int b = 0; b < 2; ++b) { a.Uhi[b] = reinterpret_cast<__half*>(Uhi[b]); a.Ulo[b] = reinterpret_cast<__half*>(Ulo[b]); }
  a.Fpart = Fpart;
  a.n_hidden = last; a.skip_layer = lay->skip_layer; a.Epad = lay->in_pad[0]; a.Edim = lay->pe_dim; a.H = H; a.C = C;
  a.nbl2e = (float)(-(double)lay->beta * 1.4426950408889634); a.ln2_ib = (float)(0.6931471805599453 / (double)lay->beta);
  a.m_dev = m_dev; a.m_mul = m_mul; a.rows_cap = rows_cap; a.cap = cap;
  a.nhh = mlp16_nhh();
  for (int l = 0; l < last; ++l)
    a.gam[l] = mlp16_debias() * (float)((lay->in_pad[l] + 15) / 16) / (float)a.nhh * 5.9604645e-8f;   // g * n_steps * 2^-24
  a.dbg = g_mlp_dbg;
  return rn == 64 ? launch_variant<64>(maps, a, C, rows_cap, st) : launch_variant<128>(maps, a, C, rows_cap, st);
}

}  // namespace ironb

extern "C" float ironb_set_mlp_debias(float g) { return ironb::mlp16_set_debias(g); }
extern "C" int ironb_set_mlp_rn(int rn) { return ironb::mlp16_set_rn(rn); }
// diagnostic: co-resident clusters of the MLP kernel variant `rn` for hidden width H on the current device (< 0: unschedulable)
extern "C" int ironb_debug_mlp_resident_clusters(int rn, int H) {
  if (rn == 64) return ironb::resident_clusters<64>(H / 64);
  if (rn == 128) return ironb::resident_clusters<128>(H / 128);
  return -1;
}

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

namespace ironb { IRONB_DEFINE_HANG_SETTER(hang_set_mlp) }
