// tcgen05 (5th-gen tensor core) tile GEMM with fp32-grade accuracy:  C[m][n] = epi( sum_k A[m][k] * B[n][k] )
//
//   * A [M][lda] and B [N][ldb] are fp32, K-major, and stay ONE fp32 copy in global memory / L2.
//   * TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages 128 x 32 fp32 boxes of A and B into a 3-stage
//     shared-memory ring (mbarrier complete_tx).
//   * four "splitter" warps rewrite each staged tile in place as hi = tf32_rn(x) and write lo = tf32_rn(x - hi)
//     next to it (same swizzled offsets, so no address math), then fence to the async proxy.
//   * one thread issues, per 8-wide k-step, THREE tcgen05.mma.kind::tf32 (A_lo*B_hi + A_hi*B_lo + A_hi*B_hi): the
//     3xTF32 split whose products are exact and whose sum is accumulated in fp32 in TMEM (128 lanes x 128 cols).
//     tcgen05.commit releases the ring slot back to the TMA producer and finally publishes the accumulator.
//   * the splitter warps then become the epilogue: tcgen05.ld 32x32b.x32 (TMEM -> registers), and the same
//     epilogue functors as the SIMT GEMM (bias / softplus / skip concat / backward elementwise) write the result.
//
// Why split in shared memory instead of pre-splitting in HBM: 3xTF32 operands are 8 B per element; at 128 x 128
// tiles that is ~84 B/clk/SM of L2 traffic, above the measured L2 slice throughput, while one fp32 copy is 42 B/clk.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"

// programmatic-dependent-launch hooks (no-ops unless built with -DIRONB_ENABLE_PDL; see gemm_tc.cu: measured slower)
#ifdef IRONB_ENABLE_PDL
#define IRONB_GDC_LAUNCH() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define IRONB_GDC_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#else
#define IRONB_GDC_LAUNCH() ((void)0)
#define IRONB_GDC_WAIT() ((void)0)
#endif

namespace ironb {
extern long long* g_mlp_dbg;   // mlp_tc.cu: optional clock64 stamp buffer (ironb_debug_mlp_timeline)
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;            // 16 KiB: one operand tile (128 rows x 128 B)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;        // [A_hi][B_hi][A_lo][B_lo]
constexpr int HI_BYTES = 2 * TILE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers + tmem ptr*/;
constexpr int NTHREADS = 320;                      // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2-9 split + epilogue
constexpr int NSPLIT = 256;
constexpr int NGRP = 2;                            // splitter warps work as NGRP groups on alternate stages, so one group's
constexpr int NSPLIT_G = NSPLIT / NGRP;            // proxy fence + barrier latency overlaps the other group's conversion
constexpr uint32_t TMEM_COLS = 512;               // three 128-column accumulators (even k-steps, odd k-steps, lo-terms)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// -DIRONB_DEBUG_HANG: every mbarrier wait is bounded.  A wait that polls for ~2 s writes a record {tag, block, thread,
// parity, extra} through d_hang_buf (a pointer to MAPPED PINNED host memory, still readable after the context is lost) and
// traps, so a protocol slip becomes a located fault instead of a silent 100 %-busy hang.  ironb_debug_hang_buffer() installs
// the buffer in every translation unit (each has its own copy of the static symbol).  Tags: 1x gemm_nt_tc (11 empty,
// 12 conv, 13 full, 14 acc, 15 conv of the other group), 2x gemm_h16 (21 empty, 22 full, 24 acc), 3x mlp_h16 (31 empty, 32 ready, 33 tmem_free, 34 full,
// 35 stg_full, 36 acc_full, 37 stg_empty).
#ifdef IRONB_DEBUG_HANG
static __device__ unsigned long long* d_hang_buf = nullptr;
__device__ __forceinline__ void hang_report(uint32_t tag, uint32_t parity, uint32_t extra) {
  unsigned long long* b = d_hang_buf;
  if (b != nullptr) {
    const unsigned long long slot = atomicAdd(b, 1ull);
    if (slot < 63) {
      unsigned long long* r = b + 8 + slot * 8;
      r[0] = tag; r[1] = blockIdx.x | ((unsigned long long)blockIdx.y << 20) | ((unsigned long long)blockIdx.z << 40);
      r[2] = threadIdx.x; r[3] = parity; r[4] = extra; r[5] = gridDim.x | ((unsigned long long)gridDim.y << 20);
      r[6] = clock64();
      __threadfence_system();
    }
  }
  __trap();
}
#endif
#ifdef IRONB_DEBUG_HANG
#define IRONB_DEFINE_HANG_SETTER(name) \
  int name(unsigned long long* p) { return (int)cudaMemcpyToSymbol(::ironb::tc::d_hang_buf, &p, sizeof(p)); }
#else
#define IRONB_DEFINE_HANG_SETTER(name) \
  int name(unsigned long long*) { return 0; }
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag = 0, uint32_t extra = 0) {
  uint32_t ok;
#ifdef IRONB_DEBUG_HANG
  unsigned long long polls = 0;
#endif
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// one lane of a converged warp: control warps run their loops warp-uniformly and predicate only the issue on it
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Round-to-nearest split: hi = tf32_rna(x), lo = tf32_rna(x - hi).  x - hi is exact in fp32 and SIGNED, so the term the
// 3xTF32 product drops (lo*lo) is zero-mean; rounding a magnitude half-up to 11 significant bits is one integer add + mask.
__device__ __forceinline__ void split1(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xffffe000u);
}
// Truncating split for operands that stay raw fp32 in shared memory (kind::tf32 ignores the 13 low mantissa bits, i.e. the
// tensor core itself sees hi = trunc(x)): lo = tf32_rna(x - trunc(x)).  Cheaper (no hi write-back), but lo has the sign of
// x, so the dropped lo*lo term is a small systematic underestimate of |a*b|.
__device__ __forceinline__ void split1_trunc(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xffffe000u);
}
// Split `per_thread` 16-byte chunks (stride nthreads) of a staged [A|B] tile pair in place (hi) and into lo.
// write_hi == 0 leaves the raw fp32 in the hi tile and relies on kind::tf32 ignoring the 13 low mantissa bits.
__device__ __forceinline__ void split_stage(float4* __restrict__ hi, float4* __restrict__ lo, int t, int nthreads,
                                            int per_thread, int write_hi) {
#pragma unroll 4
  for (int j = 0; j < per_thread; ++j) {
    const int c = t + j * nthreads;
    const float4 v = hi[c];
    float4 h, l;
    if (write_hi) {
      split1(v.x, h.x, l.x); split1(v.y, h.y, l.y); split1(v.z, h.z, l.z); split1(v.w, h.w, l.w);
      hi[c] = h;
    } else {
      split1_trunc(v.x, h.x, l.x); split1_trunc(v.y, h.y, l.y); split1_trunc(v.z, h.z, l.z); split1_trunc(v.w, h.w, l.w);
    }
    lo[c] = l;
  }
}

template <class Epi>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_nt_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int M, int N, int K,
                  Epi epi, const int* __restrict__ m_dev, int m_mul, int write_hi, int k_chunk, long long* dbg, int strict) {
  const bool stamp = dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (stamp && threadIdx.x == 0) dbg[0] = clock64();
  // Programmatic dependent launch: the next kernel of the stream may start its prologue now; this kernel touches global
  // memory only after griddepcontrol.wait (the predecessor has completed and its writes are visible).
  IRONB_GDC_LAUNCH();
  if (m_dev != nullptr) {   // the device-side row count is produced by a predecessor
    IRONB_GDC_WAIT();   // row count produced on the device (compacted ray lists): whole CTAs beyond it leave at once
    const int md = *m_dev * m_mul;
    if (md < M) M = md;
  }
  if ((int)blockIdx.y * BM >= M) return;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B tiles need 1024 B alignment
  unsigned char* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * STAGE_BYTES;                  // full[3] conv[3] empty[3] acc[1] | tmem ptr
  auto full = [&](int s) { return bars + 8u * s; };
  auto conv = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  const uint32_t acc_bar = bars + 8u * (3 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // split-K (weight gradients): blockIdx.z owns k in [kbeg, kend), k_chunk a multiple of BK; the epilogue accumulates
  const int kbeg = (k_chunk > 0) ? (int)blockIdx.z * k_chunk : 0;
  const int kend = (k_chunk > 0) ? min(K, kbeg + k_chunk) : K;
  const int nk = (kend - kbeg + BK - 1) / BK;
  if (nk <= 0) return;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), NSPLIT_G);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: 128 columns x 128 lanes of fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot_ptr;
  IRONB_GDC_WAIT();        // barriers, TMEM and tensor maps are set up: now wait for the data
  if (stamp && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0) {
    // ================= TMA producer (warp-uniform loop: addresses stay in uniform registers; elected lane issues) ====
    const bool leader = elect_one();
    for (int it = 0; it < nk; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(empty(s), ph ^ 1, 11, it);
      const uint32_t st = base + s * STAGE_BYTES;
      if (leader) {
        mbar_arrive_expect_tx(full(s), HI_BYTES);
        tma_load_2d(st, &mapA, kbeg + it * BK, m0, full(s));
        tma_load_2d(st + TILE_BYTES, &mapB, kbeg + it * BK, n0, full(s));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (warp-uniform loop, elected lane issues) =================
    const bool leader = elect_one();
    for (int it = 0; it < nk; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(conv(s), ph, 12, it);
      if (stamp && leader && it == 0) dbg[3] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t st = base + s * STAGE_BYTES;
      const uint64_t a_hi = make_desc(st), b_hi = make_desc(st + TILE_BYTES);
      const uint64_t a_lo = make_desc(st + 2 * TILE_BYTES), b_lo = make_desc(st + 3 * TILE_BYTES);
#pragma unroll
      for (int kk = 0; kk < BK / 8; ++kk) {
        const uint64_t adv = (uint64_t)(kk * 2);          // 8 tf32 = 32 B = 2 x 16 B along the swizzled row
        const int g = it * (BK / 8) + kk;                 // global k-step
        // The tensor core accumulates into TMEM with truncation, so every accumulation step costs up to 1 ulp
        // of the running sum, always towards zero.  Keep the chains short and the small terms apart:
        // hi*hi alternates between two accumulators, the two lo cross terms go to a third; the epilogue adds
        // the three in fp32 round-to-nearest.
        if (leader) {
          tc_mma_tf32(tmem + ((g & 1) ? 128u : 0u), a_hi + adv, b_hi + adv, IDESC, g >= 2 ? 1u : 0u);
          tc_mma_tf32(tmem + 256u, a_lo + adv, b_hi + adv, IDESC, g >= 1 ? 1u : 0u);
          tc_mma_tf32(tmem + 256u, a_hi + adv, b_lo + adv, IDESC, 1u);
        }
      }
      if (leader) tc_commit(empty(s));          // slot free once these MMAs have read it
    }
    if (leader) tc_commit(acc_bar);             // accumulator complete
    if (stamp && leader) dbg[4] = clock64();
    __syncwarp();
  } else {
    // ================= splitter, then epilogue (warps 2..9) =================
    const int t = threadIdx.x - 64;   // 0..255
    const int grp = t / NSPLIT_G, tg = t % NSPLIT_G;
    for (int it = grp; it < nk; it += NGRP) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      // A parity wait is exact only if the barrier cannot be a whole phase BEHIND the one waited for: try_wait.parity also
      // succeeds when `ph` is the parity of the phase that precedes the current one.  The two splitter groups take alternate
      // iterations, so this group last saw stage s two uses ago; the use in between (iteration it - STAGES) belongs to the
      // OTHER group.  If this group gets here before that iteration's TMA load has landed (loads of different stages can
      // complete out of order under memory contention, e.g. next to another stream's grid), full(s) still sits in the phase
      // of iteration it - STAGES and a wait on parity `ph` passes at once: the group would split a stage that is still being
      // written, complete conv(s) a phase early and desynchronise the pipeline (wrong sums, a hang, or a TMA load landing
      // in the shared memory of an exited CTA = "unspecified launch failure": the multi-stream fault of round 1, root cause).
      // conv(s) of iteration it - STAGES completes only after that iteration's load has landed and cannot advance further
      // without this group's own arrivals, so waiting for it first makes the wait on full(s) exact.  It costs nothing: the
      // load of iteration `it` is only issued after the MMAs of it - STAGES, which need that conv(s) phase anyway.
      if (strict && it >= STAGES) mbar_wait(conv(s), ph ^ 1, 15, it);
      mbar_wait(full(s), ph, 13, it);
      if (stamp && t == 0 && it == 0) dbg[2] = clock64();
      float4* hi = reinterpret_cast<float4*>(base_ptr + s * STAGE_BYTES);
      float4* lo = reinterpret_cast<float4*>(base_ptr + s * STAGE_BYTES + HI_BYTES);
      split_stage(hi, lo, tg, NSPLIT_G, HI_BYTES / 16 / NSPLIT_G, write_hi);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      mbar_arrive(conv(s));
    }
    // epilogue, phase 1: TMEM (lane = tile row; warp%4 selects the 32-lane quarter, (warp-2)/4 the 64-column half) ->
    // registers (the three accumulators are added in fp32 RN) -> a row-major fp32 tile in the idle pipeline stages
    if (stamp && t == 0) dbg[5] = clock64();
    mbar_wait(acc_bar, 0, 14, nk);
    if (stamp && t == 0) dbg[6] = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr int TLD = BN + 4;                       // padded row pitch: conflict-free 128-bit rows
    float* tile = reinterpret_cast<float*>(base_ptr);
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = half * 64 + cc * 32;
      uint32_t r[32], r1[32], r2[32];
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      tmem_ld32(taddr, r);
      tmem_ld32(taddr + 128u, r1);
      tmem_ld32(taddr + 256u, r2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[j] = __fadd_rn(__fadd_rn(__uint_as_float(r[g * 4 + j]), __uint_as_float(r1[g * 4 + j])), __uint_as_float(r2[g * 4 + j]));
        *reinterpret_cast<float4*>(tile + row * TLD + c0 + g * 4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (stamp && t == 0) dbg[7] = clock64();
    // phase 2: one warp per tile row, lanes along the columns: every global access of the epilogue functor (bias,
    // saved activations, outputs) is a coalesced 512-byte row segment
    const int w8 = warp - 2;
#pragma unroll 2
    for (int rr = 0; rr < BM / 8; ++rr) {
      const int trow = w8 * (BM / 8) + rr;
      const int m = m0 + trow;
      const int n = n0 + lane * 4;
      if (m < M && n < N) {
        const float4 v4 = *reinterpret_cast<const float4*>(tile + trow * TLD + lane * 4);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        epi(m, n, v);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (stamp && threadIdx.x == 0) dbg[8] = clock64();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side: tensor maps through the driver entry point (no link-time libcuda dependency) ----
typedef CUtensorMap CUtensorMapAlias;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();
// rows x K fp32 matrix with row pitch ld floats -> 2-D map with a (BK x 128) SWIZZLE_128B box, zero fill out of bounds
int make_map(CUtensorMap* map, const float* ptr, int rows, int K, int ld);
int make_map_h(CUtensorMap* map, const void* ptr, int rows, int K, int ld, int box_rows = 128);   // fp16, 64 x box_rows box
bool tc_enabled();
bool pdl_enabled();       // programmatic dependent launch of the GEMM kernels (IRONB_PDL=1 switches it on; see gemm_tc.cu)
bool split_strict();      // 1 (default): exact phase tracking of the splitter groups; IRONB_SPLIT_STRICT=0 restores the round-1 race (diagnostic)
bool split_writes_hi();   // 0 (default): raw fp32 stays as the hi operand; 1 (IRONB_SPLIT_WRITE_HI=1): store the truncated hi back

template <class Epi>
int launch_gemm_nt_tc_maps(const CUtensorMap& mA, const CUtensorMap& mB, int M, int N, int K, const Epi& epi,
                           const int* m_dev, int m_mul, cudaStream_t st, const char* what, int write_hi = -1, int k_chunk = 0) {
  if (write_hi < 0) write_hi = split_writes_hi() ? 1 : 0;
  if (M <= 0 || N <= 0) return IRONB_OK;
  auto kern = gemm_nt_tc_kernel<Epi>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  dim3 grid((unsigned)ceil_div64(N, BN), (unsigned)ceil_div64(M, BM), (unsigned)(k_chunk > 0 ? ceil_div64(K, k_chunk) : 1));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;          // no attribute at all unless PDL is switched on
  long long* dbg = g_mlp_dbg ? g_mlp_dbg + 256 : nullptr;   // IRONB debug timeline (last launch wins)
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, mA, mB, M, N, K, epi, m_dev, m_mul, write_hi, k_chunk, dbg, split_strict() ? 1 : 0);
  if (le != cudaSuccess) { (void)cudaGetLastError(); note_launch(); set_error("%s: %s", what, cudaGetErrorString(le)); return (int)le; }
  IRONB_CHECK_LAUNCH(what);
  return IRONB_OK;
}

template <class Epi>
int launch_gemm_nt_tc(const float* A, int lda, const float* B, int ldb, int M, int N, int K, const Epi& epi, cudaStream_t st,
                      const char* what, int write_hi = -1) {
  if (M <= 0 || N <= 0) return IRONB_OK;
  CUtensorMap mA, mB;
  int rc = make_map(&mA, A, M, K, lda);
  if (rc) return rc;
  rc = make_map(&mB, B, N, K, ldb);
  if (rc) return rc;
  return launch_gemm_nt_tc_maps(mA, mB, M, N, K, epi, nullptr, 1, st, what, write_hi);
}

}  // namespace tc

// fp32 FFMA tiles (exact association) or tcgen05 3xTF32 tiles, per ironb_set_gemm_mode / IRONB_GEMM.
// ---- weight gradients on the tensor cores:  C[n][k] += sum_m A[m][n] * B[m][k] ---------------------------------------
// tcgen05.mma.kind::tf32 wants K-major 32-bit operands here (the MN-major path with the 16-byte SWIZZLE_128B atom produced
// nothing on sm_100a), so both operands are first transposed into scratch ([cols][round4(M)], zero padded) by a coalesced
// smem-tile transpose, then the K-major GEMM above runs split over the row range with an atomically accumulating epilogue.
struct EpiAtomicAdd {
  float* C;
  int ldc;
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[4]) const {
    float* dst = C + (size_t)m * ldc + n0;    // 16-byte aligned: ldc and n0 are multiples of 4
    // one 128-bit reduction (REDG.E.ADD.F32x4) instead of four scalar atomics: the split-K epilogues of one weight gradient
    // issue 2.4 M atomic operations on 262 k addresses otherwise
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(acc[0]), "f"(acc[1]), "f"(acc[2]), "f"(acc[3])
                 : "memory");
  }
};

// Both operands of a weight gradient in one launch (blockIdx.z = 0: A, 1: B):  dst[c][m] = src[m][c] for c < cols, m < M;
// dst row pitch ldt >= round4(M), columns m in [M, ldt) zeroed.  Optionally the column sums of A (the bias gradient of the
// same layer: csum[c] += scale * sum_m A[m][c], c < csum_cols) are taken from the tile while it sits in shared memory.
static __global__ void __launch_bounds__(256) transpose2_kernel(const float* __restrict__ A, int lda, int colsA,
                                                                float* __restrict__ At, const float* __restrict__ B, int ldb,
                                                                int colsB, float* __restrict__ Bt, int M, int ldt,
                                                                float* __restrict__ csum, int csum_cols, float csum_scale) {
  __shared__ float tile[32][33];
  IRONB_GDC_LAUNCH();   // the split-K GEMM that follows may set itself up
  const bool second = blockIdx.z != 0;
  const float* __restrict__ src = second ? B : A;
  float* __restrict__ dst = second ? Bt : At;
  const int lds = second ? ldb : lda, cols = second ? colsB : colsA;
  const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (c0 >= cols) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + i * 8, c = c0 + tx;
    tile[ty + i * 8][tx] = (m < M && c < cols) ? src[(size_t)m * lds + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, m = m0 + tx;
    if (c < cols && m < ldt) dst[(size_t)c * ldt + m] = tile[tx][ty + i * 8];
  }
  if (!second && csum != nullptr && ty == 0 && c0 + tx < csum_cols && m0 < M) {
    float t = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) t += tile[r][tx];
    atomicAdd(csum + c0 + tx, t * csum_scale);
  }
}

inline int64_t wgrad_scratch_floats(int64_t M, int max_cols) { return 2 * (int64_t)max_cols * ((M + 3) / 4 * 4); }

// scratch: wgrad_scratch_floats(M, max(Nd, Kd)) floats
inline int launch_wgrad_tc(const float* A, int lda, const float* B, int ldb, int M, int Nd, int Kd, float* C, int ldc,
                           float* scratch, cudaStream_t st, const char* what, float* csum = nullptr, int csum_cols = 0,
                           float csum_scale = 1.f) {
  if (M <= 0 || Nd <= 0 || Kd <= 0) return IRONB_OK;
  const int ldt = (M + 3) / 4 * 4;
  float* At = scratch;
  float* Bt = scratch + (size_t)(Nd > Kd ? Nd : Kd) * ldt;
  dim3 g2((unsigned)ceil_div64(ldt, 32), (unsigned)ceil_div64(Nd > Kd ? Nd : Kd, 32), 2);
  transpose2_kernel<<<g2, 256, 0, st>>>(A, lda, Nd, At, B, ldb, Kd, Bt, M, ldt, csum, csum_cols, csum_scale);
  IRONB_CHECK_LAUNCH("transpose2_kernel");
  tc::CUtensorMapAlias mA, mB;
  int rc = tc::make_map(&mA, At, Nd, ldt, ldt);
  if (rc) return rc;
  rc = tc::make_map(&mB, Bt, Kd, ldt, ldt);
  if (rc) return rc;
  const int64_t tiles = ceil_div64(Nd, tc::BM) * ceil_div64(Kd, tc::BN);
  int64_t splits = num_sms() / tiles;
  const int64_t maxs = ceil_div64(M, 128);
  if (splits > maxs) splits = maxs;
  if (splits < 1) splits = 1;
  const int k_chunk = (int)(ceil_div64(ceil_div64(ldt, splits), tc::BK) * tc::BK);
  EpiAtomicAdd ep{C, ldc};
  return tc::launch_gemm_nt_tc_maps(mA, mB, Nd, Kd, ldt, ep, nullptr, 1, st, what, -1, k_chunk);
}

// csum (optional): the bias gradient of the same layer, csum[c] += csum_scale * sum_m A[m][c] for c < csum_cols
inline int launch_wgrad_auto(const float* A, int lda, const float* B, int ldb, int M, int Nd, int Kd, float* C, int ldc,
                             float* scratch, cudaStream_t st, const char* what, float* csum = nullptr, int csum_cols = 0,
                             float csum_scale = 1.f) {
  static const bool wgrad_simt = [] { const char* e = getenv("IRONB_WGRAD"); return e && (e[0] == 's' || e[0] == 'S'); }();   // diagnostic
  if (tc::tc_enabled() && scratch != nullptr && !wgrad_simt)
    return launch_wgrad_tc(A, lda, B, ldb, M, Nd, Kd, C, ldc, scratch, st, what, csum, csum_cols, csum_scale);
  int rc = launch_gemm_tn(A, lda, B, ldb, M, Nd, Kd, C, ldc, st, what);
  if (rc || csum == nullptr) return rc;
  return launch_colsum(A, lda, M, csum_cols, csum_scale, csum, st, what);
}

// Fork / join between the chain's stream and the weight-gradient stream: a cycling pool of timing-free events (an event may
// be re-recorded once the wait that used it has been ENQUEUED; under stream capture both calls become graph edges).
struct EventPool {
  cudaEvent_t ev[128];
  int n = 0, next = 0;
  cudaEvent_t get() {
    if (n < 128) {
      if (cudaEventCreateWithFlags(&ev[n], cudaEventDisableTiming) != cudaSuccess) return nullptr;
      return ev[n++];
    }
    next = (next + 1) % 128;
    return ev[next];
  }
};
inline int fork_to(cudaStream_t from, cudaStream_t to) {     // `to` continues after everything enqueued on `from` so far
  if (from == to) return IRONB_OK;
  static thread_local EventPool pool;
  cudaEvent_t e = pool.get();
  if (e == nullptr) { set_error("cudaEventCreate failed"); return IRONB_EINVAL; }
  IRONB_CUDA(cudaEventRecord(e, from));
  IRONB_CUDA(cudaStreamWaitEvent(to, e, 0));
  return IRONB_OK;
}

template <class Epi>
int launch_gemm_nt_auto(const float* A, int lda, const float* B, int ldb, int M, int N, int K, const Epi& epi,
                        cudaStream_t st, const char* what) {
  if (tc::tc_enabled()) return tc::launch_gemm_nt_tc(A, lda, B, ldb, M, N, K, epi, st, what);
  return launch_gemm_nt(A, lda, B, ldb, M, N, K, epi, st, what);
}
}  // namespace ironb
