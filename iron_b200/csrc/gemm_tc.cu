// Host side of the tcgen05 GEMM: TMA tensor-map encoding through the driver entry point, the GEMM mode switch,
// and a plain C = A * B^T entry used by the unit tests of both GEMM implementations.
#include <stdlib.h>

#include <atomic>

#include "gemm.cuh"
#include "gemm_h16.cuh"

namespace ironb {
namespace tc {

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map(CUtensorMap* map, const float* ptr, int rows, int K, int ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("tcgen05 gemm: cuTensorMapEncodeTiled is unavailable"); return IRONB_ENOSUP; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || (ld & 3) || K <= 0 || rows <= 0) {
    set_error("tcgen05 gemm: operand must be 16-byte aligned with a row pitch that is a multiple of 4 floats");
    return IRONB_EINVAL;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, 128u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tcgen05 gemm: cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld); return IRONB_EINVAL; }
  return IRONB_OK;
}

// rows x K fp16 matrix with row pitch ld halfs -> 2-D map with a (64 x box_rows) SWIZZLE_128B box (mlp_h16.cu operands)
int make_map_h(CUtensorMap* map, const void* ptr, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("tcgen05 mlp: cuTensorMapEncodeTiled is unavailable"); return IRONB_ENOSUP; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || (ld & 7) || K <= 0 || rows <= 0) {
    set_error("tcgen05 mlp: fp16 operand must be 16-byte aligned with a row pitch that is a multiple of 8 halfs");
    return IRONB_EINVAL;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2u};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tcgen05 mlp: cuTensorMapEncodeTiled(fp16) failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld); return IRONB_EINVAL; }
  return IRONB_OK;
}

static std::atomic<int> g_mode{-1};
static int mode_now() {
  int m = g_mode.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = getenv("IRONB_GEMM");
    // IRONB_GEMM=simt: FFMA GEMMs; =tf32: 3xTF32 everywhere; default: fp16x2 forward-type products + 3xTF32 gradients
    m = (e && (e[0] == 's' || e[0] == 'S' || e[0] == '0')) ? 0 : (e && (e[0] == 't' || e[0] == 'T' || e[0] == '1')) ? 1 : 2;
    g_mode.store(m, std::memory_order_relaxed);
  }
  return m;
}
bool tc_enabled() { return mode_now() >= 1; }
bool pdl_enabled() {
#ifdef IRONB_ENABLE_PDL
  // measured (bench.py, graph replay with three parallel streams): 5.12 ms/step with PDL, 4.99 without -- the early CTAs of
  // a dependent GEMM hold whole SMs (192 KiB of shared memory each) while they wait, which costs the other streams more
  // than the hidden prologue gains; isolated eager chains do gain (wgrad 40 -> 36 us).  Needs the build flag (the
  // griddepcontrol instructions are compiled out otherwise, and the launch attribute without them would be a race) AND
  // IRONB_PDL=1.
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("IRONB_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
#else
  return false;
#endif
}
static std::atomic<int> g_write_hi{-1};
bool split_writes_hi() {
  int m = g_write_hi.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = getenv("IRONB_SPLIT_WRITE_HI");
    m = (e && e[0] == '1') ? 1 : 0;   // default 0: kind::tf32 ignores the 13 low mantissa bits (tests/test_gemm_gpu.py mode 2)
    g_write_hi.store(m, std::memory_order_relaxed);
  }
  return m == 1;
}
bool split_strict() {
  static const bool v = [] { const char* e = getenv("IRONB_SPLIT_STRICT"); return !(e && e[0] == '0'); }();
  return v;
}
int set_mode(int mode) {
  int prev = mode_now();
  g_mode.store(mode <= 0 ? 0 : (mode == 1 ? 1 : 2), std::memory_order_relaxed);
  return prev;
}
int mode() { return mode_now(); }

}  // namespace tc

int gemm_mode() { return tc::mode(); }

namespace {
struct EpiStore {
  float* C;
  int ldc;
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[4]) const {
    *reinterpret_cast<float4*>(C + (int64_t)m * ldc + n0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
};
// y = act(x W^T + b): a plain nn.Linear (+ ReLU) -- the background NeRF's layers (models/fields.py:281-322)
struct EpiBiasAct {
  const float* bias;     // [N] or null
  float* C;
  int ldc, relu;
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[4]) const {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[j] + (bias ? __ldg(bias + n0 + j) : 0.f);
      if (relu) v[j] = fmaxf(v[j], 0.f);
    }
    *reinterpret_cast<float4*>(C + (int64_t)m * ldc + n0) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
// dx = (dy masked by the ReLU of the SAME layer's output y) W: the mask is applied to the A operand beforehand (relu_mask_kernel)
__global__ void __launch_bounds__(256) relu_mask_kernel(const float* __restrict__ dy, const float* __restrict__ y, int64_t n,
                                                        float* __restrict__ out) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 g = *reinterpret_cast<const float4*>(dy + i), a = *reinterpret_cast<const float4*>(y + i);
    *reinterpret_cast<float4*>(out + i) = make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f,
                                                      a.w > 0.f ? g.w : 0.f);
  } else {
    for (int64_t k = i; k < n; ++k) out[k] = y[k] > 0.f ? dy[k] : 0.f;
  }
}
}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int ironb_set_gemm_mode(int mode) { return tc::set_mode(mode); }

extern "C" int ironb_gemm_nt(const float* A, int lda, const float* B, int ldb, int M, int N, int K, float* C, int ldc,
                             int mode, void* stream) {
  IRONB_REQUIRE(A && B && C, "gemm_nt: null pointer");
  IRONB_REQUIRE((N & 3) == 0 && (K & 3) == 0 && (ldc & 3) == 0, "gemm_nt: N, K, ldc must be multiples of 4");
  EpiStore ep{C, ldc};
  if (mode == 1) return tc::launch_gemm_nt_tc(A, lda, B, ldb, M, N, K, ep, as_stream(stream), "gemm_nt (tcgen05)", 1);
  if (mode == 2) return tc::launch_gemm_nt_tc(A, lda, B, ldb, M, N, K, ep, as_stream(stream), "gemm_nt (tcgen05, raw hi)", 0);
  return launch_gemm_nt(A, lda, B, ldb, M, N, K, ep, as_stream(stream), "gemm_nt (simt)");
}

// The fp16x2 pre-split GEMM (gemm_h16.cuh): operands as hi / lo half arrays (hi = fp16(x), lo = fp16((x - hi) * 2^11)), K-major,
// pitches in halfs (multiples of 8).  Unit-test entry.
extern "C" int ironb_gemm_nt_h16(const void* Ah, const void* Al, int lda, const void* Bh, const void* Bl, int ldb, int M, int N,
                                 int K, float* C, int ldc, void* stream) {
  IRONB_REQUIRE(Ah && Al && Bh && Bl && C, "gemm_nt_h16: null pointer");
  IRONB_REQUIRE((N & 3) == 0 && (lda & 7) == 0 && (ldb & 7) == 0 && (ldc & 3) == 0, "gemm_nt_h16: N % 4, lda % 8, ldb % 8, ldc % 4 must be 0");
  EpiStore ep{C, ldc};
  return h16::launch_gemm_h16(reinterpret_cast<const __half*>(Ah), reinterpret_cast<const __half*>(Al), lda,
                              reinterpret_cast<const __half*>(Bh), reinterpret_cast<const __half*>(Bl), ldb, M, N, K, ep,
                              as_stream(stream), "gemm_nt_h16");
}

// C[Nd][ldc] += A[M][lda]^T * B[M][ldb]  (weight-gradient shape): unit-test entry; mode 1 = tcgen05 (transpose + split-K
// K-major GEMM, needs ironb_gemm_tn_scratch_bytes of scratch), 0 = FFMA tiles
extern "C" int64_t ironb_gemm_tn_scratch_bytes(int M, int Nd, int Kd) {
  return wgrad_scratch_floats(M, Nd > Kd ? Nd : Kd) * (int64_t)sizeof(float);
}
extern "C" int ironb_gemm_tn(const float* A, int lda, const float* B, int ldb, int M, int Nd, int Kd, float* C, int ldc,
                             int mode, void* scratch, void* stream) {
  IRONB_REQUIRE(A && B && C, "gemm_tn: null pointer");
  IRONB_REQUIRE((Nd & 3) == 0 && (Kd & 3) == 0 && (ldc & 3) == 0, "gemm_tn: Nd, Kd, ldc must be multiples of 4");
  if (mode == 1) {
    IRONB_REQUIRE(scratch != nullptr, "gemm_tn: tcgen05 mode needs scratch");
    return launch_wgrad_tc(A, lda, B, ldb, M, Nd, Kd, C, ldc, reinterpret_cast<float*>(scratch), as_stream(stream), "gemm_tn (tcgen05)");
  }
  return launch_gemm_tn(A, lda, B, ldb, M, Nd, Kd, C, ldc, as_stream(stream), "gemm_tn (simt)");
}

// ---- plain nn.Linear layers on the tensor cores (the stage-1 background NeRF: models/fields.py:241-322) -------------------
// y[M][ldc] = act(x[M][lda] W[N][ldb]^T + b): N, K and the pitches multiples of 4.
extern "C" int ironb_linear_fwd(const float* x, int lda, const float* W, int ldb, const float* bias, int M, int N, int K, int relu,
                                float* y, int ldc, void* stream) {
  IRONB_REQUIRE(x && W && y, "linear_fwd: null pointer");
  IRONB_REQUIRE((N & 3) == 0 && (K & 3) == 0 && (lda & 3) == 0 && (ldb & 3) == 0 && (ldc & 3) == 0,
                "linear_fwd: N, K and the row pitches must be multiples of 4");
  EpiBiasAct ep{bias, y, ldc, relu};
  if (tc::tc_enabled()) return tc::launch_gemm_nt_tc(x, lda, W, ldb, M, N, K, ep, as_stream(stream), "linear_fwd (tcgen05)", 0);
  return launch_gemm_nt(x, lda, W, ldb, M, N, K, ep, as_stream(stream), "linear_fwd (simt)");
}
// out = dy where y > 0 else 0 (n elements): the ReLU backward of a layer, applied before its dgrad / wgrad products.
extern "C" int ironb_relu_mask(const float* dy, const float* y, int64_t n, float* out, void* stream) {
  IRONB_REQUIRE(dy && y && out && n >= 0, "relu_mask: bad arguments");
  IRONB_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
                "relu_mask: 16-byte aligned buffers");
  if (n == 0) return IRONB_OK;
  relu_mask_kernel<<<(unsigned)ceil_div64(ceil_div64(n, 4), 256), 256, 0, as_stream(stream)>>>(dy, y, n, out);
  IRONB_CHECK_LAUNCH("relu_mask_kernel");
  return IRONB_OK;
}
// dW[N][ldc] += dy[M][lda]^T x[M][ldb], db[N] += column sums of dy (db may be NULL): zero both first.  scratch:
// ironb_gemm_tn_scratch_bytes(M, N, K).
extern "C" int ironb_linear_wgrad(const float* dy, int lda, const float* x, int ldb, int M, int N, int K, float* dW, int ldc,
                                  float* db, void* scratch, void* stream) {
  IRONB_REQUIRE(dy && x && dW, "linear_wgrad: null pointer");
  IRONB_REQUIRE((N & 3) == 0 && (K & 3) == 0 && (ldc & 3) == 0, "linear_wgrad: N, K, ldc must be multiples of 4");
  return launch_wgrad_auto(dy, lda, x, ldb, M, N, K, dW, ldc, reinterpret_cast<float*>(scratch), as_stream(stream),
                           "linear_wgrad", db, db ? N : 0, 1.f);
}

namespace ironb {
IRONB_DEFINE_HANG_SETTER(hang_set_gemm)
int hang_set_sdf(unsigned long long*);
int hang_set_matnet(unsigned long long*);
int hang_set_mlp(unsigned long long*);
}  // namespace ironb

// Debug builds (-DIRONB_DEBUG_HANG): installs a MAPPED PINNED host buffer (>= 4 KiB, zeroed) that bounded mbarrier waits
// report into before they trap (gemm_tc.cuh).  Returns 1 if the library was built with the hooks, 0 if not.
extern "C" int ironb_debug_hang_buffer(void* mapped_host) {
#ifdef IRONB_DEBUG_HANG
  unsigned long long* p = reinterpret_cast<unsigned long long*>(mapped_host);
  int rc = ironb::hang_set_gemm(p) | ironb::hang_set_sdf(p) | ironb::hang_set_matnet(p) | ironb::hang_set_mlp(p);
  return rc == 0 ? 1 : -1;
#else
  (void)mapped_host;
  return 0;
#endif
}
