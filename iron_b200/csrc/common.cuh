// Shared helpers for the iron_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/iron_b200.h"

namespace ironb {

void set_error(const char* fmt, ...);
void note_launch();   // bumps the process-wide kernel-launch counter (ironb_launch_count)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define IRONB_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::ironb::set_error(__VA_ARGS__);      \
      return IRONB_EINVAL;                  \
    }                                       \
  } while (0)

#define IRONB_CHECK_LAUNCH(what)                                               \
  do {                                                                         \
    cudaError_t e__ = cudaGetLastError();                                      \
    ::ironb::note_launch();                                                    \
    if (e__ != cudaSuccess) {                                                  \
      ::ironb::set_error("%s: %s", what, cudaGetErrorString(e__));             \
      return (int)e__;                                                         \
    }                                                                          \
  } while (0)

#define IRONB_CUDA(call)                                                       \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ::ironb::set_error("%s: %s", #call, cudaGetErrorString(e__));            \
      return (int)e__;                                                         \
    }                                                                          \
  } while (0)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

int num_sms();

// tracer implementations (trace.cu: fused persistent FFMA; trace_batched.cu: batched tcgen05)
int trace_mode();   // 2 = batched tcgen05 fp16x2 split (default), 1 = batched tcgen05 3xTF32, 0 = fused FFMA
bool trace_mlp_fused_supported(const ironb_mlp_layout* lay);   // shapes the tcgen05 tracer handles (else: FFMA tracer)
int64_t trace_batched_workspace_bytes(const ironb_mlp_layout* lay, int64_t N);
int trace_batched(const ironb_mlp_layout* lay, const float* packed, const float* ray_o, const float* ray_d,
                  const float* min_dis, const float* max_dis, const uint8_t* work_mask, int64_t N, float thr, int iters,
                  int n_steps, const float* linspace, uint8_t* conv, float* points, float* sdf, float* dist,
                  int64_t* stats, void* ws, int64_t ws_bytes, cudaStream_t st);

// ---- the two scalar activations of the SDF net, written exactly as torch evaluates them ----
// nn.Softplus(beta): x*beta > 20 ? x : log1p(exp(x*beta))/beta        (models/fields.py:80)
__device__ __forceinline__ float softplus_beta(float z, float beta) {
  float bz = z * beta;
  return bz > 20.f ? z : __fdiv_rn(log1pf(expf(bz)), beta);
}
// d/dz softplus = sigmoid(beta z); 1 beyond the threshold (autograd of the thresholded branch)
__device__ __forceinline__ float softplus_d1(float z, float beta) {
  float bz = z * beta;
  return bz > 20.f ? 1.f : __fdiv_rn(1.f, 1.f + expf(-bz));
}

// The same two functions on the MUFU units, for the GEMM epilogues of the differentiable path (sdf.cu), where the libm
// versions above (expf + log1pf + IEEE division: ~100 instructions per element, 64 elements per epilogue thread) cost more
// than the whole tensor-core mainloop of a tile (profiles/r2e_launches.md):
//   softplus: max(z, 0) + log2(1 + 2^(-|beta z| log2 e)) * ln2 / beta     -- the log term is <= ln2 / beta, so the 2^-22
//             relative error of ex2 / lg2 is ~1.6e-9 absolute; beyond beta z > 20 it is exactly z (torch's threshold)
//   sigmoid : t = 2^(-|beta z| log2 e);  1 / (1 + t)  or  t / (1 + t)      -- 2^-22 relative; exactly 1 beyond the threshold
__device__ __forceinline__ float softplus_beta_fast(float z, float beta) {
  float t, L;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(z * beta) * -1.4426950408889634f));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(L) : "f"(1.f + t));
  return fmaf(L, __fdividef(0.6931471805599453f, beta), fmaxf(z, 0.f));
}
__device__ __forceinline__ float softplus_d1_fast(float z, float beta) {
  const float bz = z * beta;
  float t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(bz) * -1.4426950408889634f));
  const float r = __fdividef(1.f, 1.f + t);
  return bz >= 0.f ? r : t * r;
}

#define IRONB_SQRT2F 1.41421356237309515f  // float(np.sqrt(2)); the reference DIVIDES by it (fields.py:89)

}  // namespace ironb
