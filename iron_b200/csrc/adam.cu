// Multi-tensor Adam: every parameter tensor of every network of the stage-2 step in ONE launch (the reference steps six
// torch.optim.Adam instances, render_surface.py:112-113, 651-653, models/network_conf.py:707-716; torch's single-tensor
// path costs ~8 launches per tensor and step).  Arithmetic follows torch.optim.Adam (torch/optim/adam.py,
// _single_tensor_adam, amsgrad = False, maximize = False):
//     g   = grad + weight_decay * p
//     m   = m + (1 - beta1) (g - m)                         (lerp)
//     v   = beta2 v + (1 - beta2) g g
//     p   = p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// with the bias corrections evaluated in double like the Python scalars.  The step count t lives on the device (one int32
// per call site) so the launch can sit inside a CUDA graph: the kernel reads t + 1, a one-thread kernel advances it.
#include "common.cuh"

namespace ironb {
namespace {

struct AdamArgs {
  const ironb_adam_tensor* tensors;
  const int* step;
  double beta1, beta2;        // doubles, like the Python scalars of the reference: 1 - beta and beta^t are formed in double
  float eps;
};

__global__ void __launch_bounds__(256) adam_kernel(AdamArgs a) {
  const ironb_adam_tensor T = a.tensors[blockIdx.y];
  const int t = *a.step + 1;
  const double bc1 = 1.0 - pow(a.beta1, (double)t);
  const double bc2 = 1.0 - pow(a.beta2, (double)t);
  const float step_size = (float)((double)T.lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float w1 = (float)(1.0 - a.beta1), w2 = (float)(1.0 - a.beta2), b2 = (float)a.beta2;
  float* __restrict__ p = T.param;
  const float* __restrict__ g = T.grad;
  float* __restrict__ m = T.exp_avg;
  float* __restrict__ v = T.exp_avg_sq;
  if (g == nullptr) return;                    // parameter without a gradient this step: skipped, like torch
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T.numel; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (T.weight_decay != 0.f) gi = fmaf(T.weight_decay, pi, gi);
    const float mi = fmaf(w1, gi - m[i], m[i]);
    const float vi = fmaf(w2 * gi, gi, b2 * v[i]);
    m[i] = mi;
    v[i] = vi;
    const float denom = __fdiv_rn(sqrtf(vi), bc2_sqrt) + a.eps;
    p[i] = fmaf(-step_size, __fdiv_rn(mi, denom), pi);
  }
}

__global__ void adam_advance_kernel(int* step) { *step += 1; }

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int ironb_adam_step(const ironb_adam_tensor* tensors_dev, int n_tensors, int64_t max_numel, double beta1,
                               double beta2, double eps, int* step_dev, void* stream) {
  IRONB_REQUIRE(tensors_dev && step_dev, "adam_step: null pointer");
  IRONB_REQUIRE(n_tensors >= 0 && n_tensors <= 65535, "adam_step: at most 65535 tensors per call");
  if (n_tensors == 0) return IRONB_OK;
  cudaStream_t st = as_stream(stream);
  int64_t bx = ceil_div64(max_numel, 256 * 4);          // ~4 elements per thread for the largest tensor
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  AdamArgs a{tensors_dev, step_dev, beta1, beta2, (float)eps};
  adam_kernel<<<dim3((unsigned)bx, (unsigned)n_tensors), 256, 0, st>>>(a);
  IRONB_CHECK_LAUNCH("adam_kernel");
  adam_advance_kernel<<<1, 1, 0, st>>>(step_dev);
  IRONB_CHECK_LAUNCH("adam_advance_kernel");
  return IRONB_OK;
}
