// Ray generation, unit-sphere clip and the hit-point compaction / gather / scatter glue.
//   Camera.get_rays      models/raytracer.py:254-286
//   intersect_sphere     models/raytracer.py:223-237
//   boolean-mask indexing and masked scatter of render_normal_and_color / render_fn
//                        models/raytracer.py:617-621, render_surface.py:140-146
// All HBM-bound elementwise kernels.  Products and sums are spelled with explicit round-to-nearest mul/add
// so no FMA contraction changes the reference's fp32 values.
#include "common.cuh"

namespace ironb {
namespace {

__device__ __forceinline__ float dot3(const float a[3], const float b[3]) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
}

__device__ __forceinline__ void sphere_clip(const float o[3], const float d[3], float r, uint8_t* hit, float* tmin,
                                            float* tmax) {
  float d1 = __fdiv_rn(-dot3(d, o), dot3(d, d));
  float p[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) p[c] = __fadd_rn(o[c], __fmul_rn(d1, d[c]));
  float tmp = __fsub_rn(__fmul_rn(r, r), dot3(p, p));
  *hit = tmp > 0.f ? 1 : 0;
  float d2 = __fdiv_rn(sqrtf(fmaxf(tmp, 0.f)), sqrtf(dot3(d, d)));
  *tmin = fmaxf(__fsub_rn(d1, d2), 0.f);
  *tmax = __fadd_rn(d1, d2);
}

struct Cam {
  float kinv[9], rot[9], org[3];
};

__global__ void __launch_bounds__(256) camera_rays_kernel(const float* __restrict__ uv, int64_t N,
                                                          const float* __restrict__ Kinv3, const float* __restrict__ R,
                                                          const float* __restrict__ origin, float r,
                                                          float* __restrict__ ray_o, float* __restrict__ ray_d,
                                                          float* __restrict__ ray_d_norm, uint8_t* __restrict__ hit,
                                                          float* __restrict__ min_dis, float* __restrict__ max_dis) {
  __shared__ Cam cam;
  if (threadIdx.x < 9) { cam.kinv[threadIdx.x] = Kinv3[threadIdx.x]; cam.rot[threadIdx.x] = R[threadIdx.x]; }
  if (threadIdx.x < 3) cam.org[threadIdx.x] = origin[threadIdx.x];
  __syncthreads();
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float u[3] = {uv[i * 2], uv[i * 2 + 1], 1.f};
  float t[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) t[c] = dot3(u, cam.kinv + c * 3);   // uv1 @ K_inv[:3,:3]^T
#pragma unroll
  for (int c = 0; c < 3; ++c) d[c] = dot3(t, cam.rot + c * 3);    // ... @ C2W[:3,:3]^T
  float nrm = sqrtf(dot3(d, d));
#pragma unroll
  for (int c = 0; c < 3; ++c) d[c] = __fdiv_rn(d[c], nrm);
  const float o[3] = {cam.org[0], cam.org[1], cam.org[2]};
#pragma unroll
  for (int c = 0; c < 3; ++c) { ray_o[i * 3 + c] = o[c]; ray_d[i * 3 + c] = d[c]; }
  ray_d_norm[i] = nrm;
  if (min_dis != nullptr) {
    uint8_t h; float a, b;
    sphere_clip(o, d, r, &h, &a, &b);
    hit[i] = h; min_dis[i] = a; max_dis[i] = b;
  }
}

__global__ void __launch_bounds__(256) intersect_kernel(const float* __restrict__ ray_o, const float* __restrict__ ray_d,
                                                        int64_t N, float r, uint8_t* __restrict__ hit,
                                                        float* __restrict__ min_dis, float* __restrict__ max_dis) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float o[3] = {ray_o[i * 3], ray_o[i * 3 + 1], ray_o[i * 3 + 2]};
  const float d[3] = {ray_d[i * 3], ray_d[i * 3 + 1], ray_d[i * 3 + 2]};
  uint8_t h; float a, b;
  sphere_clip(o, d, r, &h, &a, &b);
  hit[i] = h; min_dis[i] = a; max_dis[i] = b;
}

// ---- stable compaction: three small kernels (block counts -> serial scan of <= 4096 partials -> scatter)
constexpr int CB = 1024;   // elements per block
__global__ void __launch_bounds__(256) compact_count_kernel(const uint8_t* __restrict__ mask, int64_t N,
                                                            int32_t* __restrict__ block_counts) {
  int64_t base = (int64_t)blockIdx.x * CB;
  int c = 0;
  for (int i = threadIdx.x; i < CB; i += 256) {
    int64_t j = base + i;
    c += (j < N && mask[j]) ? 1 : 0;
  }
  __shared__ int red[256];
  red[threadIdx.x] = c;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_counts[blockIdx.x] = red[0];
}

__global__ void compact_scan_kernel(int32_t* __restrict__ block_counts, int nblocks, int32_t* __restrict__ count) {
  // exclusive scan by one warp, 32 values per step
  int lane = threadIdx.x;
  int running = 0;
  for (int base = 0; base < nblocks; base += 32) {
    int i = base + lane;
    int v = i < nblocks ? block_counts[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (i < nblocks) block_counts[i] = running + inc - v;
    running += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) *count = running;
}

__global__ void __launch_bounds__(256) compact_scatter_kernel(const uint8_t* __restrict__ mask, int64_t N,
                                                              const int32_t* __restrict__ block_offsets,
                                                              int32_t* __restrict__ idx) {
  // each warp owns a contiguous run of 128 elements: ballot + popc keeps ascending order
  __shared__ int warp_tot[8];
  int64_t base = (int64_t)blockIdx.x * CB;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int flags[4], pre[4], tot = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int64_t j = base + warp * 128 + q * 32 + lane;
    bool f = (j < N) && mask[j];
    unsigned m = __ballot_sync(0xffffffffu, f);
    flags[q] = f;
    pre[q] = tot + __popc(m & ((1u << lane) - 1u));
    tot += __popc(m);
  }
  if (lane == 0) warp_tot[warp] = tot;
  __syncthreads();
  int woff = block_offsets[blockIdx.x];
  for (int w = 0; w < warp; ++w) woff += warp_tot[w];
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (flags[q]) idx[woff + pre[q]] = (int32_t)(base + warp * 128 + q * 32 + lane);
}

// 3x3 flat grey-scale dilation (is_max) / erosion over the valid neighbours of each pixel: kornia.morphology with the
// 'geodesic' border (the image border never wins), which raytrace_camera's hole filling uses (models/raytracer.py:554-557)
__global__ void __launch_bounds__(256) morph3x3_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W,
                                                       int is_max) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  int y = i / W, x = i - y * W;
  float v = src[i];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      float u = src[yy * W + xx];
      v = is_max ? fmaxf(v, u) : fminf(v, u);
    }
  dst[i] = v;
}

// Sobel gradient magnitude of the depth map: kornia.filters.sobel (normalized 3x3 Sobel / 8 on a replicate-padded image,
// sqrt(gx^2 + gy^2 + 1e-6)), the depth-edge detector of raytrace_camera (models/raytracer.py:569)
__global__ void __launch_bounds__(256) sobel_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  int y = i / W, x = i - y * W;
  float v[3][3];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      int yy = min(max(y + dy, 0), H - 1), xx = min(max(x + dx, 0), W - 1);
      v[dy + 1][dx + 1] = src[yy * W + xx];
    }
  float gx = ((v[0][2] - v[0][0]) + 2.f * (v[1][2] - v[1][0]) + (v[2][2] - v[2][0])) * 0.125f;
  float gy = ((v[2][0] - v[0][0]) + 2.f * (v[2][1] - v[0][1]) + (v[2][2] - v[0][2])) * 0.125f;
  dst[i] = sqrtf(gx * gx + gy * gy + 1e-6f);
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                          int64_t M, int width, float* __restrict__ dst) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M * width) return;
  int64_t m = i / width;
  int c = (int)(i - m * width);
  dst[i] = src[(int64_t)idx[m] * width + c];
}
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                           int64_t M, int width, float* __restrict__ dst) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M * width) return;
  int64_t m = i / width;
  int c = (int)(i - m * width);
  dst[(int64_t)idx[m] * width + c] = src[i];
}

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int ironb_camera_rays(const float* uv, int64_t N, const float* Kinv3, const float* R_c2w,
                                 const float* origin, float r, float* ray_o, float* ray_d, float* ray_d_norm,
                                 uint8_t* hit, float* min_dis, float* max_dis, void* stream) {
  IRONB_REQUIRE(N >= 0, "camera_rays: N < 0");
  if (N == 0) return IRONB_OK;
  IRONB_REQUIRE(uv && Kinv3 && R_c2w && origin && ray_o && ray_d && ray_d_norm, "camera_rays: null pointer");
  IRONB_REQUIRE(min_dis == nullptr || (hit && max_dis), "camera_rays: hit/min_dis/max_dis must come together");
  camera_rays_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, as_stream(stream)>>>(uv, N, Kinv3, R_c2w, origin, r, ray_o,
                                                                                ray_d, ray_d_norm, hit, min_dis, max_dis);
  IRONB_CHECK_LAUNCH("camera_rays_kernel");
  return IRONB_OK;
}

extern "C" int ironb_intersect_sphere(const float* ray_o, const float* ray_d, int64_t N, float r, uint8_t* hit,
                                      float* min_dis, float* max_dis, void* stream) {
  IRONB_REQUIRE(N >= 0, "intersect_sphere: N < 0");
  if (N == 0) return IRONB_OK;
  IRONB_REQUIRE(ray_o && ray_d && hit && min_dis && max_dis, "intersect_sphere: null pointer");
  intersect_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, as_stream(stream)>>>(ray_o, ray_d, N, r, hit, min_dis, max_dis);
  IRONB_CHECK_LAUNCH("intersect_kernel");
  return IRONB_OK;
}

// idx must hold N int32 plus ceil(N/1024) scratch int32 after them (idx[N ..]).
extern "C" int ironb_compact_mask(const uint8_t* mask, int64_t N, int32_t* idx, int32_t* count, void* stream) {
  IRONB_REQUIRE(N >= 0 && N < (1ll << 31), "compact_mask: N out of range");
  IRONB_REQUIRE(idx && count, "compact_mask: null pointer");
  cudaStream_t st = as_stream(stream);
  if (N == 0) { IRONB_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st)); return IRONB_OK; }
  IRONB_REQUIRE(mask != nullptr, "compact_mask: null mask");
  int nblocks = (int)ceil_div64(N, CB);
  int32_t* scratch = idx + N;
  compact_count_kernel<<<nblocks, 256, 0, st>>>(mask, N, scratch);
  IRONB_CHECK_LAUNCH("compact_count_kernel");
  compact_scan_kernel<<<1, 32, 0, st>>>(scratch, nblocks, count);
  IRONB_CHECK_LAUNCH("compact_scan_kernel");
  compact_scatter_kernel<<<nblocks, 256, 0, st>>>(mask, N, scratch, idx);
  IRONB_CHECK_LAUNCH("compact_scatter_kernel");
  return IRONB_OK;
}

extern "C" int ironb_gather_rows(const float* src, const int32_t* idx, int64_t M, int width, float* dst, void* stream) {
  IRONB_REQUIRE(M >= 0 && width > 0, "gather_rows: bad size");
  if (M == 0) return IRONB_OK;
  IRONB_REQUIRE(src && idx && dst, "gather_rows: null pointer");
  gather_rows_kernel<<<(unsigned)ceil_div64(M * width, 256), 256, 0, as_stream(stream)>>>(src, idx, M, width, dst);
  IRONB_CHECK_LAUNCH("gather_rows_kernel");
  return IRONB_OK;
}

extern "C" int ironb_scatter_rows(const float* src, const int32_t* idx, int64_t M, int width, float* dst, void* stream) {
  IRONB_REQUIRE(M >= 0 && width > 0, "scatter_rows: bad size");
  if (M == 0) return IRONB_OK;
  IRONB_REQUIRE(src && idx && dst, "scatter_rows: null pointer");
  scatter_rows_kernel<<<(unsigned)ceil_div64(M * width, 256), 256, 0, as_stream(stream)>>>(src, idx, M, width, dst);
  IRONB_CHECK_LAUNCH("scatter_rows_kernel");
  return IRONB_OK;
}

// closing = erosion(dilation(depth)) with a 3x3 all-ones structuring element; tmp and out are H*W floats
extern "C" int ironb_depth_closing(const float* depth, int H, int W, float* tmp, float* out, void* stream) {
  IRONB_REQUIRE(H > 0 && W > 0 && (int64_t)H * W < (1ll << 31), "depth_closing: bad size");
  IRONB_REQUIRE(depth && tmp && out, "depth_closing: null pointer");
  unsigned nb = (unsigned)ceil_div64((int64_t)H * W, 256);
  morph3x3_kernel<<<nb, 256, 0, as_stream(stream)>>>(depth, tmp, H, W, 1);
  IRONB_CHECK_LAUNCH("morph3x3_kernel (dilate)");
  morph3x3_kernel<<<nb, 256, 0, as_stream(stream)>>>(tmp, out, H, W, 0);
  IRONB_CHECK_LAUNCH("morph3x3_kernel (erode)");
  return IRONB_OK;
}

extern "C" int ironb_sobel_depth(const float* depth, int H, int W, float* out, void* stream) {
  IRONB_REQUIRE(H > 0 && W > 0 && (int64_t)H * W < (1ll << 31), "sobel_depth: bad size");
  IRONB_REQUIRE(depth && out, "sobel_depth: null pointer");
  sobel_kernel<<<(unsigned)ceil_div64((int64_t)H * W, 256), 256, 0, as_stream(stream)>>>(depth, out, H, W);
  IRONB_CHECK_LAUNCH("sobel_kernel");
  return IRONB_OK;
}
