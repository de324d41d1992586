// Patch losses of the stage-2 loop: PyramidL2Loss and the masked SSIM loss, value AND gradient w.r.t. the
// rendered patch in one pass each (both are terminal losses: their autograd graphs end in a scalar, so the gradient is
// produced together with the value and the module's backward only scales it by the upstream scalar).
//
//   PyramidL2Loss.forward   models/image_losses.py:29-48    d_0 = pred - trgt;  d_{k+1} = avgpool2(conv7x7_gauss(d_k, pad 3));
//                                                           loss = sum_k sum(d_k^2) / ((h / 2^k)(w / 2^k)),  k = 0..4
//   ssim_loss_fn            models/image_losses.py:97-158   separable 11-tap Gaussian statistics without padding, channel mean,
//                                                           map padded back with 1.0, mean over the 11x11-ERODED hit mask
//
// The patches are small (64^2 .. 256^2 x 3): everything here is launch- and latency-bound, so the work is arranged as few
// launches (pyramid: 4 down + 4 up; SSIM: erode/count, statistics + derivative maps, transposed filter), all stream-ordered
// with no host read-back, i.e. capturable into the step's CUDA graph.  Inputs are addressed through element strides, so the
// [1, 3, H, W] view of the renderer's [H, W, 3] buffer (render_surface.py:594-596) is read in place.
#include "common.cuh"

namespace ironb {
namespace {

constexpr int MAXC = 4;
constexpr int PT = 16;          // pyramid: output tile edge
constexpr int ST = 16;          // SSIM: output tile edge
constexpr int WIN = 11;

struct Img {                    // a [C][H][W] image addressed through element strides
  const float* p;
  long long sc, sy, sx;
  __device__ __forceinline__ float at(int c, int y, int x) const { return p[c * sc + y * sy + x * sx]; }
};
struct ImgOut {
  float* p;
  long long sc, sy, sx;
};

struct Gauss7 { float w[7][7]; };

__device__ __forceinline__ float block_sum(float v, float* red) {   // blockDim.x == 256
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 8) t = red[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;   // valid in thread 0
}

// ---------------------------------------------------------------------------------------------- pyramid, downward
// out[c][y][x] = mean over the 2x2 block of conv7(in)(2y + dy, 2x + dx), zero padding 3.  FIRST: `in` is pred - trgt taken
// on the fly, and the level-0 sum of squares of the tile's own 32x32 input region is accumulated too (the last tile row /
// column also owns the odd leftover row / column the pooling drops).  loss += c_in * sum(in^2) [FIRST] + c_out * sum(out^2).
template <bool FIRST>
__global__ void __launch_bounds__(256) pyr_down_kernel(Img a, Img b, const float* __restrict__ in, int C, int h, int w,
                                                       float* __restrict__ out, int ho, int wo, Gauss7 G, float c_in,
                                                       float c_out, float* __restrict__ loss) {
  __shared__ float tile[2 * PT + 6][2 * PT + 7];
  __shared__ float red[8];
  const int c = blockIdx.z;
  const int ty0 = blockIdx.y * PT, tx0 = blockIdx.x * PT;           // output tile origin
  const int iy0 = 2 * ty0 - 3, ix0 = 2 * tx0 - 3;                   // input tile origin (with halo)
  float acc_in = 0.f;
  const bool last_y = (ty0 + PT >= ho), last_x = (tx0 + PT >= wo);
  for (int i = threadIdx.x; i < (2 * PT + 6) * (2 * PT + 6); i += 256) {
    const int r = i / (2 * PT + 6), q = i - r * (2 * PT + 6);
    const int y = iy0 + r, x = ix0 + q;
    float v = 0.f;
    if (y >= 0 && y < h && x >= 0 && x < w) {
      v = FIRST ? (a.at(c, y, x) - b.at(c, y, x)) : in[((size_t)c * h + y) * w + x];
      if (FIRST) {      // ownership of input pixels for the level-0 term: the tile's 2PT x 2PT core, extended at the borders
        const bool own_y = (r >= 3 && r < 3 + 2 * PT) || (last_y && r >= 3 + 2 * PT);
        const bool own_x = (q >= 3 && q < 3 + 2 * PT) || (last_x && q >= 3 + 2 * PT);
        if (own_y && own_x) acc_in += v * v;
      }
    }
    tile[r][q] = v;
  }
  __syncthreads();
  const int ly = threadIdx.x / PT, lx = threadIdx.x % PT;
  const int oy = ty0 + ly, ox = tx0 + lx;
  float acc_out = 0.f;
  if (oy < ho && ox < wo) {
    float s = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float t = 0.f;
#pragma unroll
        for (int u = 0; u < 7; ++u)
#pragma unroll
          for (int v = 0; v < 7; ++v) t = fmaf(G.w[u][v], tile[2 * ly + dy + u][2 * lx + dx + v], t);
        s += t;
      }
    s *= 0.25f;
    out[((size_t)c * ho + oy) * wo + ox] = s;
    acc_out = s * s;
  }
  const float tot = block_sum(acc_out * c_out + (FIRST ? acc_in * c_in : 0.f), red);
  if (threadIdx.x == 0 && tot != 0.f) atomicAdd(loss, tot);
}

// ---------------------------------------------------------------------------------------------- pyramid, upward
// g_k = 2 c_k d_k + conv7( up(g_{k+1}) / 4 ),  up(y', x') = g_{k+1}[y'/2][x'/2] inside [0, 2 ho) x [0, 2 wo), else 0.
// TOP: the coarse input is d_4 itself and g_4 = 2 c_4 d_4 is formed on the fly.  LAST: d_0 = pred - trgt on the fly, the
// result goes to the gradient image through its strides.
template <bool LAST>
__global__ void __launch_bounds__(256) pyr_up_kernel(Img a, Img b, const float* __restrict__ dk, const float* __restrict__ gnext,
                                                     int C, int h, int w, int ho, int wo, Gauss7 G, float two_ck,
                                                     float top_scale, float* __restrict__ gk, ImgOut go) {
  __shared__ float tile[PT + 6][PT + 7];
  const int c = blockIdx.z;
  const int y0 = blockIdx.y * PT, x0 = blockIdx.x * PT;
  for (int i = threadIdx.x; i < (PT + 6) * (PT + 6); i += 256) {
    const int r = i / (PT + 6), q = i - r * (PT + 6);
    const int y = y0 - 3 + r, x = x0 - 3 + q;
    float v = 0.f;
    if (y >= 0 && y < 2 * ho && x >= 0 && x < 2 * wo) v = gnext[((size_t)c * ho + (y >> 1)) * wo + (x >> 1)] * top_scale;
    tile[r][q] = v;
  }
  __syncthreads();
  const int ly = threadIdx.x / PT, lx = threadIdx.x % PT;
  const int y = y0 + ly, x = x0 + lx;
  if (y >= h || x >= w) return;
  float t = 0.f;
#pragma unroll
  for (int u = 0; u < 7; ++u)
#pragma unroll
    for (int v = 0; v < 7; ++v) t = fmaf(G.w[u][v], tile[ly + u][lx + v], t);
  const float d = LAST ? (a.at(c, y, x) - b.at(c, y, x)) : dk[((size_t)c * h + y) * w + x];
  const float g = fmaf(two_ck, d, t);
  if (LAST) go.p[c * go.sc + y * go.sy + x * go.sx] = g;
  else gk[((size_t)c * h + y) * w + x] = g;
}

// ---------------------------------------------------------------------------------------------- SSIM
struct Win { float g[WIN]; };

// eroded[y][x] = all mask pixels of the 11x11 window clipped to the image are set (kornia's geodesic border).  Also the
// selection count and the contribution of the border ring, whose padded ssim value is exactly 1.0:
//   acc[1] = sum eroded,   acc[0] += sum over the ring of eroded * 1.0
__global__ void __launch_bounds__(256) ssim_erode_kernel(const uint8_t* __restrict__ mask, int h, int w,
                                                         uint8_t* __restrict__ eroded, float* __restrict__ acc) {
  __shared__ float red[8];
  const int i = blockIdx.x * 256 + threadIdx.x;
  float cnt = 0.f, ring = 0.f;
  if (i < h * w) {
    const int y = i / w, x = i - y * w;
    bool e = true;
    if (mask != nullptr) {
      const int ya = max(y - WIN / 2, 0), yb = min(y + WIN / 2, h - 1), xa = max(x - WIN / 2, 0), xb = min(x + WIN / 2, w - 1);
      for (int yy = ya; yy <= yb && e; ++yy)
        for (int xx = xa; xx <= xb; ++xx)
          if (!mask[(size_t)yy * w + xx]) { e = false; break; }
    }
    const bool border = (y < WIN / 2 || y >= h - WIN / 2 || x < WIN / 2 || x >= w - WIN / 2);
    if (mask == nullptr && border) e = false;        // no mask: the mean runs over the unpadded map only
    eroded[i] = e ? 1 : 0;
    cnt = e ? 1.f : 0.f;
    ring = (e && border) ? 1.f : 0.f;
  }
  const float tc = block_sum(cnt, red);
  const float tr = block_sum(ring, red);
  if (threadIdx.x == 0) {
    if (tc != 0.f) atomicAdd(acc + 1, tc);
    if (tr != 0.f) atomicAdd(acc + 0, tr);
  }
}

// One 16x16 tile of the valid ssim map per CTA, channels in sequence.  Filtering order as the reference (H first, then W).
// acc[0] += sum_p w_p * mean_c S;  dmap[q][c][p] = w_p * dS/d{mu1, e11, e12} for the transposed filter of the backward.
__global__ void __launch_bounds__(256) ssim_fwd_kernel(Img X, Img Y, int C, int h, int w, Win Wn, float C1, float C2,
                                                       const uint8_t* __restrict__ eroded, float* __restrict__ acc,
                                                       float* __restrict__ dmap) {
  __shared__ float sx[ST + WIN - 1][ST + WIN];
  __shared__ float sy[ST + WIN - 1][ST + WIN];
  __shared__ float v[5][ST][ST + WIN];                     // vertically filtered: [quantity][out row][in col]
  __shared__ float red[8];
  const int hv = h - WIN + 1, wv = w - WIN + 1;
  const int y0 = blockIdx.y * ST, x0 = blockIdx.x * ST;
  const int ly = threadIdx.x / ST, lx = threadIdx.x % ST;
  const int py = y0 + ly, px = x0 + lx;
  const bool valid = py < hv && px < wv;
  const float wsel = valid ? (float)eroded[(size_t)(py + WIN / 2) * w + (px + WIN / 2)] : 0.f;
  float ssum = 0.f;
  for (int c = 0; c < C; ++c) {
    __syncthreads();
    for (int i = threadIdx.x; i < (ST + WIN - 1) * (ST + WIN - 1); i += 256) {
      const int r = i / (ST + WIN - 1), q = i - r * (ST + WIN - 1);
      const int y = y0 + r, x = x0 + q;
      const bool in = y < h && x < w;
      sx[r][q] = in ? X.at(c, y, x) : 0.f;
      sy[r][q] = in ? Y.at(c, y, x) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ST * (ST + WIN - 1); i += 256) {
      const int r = i / (ST + WIN - 1), q = i - r * (ST + WIN - 1);
      float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
      for (int k = 0; k < WIN; ++k) {
        const float a = sx[r + k][q], b = sy[r + k][q], g = Wn.g[k];
        m1 = fmaf(g, a, m1); m2 = fmaf(g, b, m2);
        e11 = fmaf(g, a * a, e11); e22 = fmaf(g, b * b, e22); e12 = fmaf(g, a * b, e12);
      }
      v[0][r][q] = m1; v[1][r][q] = m2; v[2][r][q] = e11; v[3][r][q] = e22; v[4][r][q] = e12;
    }
    __syncthreads();
    if (valid) {
      float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
      for (int k = 0; k < WIN; ++k) {
        const float g = Wn.g[k];
        m1 = fmaf(g, v[0][ly][lx + k], m1); m2 = fmaf(g, v[1][ly][lx + k], m2);
        e11 = fmaf(g, v[2][ly][lx + k], e11); e22 = fmaf(g, v[3][ly][lx + k], e22); e12 = fmaf(g, v[4][ly][lx + k], e12);
      }
      const float s1 = e11 - m1 * m1, s2 = e22 - m2 * m2, s12 = e12 - m1 * m2;
      const float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * s12 + C2, B1 = m1 * m1 + m2 * m2 + C1, B2 = s1 + s2 + C2;
      const float iB1 = 1.f / B1, iB2 = 1.f / B2;
      const float S = (A1 * iB1) * (A2 * iB2);
      ssum += S;
      if (dmap != nullptr) {
        // S = (A1/B1)(A2/B2) with s1 = e11 - mu1^2, s12 = e12 - mu1 mu2:
        //   dS/d e11 = -S / B2;   dS/d e12 = 2 A1 / (B1 B2);
        //   dS/d mu1 = 2 mu2 A2/(B1 B2) - 2 mu1 S/B1 + 2 mu1 S/B2 - 2 mu2 A1/(B1 B2)
        const float de11 = -S * iB2;
        const float de12 = 2.f * A1 * iB1 * iB2;
        const float dm1 = 2.f * m2 * A2 * iB1 * iB2 - 2.f * m1 * S * iB1 - 2.f * m1 * de11 - m2 * de12;
        const size_t plane = (size_t)hv * wv, o = (size_t)py * wv + px;
        dmap[(0 * C + c) * plane + o] = wsel * dm1;
        dmap[(1 * C + c) * plane + o] = wsel * de11;
        dmap[(2 * C + c) * plane + o] = wsel * de12;
      }
    }
  }
  const float tot = block_sum(wsel * ssum / (float)C, red);
  if (threadIdx.x == 0 && tot != 0.f) atomicAdd(acc, tot);
}

// dL/dX_c[q] = -(1 / (C cnt)) * sum_p G(q - p) (a_p + 2 X[q] b_p + Y[q] c_p): the transposed (full) separable filter of the
// three derivative maps.  Block (0, 0, 0) also finalises the loss value: loss = 1 - acc[0] / acc[1].
__global__ void __launch_bounds__(256) ssim_bwd_kernel(Img X, Img Y, int C, int h, int w, Win Wn, const float* __restrict__ dmap,
                                                       const float* __restrict__ acc, float* __restrict__ loss, ImgOut go) {
  __shared__ float t[3][ST + WIN - 1][ST + WIN];            // derivative maps around the tile (zero outside the valid map)
  __shared__ float u[3][ST + WIN - 1][ST + 1];              // horizontally filtered
  const int hv = h - WIN + 1, wv = w - WIN + 1;
  const int c = blockIdx.z;
  const int y0 = blockIdx.y * ST, x0 = blockIdx.x * ST;     // input-pixel tile
  const float cnt = acc[1];
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) *loss = 1.f - acc[0] / cnt;
  if (go.p == nullptr) return;
  const size_t plane = (size_t)hv * wv;
  for (int i = threadIdx.x; i < 3 * (ST + WIN - 1) * (ST + WIN - 1); i += 256) {
    const int m = i / ((ST + WIN - 1) * (ST + WIN - 1)), rem = i - m * (ST + WIN - 1) * (ST + WIN - 1);
    const int r = rem / (ST + WIN - 1), q = rem - r * (ST + WIN - 1);
    const int py = y0 - (WIN - 1) + r, px = x0 - (WIN - 1) + q;     // map position
    t[m][r][q] = (py >= 0 && py < hv && px >= 0 && px < wv) ? dmap[(m * C + c) * plane + (size_t)py * wv + px] : 0.f;
  }
  __syncthreads();
  // horizontal: u[m][r][lx] = sum_k g[k] t[m][r][lx + (WIN-1) - k]   (input x = x0 + lx receives map px = x - k)
  for (int i = threadIdx.x; i < 3 * (ST + WIN - 1) * ST; i += 256) {
    const int m = i / ((ST + WIN - 1) * ST), rem = i - m * (ST + WIN - 1) * ST;
    const int r = rem / ST, lx = rem - r * ST;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < WIN; ++k) s = fmaf(Wn.g[k], t[m][r][lx + (WIN - 1) - k], s);
    u[m][r][lx] = s;
  }
  __syncthreads();
  const int ly = threadIdx.x / ST, lx = threadIdx.x % ST;
  const int y = y0 + ly, x = x0 + lx;
  if (y >= h || x >= w) return;
  float fa = 0.f, fb = 0.f, fc = 0.f;
#pragma unroll
  for (int k = 0; k < WIN; ++k) {
    const float g = Wn.g[k];
    fa = fmaf(g, u[0][ly + (WIN - 1) - k][lx], fa);
    fb = fmaf(g, u[1][ly + (WIN - 1) - k][lx], fb);
    fc = fmaf(g, u[2][ly + (WIN - 1) - k][lx], fc);
  }
  const float scale = cnt > 0.f ? -1.f / ((float)C * cnt) : 0.f;
  go.p[c * go.sc + y * go.sy + x * go.sx] = scale * (fa + 2.f * X.at(c, y, x) * fb + Y.at(c, y, x) * fc);
}

Gauss7 make_gauss7() {
  // scipy.ndimage.gaussian_filter(dirac 7x7, sigma = 1): the 1-D sigma-1 Gaussian truncated at radius 4, normalised, applied
  // to a 7-wide dirac with scipy's 'reflect' border, per axis (models/image_losses.py:17-20)
  const int r = 4;
  double wgt[9], sum = 0.0;
  for (int d = -r; d <= r; ++d) { wgt[d + r] = exp(-0.5 * d * d); sum += wgt[d + r]; }
  for (int i = 0; i < 9; ++i) wgt[i] /= sum;
  double g1[7];
  for (int j = 0; j < 7; ++j) {
    g1[j] = 0.0;
    for (int d = -r; d <= r; ++d) {
      int i = j + d;
      while (i < 0 || i > 6) i = (i < 0) ? -i - 1 : 13 - i;
      if (i == 3) g1[j] += wgt[d + r];
    }
  }
  Gauss7 G;
  for (int a = 0; a < 7; ++a)
    for (int b = 0; b < 7; ++b) G.w[a][b] = (float)(g1[a] * g1[b]);
  return G;
}

struct Levels { int h[6], w[6], n; int64_t off_d[6], off_g[6], floats; };
Levels pyr_levels(int C, int H, int W) {
  Levels L;
  memset(&L, 0, sizeof(L));
  L.h[0] = H; L.w[0] = W; L.n = 1;
  int64_t off = 0;
  for (int k = 1; k <= 4; ++k) {
    L.h[k] = L.h[k - 1] / 2; L.w[k] = L.w[k - 1] / 2;
    if (L.h[k] < 1 || L.w[k] < 1) break;
    L.off_d[k] = off; off += ((int64_t)C * L.h[k] * L.w[k] + 63) / 64 * 64;
    L.off_g[k] = off; off += ((int64_t)C * L.h[k] * L.w[k] + 63) / 64 * 64;
    L.n = k + 1;
  }
  L.floats = off;
  return L;
}

}  // namespace
}  // namespace ironb

using namespace ironb;

extern "C" int64_t ironb_patch_loss_workspace_bytes(int C, int H, int W) {
  if (C < 1 || C > MAXC || H < 1 || W < 1) return -1;
  const int64_t pyr = pyr_levels(C, H, W).floats * 4;
  const int hv = H - WIN + 1 > 0 ? H - WIN + 1 : 0, wv = W - WIN + 1 > 0 ? W - WIN + 1 : 0;
  const int64_t ssim = 256 /*acc*/ + ((int64_t)H * W + 255) / 256 * 256 /*eroded*/ + (int64_t)3 * C * hv * wv * 4;
  return (pyr > ssim ? pyr : ssim) + 256;
}

// PyramidL2Loss.forward (models/image_losses.py:29-48) for one [C][H][W] image pair.  *loss is ACCUMULATED (caller zeroes
// it); grad (may be NULL) receives d loss / d pred through its strides.
extern "C" int ironb_pyramid_l2(const float* pred, const int64_t* pstr, const float* trgt, const int64_t* tstr, int C, int H,
                                int W, float* loss, float* grad, const int64_t* gstr, void* ws, int64_t ws_bytes,
                                void* stream) {
  IRONB_REQUIRE(pred && trgt && loss && pstr && tstr, "pyramid_l2: null pointer");
  IRONB_REQUIRE(C >= 1 && C <= MAXC && H >= 1 && W >= 1, "pyramid_l2: bad shape");
  IRONB_REQUIRE(grad == nullptr || gstr != nullptr, "pyramid_l2: gradient strides missing");
  cudaStream_t st = as_stream(stream);
  Img a{pred, pstr[0], pstr[1], pstr[2]}, b{trgt, tstr[0], tstr[1], tstr[2]};
  ImgOut go{grad, grad ? gstr[0] : 0, grad ? gstr[1] : 0, grad ? gstr[2] : 0};
  const Levels L = pyr_levels(C, H, W);
  IRONB_REQUIRE(L.n == 5, "pyramid_l2: the patch needs at least 16 x 16 pixels (four 2x poolings, as in the reference)");
  IRONB_REQUIRE(ws != nullptr && ws_bytes >= L.floats * 4, "pyramid_l2: workspace too small");
  float* base = reinterpret_cast<float*>(ws);
  static const Gauss7 G = make_gauss7();
  double ck[5];
  for (int k = 0; k < 5; ++k) ck[k] = 1.0 / (((double)H / (double)(1 << k)) * ((double)W / (double)(1 << k)));
  // downward: d_1 .. d_4 and the loss
  for (int k = 0; k < 4; ++k) {
    dim3 grid((unsigned)ceil_div64(L.w[k + 1], PT), (unsigned)ceil_div64(L.h[k + 1], PT), (unsigned)C);
    if (k == 0)
      pyr_down_kernel<true><<<grid, 256, 0, st>>>(a, b, nullptr, C, L.h[0], L.w[0], base + L.off_d[1], L.h[1], L.w[1], G,
                                                  (float)ck[0], (float)ck[1], loss);
    else
      pyr_down_kernel<false><<<grid, 256, 0, st>>>(a, b, base + L.off_d[k], C, L.h[k], L.w[k], base + L.off_d[k + 1],
                                                   L.h[k + 1], L.w[k + 1], G, 0.f, (float)ck[k + 1], loss);
    IRONB_CHECK_LAUNCH("pyr_down_kernel");
  }
  if (grad == nullptr) return IRONB_OK;
  // upward: g_3 .. g_0
  for (int k = 3; k >= 0; --k) {
    dim3 grid((unsigned)ceil_div64(L.w[k], PT), (unsigned)ceil_div64(L.h[k], PT), (unsigned)C);
    const float* gnext = (k == 3) ? base + L.off_d[4] : base + L.off_g[k + 1];
    const float top_scale = (k == 3) ? (float)(2.0 * ck[4]) * 0.25f : 0.25f;
    if (k == 0)
      pyr_up_kernel<true><<<grid, 256, 0, st>>>(a, b, nullptr, gnext, C, L.h[0], L.w[0], L.h[1], L.w[1], G, (float)(2.0 * ck[0]),
                                                top_scale, nullptr, go);
    else
      pyr_up_kernel<false><<<grid, 256, 0, st>>>(a, b, base + L.off_d[k], gnext, C, L.h[k], L.w[k], L.h[k + 1], L.w[k + 1], G,
                                                 (float)(2.0 * ck[k]), top_scale, base + L.off_g[k], go);
    IRONB_CHECK_LAUNCH("pyr_up_kernel");
  }
  return IRONB_OK;
}

// ssim_loss_fn (models/image_losses.py:97-158) for one [C][H][W] image pair; mask: [H][W] bytes or NULL.  *loss is WRITTEN.
extern "C" int ironb_ssim_loss(const float* X, const int64_t* xstr, const float* Y, const int64_t* ystr, const uint8_t* mask,
                               int C, int H, int W, float data_range, int win_size, float win_sigma, float K1, float K2,
                               float* loss, float* grad, const int64_t* gstr, void* ws, int64_t ws_bytes, void* stream) {
  IRONB_REQUIRE(X && Y && loss && xstr && ystr, "ssim_loss: null pointer");
  IRONB_REQUIRE(C >= 1 && C <= MAXC, "ssim_loss: bad channel count");
  IRONB_REQUIRE(win_size == WIN, "ssim_loss: only the reference's 11-tap window is built");
  IRONB_REQUIRE(H >= WIN && W >= WIN, "ssim_loss: the patch must be at least 11 x 11 (the reference skips the smoothing below that)");
  IRONB_REQUIRE(grad == nullptr || gstr != nullptr, "ssim_loss: gradient strides missing");
  IRONB_REQUIRE(ws != nullptr && ws_bytes >= ironb_patch_loss_workspace_bytes(C, H, W) - 256, "ssim_loss: workspace too small");
  cudaStream_t st = as_stream(stream);
  Img a{X, xstr[0], xstr[1], xstr[2]}, b{Y, ystr[0], ystr[1], ystr[2]};
  ImgOut go{grad, grad ? gstr[0] : 0, grad ? gstr[1] : 0, grad ? gstr[2] : 0};
  Win Wn;
  {   // _fspecial_gauss_1d (:51-66), fp32 arithmetic like torch
    float s = 0.f;
    for (int i = 0; i < WIN; ++i) { const float c = (float)(i - WIN / 2); Wn.g[i] = expf(-(c * c) / (2.f * win_sigma * win_sigma)); s += Wn.g[i]; }
    for (int i = 0; i < WIN; ++i) Wn.g[i] /= s;
  }
  const float C1 = (K1 * data_range) * (K1 * data_range), C2 = (K2 * data_range) * (K2 * data_range);
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  float* acc = reinterpret_cast<float*>(base);
  uint8_t* eroded = base + 256;
  float* dmap = reinterpret_cast<float*>(base + 256 + ((int64_t)H * W + 255) / 256 * 256);
  const int hv = H - WIN + 1, wv = W - WIN + 1;
  IRONB_CUDA(cudaMemsetAsync(acc, 0, 256, st));
  ssim_erode_kernel<<<(unsigned)ceil_div64((int64_t)H * W, 256), 256, 0, st>>>(mask, H, W, eroded, acc);
  IRONB_CHECK_LAUNCH("ssim_erode_kernel");
  dim3 gf((unsigned)ceil_div64(wv, ST), (unsigned)ceil_div64(hv, ST), 1);
  ssim_fwd_kernel<<<gf, 256, 0, st>>>(a, b, C, H, W, Wn, C1, C2, eroded, acc, grad ? dmap : nullptr);
  IRONB_CHECK_LAUNCH("ssim_fwd_kernel");
  dim3 gb((unsigned)ceil_div64(W, ST), (unsigned)ceil_div64(H, ST), (unsigned)C);
  if (grad == nullptr) gb = dim3(1, 1, 1);
  ssim_bwd_kernel<<<gb, 256, 0, st>>>(a, b, C, H, W, Wn, dmap, acc, loss, go);
  IRONB_CHECK_LAUNCH("ssim_bwd_kernel");
  return IRONB_OK;
}
