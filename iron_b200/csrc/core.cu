// Error string, device query, and the packed-weight layouts (host code only).
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace ironb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static void finish_layout(ironb_mlp_layout* L) {
  int64_t off = 0;
  for (int l = 0; l < L->n_lin; ++l) {
    L->in_pad[l] = round_up(L->in_dim[l], 8);
  }
  for (int l = 0; l < L->n_lin; ++l) {
    L->out_pad[l] = (l + 1 < L->n_lin) ? L->in_pad[l + 1] : round_up(L->out_dim[l], 8);
    int64_t wsz = (int64_t)L->out_pad[l] * L->in_pad[l];
    L->off_w[l] = off;  off += wsz;
    L->off_wt[l] = off; off += wsz;
    L->off_b[l] = off;  off += L->out_pad[l];
    off = (off + 63) / 64 * 64;   // keep every block 256-byte aligned
  }
  L->packed_floats = off;
  // fp16x2-split copies of the weights (tensor-core operands of the tracer's MLP kernel and of the forward-type GEMMs),
  // written by the fold: W_l for every layer; W_l^T too for SDF nets (the input-gradient chain multiplies by it)
  for (int l = 0; l < L->n_lin; ++l) {
    L->off_h16[l] = off;
    off += (int64_t)L->out_pad[l] * L->in_pad[l];       // hi + lo halfs = out_pad * in_pad floats
    off = (off + 63) / 64 * 64;
    L->off_h16t[l] = 0;
    if (L->kind == 0) {
      L->off_h16t[l] = off;
      off += (int64_t)L->out_pad[l] * L->in_pad[l];
      off = (off + 63) / 64 * 64;
    }
  }
  L->packed_total_floats = off;
}

}  // namespace ironb

using namespace ironb;

extern "C" const char* ironb_last_error(void) { return g_err; }
extern "C" int ironb_version(void) { return 100; }
extern "C" int64_t ironb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Layer shapes of SDFNetwork.__init__ (models/fields.py:26-45): dims = [E] + [H]*n_layers + [d_out];
// the layer feeding the skip layer emits H - E features.
extern "C" int ironb_sdf_layout(int d_in, int d_out, int d_hidden, int n_layers, int skip_layer, int multires,
                                float scale, float beta, ironb_mlp_layout* out) {
  IRONB_REQUIRE(out != nullptr, "sdf_layout: null out");
  IRONB_REQUIRE(d_in == 3, "sdf_layout: d_in must be 3");
  IRONB_REQUIRE(n_layers + 1 <= IRONB_MAX_LIN && n_layers >= 1, "sdf_layout: n_layers out of range");
  IRONB_REQUIRE(multires >= 0 && multires <= 10, "sdf_layout: multires out of range");
  memset(out, 0, sizeof(*out));
  const int E = multires > 0 ? d_in * (1 + 2 * multires) : d_in;
  IRONB_REQUIRE(d_hidden > E, "sdf_layout: d_hidden must exceed the encoding width");
  IRONB_REQUIRE(skip_layer < 0 || (skip_layer >= 1 && skip_layer <= n_layers), "sdf_layout: bad skip layer");
  out->n_lin = n_layers + 1;
  out->kind = 0;
  out->d_in = d_in;
  out->multires = multires;
  out->pe_dim = E;
  out->skip_layer = skip_layer < 0 ? -1 : skip_layer;
  out->d_hidden = d_hidden;
  out->d_out = d_out;
  out->scale = scale;
  out->beta = beta;
  for (int l = 0; l < out->n_lin; ++l) {
    out->in_dim[l] = (l == 0) ? E : d_hidden;
    int o = (l == out->n_lin - 1) ? d_out : d_hidden;
    if (l + 1 == out->skip_layer) o -= E;
    out->out_dim[l] = o;
  }
  finish_layout(out);
  return IRONB_OK;
}

// RenderingNetwork.__init__ (models/fields.py:163-195): dims = [in_dim0] + [H]*n_layers + [d_out]; a skip layer s (at most
// one, 1 <= s <= n_layers) takes cat(h, input) / sqrt(2), i.e. its fan-in is H + in_dim0 while the layer before it still emits
// H features (:179-187) -- unlike the SDF net, whose layer before the skip shrinks.
extern "C" int ironb_matnet_layout_skip(int in_dim0, int d_out, int d_hidden, int n_layers, int skip_layer,
                                        ironb_mlp_layout* out) {
  IRONB_REQUIRE(out != nullptr, "matnet_layout: null out");
  IRONB_REQUIRE(n_layers + 1 <= IRONB_MAX_LIN && n_layers >= 0, "matnet_layout: n_layers out of range");
  IRONB_REQUIRE(in_dim0 > 0 && d_out > 0 && d_hidden > 0, "matnet_layout: bad dims");
  IRONB_REQUIRE(skip_layer < 0 || (skip_layer >= 1 && skip_layer <= n_layers && d_hidden % 8 == 0),
                "matnet_layout: skip layer must be in [1, n_layers] and d_hidden a multiple of 8");
  memset(out, 0, sizeof(*out));
  out->n_lin = n_layers + 1;
  out->kind = 1;
  out->d_in = 3;
  out->skip_layer = skip_layer < 0 ? -1 : skip_layer;
  out->d_hidden = d_hidden;
  out->d_out = d_out;
  out->scale = 1.f;
  out->beta = 0.f;
  for (int l = 0; l < out->n_lin; ++l) {
    out->in_dim[l] = (l == 0) ? in_dim0 : d_hidden + (l == out->skip_layer ? in_dim0 : 0);
    out->out_dim[l] = (l == out->n_lin - 1) ? d_out : d_hidden;
  }
  finish_layout(out);
  return IRONB_OK;
}
extern "C" int ironb_matnet_layout(int in_dim0, int d_out, int d_hidden, int n_layers, ironb_mlp_layout* out) {
  return ironb_matnet_layout_skip(in_dim0, d_out, d_hidden, n_layers, -1, out);
}
