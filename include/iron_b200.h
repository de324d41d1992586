/*
 * iron_b200 C ABI -- the drop-in boundary of the B200-native surface-rendering hot path.
 *
 * The reference (arthurlirui/IRON) has no FFI layer: its boundary is the Python module API
 * (SURVEY.md section 8b).  These entry points are what a binding for that API calls; the Python
 * modules in iron_b200/ (same class names / signatures as the reference) are the first such binding
 * (ctypes).  Every function:
 *   - takes raw DEVICE pointers (fp32 unless stated), element counts, and a CUDA stream handle
 *     (cudaStream_t passed as void*; NULL = legacy default stream),
 *   - allocates nothing: workspaces are sized by the *_workspace_bytes functions and passed in,
 *   - keeps no global state besides the last error string,
 *   - returns 0 on success, a negative IRONB_E* code for argument errors, or the positive
 *     cudaError_t of a failed launch.  There is no CPU fallback anywhere.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   ironb_ggx_fwd / ironb_ggx_bwd        GGXColocatedRenderer.forward        models/renderer_ggx.py:82-146 (+ autograd)
 *   ironb_mlp_fold / ironb_mlp_fold_bwd  nn.utils.weight_norm on every layer  models/fields.py:75-76, 194-195
 *   ironb_sdf_getall_fwd / _bwd          SDFNetwork.forward/sdf/gradient/get_all  models/fields.py:82-137
 *                                        (+ the double backward autograd runs through them)
 *   ironb_matnet_fwd / _bwd              RenderingNetwork.forward             models/fields.py:203-239 (+ autograd)
 *   ironb_camera_rays                    Camera.get_rays + intersect_sphere   models/raytracer.py:254-286, 223-237
 *   ironb_trace                          RayTracer.forward (sphere_tracing, ray_sampler, rootfind)
 *                                                                             models/raytracer.py:45-220
 *   ironb_depth_closing / ironb_sobel_depth  kornia closing / sobel in raytrace_camera (fill_holes, edge detection)
 *                                                                             models/raytracer.py:554-557, 569
 *   ironb_composite_fwd / _bwd           CompositeRenderer.forward            models/renderer_ggx.py:781-858 (+ autograd)
 *   ironb_pyramid_l2 / ironb_ssim_loss   PyramidL2Loss, ssim_loss_fn          models/image_losses.py:13-158 (+ autograd)
 *   ironb_adam_step                      six torch.optim.Adam instances       render_surface.py:112-113, 651-653
 *   ironb_neus_composite_fwd / _bwd, ironb_neus_upsample, ironb_neus_merge, ironb_neus_sections
 *                                        NeuSRenderer.render_core / up_sample / cat_z_vals / sample_pdf
 *                                                                             models/renderer.py:43-73, 192-351 (+ autograd)
 *   ironb_linear_fwd / _wgrad, ironb_relu_mask  the nn.Linear layers of the background NeRF  models/fields.py:241-322
 *   ironb_pack_tensors                   (new: the gradient bucket of the data-parallel step; the reference is single-GPU)
 */
#ifndef IRON_B200_H
#define IRON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IRONB_MAX_LIN 12

#define IRONB_OK 0
#define IRONB_EINVAL (-1)   /* bad argument (null pointer, negative size, unsupported shape) */
#define IRONB_ENOSUP (-2)   /* configuration outside what the kernels were built for */

/* Geometry of one weight-normalised MLP and where its folded weights live inside one packed fp32
 * device buffer.  Built by ironb_sdf_layout / ironb_matnet_layout, then treated as read-only. */
typedef struct ironb_mlp_layout {
  int32_t n_lin;                   /* number of linear layers: 9 for the SDF net, 5 for a material net */
  int32_t kind;                    /* 0 = SDF net (softplus, PE input, skip), 1 = material net (ReLU) */
  int32_t d_in;                    /* raw coordinate dim (3) */
  int32_t multires;                /* PE frequencies of the SDF input (6); 0 = none */
  int32_t pe_dim;                  /* E = d_in*(1+2*multires) (39) */
  int32_t skip_layer;              /* layer whose input is cat(h, PE)/sqrt(2); -1 = none   fields.py:88-89 */
  int32_t d_hidden;
  int32_t d_out;                   /* 257 for the SDF net */
  float scale;                     /* SDFNetwork.scale   fields.py:83,98 */
  float beta;                      /* Softplus beta (100)  fields.py:80 */
  int32_t in_dim[IRONB_MAX_LIN];   /* true fan-in  K_l */
  int32_t out_dim[IRONB_MAX_LIN];  /* true fan-out N_l (H-E for the layer before the skip) */
  int32_t in_pad[IRONB_MAX_LIN];   /* K_l rounded up to 8 (zero filled) */
  int32_t out_pad[IRONB_MAX_LIN];  /* for l < n_lin-1: in_pad[l+1]; last layer: N rounded up to 8 */
  int64_t off_w[IRONB_MAX_LIN];    /* W_l  [out_pad][in_pad] row-major, float offset into the packed buffer */
  int64_t off_wt[IRONB_MAX_LIN];   /* W_l^T [in_pad][out_pad] */
  int64_t off_b[IRONB_MAX_LIN];    /* b_l  [out_pad] */
  int64_t packed_floats;           /* size of the fp32 part (W, W^T, b of every layer): the size of a dpacked buffer */
  int64_t off_h16[IRONB_MAX_LIN];  /* fp16x2-split copy of W_l written by ironb_mlp_fold -- hi [out_pad][in_pad] halfs, then
                                      lo = fp16((w - hi) * 2^11), same shape; float offset into the packed buffer.  The
                                      tensor-core operands of the default tracer (csrc/mlp_h16.cu) and of the forward-type
                                      GEMMs (csrc/gemm_h16.cuh) */
  int64_t off_h16t[IRONB_MAX_LIN]; /* the same for W_l^T [in_pad][out_pad] (SDF nets: the input-gradient chain); 0 = absent */
  int64_t packed_total_floats;     /* size of the packed buffer ironb_mlp_fold writes (fp32 part + fp16 copies), in floats */
} ironb_mlp_layout;

const char* ironb_last_error(void);
int ironb_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t ironb_launch_count(void);

/* GEMM arithmetic of the differentiable part (get_all forward / double backward, material MLPs):
 * 2 = tcgen05 with fp16x2 pre-split operands for the forward-type products (get_all forward, input-gradient chain,
 *     material forward: O(1) operands) and the 3xTF32 split for products with gradient operands (the default),
 * 1 = tcgen05 3xTF32 split everywhere, 0 = fp32 FFMA tiles.  All fp32-grade.  Returns the previous mode.
 * IRONB_GEMM=simt / tf32 selects 0 / 1 at start-up. */
int ironb_set_gemm_mode(int mode);
/* Tracer implementation: 2 = batched tcgen05 rounds, fp16x2-split operands (default); 0 = fused persistent fp32-FFMA
 * kernels (exact fp32 association).  Returns the previous mode.  IRONB_TRACE=fused selects 0 at start-up.
 * (Mode 1, the 3xTF32 predecessor of mode 2, was retired in round 2; 1 is accepted and means 2.) */
int ironb_set_trace_mode(int mode);
/* Truncation de-bias factor of the default tracer's tcgen05 accumulators (csrc/mlp_h16.cu: mlp16_debias); g < 0 only
 * queries.  Returns the previous factor.  IRONB_MLP_DEBIAS sets it at start-up; 0 switches the correction off. */
float ironb_set_mlp_debias(float g);
/* Shape of the default tracer's fused MLP kernel: 0 = automatic (64 output features per CTA in clusters of H/64 CTAs for
 * tracer calls of at most ~8,192 rays -- the latency shape -- else 128 per CTA in clusters of H/128), 64 / 128 force one.
 * Returns the previous setting.  IRONB_MLP_RN sets it at start-up. */
int ironb_set_mlp_rn(int rn);
/* Debug builds (IRONB_NVCC_EXTRA=-DIRONB_DEBUG_HANG): every mbarrier wait of the tcgen05 kernels is bounded (~2 s); a wait that
 * expires writes {tag, block, thread, parity, position} into this MAPPED PINNED host buffer (>= 4 KiB, zeroed; word 0 counts
 * the records, record i starts at word 8 + 8 i) and traps, so a pipeline slip is a located fault instead of a silent hang.
 * Returns 1 if the hooks are compiled in, 0 if not. */
int ironb_debug_hang_buffer(void* mapped_host);
/* C[M][ldc] = A[M][lda] * B[N][ldb]^T (fp32, K-major operands, N/K/ld multiples of 4): unit-test entry of both GEMMs. */
int ironb_gemm_nt(const float* A, int lda, const float* B, int ldb, int M, int N, int K, float* C, int ldc,
                  int mode, void* stream);

/* The same product on fp16x2 PRE-SPLIT operands (csrc/gemm_h16.cuh): hi = fp16(x), lo = fp16((x - hi) * 2^11) as separate
 * K-major half arrays, pitches in halfs (multiples of 8).  Unit-test entry of the forward-type tensor-core GEMM. */
int ironb_gemm_nt_h16(const void* Ah, const void* Al, int lda, const void* Bh, const void* Bl, int ldb, int M, int N, int K,
                      float* C, int ldc, void* stream);

/* C[Nd][ldc] += A[M][lda]^T * B[M][ldb] (the weight-gradient product; C is accumulated): unit-test entry. */
int64_t ironb_gemm_tn_scratch_bytes(int M, int Nd, int Kd);
int ironb_gemm_tn(const float* A, int lda, const float* B, int ldb, int M, int Nd, int Kd, float* C, int ldc,
                  int mode, void* scratch, void* stream);

/* ---------------------------------------------------------------- layouts (host only) */
int ironb_sdf_layout(int d_in, int d_out, int d_hidden, int n_layers, int skip_layer, int multires,
                     float scale, float beta, ironb_mlp_layout* out);
int ironb_matnet_layout(int in_dim0, int d_out, int d_hidden, int n_layers, ironb_mlp_layout* out);
/* The same with RenderingNetwork(skip_in=[s]) (models/fields.py:179-187, 226-227: layer s reads cat(h, input) / sqrt 2; the
 * stage-1 colour network uses n_layers = 8, skip_in = [4]); skip_layer < 0 = none. */
int ironb_matnet_layout_skip(int in_dim0, int d_out, int d_hidden, int n_layers, int skip_layer, ironb_mlp_layout* out);

/* ---------------------------------------------------------------- weight norm
 * W = v * (g / ||v||_row)  (old-style nn.utils.weight_norm, dim=0).  v/g/b: arrays of n_lin device
 * pointers (v[l]: [out,in] row-major, g[l]: [out] or NULL = no weight norm, b[l]: [out]). */
int ironb_mlp_fold(const ironb_mlp_layout* lay, const float* const* v, const float* const* g,
                   const float* const* b, float* packed, void* stream);
/* dpacked has the layout of `packed` and holds dL/dW at off_w and dL/db at off_b.
 * Writes dv[l] ([out,in]), dg[l] ([out], may be NULL when g[l] is NULL), db[l] ([out]). */
int ironb_mlp_fold_bwd(const ironb_mlp_layout* lay, const float* const* v, const float* const* g,
                       const float* dpacked, float* const* dv, float* const* dg, float* const* db,
                       void* stream);

/* ---------------------------------------------------------------- SDF network
 * Forward over M points x[M,3].  Outputs (any may be NULL): y[M] (sdf), feat[M,d_out-1], grad[M,3]
 * (= d y / d x, closed form of autograd.grad in SDFNetwork.gradient/get_all).
 * save != 0 keeps per-layer state in `ws` for ironb_sdf_getall_bwd. */
int64_t ironb_sdf_getall_workspace_bytes(const ironb_mlp_layout* lay, int64_t M, int want_grad, int save);
int ironb_sdf_getall_fwd(const ironb_mlp_layout* lay, const float* packed, const float* x, int64_t M,
                         float* y, float* feat, float* grad, int save, void* ws, int64_t ws_bytes,
                         void* stream);
/* Gradients of <ybar,y> + <fbar,feat> + <nbar,grad> (each upstream may be NULL) w.r.t. the folded
 * weights/biases, ACCUMULATED into dpacked (caller zeroes it).  This is the double backward of
 * fields.py:106-137 in closed form. `ws` is the saved forward workspace and
 * is clobbered.  wgrad_stream (may be NULL = `stream`): a second stream for the 17 weight-gradient products, which only
 * consume what the two back-propagation chains produce; the call forks to it with events and joins before returning its
 * last work to `stream` (capturable; both streams are ordered after the call like one). */
int ironb_sdf_getall_bwd(const ironb_mlp_layout* lay, const float* packed, const float* x, int64_t M,
                         const float* ybar, const float* fbar, const float* nbar, void* ws,
                         int64_t ws_bytes, float* dpacked, void* stream, void* wgrad_stream);

/* ---------------------------------------------------------------- material networks (RenderingNetwork)
 * mode 0 = 'idr'         input = cat(PE_p(points), PE_v(view_dirs), normals, feats)
 * mode 1 = 'no_view_dir' input = cat(PE_p(points), normals, feats)
 * mode 2 = 'no_normal'   input = cat(PE_p(points), PE_v(view_dirs), feats)
 * mode 3 = 'points_only' input = cat(PE_p(points), feats)                       fields.py:209-222
 * out = out_scale * (lin_last(..) + out_bias); squeeze != 0 applies squeeze_scale*sigmoid   fields.py:232-238 */
typedef struct ironb_matnet_cfg {
  int32_t mode;
  int32_t multires;        /* PE of points, 0 = raw */
  int32_t multires_view;   /* PE of view dirs, 0 = raw */
  int32_t d_feature;
  int32_t squeeze;
  float out_bias, out_scale, squeeze_scale;
} ironb_matnet_cfg;

int ironb_matnet_in_dim(const ironb_matnet_cfg* cfg);
int64_t ironb_matnet_workspace_bytes(const ironb_mlp_layout* lay, int64_t M);
int ironb_matnet_fwd(const ironb_mlp_layout* lay, const ironb_matnet_cfg* cfg, const float* packed,
                     const float* points, const float* normals, const float* view_dirs,
                     const float* feats, int64_t M, float* out, void* ws, int64_t ws_bytes, void* stream);
/* d_points/d_normals/d_view/d_feats may be NULL; written (not accumulated). dpacked accumulated.
 * wgrad_stream: as for ironb_sdf_getall_bwd (may be NULL). */
int ironb_matnet_bwd(const ironb_mlp_layout* lay, const ironb_matnet_cfg* cfg, const float* packed,
                     int64_t M, const float* out, const float* gout, void* ws, int64_t ws_bytes,
                     float* dpacked, float* d_points, float* d_normals, float* d_view, float* d_feats,
                     void* stream, void* wgrad_stream);

/* ---------------------------------------------------------------- GGX colocated shading
 * light: device scalar.  dist[M], normal[M,3], viewdir[M,3], kd[M,3], ks[M,3], alpha[M].
 * trans[5000], diff_trans[50]: the Mitsuba rough-plastic tables (renderer_ggx.py:65-74). */
int ironb_ggx_fwd(const float* light, const float* dist, const float* normal, const float* viewdir,
                  const float* kd, const float* ks, const float* alpha, const float* trans,
                  const float* diff_trans, int64_t M, float* diffuse_rgb, float* specular_rgb,
                  float* rgb, void* stream);
/* Upstream grads g_diffuse/g_specular/g_rgb [M,3] may each be NULL.  d_light is ACCUMULATED
 * (atomicAdd; caller zeroes), the others are written.  d_viewdir may be NULL. */
int ironb_ggx_bwd(const float* light, const float* dist, const float* normal, const float* viewdir,
                  const float* kd, const float* ks, const float* alpha, const float* trans,
                  const float* diff_trans, int64_t M, const float* g_diffuse, const float* g_specular,
                  const float* g_rgb, float* d_light, float* d_dist, float* d_normal, float* d_viewdir,
                  float* d_kd, float* d_ks, float* d_alpha, void* stream);

/* ---------------------------------------------------------------- camera rays + unit-sphere clip
 * uv[N,2] pixel coordinates; Kinv3[9], R_c2w[9] row-major 3x3 blocks of K^-1 and C2W; origin[3].
 * All on the device.  Outputs ray_o[N,3], ray_d[N,3], ray_d_norm[N]; if min_dis != NULL also the
 * sphere clip (hit[N] u8, min_dis[N], max_dis[N]) for radius r. */
int ironb_camera_rays(const float* uv, int64_t N, const float* Kinv3, const float* R_c2w,
                      const float* origin, float r, float* ray_o, float* ray_d, float* ray_d_norm,
                      uint8_t* hit, float* min_dis, float* max_dis, void* stream);
int ironb_intersect_sphere(const float* ray_o, const float* ray_d, int64_t N, float r, uint8_t* hit,
                           float* min_dis, float* max_dis, void* stream);

/* ---------------------------------------------------------------- tracer
 * One call == one RayTracer.forward (the bisection count is coupled across the rays of a call,
 * raytracer.py:204-214).  linspace: the n_steps sample positions (torch.linspace(0,1,n_steps)).
 * Outputs: conv[N] u8, points[N,3], sdf[N], dist[N].  stats (device, int64[8], may be NULL):
 * [0] evals sphere tracing, [1] evals sampler, [2] evals bisection, [3] sampler rays, [4] root rays,
 * [5] k_max, [6] tile evaluations. */
int64_t ironb_trace_workspace_bytes(const ironb_mlp_layout* lay, int64_t N);
int ironb_trace(const ironb_mlp_layout* lay, const float* packed, const float* ray_o,
                const float* ray_d, const float* min_dis, const float* max_dis,
                const uint8_t* work_mask, int64_t N, float sdf_threshold, int sphere_tracing_iters,
                int n_steps, const float* linspace, uint8_t* conv, float* points, float* sdf,
                float* dist, int64_t* stats, void* ws, int64_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- hit-point glue (fused elementwise)
 * Stable compaction of the hit mask: idx[0..count) = ascending ray ids with mask != 0. count: device int32. */
int ironb_compact_mask(const uint8_t* mask, int64_t N, int32_t* idx, int32_t* count, void* stream);
/* Gather rows: dst[i,:] = src[idx[i],:]  (width floats per row). */
int ironb_gather_rows(const float* src, const int32_t* idx, int64_t M, int width, float* dst, void* stream);
/* Scatter rows into a zero-initialised dense buffer: dst[idx[i],:] = src[i,:]. */
int ironb_scatter_rows(const float* src, const int32_t* idx, int64_t M, int width, float* dst, void* stream);

/* Hole filling of raytrace_camera (models/raytracer.py:554-557): 3x3 morphological closing of the [H,W] depth map
 * (kornia.morphology.closing with an all-ones kernel, geodesic border).  tmp / out: H*W floats each. */
int ironb_depth_closing(const float* depth, int H, int W, float* tmp, float* out, void* stream);
/* Depth-edge detector of raytrace_camera (models/raytracer.py:569): kornia.filters.sobel magnitude of the [H,W] depth map. */
int ironb_sobel_depth(const float* depth, int H, int W, float* out, void* stream);

/* ---------------------------------------------------------------- optimiser
 * Multi-tensor Adam with torch.optim.Adam's arithmetic (amsgrad = False): replaces the six torch.optim.Adam instances the
 * stage-2 loop steps (render_surface.py:112-113, 651-653; models/network_conf.py:707-716) with one launch.
 * tensors_dev: DEVICE array of n_tensors descriptors (grad == NULL: tensor skipped); max_numel: largest numel (grid sizing);
 * step_dev: device int32 holding the number of steps taken so far (read as t = *step + 1, then advanced by one). */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
  float lr;
  float weight_decay;
} ironb_adam_tensor;
int ironb_adam_step(const ironb_adam_tensor* tensors_dev, int n_tensors, int64_t max_numel, double beta1, double beta2,
                    double eps, int* step_dev, void* stream);

/* ---------------------------------------------------------------- CompositeRenderer ("comp2")
 * models/renderer_ggx.py:781-858: rough-plastic diffuse lobe (the GGX tables) + exact-Fresnel conductor lobe + dielectric
 * microfacet lobe, colocated flash.  M points; kd/ks [M,3]; alpha, metallic_eta, metallic_k, dielectric_eta, dist [M];
 * normal, viewdir [M,3]; light: device scalar.  Outputs [M,3] each.  The reference accumulates rgb in place into its
 * diffuse tensor, so its "diffuse_rgb" output IS rgb: pass the sum of both upstream gradients as g_rgb.
 * Backward: any upstream may be NULL; d_light is ACCUMULATED (caller zeroes it), the others are written. */
int ironb_composite_fwd(const float* light, const float* dist, const float* normal, const float* viewdir, const float* kd,
                        const float* ks, const float* alpha, const float* metallic_eta, const float* metallic_k,
                        const float* dielectric_eta, const float* trans, const float* diff_trans, int64_t M, float* rgb,
                        float* specular_rgb, float* metallic_rgb, float* dielectric_rgb, void* stream);
int ironb_composite_bwd(const float* light, const float* dist, const float* normal, const float* viewdir, const float* kd,
                        const float* ks, const float* alpha, const float* metallic_eta, const float* metallic_k,
                        const float* dielectric_eta, const float* trans, const float* diff_trans, int64_t M,
                        const float* g_rgb, const float* g_specular, const float* g_metallic, const float* g_dielectric,
                        float* d_light, float* d_dist, float* d_normal, float* d_kd, float* d_ks, float* d_alpha,
                        float* d_metallic_eta, float* d_metallic_k, float* d_dielectric_eta, void* stream);

/* ---------------------------------------------------------------- glue of a shading step (csrc/glue.cu)
 * The element-wise chains between the big kernels, one launch forward and one backward each (the reference issues them as
 * separate ATen ops).  M rows; [M,3] / [M] contiguous fp32; any upstream gradient pointer may be NULL (= zeros).
 *   unit_dist : n = g / (|g| + 1e-10), dist = |x - o|                          render_surface.py:135-146
 *   reparam   : forward value == x; d f = -sum_c v_c / clamp(g.v, 1e-4) d x_c   models/raytracer.py:17-24
 *   matpost   : kd = |a|, ks = mean_c |b_c| (|b| if is_metal), alpha = |c| + 0.01   models/rendering_func.py:5-16
 *   eik_sum   : out += sum_m w_m (|g_m| - 1)^2  (w NULL = ones)                 render_surface.py:580-583, 601-603
 *   roughrange: loss = weight * mean_{w_m != 0, r_m > value} (r_m - value), 0 if empty; acc = 2 zeroed floats   :609-613
 *   mask_rows : dst_t[m][:] = src_t[m][:] * w[m] for n <= 8 row tensors       (dense shading: zero the non-hit pixels) */
int ironb_unit_dist_fwd(const float* g, const float* x, const float* o, int64_t M, float* n, float* dist, void* stream);
int ironb_unit_dist_bwd(const float* g, const float* x, const float* o, const float* dn, const float* ddist, int64_t M,
                        float* dg, float* dx, void* stream);
int ironb_reparam_bwd(const float* g, const float* v, const float* dx, int64_t M, float* df, void* stream);
int ironb_matpost_fwd(const float* a, const float* b, const float* c, int64_t M, int is_metal, float* kd, float* ks, float* alpha,
                      void* stream);
int ironb_matpost_bwd(const float* a, const float* b, const float* c, const float* dkd, const float* dks, const float* dalpha,
                      int64_t M, int is_metal, float* da, float* db, float* dc, void* stream);
int ironb_eik_sum_fwd(const float* g, const float* w, int64_t M, float* out, void* stream);
int ironb_eik_sum_bwd(const float* g, const float* w, const float* up, int64_t M, float* dg, void* stream);
int ironb_roughrange_fwd(const float* r, const float* w, int64_t M, float value, float weight, float* acc, float* loss,
                         void* stream);
int ironb_roughrange_bwd(const float* r, const float* w, const float* acc, const float* up, int64_t M, float value, float weight,
                         float* dr, void* stream);
int ironb_mask_rows(const float* const* src, float* const* dst, const int* width, int n, const float* w, int64_t M, void* stream);
/* ---------------------------------------------------------------- stage-1 NeuS volume renderer (csrc/neus.cu)
 * The per-ray part of NeuSRenderer.render_core (models/renderer.py:248-351): SDF values / SDF gradients / colours at the n
 * section midpoints of N rays (row-major [N][n], [N][n][3]) -> composited colour [N][3], weights [N][n_tot], cdf [N][n],
 * inside_sphere [N][n] and the eikonal term gradient_error [1]; acc [2] keeps the two eikonal sums for the backward.
 * bg_density [N][n_tot] / bg_dists [N][n_tot] / bg_color [N][n_tot][3]: the background NeRF's raw density, the lengths of its
 * sections and its colour (render_core_outside, :145-170: alpha = 1 - exp(-softplus(density) dist) is formed in the kernel),
 * NULL without a background model (then n_tot == n); bg_rgb [3]: fixed background colour or NULL; inv_s [1] on the device. */
int ironb_neus_composite_fwd(const float* ray_o, const float* ray_d, const float* mid_z, const float* dists, const float* sdf,
                             const float* grad, const float* color, const float* inv_s, const float* bg_density,
                             const float* bg_dists, const float* bg_color, const float* bg_rgb, int64_t N, int n, int n_tot,
                             float cos_anneal_ratio, float* out_color, float* weights, float* cdf, float* inside, float* acc,
                             float* gradient_error, void* stream);
/* Backward of the above (replaces autograd through :277-322): upstream d_color [N][3], d_weights [N][n_tot] or NULL,
 * d_gradient_error [1] or NULL -> d_sdf [N][n], d_grad [N][n][3], d_colors [N][n][3], d_inv_s [1], and with a background
 * d_bg_density [N][n_tot], d_bg_color [N][n_tot][3].  The forward quantities are recomputed from the same inputs. */
int ironb_neus_composite_bwd(const float* ray_o, const float* ray_d, const float* mid_z, const float* dists, const float* sdf,
                             const float* grad, const float* color, const float* inv_s, const float* bg_density,
                             const float* bg_dists, const float* bg_color, const float* bg_rgb, int64_t N, int n, int n_tot,
                             float cos_anneal_ratio, const float* weights, const float* acc, const float* d_color,
                             const float* d_weights, const float* d_gradient_error, float* d_sdf, float* d_grad,
                             float* d_colors, float* d_inv_s, float* d_bg_density, float* d_bg_color, void* stream);
/* Sections of a ray batch (:254-262, :145-160): dists [N][n] (last = sample_dist), mid [N][n], the section midpoints pts
 * [N][n][3] (outside != 0: the background model's inverted-sphere points, [N][n][4]) and, if dirs != NULL, the per-point view
 * directions [N][n][3]. */
int ironb_neus_sections(const float* ray_o, const float* ray_d, const float* z, int64_t N, int n, float sample_dist, int outside,
                        float* dists, float* mid, float* pts, float* dirs, void* stream);
/* One hierarchical-sampling step, NeuSRenderer.up_sample + sample_pdf(det=True) (:192-232, :43-73): from n ascending samples
 * z [N][n] with SDF values sdf [N][n] and a fixed sharpness inv_s to m new samples new_z [N][m] at the quantiles of the interval
 * weights' CDF, and (new_pts != NULL) their points o + d z [N][m][3].  n <= 256. */
int ironb_neus_upsample(const float* ray_o, const float* ray_d, const float* z, const float* sdf, int64_t N, int n, int m,
                        float inv_s, float* new_z, float* new_pts, void* stream);
/* cat_z_vals (:234-246): merges the ascending lists z [N][n] and new_z [N][m] into out_z [N][n+m]; out_sdf != NULL: the SDF values
 * (sdf [N][n], new_sdf [N][m]) travel with their samples. */
int ironb_neus_merge(const float* z, const float* sdf, int n, const float* new_z, const float* new_sdf, int m, int64_t N,
                     float* out_z, float* out_sdf, void* stream);
/* ---------------------------------------------------------------- plain nn.Linear layers on the tensor cores
 * The stage-1 background model (NeRF, models/fields.py:241-322) is plain Linear + ReLU layers:
 *   ironb_linear_fwd    y[M][ldc] = act(x[M][lda] W[N][ldb]^T + b)           (relu != 0: ReLU; bias may be NULL)
 *   ironb_relu_mask     out = dy where y > 0 else 0                            (n elements, 16-byte aligned buffers)
 *   ironb_linear_wgrad  dW[N][ldc] += dy[M][lda]^T x[M][ldb], db[N] += column sums of dy (zero both first; db may be NULL;
 *                       scratch: ironb_gemm_tn_scratch_bytes(M, N, K))
 * the input gradient is ironb_gemm_nt(dy, W^T).  N, K and the pitches multiples of 4; fp32-grade (3xTF32) arithmetic. */
int ironb_linear_fwd(const float* x, int lda, const float* W, int ldb, const float* bias, int M, int N, int K, int relu, float* y,
                     int ldc, void* stream);
int ironb_relu_mask(const float* dy, const float* y, int64_t n, float* out, void* stream);
int ironb_linear_wgrad(const float* dy, int lda, const float* x, int ldb, int M, int N, int K, float* dW, int ldc, float* db,
                       void* scratch, void* stream);
/* Gradient bucket of the data-parallel step (the reference is single-GPU, no interface replaced):
 * flat[off[t] .. off[t] + off[n + t]) = scale * src[t][:] for t < n, zeros where src[t] is NULL.  src_dev (n pointers) and
 * off_dev (n element offsets followed by n element counts) are DEVICE arrays, so the launch can be a CUDA-graph node.  max_numel sizes the grid. */
int ironb_pack_tensors(const float* const* src_dev, const int64_t* off_dev, int n, int64_t max_numel, float* flat, float scale,
                       void* stream);

/* ---------------------------------------------------------------- patch losses
 * One [C][H][W] image pair (C <= 4), addressed through ELEMENT strides {channel, row, column}, so the [1,3,H,W] view of
 * the renderer's [H,W,3] buffer is read in place.  Value and gradient w.r.t. the first image in one call; `ws` from
 * ironb_patch_loss_workspace_bytes (one per call in flight).  No host read-back: capturable into a CUDA graph.
 *
 * ironb_pyramid_l2: PyramidL2Loss.forward, models/image_losses.py:29-48 (7x7 sigma-1 Gaussian, 2x average pooling, five
 * levels; H, W >= 16).  *loss is ACCUMULATED (the caller zeroes it).  grad may be NULL.
 * ironb_ssim_loss: ssim_loss_fn, models/image_losses.py:97-158 (11-tap sigma-1.5 separable window without padding,
 * channel mean, padded back with 1.0, mean over the 11x11-eroded mask; mask: H*W bytes or NULL = unmasked mean of the
 * unpadded map).  *loss is WRITTEN.  grad may be NULL. */
int64_t ironb_patch_loss_workspace_bytes(int C, int H, int W);
int ironb_pyramid_l2(const float* pred, const int64_t* pred_strides, const float* trgt, const int64_t* trgt_strides, int C,
                     int H, int W, float* loss, float* grad, const int64_t* grad_strides, void* ws, int64_t ws_bytes,
                     void* stream);
int ironb_ssim_loss(const float* X, const int64_t* x_strides, const float* Y, const int64_t* y_strides, const uint8_t* mask,
                    int C, int H, int W, float data_range, int win_size, float win_sigma, float K1, float K2, float* loss,
                    float* grad, const int64_t* grad_strides, void* ws, int64_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IRON_B200_H */
