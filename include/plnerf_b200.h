/*
 * plnerf_b200 -- C ABI of the B200-native (sm_100a) PL-NeRF ray-rendering hot path.
 *
 * The reference (mikacuy/PL-NeRF) is pure Python/PyTorch and has no FFI layer; its pluggable seams
 * are Python callables.  Each entry point below replaces one of those
 * callables; the citation is the reference interface it stands in for (paths relative to the
 * reference repo root).  The Python host side (pl-nerf_b200/run_plnerf.py, run_nerf_helpers.py)
 * binds these through ctypes with torch tensors as device buffers; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data unless stated; buffers are caller-owned,
 *     row-major, contiguous unless a stride is given;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - every function returns 0 on success, <0 on error (PLNERF_E_*), never throws or aborts;
 *     plnerf_last_error() returns a thread-local message for the last failure;
 *   - optional pointers may be NULL where documented;
 *   - there is no CPU fallback: without a CUDA device every compute entry returns PLNERF_E_CUDA.
 */
#ifndef PLNERF_B200_H
#define PLNERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLNERF_ABI_VERSION 2

enum {
  PLNERF_OK = 0,
  PLNERF_E_BADARG = -1,      /* null pointer, negative size, misaligned buffer */
  PLNERF_E_UNSUPPORTED = -2, /* shape outside what the sm_100a kernels implement */
  PLNERF_E_CUDA = -3,        /* CUDA runtime error (message in plnerf_last_error) */
  PLNERF_E_WORKSPACE = -4    /* workspace too small */
};

enum { PLNERF_MODE_CONSTANT = 0, PLNERF_MODE_LINEAR = 1 };      /* --mode, run_plnerf.py:906 */
enum { PLNERF_COLOR_MIDPOINT = 0, PLNERF_COLOR_LEFT = 1 };      /* --color_mode, run_plnerf.py:908 */
enum { PLNERF_PREC_BF16 = 0,    /* one bf16 tcgen05 MMA per product, fp32 accumulate (fast)      */
       PLNERF_PREC_BF16X3 = 1   /* hi/lo bf16 split, 3 MMAs per product (~fp32 products, parity) */ };

#define PLNERF_MAX_DEPTH 16

/* Architecture of one NeRF MLP: mirrors NeRF.__init__ (run_nerf_helpers.py:77-103). */
typedef struct plnerf_net_desc {
  int32_t D;               /* trunk depth (8) */
  int32_t W;               /* trunk width (256; the only width the tcgen05 kernel implements) */
  int32_t input_ch;        /* 3 + 6*multires (63) or 3 */
  int32_t input_ch_views;  /* 3 + 6*multires_views (27), 3, or 0 */
  int32_t output_ch;       /* only read when use_viewdirs == 0 (4 or 5) */
  int32_t use_viewdirs;
  int32_t n_skips;
  int32_t skips[PLNERF_MAX_DEPTH]; /* layer indices i after which [input_pts, h] is concatenated */
} plnerf_net_desc;

/* fp32 parameters in the reference's state_dict layout ([out,in] row-major weights), i.e. the
 * tensors of NeRF.state_dict() (run_nerf_helpers.py:85-103), passed by device pointer. */
typedef struct plnerf_net_params {
  const float* pts_w[PLNERF_MAX_DEPTH]; /* pts_linears.i.weight */
  const float* pts_b[PLNERF_MAX_DEPTH]; /* pts_linears.i.bias   */
  const float* views_w;   /* views_linears.0.weight [W/2, W + input_ch_views] */
  const float* views_b;
  const float* feature_w; /* feature_linear [W,W]   (use_viewdirs) */
  const float* feature_b;
  const float* alpha_w;   /* alpha_linear [1,W]     (use_viewdirs) */
  const float* alpha_b;
  const float* rgb_w;     /* rgb_linear [3,W/2]     (use_viewdirs) */
  const float* rgb_b;
  const float* output_w;  /* output_linear [output_ch,W] (!use_viewdirs) */
  const float* output_b;
} plnerf_net_params;

/* Gradient buffers with the same names / shapes as plnerf_net_params (fp32, accumulated with +=, so
 * the caller zeroes them or lets them alias param.grad). */
typedef struct plnerf_net_grads {
  float* pts_w[PLNERF_MAX_DEPTH];
  float* pts_b[PLNERF_MAX_DEPTH];
  float* views_w;
  float* views_b;
  float* feature_w;
  float* feature_b;
  float* alpha_w;
  float* alpha_b;
  float* rgb_w;
  float* rgb_b;
  float* output_w;
  float* output_b;
} plnerf_net_grads;

/* Everything render_rays() receives besides tensors (run_plnerf.py:627-646). */
typedef struct plnerf_render_cfg {
  int32_t N_samples;
  int32_t N_importance;
  int32_t mode;           /* PLNERF_MODE_*; constant_init is applied by the caller (mode=constant) */
  int32_t color_mode;     /* PLNERF_COLOR_* */
  int32_t white_bkgd;
  int32_t lindisp;
  int32_t farcolorfix;
  int32_t perturb;        /* != 0: stratified jitter + random u (run_plnerf.py:691-705) */
  float raw_noise_std;
  float zero_tol;         /* 1e-4 */
  float epsilon;          /* 1e-3 */
  int32_t multires;       /* xyz PE frequencies (10); -1 = identity */
  int32_t multires_views; /* dir PE frequencies (4);  -1 = identity */
  int32_t precision;      /* PLNERF_PREC_* */
  uint64_t seed;          /* Philox key for draws not supplied explicitly */
  uint64_t ray_id_offset; /* global index of ray 0 (keeps RNG invariant to chunking / sharding) */
} plnerf_render_cfg;

/* Output / intermediate buffers of one render_rays call.  NULL = not wanted.  */
typedef struct plnerf_render_out {
  float* rgb_map;   /* [n,3]  */
  float* disp_map;  /* [n]    */
  float* acc_map;   /* [n]    */
  float* depth_map; /* [n]    */
  float* raw;       /* [n, S_last, 4] raw of the last pass (retraw), may be NULL */
  float* rgb0;      /* [n,3]  coarse outputs, only written when N_importance > 0 */
  float* disp0;
  float* acc0;
  float* depth0;
  float* z_std;     /* [n]    */
  float* z_vals;    /* [n, N_samples+N_importance] merged depths (or [n,N_samples]), may be NULL */
  int64_t* inds;    /* [n, N_importance] searchsorted indices, may be NULL */
} plnerf_render_out;

const char* plnerf_last_error(void);
int plnerf_abi_version(void);
/* Number of CUDA kernels this library has launched since load (claim for bench.py gpu_launches). */
uint64_t plnerf_launch_count(void);

/* ---- a5: Embedder.embed / get_embedder (run_nerf_helpers.py:24-72) ---------------------------
 * x [n,3] -> out [n, 3+6*multires]  (multires < 0: copy). */
int plnerf_encode(const float* x, int64_t n, int multires, float* out, void* stream);

/* ---- f-1 (next row): ray generation and packing done by render(), run_plnerf.py:138-164 ----------
 * get_rays (run_nerf_helpers.py:162-171) from a pose c2w [3, >=4] (row stride c2w_ld), or the given
 * rays_o / rays_d [n,3] when c2w == NULL; viewdirs = d/|d| taken before NDC (:145-150, from c2w even
 * when c2w_staticcam supplies origins and directions); ndc_rays (:184-201) with the near plane
 * ndc_near (render() passes 1.0) and the scalar factors ndc_cx = -1/(W/(2 focal)), ndc_cy =
 * -1/(H/(2 focal)) the caller computed in double; near/far columns.  out [n, 8 | 11] row-major with
 * row stride `stride` floats.  n = H*W when c2w != NULL. */
int plnerf_pack_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int c2w_ld,
                     const float* c2w_staticcam, int c2w_staticcam_ld, const float* rays_o,
                     const float* rays_d, int64_t n, int ndc, float ndc_cx, float ndc_cy, float ndc_near,
                     float near, float far, int use_viewdirs, float* out, int stride, void* stream);

/* ---- f-2 (next row): the training loop's per-iteration ray selection, run_plnerf.py:1259-1280 ----
 * The reference builds get_rays for the WHOLE image (H*W*6 floats), a [H*W,2] coordinate grid, then
 * gathers the N_rand chosen pixels.  Here the rays of the chosen pixels only: pix [n] int64 flat pixel
 * ids (row*W + col, each in [0, H*W); the caller guarantees the range), same arithmetic per ray as
 * plnerf_pack_rays with a pose, so out[i] is bit-identical to row pix[i] of the full-image packing. */
int plnerf_pack_pixel_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int c2w_ld,
                           const int64_t* pix, int64_t n, int ndc, float ndc_cx, float ndc_cy, float ndc_near,
                           float near, float far, int use_viewdirs, float* out, int stride, void* stream);

/* ---- a3 (first part): stratified depths, render_rays (run_plnerf.py:683-705) -----------------
 * rays [n, stride] (cols 6,7 = near, far).  t_rand [n,Ns] explicit jitter in [0,1) or NULL;
 * with t_rand == NULL, perturb != 0 draws Philox(seed, ray id, sample) jitter. */
int plnerf_stratified_z(const float* rays, int64_t n, int stride, int N_samples, int lindisp,
                        int perturb, const float* t_rand, uint64_t seed, uint64_t ray_id_offset,
                        float* z_vals, void* stream);

/* ---- weights: repack NeRF.state_dict() tensors into the kernel's streaming layout -------------
 * (bf16 hi[/lo] K-major UMMA core-matrix panels in layer order + fp32 bias/head block).  Must be
 * re-run whenever the parameters change (after optimizer.step()). */
size_t plnerf_packed_bytes(const plnerf_net_desc* desc, int precision);
int plnerf_pack_weights(const plnerf_net_desc* desc, const plnerf_net_params* params, int precision,
                        void* packed, void* stream);

/* ---- a4+a5+a7: network_query_fn == run_network (run_plnerf.py:78-92) --------------------------
 * pts = rays_o + rays_d * z (never materialised), PE, MLP.  rays [n, stride] with cols 0-2 origin,
 * 3-5 direction, last 3 = unit viewdir when desc->use_viewdirs.  z [n,S] -> raw [n,S,4].
 * ws: workspace of plnerf_query_workspace_bytes(desc, n) bytes. */
size_t plnerf_query_workspace_bytes(const plnerf_net_desc* desc, int64_t n_rays);
int plnerf_network_query(const plnerf_net_desc* desc, const void* packed, int precision,
                         int multires, int multires_views, const float* rays, int64_t n, int stride,
                         const float* z, int S, float* raw, void* ws, size_t ws_bytes, void* stream);

/* ---- training: what loss.backward() replays through run_network / NeRF.forward (autograd) --------
 * plnerf_network_query_train = plnerf_network_query (bf16) that additionally stashes every layer's
 * bf16 activations + ReLU masks (stash: plnerf_train_stash_bytes(desc, n, S) bytes, 1 KiB aligned).
 * plnerf_network_query_bwd consumes that stash and g_raw = dL/draw [n*S, g_stride>=4] and ADDS the
 * parameter gradients into `grads` (fp32, reference state_dict layout).  packed_bwd: transposed weight
 * stream from plnerf_pack_weights_bwd (plnerf_packed_bwd_bytes).  Networks with view directions
 * (raw [n,S,4]) and without (output_linear head, raw [n,S,output_ch], 4 <= output_ch <= 8; channels past the fourth get no
 * gradient, like the unused views_linears); the sample positions carry no gradient (run_plnerf.py:728), so no input gradient is produced. */
size_t plnerf_train_stash_bytes(const plnerf_net_desc* desc, int64_t n_rays, int S);
int plnerf_network_query_train(const plnerf_net_desc* desc, const void* packed, int multires,
                               int multires_views, const float* rays, int64_t n, int stride,
                               const float* z, int S, float* raw, void* stash, size_t stash_bytes,
                               void* ws, size_t ws_bytes, void* stream);
size_t plnerf_packed_bwd_bytes(const plnerf_net_desc* desc);
int plnerf_pack_weights_bwd(const plnerf_net_desc* desc, const plnerf_net_params* params,
                            void* packed_bwd, void* stream);
/* The repack a training loop needs after optimizer.step() (run_plnerf.py:1302-1303 changes every parameter): for each of
 * n_nets (1 or 2) networks, plnerf_pack_weights(bf16) into packed[i] and plnerf_pack_weights_bwd into packed_bwd[i], all in
 * ONE kernel launch. */
int plnerf_pack_weights_train(int n_nets, const plnerf_net_desc* const* descs,
                              const plnerf_net_params* const* params, void* const* packed,
                              void* const* packed_bwd, void* stream);
int plnerf_network_query_bwd(const plnerf_net_desc* desc, const void* packed, const void* packed_bwd,
                             int64_t n, int S, const float* g_raw, int g_stride, void* stash,
                             size_t stash_bytes, const plnerf_net_grads* grads, void* stream);

/* ---- a7: NeRF.forward (run_nerf_helpers.py:105-128) on already-embedded rows ------------------
 * x [m, input_ch + input_ch_views] -> out [m,4] (use_viewdirs) or [m,output_ch].
 * ws: plnerf_query_workspace_bytes(desc, m). */
int plnerf_mlp_forward(const plnerf_net_desc* desc, const void* packed, int precision,
                       const float* x, int64_t m, float* out, void* ws, size_t ws_bytes, void* stream);

/* ---- a8/a9/a10: raw2outputs + compute_weights{,_piecewise_linear} (run_plnerf.py:504-624) -----
 * raw [n,S,raw_stride>=4], z [n,S], rays [n,stride] (dir cols 3-5, near/far cols 6,7).
 * noise [n,S] additive density noise already scaled by raw_noise_std, or NULL.
 * Optional outputs: weights [n,S+1] (linear) / [n,S] (constant); tau,T [n,S+2] (linear only). */
int plnerf_raw2outputs(const float* raw, int raw_stride, const float* z, const float* rays,
                       int64_t n, int stride, int S, int mode, int color_mode, int white_bkgd,
                       int farcolorfix, const float* noise, float* rgb_map, float* disp_map,
                       float* acc_map, float* depth_map, float* weights, float* tau, float* T,
                       void* stream);

/* ---- backward of a8/a9/a10 (what autograd computes through raw2outputs, run_plnerf.py:553-624) ----
 * Upstream gradients of rgb_map [n,3], depth_map/acc_map/disp_map [n] (any may be NULL) ->
 * g_raw [n,S,raw_stride] (channels 0-3; further channels are zeroed).  z_vals carry no gradient
 * (the reference detaches the importance samples, run_plnerf.py:728). */
int plnerf_raw2outputs_bwd(const float* raw, int raw_stride, const float* z, const float* rays,
                           int64_t n, int stride, int S, int mode, int color_mode, int white_bkgd,
                           int farcolorfix, const float* noise, const float* g_rgb_map,
                           const float* g_depth_map, const float* g_acc_map, const float* g_disp_map,
                           float* g_raw, void* stream);

/* ---- a11: sample_pdf_reformulation (+pw_linear_sample_*) (run_nerf_helpers.py:340-445) --------
 * z [n,S], weights [n,S+1], tau,T [n,S+2], rays (near/far cols 6,7), u [n,Ni] in [0,1) or NULL
 * (Philox).  -> samples [n,Ni] (unclamped, unsorted), inds [n,Ni] int64 or NULL. */
int plnerf_sample_pdf_pl(const float* z, const float* weights, const float* tau, const float* T,
                         const float* rays, int64_t n, int stride, int S, int Ni, const float* u,
                         uint64_t seed, uint64_t ray_id_offset, float zero_tol, float epsilon,
                         float* samples, int64_t* inds, void* stream);

/* ---- a12: sample_pdf (run_nerf_helpers.py:241-284) ---------------------------------------------
 * bins [n,nb], weights [n,nb-1] (raw, the +1e-5 and normalisation happen inside). */
int plnerf_sample_pdf(const float* bins, const float* weights, int64_t n, int nb, int Ni,
                      const float* u, uint64_t seed, uint64_t ray_id_offset, float* samples,
                      int64_t* inds, void* stream);

/* ---- f-4 (next row): the depth-experiment sampler variants, forward ------------------------------
 * sample_pdf_reformulation_return_u (run_nerf_helpers.py:448-533) and sample_pdf_return_u (:286-337): the same
 * inverse-CDF samplers with `load_u` (explicit u, or NULL = Philox draws) that additionally return what the depth
 * losses consume: T, tau and the knot at the lower bracket index (`T_below`, `tau_below`, `bin_below`, each [n,Ni])
 * and the u that was used (`u_out` [n,Ni]).  Any extra output may be NULL. */
int plnerf_sample_pdf_pl_return_u(const float* z, const float* weights, const float* tau, const float* T,
                                  const float* rays, int64_t n, int stride, int S, int Ni, const float* load_u,
                                  uint64_t seed, uint64_t ray_id_offset, float zero_tol, float epsilon,
                                  float* samples, float* T_below, float* tau_below, float* bin_below, float* u_out,
                                  int64_t* inds, void* stream);
int plnerf_sample_pdf_return_u(const float* bins, const float* weights, int64_t n, int nb, int Ni,
                               const float* load_u, uint64_t seed, uint64_t ray_id_offset, float* samples,
                               float* u_out, int64_t* inds, void* stream);

/* ---- f-4, differentiable form: what autograd computes through the two return_u samplers (the depth experiments,
 * depth_supervised_exps/run_nerf_sample_based_depth.py:881-932, back-propagate through the samples).  The bracket comes from
 * searchsorted (no gradient); each sample sends gradient to the two knots of its bracket.  u = the forward's draws (u_out or
 * load_u).  Cotangents g_samples / g_T_below / g_tau_below / g_bin_below [n,Ni] (NULL = zero).  Outputs are WRITTEN, any may
 * be NULL: g_z [n,S], g_near [n], g_far [n] (the knots s = [near, z, far]), g_tau [n,S+2], g_T [n,S+2]; the weights get no
 * gradient in the piecewise-linear variant (they only enter the cdf).  Rules at kinks follow torch: max(eps, x) splits
 * ties, clamp(t, eps, ds) sends the gradient to the bound it returns, a NaN sample's gradient goes to s_left.
 * plnerf_sample_pdf_return_u_bwd: g_bins [n,nb] and g_weights [n,nb-1] (through pdf = (w + 1e-5) / sum and its cumsum). */
int plnerf_sample_pdf_pl_return_u_bwd(const float* z, const float* weights, const float* tau, const float* T, const float* rays,
                                      int64_t n, int stride, int S, int Ni, const float* u, float zero_tol, float epsilon,
                                      const float* g_samples, const float* g_T_below, const float* g_tau_below,
                                      const float* g_bin_below, float* g_z, float* g_near, float* g_far, float* g_tau, float* g_T,
                                      void* stream);
int plnerf_sample_pdf_return_u_bwd(const float* bins, const float* weights, int64_t n, int nb, int Ni, const float* u,
                                   const float* g_samples, float* g_bins, float* g_weights, void* stream);

/* ---- a13: clamp + sort-merge + z_std (run_plnerf.py:728-734, :752) ----------------------------
 * z [n,S] ascending, samples [n,Ni] -> z_out [n,S+Ni] ascending; z_std [n] or NULL. */
int plnerf_merge_samples(const float* z, const float* samples, const float* rays, int64_t n,
                         int stride, int S, int Ni, float* z_out, float* z_std, void* stream);

/* ---- a3: render_rays (run_plnerf.py:627-758), forward ------------------------------------------
 * rays [n, stride] = [o(3), d(3), near, far, (viewdir(3))].  Explicit draws (any may be NULL):
 * t_rand [n,Ns], u [n,Ni], noise0 [n,Ns], noise1 [n,Ns+Ni] (already scaled).
 * fine_desc/fine_packed NULL -> the coarse network is used for the fine pass (network_fine=None).
 */
size_t plnerf_render_workspace_bytes(const plnerf_render_cfg* cfg, const plnerf_net_desc* desc,
                                     int64_t n_rays);
int plnerf_render_rays_fwd(const plnerf_render_cfg* cfg, const plnerf_net_desc* coarse_desc,
                           const void* coarse_packed, const plnerf_net_desc* fine_desc,
                           const void* fine_packed, const float* rays, int64_t n, int stride,
                           const float* t_rand, const float* u, const float* noise0,
                           const float* noise1, const plnerf_render_out* out, void* ws,
                           size_t ws_bytes, void* stream);

/* ---- a3, training: render_rays forward that keeps what loss.backward() needs, and its backward -----------------
 * The reference's contract (run_plnerf.py:1283-1300): render(..., retraw=True) -> loss = img2mse(rgb) + img2mse(rgb0)
 * -> loss.backward() reaches both networks' parameters; the importance samples carry no gradient (:728).
 * plnerf_render_rays_fwd_train == plnerf_render_rays_fwd with the fused MLP in stash mode; everything the backward
 * reads (depths, raw of both passes, the two stashes) stays in `ws` (plnerf_render_train_workspace_bytes, 1 KiB
 * aligned), which the caller keeps untouched until plnerf_render_rays_bwd has been enqueued on the same stream.
 * plnerf_render_rays_bwd: upstream gradients of the eight maps (any may be NULL) -> parameter gradients ADDED into
 * grads_coarse / grads_fine (fp32, state_dict layout; fine_* NULL = the coarse network served the fine pass).
 * `*_packed_bwd` = plnerf_pack_weights_bwd.  noise0 / noise1: the same explicit arrays as in the forward (or NULL: the
 * forward's Philox draws are regenerated from cfg->seed).  bf16 precision; both networks with or both without view directions.
 * With a fine pass, plnerf_render_rays_bwd enqueues the coarse pass's backward on a side stream the library owns (one per
 * device and host thread), forked from `stream` and joined back into it before the call returns: the two passes are
 * independent (:728) and their kernels fill each other's idle SMs.  The caller sees ordinary stream order. */
typedef struct plnerf_render_grads {
  const float *g_rgb_map, *g_disp_map, *g_acc_map, *g_depth_map;   /* [n,3], [n], [n], [n] */
  const float *g_rgb0, *g_disp0, *g_acc0, *g_depth0;               /* coarse maps, read when N_importance > 0 */
} plnerf_render_grads;
size_t plnerf_render_train_workspace_bytes(const plnerf_render_cfg* cfg, const plnerf_net_desc* coarse_desc,
                                           const plnerf_net_desc* fine_desc, int64_t n_rays);
int plnerf_render_rays_fwd_train(const plnerf_render_cfg* cfg, const plnerf_net_desc* coarse_desc,
                                 const void* coarse_packed, const plnerf_net_desc* fine_desc, const void* fine_packed,
                                 const float* rays, int64_t n, int stride, const float* t_rand, const float* u,
                                 const float* noise0, const float* noise1, const plnerf_render_out* out, void* ws,
                                 size_t ws_bytes, void* stream);
int plnerf_render_rays_bwd(const plnerf_render_cfg* cfg, const plnerf_net_desc* coarse_desc, const void* coarse_packed,
                           const void* coarse_packed_bwd, const plnerf_net_desc* fine_desc, const void* fine_packed,
                           const void* fine_packed_bwd, const float* rays, int64_t n, int stride, const float* noise0,
                           const float* noise1, const plnerf_render_grads* g, const plnerf_net_grads* grads_coarse,
                           const plnerf_net_grads* grads_fine, void* ws, size_t ws_bytes, void* stream);

/* ---- f-2: the loss and optimiser steps of the training loop (run_plnerf.py:1289-1315) ------------------------------
 * plnerf_mse_loss_grad: img_loss = img2mse(rgb, target_s), img_loss0 = img2mse(rgb0, target_s) (run_nerf_helpers.py:17,
 * run_plnerf.py:1289-1297) and the start of loss.backward(): g_rgb = scale (rgb - t), g_rgb0 = scale (rgb0 - t) with
 * scale = 2 / (3 B) for the mean over a global batch of B rays; the sums of squared errors are ADDED to sqerr[0] (fine) and
 * sqerr[1] (coarse) in a fixed order (reproducible).  The target row of ray i is target[pix[i]] (the gather :1280) or
 * target[i] when pix is NULL.  rgb0 / g_rgb0 may be NULL (N_importance == 0).
 * plnerf_adam_step: torch.optim.Adam(betas=(beta1, beta2), eps, amsgrad=False, weight_decay=0) (:431-447, :1302-1303) on
 * one flat fp32 segment (parameters, gradients and both moments contiguous, same length); step = the 1-based count of this
 * update; zero_grads != 0 also clears the gradients (the next iteration's optimizer.zero_grad(), :1286). */
int plnerf_mse_loss_grad(const float* rgb, const float* rgb0, const float* target, const int64_t* pix, int64_t n,
                         float scale, float* g_rgb, float* g_rgb0, float* sqerr, void* stream);
int plnerf_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                     double beta2, double eps, int64_t step, int zero_grads, void* stream);
/* plnerf_train_rays_mse: the device work of one iteration of the reference loop on a ray batch (run_plnerf.py:1283-1299:
 * render -> img2mse(rgb) + img2mse(rgb0) -> loss.backward()) in ONE call = plnerf_render_rays_fwd_train +
 * plnerf_mse_loss_grad + plnerf_render_rays_bwd with the same arguments (ws: plnerf_render_train_workspace_bytes), except
 * that the coarse pass's loss and backward are enqueued on an internal side stream as soon as the coarse maps exist and run
 * BESIDE the fine pass (the importance samples are detached, :728: the passes are independent); the side stream is joined
 * into `stream` before the call returns.  Gradients are ADDED into grads_*; the sums of squared errors into sqerr[0] (fine;
 * the only map when N_importance == 0) and sqerr[1] (coarse).  `out` may be NULL, and so may any of its members: maps the
 * caller does not ask for stay in the workspace. */
int plnerf_train_rays_mse(const plnerf_render_cfg* cfg, const plnerf_net_desc* coarse_desc, const void* coarse_packed,
                          const void* coarse_packed_bwd, const plnerf_net_desc* fine_desc, const void* fine_packed,
                          const void* fine_packed_bwd, const float* rays, int64_t n, int stride, const float* t_rand,
                          const float* u, const float* noise0, const float* noise1, const float* target,
                          const int64_t* pix, float scale, float* sqerr, const plnerf_render_out* out,
                          const plnerf_net_grads* grads_coarse, const plnerf_net_grads* grads_fine, void* ws,
                          size_t ws_bytes, void* stream);

/* ---- measurement hooks (bench.py): time every fused-MLP launch with CUDA events on its own stream --
 * plnerf_profile_enable(1) starts recording (and clears old records); plnerf_profile_read
 * synchronises the recorded events and returns the summed device time, launch count and rows. */
int plnerf_profile_enable(int on);
int plnerf_profile_read(double* mlp_ms_sum, int64_t* mlp_launches, int64_t* mlp_rows);

#ifdef __cplusplus
}
#endif
#endif /* PLNERF_B200_H */
