// Internal launch wrappers shared by api.cu / render.cu (device pointers, explicit stream).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/plnerf_b200.h"

namespace plnerf {

int launch_encode(const float* x, int64_t n, int L, float* out, cudaStream_t st);
int launch_stratified_z(const float* rays, int64_t n, int stride, int Ns, int lindisp, int perturb,
                        const float* t_rand, uint64_t seed, uint64_t ray0, float* z, cudaStream_t st);
int launch_composite(const float* raw, int raw_stride, const float* z, const float* rays, int64_t n,
                     int stride, int S, int mode, int color_mode, int white_bkgd, int farcolorfix,
                     const float* noise, float noise_std, uint64_t seed, uint64_t ray0, uint32_t noise_stream,
                     float* rgb_map, float* disp_map, float* acc_map, float* depth_map, float* weights,
                     float* tau, float* T, cudaStream_t st);
int launch_composite_bwd(const float* raw, int raw_stride, const float* z, const float* rays, int64_t n, int stride,
                         int S, int mode, int color_mode, int white_bkgd, int farcolorfix, const float* noise,
                         const float* g_rgb, const float* g_depth, const float* g_acc, const float* g_disp,
                         float* g_raw, cudaStream_t st, float noise_std = 0.f, uint64_t seed = 0, uint64_t ray0 = 0,
                         uint32_t noise_stream = 0);
int launch_sample_pl(const float* z, const float* w, const float* tau, const float* T, const float* rays,
                     int64_t n, int stride, int S, int Ni, const float* u, uint64_t seed, uint64_t ray0,
                     float zero_tol, float eps, float* samples, int64_t* inds, cudaStream_t st, float* T_below = nullptr,
                     float* tau_below = nullptr, float* bin_below = nullptr, float* u_out = nullptr);
int launch_sample_const(const float* bins, int bins_stride, int bins_mid, const float* w, int w_stride,
                        int64_t n, int nb, int Ni, const float* u, uint64_t seed, uint64_t ray0,
                        float* samples, int64_t* inds, cudaStream_t st, float* u_out = nullptr);
// f-4: gradients of the return_u sampler variants (the depth experiments differentiate through the samples)
int launch_sample_pl_bwd(const float* z, const float* w, const float* tau, const float* T, const float* rays, int64_t n,
                         int stride, int S, int Ni, const float* u, float zero_tol, float eps, const float* g_samples,
                         const float* g_T_below, const float* g_tau_below, const float* g_bin_below, float* g_z, float* g_near,
                         float* g_far, float* g_tau, float* g_T, cudaStream_t st);
int launch_sample_const_bwd(const float* bins, const float* w, int64_t n, int nb, int Ni, const float* u, const float* g_samples,
                            float* g_bins, float* g_w, cudaStream_t st);
int launch_merge(const float* z, const float* samples, const float* rays, int64_t n, int stride, int S,
                 int Ni, float* z_out, float* z_std, cudaStream_t st);
// importance sampling (linear: sample_pdf_reformulation; else sample_pdf on z_mid / weights[1:-1] of a constant-mode weight
// row [n,S]) + clamp + sort-merge + z_std in one kernel
int launch_sample_merge(int linear, const float* z, const float* w, const float* tau, const float* T, const float* rays, int64_t n,
                        int stride, int S, int Ni, const float* u, uint64_t seed, uint64_t ray0, float zero_tol, float eps,
                        float* z_out, float* z_std, int64_t* inds, cudaStream_t st);

int launch_pack_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int c2w_ld,
                     const float* c2w_static, int c2w_static_ld, const float* rays_o, const float* rays_d,
                     const int64_t* pix, int64_t n,
                     int ndc, float ndc_cx, float ndc_cy, float ndc_near, float near, float far, int use_viewdirs,
                     float* out, int stride, cudaStream_t st);

// f-2: loss gradient of the two MSE terms and the flat Adam step of the training loop
int launch_mse_loss_grad(const float* rgb, const float* rgb0, const float* target, const int64_t* pix, int64_t n, float scale,
                         float* g_rgb, float* g_rgb0, float* sqerr, cudaStream_t st);
int launch_adam_flat(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                     double beta2, double eps, int64_t step, int zero_grads, cudaStream_t st);

// ---- fused MLP (mlp_fwd.cu) ----------------------------------------------------------------
size_t mlp_packed_bytes(const plnerf_net_desc* d, int precision);
int mlp_pack(const plnerf_net_desc* d, const plnerf_net_params* p, int precision, void* packed, cudaStream_t st);
size_t mlp_workspace_bytes(const plnerf_net_desc* d, int64_t n_rays);
// Quadrature to run INSIDE the fused query (k_mlp3's compositor warps): the arguments of launch_composite.
struct FusedComposite {
  int mode, color_mode, white_bkgd, farcolorfix;
  const float* noise; float noise_std; uint64_t seed, ray0; uint32_t noise_stream;
  float *rgb_map, *disp_map, *acc_map, *depth_map, *weights, *tau, *T;
};
// Fused query: rows = n*S samples of rays (PE computed in-kernel).  With `fc`, the quadrature of the rays runs inside the
// kernel when the configuration allows it (bf16, use_viewdirs network, S <= 256, k_mlp3): *fused tells whether it did
// (otherwise the caller composites `raw` itself); `raw` is then written only if need_raw.
int mlp_query(const plnerf_net_desc* d, const void* packed, int precision, int multires, int multires_views,
              const float* rays, int64_t n, int stride, const float* z, int S, float* raw, int raw_stride,
              void* ws, size_t ws_bytes, cudaStream_t st, const FusedComposite* fc = nullptr, bool need_raw = true,
              bool* fused = nullptr, const float* viewbias_pre = nullptr);
// stratified depths + the per-ray view bias of both networks in one launch (returns 1 when the configuration is not covered)
int launch_ray_setup(const plnerf_net_desc* cd, const void* cpacked, const plnerf_net_desc* fd, const void* fpacked, int precision,
                     int multires_views, const float* rays, int64_t n, int stride, int Ns, int lindisp, int perturb,
                     const float* t_rand, uint64_t seed, uint64_t ray0, float* z, float* vb_c, float* vb_f, cudaStream_t st,
                     float* dirpe = nullptr);
// NeRF.forward on embedded rows x [m, input_ch + input_ch_views].
int mlp_forward_embedded(const plnerf_net_desc* d, const void* packed, int precision, const float* x, int64_t m,
                         float* out, void* ws, size_t ws_bytes, cudaStream_t st);
size_t mlp_train_stash_bytes(const plnerf_net_desc* d, int64_t n_rays, int S);
int mlp_query_train(const plnerf_net_desc* d, const void* packed, int multires, int multires_views, const float* rays,
                    int64_t n, int stride, const float* z, int S, float* raw, int raw_stride, void* stash,
                    size_t stash_bytes, void* ws, size_t ws_bytes, cudaStream_t st, const float* viewbias_pre = nullptr,
                    const float* dirpe_pre = nullptr);
size_t mlp_packed_bwd_bytes(const plnerf_net_desc* d);
int mlp_pack_bwd(const plnerf_net_desc* d, const plnerf_net_params* p, void* packed, cudaStream_t st);
// forward stream + tail + transposed stream of 1 or 2 networks in one launch (the training step's repack after Adam)
int mlp_pack_train(int n_nets, const plnerf_net_desc* const* descs, const plnerf_net_params* const* params, void* const* packed,
                   void* const* packed_bwd, cudaStream_t st);
int mlp_query_bwd(const plnerf_net_desc* d, const void* packed_fwd, const void* packed_bwd, int64_t n, int S,
                  const float* g_raw, int g_stride, void* stash, size_t stash_bytes, const plnerf_net_grads* g,
                  cudaStream_t st);
int profile_enable(int on);
int profile_read(double* ms_sum, int64_t* launches, int64_t* rows);

}  // namespace plnerf
