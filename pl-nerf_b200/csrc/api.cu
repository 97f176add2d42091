// C ABI of libplnerf_b200 (declared in include/plnerf_b200.h): argument validation, error plumbing
// and the render_rays orchestration (reference run_plnerf.py:627-758) on one CUDA stream.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "ops.cuh"

namespace plnerf {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return PLNERF_E_CUDA;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#ifdef PLNERF_DEBUG
int debug_umma_gemm(const float* A, const float* B, int N, int K, float* D, cudaStream_t st);
int debug_umma_gemm_ex(const float* A, const float* B, int N, int K, int a_mode, uint32_t lbo, uint32_t sbo, float* D, cudaStream_t st);
int debug_mma_rate(int mode, int iters, int grid, long long* cycles_out, cudaStream_t st);
int debug_set_trace(long long* buf);
int debug_umma_gemm_mn(const float* X, const float* Y, int N, int K, uint32_t lbo, uint32_t sbo, float* D, cudaStream_t st);
#endif

// workspace carving for render_rays
struct RenderWs {
  float *z0, *raw0, *w0, *tau0, *T0, *zs, *z1, *raw1, *vb_f;
  void* mlp_ws; size_t mlp_ws_bytes;
  size_t total;
};
static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
static RenderWs carve(const plnerf_render_cfg* c, const plnerf_net_desc* d, int64_t n, uint8_t* base) {
  RenderWs w;
  size_t off = 0;
  const int Ns = c->N_samples, Ni = c->N_importance, S1 = Ns + Ni;
  auto take = [&](size_t floats) { float* p = reinterpret_cast<float*>(base + off); off += align_up(floats * sizeof(float)); return p; };
  const int ch0 = 4 > (d->use_viewdirs ? 4 : d->output_ch) ? 4 : (d->use_viewdirs ? 4 : d->output_ch);
  w.z0 = take((size_t)n * Ns);
  w.raw0 = take((size_t)n * Ns * ch0);
  w.w0 = take((size_t)n * (Ns + 1));
  w.tau0 = take((size_t)n * (Ns + 2));
  w.T0 = take((size_t)n * (Ns + 2));
  w.zs = take((size_t)n * (Ni > 0 ? Ni : 1));
  w.z1 = take((size_t)n * S1);
  w.raw1 = take((size_t)n * S1 * ch0);
  w.vb_f = take(d->use_viewdirs ? (size_t)n * 128 : 0);      // the fine network's per-ray view bias (launch_ray_setup)
  w.mlp_ws = base + off;
  w.mlp_ws_bytes = mlp_workspace_bytes(d, n);
  off += align_up(w.mlp_ws_bytes);
  w.total = off;
  return w;
}

}  // namespace plnerf

using namespace plnerf;

extern "C" {

const char* plnerf_last_error(void) { return g_err; }
int plnerf_abi_version(void) { return PLNERF_ABI_VERSION; }
uint64_t plnerf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int plnerf_encode(const float* x, int64_t n, int multires, float* out, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (x && out)), "encode: null argument");
  PLNERF_CHECK_ARG(multires <= 16, "encode: multires too large");
  return launch_encode(x, n, multires, out, (cudaStream_t)stream);
}

int plnerf_stratified_z(const float* rays, int64_t n, int stride, int N_samples, int lindisp, int perturb,
                        const float* t_rand, uint64_t seed, uint64_t ray_id_offset, float* z_vals, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (rays && z_vals)), "stratified_z: null argument");
  PLNERF_CHECK_ARG(stride >= 8 && N_samples >= 1, "stratified_z: need stride >= 8 and N_samples >= 1");
  return launch_stratified_z(rays, n, stride, N_samples, lindisp, perturb, t_rand, seed, ray_id_offset, z_vals,
                             (cudaStream_t)stream);
}

size_t plnerf_packed_bytes(const plnerf_net_desc* desc, int precision) { return mlp_packed_bytes(desc, precision); }

int plnerf_pack_weights(const plnerf_net_desc* desc, const plnerf_net_params* params, int precision, void* packed,
                        void* stream) {
  return mlp_pack(desc, params, precision, packed, (cudaStream_t)stream);
}

size_t plnerf_query_workspace_bytes(const plnerf_net_desc* desc, int64_t n_rays) { return mlp_workspace_bytes(desc, n_rays); }

int plnerf_network_query(const plnerf_net_desc* desc, const void* packed, int precision, int multires,
                         int multires_views, const float* rays, int64_t n, int stride, const float* z, int S,
                         float* raw, void* ws, size_t ws_bytes, void* stream) {
  PLNERF_CHECK_ARG(desc, "network_query: null desc");
  return mlp_query(desc, packed, precision, multires, multires_views, rays, n, stride, z, S, raw,
                   desc->use_viewdirs ? 4 : desc->output_ch, ws, ws_bytes, (cudaStream_t)stream);
}

size_t plnerf_train_stash_bytes(const plnerf_net_desc* desc, int64_t n_rays, int S) { return mlp_train_stash_bytes(desc, n_rays, S); }

int plnerf_network_query_train(const plnerf_net_desc* desc, const void* packed, int multires, int multires_views,
                               const float* rays, int64_t n, int stride, const float* z, int S, float* raw, void* stash,
                               size_t stash_bytes, void* ws, size_t ws_bytes, void* stream) {
  PLNERF_CHECK_ARG(desc, "network_query_train: null desc");
  return mlp_query_train(desc, packed, multires, multires_views, rays, n, stride, z, S, raw, desc->use_viewdirs ? 4 : desc->output_ch, stash, stash_bytes, ws,
                         ws_bytes, (cudaStream_t)stream);
}

size_t plnerf_packed_bwd_bytes(const plnerf_net_desc* desc) { return mlp_packed_bwd_bytes(desc); }

int plnerf_pack_weights_bwd(const plnerf_net_desc* desc, const plnerf_net_params* params, void* packed_bwd, void* stream) {
  return mlp_pack_bwd(desc, params, packed_bwd, (cudaStream_t)stream);
}

int plnerf_pack_weights_train(int n_nets, const plnerf_net_desc* const* descs, const plnerf_net_params* const* params,
                              void* const* packed, void* const* packed_bwd, void* stream) {
  return mlp_pack_train(n_nets, descs, params, packed, packed_bwd, (cudaStream_t)stream);
}

int plnerf_network_query_bwd(const plnerf_net_desc* desc, const void* packed, const void* packed_bwd, int64_t n, int S,
                             const float* g_raw, int g_stride, void* stash, size_t stash_bytes,
                             const plnerf_net_grads* grads, void* stream) {
  return mlp_query_bwd(desc, packed, packed_bwd, n, S, g_raw, g_stride, stash, stash_bytes, grads, (cudaStream_t)stream);
}

int plnerf_mlp_forward(const plnerf_net_desc* desc, const void* packed, int precision, const float* x, int64_t m,
                       float* out, void* ws, size_t ws_bytes, void* stream) {
  return mlp_forward_embedded(desc, packed, precision, x, m, out, ws, ws_bytes, (cudaStream_t)stream);
}

int plnerf_raw2outputs(const float* raw, int raw_stride, const float* z, const float* rays, int64_t n, int stride,
                       int S, int mode, int color_mode, int white_bkgd, int farcolorfix, const float* noise,
                       float* rgb_map, float* disp_map, float* acc_map, float* depth_map, float* weights, float* tau,
                       float* T, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (raw && z && rays)), "raw2outputs: null argument");
  PLNERF_CHECK_ARG(raw_stride >= 4 && stride >= 8 && S >= 1, "raw2outputs: need raw_stride >= 4, stride >= 8, S >= 1");
  PLNERF_CHECK_ARG(mode == PLNERF_MODE_LINEAR || mode == PLNERF_MODE_CONSTANT, "raw2outputs: bad mode %d", mode);
  PLNERF_CHECK_ARG(color_mode == PLNERF_COLOR_MIDPOINT || color_mode == PLNERF_COLOR_LEFT, "raw2outputs: bad color_mode %d", color_mode);
  return launch_composite(raw, raw_stride, z, rays, n, stride, S, mode, color_mode, white_bkgd, farcolorfix, noise,
                          0.f, 0, 0, 0, rgb_map, disp_map, acc_map, depth_map, weights, tau, T, (cudaStream_t)stream);
}

int plnerf_raw2outputs_bwd(const float* raw, int raw_stride, const float* z, const float* rays, int64_t n, int stride,
                           int S, int mode, int color_mode, int white_bkgd, int farcolorfix, const float* noise,
                           const float* g_rgb_map, const float* g_depth_map, const float* g_acc_map,
                           const float* g_disp_map, float* g_raw, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (raw && z && rays && g_raw)), "raw2outputs_bwd: null argument");
  PLNERF_CHECK_ARG(raw_stride >= 4 && stride >= 8 && S >= 1, "raw2outputs_bwd: need raw_stride >= 4, stride >= 8, S >= 1");
  PLNERF_CHECK_ARG(mode == PLNERF_MODE_LINEAR || mode == PLNERF_MODE_CONSTANT, "raw2outputs_bwd: bad mode %d", mode);
  return launch_composite_bwd(raw, raw_stride, z, rays, n, stride, S, mode, color_mode, white_bkgd, farcolorfix, noise,
                              g_rgb_map, g_depth_map, g_acc_map, g_disp_map, g_raw, (cudaStream_t)stream);
}

int plnerf_sample_pdf_pl(const float* z, const float* weights, const float* tau, const float* T, const float* rays,
                         int64_t n, int stride, int S, int Ni, const float* u, uint64_t seed, uint64_t ray_id_offset,
                         float zero_tol, float epsilon, float* samples, int64_t* inds, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (z && weights && tau && T && rays && samples)), "sample_pdf_pl: null argument");
  PLNERF_CHECK_ARG(stride >= 8 && S >= 1 && Ni >= 0, "sample_pdf_pl: bad sizes");
  return launch_sample_pl(z, weights, tau, T, rays, n, stride, S, Ni, u, seed, ray_id_offset, zero_tol, epsilon,
                          samples, inds, (cudaStream_t)stream);
}

int plnerf_sample_pdf(const float* bins, const float* weights, int64_t n, int nb, int Ni, const float* u,
                      uint64_t seed, uint64_t ray_id_offset, float* samples, int64_t* inds, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (bins && weights && samples)), "sample_pdf: null argument");
  return launch_sample_const(bins, nb, 0, weights, nb - 1, n, nb, Ni, u, seed, ray_id_offset, samples, inds,
                             (cudaStream_t)stream);
}

int plnerf_sample_pdf_pl_return_u(const float* z, const float* weights, const float* tau, const float* T, const float* rays,
                                  int64_t n, int stride, int S, int Ni, const float* load_u, uint64_t seed,
                                  uint64_t ray_id_offset, float zero_tol, float epsilon, float* samples, float* T_below,
                                  float* tau_below, float* bin_below, float* u_out, int64_t* inds, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (z && weights && tau && T && rays && samples)), "sample_pdf_pl_return_u: null argument");
  PLNERF_CHECK_ARG(stride >= 8 && S >= 1 && Ni >= 0, "sample_pdf_pl_return_u: bad sizes");
  return launch_sample_pl(z, weights, tau, T, rays, n, stride, S, Ni, load_u, seed, ray_id_offset, zero_tol, epsilon,
                          samples, inds, (cudaStream_t)stream, T_below, tau_below, bin_below, u_out);
}

int plnerf_sample_pdf_return_u(const float* bins, const float* weights, int64_t n, int nb, int Ni, const float* load_u,
                               uint64_t seed, uint64_t ray_id_offset, float* samples, float* u_out, int64_t* inds,
                               void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (bins && weights && samples)), "sample_pdf_return_u: null argument");
  return launch_sample_const(bins, nb, 0, weights, nb - 1, n, nb, Ni, load_u, seed, ray_id_offset, samples, inds,
                             (cudaStream_t)stream, u_out);
}

int plnerf_sample_pdf_pl_return_u_bwd(const float* z, const float* weights, const float* tau, const float* T, const float* rays,
                                      int64_t n, int stride, int S, int Ni, const float* u, float zero_tol, float epsilon,
                                      const float* g_samples, const float* g_T_below, const float* g_tau_below,
                                      const float* g_bin_below, float* g_z, float* g_near, float* g_far, float* g_tau, float* g_T,
                                      void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (z && weights && tau && T && rays && u)), "sample_pdf_pl_return_u_bwd: null argument");
  PLNERF_CHECK_ARG(stride >= 8 && S >= 1 && Ni >= 0, "sample_pdf_pl_return_u_bwd: bad sizes");
  return launch_sample_pl_bwd(z, weights, tau, T, rays, n, stride, S, Ni, u, zero_tol, epsilon, g_samples, g_T_below, g_tau_below,
                              g_bin_below, g_z, g_near, g_far, g_tau, g_T, (cudaStream_t)stream);
}

int plnerf_sample_pdf_return_u_bwd(const float* bins, const float* weights, int64_t n, int nb, int Ni, const float* u,
                                   const float* g_samples, float* g_bins, float* g_weights, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (bins && weights && u && g_samples)), "sample_pdf_return_u_bwd: null argument");
  return launch_sample_const_bwd(bins, weights, n, nb, Ni, u, g_samples, g_bins, g_weights, (cudaStream_t)stream);
}

int plnerf_merge_samples(const float* z, const float* samples, const float* rays, int64_t n, int stride, int S,
                         int Ni, float* z_out, float* z_std, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (z && samples && rays && z_out)), "merge_samples: null argument");
  PLNERF_CHECK_ARG(stride >= 8 && S >= 1 && Ni >= 1, "merge_samples: bad sizes");
  return launch_merge(z, samples, rays, n, stride, S, Ni, z_out, z_std, (cudaStream_t)stream);
}

size_t plnerf_render_workspace_bytes(const plnerf_render_cfg* cfg, const plnerf_net_desc* desc, int64_t n_rays) {
  if (!cfg || !desc || n_rays < 0) return 0;
  return carve(cfg, desc, n_rays, nullptr).total;
}

int plnerf_render_rays_fwd(const plnerf_render_cfg* cfg, const plnerf_net_desc* cdesc, const void* cpacked,
                           const plnerf_net_desc* fdesc, const void* fpacked, const float* rays, int64_t n, int stride,
                           const float* t_rand, const float* u, const float* noise0, const float* noise1,
                           const plnerf_render_out* out, void* ws, size_t ws_bytes, void* stream) {
  PLNERF_CHECK_ARG(cfg && cdesc && cpacked && out, "render_rays: null argument");
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || rays), "render_rays: null rays");
  PLNERF_CHECK_ARG(stride >= 8, "render_rays: ray stride must be >= 8");
  PLNERF_CHECK_ARG(cfg->N_samples >= 2 && cfg->N_importance >= 0, "render_rays: need N_samples >= 2, N_importance >= 0");
  PLNERF_CHECK_ARG(cfg->mode == PLNERF_MODE_LINEAR || cfg->mode == PLNERF_MODE_CONSTANT, "render_rays: bad mode");
  PLNERF_CHECK_ARG(out->rgb_map && out->disp_map && out->acc_map && out->depth_map, "render_rays: rgb/disp/acc/depth outputs are required");
  if (n == 0) return PLNERF_OK;
  if (!fdesc || !fpacked) { fdesc = cdesc; fpacked = cpacked; }
  PLNERF_CHECK_ARG(fdesc->use_viewdirs == cdesc->use_viewdirs, "render_rays: coarse/fine use_viewdirs differ");
  const size_t need = carve(cfg, cdesc, n, nullptr).total;
  if (!ws || ws_bytes < need) { set_error("render_rays: workspace too small: need %zu bytes, got %zu", need, ws_bytes); return PLNERF_E_WORKSPACE; }
  PLNERF_CHECK_ARG(((uintptr_t)ws & 255) == 0, "render_rays: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  RenderWs w = carve(cfg, cdesc, n, static_cast<uint8_t*>(ws));
  const int Ns = cfg->N_samples, Ni = cfg->N_importance, S1 = Ns + Ni;
  const int chc = cdesc->use_viewdirs ? 4 : cdesc->output_ch, chf = fdesc->use_viewdirs ? 4 : fdesc->output_ch;
  const bool fine = Ni > 0;
  int rc;
  // (1) stratified depths (run_plnerf.py:683-705) + the per-ray view bias of both networks: one launch where covered
  float* z0 = (!fine && out->z_vals) ? out->z_vals : w.z0;
  const bool two_nets = fine && fpacked != cpacked;
  float* vb_c = static_cast<float*>(w.mlp_ws);
  rc = launch_ray_setup(cdesc, cpacked, two_nets ? fdesc : nullptr, fpacked, cfg->precision, cfg->multires_views, rays, n, stride, Ns,
                        cfg->lindisp, cfg->perturb, t_rand, cfg->seed, cfg->ray_id_offset, z0, vb_c, w.vb_f, st);
  if (rc < 0) return rc;
  const bool have_vb = (rc == 0);
  if (!have_vb) {
    rc = launch_stratified_z(rays, n, stride, Ns, cfg->lindisp, cfg->perturb, t_rand, cfg->seed, cfg->ray_id_offset, z0, st);
    if (rc) return rc;
  }
  // (2) coarse network query (:714)
  // (2)+(3) coarse network query (:714) with the quadrature (:715) fused into the kernel where the configuration allows
  float* raw0 = (!fine && out->raw) ? out->raw : w.raw0;
  const float std0 = (!noise0) ? cfg->raw_noise_std : 0.f;
  const FusedComposite fc0{cfg->mode, cfg->color_mode, cfg->white_bkgd, cfg->farcolorfix, noise0, std0, cfg->seed,
                           cfg->ray_id_offset, RNG_STREAM_NOISE0, fine ? out->rgb0 : out->rgb_map,
                           fine ? out->disp0 : out->disp_map, fine ? out->acc0 : out->acc_map,
                           fine ? out->depth0 : out->depth_map, fine ? w.w0 : nullptr, fine ? w.tau0 : nullptr,
                           fine ? w.T0 : nullptr};
  bool fused0 = false;
  rc = mlp_query(cdesc, cpacked, cfg->precision, cfg->multires, cfg->multires_views, rays, n, stride, z0, Ns, raw0, chc,
                 w.mlp_ws, w.mlp_ws_bytes, st, &fc0, /*need_raw=*/(!fine && out->raw != nullptr), &fused0, have_vb ? vb_c : nullptr);
  if (rc) return rc;
  if (!fused0) {
    rc = launch_composite(raw0, chc, z0, rays, n, stride, Ns, cfg->mode, cfg->color_mode, cfg->white_bkgd, cfg->farcolorfix,
                          noise0, std0, cfg->seed, cfg->ray_id_offset, RNG_STREAM_NOISE0, fc0.rgb_map, fc0.disp_map,
                          fc0.acc_map, fc0.depth_map, fc0.weights, fc0.tau, fc0.T, st);
    if (rc) return rc;
  }
  if (!fine) return PLNERF_OK;
  // (4)+(5) importance sampling (:721-726), detach / clamp / sort-merge / z_std (:728-734, :752): one kernel
  float* z1 = out->z_vals ? out->z_vals : w.z1;
  rc = launch_sample_merge(cfg->mode == PLNERF_MODE_LINEAR, z0, w.w0, w.tau0, w.T0, rays, n, stride, Ns, Ni, u, cfg->seed,
                           cfg->ray_id_offset, cfg->zero_tol, cfg->epsilon, z1, out->z_std, out->inds, st);
  if (rc) return rc;
  // (6) fine network query (:737-739) and quadrature (:741)
  float* raw1 = out->raw ? out->raw : w.raw1;
  const float std1 = (!noise1) ? cfg->raw_noise_std : 0.f;
  const FusedComposite fc1{cfg->mode, cfg->color_mode, cfg->white_bkgd, cfg->farcolorfix, noise1, std1, cfg->seed,
                           cfg->ray_id_offset, RNG_STREAM_NOISE1, out->rgb_map, out->disp_map, out->acc_map, out->depth_map,
                           nullptr, nullptr, nullptr};
  bool fused1 = false;
  rc = mlp_query(fdesc, fpacked, cfg->precision, cfg->multires, cfg->multires_views, rays, n, stride, z1, S1, raw1, chf,
                 w.mlp_ws, w.mlp_ws_bytes, st, &fc1, /*need_raw=*/out->raw != nullptr, &fused1,
                 have_vb ? (two_nets ? w.vb_f : vb_c) : nullptr);
  if (rc) return rc;
  if (!fused1)
    rc = launch_composite(raw1, chf, z1, rays, n, stride, S1, cfg->mode, cfg->color_mode, cfg->white_bkgd, cfg->farcolorfix,
                          noise1, std1, cfg->seed, cfg->ray_id_offset, RNG_STREAM_NOISE1, out->rgb_map, out->disp_map,
                          out->acc_map, out->depth_map, nullptr, nullptr, nullptr, st);
  return rc;
}

// One non-blocking side stream + fork / join events per device and host thread (created on first use, kept for the life of
// the thread): concurrent callers on different host threads never share fork / join events.
namespace plnerf {
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; bool ok; };
static SideStream* side_stream() {
  static thread_local SideStream tab[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("side_stream: bad device"); return nullptr; }
  SideStream& s = tab[dev];
  if (!s.ok) {
    cudaError_t e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming);
    if (e != cudaSuccess) { cuda_fail(e, "side_stream"); return nullptr; }
    s.ok = true;
  }
  return &s;
}
}  // namespace plnerf

// ---- training: forward with stash + the whole backward of a ray batch, one ABI call each --------------------------
namespace plnerf {
struct TrainWs {
  float *z0, *raw0, *w0, *tau0, *T0, *zs, *z1, *raw1, *graw, *graw0, *vb_f, *dirpe, *g_rgb, *g_rgb0, *maps;
  void* mlp_ws; size_t mlp_ws_bytes;
  uint8_t *stash0, *stash1; size_t stash0_bytes, stash1_bytes;
  size_t total;
};
static TrainWs carve_train(const plnerf_render_cfg* c, const plnerf_net_desc* cd, const plnerf_net_desc* fd, int64_t n, uint8_t* base) {
  TrainWs w;
  size_t off = 0;
  const int Ns = c->N_samples, Ni = c->N_importance, S1 = Ns + Ni;
  const int chc = cd->use_viewdirs ? 4 : cd->output_ch, chf = fd->use_viewdirs ? 4 : fd->output_ch;   // raw channels per sample
  auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  auto take = [&](size_t floats) { float* p = reinterpret_cast<float*>(base + off); off += up(floats * sizeof(float)); return p; };
  w.z0 = take((size_t)n * Ns);
  w.raw0 = take((size_t)n * Ns * chc);
  w.w0 = take((size_t)n * (Ns + 1));
  w.tau0 = take((size_t)n * (Ns + 2));
  w.T0 = take((size_t)n * (Ns + 2));
  w.zs = take((size_t)n * (Ni > 0 ? Ni : 1));
  w.z1 = take((size_t)n * S1);
  w.raw1 = take((size_t)n * S1 * chf);
  w.graw = take((size_t)n * S1 * (chc > chf ? chc : chf));     // k_composite_bwd writes d raw in raw's own layout
  w.graw0 = take(Ni > 0 ? (size_t)n * Ns * chc : 1);          // the coarse pass's own d raw: its backward runs beside the fine one
  w.vb_f = take((size_t)n * 128);
  w.dirpe = take((size_t)n * 32);
  w.g_rgb = take((size_t)n * 3);                               // plnerf_train_rays_mse: the loss gradients of the two rgb maps
  w.g_rgb0 = take((size_t)n * 3);
  w.maps = take((size_t)n * 13);                               // ... and the maps themselves when the caller does not ask for them
  w.mlp_ws = base + off;
  w.mlp_ws_bytes = mlp_workspace_bytes(cd, n);
  off += up(w.mlp_ws_bytes);
  w.stash0 = base + off; w.stash0_bytes = mlp_train_stash_bytes(cd, n, Ns); off += up(w.stash0_bytes);
  w.stash1 = base + off; w.stash1_bytes = Ni > 0 ? mlp_train_stash_bytes(fd, n, S1) : 0; off += up(w.stash1_bytes);
  w.total = off;
  return w;
}
}  // namespace plnerf

size_t plnerf_render_train_workspace_bytes(const plnerf_render_cfg* cfg, const plnerf_net_desc* coarse_desc,
                                           const plnerf_net_desc* fine_desc, int64_t n_rays) {
  if (!cfg || !coarse_desc || n_rays < 0) return 0;
  if (!fine_desc) fine_desc = coarse_desc;
  if (mlp_train_stash_bytes(coarse_desc, 1, 1) == 0) return 0;     // unsupported network (error text is set)
  return carve_train(cfg, coarse_desc, fine_desc, n_rays, nullptr).total;
}

static int check_train_args(const plnerf_render_cfg* cfg, const plnerf_net_desc* cdesc, const plnerf_net_desc* fdesc, int64_t n,
                            const float* rays, int stride, const void* ws, size_t ws_bytes, const char* who) {
  PLNERF_CHECK_ARG(cfg && cdesc && fdesc, "%s: null argument", who);
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || rays), "%s: null rays", who);
  PLNERF_CHECK_ARG(cdesc->use_viewdirs == fdesc->use_viewdirs, "%s: coarse/fine use_viewdirs differ", who);
  PLNERF_CHECK_ARG(stride >= (cdesc->use_viewdirs ? 11 : 8), "%s: ray stride too small (networks with view directions need stride >= 11)", who);
  PLNERF_CHECK_ARG(cfg->N_samples >= 2 && cfg->N_importance >= 0, "%s: need N_samples >= 2, N_importance >= 0", who);
  PLNERF_CHECK_ARG(cfg->mode == PLNERF_MODE_LINEAR || cfg->mode == PLNERF_MODE_CONSTANT, "%s: bad mode", who);
  if (cfg->precision != PLNERF_PREC_BF16) { set_error("%s: gradients are implemented for PLNERF_PREC_BF16 only", who); return PLNERF_E_UNSUPPORTED; }
  const size_t need = carve_train(cfg, cdesc, fdesc, n, nullptr).total;
  if (!ws || ws_bytes < need) { set_error("%s: workspace too small: need %zu bytes, got %zu", who, need, ws_bytes); return PLNERF_E_WORKSPACE; }
  PLNERF_CHECK_ARG(((uintptr_t)ws & 1023) == 0, "%s: workspace must be 1024-byte aligned", who);
  return PLNERF_OK;
}

// The loss and the coarse pass's backward as the fused training entry runs them: forked onto the side stream as soon as the
// coarse maps exist (nothing of the fine pass feeds them: the importance samples are detached, run_plnerf.py:728).
namespace plnerf {
struct CoarseBackward {
  const float* target; const int64_t* pix; float scale; float* sqerr;
  const void* cpacked_bwd; const plnerf_net_grads* grads_coarse; SideStream* side;
};
static int coarse_backward(const plnerf_render_cfg* cfg, const plnerf_net_desc* cdesc, const void* cpacked, const float* rays, int64_t n,
                           int stride, const float* noise0, const float* rgb0, const TrainWs& w, const CoarseBackward& cb, cudaStream_t sc,
                           float* graw_c, float* sqerr_slot) {
  const int Ns = cfg->N_samples, chc = cdesc->use_viewdirs ? 4 : cdesc->output_ch;
  int rc = launch_mse_loss_grad(rgb0, nullptr, cb.target, cb.pix, n, cb.scale, w.g_rgb0, nullptr, sqerr_slot, sc);
  if (!rc) rc = launch_composite_bwd(w.raw0, chc, w.z0, rays, n, stride, Ns, cfg->mode, cfg->color_mode, cfg->white_bkgd, cfg->farcolorfix,
                                     noise0, w.g_rgb0, nullptr, nullptr, nullptr, graw_c, sc, noise0 ? 0.f : cfg->raw_noise_std, cfg->seed,
                                     cfg->ray_id_offset, RNG_STREAM_NOISE0);
  if (!rc) rc = mlp_query_bwd(cdesc, cpacked, cb.cpacked_bwd, n, Ns, graw_c, chc, w.stash0, w.stash0_bytes, cb.grads_coarse, sc);
  return rc;
}
}  // namespace plnerf

static int train_forward(const plnerf_render_cfg* cfg, const plnerf_net_desc* cdesc, const void* cpacked,
                         const plnerf_net_desc* fdesc, const void* fpacked, const float* rays, int64_t n, int stride,
                         const float* t_rand, const float* u, const float* noise0, const float* noise1,
                         const plnerf_render_out* out, const TrainWs& w, cudaStream_t st, const CoarseBackward* cb) {
  int rc = PLNERF_OK;
  const int Ns = cfg->N_samples, Ni = cfg->N_importance, S1 = Ns + Ni;
  const int chc = cdesc->use_viewdirs ? 4 : cdesc->output_ch, chf = fdesc->use_viewdirs ? 4 : fdesc->output_ch;
  const bool fine = Ni > 0;
  // depths + the view bias of both networks + the direction encoding in one launch where covered (as in the inference entry)
  const bool two_nets = fine && fpacked != cpacked;
  float* vb_c = static_cast<float*>(w.mlp_ws);
  rc = launch_ray_setup(cdesc, cpacked, two_nets ? fdesc : nullptr, fpacked, cfg->precision, cfg->multires_views, rays, n, stride, Ns,
                        cfg->lindisp, cfg->perturb, t_rand, cfg->seed, cfg->ray_id_offset, w.z0, vb_c, w.vb_f, st, w.dirpe);
  if (rc < 0) return rc;
  const bool have_vb = (rc == 0);
  if (!have_vb) {
    rc = launch_stratified_z(rays, n, stride, Ns, cfg->lindisp, cfg->perturb, t_rand, cfg->seed, cfg->ray_id_offset, w.z0, st);
    if (rc) return rc;
  }
  rc = mlp_query_train(cdesc, cpacked, cfg->multires, cfg->multires_views, rays, n, stride, w.z0, Ns, w.raw0, chc, w.stash0,
                       w.stash0_bytes, w.mlp_ws, w.mlp_ws_bytes, st, have_vb ? vb_c : nullptr, have_vb ? w.dirpe : nullptr);
  if (rc) return rc;
  rc = launch_composite(w.raw0, chc, w.z0, rays, n, stride, Ns, cfg->mode, cfg->color_mode, cfg->white_bkgd, cfg->farcolorfix,
                        noise0, noise0 ? 0.f : cfg->raw_noise_std, cfg->seed, cfg->ray_id_offset, RNG_STREAM_NOISE0,
                        fine ? out->rgb0 : out->rgb_map, fine ? out->disp0 : out->disp_map, fine ? out->acc0 : out->acc_map,
                        fine ? out->depth0 : out->depth_map, fine ? w.w0 : nullptr, fine ? w.tau0 : nullptr,
                        fine ? w.T0 : nullptr, st);
  if (rc) return rc;
  if (!fine) {
    if (out->raw) PLNERF_CUDA(cudaMemcpyAsync(out->raw, w.raw0, (size_t)n * Ns * chc * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (out->z_vals) PLNERF_CUDA(cudaMemcpyAsync(out->z_vals, w.z0, (size_t)n * Ns * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return PLNERF_OK;
  }
  PLNERF_CHECK_ARG(out->rgb0 && out->disp0 && out->acc0 && out->depth0, "render_rays_fwd_train: coarse outputs are required when N_importance > 0");
  if (cb) {      // fused training entry: the coarse loss + backward start here, beside the fine pass
    PLNERF_CUDA(cudaEventRecord(cb->side->fork, st));
    PLNERF_CUDA(cudaStreamWaitEvent(cb->side->stream, cb->side->fork, 0));
    rc = coarse_backward(cfg, cdesc, cpacked, rays, n, stride, noise0, out->rgb0, w, *cb, cb->side->stream, w.graw0, cb->sqerr + 1);
    if (rc) return rc;
  }
  rc = launch_sample_merge(cfg->mode == PLNERF_MODE_LINEAR, w.z0, w.w0, w.tau0, w.T0, rays, n, stride, Ns, Ni, u, cfg->seed,
                           cfg->ray_id_offset, cfg->zero_tol, cfg->epsilon, w.z1, out->z_std, out->inds, st);
  if (rc) return rc;
  rc = mlp_query_train(fdesc, fpacked, cfg->multires, cfg->multires_views, rays, n, stride, w.z1, S1, w.raw1, chf, w.stash1,
                       w.stash1_bytes, w.mlp_ws, w.mlp_ws_bytes, st, have_vb ? (two_nets ? w.vb_f : vb_c) : nullptr,
                       have_vb ? w.dirpe : nullptr);
  if (rc) return rc;
  rc = launch_composite(w.raw1, chf, w.z1, rays, n, stride, S1, cfg->mode, cfg->color_mode, cfg->white_bkgd, cfg->farcolorfix,
                        noise1, noise1 ? 0.f : cfg->raw_noise_std, cfg->seed, cfg->ray_id_offset, RNG_STREAM_NOISE1,
                        out->rgb_map, out->disp_map, out->acc_map, out->depth_map, nullptr, nullptr, nullptr, st);
  if (rc) return rc;
  if (out->raw) PLNERF_CUDA(cudaMemcpyAsync(out->raw, w.raw1, (size_t)n * S1 * chf * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (out->z_vals) PLNERF_CUDA(cudaMemcpyAsync(out->z_vals, w.z1, (size_t)n * S1 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return PLNERF_OK;
}

int plnerf_render_rays_fwd_train(const plnerf_render_cfg* cfg, const plnerf_net_desc* cdesc, const void* cpacked,
                                 const plnerf_net_desc* fdesc, const void* fpacked, const float* rays, int64_t n, int stride,
                                 const float* t_rand, const float* u, const float* noise0, const float* noise1,
                                 const plnerf_render_out* out, void* ws, size_t ws_bytes, void* stream) {
  PLNERF_CHECK_ARG(cpacked && out, "render_rays_fwd_train: null argument");
  if (!fdesc || !fpacked) { fdesc = cdesc; fpacked = cpacked; }
  int rc = check_train_args(cfg, cdesc, fdesc, n, rays, stride, ws, ws_bytes, "render_rays_fwd_train");
  if (rc) return rc;
  PLNERF_CHECK_ARG(out->rgb_map && out->disp_map && out->acc_map && out->depth_map, "render_rays_fwd_train: rgb/disp/acc/depth outputs are required");
  if (n == 0) return PLNERF_OK;
  const TrainWs w = carve_train(cfg, cdesc, fdesc, n, static_cast<uint8_t*>(ws));
  return train_forward(cfg, cdesc, cpacked, fdesc, fpacked, rays, n, stride, t_rand, u, noise0, noise1, out, w, (cudaStream_t)stream, nullptr);
}

int plnerf_train_rays_mse(const plnerf_render_cfg* cfg, const plnerf_net_desc* cdesc, const void* cpacked, const void* cpacked_bwd,
                          const plnerf_net_desc* fdesc, const void* fpacked, const void* fpacked_bwd, const float* rays,
                          int64_t n, int stride, const float* t_rand, const float* u, const float* noise0, const float* noise1,
                          const float* target, const int64_t* pix, float scale, float* sqerr, const plnerf_render_out* out,
                          const plnerf_net_grads* grads_coarse, const plnerf_net_grads* grads_fine, void* ws, size_t ws_bytes,
                          void* stream) {
  PLNERF_CHECK_ARG(cpacked && cpacked_bwd && grads_coarse && sqerr && (n == 0 || target), "train_rays_mse: null argument");
  if (!fdesc || !fpacked) { fdesc = cdesc; fpacked = cpacked; fpacked_bwd = cpacked_bwd; grads_fine = grads_coarse; }
  PLNERF_CHECK_ARG(fpacked_bwd && grads_fine, "train_rays_mse: the fine network needs its transposed weights and gradient buffers");
  int rc = check_train_args(cfg, cdesc, fdesc, n, rays, stride, ws, ws_bytes, "train_rays_mse");
  if (rc) return rc;
  if (n == 0) return PLNERF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const TrainWs w = carve_train(cfg, cdesc, fdesc, n, static_cast<uint8_t*>(ws));
  const int Ns = cfg->N_samples, Ni = cfg->N_importance, S1 = Ns + Ni;
  const int chc = cdesc->use_viewdirs ? 4 : cdesc->output_ch, chf = fdesc->use_viewdirs ? 4 : fdesc->output_ch;
  // maps the caller does not ask for live in the workspace
  plnerf_render_out o;
  memset(&o, 0, sizeof(o));
  if (out) o = *out;
  float* m = w.maps;
  auto own = [&](float*& p, size_t per_ray) { if (!p) p = m; m += (size_t)n * per_ray; };
  own(o.rgb_map, 3); own(o.disp_map, 1); own(o.acc_map, 1); own(o.depth_map, 1);
  if (Ni > 0) { own(o.rgb0, 3); own(o.disp0, 1); own(o.acc0, 1); own(o.depth0, 1); }
  CoarseBackward cb{target, pix, scale, sqerr, cpacked_bwd, grads_coarse, nullptr};
  if (Ni > 0) {
    cb.side = side_stream();
    if (!cb.side) return PLNERF_E_CUDA;
  }
  rc = train_forward(cfg, cdesc, cpacked, fdesc, fpacked, rays, n, stride, t_rand, u, noise0, noise1, &o, w, st, Ni > 0 ? &cb : nullptr);
  if (!rc) {
    if (Ni > 0) {
      // fine pass: img2mse(rgb_map) (run_plnerf.py:1289), d(maps)/d(raw1), the fine network's parameter gradients
      rc = launch_mse_loss_grad(o.rgb_map, nullptr, target, pix, n, scale, w.g_rgb, nullptr, sqerr, st);
      if (!rc) rc = launch_composite_bwd(w.raw1, chf, w.z1, rays, n, stride, S1, cfg->mode, cfg->color_mode, cfg->white_bkgd, cfg->farcolorfix,
                                         noise1, w.g_rgb, nullptr, nullptr, nullptr, w.graw, st, noise1 ? 0.f : cfg->raw_noise_std, cfg->seed,
                                         cfg->ray_id_offset, RNG_STREAM_NOISE1);
      if (!rc) rc = mlp_query_bwd(fdesc, fpacked, fpacked_bwd, n, S1, w.graw, chf, w.stash1, w.stash1_bytes, grads_fine, st);
    } else {
      rc = coarse_backward(cfg, cdesc, cpacked, rays, n, stride, noise0, o.rgb_map, w, cb, st, w.graw, sqerr);
    }
  }
  if (cb.side) {      // always join (also after an error, once the fork may have happened)
    cudaError_t e = cudaEventRecord(cb.side->join, cb.side->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, cb.side->join, 0);
    if (e != cudaSuccess && !rc) rc = cuda_fail(e, "train_rays_mse: join");
  }
  return rc;
}

int plnerf_render_rays_bwd(const plnerf_render_cfg* cfg, const plnerf_net_desc* cdesc, const void* cpacked, const void* cpacked_bwd,
                           const plnerf_net_desc* fdesc, const void* fpacked, const void* fpacked_bwd, const float* rays,
                           int64_t n, int stride, const float* noise0, const float* noise1, const plnerf_render_grads* g,
                           const plnerf_net_grads* grads_coarse, const plnerf_net_grads* grads_fine, void* ws, size_t ws_bytes,
                           void* stream) {
  PLNERF_CHECK_ARG(cpacked && cpacked_bwd && g && grads_coarse, "render_rays_bwd: null argument");
  if (!fdesc || !fpacked) { fdesc = cdesc; fpacked = cpacked; fpacked_bwd = cpacked_bwd; grads_fine = grads_coarse; }
  PLNERF_CHECK_ARG(fpacked_bwd && grads_fine, "render_rays_bwd: the fine network needs its transposed weights and gradient buffers");
  int rc = check_train_args(cfg, cdesc, fdesc, n, rays, stride, ws, ws_bytes, "render_rays_bwd");
  if (rc) return rc;
  if (n == 0) return PLNERF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  TrainWs w = carve_train(cfg, cdesc, fdesc, n, static_cast<uint8_t*>(ws));
  const int Ns = cfg->N_samples, Ni = cfg->N_importance, S1 = Ns + Ni;
  const int chc = cdesc->use_viewdirs ? 4 : cdesc->output_ch, chf = fdesc->use_viewdirs ? 4 : fdesc->output_ch;
  // coarse pass (the importance samples are detached, run_plnerf.py:728: no gradient reaches it from the fine maps, so
  // the two passes are independent).  With both, the coarse pass is enqueued on a side stream forked from `stream` and
  // joined before returning: its CTAs take the SMs the fine pass's kernels leave idle (ragged last rounds, fill / drain).
  const float *gr = Ni > 0 ? g->g_rgb0 : g->g_rgb_map, *gd = Ni > 0 ? g->g_depth0 : g->g_depth_map;
  const float *ga = Ni > 0 ? g->g_acc0 : g->g_acc_map, *gp = Ni > 0 ? g->g_disp0 : g->g_disp_map;
  const bool coarse = (Ni == 0 || gr || gd || ga || gp);
  cudaStream_t sc = st;
  SideStream* side = nullptr;
  bool fork = Ni > 0 && coarse;
#ifdef PLNERF_DEBUG
  if (getenv("PLNERF_NO_SIDE")) fork = false;      // A/B switch, developer library only
#endif
  if (fork) {
    side = side_stream();
    if (!side) return PLNERF_E_CUDA;
    PLNERF_CUDA(cudaEventRecord(side->fork, st));
    PLNERF_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    sc = side->stream;
  }
  if (Ni > 0) {
    // fine pass: d(maps)/d(raw1) (run_plnerf.py:741 through autograd), then the fine network's parameter gradients
    rc = launch_composite_bwd(w.raw1, chf, w.z1, rays, n, stride, S1, cfg->mode, cfg->color_mode, cfg->white_bkgd, cfg->farcolorfix,
                              noise1, g->g_rgb_map, g->g_depth_map, g->g_acc_map, g->g_disp_map, w.graw, st,
                              noise1 ? 0.f : cfg->raw_noise_std, cfg->seed, cfg->ray_id_offset, RNG_STREAM_NOISE1);
    if (!rc) rc = mlp_query_bwd(fdesc, fpacked, fpacked_bwd, n, S1, w.graw, chf, w.stash1, w.stash1_bytes, grads_fine, st);
  }
  if (!rc && coarse) {
    float* graw_c = Ni > 0 ? w.graw0 : w.graw;
    rc = launch_composite_bwd(w.raw0, chc, w.z0, rays, n, stride, Ns, cfg->mode, cfg->color_mode, cfg->white_bkgd, cfg->farcolorfix,
                              noise0, gr, gd, ga, gp, graw_c, sc, noise0 ? 0.f : cfg->raw_noise_std, cfg->seed,
                              cfg->ray_id_offset, RNG_STREAM_NOISE0);
    if (!rc) rc = mlp_query_bwd(cdesc, cpacked, cpacked_bwd, n, Ns, graw_c, chc, w.stash0, w.stash0_bytes, grads_coarse, sc);
  }
  if (side) {      // always join, also after an error: `stream` must not be left with a dangling fork
    cudaError_t e = cudaEventRecord(side->join, side->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, side->join, 0);
    if (e != cudaSuccess && !rc) rc = cuda_fail(e, "render_rays_bwd: join");
  }
  return rc;
}

int plnerf_mse_loss_grad(const float* rgb, const float* rgb0, const float* target, const int64_t* pix, int64_t n,
                         float scale, float* g_rgb, float* g_rgb0, float* sqerr, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (rgb && target && g_rgb && sqerr)), "mse_loss_grad: null argument");
  PLNERF_CHECK_ARG((rgb0 == nullptr) == (g_rgb0 == nullptr), "mse_loss_grad: rgb0 and g_rgb0 go together");
  return launch_mse_loss_grad(rgb, rgb0, target, pix, n, scale, g_rgb, g_rgb0, sqerr, (cudaStream_t)stream);
}

int plnerf_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                     double beta2, double eps, int64_t step, int zero_grads, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (params && grads && exp_avg && exp_avg_sq)), "adam_step: null argument");
  PLNERF_CHECK_ARG(step >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "adam_step: need step >= 1 and betas in [0, 1)");
  return launch_adam_flat(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, zero_grads, (cudaStream_t)stream);
}

int plnerf_profile_enable(int on) { return profile_enable(on); }
int plnerf_profile_read(double* mlp_ms_sum, int64_t* mlp_launches, int64_t* mlp_rows) {
  return profile_read(mlp_ms_sum, mlp_launches, mlp_rows);
}

int plnerf_pack_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int c2w_ld,
                     const float* c2w_staticcam, int c2w_staticcam_ld, const float* rays_o, const float* rays_d,
                     int64_t n, int ndc, float ndc_cx, float ndc_cy, float ndc_near, float near, float far,
                     int use_viewdirs, float* out, int stride, void* stream) {
  PLNERF_CHECK_ARG(out && n >= 0, "pack_rays: null output or negative size");
  PLNERF_CHECK_ARG(c2w || (rays_o && rays_d), "pack_rays: need a pose or rays_o/rays_d");
  PLNERF_CHECK_ARG(!c2w || (H > 0 && W > 0 && n == (int64_t)H * W && c2w_ld >= 4), "pack_rays: with a pose, n must be H*W and c2w_ld >= 4");
  PLNERF_CHECK_ARG(!c2w_staticcam || (c2w && c2w_staticcam_ld >= 4), "pack_rays: c2w_staticcam needs c2w");
  PLNERF_CHECK_ARG(stride >= (use_viewdirs ? 11 : 8), "pack_rays: row stride too small");
  return launch_pack_rays(H, W, fx, fy, cx, cy, c2w, c2w_ld, c2w_staticcam, c2w_staticcam_ld, rays_o, rays_d, nullptr, n, ndc,
                          ndc_cx, ndc_cy, ndc_near, near, far, use_viewdirs, out, stride, (cudaStream_t)stream);
}

int plnerf_pack_pixel_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int c2w_ld,
                           const int64_t* pix, int64_t n, int ndc, float ndc_cx, float ndc_cy, float ndc_near, float near,
                           float far, int use_viewdirs, float* out, int stride, void* stream) {
  PLNERF_CHECK_ARG(n >= 0 && (n == 0 || (out && pix && c2w)), "pack_pixel_rays: null argument");
  PLNERF_CHECK_ARG(H > 0 && W > 0 && c2w_ld >= 4, "pack_pixel_rays: need H, W > 0 and c2w_ld >= 4");
  PLNERF_CHECK_ARG(stride >= (use_viewdirs ? 11 : 8), "pack_pixel_rays: row stride too small");
  return launch_pack_rays(H, W, fx, fy, cx, cy, c2w, c2w_ld, nullptr, 0, nullptr, nullptr, pix, n, ndc, ndc_cx, ndc_cy,
                          ndc_near, near, far, use_viewdirs, out, stride, (cudaStream_t)stream);
}

// ---- developer library only (-DPLNERF_DEBUG, libplnerf_b200_debug.so): bring-up / measurement entry points.  They are NOT
// part of the ABI in include/plnerf_b200.h and the product library does not export them.
#ifdef PLNERF_DEBUG
int plnerf_debug_umma_gemm(const float* A, const float* B, int N, int K, float* D, void* stream) {
  return debug_umma_gemm(A, B, N, K, D, (cudaStream_t)stream);
}

int plnerf_debug_umma_gemm_mn(const float* X, const float* Y, int N, int K, uint32_t lbo, uint32_t sbo, float* D, void* stream) {
  return plnerf::debug_umma_gemm_mn(X, Y, N, K, lbo, sbo, D, (cudaStream_t)stream);
}

// debug timeline buffer: 3 regions x 256 events x (clock, code) int64 (not part of the product ABI)
int plnerf_debug_set_trace(long long* buf) { return plnerf::debug_set_trace(buf); }

// bring-up microbenchmark (not part of the product ABI)
int plnerf_debug_mma_rate(int mode, int iters, int grid, long long* cycles_out, void* stream) {
  return plnerf::debug_mma_rate(mode, iters, grid, cycles_out, (cudaStream_t)stream);
}

// test-only variant with explicit operand mode and descriptor strides (not part of the product ABI)
int plnerf_debug_umma_gemm_ex(const float* A, const float* B, int N, int K, int a_mode, uint32_t lbo, uint32_t sbo,
                              float* D, void* stream) {
  return debug_umma_gemm_ex(A, B, N, K, a_mode, lbo, sbo, D, (cudaStream_t)stream);
}

#endif  // PLNERF_DEBUG

}  // extern "C"
