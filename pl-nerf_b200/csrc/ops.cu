// Non-GEMM operators of the PL-NeRF hot path (sm_100a): positional encoding, stratified depths,
// piecewise-linear / piecewise-constant quadrature (raw2outputs), the inverse-CDF samplers and the
// clamp + sort-merge.  One warp per ray, lanes striped along the sample axis (coalesced), prefix
// products / sums as warp-shuffle scans in fp64 rounded per element to fp32 (what torch's CPU
// cumprod/cumsum produce).  All of these are HBM-bound streaming kernels.
//
// Reference call sites are cited per kernel (paths relative to the reference repo).
#include <math.h>

#include "common.cuh"
#include "composite.cuh"
#include "ops.cuh"

namespace plnerf {

// =============================================================================================
// a5  Embedder.embed  (run_nerf_helpers.py:36-54)
// =============================================================================================
__global__ void k_encode(const float* __restrict__ x, int64_t n, int L, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (row, xyz)
  if (idx >= n * 3) return;
  const int64_t row = idx / 3;
  const int c = (int)(idx - row * 3);
  const int od = (L < 0) ? 3 : 3 + 6 * L;
  const float v = x[idx];
  float* o = out + row * od;
  o[c] = v;
  float f = 1.0f;
  for (int k = 0; k < L; ++k) {
    const float a = v * f;  // exact: power-of-two scale
    o[3 + 6 * k + c] = sinf(a);
    o[3 + 6 * k + 3 + c] = cosf(a);
    f *= 2.0f;
  }
}

// =============================================================================================
// a3  stratified depths  (run_plnerf.py:683-705)
// =============================================================================================
__device__ __forceinline__ float linspace01(int i, int n, float step) {
  // torch.linspace(0,1,n): first half start + step*i, second half end - step*(n-1-i), one rounding
  return (i < n / 2) ? step * (float)i : fmaf(-step, (float)(n - 1 - i), 1.0f);
}

__device__ __forceinline__ float base_z(float near, float far, float t, int lindisp) {
  const float omt = __fsub_rn(1.0f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
  const float a = __fmul_rn(__fdiv_rn(1.0f, near), omt);
  const float b = __fmul_rn(__fdiv_rn(1.0f, far), t);
  return __fdiv_rn(1.0f, __fadd_rn(a, b));
}

__global__ void k_stratified_z(const float* __restrict__ rays, int64_t n, int stride, int Ns, int lindisp,
                               int perturb, const float* __restrict__ t_rand, uint64_t seed,
                               uint64_t ray0, float* __restrict__ z_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * Ns) return;
  const int64_t r = idx / Ns;
  const int i = (int)(idx - r * Ns);
  const float near = rays[r * stride + 6], far = rays[r * stride + 7];
  const float step = (Ns > 1) ? __fdiv_rn(1.0f, (float)(Ns - 1)) : 0.0f;
  const float zi = base_z(near, far, linspace01(i, Ns, step), lindisp);
  float z = zi;
  if (perturb) {
    float lower = zi, upper = zi;
    if (i > 0) {
      const float zp = base_z(near, far, linspace01(i - 1, Ns, step), lindisp);
      lower = __fmul_rn(0.5f, __fadd_rn(zi, zp));
    }
    if (i < Ns - 1) {
      const float zn = base_z(near, far, linspace01(i + 1, Ns, step), lindisp);
      upper = __fmul_rn(0.5f, __fadd_rn(zn, zi));
    }
    const float t = t_rand ? t_rand[idx] : philox_uniform(seed, ray0 + (uint64_t)r, RNG_STREAM_TRAND, (uint32_t)i);
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t));
  }
  z_out[idx] = z;
}

// =============================================================================================
// a8/a9/a10  raw2outputs  (run_plnerf.py:504-624)
// =============================================================================================
template <int MODE>
__global__ void __launch_bounds__(256) k_composite(CompositeArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= a.n) return;
  composite_ray<MODE>(a, r, lane, RawGlobal{a.raw + r * (int64_t)a.S * a.raw_stride, a.raw_stride}, a.z + r * (int64_t)a.S);
}

int launch_composite(const float* raw, int raw_stride, const float* z, const float* rays, int64_t n,
                     int stride, int S, int mode, int color_mode, int white_bkgd, int farcolorfix,
                     const float* noise, float noise_std, uint64_t seed, uint64_t ray0, uint32_t noise_stream,
                     float* rgb_map, float* disp_map, float* acc_map, float* depth_map, float* weights,
                     float* tau, float* T, cudaStream_t st) {
  if (n == 0) return PLNERF_OK;
  CompositeArgs a;
  a.raw = raw; a.raw_stride = raw_stride; a.z = z; a.rays = rays; a.n = n; a.stride = stride; a.S = S;
  a.color_mode = color_mode; a.white_bkgd = white_bkgd; a.farcolorfix = farcolorfix;
  a.noise = noise; a.noise_std = noise_std; a.seed = seed; a.ray0 = ray0; a.noise_stream = noise_stream;
  a.rgb_map = rgb_map; a.disp_map = disp_map; a.acc_map = acc_map; a.depth_map = depth_map;
  a.weights = weights; a.tau = tau; a.T = T;
  const int threads = 256;
  const unsigned blocks = (unsigned)ceil_div(n * 32, threads);
  if (mode == PLNERF_MODE_LINEAR) k_composite<PLNERF_MODE_LINEAR><<<blocks, threads, 0, st>>>(a);
  else k_composite<PLNERF_MODE_CONSTANT><<<blocks, threads, 0, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_composite");
  return PLNERF_OK;
}

// =============================================================================================
// backward of raw2outputs (what autograd differentiates in the reference):
// upstream grads of rgb_map/depth_map/acc_map/disp_map -> grad of raw[..., :4].
//   linear:   dL/da_k = e_k T_k G_k - sum_{j>k} w_j G_j ;  dL/dtau_i = .5 (D_{i-1} dL/da_{i-1} + D_i dL/da_i)
//   constant: dL/dalpha_i = T_i G_i - (sum_{j>i} w_j G_j) / (1 - alpha_i + 1e-10)
// with G_j = gC.m_j + gD' zmid_j + gA' (m_j the colour of interval j).  One warp per ray; the forward
// scan is recomputed, the suffix sum is a reverse warp scan.
// =============================================================================================
struct CompositeBwdArgs {
  const float* raw; int raw_stride;
  const float* z; const float* rays; int64_t n; int stride; int S;
  int color_mode, white_bkgd, farcolorfix;
  const float* noise;
  const float *g_rgb, *g_depth, *g_acc, *g_disp;   // any may be null
  float* g_raw;                                      // [n,S,raw_stride]
  float noise_std; uint64_t seed, ray0; uint32_t noise_stream;   // noise == null && noise_std > 0: the forward's Philox draws
};

template <int MODE>
__global__ void __launch_bounds__(128) k_composite_bwd(CompositeBwdArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (r >= a.n) return;
  const int S = a.S;
  const int nI = (MODE == PLNERF_MODE_LINEAR) ? S + 1 : S;
  float* sw = smem + (size_t)wib * 4 * (S + 2);   // w_j
  float* sx = sw + (S + 2);                        // e_j T_j (linear) / T_j (constant)
  float* sy = sx + (S + 2);                        // delta_j (interval length * |d|) ; later dL/da_j
  float* sg = sy + (S + 2);                        // suffix sums of w_j G_j (exclusive)
  const float* ray = a.rays + r * a.stride;
  const float dx = ray[3], dy = ray[4], dz = ray[5];
  const float near = ray[6], far = ray[7];
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const float* raw = a.raw + r * (int64_t)S * a.raw_stride;
  const float* z = a.z + r * (int64_t)S;
  auto sigma_at = [&](int k) -> float {
    float s = raw[(int64_t)k * a.raw_stride + 3];
    if (a.noise) s += a.noise[r * (int64_t)S + k];
    else if (a.noise_std > 0.f)
      s = __fadd_rn(s, __fmul_rn(philox_normal(a.seed, a.ray0 + (uint64_t)r, a.noise_stream, (uint32_t)k), a.noise_std));
    return s;
  };
  auto color_at = [&](int k, int c) -> float { return sigmoidf_(raw[(int64_t)k * a.raw_stride + c]); };
  auto knot = [&](int k) -> float { return k == 0 ? near : (k == S + 1 ? far : z[k - 1]); };
  auto tau_at = [&](int k) -> float { return k == 0 ? 1e-10f : (k == S + 1 ? 1e10f : fmaxf(sigma_at(k - 1), 0.0f)); };

  // ---- pass 1: forward scan (same arithmetic as k_composite), totals
  float acc_d = 0.f, acc_w = 0.f;
  double carry = 1.0;
  for (int base = 0; base < nI; base += 32) {
    const int i = base + lane;
    const bool valid = i < nI;
    float e = 1.0f, delta = 0.f, zm = 0.f, fac = 0.f;
    if (valid) {
      if (MODE == PLNERF_MODE_LINEAR) {
        const float s0 = knot(i), s1 = knot(i + 1);
        delta = (s1 - s0) * dnorm;
        e = expf(-0.5f * (tau_at(i + 1) + tau_at(i)) * delta);
        fac = 1.0f - e;
        zm = 0.5f * (s1 + s0);
      } else {
        zm = z[i];
        delta = ((i < S - 1) ? (z[i + 1] - zm) : 1e10f) * dnorm;
        const float alpha = 1.0f - expf(-fmaxf(sigma_at(i), 0.f) * delta);
        fac = alpha;
        e = (1.0f - alpha) + 1e-10f;
      }
    }
    const double incl = warp_incl_scan_mul((double)e, lane) * carry;
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = carry;
    carry = __shfl_sync(0xffffffffu, incl, 31);
    if (valid) {
      const float T = (float)excl;
      const float w = fac * T;
      sw[i] = w;
      sx[i] = (MODE == PLNERF_MODE_LINEAR) ? e * T : T;
      sy[i] = delta;
      acc_d += w * zm;
      acc_w += w;
    }
  }
  acc_d = warp_sum(acc_d); acc_w = warp_sum(acc_w);
  __syncwarp();
  // ---- upstream scalars
  float gC[3] = {0.f, 0.f, 0.f};
  if (a.g_rgb) { gC[0] = a.g_rgb[r * 3]; gC[1] = a.g_rgb[r * 3 + 1]; gC[2] = a.g_rgb[r * 3 + 2]; }
  float gD = a.g_depth ? a.g_depth[r] : 0.f;
  float gA = a.g_acc ? a.g_acc[r] : 0.f;
  if (a.g_disp) {
    const float q = acc_d / acc_w;
    if (q > 1e-10f) {   // disp = 1/q ; below the clamp the output is constant
      const float gq = -a.g_disp[r] / (q * q);
      gD += gq / acc_w;
      gA += -gq * acc_d / (acc_w * acc_w);
    }
  }
  if (a.white_bkgd) gA -= gC[0] + gC[1] + gC[2];
  // colour of interval / sample j as used by the forward
  auto mcol = [&](int j, int c) -> float {
    if (MODE != PLNERF_MODE_LINEAR) return color_at(j, c);
    const int kl = max(j - 1, 0), kr = min(j, S - 1);
    if (a.color_mode == PLNERF_COLOR_LEFT) return color_at(kl, c);
    const float right = (a.farcolorfix && j == S) ? 0.f : color_at(kr, c);
    return 0.5f * (right + color_at(kl, c));
  };
  // ---- pass 2: reverse (suffix) scan of w_j G_j, exclusive
  float rcarry = 0.f;
  for (int top = ((nI - 1) / 32) * 32; top >= 0; top -= 32) {
    const int i = top + lane;
    float v = 0.f;
    if (i < nI) {
      float zm;
      if (MODE == PLNERF_MODE_LINEAR) zm = 0.5f * (knot(i + 1) + knot(i)); else zm = z[i];
      const float G = gC[0] * mcol(i, 0) + gC[1] * mcol(i, 1) + gC[2] * mcol(i, 2) + gD * zm + gA;
      v = sw[i] * G;
      sg[i] = G;   // stash G_i, replaced below
    }
    float incl = v;   // inclusive suffix within the chunk
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += t;
    }
    const float suffix_excl = incl - v + rcarry;
    rcarry += __shfl_sync(0xffffffffu, incl, 0);
    if (i < nI) {
      const float G = sg[i];
      if (MODE == PLNERF_MODE_LINEAR) sy[i] = sy[i] * (sx[i] * G - suffix_excl);   // delta_i * dL/da_i
      else sg[i] = suffix_excl;                                                    // constant mode: finished in pass 3
    }
  }
  __syncwarp();
  // ---- pass 3: per-sample gradients
  float* graw = a.g_raw + r * (int64_t)S * a.raw_stride;
  for (int i = lane; i < S; i += 32) {
    float gs = 0.f, gcr[3];
    const float sraw = sigma_at(i);
    if (MODE == PLNERF_MODE_LINEAR) {
      // sample i is knot i+1: dL/dtau = .5 (delta_i dL/da_i + delta_{i+1} dL/da_{i+1})
      gs = (sraw > 0.f) ? 0.5f * (sy[i] + sy[i + 1]) : 0.f;
      float cw;   // total weight multiplying c_i
      if (a.color_mode == PLNERF_COLOR_LEFT) cw = sw[i + 1] + (i == 0 ? sw[0] : 0.f);
      else cw = 0.5f * (sw[i] + sw[i + 1]) + (i == 0 ? 0.5f * sw[0] : 0.f) + ((i == S - 1 && !a.farcolorfix) ? 0.5f * sw[S] : 0.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) gcr[c] = cw * gC[c];
    } else {
      const float zi = z[i];
      const float delta = sy[i];
      const float sg_relu = fmaxf(sraw, 0.f);
      const float ex = expf(-sg_relu * delta);           // 1 - alpha
      const float G = gC[0] * color_at(i, 0) + gC[1] * color_at(i, 1) + gC[2] * color_at(i, 2) + gD * zi + gA;
      const float dLdalpha = sx[i] * G - sg[i] / (ex + 1e-10f);
      gs = (sraw > 0.f) ? dLdalpha * delta * ex : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) gcr[c] = sw[i] * gC[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float col = color_at(i, c);
      graw[(int64_t)i * a.raw_stride + c] = gcr[c] * col * (1.0f - col);
    }
    graw[(int64_t)i * a.raw_stride + 3] = gs;
    for (int c = 4; c < a.raw_stride; ++c) graw[(int64_t)i * a.raw_stride + c] = 0.f;
  }
}

int launch_composite_bwd(const float* raw, int raw_stride, const float* z, const float* rays, int64_t n, int stride,
                         int S, int mode, int color_mode, int white_bkgd, int farcolorfix, const float* noise,
                         const float* g_rgb, const float* g_depth, const float* g_acc, const float* g_disp,
                         float* g_raw, cudaStream_t st, float noise_std, uint64_t seed, uint64_t ray0, uint32_t noise_stream) {
  if (n == 0) return PLNERF_OK;
  CompositeBwdArgs a{raw, raw_stride, z, rays, n, stride, S, color_mode, white_bkgd, farcolorfix, noise,
                     g_rgb, g_depth, g_acc, g_disp, g_raw, noise_std, seed, ray0, noise_stream};
  const int wpb = 4;
  const size_t smem = (size_t)wpb * 4 * (S + 2) * sizeof(float);
  if (smem > 48 * 1024) { set_error("raw2outputs_bwd: N_samples=%d too large", S); return PLNERF_E_UNSUPPORTED; }
  const unsigned blocks = (unsigned)ceil_div(n, wpb);
  if (mode == PLNERF_MODE_LINEAR) k_composite_bwd<PLNERF_MODE_LINEAR><<<blocks, wpb * 32, smem, st>>>(a);
  else k_composite_bwd<PLNERF_MODE_CONSTANT><<<blocks, wpb * 32, smem, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_composite_bwd");
  return PLNERF_OK;
}

// =============================================================================================
// samplers + sort-merge.  One warp per ray; the per-ray steps are device functions shared by the op-level kernels
// (k_sample_pl, k_sample_const, k_merge) and by the fused k_sample_merge that render_rays uses (samples never leave the SM).
// =============================================================================================
// torch.searchsorted(cdf, u, right=True): ATen's upper-bound loop, reproduced step for step so the
// result is identical even where rounding makes cdf non-monotone at its forced last element.
__device__ __forceinline__ int upper_bound_torch(const float* cdf, int n, float u) {
  int start = 0, end = n;
  while (start < end) {
    const int mid = start + ((end - start) >> 1);
    if (!(cdf[mid] > u)) start = mid + 1; else end = mid;
  }
  return start;
}

// ---- a11  sample_pdf_reformulation + pw_linear_sample_{in,de}creasing  (run_nerf_helpers.py:340-445)
struct SamplePLArgs {
  const float *z, *w, *tau, *T, *rays;
  int64_t n; int stride, S, Ni;
  const float* u; uint64_t seed, ray0;
  float zero_tol, eps;
  float* samples; int64_t* inds;
  // f-4 (sample_pdf_reformulation_return_u, run_nerf_helpers.py:448-533): optional extra returns, each [n,Ni]
  float *T_below, *tau_below, *bin_below, *u_out;
};

// knots s = [near, z.., far], T, tau -> shared memory; cdf = [0, cumsum(w)] with fp64 accumulation, last forced to 1
__device__ __forceinline__ void pl_build(const SamplePLArgs& a, int64_t r, int lane, float* cdf, float* s, float* T, float* tau) {
  const int S = a.S, nk = S + 2;
  const float near = a.rays[r * a.stride + 6], far = a.rays[r * a.stride + 7];
  for (int k = lane; k < nk; k += 32) {
    s[k] = (k == 0) ? near : (k == S + 1 ? far : a.z[r * (int64_t)S + k - 1]);
    T[k] = a.T[r * (int64_t)nk + k];
    tau[k] = a.tau[r * (int64_t)nk + k];
  }
  double carry = 0.0;
  if (lane == 0) cdf[0] = 0.0f;
  for (int base = 0; base < S + 1; base += 32) {
    const int i = base + lane;
    const double wv = (i < S + 1) ? (double)a.w[r * (int64_t)(S + 1) + i] : 0.0;
    const double incl = warp_incl_scan_add(wv, lane) + carry;
    carry = __shfl_sync(0xffffffffu, incl, 31);
    if (i < S + 1) cdf[i + 1] = (i == S) ? 1.0f : (float)incl;
  }
  __syncwarp();
}

// sample k of ray r: the draw, the bracket, the closed-form inverse; writes the optional per-sample outputs
__device__ __forceinline__ float pl_sample(const SamplePLArgs& a, int64_t r, int k, const float* cdf, const float* s,
                                           const float* T, const float* tau, const float* u_drawn = nullptr) {
  const int S = a.S, nk = S + 2;
  const float eps = a.eps, tol = a.zero_tol;
  const float u = u_drawn ? u_drawn[k]
                          : (a.u ? a.u[r * (int64_t)a.Ni + k] : philox_uniform(a.seed, a.ray0 + (uint64_t)r, RNG_STREAM_U, (uint32_t)k));
  const int ind = upper_bound_torch(cdf, nk, u);
  const int below = max(0, ind - 1);
  const int above = min(nk - 1, ind);
  const float s_l = s[below], s_r = s[above];
  const float T_l = T[below];
  const float tau_l = tau[below], tau_r = tau[above];
  // tau_diff gathered at `below` from tau[1:]-tau[:-1] (size S+1); the reference raises for
  // below == S+1 (only reachable with u >= 1): we clamp instead of faulting.
  const int bd = min(below, S);
  const float dtau = __fsub_rn(tau[bd + 1], tau[bd]);
  float x;
  if (dtau < tol && dtau > -tol) {
    x = s_l;
  } else {
    const float ln_term = -logf(fmaxf(eps, __fdiv_rn(__fsub_rn(1.0f, u), fmaxf(eps, T_l))));
    const float ds = __fsub_rn(s_r, s_l);
    float t;
    if (dtau >= tol) {
      const float disc = __fadd_rn(__fmul_rn(tau_l, tau_l),
                                   __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, __fsub_rn(tau_r, tau_l)), ln_term), fmaxf(eps, ds)));
      t = __fdiv_rn(__fmul_rn(ds, __fadd_rn(-tau_l, sqrtf(fmaxf(eps, disc)))), fmaxf(eps, __fsub_rn(tau_r, tau_l)));
    } else {
      const float disc = __fsub_rn(__fmul_rn(tau_l, tau_l),
                                   __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, __fsub_rn(tau_l, tau_r)), ln_term), fmaxf(eps, ds)));
      t = __fdiv_rn(__fmul_rn(ds, __fsub_rn(tau_l, sqrtf(fmaxf(eps, disc)))), fmaxf(eps, __fsub_rn(tau_l, tau_r)));
    }
    // torch.clamp(t, min=eps, max=ds): min(max(t, eps), ds) -- max wins, NaN propagates
    if (t == t) t = fminf(fmaxf(t, eps), ds);
    x = __fadd_rn(s_l, t);
    if (x != x) x = s_l;
  }
  if (a.inds) a.inds[r * (int64_t)a.Ni + k] = (int64_t)ind;
  if (a.T_below) a.T_below[r * (int64_t)a.Ni + k] = T_l;
  if (a.tau_below) a.tau_below[r * (int64_t)a.Ni + k] = tau_l;
  if (a.bin_below) a.bin_below[r * (int64_t)a.Ni + k] = s_l;
  if (a.u_out) a.u_out[r * (int64_t)a.Ni + k] = u;
  return x;
}

__global__ void __launch_bounds__(128) k_sample_pl(SamplePLArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (r >= a.n) return;
  const int nk = a.S + 2;
  float* cdf = smem + (size_t)wib * 4 * nk;
  float *s = cdf + nk, *T = s + nk, *tau = T + nk;
  pl_build(a, r, lane, cdf, s, T, tau);
  for (int k = lane; k < a.Ni; k += 32) a.samples[r * (int64_t)a.Ni + k] = pl_sample(a, r, k, cdf, s, T, tau);
}

int launch_sample_pl(const float* z, const float* w, const float* tau, const float* T, const float* rays,
                     int64_t n, int stride, int S, int Ni, const float* u, uint64_t seed, uint64_t ray0,
                     float zero_tol, float eps, float* samples, int64_t* inds, cudaStream_t st, float* T_below,
                     float* tau_below, float* bin_below, float* u_out) {
  if (n == 0 || Ni == 0) return PLNERF_OK;
  SamplePLArgs a{z, w, tau, T, rays, n, stride, S, Ni, u, seed, ray0, zero_tol, eps, samples, inds, T_below, tau_below, bin_below, u_out};
  const int wpb = 4;
  const size_t smem = (size_t)wpb * 4 * (S + 2) * sizeof(float);
  if (smem > 48 * 1024) { set_error("sample_pdf_pl: N_samples=%d too large", S); return PLNERF_E_UNSUPPORTED; }
  k_sample_pl<<<(unsigned)ceil_div(n, wpb), wpb * 32, smem, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_sample_pl");
  return PLNERF_OK;
}

// ---- a12  sample_pdf  (run_nerf_helpers.py:241-284)
struct SampleConstArgs {
  const float* bins; int bins_stride; int bins_mid;   // bins_mid: bins_k = .5*(z[k+1]+z[k]) from z rows
  const float* w; int w_stride;                      // weights row stride (slice [1:-1] of a wider row)
  int64_t n; int nb, Ni;
  const float* u; uint64_t seed, ray0;
  float* samples; int64_t* inds;
  float* u_out;                                      // f-4 (sample_pdf_return_u, run_nerf_helpers.py:286-337): optional [n,Ni]
};

__device__ __forceinline__ void const_build(const SampleConstArgs& a, int64_t r, int lane, float* cdf, float* bins) {
  const int nb = a.nb, nw = nb - 1;
  const float* brow = a.bins + r * (int64_t)a.bins_stride;
  for (int k = lane; k < nb; k += 32)
    bins[k] = a.bins_mid ? __fmul_rn(0.5f, __fadd_rn(brow[k + 1], brow[k])) : brow[k];
  const float* wrow = a.w + r * (int64_t)a.w_stride;
  float tot = 0.f;
  for (int k = lane; k < nw; k += 32) tot += __fadd_rn(wrow[k], 1e-5f);
  tot = warp_sum(tot);
  double carry = 0.0;
  if (lane == 0) cdf[0] = 0.0f;
  for (int base = 0; base < nw; base += 32) {
    const int i = base + lane;
    const double pv = (i < nw) ? (double)__fdiv_rn(__fadd_rn(wrow[i], 1e-5f), tot) : 0.0;
    const double incl = warp_incl_scan_add(pv, lane) + carry;
    carry = __shfl_sync(0xffffffffu, incl, 31);
    if (i < nw) cdf[i + 1] = (float)incl;
  }
  __syncwarp();
}

__device__ __forceinline__ float const_sample(const SampleConstArgs& a, int64_t r, int k, const float* cdf, const float* bins,
                                              const float* u_drawn = nullptr) {
  const int nb = a.nb;
  const float u = u_drawn ? u_drawn[k]
                          : (a.u ? a.u[r * (int64_t)a.Ni + k] : philox_uniform(a.seed, a.ray0 + (uint64_t)r, RNG_STREAM_U, (uint32_t)k));
  const int ind = upper_bound_torch(cdf, nb, u);
  const int below = max(0, ind - 1);
  const int above = min(nb - 1, ind);
  float denom = __fsub_rn(cdf[above], cdf[below]);
  if (denom < 1e-5f) denom = 1.0f;
  const float t = __fdiv_rn(__fsub_rn(u, cdf[below]), denom);
  if (a.inds) a.inds[r * (int64_t)a.Ni + k] = (int64_t)ind;
  if (a.u_out) a.u_out[r * (int64_t)a.Ni + k] = u;
  return __fadd_rn(bins[below], __fmul_rn(t, __fsub_rn(bins[above], bins[below])));
}

__global__ void __launch_bounds__(128) k_sample_const(SampleConstArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (r >= a.n) return;
  float* cdf = smem + (size_t)wib * 2 * a.nb;
  float* bins = cdf + a.nb;
  const_build(a, r, lane, cdf, bins);
  for (int k = lane; k < a.Ni; k += 32) a.samples[r * (int64_t)a.Ni + k] = const_sample(a, r, k, cdf, bins);
}

int launch_sample_const(const float* bins, int bins_stride, int bins_mid, const float* w, int w_stride,
                        int64_t n, int nb, int Ni, const float* u, uint64_t seed, uint64_t ray0,
                        float* samples, int64_t* inds, cudaStream_t st, float* u_out) {
  if (n == 0 || Ni == 0) return PLNERF_OK;
  if (nb < 2) { set_error("sample_pdf: need at least 2 bins"); return PLNERF_E_BADARG; }
  SampleConstArgs a{bins, bins_stride, bins_mid, w, w_stride, n, nb, Ni, u, seed, ray0, samples, inds, u_out};
  const int wpb = 4;
  const size_t smem = (size_t)wpb * 2 * nb * sizeof(float);
  if (smem > 48 * 1024) { set_error("sample_pdf: %d bins too many", nb); return PLNERF_E_UNSUPPORTED; }
  k_sample_const<<<(unsigned)ceil_div(n, wpb), wpb * 32, smem, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_sample_const");
  return PLNERF_OK;
}

// ---- f-4  gradients of the depth-experiment samplers (what autograd computes through
// sample_pdf_reformulation_return_u, run_nerf_helpers.py:448-533, and sample_pdf_return_u, :286-337): the bracket comes
// from searchsorted (no gradient); the sample is a closed-form function of the gathered knots, so each sample sends
// gradient to the two knots of its bracket.  One warp per ray: pass 1 leaves every sample's contributions in shared memory,
// pass 2 gives each knot to one lane, which adds its samples' contributions in sample order (deterministic, no atomics).
// max(eps, x) routes like torch.max (ties 1/2), clamp(t, min=eps, max=ds) like torch.clamp with tensor bounds
// (max < min or t > max: to the bound; min <= t <= max: to t), where(isnan) sends a NaN sample's gradient to s_left.
struct SamplePLBwdArgs {
  SamplePLArgs f;                                                    // forward inputs; f.u = the forward's u (u_out / load_u)
  const float *g_samples, *g_T_below, *g_tau_below, *g_bin_below;   // [n,Ni] cotangents, any may be null
  float *g_z, *g_near, *g_far, *g_tau, *g_T;                         // [n,S], [n], [n], [n,S+2], [n,S+2]: written; any may be null
};
__device__ __forceinline__ float max_grad_first(float c, float x) {   // d max(c, x) / dx with torch's tie rule
  return x > c ? 1.0f : (x == c ? 0.5f : 0.0f);
}
__global__ void __launch_bounds__(128) k_sample_pl_bwd(const __grid_constant__ SamplePLBwdArgs a) {
  extern __shared__ float smem[];
  const SamplePLArgs& f = a.f;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (r >= f.n) return;
  const int S = f.S, nk = S + 2, Ni = f.Ni;
  float* cdf = smem + (size_t)wib * (4 * nk + 7 * Ni);
  float *s = cdf + nk, *T = s + nk, *tau = T + nk;
  int* rb = reinterpret_cast<int*>(tau + nk);          // per sample: below, above, then 5 gradient values
  int* ra = rb + Ni;
  float *r_sl = reinterpret_cast<float*>(ra + Ni), *r_sr = r_sl + Ni, *r_T = r_sr + Ni, *r_tl = r_T + Ni, *r_tr = r_tl + Ni;
  pl_build(f, r, lane, cdf, s, T, tau);
  const float eps = f.eps, tol = f.zero_tol;
  for (int k = lane; k < Ni; k += 32) {
    const int64_t o = r * (int64_t)Ni + k;
    const float u = f.u[o];
    const int ind = upper_bound_torch(cdf, nk, u);
    const int below = max(0, ind - 1), above = min(nk - 1, ind);
    const float s_l = s[below], s_r = s[above], T_l = T[below], tau_l = tau[below], tau_r = tau[above];
    const int bd = min(below, S);
    const float dtau = __fsub_rn(tau[bd + 1], tau[bd]);
    const float gx = a.g_samples ? a.g_samples[o] : 0.f;
    float g_sl = gx, g_sr = 0.f, g_Tl = 0.f, g_tl = 0.f, g_tr = 0.f;
    if (!(dtau < tol && dtau > -tol)) {
      const bool inc = dtau >= tol;
      const float Tm = fmaxf(eps, T_l), c1 = (1.0f - u) / Tm, m1 = fmaxf(eps, c1), L = -logf(m1);
      const float dsr = s_r - s_l, dsm = fmaxf(eps, dsr);
      const float A = inc ? tau_r - tau_l : tau_l - tau_r;          // the positive slope term of either branch
      const float sgn = inc ? 1.0f : -1.0f;
      const float D = tau_l * tau_l + sgn * (2.0f * A * L) / dsm;
      const float sq = sqrtf(fmaxf(eps, D)), dt = fmaxf(eps, A);
      const float num = inc ? (-tau_l + sq) : (tau_l - sq);
      const float t0 = (dsr * num) / dt;
      float t = t0;
      if (t == t) t = fminf(fmaxf(t, eps), dsr);
      const float x = s_l + t;
      if (x == x) {
        float g_dsr = 0.f, g_t0 = 0.f;
        if (dsr < eps || t0 > dsr) g_dsr = gx;                     // the result is the max bound
        else if (t0 >= eps) g_t0 = gx;                              // (t0 < eps: the constant min bound)
        g_dsr += g_t0 * (num / dt);
        const float g_num = g_t0 * (dsr / dt);
        const float g_dt = -g_t0 * (t0 / dt);
        g_tl += inc ? -g_num : g_num;
        const float g_sq = inc ? g_num : -g_num;
        const float g_D = g_sq * (0.5f / sq) * max_grad_first(eps, D);
        g_tl += 2.0f * tau_l * g_D;
        float g_A = sgn * (2.0f * L / dsm) * g_D;
        const float g_L = sgn * (2.0f * A / dsm) * g_D;
        const float g_dsm = -sgn * (2.0f * A * L) / (dsm * dsm) * g_D;
        g_A += max_grad_first(eps, A) * g_dt;
        if (inc) { g_tr += g_A; g_tl -= g_A; } else { g_tl += g_A; g_tr -= g_A; }
        g_dsr += max_grad_first(eps, dsr) * g_dsm;
        g_sr += g_dsr; g_sl -= g_dsr;
        const float g_m1 = -g_L / m1;
        const float g_c1 = max_grad_first(eps, c1) * g_m1;
        const float g_Tm = -g_c1 * (c1 / Tm);
        g_Tl += max_grad_first(eps, T_l) * g_Tm;
      }
    }
    if (a.g_bin_below) g_sl += a.g_bin_below[o];
    if (a.g_T_below) g_Tl += a.g_T_below[o];
    if (a.g_tau_below) g_tl += a.g_tau_below[o];
    rb[k] = below; ra[k] = above;
    r_sl[k] = g_sl; r_sr[k] = g_sr; r_T[k] = g_Tl; r_tl[k] = g_tl; r_tr[k] = g_tr;
  }
  __syncwarp();
  for (int j = lane; j < nk; j += 32) {
    float gs = 0.f, gT = 0.f, gt = 0.f;
    for (int k = 0; k < Ni; ++k) {
      if (rb[k] == j) { gs += r_sl[k]; gT += r_T[k]; gt += r_tl[k]; }
      if (ra[k] == j) { gs += r_sr[k]; gt += r_tr[k]; }
    }
    if (a.g_T) a.g_T[r * (int64_t)nk + j] = gT;
    if (a.g_tau) a.g_tau[r * (int64_t)nk + j] = gt;
    if (j == 0) { if (a.g_near) a.g_near[r] = gs; }
    else if (j == S + 1) { if (a.g_far) a.g_far[r] = gs; }
    else if (a.g_z) a.g_z[r * (int64_t)S + j - 1] = gs;
  }
}

int launch_sample_pl_bwd(const float* z, const float* w, const float* tau, const float* T, const float* rays, int64_t n,
                         int stride, int S, int Ni, const float* u, float zero_tol, float eps, const float* g_samples,
                         const float* g_T_below, const float* g_tau_below, const float* g_bin_below, float* g_z, float* g_near,
                         float* g_far, float* g_tau, float* g_T, cudaStream_t st) {
  if (n == 0) return PLNERF_OK;
  SamplePLBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.f = SamplePLArgs{z, w, tau, T, rays, n, stride, S, Ni, u, 0, 0, zero_tol, eps, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  a.g_samples = g_samples; a.g_T_below = g_T_below; a.g_tau_below = g_tau_below; a.g_bin_below = g_bin_below;
  a.g_z = g_z; a.g_near = g_near; a.g_far = g_far; a.g_tau = g_tau; a.g_T = g_T;
  const int wpb = 4;
  const size_t smem = (size_t)wpb * (4 * (S + 2) + 7 * Ni) * sizeof(float);
  if (smem > 48 * 1024) { set_error("sample_pdf_pl backward: S=%d Ni=%d too large", S, Ni); return PLNERF_E_UNSUPPORTED; }
  k_sample_pl_bwd<<<(unsigned)ceil_div(n, wpb), wpb * 32, smem, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_sample_pl_bwd");
  return PLNERF_OK;
}

// sample_pdf_return_u (:286-337): x = bins_b + t (bins_a - bins_b), t = (u - cdf_b) / denom, cdf = [0, cumsum(pdf)],
// pdf = (w + 1e-5) / sum(w + 1e-5): gradient reaches the bins directly and the weights through the cdf.
struct SampleConstBwdArgs {
  SampleConstArgs f;              // forward inputs (bins_mid must be 0: explicit bins); f.u = the forward's u
  const float* g_samples;         // [n,Ni]
  float *g_bins, *g_w;            // [n,nb], [n,nb-1]: written; either may be null
};
__global__ void __launch_bounds__(128) k_sample_const_bwd(const __grid_constant__ SampleConstBwdArgs a) {
  extern __shared__ float smem[];
  const SampleConstArgs& f = a.f;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (r >= f.n) return;
  const int nb = f.nb, nw = nb - 1, Ni = f.Ni;
  float* cdf = smem + (size_t)wib * (3 * nb + 6 * Ni);
  float* bins = cdf + nb;
  float* gcdf = bins + nb;                              // d L / d cdf, then (in place) d L / d pdf
  int* rb = reinterpret_cast<int*>(gcdf + nb);
  int* ra = rb + Ni;
  float *r_bb = reinterpret_cast<float*>(ra + Ni), *r_ba = r_bb + Ni, *r_cb = r_ba + Ni, *r_ca = r_cb + Ni;
  const_build(f, r, lane, cdf, bins);
  for (int k = lane; k < Ni; k += 32) {
    const int64_t o = r * (int64_t)Ni + k;
    const float u = f.u[o], gx = a.g_samples[o];
    const int ind = upper_bound_torch(cdf, nb, u);
    const int below = max(0, ind - 1), above = min(nb - 1, ind);
    const float denom0 = __fsub_rn(cdf[above], cdf[below]);
    const float denom = denom0 < 1e-5f ? 1.0f : denom0;
    const float t = (u - cdf[below]) / denom;
    const float g_t = gx * (bins[above] - bins[below]);
    float g_cb = -g_t / denom, g_ca = 0.f;
    if (!(denom0 < 1e-5f)) { const float g_den = -g_t * t / denom; g_ca = g_den; g_cb -= g_den; }
    rb[k] = below; ra[k] = above;
    r_bb[k] = gx * (1.0f - t); r_ba[k] = gx * t; r_cb[k] = g_cb; r_ca[k] = g_ca;
  }
  __syncwarp();
  for (int j = lane; j < nb; j += 32) {
    float gb = 0.f, gc = 0.f;
    for (int k = 0; k < Ni; ++k) {
      if (rb[k] == j) { gb += r_bb[k]; gc += r_cb[k]; }
      if (ra[k] == j) { gb += r_ba[k]; gc += r_ca[k]; }
    }
    if (a.g_bins) a.g_bins[r * (int64_t)nb + j] = gb;
    gcdf[j] = gc;
  }
  __syncwarp();
  if (!a.g_w) return;
  // cdf[i] = sum_{j < i} pdf[j]  ->  d/d pdf[j] = sum_{i > j} gcdf[i] (suffix sums, sequential per lane chunk: nb is small);
  // pdf = wt / W  ->  d/d w[j] = (gpdf[j] - sum_k gpdf[k] pdf[k]) / W
  const float* wrow = f.w + r * (int64_t)f.w_stride;
  float tot = 0.f;
  for (int k = lane; k < nw; k += 32) tot += __fadd_rn(wrow[k], 1e-5f);
  tot = warp_sum(tot);
  if (lane == 0) {
    float run = 0.f, above = gcdf[nw];
    for (int j = nw - 1; j >= 0; --j) { const float own = gcdf[j]; run += above; gcdf[j] = run; above = own; }   // gcdf[j] := d L / d pdf[j]
  }
  __syncwarp();
  float dot = 0.f;
  for (int k = lane; k < nw; k += 32) dot += gcdf[k] * (__fadd_rn(wrow[k], 1e-5f) / tot);
  dot = warp_sum(dot);
  for (int k = lane; k < nw; k += 32) a.g_w[r * (int64_t)nw + k] = (gcdf[k] - dot) / tot;
}

int launch_sample_const_bwd(const float* bins, const float* w, int64_t n, int nb, int Ni, const float* u, const float* g_samples,
                            float* g_bins, float* g_w, cudaStream_t st) {
  if (n == 0) return PLNERF_OK;
  if (nb < 2) { set_error("sample_pdf backward: need at least 2 bins"); return PLNERF_E_BADARG; }
  SampleConstBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.f = SampleConstArgs{bins, nb, 0, w, nb - 1, n, nb, Ni, u, 0, 0, nullptr, nullptr, nullptr};
  a.g_samples = g_samples; a.g_bins = g_bins; a.g_w = g_w;
  const int wpb = 4;
  const size_t smem = (size_t)wpb * (3 * nb + 6 * Ni) * sizeof(float);
  if (smem > 48 * 1024) { set_error("sample_pdf backward: nb=%d Ni=%d too large", nb, Ni); return PLNERF_E_UNSUPPORTED; }
  k_sample_const_bwd<<<(unsigned)ceil_div(n, wpb), wpb * 32, smem, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_sample_const_bwd");
  return PLNERF_OK;
}

// ---- a13  clamp + sort(cat(z_vals, z_samples)) + std   (run_plnerf.py:728-734, :752)
// Sort the Ni clamped samples inside the warp (bitonic, padded to a power of two), then merge the two ascending lists by
// binary searches (coarse depths first on ties) -- the output is the sorted multiset, like torch.sort.
__device__ __forceinline__ float clamp_sample(float x, float near, float far) {
  if (x == x) x = fminf(fmaxf(x, near), far);  // torch.clamp(x, near, far); NaN propagates
  return x;
}

// Ascending bitonic sort of 32 * KPL keys held KPL per lane (element index lane * KPL + j); no NaNs.
template <int KPL>
__device__ __forceinline__ void warp_sort_regs(float (&v)[KPL], int lane) {
#pragma unroll
  for (int size = 2; size <= 32 * KPL; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= KPL) {
        // partner element = same j in lane ^ (stride / KPL); bit `size` of the element index comes from the lane (size > KPL)
        const int ls = stride / KPL;
        const bool up = ((lane * KPL) & size) == 0, lower = (lane & ls) == 0;
        const bool keep_min = (lower == up);
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          const float o = __shfl_xor_sync(0xffffffffu, v[j], ls);
          v[j] = keep_min ? fminf(v[j], o) : fmaxf(v[j], o);
        }
      } else {
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          if ((j & stride) == 0) {
            const bool up = ((lane * KPL + j) & size) == 0;
            const float p = v[j], q = v[j ^ stride];
            const float mn = fminf(p, q), mx = fmaxf(p, q);
            v[j] = up ? mn : mx;
            v[j ^ stride] = up ? mx : mn;
          }
        }
      }
    }
  }
}
// xs[0..Ni) (shared memory, no NaNs) -> xsorted[0..32 KPL) ascending, +inf beyond the Ni keys
template <int KPL>
__device__ __forceinline__ void warp_sort_to(const float* xs, float* xsorted, int Ni, int lane) {
  float v[KPL];
#pragma unroll
  for (int j = 0; j < KPL; ++j) { const int e = lane * KPL + j; v[j] = (e < Ni) ? xs[e] : __int_as_float(0x7f800000); }
  warp_sort_regs<KPL>(v, lane);
#pragma unroll
  for (int j = 0; j < KPL; ++j) xsorted[lane * KPL + j] = v[j];
  __syncwarp();
}

// xs[0..Ni) clamped samples (unsorted), xsorted[0..np2) scratch, zc[0..S) ascending coarse depths (all shared memory)
__device__ __forceinline__ void merge_ray(const float* zc, const float* xs, float* xsorted, int S, int Ni, int np2, int lane,
                                          float* out, float* z_std_out) {
  if (z_std_out) {  // torch.std(z_samples, unbiased=False)
    float sum = 0.f;
    for (int k = lane; k < Ni; k += 32) sum += xs[k];
    const float mean = __fdiv_rn(warp_sum(sum), (float)Ni);
    float v = 0.f;
    for (int k = lane; k < Ni; k += 32) { const float d = xs[k] - mean; v += d * d; }
    v = warp_sum(v);
    if (lane == 0) *z_std_out = sqrtf(__fdiv_rn(v, (float)Ni));
  }
  // No NaN among the samples (the rule): sort in registers, KPL keys per lane (element lane * KPL + j), +inf pads -- the
  // bitonic network's strides below KPL are register compare-exchanges, the others one shuffle + one min/max per key
  // (equal keys have equal bits, so min / max give torch.sort's output).  ~300 instructions per lane for 128 keys against
  // ~1000 plus shared-memory latency for the network run in shared memory, which remains for NaN inputs and Ni > 256.
  bool has_nan = false;
  for (int k = lane; k < Ni; k += 32) has_nan |= (xs[k] != xs[k]);
  has_nan = __any_sync(0xffffffffu, has_nan);
  bool sorted_in_regs = false;
  if (!has_nan) {
    sorted_in_regs = true;
    if (np2 == 32) warp_sort_to<1>(xs, xsorted, Ni, lane);
    else if (np2 == 64) warp_sort_to<2>(xs, xsorted, Ni, lane);
    else if (np2 == 128) warp_sort_to<4>(xs, xsorted, Ni, lane);
    else if (np2 == 256) warp_sort_to<8>(xs, xsorted, Ni, lane);
    else sorted_in_regs = false;
  }
  if (!sorted_in_regs) {
    for (int k = lane; k < np2; k += 32) xsorted[k] = (k < Ni) ? xs[k] : __int_as_float(0x7fffffff);   // pad key: sorts after everything
    __syncwarp();
    auto gt = [](float p, float q) {   // total order: numbers < NaNs (like torch.sort) < pad keys
      const bool pn = (p != p), qn = (q != q);
      if (pn || qn) return pn && (!qn || __float_as_int(p) > __float_as_int(q));
      return p > q;
    };
    for (int size = 2; size <= np2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = lane; t < (np2 >> 1); t += 32) {
          const int lo = 2 * t - (t & (stride - 1));       // index with the `stride` bit clear
          const int hi = lo + stride;
          const bool up = ((lo & size) == 0);
          const float p = xsorted[lo], q = xsorted[hi];
          if (gt(p, q) == up) { xsorted[lo] = q; xsorted[hi] = p; }
        }
        __syncwarp();
      }
    }
  }
  for (int i = lane; i < S; i += 32) {  // coarse i lands after every sample strictly smaller
    const float v = zc[i];
    int lo = 0, hi = Ni;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (xsorted[mid] < v) lo = mid + 1; else hi = mid; }
    out[i + lo] = v;
  }
  for (int k = lane; k < Ni; k += 32) {  // sample k lands after every coarse depth <= it
    const float v = xsorted[k];
    int lo = 0, hi = S;
    if (v == v) { while (lo < hi) { const int mid = (lo + hi) >> 1; if (zc[mid] <= v) lo = mid + 1; else hi = mid; } }
    else lo = S;
    out[k + lo] = v;
  }
}

struct MergeArgs {
  const float *z, *samples, *rays; int64_t n; int stride, S, Ni; float *z_out, *z_std;
};

__global__ void __launch_bounds__(128) k_merge(MergeArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (r >= a.n) return;
  const int S = a.S, Ni = a.Ni;
  int np2 = 32;                      // sort scratch: a power of two >= max(N_importance, 32) (the register sort writes whole warps)
  while (np2 < Ni) np2 <<= 1;
  float* zc = smem + (size_t)wib * (S + Ni + np2);
  float* xs = zc + S;
  float* xsorted = xs + Ni;
  const float near = a.rays[r * a.stride + 6], far = a.rays[r * a.stride + 7];
  for (int k = lane; k < S; k += 32) zc[k] = a.z[r * (int64_t)S + k];
  for (int k = lane; k < Ni; k += 32) xs[k] = clamp_sample(a.samples[r * (int64_t)Ni + k], near, far);
  __syncwarp();
  merge_ray(zc, xs, xsorted, S, Ni, np2, lane, a.z_out + r * (int64_t)(S + Ni), a.z_std ? a.z_std + r : nullptr);
}

int launch_merge(const float* z, const float* samples, const float* rays, int64_t n, int stride, int S,
                 int Ni, float* z_out, float* z_std, cudaStream_t st) {
  if (n == 0) return PLNERF_OK;
  MergeArgs a{z, samples, rays, n, stride, S, Ni, z_out, z_std};
  const int wpb = 4;
  int np2 = 32;                      // sort scratch: a power of two >= max(N_importance, 32) (the register sort writes whole warps)
  while (np2 < Ni) np2 <<= 1;
  const size_t smem = (size_t)wpb * (S + Ni + np2) * sizeof(float);
  if (smem > 48 * 1024) { set_error("merge: S=%d Ni=%d too large", S, Ni); return PLNERF_E_UNSUPPORTED; }
  k_merge<<<(unsigned)ceil_div(n, wpb), wpb * 32, smem, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_merge");
  return PLNERF_OK;
}

// ---- fused: importance sampling + clamp + sort-merge + z_std of render_rays (run_plnerf.py:721-734, :752) in one kernel.
// The Ni samples stay in shared memory between the inverse-CDF step and the merge (the unfused pair writes and re-reads
// [n, Ni] floats and launches twice).  LINEAR: sample_pdf_reformulation on (z, weights, tau, T); otherwise sample_pdf on
// bins = z_mid, weights[..., 1:-1].
struct SampleMergeArgs {
  SamplePLArgs pl; SampleConstArgs cs;
  const float *z, *rays; int64_t n; int stride, S, Ni;
  float *z_out, *z_std;
};

template <bool LINEAR>
__global__ void __launch_bounds__(128) k_sample_merge(const __grid_constant__ SampleMergeArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (r >= a.n) return;
  const int S = a.S, Ni = a.Ni, nk = S + 2;
  int np2 = 32;                      // sort scratch: a power of two >= max(N_importance, 32) (the register sort writes whole warps)
  while (np2 < Ni) np2 <<= 1;
  const int per_warp = 4 * nk + S + Ni + np2;
  float* cdf = smem + (size_t)wib * per_warp;
  float *s = cdf + nk, *T = s + nk, *tau = T + nk;     // constant mode: s = bins (nb = S - 1), T / tau unused
  float* zc = tau + nk;
  float* xs = zc + S;
  float* xsorted = xs + Ni;
  const float near = a.rays[r * a.stride + 6], far = a.rays[r * a.stride + 7];
  for (int k = lane; k < S; k += 32) zc[k] = a.z[r * (int64_t)S + k];
  // device-side draws: one Philox block yields the draws of four consecutive samples (philox_uniform(k) = word k & 3 of block
  // k >> 2); they wait in xs[], where sample k's lane replaces its own draw by the sample
  const float* u_drawn = nullptr;
  if (!a.pl.u) {
    for (int b = lane; 4 * b < Ni; b += 32) {
      uint32_t w4[4];
      philox4x32(a.pl.seed, a.pl.ray0 + (uint64_t)r, RNG_STREAM_U, (uint32_t)b, w4);
#pragma unroll
      for (int i = 0; i < 4; ++i) if (4 * b + i < Ni) xs[4 * b + i] = (float)(w4[i] >> 8) * (1.0f / 16777216.0f);
    }
    u_drawn = xs;
  }
  if (LINEAR) {
    pl_build(a.pl, r, lane, cdf, s, T, tau);       // (ends with __syncwarp: the draws are visible)
    for (int k = lane; k < Ni; k += 32) {
      const float x = pl_sample(a.pl, r, k, cdf, s, T, tau, u_drawn);
      if (a.pl.samples) a.pl.samples[r * (int64_t)Ni + k] = x;
      xs[k] = clamp_sample(x, near, far);
    }
  } else {
    const_build(a.cs, r, lane, cdf, s);
    for (int k = lane; k < Ni; k += 32) {
      const float x = const_sample(a.cs, r, k, cdf, s, u_drawn);
      if (a.cs.samples) a.cs.samples[r * (int64_t)Ni + k] = x;
      xs[k] = clamp_sample(x, near, far);
    }
  }
  __syncwarp();
  merge_ray(zc, xs, xsorted, S, Ni, np2, lane, a.z_out + r * (int64_t)(S + Ni), a.z_std ? a.z_std + r : nullptr);
}

int launch_sample_merge(int linear, const float* z, const float* w, const float* tau, const float* T, const float* rays, int64_t n,
                        int stride, int S, int Ni, const float* u, uint64_t seed, uint64_t ray0, float zero_tol, float eps,
                        float* z_out, float* z_std, int64_t* inds, cudaStream_t st) {
  if (n == 0) return PLNERF_OK;
  SampleMergeArgs a;
  memset(&a, 0, sizeof(a));
  a.pl = SamplePLArgs{z, w, tau, T, rays, n, stride, S, Ni, u, seed, ray0, zero_tol, eps, nullptr, inds, nullptr, nullptr, nullptr, nullptr};
  // constant mode (run_plnerf.py:726): bins = z_mid [S-1] from the z rows, weights[..., 1:-1] (row stride S)
  a.cs = SampleConstArgs{z, S, 1, w ? w + 1 : nullptr, S, n, S - 1, Ni, u, seed, ray0, nullptr, inds, nullptr};
  a.z = z; a.rays = rays; a.n = n; a.stride = stride; a.S = S; a.Ni = Ni; a.z_out = z_out; a.z_std = z_std;
  const int wpb = 4;
  int np2 = 32;                      // sort scratch: a power of two >= max(N_importance, 32) (the register sort writes whole warps)
  while (np2 < Ni) np2 <<= 1;
  const size_t smem = (size_t)wpb * (4 * (S + 2) + S + Ni + np2) * sizeof(float);
  if (smem > 48 * 1024) { set_error("sample+merge: S=%d Ni=%d too large", S, Ni); return PLNERF_E_UNSUPPORTED; }
  if (linear) k_sample_merge<true><<<(unsigned)ceil_div(n, wpb), wpb * 32, smem, st>>>(a);
  else {
    if (S < 3) { set_error("sample_pdf: need at least 2 bins"); return PLNERF_E_BADARG; }
    k_sample_merge<false><<<(unsigned)ceil_div(n, wpb), wpb * 32, smem, st>>>(a);
  }
  PLNERF_LAUNCH_CHECK("k_sample_merge");
  return PLNERF_OK;
}

int launch_encode(const float* x, int64_t n, int L, float* out, cudaStream_t st) {
  if (n == 0) return PLNERF_OK;
  k_encode<<<(unsigned)ceil_div(n * 3, 256), 256, 0, st>>>(x, n, L, out);
  PLNERF_LAUNCH_CHECK("k_encode");
  return PLNERF_OK;
}

int launch_stratified_z(const float* rays, int64_t n, int stride, int Ns, int lindisp, int perturb,
                        const float* t_rand, uint64_t seed, uint64_t ray0, float* z, cudaStream_t st) {
  if (n == 0) return PLNERF_OK;
  k_stratified_z<<<(unsigned)ceil_div(n * Ns, 256), 256, 0, st>>>(rays, n, stride, Ns, lindisp, perturb, t_rand,
                                                                   seed, ray0, z);
  PLNERF_LAUNCH_CHECK("k_stratified_z");
  return PLNERF_OK;
}

// =============================================================================================
// f-1: ray generation + packing.  One thread per ray does what render() spends ~15 full-image torch ops on
// (run_plnerf.py:138-164): get_rays from a camera pose (run_nerf_helpers.py:162-171) or a given (o, d) pair,
// viewdirs = d / |d| taken BEFORE the NDC warp (:145-150), ndc_rays (:184-201), and the [o, d, near, far, viewdir]
// row.  Every operation is a separately rounded fp32 op in the reference's order (no FMA contraction).
// =============================================================================================
struct PackRaysArgs {
  int H, W;
  float fx, fy, cx, cy;         // K[0][0], K[1][1], K[0][2], K[1][2] rounded to fp32 like torch does with python scalars
  const float* c2w; int c2w_ld;         // [3, >=4] pose (rows strided by c2w_ld) or null
  const float* c2w_static; int c2w_static_ld;   // optional: origins/directions from this pose, viewdirs from c2w
  const float* rays_o; const float* rays_d;     // [n,3] each, used when c2w == null
  const int64_t* pix;           // optional [n] flat pixel ids (row*W + col) with a pose: rays of those pixels only; null = all H*W in order
  int64_t n;
  int ndc, use_viewdirs;
  float ndc_cx, ndc_cy;         // fl32(-1/(W/(2 focal))), fl32(-1/(H/(2 focal)))  (python float64 arithmetic, then fp32)
  float ndc_near;               // the near plane passed to ndc_rays (1.0 in render())
  float near, far;
  float* out; int stride;       // [n, 8 | 11]
};

__device__ __forceinline__ void pose_ray(const float* c2w, int ld, float dx, float dy, float dz, float (&o)[3], float (&d)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // torch.sum(dirs[..., None, :] * c2w[:3,:3], -1): three rounded products, summed left to right
    d[c] = __fadd_rn(__fadd_rn(__fmul_rn(dx, c2w[c * ld + 0]), __fmul_rn(dy, c2w[c * ld + 1])), __fmul_rn(dz, c2w[c * ld + 2]));
    o[c] = c2w[c * ld + 3];
  }
}

__global__ void __launch_bounds__(256) k_pack_rays(const PackRaysArgs a) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n) return;
  float o[3], d[3], vd[3] = {0.f, 0.f, 0.f};
  if (a.c2w) {
    const int64_t px = a.pix ? a.pix[r] : r;
    const float i = (float)(px % a.W), j = (float)(px / a.W);        // torch.linspace(0, W-1, W) is exact on integers
    const float dx = __fdiv_rn(__fsub_rn(i, a.cx), a.fx);
    const float dy = -__fdiv_rn(__fsub_rn(j, a.cy), a.fy);
    pose_ray(a.c2w, a.c2w_ld, dx, dy, -1.f, o, d);
    if (a.use_viewdirs) { vd[0] = d[0]; vd[1] = d[1]; vd[2] = d[2]; }
    if (a.c2w_static) pose_ray(a.c2w_static, a.c2w_static_ld, dx, dy, -1.f, o, d);
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) { o[c] = a.rays_o[r * 3 + c]; d[c] = a.rays_d[r * 3 + c]; }
    if (a.use_viewdirs) { vd[0] = d[0]; vd[1] = d[1]; vd[2] = d[2]; }
  }
  if (a.use_viewdirs) {
    // torch.norm(dim=-1): sqrt of the fp32 sum of squares
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vd[0], vd[0]), __fmul_rn(vd[1], vd[1])), __fmul_rn(vd[2], vd[2])));
#pragma unroll
    for (int c = 0; c < 3; ++c) vd[c] = __fdiv_rn(vd[c], nrm);
  }
  if (a.ndc) {
    const float t = __fdiv_rn(-__fadd_rn(a.ndc_near, o[2]), d[2]);
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = __fadd_rn(o[c], __fmul_rn(t, d[c]));
    const float o0 = __fdiv_rn(__fmul_rn(a.ndc_cx, o[0]), o[2]);
    const float o1 = __fdiv_rn(__fmul_rn(a.ndc_cy, o[1]), o[2]);
    const float o2 = __fadd_rn(1.f, __fdiv_rn(__fmul_rn(2.f, a.ndc_near), o[2]));
    const float d0 = __fmul_rn(a.ndc_cx, __fsub_rn(__fdiv_rn(d[0], d[2]), __fdiv_rn(o[0], o[2])));
    const float d1 = __fmul_rn(a.ndc_cy, __fsub_rn(__fdiv_rn(d[1], d[2]), __fdiv_rn(o[1], o[2])));
    const float d2 = __fdiv_rn(__fmul_rn(-2.f, a.ndc_near), o[2]);
    o[0] = o0; o[1] = o1; o[2] = o2; d[0] = d0; d[1] = d1; d[2] = d2;
  }
  float* row = a.out + r * (int64_t)a.stride;
  row[0] = o[0]; row[1] = o[1]; row[2] = o[2]; row[3] = d[0]; row[4] = d[1]; row[5] = d[2];
  row[6] = a.near; row[7] = a.far;
  if (a.use_viewdirs) { row[8] = vd[0]; row[9] = vd[1]; row[10] = vd[2]; }
}

int launch_pack_rays(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int c2w_ld,
                     const float* c2w_static, int c2w_static_ld, const float* rays_o, const float* rays_d,
                     const int64_t* pix, int64_t n,
                     int ndc, float ndc_cx, float ndc_cy, float ndc_near, float near, float far, int use_viewdirs,
                     float* out, int stride, cudaStream_t st) {
  if (n == 0) return PLNERF_OK;
  PackRaysArgs a;
  a.H = H; a.W = W; a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.c2w = c2w; a.c2w_ld = c2w_ld;
  a.c2w_static = c2w_static; a.c2w_static_ld = c2w_static_ld; a.rays_o = rays_o; a.rays_d = rays_d; a.pix = pix; a.n = n;
  a.ndc = ndc; a.use_viewdirs = use_viewdirs; a.ndc_cx = ndc_cx; a.ndc_cy = ndc_cy; a.ndc_near = ndc_near;
  a.near = near; a.far = far; a.out = out; a.stride = stride;
  k_pack_rays<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_pack_rays");
  return PLNERF_OK;
}

// ---------------------------------------------------------------------------------------------
// f-2: the training loop's loss and optimiser steps (run_plnerf.py:1289-1315)
// ---------------------------------------------------------------------------------------------
// img2mse twice + the start of loss.backward() (run_plnerf.py:1289-1297, run_nerf_helpers.py:17): d = rgb - target,
// g = scale * d (scale = 2 / (3 B) for the mean over the global batch) for the fine and the coarse map, and the two sums of
// squares ADDED to sqerr[0:2].  One block: the reduction order is fixed (the loss is reproducible bit for bit); the target
// row of ray i is target[pix[i]] when pixel ids are given (the gather target_s = target[select_coords], :1280).
__global__ void __launch_bounds__(1024) k_mse_loss_grad(const float* __restrict__ rgb, const float* __restrict__ rgb0,
                                                        const float* __restrict__ target, const int64_t* __restrict__ pix,
                                                        int64_t n, float scale, float* __restrict__ g_rgb,
                                                        float* __restrict__ g_rgb0, float* __restrict__ sqerr) {
  __shared__ float red[2][32];
  float s = 0.f, s0 = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float* t = target + 3 * (pix ? pix[i] : i);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float tc = t[c];
      const float d = rgb[3 * i + c] - tc;
      g_rgb[3 * i + c] = d * scale;
      s = fmaf(d, d, s);
      if (rgb0) {
        const float d0 = rgb0[3 * i + c] - tc;
        g_rgb0[3 * i + c] = d0 * scale;
        s0 = fmaf(d0, d0, s0);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s0 += __shfl_xor_sync(0xffffffffu, s0, o); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s; red[1][warp] = s0; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    s = lane < nw ? red[0][lane] : 0.f;
    s0 = lane < nw ? red[1][lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s0 += __shfl_xor_sync(0xffffffffu, s0, o); }
    if (lane == 0) { sqerr[0] += s; if (rgb0) sqerr[1] += s0; }
  }
}

int launch_mse_loss_grad(const float* rgb, const float* rgb0, const float* target, const int64_t* pix, int64_t n, float scale,
                         float* g_rgb, float* g_rgb0, float* sqerr, cudaStream_t st) {
  if (n == 0) return PLNERF_OK;
  k_mse_loss_grad<<<1, 1024, 0, st>>>(rgb, rgb0, target, pix, n, scale, g_rgb, g_rgb0, sqerr);
  PLNERF_LAUNCH_CHECK("k_mse_loss_grad");
  return PLNERF_OK;
}

// torch.optim.Adam (amsgrad=False, weight_decay=0; run_plnerf.py:431-447, :1302-1303) over ONE flat fp32 segment instead of
// 24 tensors per network: exp_avg <- lerp(exp_avg, g, 1 - beta1); exp_avg_sq <- beta2 exp_avg_sq + (1 - beta2) g^2;
// p <- p - step_size * exp_avg / (sqrt(exp_avg_sq) / sqrt(bias_correction2) + eps).  Optionally clears the gradient (the
// optimizer.zero_grad() of the next iteration, :1286).
__global__ void __launch_bounds__(256) k_adam_flat(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                   float4* __restrict__ v, int64_t n4, float* __restrict__ pt,
                                                   float* __restrict__ gt, float* __restrict__ mt, float* __restrict__ vt,
                                                   int n_tail, float w1, float beta2, float w2, float step_size,
                                                   float bc2_sqrt, float eps, int zero_grads) {
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    mm = mm + w1 * (gg - mm);
    vv = beta2 * vv + w2 * gg * gg;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp = pp - step_size * (mm / denom);
  };
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = g[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (zero_grads) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (i - n4 < n_tail) {
    const int k = (int)(i - n4);
    float pp = pt[k], mm = mt[k], vv = vt[k];
    upd(pp, gt[k], mm, vv);
    pt[k] = pp; mt[k] = mm; vt[k] = vv;
    if (zero_grads) gt[k] = 0.f;
  }
}

int launch_adam_flat(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                     double beta2, double eps, int64_t step, int zero_grads, cudaStream_t st) {
  if (n == 0) return PLNERF_OK;
  // scalar terms in double, like torch's fused kernel
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  // 16-byte vector body when the four arrays are 16-byte aligned (flat buffers are), scalar tail; otherwise all scalar
  const bool vec = (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0, tail = n - 4 * n4;
  if (tail >= ((int64_t)1 << 31)) { set_error("adam_step: unaligned segments above 2^31 elements are not supported"); return PLNERF_E_UNSUPPORTED; }
  const int64_t o = 4 * n4;
  k_adam_flat<<<(unsigned)ceil_div(n4 + tail, 256), 256, 0, st>>>(
      reinterpret_cast<float4*>(params), reinterpret_cast<float4*>(grads), reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), n4, params + o, grads + o, exp_avg + o, exp_avg_sq + o, (int)tail,
      (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), step_size, bc2_sqrt, (float)eps, zero_grads);
  PLNERF_LAUNCH_CHECK("k_adam_flat");
  return PLNERF_OK;
}

}  // namespace plnerf
