// Quadrature of one ray by one warp (raw2outputs + compute_weights{,_piecewise_linear}, reference run_plnerf.py:504-624):
// shared by the standalone kernel k_composite (ops.cu) and by the compositor warps of the fused MLP kernel (mlp_fwd3.cuh),
// so the two produce bit-identical maps.  Lanes are striped along the sample axis; prefix products are warp-shuffle scans
// in fp64 rounded per element to fp32 (what torch's CPU cumprod produces).
#pragma once
#include <math.h>

#include "common.cuh"

namespace plnerf {

__device__ __forceinline__ float sigmoidf_(float x) { return __fdiv_rn(1.0f, 1.0f + expf(-x)); }

struct CompositeArgs {
  const float* raw; int raw_stride;
  const float* z; const float* rays; int64_t n; int stride; int S;
  int color_mode, white_bkgd, farcolorfix;
  const float* noise;       // explicit additive noise [n,S] or null
  float noise_std;          // >0 with noise == null: Philox normal * std
  uint64_t seed, ray0; uint32_t noise_stream;
  float *rgb_map, *disp_map, *acc_map, *depth_map, *weights, *tau, *T;
};

// raw values of the ray's samples: global memory [S, raw_stride] ...
struct RawGlobal {
  const float* p; int stride;
  __device__ __forceinline__ float operator()(int k, int c) const { return p[(int64_t)k * stride + c]; }
};
// ... or a shared-memory ring of float4 rows (the fused kernel): sample k of the ray sits at ring row (first + k) mod rows
struct RawRing {
  const float4* ring; int first, rows;
  __device__ __forceinline__ float operator()(int k, int c) const {
    int i = first + k;
    if (i >= rows) i -= rows;
    return reinterpret_cast<const float*>(ring + i)[c];
  }
};

// One warp, ray r (all 32 lanes call it converged).  `raw` yields raw[k][c] of the ray's samples, z its S depths.
template <int MODE, class Raw>
__device__ __forceinline__ void composite_ray(const CompositeArgs& a, int64_t r, int lane, const Raw raw, const float* z) {
  const int S = a.S;
  const float* ray = a.rays + r * a.stride;
  const float dx = ray[3], dy = ray[4], dz = ray[5];
  const float near = ray[6], far = ray[7];
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));

  auto sigma_at = [&](int k) -> float {  // density of sample k incl. noise, before relu
    float s = raw(k, 3);
    if (a.noise) s = __fadd_rn(s, a.noise[r * (int64_t)S + k]);
    else if (a.noise_std > 0.f)
      s = __fadd_rn(s, __fmul_rn(philox_normal(a.seed, a.ray0 + (uint64_t)r, a.noise_stream, (uint32_t)k), a.noise_std));
    return s;
  };
  auto color_at = [&](int k, int c) -> float { return sigmoidf_(raw(k, c)); };

  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_w = 0.f;
  double carry = 1.0;

  if (MODE == PLNERF_MODE_LINEAR) {
    // knots s = [near, z_0..z_{S-1}, far] (S+2), tau = relu([1e-10, sigma.., 1e10]); S+1 intervals
    auto knot = [&](int k) -> float { return k == 0 ? near : (k == S + 1 ? far : z[k - 1]); };
    auto tau_at = [&](int k) -> float {
      return k == 0 ? 1e-10f : (k == S + 1 ? 1e10f : fmaxf(sigma_at(k - 1), 0.0f));
    };
    const int nI = S + 1;
    if (lane == 0 && a.T) a.T[r * (int64_t)(S + 2)] = 1.0f;
    for (int base = 0; base < nI; base += 32) {
      const int i = base + lane;
      const bool valid = i < nI;
      float e = 1.0f, s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
      if (valid) {
        s0 = knot(i); s1 = knot(i + 1);
        t0 = tau_at(i); t1 = tau_at(i + 1);
        const float dist = __fmul_rn(__fsub_rn(s1, s0), dnorm);
        const float ave = __fmul_rn(0.5f, __fadd_rn(t1, t0));
        e = expf(__fmul_rn(-ave, dist));
      }
      const double incl = warp_incl_scan_mul((double)e, lane) * carry;
      double excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = carry;
      carry = __shfl_sync(0xffffffffu, incl, 31);
      if (valid) {
        const float Ti = (float)excl;
        const float w = __fmul_rn(__fsub_rn(1.0f, e), Ti);
        if (a.weights) a.weights[r * (int64_t)(S + 1) + i] = w;
        if (a.T) a.T[r * (int64_t)(S + 2) + i + 1] = (float)incl;
        if (a.tau) {
          a.tau[r * (int64_t)(S + 2) + i] = t0;
          if (i == S) a.tau[r * (int64_t)(S + 2) + S + 1] = t1;
        }
        // colours: cc = [c_0, c_0..c_{S-1}, c_{S-1}|0] (midpoint) or [c_0, c_0..c_{S-1}] (left)
        const int kl = max(i - 1, 0);
        const int kr = min(i, S - 1);
        float cr, cg, cb;
        if (a.color_mode == PLNERF_COLOR_MIDPOINT) {
          const bool zero_right = a.farcolorfix && (i == S);
          const float lr = color_at(kl, 0), lg = color_at(kl, 1), lb = color_at(kl, 2);
          float rr = 0.f, rg = 0.f, rb = 0.f;
          if (!zero_right) {
            if (kr == kl) { rr = lr; rg = lg; rb = lb; }
            else { rr = color_at(kr, 0); rg = color_at(kr, 1); rb = color_at(kr, 2); }
          }
          cr = __fmul_rn(0.5f, __fadd_rn(rr, lr));
          cg = __fmul_rn(0.5f, __fadd_rn(rg, lg));
          cb = __fmul_rn(0.5f, __fadd_rn(rb, lb));
        } else {
          cr = color_at(kl, 0); cg = color_at(kl, 1); cb = color_at(kl, 2);
        }
        acc_r += w * cr; acc_g += w * cg; acc_b += w * cb;
        acc_d += w * __fmul_rn(0.5f, __fadd_rn(s1, s0));
        acc_w += w;
      }
    }
  } else {
    // constant: dists = [z_{i+1}-z_i, 1e10]*|d|, alpha = 1-exp(-relu(sigma)*dist),
    // w = alpha * cumprod([1, 1-alpha+1e-10])[:-1]
    for (int base = 0; base < S; base += 32) {
      const int i = base + lane;
      const bool valid = i < S;
      float alpha = 0.f, om = 1.0f, zi = 0.f;
      if (valid) {
        zi = z[i];
        const float d0 = (i < S - 1) ? __fsub_rn(z[i + 1], zi) : 1e10f;
        const float dist = __fmul_rn(d0, dnorm);
        const float sg = fmaxf(sigma_at(i), 0.0f);
        alpha = __fsub_rn(1.0f, expf(__fmul_rn(-sg, dist)));
        om = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      }
      const double incl = warp_incl_scan_mul((double)om, lane) * carry;
      double excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = carry;
      carry = __shfl_sync(0xffffffffu, incl, 31);
      if (valid) {
        const float w = __fmul_rn(alpha, (float)excl);
        if (a.weights) a.weights[r * (int64_t)S + i] = w;
        acc_r += w * color_at(i, 0); acc_g += w * color_at(i, 1); acc_b += w * color_at(i, 2);
        acc_d += w * zi;
        acc_w += w;
      }
    }
  }
  acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b);
  acc_d = warp_sum(acc_d); acc_w = warp_sum(acc_w);
  if (lane == 0) {
    const float q = __fdiv_rn(acc_d, acc_w);
    const float disp = (q != q) ? q : __fdiv_rn(1.0f, fmaxf(1e-10f, q));
    if (a.white_bkgd) {
      const float bg = __fsub_rn(1.0f, acc_w);
      acc_r += bg; acc_g += bg; acc_b += bg;
    }
    if (a.rgb_map) { a.rgb_map[r * 3 + 0] = acc_r; a.rgb_map[r * 3 + 1] = acc_g; a.rgb_map[r * 3 + 2] = acc_b; }
    if (a.disp_map) a.disp_map[r] = disp;
    if (a.acc_map) a.acc_map[r] = acc_w;
    if (a.depth_map) a.depth_map[r] = acc_d;
  }
}

}  // namespace plnerf
