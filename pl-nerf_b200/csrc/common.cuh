// Shared helpers for the plnerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/plnerf_b200.h"

namespace plnerf {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);

#define PLNERF_CHECK_ARG(cond, ...)                  \
  do {                                               \
    if (!(cond)) {                                   \
      plnerf::set_error(__VA_ARGS__);                \
      return PLNERF_E_BADARG;                        \
    }                                                \
  } while (0)

#define PLNERF_CUDA(call)                                            \
  do {                                                               \
    cudaError_t _e = (call);                                         \
    if (_e != cudaSuccess) return plnerf::cuda_fail(_e, #call);      \
  } while (0)

#define PLNERF_LAUNCH_CHECK(name)                                    \
  do {                                                               \
    plnerf::count_launch();                                          \
    cudaError_t _e = cudaGetLastError();                             \
    if (_e != cudaSuccess) return plnerf::cuda_fail(_e, name);       \
  } while (0)

// ---- Philox4x32-10: counter-based RNG keyed by (seed, ray id, stream, sample/4) ----------------
// Draws are a pure function of (seed, global ray index, stream id, sample index) so results do
// not depend on chunking or on how rays are sharded over GPUs.
enum { RNG_STREAM_TRAND = 0, RNG_STREAM_U = 1, RNG_STREAM_NOISE0 = 2, RNG_STREAM_NOISE1 = 3 };

__host__ __device__ inline void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                             uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__host__ __device__ inline void philox4x32(uint64_t seed, uint64_t ray, uint32_t stream, uint32_t blk,
                                           uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ray, c1 = (uint32_t)(ray >> 32), c2 = stream, c3 = blk;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in [0,1): 24 random bits
__host__ __device__ inline float philox_uniform(uint64_t seed, uint64_t ray, uint32_t stream, uint32_t idx) {
  uint32_t r[4];
  philox4x32(seed, ray, stream, idx >> 2, r);
  return (float)(r[idx & 3] >> 8) * (1.0f / 16777216.0f);
}

#ifdef __CUDACC__
// standard normal via Box-Muller on two 24-bit uniforms
__device__ inline float philox_normal(uint64_t seed, uint64_t ray, uint32_t stream, uint32_t idx) {
  uint32_t r[4];
  philox4x32(seed, ray, stream, idx >> 1, r);
  const int h = (idx & 1) * 2;
  const float u1 = ((float)(r[h] >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
  const float u2 = (float)(r[h + 1] >> 8) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inclusive warp scans in fp64 (the CPU reference accumulates cumsum/cumprod in fp64 and rounds
// each prefix to fp32)
__device__ __forceinline__ double warp_incl_scan_add(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ double warp_incl_scan_mul(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
#endif

#ifdef __CUDACC__
// ---- positional-encoding angles --------------------------------------------------------------------
// The encoding needs sin/cos of p * 2^k for k = 0..L-1 (run_nerf_helpers.py:45-48).  Scaling by 2^k is exact in
// fp32, so the reference's argument is the real number p * 2^k.  pe_turns() converts p once to a 32-bit fixed-point
// fraction of a full turn, U = frac(p / 2pi) * 2^32 (64x24-bit integer product: 2^-32 turns absolute error); the angle of octave k
// is then the EXACT left shift U << k (mod 1 turn), and pe_sin / pe_cos evaluate it with the SFU on [-pi, pi), where
// sin.approx / cos.approx are accurate to 2^-21.4 absolute.  Total error <= ~1e-6 absolute at k = 9 (shifted
// conversion error 2^-23 turns + SFU), i.e. the fp32-reference encoding to within the 1e-4 parity budget of the
// network outputs, at ~6 instructions per sin/cos pair instead of ~60 for two precise sinf/cosf calls.
__device__ __forceinline__ uint32_t pe_turns(float p) {
  // integer-only: p = +-m * 2^(e-150); frac(|p| / 2pi) * 2^32 = (m * C) >> (182 - e) mod 2^32 with C = 2^64 / (2 pi)
  // (truncation error < 2^-32 turns).  No fp64: the double-precision multiply / 64-bit convert pair costs this kernel's
  // 16 epilogue warps ~3000 cycles per tile on B200.
  const uint32_t bits = __float_as_uint(p);
  uint32_t e = (bits >> 23) & 255u;
  const uint32_t m = (bits & 0x7FFFFFu) | (e ? 0x800000u : 0u);
  if (e == 0) e = 1;
  const unsigned long long C = 0x28BE60DB9391054AULL;
  const unsigned long long lo = (unsigned long long)m * C;          // low 64 bits of the 88-bit product
  const unsigned long long hi = __umul64hi((unsigned long long)m, C);
  const int sh = 182 - (int)e;                                       // right shift of the 128-bit value hi:lo
  uint32_t u;
  if (sh >= 96) u = 0u;
  else if (sh >= 64) u = (uint32_t)(hi >> (sh - 64));
  else if (sh > 0) u = (uint32_t)((lo >> sh) | (hi << (64 - sh)));
  else u = (uint32_t)(lo << (-sh));                                  // |p| >= 2^32: only the low product bits matter
  return (bits >> 31) ? (0u - u) : u;
}
__device__ __forceinline__ float pe_angle(uint32_t turns, int k) {
  return (float)(int32_t)(turns << k) * 1.4629180792671596e-9f;   // 2 pi / 2^32: radians in [-pi, pi)
}
__device__ __forceinline__ float pe_sin(uint32_t turns, int k) { return __sinf(pe_angle(turns, k)); }
__device__ __forceinline__ float pe_cos(uint32_t turns, int k) { return __cosf(pe_angle(turns, k)); }

// Elements [16*G, 16*G + 16) of the 3 + 6*L wide encoding [x, y, z, sin(2^0 p), cos(2^0 p), sin(2^1 p), ...] of one point,
// fully unrolled for a compile-time group G (no dynamic indexing, no integer division, one angle per sin/cos pair that
// falls into the group): out[e] = element 16*G + e, zero beyond n_ch.
// Per coordinate, the LOWEST octave the group touches is evaluated exactly (shifted turn fraction + SFU); the group's
// higher octaves (at most 3 more) follow from the double-angle recurrence sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a
// on the FMA pipe -- the SFU / conversion pipe (16 lanes per clock per SM) is what bounds the encoding.  Each doubling
// multiplies the absolute error by 3-4: <= 2e-5 at the end of a group -- invisible after the bf16 rounding of the
// operand (RECUR = true, bf16 modes) but not for the hi+lo split mode, which evaluates every octave exactly (RECUR = false).
template <int G, bool RECUR>
__device__ __forceinline__ void pe_group16(const float (&p)[3], const uint32_t (&turns)[3], int n_ch, float (&out)[16]) {
  constexpr int i_lo = (16 * G < 3) ? 3 : 16 * G, i_hi = 16 * G + 15;
  constexpr int k_lo = (i_lo - 3) / 6, k_hi = (i_hi - 3) / 6;                 // octaves touched by elements [i_lo, i_hi]
  float sn[3][k_hi - k_lo + 1], cs[3][k_hi - k_lo + 1];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // first octave whose sin or cos of coordinate c falls into the group
    const int k0 = (3 + 6 * k_lo + 3 + c >= i_lo) ? k_lo : k_lo + 1;   // cos(c, k_lo) index = 3 + 6 k_lo + 3 + c
#pragma unroll
    for (int k = k_lo; k <= k_hi; ++k) {
      if (k < k0) { sn[c][k - k_lo] = 0.f; cs[c][k - k_lo] = 0.f; }
      else if (k == k0 || !RECUR) {
        const float ang = pe_angle(turns[c], k);
        sn[c][k - k_lo] = __sinf(ang); cs[c][k - k_lo] = __cosf(ang);
      } else {
        const float s1 = sn[c][k - 1 - k_lo], c1 = cs[c][k - 1 - k_lo];
        sn[c][k - k_lo] = 2.f * s1 * c1;
        cs[c][k - k_lo] = fmaf(-2.f * s1, s1, 1.f);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int idx = 16 * G + e;
    float v;
    if (idx < 3) {
      v = p[idx];
    } else {
      const int t = idx - 3, k = t / 6, r = t % 6, c = r % 3;      // compile-time after unrolling
      v = (r >= 3) ? cs[c][k - k_lo] : sn[c][k - k_lo];
    }
    out[e] = (idx < n_ch) ? v : 0.f;
  }
}
template <bool RECUR>
__device__ __forceinline__ void pe_group16(int g, const float (&p)[3], const uint32_t (&turns)[3], int n_ch, float (&out)[16]) {
  switch (g) {
    case 0: pe_group16<0, RECUR>(p, turns, n_ch, out); break;
    case 1: pe_group16<1, RECUR>(p, turns, n_ch, out); break;
    case 2: pe_group16<2, RECUR>(p, turns, n_ch, out); break;
    default: pe_group16<3, RECUR>(p, turns, n_ch, out); break;
  }
}
#endif

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace plnerf
