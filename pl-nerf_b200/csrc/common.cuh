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

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace plnerf
