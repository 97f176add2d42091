// Bring-up / measurement kernels of the tcgen05 path.  Compiled ONLY into the developer library (-DPLNERF_DEBUG,
// libplnerf_b200_debug.so): the product library carries no debug entry points, switches or environment knobs.
// Included by mlp_fwd.cu inside namespace plnerf (after the fused kernels, whose helpers it uses).
#pragma once

namespace {

// =============================================================================================
// debug: single tile GEMM  D[128,N] = A[128,K] * B[N,K]^T  through the same primitives
// (N in {128,256}, K % 16 == 0, K <= 256).  a_mode 0 = A from shared memory panels (SS),
// 1 = A from tensor memory (TS).  lbo/sbo are passed explicitly so tests can pin the encoding.
// =============================================================================================
__global__ void __launch_bounds__(128, 1) k_debug_gemm(const float* __restrict__ Ag, const float* __restrict__ Bg, int N, int K,
                                                       int a_mode, uint32_t lbo, uint32_t sbo, float* __restrict__ Dg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: A panels [K/8][128 rows][16B] | B panels per 128-row half: [half][K/8][128][16B] | barrier | tmem slot
  const int kp = K / 8;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)kp * 2048;
  const int nh = N / 128;
  uint8_t* sBar = sB + (size_t)nh * kp * 2048;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sBar + 16);
  const uint32_t bar = ptx::smem_u32(sBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
  // B -> smem panels (generic-proxy stores + proxy fence)
  for (int idx = threadIdx.x; idx < N * kp; idx += 128) {
    const int n = idx % N, p = idx / N;
    uint32_t w[4];
    for (int e = 0; e < 4; ++e) w[e] = ptx::pack_bf16(Bg[(size_t)n * K + p * 8 + 2 * e], Bg[(size_t)n * K + p * 8 + 2 * e + 1]);
    *reinterpret_cast<uint4*>(sB + ((size_t)(n / 128) * kp + p) * 2048 + (n % 128) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;
  // A -> smem panels or TMEM columns [256, 256+K/2)
  if (a_mode == 0) {
    for (int p = 0; p < kp; ++p) {
      uint32_t w[4];
      for (int e = 0; e < 4; ++e) w[e] = ptx::pack_bf16(Ag[(size_t)row * K + p * 8 + 2 * e], Ag[(size_t)row * K + p * 8 + 2 * e + 1]);
      *reinterpret_cast<uint4*>(sA + (size_t)p * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = 2 * (c0 + i);
        pk[i] = (k < K) ? ptx::pack_bf16(Ag[(size_t)row * K + k], Ag[(size_t)row * K + k + 1]) : 0u;
      }
      ptx::tmem_st16(tmem + lane_addr + 256u + (uint32_t)c0, pk);
    }
    ptx::tmem_st_wait();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::idesc_bf16_f32(128, 128);
    for (int h = 0; h < nh; ++h) {
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t bd = ptx::smem_desc(ptx::smem_u32(sB + ((size_t)h * kp + 2 * ks) * 2048), lbo, sbo);
        if (a_mode == 0) ptx::mma_ss(tmem + 128u * h, ptx::smem_desc(ptx::smem_u32(sA + (size_t)(2 * ks) * 2048), lbo, sbo), bd, idesc, ks > 0);
        else ptx::mma_ts(tmem + 128u * h, tmem + 256u + 8u * ks, bd, idesc, ks > 0);
      }
    }
    ptx::mma_commit(bar);
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  for (int c = 0; c < N / 32; ++c) {
    uint32_t r[32];
    ptx::tmem_ld32(tmem + lane_addr + 32u * c, r);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) Dg[(size_t)row * N + c * 32 + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// =============================================================================================
// debug: MN-major operands (what the weight-gradient GEMM uses).  D[128, N] = sum_k X[k, m] Y[k, n],
// X [K,128] and Y [K,N] row-major fp32 (so the contraction index k is the strided one).  Operands are
// staged as no-swizzle MN-major core matrices: block (mn8, k8) = 8 k-rows x 8 mn-values (mn fastest,
// 128 contiguous bytes) at ((mn8 * K/8) + k8) * 128; descriptor SBO = (K/8)*128 (MN direction),
// LBO = 128 (K direction).
// =============================================================================================
__global__ void __launch_bounds__(128, 1) k_debug_gemm_mn(const float* __restrict__ Xg, const float* __restrict__ Yg, int N, int K,
                                                          uint32_t lbo, uint32_t sbo, float* __restrict__ Dg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int k8n = K / 8;
  uint8_t* sA = smem;                                  // 16 mn8 blocks x k8n x 128 B
  uint8_t* sB = smem + (size_t)16 * k8n * 128;         // N/8 blocks x k8n x 128 B
  uint8_t* sBar = sB + (size_t)(N / 8) * k8n * 128;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sBar + 16);
  const uint32_t bar = ptx::smem_u32(sBar);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
  for (int idx = threadIdx.x; idx < K * 128; idx += 128) {
    const int k = idx / 128, m = idx % 128;
    reinterpret_cast<__nv_bfloat16*>(sA + ((size_t)(m / 8) * k8n + k / 8) * 128)[(k % 8) * 8 + (m % 8)] = __float2bfloat16_rn(Xg[idx]);
  }
  for (int idx = threadIdx.x; idx < K * N; idx += 128) {
    const int k = idx / N, n = idx % N;
    reinterpret_cast<__nv_bfloat16*>(sB + ((size_t)(n / 8) * k8n + k / 8) * 128)[(k % 8) * 8 + (n % 8)] = __float2bfloat16_rn(Yg[idx]);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::idesc_bf16_f32_mn(128, N);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = ptx::smem_desc(ptx::smem_u32(sA) + ks * 256, lbo, sbo);
      const uint64_t bd = ptx::smem_desc(ptx::smem_u32(sB) + ks * 256, lbo, sbo);
      ptx::mma_ss(tmem, ad, bd, idesc, ks > 0);
    }
    ptx::mma_commit(bar);
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;
  for (int c = 0; c < N / 32; ++c) {
    uint32_t r[32];
    ptx::tmem_ld32(tmem + lane_addr + 32u * c, r);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) Dg[(size_t)threadIdx.x * N + c * 32 + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// =============================================================================================
// debug: raw tcgen05.mma issue/execute rate.  mode 0: TS N=128, 1: TS N=256, 2: SS N=128, 3: SS N=256.
// One CTA per SM issues `iters` x 16 back-to-back MMAs on garbage operands; reports cycles per MMA.
// =============================================================================================
__global__ void __launch_bounds__(640, 1) k_debug_mma_rate(int mode, int iters, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 160 * 1024);
  const uint32_t bar = ptx::smem_u32(smem + 160 * 1024 + 16);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar + 8, 1); ptx::mbar_init(bar + 16, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp >= 4) {
    // mode 13/14: extra warps polling an mbarrier (what the epilogue warps of the fused kernels do while they wait)
    ptx::mbar_wait(bar + 16, 0);
  } else if (mode >= 20) {
    // CUDA-core conversion throughput (all 4 warps): 20 = cvt.rn.relu.bf16x2.f32, 21 = max + integer round-half-up + PRMT,
    // 22 = packed fp32 add (baseline), 23 = cvt.rn.bf16x2.f32 (no relu).  Reports cycles per warp-level "pair" operation.
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (float)(threadIdx.x * 32 + i) * 1.0001f - 1000.f;
    uint32_t sink = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t d;
        if (mode == 20) {
          d = pack_bf16_relu(v[2 * i], v[2 * i + 1]);
        } else if (mode == 23) {
          d = ptx::pack_bf16(v[2 * i], v[2 * i + 1]);
        } else if (mode == 21) {
          const uint32_t a = __float_as_uint(fmaxf(v[2 * i], 0.f)) + 0x8000u, b = __float_as_uint(fmaxf(v[2 * i + 1], 0.f)) + 0x8000u;
          d = __byte_perm(a, b, 0x7632);
        } else {
          float x = v[2 * i], y = v[2 * i + 1];
          add2(x, y, 1.5f, 2.5f);
          d = __float_as_uint(x) ^ __float_as_uint(y);
        }
        sink ^= d;
        v[2 * i] = __uint_as_float(__float_as_uint(v[2 * i]) ^ (d & 1u));   // keep the chain data-dependent but cheap
      }
    }
    const long long t1 = clock64();
    if (sink == 0x12345678u) cycles_out[0] = 0;
    if (threadIdx.x == 32) cycles_out[blockIdx.x] = t1 - t0;
  } else if (warp == 1) {
    const int N = ((mode & 1) && mode < 4) ? 256 : 128;
    const bool ss = (mode == 2 || mode == 3);
    const uint32_t idesc = ptx::idesc_bf16_f32(128, N);
    const uint32_t sb = ptx::smem_u32(smem);
    const uint32_t lbo = N * 16;
    const uint32_t bar2 = bar + 8;   // second barrier for the per-batch commit experiments
    long long t0 = clock64(), t_issue = 0;
    if (mode < 4) {
      for (int it = 0; it < iters; ++it) {
        if (ptx::elect_one()) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint64_t bd = ptx::smem_desc(sb + 64 * 1024 + (j & 3) * (N * 32), lbo, 128);
            if (ss) ptx::mma_ss(tmem, ptx::smem_desc(sb + j * 4096, 2048, 128), bd, idesc, j > 0);
            else ptx::mma_ts(tmem, tmem + 256u + 8u * j, bd, idesc, j > 0);
          }
        }
        __syncwarp();
      }
    } else if (mode >= 6 && mode <= 14) {
      // SS-form operand-layout experiments (rate only, operands are garbage):
      //  6: N=256 no swizzle   7: N=256 A+B SWIZZLE_128B   8: N=256 A swizzled only   9: N=256 B swizzled only
      // 10: N=128 A+B SWIZZLE_128B   11: N=256 no swizzle, A fixed (same 4 KB every MMA)   12: N=256 no swizzle, B fixed
      const int N = (mode == 10) ? 128 : 256;
      const bool a_sw = (mode == 7 || mode == 8 || mode == 10), b_sw = (mode == 7 || mode == 9 || mode == 10);
      const uint32_t idesc = ptx::idesc_bf16_f32(128, N);
      const uint32_t sb = ptx::smem_u32(smem);
      auto desc = [&](uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) -> uint64_t {
        return ptx::smem_desc(addr, lbo, sbo) | ((uint64_t)layout << 61);
      };
      for (int it = 0; it < iters; ++it) {
        if (ptx::elect_one()) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int ja = (mode == 11) ? 0 : j, jb = (mode == 12) ? 0 : (j & 3);
            const uint64_t ad = a_sw ? desc(sb + (ja >> 2) * 16384 + (ja & 3) * 32, 16, 1024, 2) : desc(sb + ja * 4096, 2048, 128, 0);
            const uint64_t bd = b_sw ? desc(sb + 64 * 1024 + jb * 32, 16, 1024, 2) : desc(sb + 64 * 1024 + jb * (N * 32), N * 16, 128, 0);
            ptx::mma_ss(tmem, ad, bd, idesc, j > 0);
          }
        }
        __syncwarp();
      }
    } else if (mode == 4) {
      // like the kernel's stage loop: 8 MMAs, commit to a barrier, wait for the PREVIOUS batch's barrier
      for (int it = 0; it < iters * 2; ++it) {
        if (ptx::elect_one()) {
          issue_ts8<false>(tmem, tmem + 256u, tmem + 384u, ptx::smem_desc(sb + 64 * 1024, 2048, 128), idesc, 1);
          ptx::mma_commit(bar2);
        }
        __syncwarp();
        if (it > 0) ptx::mbar_wait(bar2, (it - 1) & 1);
        ptx::tc_fence_after();
      }
      ptx::mbar_wait(bar2, (iters * 2 - 1) & 1);
    } else {
      // mode 5: how far ahead of the tensor pipe does the issuing thread run?  (queue depth)
      for (int it = 0; it < iters; ++it) {
        long long a0 = clock64();
        if (ptx::elect_one()) {
          issue_ts8<false>(tmem, tmem + 256u, tmem + 384u, ptx::smem_desc(sb + 64 * 1024, 2048, 128), idesc, 1);
          issue_ts8<false>(tmem, tmem + 256u, tmem + 384u, ptx::smem_desc(sb + 64 * 1024, 2048, 128), idesc, 1);
        }
        __syncwarp();
        long long a1 = clock64();
        if (ptx::elect_one()) ptx::mma_commit(bar2);
        __syncwarp();
        ptx::mbar_wait(bar2, it & 1);
        t_issue += a1 - a0;
      }
    }
    if (ptx::elect_one()) ptx::mma_commit(bar);
    __syncwarp();
    ptx::mbar_wait(bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 32) cycles_out[blockIdx.x] = (mode == 5) ? t_issue : (t1 - t0);
    if (threadIdx.x == 32) ptx::mbar_arrive(bar + 16);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

}  // namespace

int debug_umma_gemm_ex(const float* A, const float* B, int N, int K, int a_mode, uint32_t lbo, uint32_t sbo, float* D, cudaStream_t st) {
  PLNERF_CHECK_ARG(A && B && D, "debug_umma_gemm: null argument");
  PLNERF_CHECK_ARG((N == 128 || N == 256) && K % 16 == 0 && K >= 16 && K <= 256, "debug_umma_gemm: N in {128,256}, K%%16==0, K<=256");
  int rc = query_device();
  if (rc) return rc;
  const size_t smem = (size_t)(K / 8) * 2048 * (1 + N / 128) + 64;
  PLNERF_CUDA(cudaFuncSetAttribute(k_debug_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
  k_debug_gemm<<<1, 128, smem, st>>>(A, B, N, K, a_mode, lbo, sbo, D);
  PLNERF_LAUNCH_CHECK("k_debug_gemm");
  return PLNERF_OK;
}

int debug_set_trace(long long* buf) { g_trace = buf; return PLNERF_OK; }

int debug_umma_gemm_mn(const float* X, const float* Y, int N, int K, uint32_t lbo, uint32_t sbo, float* D, cudaStream_t st) {
  PLNERF_CHECK_ARG(X && Y && D, "debug_umma_gemm_mn: null argument");
  PLNERF_CHECK_ARG(N % 32 == 0 && N >= 32 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 128, "debug_umma_gemm_mn: bad N/K");
  int rc = query_device();
  if (rc) return rc;
  const size_t smem = (size_t)(16 + N / 8) * (K / 8) * 128 + 64;
  PLNERF_CUDA(cudaFuncSetAttribute(k_debug_gemm_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
  k_debug_gemm_mn<<<1, 128, smem, st>>>(X, Y, N, K, lbo, sbo, D);
  PLNERF_LAUNCH_CHECK("k_debug_gemm_mn");
  return PLNERF_OK;
}

int debug_mma_rate(int mode, int iters, int grid, long long* cycles_out, cudaStream_t st) {
  int rc = query_device();
  if (rc) return rc;
  PLNERF_CUDA(cudaFuncSetAttribute(k_debug_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
  const int threads = (mode == 13) ? 640 : (mode == 14 ? 256 : 128);   // 13: 16 polling warps, 14: 4 polling warps
  k_debug_mma_rate<<<grid, threads, 160 * 1024 + 64, st>>>(mode, iters, cycles_out);
  PLNERF_LAUNCH_CHECK("k_debug_mma_rate");
  return PLNERF_OK;
}

int debug_umma_gemm(const float* A, const float* B, int N, int K, float* D, cudaStream_t st) {
  return debug_umma_gemm_ex(A, B, N, K, 0, 2048, 128, D, st);
}

