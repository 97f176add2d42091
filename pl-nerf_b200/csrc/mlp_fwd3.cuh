// k_mlp3 -- bf16 inference generation 3 of the fused PE + MLP kernel (included by mlp_fwd.cu).
//
// Why: k_mlp_fwd keeps ONE 128-row tile in flight, so the tensor pipe and the epilogue warps wait for each other once
// per layer (layer period ~2800 cycles against 2080 cycles of MMA).  k_mlp3 keeps TWO tiles ("X" and "Y") in flight at
// the SAME layer and alternates them on the tensor pipe half by half:
//
//      pipe:      X(l,a)  Y(l,a)  X(l,b)  Y(l,b)  X(l+1,a) ...
//      epilogue:          X(l,a)  Y(l,a)  X(l,b)  Y(l,b)   ...
//
// so every half-epilogue (tcgen05.ld, bias, ReLU, bf16 pack, tcgen05.st) hides behind the other tile's 16 MMAs, and both
// tiles consume every weight stage (L2 -> shared-memory traffic per row halves).
//
// Tensor memory (512 columns): tile t owns  D_t = [256t, 256t+128)  one fp32 accumulator HALF (128 neurons)
//                                           A_t = [256t+128, 256t+256)  the layer input, 128x256 bf16 (TS-form A operand).
// A_t is single-buffered: the bf16 output of half a cannot overwrite it while half b's MMAs still read it, so the
// epilogue threads HOLD half a (16 packed registers per thread) and store both halves after half b completed.
// Layers that do not read A_t (layer 0: the encoding is a shared-memory operand) store directly.
//
// Warp roles (20 warps): 0-7 epilogue of tile X, 8-15 epilogue of tile Y (TMEM lane quarter = warp % 4, column half =
// (warp / 4) % 2: one row x 64 accumulator columns per thread) -- two independent instruction streams, so X's and Y's
// epilogues overlap each other as well as the MMAs; 16 / 17 = MMA issuer of tile X / Y (16 allocates TMEM, 17 also
// streams the weights: one bulk-TMA refill per stage it issues); 18-19 =
// helpers that encode the NEXT pair's rows (one thread per row: the angle reduction is done once per row, not once
// per column group) and pull the pair after that into L2.
//
// Rows -> CTAs: every CTA owns a CONTIGUOUS range of whole rays (units of vb_div rows), tiled from its own first row
// (the last tile of a CTA may be partial), so that all samples of a ray pass through one CTA in order.  With
// A.fuse_comp the quadrature (raw2outputs, composite.cuh) runs inside the kernel: the views epilogue leaves each row's
// four outputs in a shared-memory ring (3 tile pairs deep) and the helper warps composite every ray whose last sample
// has arrived -- `raw` goes to HBM only when the caller asks for it (retraw).
//
// Weight ring: N slots of one 32 KB stage (8 K-steps of a 128-neuron half); both issuers walk it, a slot is refilled
// once both have released it (tcgen05.commit on a count-2 barrier).  Biases are added in the epilogue (a bias
// K-step would cost 6% of the tensor pipe, which is the bound here).
namespace v3 {

constexpr int MAX_SLOTS = 6;
constexpr int EPI_WARPS_PER_TILE = 8;
constexpr int WARP_MMA3 = 2 * EPI_WARPS_PER_TILE, WARP_HELP3 = WARP_MMA3 + 2, N_HELP_WARPS = 2;   // issuers: WARP_MMA3 + tile
constexpr int THREADS3 = 32 * (WARP_HELP3 + N_HELP_WARPS);   // 640
constexpr uint32_t VB_RAYS_SMEM = 3;
constexpr int RING_PAIRS = 3, RING_ROWS = RING_PAIRS * 2 * TILE_M;   // output ring: a ray of <= 256 samples spans <= 2 pairs       // rays whose view-bias rows a 128-sample tile can touch when S >= 64

// Shared-memory map: everything the epilogue warps address sits at COMPILE-TIME offsets (their addresses are instruction
// immediates, not registers -- the epilogue threads hold two tiles' state in a 96-register budget); the variable-size
// parts (fp32 constants, weight ring) come last.
template <bool VD>
struct Smem3 {
  static constexpr uint32_t bars = 0;                                    // 256 B of mbarriers + the TMEM base slot
  static constexpr uint32_t prog = 256;                                  // <= 48 stages per tile: [2][48] entries of 16 B, then the stream table [48] x 8 B
  static constexpr uint32_t stab = prog + 16u * 2u * 48u;
  static constexpr uint32_t vb0 = stab + 8u * 48u;                 // [2][3 rays][128] view-bias rows
  // VD: ring of the rows' four outputs (float4), RING_PAIRS tile pairs deep -- also the exchange buffer of the two column
  // halves' head partial sums; !VD: [2][128 rows][MAX_OUT_CH] exchange buffer only
  static constexpr uint32_t XCH_BYTES = VD ? RING_ROWS * 16u : 2u * TILE_M * MAX_OUT_CH * 4u;
  static constexpr uint32_t xch0 = vb0 + 2u * VB_RAYS_SMEM * 128u * 4u;
  static constexpr uint32_t ctrs = xch0 + XCH_BYTES;                       // flow-control counters (see the kernel)
  // per-layer facts the epilogue needs, one 16-byte entry per layer {epi | halves << 8 | flags << 16 | reads_a << 24,
  // bias offset, stash byte offset, mask word offset (-1: none)}: ONE shared-memory load per layer instead of chains of
  // register-indexed constant-bank loads (plan -> index -> offset table), which sat in the epilogue's critical path
  static constexpr uint32_t ltab = ctrs + 64u;
  static constexpr uint32_t pe0 = (ltab + 16u * MAX_LAYERS + 1023u) & ~1023u;   // [2] encoding operand tiles
  static constexpr uint32_t consts = pe0 + 2u * PE_TILE_BYTES;           // fp32 biases + heads (const_floats)
  uint32_t ring, total;
  int n_slots;
};
template <bool VD>
__host__ __device__ inline Smem3<VD> smem3_layout(int n_slots, int const_floats) {
  Smem3<VD> s;
  s.n_slots = n_slots;
  s.ring = (Smem3<VD>::consts + (uint32_t)const_floats * 4u + 1023u) & ~1023u;
  s.total = s.ring + (uint32_t)n_slots * STAGE_BYTES;
  return s;
}

// Program entry flags (see issue_tile3)
enum : uint32_t { PF_FIRST = 1u, PF_SS = 2u, PF_COMMIT = 4u, PF_WAIT_EP = 8u, PF_WAIT_PE = 16u, PF_WAIT_EF = 32u };

// Per-issuer state carried across tile pairs: consumer ring cursor + phase, parities of the tile's epilogue / encoding
// barriers; producer duty (tile Y's issuer only): ring cursor + phase, next stage of the weight stream, stages left.
struct Issue3 { uint32_t sl, ph, c, q, psl, pph, pj, prem; };   // c: bit 0 / 1 = parity of the early / full epilogue barrier

#define PLNERF3_MMA_TS(PRED) \
  "tcgen05.mma.cta_group::1.kind::f16 [ed], [ea], bd, %9, " PRED ";\n\t" \
  "add.u64 bd, bd, 256;\n\tadd.u32 ea, ea, 8;\n\t"
#if defined(PLNERF_DEBUG) && defined(PLNERF_ENABLE_TRACE)
#define PLNERF3_STAMP "@ptr mov.u32 t, %%clock;\n\t@ptr st.global.u32 [tr], t;\n\t@ptr add.u64 tr, tr, 4;\n\t"
#else
#define PLNERF3_STAMP ""
#endif

// One tile's whole stage program of a pair from one PTX loop (an issuing warp retires one dependent instruction per ~6.5
// cycles: a stage's bookkeeping has to stay at a few dozen instructions, and ONE issuing thread cannot feed the pipe for
// two tiles -- measured 730 cycles per 8-MMA stage against 520 cycles of tensor work -- hence one issuer warp per tile).
// Entry (16 bytes): {accumulator tmem address, A operand (tmem address | low word of the shared-memory descriptor),
// flags, accumulator-full barrier}.  flags: PF_FIRST first MMA overwrites the accumulator, PF_SS shared-memory A operand
// with bits 8-11 K-steps, PF_COMMIT commit the accumulator-full barrier after the stage, PF_WAIT_EP wait for the tile's
// "early" epilogue barrier before the stage (accumulator drained, first half of the layer input stored; + PF_WAIT_PE: and
// for its encoding operand), PF_WAIT_EF wait for the "full" epilogue barrier (whole layer input stored) -- both barriers
// advance one phase per half-epilogue, so one parity bit per barrier tracks them.  Every stage waits for its ring slot's
// weights and releases the slot with a commit (the slot's empty barrier counts both tiles' issuers).
// Producer duty (prem > 0): after each stage, refill the ring slot released `lag` stages ago with the next stage of the
// packed stream ({global byte offset, bytes} table at stab_addr, n_stab entries per pair).
__device__ __forceinline__ void issue_tile3(Issue3& st, uint32_t prog_addr, uint32_t n_entries, uint64_t ring_desc, uint32_t idesc,
                                            uint32_t desc_hi, uint32_t wfull0, uint32_t wempty0, uint32_t epearly, uint32_t epfull, uint32_t peready,
                                            uint32_t n_slots, uint32_t stab_addr, uint32_t n_stab, uint32_t ring_addr,
                                            unsigned long long wbase, unsigned long long trace_ptr) {
  const uint64_t wpolicy = ptx::l2_policy_evict_last();      // the weight stream stays in L2 (every CTA re-reads it per tile pair)
  asm volatile(
      "{\n\t.reg .pred p, pacc, pt, ptr;\n\t"
      ".reg .b32 sl, ph, n, pa, ed, ea, ef, eb, t, k, wb, c, q, psl, pph, pj, prem, so2, sb2;\n\t.reg .b64 bd, so, ad, tr, ga;\n\t"
      "mov.b64 tr, %19;\n\tsetp.ne.b64 ptr, tr, 0;\n\t"
      "mov.b32 sl, %0;\n\tmov.b32 ph, %1;\n\tmov.b32 c, %2;\n\tmov.b32 q, %3;\n\t"
      "mov.b32 psl, %4;\n\tmov.b32 pph, %5;\n\tmov.b32 pj, %6;\n\tmov.b32 prem, %7;\n\t"
      "mov.b32 n, %11;\n\tmov.b32 pa, %10;\n\tsetp.eq.b32 pt, sl, sl;\n\t"
      "LOOP3:\n\t"
      "ld.shared.v4.u32 {ed, ea, ef, eb}, [pa];\n\t"
      PLNERF3_STAMP
      // ---- the ring slot's weights (they land long before the activations: checked first, off the dependency chain)
      "shl.b32 wb, sl, 3;\n\t"
      "add.u32 t, wb, %12;\n\t"
      "WAQ3:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [t], ph;\n\t@!p bra WAQ3;\n\t"
      // ---- the tile's previous half-epilogue: early barrier (accumulator drained, first input half stored), the
      // encoding operand, full barrier (whole layer input stored)
      "and.b32 t, ef, 8;\n\tsetp.eq.b32 p, t, 0;\n\t@p bra NOEP3;\n\t"
      "and.b32 k, c, 1;\n\t"
      "WEP3:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%14], k;\n\t@!p bra WEP3;\n\t"
      "xor.b32 c, c, 1;\n\t"
      "and.b32 t, ef, 16;\n\tsetp.eq.b32 p, t, 0;\n\t@p bra NOEP3;\n\t"
      "WPE3:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%15], q;\n\t@!p bra WPE3;\n\t"
      "xor.b32 q, q, 1;\n\t"
      "NOEP3:\n\t"
      "and.b32 t, ef, 32;\n\tsetp.eq.b32 p, t, 0;\n\t@p bra NOEF3;\n\t"
      "shr.u32 k, c, 1;\n\t"
      "WEF3:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%23], k;\n\t@!p bra WEF3;\n\t"
      "xor.b32 c, c, 2;\n\t"
      "NOEF3:\n\t"
      "tcgen05.fence::after_thread_sync;\n\t"
      PLNERF3_STAMP
      "mul.wide.u32 so, sl, 2048;\n\tadd.u64 bd, so, %8;\n\t"
      "and.b32 t, ef, 1;\n\tsetp.eq.b32 pacc, t, 0;\n\t"
      "and.b32 t, ef, 2;\n\tsetp.ne.b32 p, t, 0;\n\t@p bra SS3;\n\t"
      PLNERF3_MMA_TS("pacc") PLNERF3_MMA_TS("pt") PLNERF3_MMA_TS("pt") PLNERF3_MMA_TS("pt")
      PLNERF3_MMA_TS("pt") PLNERF3_MMA_TS("pt") PLNERF3_MMA_TS("pt") PLNERF3_MMA_TS("pt")
      "bra DONE3;\n\t"
      "SS3:\n\t"
      "mov.b64 ad, {ea, %13};\n\tshr.u32 k, ef, 8;\n\tand.b32 k, k, 15;\n\t"
      "SSL3:\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [ed], ad, bd, %9, pacc;\n\t"
      "setp.eq.b32 pacc, sl, sl;\n\tadd.u64 ad, ad, 256;\n\tadd.u64 bd, bd, 256;\n\t"
      "sub.u32 k, k, 1;\n\tsetp.ne.b32 p, k, 0;\n\t@p bra SSL3;\n\t"
      "DONE3:\n\t"
      "add.u32 t, wb, %16;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [t];\n\t"
      "and.b32 t, ef, 4;\n\tsetp.ne.b32 p, t, 0;\n\t"
      "@p tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [eb];\n\t"
      PLNERF3_STAMP
      "add.u32 sl, sl, 1;\n\tsetp.eq.u32 p, sl, %17;\n\t@p mov.b32 sl, 0;\n\t@p xor.b32 ph, ph, 1;\n\t"
      // ---- producer duty: refill the slot both tiles released a while ago with the next stage of the stream
      "setp.eq.u32 p, prem, 0;\n\t@p bra NOPROD3;\n\t"
      "shl.b32 wb, psl, 3;\n\tadd.u32 t, wb, %16;\n\txor.b32 k, pph, 1;\n\t"
      "WPR3:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [t], k;\n\t@!p bra WPR3;\n\t"
      "shl.b32 t, pj, 3;\n\tadd.u32 t, t, %18;\n\tld.shared.v2.u32 {so2, sb2}, [t];\n\t"
      "add.u32 t, wb, %12;\n\t"
      "mbarrier.arrive.expect_tx.shared::cta.b64 _, [t], sb2;\n\t"
      "cvt.u64.u32 ga, so2;\n\tadd.u64 ga, ga, %20;\n\t"
      "shl.b32 k, psl, 15;\n\tadd.u32 k, k, %21;\n\t"
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [k], [ga], sb2, [t], %24;\n\t"
      "add.u32 psl, psl, 1;\n\tsetp.eq.u32 p, psl, %17;\n\t@p mov.b32 psl, 0;\n\t@p xor.b32 pph, pph, 1;\n\t"
      "add.u32 pj, pj, 1;\n\tsetp.eq.u32 p, pj, %22;\n\t@p mov.b32 pj, 0;\n\tsub.u32 prem, prem, 1;\n\t"
      "NOPROD3:\n\t"
      "add.u32 pa, pa, 16;\n\tsub.u32 n, n, 1;\n\tsetp.ne.b32 p, n, 0;\n\t@p bra LOOP3;\n\t"
      "mov.b32 %0, sl;\n\tmov.b32 %1, ph;\n\tmov.b32 %2, c;\n\tmov.b32 %3, q;\n\t"
      "mov.b32 %4, psl;\n\tmov.b32 %5, pph;\n\tmov.b32 %6, pj;\n\tmov.b32 %7, prem;\n\t}"
      : "+r"(st.sl), "+r"(st.ph), "+r"(st.c), "+r"(st.q), "+r"(st.psl), "+r"(st.pph), "+r"(st.pj), "+r"(st.prem)
      : "l"(ring_desc), "r"(idesc), "r"(prog_addr), "r"(n_entries), "r"(wfull0), "r"(desc_hi), "r"(epearly), "r"(peready),
        "r"(wempty0), "r"(n_slots), "r"(stab_addr), "l"(trace_ptr), "l"(wbase), "r"(ring_addr), "r"(n_stab), "r"(epfull), "l"(wpolicy)
      : "memory");
}

// Full encoding of one row (all 3 + 6L elements, one thread): p = o + d*z, one 32-bit turn fraction per coordinate, every
// octave an exact shift of it + SFU sin/cos (common.cuh) -> the row's 16-byte units of the tile's PE operand (bf16,
// K-major 8x16B core-matrix panels: panel j = columns [8j, 8j+8) at j * 2048 + row * 16).
template <bool STASH>
__device__ __forceinline__ void pe_write_row(const MlpArgs& A, uint8_t* pe_tile, int64_t g, int row, int64_t row_end) {
  const NetPlan& P = A.plan;
  const int64_t gc = (g < row_end) ? g : (row_end - 1);
  const int n_panels = 2 * P.pe_ks;
  // training: the encoding (and the ray's padded direction encoding) also go to the stash as weight-gradient operand
  // tiles; tiles are globally aligned in this mode (row0 is a multiple of 256)
  uint8_t* st_in = STASH ? A.in_stash + (g / TILE_M) * (int64_t)A.tl.in_tile_bytes : nullptr;
  if (STASH) {
    const float* dp = A.dirpe + (gc / A.vb_div) * 32;
    uint8_t* tile_dir = st_in + A.tl.in_off[A.tl.idx_dir];
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      uint4 q4;
      q4.x = ptx::pack_bf16(dp[8 * c8 + 0], dp[8 * c8 + 1]); q4.y = ptx::pack_bf16(dp[8 * c8 + 2], dp[8 * c8 + 3]);
      q4.z = ptx::pack_bf16(dp[8 * c8 + 4], dp[8 * c8 + 5]); q4.w = ptx::pack_bf16(dp[8 * c8 + 6], dp[8 * c8 + 7]);
      stash_store8(tile_dir, 32, row, c8, q4);
    }
  }
  if (A.x_emb) {
    // pre-embedded rows (NeRF.forward entry)
    const float* xr = A.x_emb + gc * (int64_t)A.x_ld;
    for (int j = 0; j < n_panels; ++j) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { const int idx = 8 * j + e; v[e] = (idx < P.input_ch) ? xr[idx] : 0.f; }
      uint4 q4;
      q4.x = ptx::pack_bf16(v[0], v[1]); q4.y = ptx::pack_bf16(v[2], v[3]); q4.z = ptx::pack_bf16(v[4], v[5]); q4.w = ptx::pack_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(pe_tile + j * 2048 + row * 16) = q4;
    }
    return;
  }
  float p[3];
  const float* rp = A.rays + (gc / A.S) * (int64_t)A.stride;
  const float zz = A.z[gc];
#pragma unroll
  for (int c = 0; c < 3; ++c) p[c] = __fadd_rn(rp[c], __fmul_rn(rp[3 + c], zz));   // o + d*z, two roundings like the reference
  const uint32_t turns[3] = {pe_turns(p[0]), pe_turns(p[1]), pe_turns(p[2])};
  const int n_ch = P.input_ch;
#pragma unroll
  for (int j = 0; j < 2 * MAX_PE_KS; ++j) {
    if (j < n_panels) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int idx = 8 * j + e;                              // compile-time after unrolling
        float x;
        if (idx < 3) x = p[idx];
        else { const int tt = idx - 3, k = tt / 6, r = tt % 6, c = r % 3; x = (r >= 3) ? pe_cos(turns[c], k) : pe_sin(turns[c], k); }
        v[e] = (idx < n_ch) ? x : 0.f;
      }
      uint4 q4;
      q4.x = ptx::pack_bf16(v[0], v[1]); q4.y = ptx::pack_bf16(v[2], v[3]); q4.z = ptx::pack_bf16(v[4], v[5]); q4.w = ptx::pack_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(pe_tile + j * 2048 + row * 16) = q4;
      if (STASH) stash_store8(st_in + A.tl.in_off[A.tl.idx_pe], A.tl.in_width[A.tl.idx_pe], row, j, q4);
    }
  }
}

// bias + (ReLU) + bf16 pack of 32 accumulator columns -> 16 packed words
template <bool RELU>
__device__ __forceinline__ void cvt32(uint32_t (&r)[32], const float* bias, uint32_t* pk, uint32_t* mask = nullptr) {
  float* val = reinterpret_cast<float*>(r);
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias + i);
    add2(val[i], val[i + 1], b4.x, b4.y);
    add2(val[i + 2], val[i + 3], b4.z, b4.w);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = RELU ? pack_bf16_relu(val[2 * i], val[2 * i + 1]) : ptx::pack_bf16(val[2 * i], val[2 * i + 1]);
  if (RELU && mask) *mask = relu_mask_from_packed(pk);     // training: the ReLU mask of these 32 columns
}
__device__ __forceinline__ float dot32_relu(const uint32_t (&r)[32], const float* w, float acc) {
  const float* val = reinterpret_cast<const float*>(r);
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 w4 = *reinterpret_cast<const float4*>(w + i);
    acc = fmaf(fmaxf(val[i], 0.f), w4.x, acc); acc = fmaf(fmaxf(val[i + 1], 0.f), w4.y, acc);
    acc = fmaf(fmaxf(val[i + 2], 0.f), w4.z, acc); acc = fmaf(fmaxf(val[i + 3], 0.f), w4.w, acc);
  }
  return acc;
}

// VD: use_viewdirs network (alpha + rgb heads behind a views layer) or not (output_linear behind the last trunk layer)
#ifdef PLNERF_DEBUG
#define PLNERF3_DBG(bit) ((A.debug_flags & (bit)) != 0)   // bring-up experiments: 1 = no epilogue TMEM traffic / math, 2 = 16-byte weight copies, 4 = no encoding, 8 = epilogue signals before it works (no dependency chain, results invalid)
#else
#define PLNERF3_DBG(bit) false
#endif
// STASH (training forward, VD only): every layer's bf16 activations + ReLU masks also go to the training stash (the
// weight-gradient operand tiles k_mlp_fwd<2> writes; same values, same layout)
template <bool VD, bool STASH = false>
__global__ void __launch_bounds__(THREADS3, 1) k_mlp3(const __grid_constant__ MlpArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const NetPlan& P = A.plan;
  const Smem3<VD> SL = smem3_layout<VD>(A.n_stages, P.const_floats);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = ptx::smem_u32(smem);
  float* consts = reinterpret_cast<float*>(smem + Smem3<VD>::consts);
  const uint32_t s_bars = sbase + Smem3<VD>::bars;
  const uint32_t w_full0 = s_bars, w_empty0 = s_bars + 8u * MAX_SLOTS;
  const uint32_t d_full0 = s_bars + 8u * (2 * MAX_SLOTS);         // [2] accumulator half of tile t is complete
  const uint32_t ep_done0 = d_full0 + 16u;                        // [2] tile t's half-epilogue is done (8 warps)
  const uint32_t pe_ready0 = d_full0 + 32u;                       // [2] tile t's encoding operand is written (helper warps)
  const uint32_t vb_full0 = d_full0 + 48u;                        // [2] tile t's view-bias rows landed (bulk copy)
  const uint32_t pe_free0 = d_full0 + 64u;                        // [2] every MMA reading tile t's encoding operand is complete
  const uint32_t ep_early0 = d_full0 + 80u;                       // [2] tile t's accumulator is drained and the first half of its layer input stored
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Smem3<VD>::bars + 8u * (2 * MAX_SLOTS) + 112u);
  const int n_slots = SL.n_slots;

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_slots; ++s) { ptx::mbar_init(w_full0 + 8u * s, 1); ptx::mbar_init(w_empty0 + 8u * s, 2); }
    for (uint32_t t = 0; t < 2; ++t) {
      ptx::mbar_init(d_full0 + 8u * t, 1);
      ptx::mbar_init(ep_done0 + 8u * t, EPI_WARPS_PER_TILE);
      ptx::mbar_init(pe_ready0 + 8u * t, N_HELP_WARPS);
      ptx::mbar_init(vb_full0 + 8u * t, 1);
      ptx::mbar_init(pe_free0 + 8u * t, EPI_WARPS_PER_TILE);
      ptx::mbar_init(ep_early0 + 8u * t, EPI_WARPS_PER_TILE);
    }
    ptx::fence_mbar_init();
    for (int i = 0; i < 16; ++i) reinterpret_cast<uint32_t*>(smem + Smem3<VD>::ctrs)[i] = 0u;
  }
  if ((int)threadIdx.x < P.n_layers) {
    const LayerPlan& L = P.L[threadIdx.x];
    int4 e;
    e.x = (int)L.epi | ((int)L.n_halves << 8) | ((int)(uint8_t)L.flags << 16) | ((L.n_h_ks > 0 ? 1 : 0) << 24);
    e.y = L.bias_off;
    e.z = (STASH && L.stash_idx >= 0) ? A.tl.in_off[L.stash_idx] : 0;
    e.w = (STASH && L.mask_idx >= 0) ? A.tl.mask_off[L.mask_idx] : -1;
    reinterpret_cast<int4*>(smem + Smem3<VD>::ltab)[threadIdx.x] = e;
  }
  if (warp == WARP_MMA3) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < P.const_floats; i += THREADS3) consts[i] = A.tail[i];
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // this CTA's contiguous range of whole units (rays of vb_div rows): rows [row0, row_end), tiled from row0
  const int64_t n_units = (A.M + A.unit_rows - 1) / A.unit_rows;     // unit = a ray (fused quadrature) or a tile pair
  const int64_t unit0 = (int64_t)blockIdx.x * A.cta_units;
  const int64_t unit1 = (unit0 + A.cta_units < n_units) ? unit0 + A.cta_units : n_units;
  const int64_t row0 = unit0 * A.unit_rows;
  const int64_t row_end = (unit1 * A.unit_rows < A.M) ? unit1 * A.unit_rows : A.M;
  const int n_pairs = (row_end > row0) ? (int)((row_end - row0 + 2 * TILE_M - 1) / (2 * TILE_M)) : 0;
  // flow control of the output ring (plain counters: the parties may run several pairs apart, which mbarrier parities
  // cannot express): tiles_done[t] += 1 per final-row warp and pair (4 per pair), comp_done += 1 per helper warp and pair
  volatile uint32_t* ctrs = reinterpret_cast<volatile uint32_t*>(smem + Smem3<VD>::ctrs);
  int l_pe_last = 0;
  for (int l = 0; l < P.n_layers; ++l) if (P.L[l].n_pe_ks > 0) l_pe_last = l;

  if (warp == WARP_MMA3 || warp == WARP_MMA3 + 1) {
    // ===================== MMA issuers: warp 16 -> tile X, warp 17 -> tile Y (+ weight stream) ============================
    const uint32_t t = (uint32_t)(warp - WARP_MMA3);
    const uint32_t idesc = ptx::idesc_bf16_f32(128, 128);
    const uint64_t desc_base = ptx::smem_desc(0, 2048, 128);
    const uint32_t desc_hi = (uint32_t)(desc_base >> 32);
    const uint32_t desc_lo0 = (uint32_t)(desc_base & 0xFFFFFFFFu);
    auto lo_of = [&](uint32_t saddr) -> uint32_t { return desc_lo0 | ((saddr & 0x3FFFFu) >> 4); };
    uint4* prog = reinterpret_cast<uint4*>(smem + Smem3<VD>::prog) + 48 * t;
    uint2* stab = reinterpret_cast<uint2*>(smem + Smem3<VD>::stab);
    int n_entries = 0;
    if (lane == 0) {
      uint32_t off = 0;
      for (int l = 0; l < P.n_layers; ++l) {
        const int n_pe = P.L[l].n_pe_ks, n_h = P.L[l].n_h_ks, nst = stages_of(n_pe, n_h);
        for (int h = 0; h < P.L[l].n_halves; ++h)
          for (int s2 = 0; s2 < nst; ++s2) {
            const StageInfo si = stage_info(n_pe, n_h, s2);
            uint32_t f = (s2 == 0 ? PF_FIRST : 0u);
            uint32_t a;
            if (si.is_pe) { f |= PF_SS | ((uint32_t)si.nks << 8); a = lo_of(sbase + Smem3<VD>::pe0 + t * PE_TILE_BYTES + (uint32_t)si.k0 * KS_BYTES); }
            else a = tmem + 256u * t + 128u + 8u * (uint32_t)si.k0;
            if (s2 == 0) { f |= PF_WAIT_EP; if (l == 0 && h == 0) f |= PF_WAIT_PE; }
            // the whole layer input is needed from hidden K-step 8 on (or by the half's only stage group when it has none)
            if ((!si.is_pe && si.k0 == KS_PER_STAGE) || (s2 == 0 && n_h <= KS_PER_STAGE)) f |= PF_WAIT_EF;
            if (s2 == nst - 1) f |= PF_COMMIT;
            if (t == 1) stab[n_entries] = make_uint2(off, PLNERF3_DBG(2) ? 16u : (uint32_t)si.nks * KS_BYTES);
            off += (uint32_t)si.nks * KS_BYTES;
            prog[n_entries++] = make_uint4(tmem + 256u * t, a, f, d_full0 + 8u * t);
          }
      }
    }
    n_entries = __shfl_sync(0xffffffffu, n_entries, 0);
    __syncwarp();
    const uint64_t ring_desc = ((uint64_t)desc_hi << 32) | (uint64_t)lo_of(sbase + SL.ring);
    Issue3 st = {0u, 0u, 0u, 0u, 0u, 0u, 0u, (t == 1) ? (uint32_t)(n_pairs * n_entries) : 0u};
    const uint32_t prog_addr = sbase + Smem3<VD>::prog + 16u * 48u * t, stab_addr = sbase + Smem3<VD>::stab;
    if (t == 1 && ptx::elect_one()) {
      // prologue of the weight stream: fill all but REFILL_LAG slots; afterwards every issued stage refills the slot both
      // tiles released REFILL_LAG stages earlier
      constexpr uint32_t REFILL_LAG = 2;
      for (uint32_t s2 = 0; s2 + REFILL_LAG < (uint32_t)n_slots && st.prem > 0; ++s2) {
        const uint2 e = stab[st.pj];
        ptx::mbar_arrive_expect_tx(w_full0 + 8u * st.psl, e.y);
        ptx::bulk_g2s_hint(sbase + SL.ring + st.psl * (uint32_t)STAGE_BYTES, A.w + e.x, e.y, w_full0 + 8u * st.psl, ptx::l2_policy_evict_last());
        if (++st.psl == (uint32_t)n_slots) { st.psl = 0; st.pph ^= 1u; }
        if (++st.pj == (uint32_t)n_entries) st.pj = 0;
        --st.prem;
      }
    }
    __syncwarp();
    for (int pr = 0; pr < n_pairs; ++pr) {
      unsigned long long trp = 0ull;
#if defined(PLNERF_DEBUG) && defined(PLNERF_ENABLE_TRACE)
      // issue-side stamps of block 0's third pair: 3 x 32-bit clocks per program entry behind the epilogue regions
      if (A.trace && blockIdx.x == 0 && pr == 2) trp = (unsigned long long)(A.trace + 4 * 256 * 2 + 128 * t);
#endif
      if (ptx::elect_one())
        issue_tile3(st, prog_addr, (uint32_t)n_entries, ring_desc, idesc, desc_hi, w_full0, w_empty0, ep_early0 + 8u * t, ep_done0 + 8u * t, pe_ready0 + 8u * t,
                    (uint32_t)n_slots, stab_addr, (uint32_t)n_entries, sbase + SL.ring, (unsigned long long)A.w, trp);
      __syncwarp();   // the state lives in the elected lane; elect.sync of a converged warp picks the same lane every time
    }
  } else if (warp >= WARP_HELP3) {
    // ===================== helper warps: encode the next pair's rows, prefetch the pair after =============================
    const int hid = threadIdx.x - 32 * WARP_HELP3;          // 0..63
    const int hw = warp - WARP_HELP3;
    auto encode_pair_tile = [&](int pair, int t) {
      const int64_t g0 = row0 + (2 * (int64_t)pair + t) * TILE_M;
#pragma unroll 1
      for (int rr = hid; rr < TILE_M; rr += 32 * N_HELP_WARPS)
        if (!PLNERF3_DBG(4)) pe_write_row<STASH>(A, smem + Smem3<VD>::pe0 + t * PE_TILE_BYTES, g0 + rr, rr, row_end);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(pe_ready0 + 8u * t);
    };
    auto prefetch_pair = [&](int pair) {
      // depths / ray rows of a later pair -> L2 (a cold DRAM read otherwise): one line per 32 depths, the rows' rays
      if (A.x_emb || pair >= n_pairs) return;
#pragma unroll 1
      for (int rr = 32 * hid; rr < 2 * TILE_M; rr += 32 * 32 * N_HELP_WARPS) {
        const int64_t gt = row0 + 2 * (int64_t)pair * TILE_M + rr, gcn = (gt < row_end) ? gt : (row_end - 1);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.z + gcn));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.rays + (gcn / A.S) * (int64_t)A.stride));
      }
    };
    if (n_pairs > 0) {
      encode_pair_tile(0, 0);
      encode_pair_tile(0, 1);
      prefetch_pair(1);
    }
    uint32_t fph = 0u;                                       // parity of pe_free[0] and pe_free[1] (they advance together)
    int cu = 0;                                              // next ray (local index) to composite; same in both helper warps
    const int n_cta_units = (int)(unit1 - unit0);
    const int64_t rows_cta = row_end - row0;
    const float4* ring = reinterpret_cast<const float4*>(smem + Smem3<VD>::xch0);
    for (int pr = 0; pr < n_pairs; ++pr) {
      if (pr + 1 < n_pairs) {
        // the NEXT pair's tile t, once pair pr's tile t has released its operand (the trunk layers that read it are complete)
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          ptx::mbar_wait(pe_free0 + 8u * t, fph);
          encode_pair_tile(pr + 1, t);
        }
        fph ^= 1u;
        prefetch_pair(pr + 2);
      }
      if (VD && A.fuse_comp) {
        // ---- quadrature of every ray whose last sample arrived with this pair (one warp per ray, alternating warps)
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const uint32_t need = 4u * (uint32_t)(pr + 1);
          while (ctrs[t] < need) __nanosleep(200);      // (a busy spin would steal issue slots from the epilogue warps)
          __threadfence_block();
          int64_t have = (2 * (int64_t)pr + t + 1) * TILE_M;          // local rows present in the ring
          if (have > rows_cta) have = rows_cta;
          while (cu < n_cta_units && (int64_t)(cu + 1) * A.S <= have) {
            if ((cu & 1) == hw) {
              const int64_t r = unit0 + cu;
              const RawRing raw{ring, (int)(((int64_t)cu * A.S) % RING_ROWS), RING_ROWS};
              if (A.comp_mode == PLNERF_MODE_LINEAR) composite_ray<PLNERF_MODE_LINEAR>(A.comp, r, lane, raw, A.z + r * (int64_t)A.S);
              else composite_ray<PLNERF_MODE_CONSTANT>(A.comp, r, lane, raw, A.z + r * (int64_t)A.S);
            }
            ++cu;
          }
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) atomicAdd(const_cast<uint32_t*>(ctrs + 2), 1u);
      }
    }
  } else {
    // ===================== epilogue warps: group t = warp / 8 serves tile t ==============================================
    const int t = warp >> 3;                   // tile of this warp
    const int q = warp & 3, ch = (warp >> 2) & 1;
    const int row = q * 32 + lane;
    const uint32_t tm_row = tmem + (((uint32_t)(q * 32)) << 16) + 256u * (uint32_t)t;
    const uint32_t tm_d = tm_row + 64u * (uint32_t)ch;             // this thread's 64 accumulator columns of a half
    const uint32_t tm_a = tm_row + 128u + 32u * (uint32_t)ch;      // its 32 packed columns inside a half of A_t
    const uint32_t d_full = d_full0 + 8u * t, ep_done = ep_done0 + 8u * t, ep_early = ep_early0 + 8u * t, vb_full = vb_full0 + 8u * t, pe_free = pe_free0 + 8u * t;
    const bool vb_smem = VD && A.viewbias && A.vb_div >= 64;
    float* xch = reinterpret_cast<float*>(smem + Smem3<VD>::xch0) + (VD ? 0 : t * TILE_M * MAX_OUT_CH);   // VD: the output ring
    const float* vbs = reinterpret_cast<const float*>(smem + Smem3<VD>::vb0 + t * (VB_RAYS_SMEM * 512u));
    uint32_t dph = 0u, vph = 0u;
    int vb_idx = 0;
    if (lane == 0) { ptx::mbar_arrive(ep_early); ptx::mbar_arrive(ep_done); }   // the tile starts with a drained accumulator
    int tcnt = 0;

    // Every half-epilogue arrives exactly once on each of the two barriers the tile's issuer waits on: `early` = the
    // accumulator is in registers (and, in a second half, the held first half of the layer input is stored), `full` = the
    // whole layer input is stored.  what: 1 = early, 2 = full, 3 = both.
    auto arrive_now = [&](int what) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (what & 1) ptx::mbar_arrive(ep_early);
        if (what & 2) ptx::mbar_arrive(ep_done);
      }
    };
    auto arrive_ep = [&](int what) { if (!PLNERF3_DBG(8)) arrive_now(what); };
    auto wait_d = [&]() {
      ptx::mbar_wait(d_full, dph); dph ^= 1u;
      ptx::tc_fence_after();
      if (PLNERF3_DBG(8)) arrive_now(3);
    };

    for (int pr = 0; pr < n_pairs; ++pr) {
      const bool trace_on = (pr == 2) && lane == 0 && q == 0 && ch == 0;
      const int64_t g = row0 + (2 * (int64_t)pr + t) * TILE_M + row;
      float alpha_acc = 0.f;
      // training stash: this tile's operand tiles and this row's mask words (layer offsets come from the layer table)
      const int64_t gtile = STASH ? g / TILE_M : 0;
      uint8_t* const st_base = STASH ? A.in_stash + gtile * (int64_t)A.tl.in_tile_bytes : nullptr;
      uint32_t* const m_base = STASH ? A.masks + gtile * (int64_t)A.tl.mask_tile_words + row : nullptr;
      // this thread's 16-byte unit inside a 256-wide operand tile (MN-major 8x8 core matrices, see stash_store8), first of
      // its eight 8-column groups of a half: every store of a trunk layer is then `st_row + layer offset + immediate`
      uint8_t* const st_row = STASH ? st_base + (((row >> 6) * 32 * 8 + ((row & 63) >> 3)) * 128 + (row & 7) * 16 + 8 * ch * 1024) : nullptr;
      for (int l = 0; l < P.n_layers; ++l) {
        const int4 lt = reinterpret_cast<const int4*>(smem + Smem3<VD>::ltab)[l];
        const int epi = lt.x & 255, n_halves = (lt.x >> 8) & 255;
        const bool reads_a = (lt.x >> 24) != 0;
        const float* bias = consts + lt.y + 64 * ch;
        if (epi == EPI_RELU_A || epi == EPI_LINEAR_A) {
          // ---- trunk / feature layer (two 128-neuron halves): bias, (ReLU), bf16 pack -> A_t
          const bool alpha_here = VD && (((lt.x >> 16) & 255) & FLAG_ALPHA);
          const float* aw = consts + P.alpha_w_off + 64 * ch;
          uint32_t held[32];                         // half a (packed), kept until half b's MMAs have read A_t
          // training stash of this layer's output: operand tile of width 256, mask words
          uint8_t* st_t = STASH ? st_row + lt.z : nullptr;
          uint32_t* st_m = (STASH && lt.w >= 0) ? m_base + lt.w + 2 * ch * 128 : nullptr;
          auto stash32 = [&](const uint32_t* pk16, uint32_t m, int h, int c2) {   // 32 columns [128 h + 64 ch + 32 c2, +32)
            if (st_m && !PLNERF3_DBG(64)) st_m[(4 * h + c2) * 128] = m;                    // (64: measurement, no mask stores)
            if (PLNERF3_DBG(32)) return;        // measurement: the stash forward without its activation stores
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(st_t + (16 * h + 4 * c2 + q4) * 1024), "r"(pk16[4 * q4]),
                           "r"(pk16[4 * q4 + 1]), "r"(pk16[4 * q4 + 2]), "r"(pk16[4 * q4 + 3]) : "memory");
          };
          // half a
          PLNERF_TRACE(t * 2 + ch, tcnt, 2000 + l * 10);
          wait_d();
          PLNERF_TRACE(t * 2 + ch, tcnt, 3000 + l * 10);
          if (PLNERF3_DBG(1)) { arrive_now(3); wait_d(); arrive_now(3); goto layer_tail; }
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            uint32_t r[32];
            ptx::tmem_ld32(tm_d + 32u * c2, r);
            ptx::tmem_ld_wait();
            if (c2 == 1) arrive_ep(reads_a ? 3 : 1);   // the accumulator is drained: half b may start (layer 0: `full` follows the store)
            uint32_t m = 0;
            if (epi == EPI_RELU_A) cvt32<true>(r, bias + 32 * c2, held + 16 * c2, (STASH && !PLNERF3_DBG(64)) ? &m : nullptr);
            else cvt32<false>(r, bias + 32 * c2, held + 16 * c2);
            if (STASH) stash32(held + 16 * c2, m, 0, c2);
            if (alpha_here) alpha_acc = dot32_relu(r, aw + 32 * c2, alpha_acc);
          }
          if (!reads_a) {
            // layer 0: nothing reads A_t, store right away
            ptx::tmem_st16(tm_a, reinterpret_cast<uint32_t(&)[16]>(held[0]));
            ptx::tmem_st16(tm_a + 16u, reinterpret_cast<uint32_t(&)[16]>(held[16]));
            ptx::tmem_st_wait();
            arrive_ep(2);
          }
          PLNERF_TRACE(t * 2 + ch, tcnt, 4000 + l * 10);
          // half b
          wait_d();
          PLNERF_TRACE(t * 2 + ch, tcnt, 3000 + l * 10 + 1);
          {
            if (reads_a) {
              ptx::tmem_st16(tm_a, reinterpret_cast<uint32_t(&)[16]>(held[0]));
              ptx::tmem_st16(tm_a + 16u, reinterpret_cast<uint32_t(&)[16]>(held[16]));
            }
            if (STASH) {
              // stash forward (the epilogue is its bound, not the pipe): one 32-column chunk at a time -- with both chunks'
              // 64 accumulator registers in flight next to the stash addresses the compiler spills more, and the reloads
              // come from L2 (the kernel leaves the L1 a few KB); the early signal moves behind the first chunk's work
              // (same-box A/B: -0.9% on the training step; the same order in the INFERENCE kernel costs +3.7%, see below)
              uint32_t pk[16], m = 0;
              {
                uint32_t r0[32];
                ptx::tmem_ld32(tm_d, r0);
                ptx::tmem_ld_wait();
                if (epi == EPI_RELU_A) cvt32<true>(r0, bias + 128, pk, !PLNERF3_DBG(64) ? &m : nullptr); else cvt32<false>(r0, bias + 128, pk);
                if (alpha_here) alpha_acc = dot32_relu(r0, aw + 128, alpha_acc);
              }
              ptx::tmem_st16(tm_a + 64u, pk);
              stash32(pk, m, 1, 0);
              {
                uint32_t r1[32];
                ptx::tmem_ld32(tm_d + 32u, r1);
                ptx::tmem_ld_wait();
                ptx::tmem_st_wait();
                arrive_ep(1);
                PLNERF_TRACE(t * 2 + ch, tcnt, 5000 + l * 10 + 1);
                if (epi == EPI_RELU_A) cvt32<true>(r1, bias + 160, pk, !PLNERF3_DBG(64) ? &m : nullptr); else cvt32<false>(r1, bias + 160, pk);
                if (alpha_here) alpha_acc = dot32_relu(r1, aw + 160, alpha_acc);
              }
              ptx::tmem_st16(tm_a + 80u, pk);
              stash32(pk, m, 1, 1);
            } else {
              // both 32-column chunks in flight (the held half's registers are free again), then the early signal: the next
              // layer's first eight K-steps need only the accumulator drained and the first input half in place
              uint32_t r0[32], r1[32];
              ptx::tmem_ld32(tm_d, r0);
              ptx::tmem_ld32(tm_d + 32u, r1);
              ptx::tmem_ld_wait();
              ptx::tmem_st_wait();
              arrive_ep(1);
              PLNERF_TRACE(t * 2 + ch, tcnt, 5000 + l * 10 + 1);
              uint32_t pk[16];
              if (epi == EPI_RELU_A) cvt32<true>(r0, bias + 128, pk, nullptr); else cvt32<false>(r0, bias + 128, pk);
              ptx::tmem_st16(tm_a + 64u, pk);
              if (epi == EPI_RELU_A) cvt32<true>(r1, bias + 160, pk, nullptr); else cvt32<false>(r1, bias + 160, pk);
              ptx::tmem_st16(tm_a + 80u, pk);
              if (alpha_here) { alpha_acc = dot32_relu(r0, aw + 128, alpha_acc); alpha_acc = dot32_relu(r1, aw + 160, alpha_acc); }
            }
          }
          ptx::tmem_st_wait();
          arrive_ep(2);
          PLNERF_TRACE(t * 2 + ch, tcnt, 4000 + l * 10 + 1);
        } else if (VD) {
          // ---- views layer (one half): per-ray bias (view-direction columns), ReLU, rgb head; then the row's 4 outputs
          wait_d();
          const float* vbr;
          if (vb_smem) {
            ptx::mbar_wait(vb_full, vph); vph ^= 1u;
            vbr = vbs + vb_idx * 128 + 64 * ch;
          } else {
            const int64_t gc = (g < row_end) ? g : (row_end - 1);
            vbr = A.viewbias + (gc / A.vb_div) * 128 + 64 * ch;
          }
          const float* rw = consts + P.rgb_w_off + 64 * ch;
          float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            uint32_t r[32];
            ptx::tmem_ld32(tm_d + 32u * c2, r);
            ptx::tmem_ld_wait();
            if (c2 == 1) arrive_ep(3);                 // the accumulator is in registers: hand it back before the head math
            const float* val = reinterpret_cast<const float*>(r);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(vbr + 32 * c2 + i);
              const float4 w0 = *reinterpret_cast<const float4*>(rw + 32 * c2 + i);
              const float4 w1 = *reinterpret_cast<const float4*>(rw + 128 + 32 * c2 + i);
              const float4 w2 = *reinterpret_cast<const float4*>(rw + 256 + 32 * c2 + i);
              const float v0 = fmaxf(val[i] + b4.x, 0.f), v1 = fmaxf(val[i + 1] + b4.y, 0.f);
              const float v2 = fmaxf(val[i + 2] + b4.z, 0.f), v3 = fmaxf(val[i + 3] + b4.w, 0.f);
              if (STASH) { r[i] = __float_as_uint(v0); r[i + 1] = __float_as_uint(v1); r[i + 2] = __float_as_uint(v2); r[i + 3] = __float_as_uint(v3); }
              h0 = fmaf(v0, w0.x, h0); h0 = fmaf(v1, w0.y, h0); h0 = fmaf(v2, w0.z, h0); h0 = fmaf(v3, w0.w, h0);
              h1 = fmaf(v0, w1.x, h1); h1 = fmaf(v1, w1.y, h1); h1 = fmaf(v2, w1.z, h1); h1 = fmaf(v3, w1.w, h1);
              h2 = fmaf(v0, w2.x, h2); h2 = fmaf(v1, w2.y, h2); h2 = fmaf(v2, w2.z, h2); h2 = fmaf(v3, w2.w, h2);
            }
            if (STASH) {
              // the views layer's output (128 wide) and its ReLU mask
              uint8_t* st_t = st_base + lt.z;
              uint32_t pk[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = ptx::pack_bf16(val[2 * i], val[2 * i + 1]);
              m_base[lt.w + (2 * ch + c2) * 128] = relu_mask_from_packed(pk);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                stash_store8(st_t, 128, row, 8 * ch + 4 * c2 + q4, make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]));
            }
          }
          // combine the two column halves (fixed order: bitwise reproducible); the row's four outputs go to the output
          // ring (slot of pair pr % RING_PAIRS; read by the compositor warps when A.fuse_comp) and, if asked for, to HBM
          float4* slot = reinterpret_cast<float4*>(xch) + ((2 * pr + t) % (2 * RING_PAIRS)) * TILE_M + row;
          if (A.fuse_comp && pr >= RING_PAIRS) {
            // the slot's previous rows (pair pr - 3) are released once every ray ending in pair pr - 2 is composited
            const uint32_t need = 2u * (uint32_t)(pr - 1);
            while (ctrs[2] < need) __nanosleep(100);
            __threadfence_block();
          }
          if (ch == 1) *slot = make_float4(h0, h1, h2, alpha_acc);
          asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "n"(32 * EPI_WARPS_PER_TILE) : "memory");
          if (ch == 0) {
            const float4 x = *slot;
            float4 acc;
            acc.x = h0 + x.x + consts[P.rgb_b_off + 0]; acc.y = h1 + x.y + consts[P.rgb_b_off + 1];
            acc.z = h2 + x.z + consts[P.rgb_b_off + 2]; acc.w = alpha_acc + x.w + consts[P.alpha_b_off];
            *slot = acc;
            if (!(A.fuse_comp && A.skip_out) && g < row_end) {
              float* o = A.out + g * (int64_t)A.out_stride;
              if (A.out_stride == 4) *reinterpret_cast<float4*>(o) = acc;
              else { o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w; }
            }
            if (A.fuse_comp) {
              __threadfence_block();
              __syncwarp();
              if (lane == 0) atomicAdd(const_cast<uint32_t*>(ctrs + t), 1u);
            }
          }
        } else {
          // ---- last trunk layer of a network without view directions: ReLU, output_linear on the fp32 values
          const int nch = P.out_ch;
          float hp[MAX_OUT_CH];
#pragma unroll
          for (int c = 0; c < MAX_OUT_CH; ++c) hp[c] = 0.f;
          for (int h = 0; h < n_halves; ++h) {
            wait_d();
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
              uint32_t r[32];
              ptx::tmem_ld32(tm_d + 32u * c2, r);
              ptx::tmem_ld_wait();
              if (c2 == 1) arrive_ep(3);
              float* val = reinterpret_cast<float*>(r);
              const float* b = bias + 128 * h + 32 * c2;
#pragma unroll
              for (int i = 0; i < 32; ++i) val[i] += b[i];
#pragma unroll
              for (int c = 0; c < MAX_OUT_CH; ++c)
                if (c < nch) hp[c] = dot32_relu(r, consts + P.out_w_off + c * 256 + 128 * h + 64 * ch + 32 * c2, hp[c]);
            }
          }
          if (ch == 1) {
#pragma unroll
            for (int c = 0; c < MAX_OUT_CH; ++c) if (c < nch) xch[row * MAX_OUT_CH + c] = hp[c];
          }
          asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "n"(32 * EPI_WARPS_PER_TILE) : "memory");
          if (ch == 0 && g < row_end) {
            float* o = A.out + g * (int64_t)A.out_stride;
#pragma unroll
            for (int c = 0; c < MAX_OUT_CH; ++c) if (c < nch) o[c] = hp[c] + xch[row * MAX_OUT_CH + c] + consts[P.out_b_off + c];
          }
        }
      layer_tail:
        // ---- bookkeeping hidden behind the MMAs in flight
        if (l == l_pe_last) {
          // every MMA that reads this tile's encoding operand has completed: the helper warps may overwrite it
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(pe_free);
        }
        if (l == 1 && vb_smem) {
          // this tile's per-ray view-bias rows -> shared memory, one bulk copy.  All 8 warps of the group are past the
          // previous pair's views epilogue here (this layer's MMAs needed their arrivals of layer 0's epilogue).
          const int64_t trow0 = row0 + (2 * (int64_t)pr + t) * TILE_M;
          const int64_t gc = (g < row_end) ? g : (row_end - 1);
          const int64_t ray0 = trow0 / A.vb_div;
          const int d = (int)(gc / A.vb_div - ray0);      // (the 64-bit divisions stay out of the views epilogue)
          vb_idx = d < 0 ? 0 : (d >= (int)VB_RAYS_SMEM ? (int)VB_RAYS_SMEM - 1 : d);
          if ((warp & 7) == 0 && lane == 0) {
            const int64_t n_rays = (A.M + A.vb_div - 1) / A.vb_div;
            if (ray0 < n_rays) {
              const int64_t nr = (n_rays - ray0 < (int64_t)VB_RAYS_SMEM) ? n_rays - ray0 : (int64_t)VB_RAYS_SMEM;
              ptx::mbar_arrive_expect_tx(vb_full, (uint32_t)nr * 512u);
              ptx::bulk_g2s(sbase + Smem3<VD>::vb0 + t * (VB_RAYS_SMEM * 512u), A.viewbias + ray0 * 128, (uint32_t)nr * 512u, vb_full);
            } else {
              ptx::mbar_arrive(vb_full);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA3) ptx::tmem_dealloc(tmem, 512);
}

}  // namespace v3
