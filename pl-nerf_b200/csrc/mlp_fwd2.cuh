// k_mlp2: second-generation fused PE + MLP forward (bf16 operands, fp32 accumulate), included by
// mlp_fwd.cu inside its anonymous namespace (shares NetPlan / MlpArgs / the packed weight stream).
//
// Why a second kernel: k_mlp_fwd keeps ONE 128-row tile in flight per SM, so every layer pays the
// full MMA -> commit -> epilogue -> arrive -> issue round trip (~1300 cycles, timeline traces in
// profiles/): the tensor pipe idles ~54% of the time even with an empty epilogue.  k_mlp2 keeps TWO
// independent tiles ("slots") in flight per CTA and strictly alternates them: while the tensor pipe
// runs layer l of slot 0, the 16 epilogue warps drain layer l of slot 1, and vice versa.
//
//   * tensor memory holds only accumulators: D_0 = columns [0,256), D_1 = [256,512) (128x256 fp32 each);
//   * activations live in SHARED memory as the SS-form A operand, 80 KB per slot: 32 hidden panels
//     + 8 positional-encoding panels (panel = 8 K-columns x 128 rows x 16 B, K-major core matrices,
//     no swizzle).  A layer's epilogue overwrites the hidden panels IN PLACE: it only starts once
//     every MMA of that layer (the only readers) has completed (tcgen05.commit -> d_full);
//   * every layer is issued as N=256 tcgen05.mma (16-20 K-steps);
//   * kCta == 2: the two CTAs of a cluster form a pair (tcgen05 cta_group::2, M=256): each CTA owns
//     128 rows of both slots and streams only ITS 128-neuron half of every weight K-step (halves the
//     L2->SM weight traffic and the shared-memory operand reads per CTA); the leader CTA issues, commits
//     are multicast to both CTAs, the peer's epilogue warps arrive remotely on the leader's barriers;
//   * biases are added INSIDE the tensor pipe: one extra K-step per layer multiplies a constant ones operand
//     with a packed [bias_hi, bias_mid, bias_lo] block (three bf16 terms = the fp32 bias exactly), so the hot epilogue is only
//     tcgen05.ld -> cvt.rn.relu.bf16x2 -> st.shared;
//   * weights: the same packed stream as k_mlp_fwd ([layer][half][K-step][panel][128 rows][16 B]),
//     pulled through a ring of 2-K-step stages by 1-D bulk TMA.
//
// Replaces run_network / NeRF.forward (reference run_plnerf.py:78-92, run_nerf_helpers.py:105-128)
// for PLNERF_PREC_BF16 inference; bf16x3 and the training modes stay on k_mlp_fwd.

namespace v2 {

#ifdef PLNERF_ENABLE_TRACE
#define V2_TRACE(region, cnt, code)                                                                  \
  do {                                                                                                \
    if (A.trace && blockIdx.x == 0 && r == 2 && l >= 1 && l <= 3 && (cnt) < 256) {                    \
      A.trace[((region) * 256 + (cnt)) * 2] = clock64();                                              \
      A.trace[((region) * 256 + (cnt)) * 2 + 1] = (code);                                             \
      ++(cnt);                                                                                        \
    }                                                                                                 \
  } while (0)
#else
#define V2_TRACE(region, cnt, code) do { (void)(cnt); } while (0)
#endif

constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 64 + EPI_THREADS;          // warps 0-15 epilogue, warp 16 = TMA producer (+TMEM alloc), warp 17 = MMA issuer / peer relay
// (the SM's warp arbiter prefers the HIGHEST warp id: the latency-critical single-thread roles sit above the epilogue warps)
constexpr int WARP_TMA = EPI_WARPS, WARP_MMA = EPI_WARPS + 1;
// A slot = [PE panels 16 KB (right-aligned: they end where the hidden panels start) | hidden panels 64 KB | ones panels 4 KB]: the K-steps of a layer ([PE,] hidden, bias)
// are CONTIGUOUS in descriptor space, so a layer is one run of K-steps for the issuer
constexpr int A_PE_OFF = 0, A_HID_OFF = PE_TILE_BYTES, A_HID_BYTES = 32 * 2048, A_ONES_OFF = A_HID_OFF + A_HID_BYTES;
constexpr int A_SLOT_BYTES = A_ONES_OFF + KS_BYTES;   // 84 KB
// K-steps per ring stage: 2 (16 KB with both halves for a single CTA, 8 KB = this CTA's half for a CTA pair)
__host__ __device__ constexpr int ks_per_stage(int kcta) { return 2; }
constexpr int MAX_ST2 = 8;

struct Smem2 { uint32_t a[2], ring, stage_bytes, n_stages, consts, lay_issue, bars, total; };
// first float of the const block the epilogue still needs (head weights / head biases); layer biases live in the MMA
__host__ __device__ inline int head_const_off(const NetPlan& P) { return P.use_viewdirs ? P.alpha_w_off : P.out_w_off; }
__host__ __device__ inline Smem2 smem2_layout(int kcta, int const_floats, int max_smem) {
  Smem2 s;
  s.a[0] = 0; s.a[1] = A_SLOT_BYTES;
  s.ring = 2 * A_SLOT_BYTES;
  s.stage_bytes = (uint32_t)(ks_per_stage(kcta) * KS_BYTES * (kcta == 1 ? 2 : 1));
  const uint32_t cbytes = (uint32_t)((const_floats * 4 + 127) & ~127);
  const uint32_t fixed = s.ring + cbytes + MAX_LAYERS * 32 + 512;
  int n = ((uint32_t)max_smem > fixed) ? (int)(((uint32_t)max_smem - fixed) / s.stage_bytes) : 0;
  if (n > MAX_ST2) n = MAX_ST2;
  n &= ~1;                                   // stages are consumed in adjacent pairs ("super-stages")
  s.n_stages = (uint32_t)n;
  s.consts = s.ring + s.n_stages * s.stage_bytes;
  s.lay_issue = s.consts + cbytes;
  s.bars = s.lay_issue + MAX_LAYERS * 32;
  s.total = s.bars + 512;
  return s;
}

// A layer's K-steps as <= 2 runs that are contiguous in the slot's A panels: run = (byte offset of its first A K-step
// inside the slot, number of K-steps).  Layer input = PE only (first layer): [PE] and [ones]; otherwise one run
// [PE,] hidden [, ones].  The B stream of the layer has the same order: PE K-steps, hidden K-steps, bias block.
struct Runs { int n; int a_off[2]; int nks[2]; };
__device__ __forceinline__ Runs layer_runs(const NetPlan& P, int l) {
  const int n_pe = P.L[l].n_pe_ks, n_h = P.L[l].n_h_ks, hb = (P.bias_block_idx[l] >= 0) ? 1 : 0;
  Runs R;
  if (n_h == 0) { R.n = 1 + hb; R.a_off[0] = A_HID_OFF - n_pe * KS_BYTES; R.nks[0] = n_pe; R.a_off[1] = A_ONES_OFF; R.nks[1] = hb; }
  else { R.n = 1; R.a_off[0] = (n_pe > 0) ? A_HID_OFF - n_pe * KS_BYTES : A_HID_OFF; R.nks[0] = n_pe + n_h + hb; R.a_off[1] = 0; R.nks[1] = 0; }
  return R;
}
// every run is cut into super-stages of 4 K-steps = two adjacent ring stages of 2 K-steps (the second may be empty)

__host__ __device__ inline int stages2_of(int KS2, int n_pe, int n_h) { return (n_pe + KS2 - 1) / KS2 + (n_h + KS2 - 1) / KS2; }
__host__ __device__ inline StageInfo stage2_info(int KS2, int n_pe, int n_h, int i) {
  const int npe_st = (n_pe + KS2 - 1) / KS2;
  StageInfo si;
  if (i < npe_st) { si.is_pe = 1; si.k0 = i * KS2; si.nks = min(KS2, n_pe - si.k0); }
  else { si.is_pe = 0; si.k0 = (i - npe_st) * KS2; si.nks = min(KS2, n_h - si.k0); }
  return si;
}

// ---- cluster / cta_group::2 primitives -----------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier given by its shared::cluster address (own CTA or the peer).  Default (.release.cta) semantics
// like cutlass::arch::ClusterBarrier::arrive(cta_id): an explicit .release.cluster costs >1000 cycles per arrive here;
// the data handed over is shared memory made visible to the async proxy by fence.proxy.async before the arrive.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
template <int kCta>
__device__ __forceinline__ void tmem_alloc_g(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (kCta == 1) {
    ptx::tmem_alloc(dst_smem, ncols);
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int kCta>
__device__ __forceinline__ void tmem_dealloc_g(uint32_t taddr, uint32_t ncols) {
  if constexpr (kCta == 1) ptx::tmem_dealloc(taddr, ncols);
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int kCta>
__device__ __forceinline__ void mma_ss_g(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kCta == 1) {
    ptx::mma_ss(d_tmem, a_desc, b_desc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// commit: the barrier (same shared offset in both CTAs of the pair when kCta == 2) gets one arrival once all
// previously issued MMAs have completed
template <int kCta>
__device__ __forceinline__ void mma_commit_g(uint32_t bar) {
  if constexpr (kCta == 1) {
    ptx::mma_commit(bar);
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
  }
}


// ---- tight SS-form MMA issue: NK consecutive K=16 steps from ONE asm block ------------------------
// A descriptor advances by one K-step (two 2 KB panels = 256 descriptor units), B by `bstep`.
#define V2_STEP(CG, PRED) \
  "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], ad, bd, %3, " PRED ";\n\t" \
  "add.u64 ad, ad, 256;\n\tadd.u64 bd, bd, %5;\n\t"
#define V2_PROLOG \
  "{\n\t.reg .pred p, pt;\n\t.reg .b64 ad, bd;\n\t" \
  "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 pt, %4, %4;\n\tmov.b64 ad, %1;\n\tmov.b64 bd, %2;\n\t"
#define V2_ARGS ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "l"(bstep) : "memory"
template <int kCta, int NK>
__device__ __forceinline__ void issue_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint64_t bstep) {
  if constexpr (kCta == 1) {
    if constexpr (NK == 4) asm volatile(V2_PROLOG V2_STEP("1", "p") V2_STEP("1", "pt") V2_STEP("1", "pt") V2_STEP("1", "pt") "}" V2_ARGS);
    else if constexpr (NK == 2) asm volatile(V2_PROLOG V2_STEP("1", "p") V2_STEP("1", "pt") "}" V2_ARGS);
    else asm volatile(V2_PROLOG V2_STEP("1", "p") "}" V2_ARGS);
  } else {
    if constexpr (NK == 4) asm volatile(V2_PROLOG V2_STEP("2", "p") V2_STEP("2", "pt") V2_STEP("2", "pt") V2_STEP("2", "pt") "}" V2_ARGS);
    else if constexpr (NK == 2) asm volatile(V2_PROLOG V2_STEP("2", "p") V2_STEP("2", "pt") "}" V2_ARGS);
    else asm volatile(V2_PROLOG V2_STEP("2", "p") "}" V2_ARGS);
  }
}

// One full super-stage from ONE asm block: K-steps 0,1 read ring stage s (B at %2), K-steps 2,3 ring stage s+1 (B at %6);
// each ring stage is released (tcgen05.commit -> its empty barrier) right after its two MMAs.
#define V2_SUPER(CG, COMMIT) \
  "{\n\t.reg .pred p, pt;\n\t.reg .b64 ad, bd;\n\t" \
  "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 pt, %4, %4;\n\tmov.b64 ad, %1;\n\tmov.b64 bd, %2;\n\t" \
  "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], ad, bd, %3, p;\n\t" \
  "add.u64 ad, ad, 256;\n\tadd.u64 bd, bd, %5;\n\t" \
  "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], ad, bd, %3, pt;\n\t" \
  COMMIT(CG, "%7") \
  "add.u64 ad, ad, 256;\n\tmov.b64 bd, %6;\n\t" \
  "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], ad, bd, %3, pt;\n\t" \
  "add.u64 ad, ad, 256;\n\tadd.u64 bd, bd, %5;\n\t" \
  "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], ad, bd, %3, pt;\n\t" \
  COMMIT(CG, "%8") "}"
#define V2_COMMIT1(CG, BAR) "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [" BAR "];\n\t"
#define V2_COMMIT2(CG, BAR) "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [" BAR "], %9;\n\t"
#define V2_NOCOMMIT(CG, BAR) ""
template <int kCta, bool kCommit>
__device__ __forceinline__ void issue_super(uint32_t d, uint64_t a, uint64_t b0, uint64_t b1, uint32_t idesc, uint32_t acc,
                                            uint64_t bstep, uint32_t empty0, uint32_t empty1) {
  const uint16_t mask = 3;
  if constexpr (kCta == 1) {
    if constexpr (kCommit) asm volatile(V2_SUPER("1", V2_COMMIT1) ::"r"(d), "l"(a), "l"(b0), "r"(idesc), "r"(acc), "l"(bstep), "l"(b1), "r"(empty0), "r"(empty1), "h"(mask) : "memory");
    else asm volatile(V2_SUPER("1", V2_NOCOMMIT) ::"r"(d), "l"(a), "l"(b0), "r"(idesc), "r"(acc), "l"(bstep), "l"(b1), "r"(empty0), "r"(empty1), "h"(mask) : "memory");
  } else {
    if constexpr (kCommit) asm volatile(V2_SUPER("2", V2_COMMIT2) ::"r"(d), "l"(a), "l"(b0), "r"(idesc), "r"(acc), "l"(bstep), "l"(b1), "r"(empty0), "r"(empty1), "h"(mask) : "memory");
    else asm volatile(V2_SUPER("2", V2_NOCOMMIT) ::"r"(d), "l"(a), "l"(b0), "r"(idesc), "r"(acc), "l"(bstep), "l"(b1), "r"(empty0), "r"(empty1), "h"(mask) : "memory");
  }
}

// ---- a whole run of K-steps from ONE asm block --------------------------------------------------------------------
// n_super full super-stages (4 K-steps over ring stages slot, slot+1) followed by an optional tail of 1-3 K-steps
// (also two ring stages, the second possibly empty) and an optional accumulator-full commit.  Everything between the
// MMAs -- ring-slot barrier addresses, waits, B descriptors, stage releases, slot/phase advance -- lives in registers
// declared inside the block, which ptxas keeps on the uniform datapath (2-4 cycles per op instead of an R2UR round
// trip per value): measured ~230 cycles of fixed cost per asm call, nothing noticeable per iteration.
// Barrier layout (8 bytes each, relative to w_full(0)): w_empty(s) = +8*MAX_ST2, w_fullp(s) = +16*MAX_ST2.
#define V2_RUN_WAIT(L, OFF) \
  L ":\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [fb+" OFF "], ph;\n\t@!p bra " L ";\n\t"
#define V2_RUN_WAITS_1 V2_RUN_WAIT("W0", "0") V2_RUN_WAIT("W1", "8")
#define V2_RUN_WAITS_2 V2_RUN_WAIT("W0", "0") V2_RUN_WAIT("W1", "8") V2_RUN_WAIT("W2", "128") V2_RUN_WAIT("W3", "136")
#define V2_TAIL_WAITS_1 V2_RUN_WAIT("X0", "0") V2_RUN_WAIT("X1", "8")
#define V2_TAIL_WAITS_2 V2_RUN_WAIT("X0", "0") V2_RUN_WAIT("X1", "8") V2_RUN_WAIT("X2", "128") V2_RUN_WAIT("X3", "136")
#define V2_RUN_MMA(CG, PRED) "tcgen05.mma.cta_group::" CG ".kind::f16 [%2], ad, bd, %3, " PRED ";\n\t"
#define V2_RUN_COMMIT_1(ADDR) "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [" ADDR "];\n\t"
#define V2_RUN_COMMIT_2(ADDR) "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [" ADDR "], cm;\n\t"
#define V2_RUN_BODY(CG, WAITS, TWAITS, COMMIT) \
  "{\n\t.reg .pred p, pacc, pt;\n\t.reg .b32 sl, ph, n, fb, eb;\n\t.reg .b64 ad, bd, so;\n\t.reg .b16 cm;\n\t" \
  "mov.b16 cm, 3;\n\tmov.b32 sl, %0;\n\tmov.b32 ph, %1;\n\tmov.b32 n, %9;\n\tmov.b64 ad, %4;\n\t" \
  "setp.ne.b32 pacc, %8, 0;\n\tsetp.eq.b32 pt, sl, sl;\n\t" \
  "setp.eq.b32 p, n, 0;\n\t@p bra TAIL;\n\t" \
  "LOOP:\n\t" \
  "shl.b32 fb, sl, 3;\n\tadd.u32 fb, fb, %11;\n\tadd.u32 eb, fb, 64;\n\t" \
  WAITS \
  "mul.wide.u32 so, sl, %7;\n\tadd.u64 bd, so, %5;\n\t" \
  V2_RUN_MMA(CG, "pacc") "add.u64 ad, ad, 256;\n\tadd.u64 bd, bd, %6;\n\t" \
  V2_RUN_MMA(CG, "pt") COMMIT("eb") \
  "add.u64 ad, ad, 256;\n\tadd.u64 so, so, %13;\n\tadd.u64 bd, so, %5;\n\t" \
  V2_RUN_MMA(CG, "pt") "add.u64 ad, ad, 256;\n\tadd.u64 bd, bd, %6;\n\t" \
  V2_RUN_MMA(CG, "pt") COMMIT("eb+8") \
  "add.u64 ad, ad, 256;\n\tsetp.eq.b32 pacc, sl, sl;\n\t" \
  "add.u32 sl, sl, 2;\n\tsetp.eq.u32 p, sl, %12;\n\t@p mov.b32 sl, 0;\n\t@p xor.b32 ph, ph, 1;\n\t" \
  "sub.u32 n, n, 1;\n\tsetp.ne.b32 p, n, 0;\n\t@p bra LOOP;\n\t" \
  "TAIL:\n\t" \
  "setp.eq.b32 p, %10, 0;\n\t@p bra DONE;\n\t" \
  "shl.b32 fb, sl, 3;\n\tadd.u32 fb, fb, %11;\n\tadd.u32 eb, fb, 64;\n\t" \
  TWAITS \
  "mul.wide.u32 so, sl, %7;\n\tadd.u64 bd, so, %5;\n\t" \
  V2_RUN_MMA(CG, "pacc") \
  "setp.lt.u32 p, %10, 2;\n\t@p bra T1;\n\t" \
  "add.u64 ad, ad, 256;\n\tadd.u64 bd, bd, %6;\n\t" V2_RUN_MMA(CG, "pt") \
  "T1:\n\t" COMMIT("eb") \
  "setp.lt.u32 p, %10, 3;\n\t@p bra T2;\n\t" \
  "add.u64 ad, ad, 256;\n\tadd.u64 so, so, %13;\n\tadd.u64 bd, so, %5;\n\t" V2_RUN_MMA(CG, "pt") \
  "T2:\n\t" COMMIT("eb+8") \
  "add.u32 sl, sl, 2;\n\tsetp.eq.u32 p, sl, %12;\n\t@p mov.b32 sl, 0;\n\t@p xor.b32 ph, ph, 1;\n\t" \
  "DONE:\n\t" \
  "setp.eq.b32 p, %14, 0;\n\t@p bra FIN;\n\t" \
  COMMIT("%14") \
  "FIN:\n\t" \
  "mov.b32 %0, sl;\n\tmov.b32 %1, ph;\n\t}"
static_assert(MAX_ST2 == 8, "V2_RUN_BODY hard-codes the barrier offsets 64 (w_empty) and 128 (w_fullp)");
template <int kCta>
__device__ __forceinline__ void issue_run(uint32_t& slot, uint32_t& phase, uint32_t d, uint32_t idesc, uint64_t a, uint64_t bbase,
                                          uint64_t bstep, uint32_t stage16, uint32_t acc, uint32_t n_super, uint32_t tail,
                                          uint32_t full0, uint32_t nst, uint32_t dfull_bar) {
  const uint64_t stage16_64 = stage16;
  if constexpr (kCta == 1) {
    asm volatile(V2_RUN_BODY("1", V2_RUN_WAITS_1, V2_TAIL_WAITS_1, V2_RUN_COMMIT_1)
                 : "+r"(slot), "+r"(phase)
                 : "r"(d), "r"(idesc), "l"(a), "l"(bbase), "l"(bstep), "r"(stage16), "r"(acc), "r"(n_super), "r"(tail), "r"(full0),
                   "r"(nst), "l"(stage16_64), "r"(dfull_bar)
                 : "memory");
  } else {
    asm volatile(V2_RUN_BODY("2", V2_RUN_WAITS_2, V2_TAIL_WAITS_2, V2_RUN_COMMIT_2)
                 : "+r"(slot), "+r"(phase)
                 : "r"(d), "r"(idesc), "l"(a), "l"(bbase), "l"(bstep), "r"(stage16), "r"(acc), "r"(n_super), "r"(tail), "r"(full0),
                   "r"(nst), "l"(stage16_64), "r"(dfull_bar)
                 : "memory");
  }
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) { ptx::tmem_ld32(taddr, r); }

// Positional encoding of one row -> this thread's two 8-column panels (c = column group 0..3) of the slot's PE panels.
__device__ __forceinline__ void pe_rows(const MlpArgs& A, uint8_t* pe_panels, int64_t g_row, int row, int c) {
  const NetPlan& P = A.plan;
  const int64_t gc = (g_row < A.M) ? g_row : (A.M - 1);
  const int n_panels = 2 * P.pe_ks;
  const int p_lo = c * n_panels / 4, p_hi = (c + 1) * n_panels / 4;
  float p[3] = {0.f, 0.f, 0.f};
  const float* xr = nullptr;
  if (A.x_emb) {
    xr = A.x_emb + gc * (int64_t)A.x_ld;
  } else {
    const int64_t ray = gc / A.S;
    const float* rp = A.rays + ray * (int64_t)A.stride;
    const float zz = A.z[gc];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = __fadd_rn(rp[k], __fmul_rn(rp[3 + k], zz));  // o + d*z, two roundings (run_plnerf.py:707)
  }
  const uint32_t turns[3] = {pe_turns(p[0]), pe_turns(p[1]), pe_turns(p[2])};
  auto elem = [&](int idx) -> float {
    if (idx >= P.input_ch) return 0.f;
    if (xr) return xr[idx];
    if (idx < 3) return p[idx];
    const int t = idx - 3, k = t / 6, r = t - 6 * k, cc = (r >= 3) ? r - 3 : r;
    return (r >= 3) ? pe_cos(turns[cc], k) : pe_sin(turns[cc], k);   // run_nerf_helpers.py:45-48; common.cuh pe_turns
  };
  if (xr == nullptr && n_panels == 8) {
    // 63-wide encoding: one fully unrolled 16-element group per thread, values stay in registers (common.cuh)
    float v16[16];
    pe_group16<false>(c, p, turns, P.input_ch, v16);
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      uint4 q;
      q.x = ptx::pack_bf16(v16[8 * h2 + 0], v16[8 * h2 + 1]); q.y = ptx::pack_bf16(v16[8 * h2 + 2], v16[8 * h2 + 3]);
      q.z = ptx::pack_bf16(v16[8 * h2 + 4], v16[8 * h2 + 5]); q.w = ptx::pack_bf16(v16[8 * h2 + 6], v16[8 * h2 + 7]);
      *reinterpret_cast<uint4*>(pe_panels + (p_lo + h2) * 2048 + row * 16) = q;
    }
    return;
  }
  for (int pnl = p_lo; pnl < p_hi; ++pnl) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = elem(8 * pnl + e);
    uint4 q;
    q.x = ptx::pack_bf16(v[0], v[1]); q.y = ptx::pack_bf16(v[2], v[3]);
    q.z = ptx::pack_bf16(v[4], v[5]); q.w = ptx::pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(pe_panels + pnl * 2048 + row * 16) = q;
  }
}

template <int kCta>
__global__ void __launch_bounds__(THREADS, 1) k_mlp2(const __grid_constant__ MlpArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const NetPlan& P = A.plan;
  const int head_off = head_const_off(P);
  const Smem2 SL = smem2_layout(kCta, P.const_floats - head_off, A.n_stages /* carries max smem */);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (kCta == 2) ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0);
  const int NST = (int)SL.n_stages;
  constexpr int KS2 = ks_per_stage(kCta);

  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t s_a[2] = {sbase + SL.a[0], sbase + SL.a[1]};
  const uint32_t s_ring = sbase + SL.ring;
  float* consts = reinterpret_cast<float*>(smem + SL.consts) - head_off;   // indexed with the plan's absolute float offsets
  const uint32_t s_bars = sbase + SL.bars;
  auto w_full = [&](int s) { return s_bars + 8u * s; };
  auto w_empty = [&](int s) { return s_bars + 8u * (MAX_ST2 + s); };
  auto w_fullp = [&](int s) { return s_bars + 8u * (2 * MAX_ST2 + s); };       // leader only: peer's stage landed
  auto d_full = [&](int t) { return s_bars + 8u * (3 * MAX_ST2 + t); };
  auto a_ready = [&](int t) { return s_bars + 8u * (3 * MAX_ST2 + 2 + t); };   // leader's copy is the one waited on
  auto pe_ready = [&](int t) { return s_bars + 8u * (3 * MAX_ST2 + 4 + t); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SL.bars + 8u * (3 * MAX_ST2 + 6));
  uint32_t* lay_off = reinterpret_cast<uint32_t*>(smem + SL.bars + 8u * (3 * MAX_ST2 + 8));   // [MAX_LAYERS] byte offset of each layer in the stream
  // per-layer issue parameters, read by the MMA issuer one layer ahead: {idesc, B K-step stride>>4, B LBO bits (lo word),
  // runs, run0 A offset>>4, run0 K-steps, run1 A offset>>4, run1 K-steps}
  uint4* lay_issue = reinterpret_cast<uint4*>(smem + SL.lay_issue);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { ptx::mbar_init(w_full(s), 1); ptx::mbar_init(w_empty(s), 1); ptx::mbar_init(w_fullp(s), 1); }
    for (int t = 0; t < 2; ++t) {
      ptx::mbar_init(d_full(t), 1);
      ptx::mbar_init(a_ready(t), EPI_WARPS * kCta);
      ptx::mbar_init(pe_ready(t), EPI_WARPS * kCta);
    }
    uint32_t off = 0;
    for (int l = 0; l < P.n_layers; ++l) {
      lay_off[l] = off; off += (uint32_t)(P.L[l].n_halves * (P.L[l].n_pe_ks + P.L[l].n_h_ks) * KS_BYTES);
      const int nh = P.L[l].n_halves;
      const uint32_t b_ks = (kCta == 1) ? (nh == 2 ? 8192u : 4096u) : (nh == 2 ? 4096u : 2048u);   // B bytes per K-step in a stage
      const Runs R = layer_runs(P, l);
      lay_issue[2 * l] = make_uint4(ptx::idesc_bf16_f32(128 * kCta, nh == 2 ? 256 : 128), b_ks >> 4, (((b_ks / 2) >> 4) & 0x3FFFu) << 16, (uint32_t)R.n);
      lay_issue[2 * l + 1] = make_uint4((uint32_t)R.a_off[0] >> 4, (uint32_t)R.nks[0], (uint32_t)R.a_off[1] >> 4, (uint32_t)R.nks[1]);
    }
    ptx::fence_mbar_init();
  }
  if (warp == WARP_TMA) tmem_alloc_g<kCta>(ptx::smem_u32(tmem_slot), 512);
  for (int i = head_off + threadIdx.x; i < P.const_floats; i += THREADS) consts[i] = A.tail[i];
  // ones operand: panel 0 = [1, 1, 1, 0, 0, 0, 0, 0] in every row (bf16 1.0 = 0x3F80), panel 1 = 0
  for (int i = threadIdx.x; i < 2 * (KS_BYTES / 16); i += THREADS) {
    const int t = i / (KS_BYTES / 16), j = i % (KS_BYTES / 16);
    reinterpret_cast<uint4*>(smem + SL.a[t] + A_ONES_OFF)[j] = (j < 128) ? make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kCta == 2) cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
#if 1   // whole-kernel SM-cycle / wall-time pair (block 0 only, when a trace buffer is attached): clock under load
  long long t_clk0 = 0, t_ns0 = 0;
  if (A.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    t_clk0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ns0));
  }
#endif

  // ---- work assignment: a "unit" = 128*kCta consecutive rows; group g (CTA or CTA pair) owns units g, g+G, ...
  const int64_t n_units = A.n_tiles;                      // launch_mlp2 sets n_tiles = ceil(M / (128*kCta))
  const int64_t G = gridDim.x / kCta, gidx = blockIdx.x / kCta;
  const int64_t my_units = (n_units > gidx) ? (n_units - gidx + G - 1) / G : 0;
  const int64_t rounds = (my_units + 1) / 2;
  // first row of this CTA's 128-row tile in slot t of round r (may lie beyond M: fully masked tile)
  auto tile_row0 = [&](int64_t r, int t) -> int64_t { return ((gidx + (2 * r + t) * G) * kCta + rank) * TILE_M; };

  if (warp == WARP_TMA) {
    // ===================== TMA producer ===========================================================
    if (lane == 0 && !(A.debug_flags & 4)) {
      uint32_t slot = 0, phase = 0;
      int tc = 0;
      const uint8_t* bias_base = A.w + P.bias_blocks_off;
      for (int64_t r = 0; r < rounds; ++r) {
        for (int l = 0; l < P.n_layers; ++l) {
          const int nh = P.L[l].n_halves, nks_tot = P.L[l].n_pe_ks + P.L[l].n_h_ks, bidx = P.bias_block_idx[l];
          const uint8_t* lbase = A.w + lay_off[l];
          const Runs R = layer_runs(P, l);
          // K-step i of the layer (PE.., hidden.., bias), 128-neuron half h -> its 4 KB block in the packed buffer
          auto src = [&](int h, int i) -> const uint8_t* {
            return (i >= nks_tot) ? bias_base + (size_t)(bidx + h) * KS_BYTES : lbase + ((size_t)h * nks_tot + i) * KS_BYTES;
          };
          for (int t = 0; t < 2; ++t) {
            int i0 = 0;     // first K-step of the run inside the layer
            for (int run = 0; run < R.n; ++run) {
              const int n_run = R.nks[run];
              for (int k0 = 0; k0 < ((n_run + 3) & ~3); k0 += KS2) {   // ring stages of the run, padded to an even count
                const int nks = max(0, min(KS2, n_run - k0));
                const uint32_t dst = s_ring + slot * SL.stage_bytes;
                ptx::mbar_wait(w_empty(slot), phase ^ 1);
                V2_TRACE(3, tc, 6000 + l * 100 + t * 50 + k0);
                if (nks == 0 || ((A.debug_flags & 1) && r > 0)) {
                  ptx::mbar_arrive(w_full(slot));            // padding stage (or bring-up experiment: no re-streaming)
                } else if (kCta == 1) {
                  if (nh == 2) {
                    // N=256 operand image per K-step: [panel][256 rows][16 B]  (two 2 KB pieces per half and K-step)
                    ptx::mbar_arrive_expect_tx(w_full(slot), (uint32_t)nks * 2 * KS_BYTES);
                    for (int k = 0; k < nks; ++k)
                      for (int h = 0; h < 2; ++h)
                        for (int pnl = 0; pnl < 2; ++pnl)
                          ptx::bulk_g2s(dst + k * 8192 + pnl * 4096 + h * 2048, src(h, i0 + k0 + k) + pnl * 2048, 2048, w_full(slot));
                  } else {
                    ptx::mbar_arrive_expect_tx(w_full(slot), (uint32_t)nks * KS_BYTES);
                    for (int k = 0; k < nks; ++k) ptx::bulk_g2s(dst + k * KS_BYTES, src(0, i0 + k0 + k), KS_BYTES, w_full(slot));
                  }
                } else {
                  if (nh == 2) {
                    // this CTA's 128-neuron half, one 4 KB block per K-step
                    ptx::mbar_arrive_expect_tx(w_full(slot), (uint32_t)nks * KS_BYTES);
                    for (int k = 0; k < nks; ++k) ptx::bulk_g2s(dst + k * KS_BYTES, src((int)rank, i0 + k0 + k), KS_BYTES, w_full(slot));
                  } else {
                    // N=128 layer: this CTA's 64 rows of both panels -> [panel][64 rows][16 B] per K-step
                    ptx::mbar_arrive_expect_tx(w_full(slot), (uint32_t)nks * (KS_BYTES / 2));
                    for (int k = 0; k < nks; ++k)
                      for (int pnl = 0; pnl < 2; ++pnl)
                        ptx::bulk_g2s(dst + k * 2048 + pnl * 1024, src(0, i0 + k0 + k) + pnl * 2048 + rank * 1024, 1024, w_full(slot));
                  }
                }
                if (++slot == (uint32_t)NST) { slot = 0; phase ^= 1; }
              }
              i0 += n_run;
            }
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    if (kCta == 2 && !leader) {
      // ===================== peer relay: tell the leader when this CTA's half of a stage has landed ====
      if (lane == 0 && !(A.debug_flags & 4)) {
        uint32_t slot = 0, phase = 0;
        for (int64_t r = 0; r < rounds; ++r)
          for (int l = 0; l < P.n_layers; ++l) {
            const Runs R = layer_runs(P, l);
            int nst = 0;
            for (int run = 0; run < R.n; ++run) nst += ((R.nks[run] + 3) & ~3) / KS2;
            for (int i = 0; i < 2 * nst; ++i) {
              ptx::mbar_wait(w_full(slot), phase);
              mbar_arrive_cluster(mapa(w_fullp(slot), 0));
              if (++slot == (uint32_t)NST) { slot = 0; phase ^= 1; }
            }
          }
      }
    } else {
      // ===================== MMA issuer (leader) ====================================================
      // Warp-uniform control flow (all lanes wait on the barriers and track the ring), ONE elected lane issues.
      // A single warp retires roughly one dependent instruction per 5-7 cycles and the tensor pipe buffers only
      // ~3-4 MMAs behind the issuing thread, so the instruction count between MMAs decides whether the pipe stays
      // fed.  Hence: a layer is one contiguous run of K-steps consumed in super-stages of 4 K-steps (two adjacent
      // ring stages) whose MMAs and stage releases come from ONE asm block; the next (layer, slot) item's
      // parameters are fetched and its readiness barriers are waited while the current item's MMAs are still queued;
      // no divergent code around tcgen05 (ptxas wraps every tcgen05 instruction of a divergent thread in an
      // ELECT / R2UR.BROADCAST loop).
      const uint64_t dfix = ptx::smem_desc(0, 0, 128);                      // SBO = 128, version; address / LBO added below
      const uint64_t a_slot[2] = {dfix | ((uint64_t)((2048 >> 4) & 0x3FFF) << 16) | (uint64_t)((s_a[0] & 0x3FFFFu) >> 4),
                                  dfix | ((uint64_t)((2048 >> 4) & 0x3FFF) << 16) | (uint64_t)((s_a[1] & 0x3FFFFu) >> 4)};
      const uint64_t ring0 = dfix | (uint64_t)((s_ring & 0x3FFFFu) >> 4);
      const uint32_t stage16 = SL.stage_bytes >> 4;
      const bool ringed = !(A.debug_flags & 4);
      uint32_t slot = 0, phase = 0;     // slot is always even here: super-stages take ring stages (slot, slot+1)
      uint32_t n_a[2] = {0, 0};
      int tc = 0;
      const int64_t n_items = rounds * P.n_layers * 2;
      // readiness of item (r, l, t): previous layer's activations written + accumulator drained; a tile's first layer
      // needs its PE panels instead of (round 0) / in addition to (the previous tile's last drain) that
      auto wait_item = [&](int64_t r, int l, int t) {
        if (l == 0) ptx::mbar_wait(pe_ready(t), (uint32_t)(r & 1));
        if (!(r == 0 && l == 0)) { ptx::mbar_wait(a_ready(t), n_a[t] & 1); ++n_a[t]; }
      };
      int64_t r = 0;
      int l = 0, t = 0;
      uint4 q0 = lay_issue[0], q1 = lay_issue[1];
      if (n_items > 0) wait_item(0, 0, 0);
      for (int64_t item = 0; item < n_items; ++item) {
        // the item after this one
        int64_t rn = r; int ln = l, tn = t ^ 1;
        if (tn == 0) { if (++ln == P.n_layers) { ln = 0; ++rn; } }
        const uint4 nq0 = lay_issue[2 * ln], nq1 = lay_issue[2 * ln + 1];
        bool prewaited = (item + 1 >= n_items);
        if (lane == 0) V2_TRACE(0, tc, 11000 + l * 100 + t * 50);
        ptx::tc_fence_after();
        if (lane == 0) V2_TRACE(0, tc, 1000 + l * 100 + t * 50);
        const uint32_t idesc = q0.x, d = tmem + 256u * t;
        const uint64_t bstep = q0.y, b_lbo = q0.z;
        uint32_t acc = 0;
        if (ringed) {
          // production path: one or two asm calls per run (the split leaves room for the next item's readiness wait)
          for (int run = 0; run < (int)q0.w; ++run) {
            uint64_t a = a_slot[t] + (uint64_t)(run == 0 ? q1.x : q1.z);
            const uint32_t nks = (run == 0 ? q1.y : q1.w);
            uint32_t n_super = nks >> 2;
            const uint32_t tail = nks & 3u;
            const bool last_run = (run == (int)q0.w - 1);
            // all lanes advance the ring cursor arithmetically; the issuing lane's asm works on a private copy
            auto advance = [&](uint32_t stages) {
              slot += stages;
              while (slot >= (uint32_t)NST) { slot -= (uint32_t)NST; phase ^= 1; }
            };
            if (last_run && n_super > 1) {
              if (ptx::elect_one()) {
                uint32_t s2 = slot, p2 = phase;
                issue_run<kCta>(s2, p2, d, idesc, a, ring0 | b_lbo, bstep, stage16, acc, n_super - 1, 0, w_full(0), (uint32_t)NST, 0);
              }
              __syncwarp();
              advance(2 * (n_super - 1));
              a += (uint64_t)(n_super - 1) * 4 * (KS_BYTES >> 4);
              n_super = 1; acc = 1;
            }
            if (last_run && !prewaited) { wait_item(rn, ln, tn); prewaited = true; }
            if (ptx::elect_one()) {
              uint32_t s2 = slot, p2 = phase;
              issue_run<kCta>(s2, p2, d, idesc, a, ring0 | b_lbo, bstep, stage16, acc, n_super, tail, w_full(0), (uint32_t)NST, last_run ? d_full(t) : 0u);
            }
            __syncwarp();
            advance(2 * (n_super + (tail ? 1u : 0u)));
            acc = 1;
          }
        } else {
        for (int run = 0; run < (int)q0.w; ++run) {
          uint64_t a = a_slot[t] + (uint64_t)(run == 0 ? q1.x : q1.z);
          for (int left = (int)(run == 0 ? q1.y : q1.w); left > 0; left -= 4) {
            // before the item's last full super-stage: make sure the NEXT item may start (its epilogue had this whole
            // item's MMA phase to finish), so that nothing but a commit separates the two items' MMAs
            if (!prewaited && run == (int)q0.w - 1 && left <= 5) { wait_item(rn, ln, tn); prewaited = true; }
            if (ringed) {
              ptx::mbar_wait(w_full(slot), phase);
              ptx::mbar_wait(w_full(slot + 1), phase);
              if (kCta == 2) { ptx::mbar_wait(w_fullp(slot), phase); ptx::mbar_wait(w_fullp(slot + 1), phase); }
            }
            if (lane == 0) V2_TRACE(0, tc, 2000 + l * 100 + t * 50 + left);
            const uint64_t b0 = (ring0 + (uint64_t)(slot * stage16)) | b_lbo, b1 = b0 + stage16;
            if (ptx::elect_one()) {
              if (left >= 4) {
                if (ringed) issue_super<kCta, true>(d, a, b0, b1, idesc, acc, bstep, w_empty(slot), w_empty(slot + 1));
                else issue_super<kCta, false>(d, a, b0, b1, idesc, acc, bstep, 0, 0);
              } else if (left == 1) {
                issue_ss<kCta, 1>(d, a, b0, idesc, acc, bstep);
                if (ringed) { mma_commit_g<kCta>(w_empty(slot)); mma_commit_g<kCta>(w_empty(slot + 1)); }
              } else {
                issue_ss<kCta, 2>(d, a, b0, idesc, acc, bstep);
                if (ringed) mma_commit_g<kCta>(w_empty(slot));
                if (left == 3) issue_ss<kCta, 1>(d, a + 512, b1, idesc, 1, bstep);
                if (ringed) mma_commit_g<kCta>(w_empty(slot + 1));
              }
            }
            __syncwarp();
            if (lane == 0) V2_TRACE(0, tc, 9000 + l * 100 + t * 50 + left);
            acc = 1;
            a += 4 * (KS_BYTES >> 4);
            slot += 2;
            if (slot == (uint32_t)NST) { slot = 0; phase ^= 1; }
          }
        }
          if (ptx::elect_one()) mma_commit_g<kCta>(d_full(t));
          __syncwarp();
        }
        if (lane == 0) V2_TRACE(0, tc, 10000 + l * 100 + t * 50);
        r = rn; l = ln; t = tn; q0 = nq0; q1 = nq1;
      }
    }
  } else {
    // ===================== epilogue warps (16) ====================================================
    // TMEM lane quarter q = warp % 4 (rows 32q..32q+31, one row per thread); column group c = warp / 4.
    const int q = warp & 3, c = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    // barriers the MMA issuer waits on live in the leader CTA
    const uint32_t a_ready_r[2] = {kCta == 2 ? mapa(a_ready(0), 0) : a_ready(0), kCta == 2 ? mapa(a_ready(1), 0) : a_ready(1)};
    const uint32_t pe_ready_r[2] = {kCta == 2 ? mapa(pe_ready(0), 0) : pe_ready(0), kCta == 2 ? mapa(pe_ready(1), 0) : pe_ready(1)};
    auto signal = [&](uint32_t bar_r) {
      __syncwarp();
      if (lane == 0) { if (kCta == 2) mbar_arrive_cluster(bar_r); else ptx::mbar_arrive(bar_r); }
    };
    int l_pe_last = 0;
    for (int l = 0; l < P.n_layers; ++l) if (P.L[l].n_pe_ks > 0) l_pe_last = l;
    uint32_t n_d[2] = {0, 0};
    int tc = 0;

    if (rounds > 0) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        pe_rows(A, smem + SL.a[t] + A_HID_OFF - P.pe_ks * KS_BYTES, tile_row0(0, t) + row, row, c);
        ptx::fence_proxy_async_smem();
        signal(pe_ready_r[t]);
      }
    }
    for (int64_t r = 0; r < rounds; ++r) {
      float alpha_acc[2] = {0.f, 0.f};
      for (int l = 0; l < P.n_layers; ++l) {
        const int epi = P.L[l].epi, flags = P.L[l].flags;
        const bool last_layer = (l == P.n_layers - 1);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int64_t g = tile_row0(r, t) + row;
          const bool valid = g < A.M;
          const int64_t gc = valid ? g : (A.M - 1);
          uint8_t* a_hid = smem + SL.a[t] + A_HID_OFF;
          float head[MAX_OUT_CH];
#pragma unroll
          for (int ch = 0; ch < MAX_OUT_CH; ++ch) head[ch] = 0.f;

          if (A.debug_flags & 2) {
            ptx::mbar_wait(d_full(t), n_d[t] & 1); ++n_d[t];   // bring-up experiment: empty epilogue
            ptx::tc_fence_after();
          } else if (epi == EPI_VIEWS) {
            // N=128 layer: this warp owns 32 columns; per-ray view bias prefetched before the accumulator wait
            const float* vbrow = A.viewbias + (gc / A.vb_div) * 128 + 32 * c;
            float4 vb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) vb[i] = *reinterpret_cast<const float4*>(vbrow + 4 * i);
            ptx::mbar_wait(d_full(t), n_d[t] & 1); ++n_d[t];
            ptx::tc_fence_after();
            uint32_t rr[32];
            ptx::tmem_ld32(tmem + lane_addr + 256u * t + 32u * c, rr);
            ptx::tmem_ld_wait();
            float* val = reinterpret_cast<float*>(rr);
            const float* rw = consts + P.rgb_w_off + 32 * c;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = vb[i >> 2];
              val[i] = fmaxf(val[i] + b4.x, 0.f); val[i + 1] = fmaxf(val[i + 1] + b4.y, 0.f);
              val[i + 2] = fmaxf(val[i + 2] + b4.z, 0.f); val[i + 3] = fmaxf(val[i + 3] + b4.w, 0.f);
              const float4 w0 = *reinterpret_cast<const float4*>(rw + i);
              const float4 w1 = *reinterpret_cast<const float4*>(rw + 128 + i);
              const float4 w2 = *reinterpret_cast<const float4*>(rw + 256 + i);
              head[0] = fmaf(val[i], w0.x, head[0]); head[0] = fmaf(val[i + 1], w0.y, head[0]);
              head[0] = fmaf(val[i + 2], w0.z, head[0]); head[0] = fmaf(val[i + 3], w0.w, head[0]);
              head[1] = fmaf(val[i], w1.x, head[1]); head[1] = fmaf(val[i + 1], w1.y, head[1]);
              head[1] = fmaf(val[i + 2], w1.z, head[1]); head[1] = fmaf(val[i + 3], w1.w, head[1]);
              head[2] = fmaf(val[i], w2.x, head[2]); head[2] = fmaf(val[i + 1], w2.y, head[2]);
              head[2] = fmaf(val[i + 2], w2.z, head[2]); head[2] = fmaf(val[i + 3], w2.w, head[2]);
            }
          } else {
            ptx::mbar_wait(d_full(t), n_d[t] & 1); ++n_d[t];
            ptx::tc_fence_after();
            if (lane == 0 && q == 0 && c < 2) V2_TRACE(1 + c, tc, 3000 + l * 100 + t * 50);
            uint32_t r0[32], r1[32];
            ptx::tmem_ld32(tmem + lane_addr + 256u * t + 64u * c, r0);
            ptx::tmem_ld32(tmem + lane_addr + 256u * t + 64u * c + 32u, r1);
            ptx::tmem_ld_wait();
            if (lane == 0 && q == 0 && c < 2) V2_TRACE(1 + c, tc, 7000 + l * 100 + t * 50);
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
              float* val = reinterpret_cast<float*>(hc ? r1 : r0);
              const int n0 = 64 * c + 32 * hc;
              // (the layer bias was added by the tensor pipe: bias K-step)
              uint32_t pk[16];
              if (epi == EPI_RELU_A && flags == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = pack_bf16_relu(val[2 * i], val[2 * i + 1]);
              } else {
                if (epi != EPI_LINEAR_A) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) val[i] = fmaxf(val[i], 0.f);
                }
                if (flags & FLAG_ALPHA) {
                  const float* aw = consts + P.alpha_w_off + n0;
                  float acc = alpha_acc[t];
#pragma unroll
                  for (int i = 0; i < 32; i += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(aw + i);
                    acc = fmaf(val[i], w4.x, acc); acc = fmaf(val[i + 1], w4.y, acc);
                    acc = fmaf(val[i + 2], w4.z, acc); acc = fmaf(val[i + 3], w4.w, acc);
                  }
                  alpha_acc[t] = acc;
                }
                if (flags & FLAG_OUTHEAD) {
#pragma unroll
                  for (int ch = 0; ch < MAX_OUT_CH; ++ch) {
                    if (ch < P.out_ch) {
                      const float* ow = consts + P.out_w_off + ch * 256 + n0;
#pragma unroll
                      for (int i = 0; i < 32; ++i) head[ch] = fmaf(val[i], ow[i], head[ch]);
                    }
                  }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = ptx::pack_bf16(val[2 * i], val[2 * i + 1]);
              }
              if (epi != EPI_RELU_HEAD) {
                // 32 columns = 4 panels; a warp's 32 rows are 512 contiguous bytes of a panel (conflict-free)
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                  *reinterpret_cast<uint4*>(a_hid + ((n0 >> 3) + q4) * 2048 + row * 16) =
                      make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
              }
            }
          }
          if (!last_layer) {
            // hand the slot back to the MMA issuer: activations visible to the async proxy, accumulator drained
            if (lane == 0 && q == 0 && c < 2) V2_TRACE(1 + c, tc, 8000 + l * 100 + t * 50);
            ptx::fence_proxy_async_smem();
            ptx::tc_fence_before();
            signal(a_ready_r[t]);
            if (lane == 0 && q == 0 && c < 2) V2_TRACE(1 + c, tc, 4000 + l * 100 + t * 50);
            if (l == l_pe_last && r + 1 < rounds) {
              // every MMA that reads this slot's PE panels has completed: encode the slot's NEXT tile now
              pe_rows(A, smem + SL.a[t] + A_HID_OFF - P.pe_ks * KS_BYTES, tile_row0(r + 1, t) + row, row, c);
              ptx::fence_proxy_async_smem();
              signal(pe_ready_r[t]);
            }
          } else {
            // ---- last layer: combine the 4 column groups' partial head sums (scratch = the slot's hidden panels,
            // free now: their only readers, this layer's MMAs, have completed) and write the row
            float* xch = reinterpret_cast<float*>(a_hid);
            if (c > 0) {
              float* x = xch + ((c - 1) * TILE_M + row) * (MAX_OUT_CH + 1);
              x[0] = alpha_acc[t];
#pragma unroll
              for (int ch = 0; ch < MAX_OUT_CH; ++ch) x[1 + ch] = head[ch];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            if (c == 0) {
              float al = alpha_acc[t];
#pragma unroll
              for (int g2 = 1; g2 < 4; ++g2) {
                const float* x = xch + ((g2 - 1) * TILE_M + row) * (MAX_OUT_CH + 1);
                al += x[0];
#pragma unroll
                for (int ch = 0; ch < MAX_OUT_CH; ++ch) head[ch] += x[1 + ch];
              }
              if (valid) {
                float* o = A.out + g * (int64_t)A.out_stride;
                if (P.use_viewdirs) {
                  const float o0 = head[0] + consts[P.rgb_b_off + 0], o1 = head[1] + consts[P.rgb_b_off + 1];
                  const float o2 = head[2] + consts[P.rgb_b_off + 2], o3 = al + consts[P.alpha_b_off];
                  if (A.out_stride == 4 && (reinterpret_cast<uintptr_t>(A.out) & 15) == 0) *reinterpret_cast<float4*>(o) = make_float4(o0, o1, o2, o3);
                  else { o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3; }
                } else {
#pragma unroll
                  for (int ch = 0; ch < MAX_OUT_CH; ++ch)
                    if (ch < P.out_ch) o[ch] = head[ch] + consts[P.out_b_off + ch];
                }
              }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // scratch is overwritten by the next tile's layer 0
            // the accumulator is drained: the slot's next tile may start (its layer 0 waits a_ready AND pe_ready)
            ptx::tc_fence_before();
            signal(a_ready_r[t]);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
#if 1
  if (A.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    long long t_ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ns1));
    A.trace[4 * 256 * 2 - 2] = clock64() - t_clk0;      // SM cycles of this CTA
    A.trace[4 * 256 * 2 - 1] = t_ns1 - t_ns0;           // nanoseconds
  }
#endif
  if constexpr (kCta == 2) cluster_sync_all();
  if (warp == WARP_TMA) tmem_dealloc_g<kCta>(tmem, 512);
}

}  // namespace v2
