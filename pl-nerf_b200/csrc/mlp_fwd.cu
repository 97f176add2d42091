// Fused NeRF MLP forward for sm_100a (B200): positional encoding + 8x256 trunk + heads in ONE
// persistent, warp-specialised kernel.   Replaces run_network / NeRF.forward
// (reference run_plnerf.py:78-92, run_nerf_helpers.py:105-128).
//
// Design (DESIGN.md section "K1"):
//   * one CTA per SM, each CTA walks 128-sample tiles (rows of the flattened [rays*S] samples);
//   * activations never leave the SM: the 128x256 hidden state lives in TENSOR MEMORY as bf16
//     (tcgen05.mma "TS" form, A operand from TMEM); only the 63-wide positional encoding tile is a
//     shared-memory ("SS") operand;
//   * weights are pre-packed (plnerf_pack_weights) into the exact shared-memory image the MMA wants
//     (K-major 8x16B core-matrix panels, no swizzle) in streaming order and pulled through a ring
//     of 16 KB stages with 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx);
//   * each 256-wide layer is issued as two N=128 halves with separate fp32 accumulators
//     (TMEM cols [0,128) and [128,256)); the epilogue of half a (bias, ReLU, bf16 pack, tcgen05.st
//     into the *other* activation buffer) runs while the tensor pipe works on half b, and the next
//     layer's first K-steps only depend on half a -- the tensor pipe does not wait for epilogues;
//   * heads with tiny N run on CUDA cores inside the epilogue from the fp32 accumulators:
//     alpha (256->1), rgb (128->3), output_linear (256->output_ch); the viewdir part of
//     views_linears is constant per ray and enters as a per-ray fp32 bias (k_viewbias);
//   * precision: PLNERF_PREC_BF16 = one MMA per product; PLNERF_PREC_BF16X3 = activations and
//     weights split hi+lo in bf16 and three MMAs (hi*hi + lo*hi + hi*lo), ~2^-16 relative product
//     error, used for the 1e-4 parity gate against the fp32 reference.
//
// TMEM map (512 columns): D_a [0,128) D_b [128,256); bf16 mode: A0 [256,384) A1 [384,512)
// (double buffered 128x256 bf16 = 128 columns each); bf16x3 mode: A_hi [256,384) A_lo [384,512).
#include <cuda_bf16.h>
#include <math.h>

#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "composite.cuh"
#include "ops.cuh"
#include "umma.cuh"

namespace plnerf {

namespace {

constexpr int TILE_M = 128;
constexpr int KS_BYTES = 4096;      // one K=16 step of one 128-neuron half (2 panels x 128 rows x 16 B)
constexpr int KS_PER_STAGE = 8;
constexpr int STAGE_BYTES = KS_PER_STAGE * KS_BYTES;  // 32 KB
constexpr int MAX_LAYERS = PLNERF_MAX_DEPTH + 2;
constexpr int MAX_PE_KS = 4;        // input_ch <= 64
constexpr int PE_TILE_BYTES = MAX_PE_KS * KS_BYTES;   // 16 KB
constexpr int MAX_CONST_FLOATS = 4096;                // biases + small heads staged in smem
constexpr int MAX_STAGES = 6;
constexpr int NGRP = 4;               // epilogue column groups: each owns 4/NGRP of the 32-column chunks of a half
constexpr int CHUNKS_PER_GRP = 4 / NGRP;
constexpr int NUM_EPI_THREADS = 128 * NGRP;
// warps 0..15 epilogue, then the TMA producer (+TMEM alloc) and the MMA issuer: the SM's warp arbiter prefers the
// HIGHEST warp id, so the two latency-critical single-thread roles sit above the epilogue warps they share SMSPs with
constexpr int WARP_TMA = NUM_EPI_THREADS / 32, WARP_MMA = WARP_TMA + 1;
constexpr int NUM_THREADS = 64 + NUM_EPI_THREADS;
constexpr int MAX_OUT_CH = 8;

enum { EPI_RELU_A = 0, EPI_LINEAR_A = 1, EPI_VIEWS = 2, EPI_RELU_HEAD = 3 };
enum { FLAG_ALPHA = 1, FLAG_OUTHEAD = 2 };

struct LayerPlan {
  int16_t n_pe_ks;   // K=16 steps taken from the PE tile (shared memory operand)
  int16_t n_h_ks;    // K=16 steps taken from the hidden state (TMEM operand)
  int8_t n_halves;   // 2 for N=256, 1 for N=128
  int8_t epi;
  int8_t flags;
  int8_t pad;
  int32_t bias_off;  // float offset into the tail/const block
  // training: stash tensor this layer's epilogue writes (forward: its output activations as a
  // weight-gradient operand; dgrad: its output gradient), relu mask it writes (forward) / applies (dgrad)
  int16_t stash_idx, mask_idx;
  // packing sources (host/pack kernel only)
  const float* W;
  int32_t ldw;
  int32_t pe_col0;   // source column of the PE part (or -1)
  int32_t h_col0;    // source column of the hidden part (or -1)
  int32_t transposed;  // dgrad: B'[k][n] = W[n][h_col0 + k]
};

// Per-tile layout of the training stash (all tensors as MN-major 8x8 core-matrix tiles
// [m-half(2)][col8(width/8)][m8(8)][8 rows x 8 cols], i.e. exactly the wgrad MMA operand image).
constexpr int MAX_STASH = 16;
struct TrainLayout {
  int32_t n_in, n_dy, n_mask;
  int32_t in_width[MAX_STASH], in_off[MAX_STASH];     // forward activations (weight-gradient B operands)
  int32_t dy_width[MAX_STASH], dy_off[MAX_STASH];     // output gradients (weight-gradient A operands)
  int32_t mask_words[MAX_STASH], mask_off[MAX_STASH]; // relu masks: [word][row] uint32, word offsets
  int32_t in_tile_bytes, dy_tile_bytes, mask_tile_words;
  int32_t idx_pe, idx_h0, idx_feat, idx_hv, idx_dir;  // In tensors
  int32_t dy_h0, dy_feat, dy_views, dy_head;           // dY tensors (dy_head: the network output's gradient, 16 wide)
  int32_t mask_views;                                  // mask index of the views layer (trunk layer l -> l)
};

struct NetPlan {
  int32_t n_layers;
  LayerPlan L[MAX_LAYERS];
  int32_t input_ch, input_ch_views, out_ch, use_viewdirs, pe_ks, precision;
  // tail (fp32) offsets, in floats
  int32_t alpha_w_off, alpha_b_off, rgb_w_off, rgb_b_off, out_w_off, out_b_off, dirw_off, views_b_off;
  int32_t const_floats;  // prefix of the tail staged in shared memory (everything but dirw)
  int32_t tail_floats;
  int64_t weight_bytes;  // bf16 stream
  int32_t is_dgrad;      // plan describes the backward (input-gradient) chain
  int32_t D;
};

// A (layer, half) streams its K-steps in stages of <= KS_PER_STAGE; the PE K-steps (shared-memory A
// operand) and the hidden K-steps (TMEM A operand) are staged separately so no stage mixes them.
struct StageInfo { int is_pe, k0, nks; };   // k0 = first K-step of the stage inside its segment
__host__ __device__ inline int stages_of(int n_pe, int n_h) {
  return (n_pe + KS_PER_STAGE - 1) / KS_PER_STAGE + (n_h + KS_PER_STAGE - 1) / KS_PER_STAGE;
}
__host__ __device__ inline StageInfo stage_info(int n_pe, int n_h, int i) {
  const int npe_st = (n_pe + KS_PER_STAGE - 1) / KS_PER_STAGE;
  StageInfo si;
  if (i < npe_st) { si.is_pe = 1; si.k0 = i * KS_PER_STAGE; si.nks = min(KS_PER_STAGE, n_pe - si.k0); }
  else { si.is_pe = 0; si.k0 = (i - npe_st) * KS_PER_STAGE; si.nks = min(KS_PER_STAGE, n_h - si.k0); }
  return si;
}

// Walks desc -> plan.  Returns 0 or an error code.
int build_plan(const plnerf_net_desc* d, int precision, const plnerf_net_params* p, NetPlan* out) {
  if (!d) { set_error("net desc is null"); return PLNERF_E_BADARG; }
  if (d->W != 256) { set_error("only W=256 is implemented by the tcgen05 kernel (got %d)", d->W); return PLNERF_E_UNSUPPORTED; }
  if (d->D < 1 || d->D > PLNERF_MAX_DEPTH) { set_error("D=%d out of range", d->D); return PLNERF_E_UNSUPPORTED; }
  if (d->input_ch < 1 || d->input_ch > 64) { set_error("input_ch=%d not in [1,64]", d->input_ch); return PLNERF_E_UNSUPPORTED; }
  if (d->use_viewdirs && (d->input_ch_views < 1 || d->input_ch_views > 64)) { set_error("input_ch_views=%d not in [1,64]", d->input_ch_views); return PLNERF_E_UNSUPPORTED; }
  if (!d->use_viewdirs && (d->output_ch < 1 || d->output_ch > MAX_OUT_CH)) { set_error("output_ch=%d not in [1,%d]", d->output_ch, MAX_OUT_CH); return PLNERF_E_UNSUPPORTED; }
  if (precision != PLNERF_PREC_BF16 && precision != PLNERF_PREC_BF16X3) { set_error("bad precision %d", precision); return PLNERF_E_BADARG; }
  if (d->n_skips < 0 || d->n_skips > PLNERF_MAX_DEPTH) { set_error("bad n_skips"); return PLNERF_E_BADARG; }
  NetPlan& P = *out;
  memset(&P, 0, sizeof(P));
  P.input_ch = d->input_ch; P.input_ch_views = d->use_viewdirs ? d->input_ch_views : 0;
  P.out_ch = d->use_viewdirs ? 4 : d->output_ch; P.use_viewdirs = d->use_viewdirs;
  P.pe_ks = (d->input_ch + 15) / 16; P.precision = precision; P.D = d->D;
  auto is_skip = [&](int i) { for (int k = 0; k < d->n_skips; ++k) if (d->skips[k] == i) return true; return false; };
  int nl = 0, foff = 0;
  for (int i = 0; i < d->D; ++i) {
    LayerPlan& L = P.L[nl++];
    const bool first = (i == 0), skip_in = (i > 0) && is_skip(i - 1);
    L.n_pe_ks = (first || skip_in) ? P.pe_ks : 0;
    L.n_h_ks = first ? 0 : 16;
    L.n_halves = 2;
    L.epi = EPI_RELU_A; L.flags = 0;
    L.bias_off = foff; foff += 256;
    L.stash_idx = (int16_t)(1 + i); L.mask_idx = (int16_t)i;   // In tensors: 0 = PE, 1+l = h_l
    L.W = p ? p->pts_w[i] : nullptr;
    L.ldw = first ? d->input_ch : (skip_in ? 256 + d->input_ch : 256);
    L.pe_col0 = (first || skip_in) ? 0 : -1;
    L.h_col0 = first ? -1 : (skip_in ? d->input_ch : 0);
    if (i == d->D - 1) {
      // the reference concatenates [input_pts, h] after a skip layer even when it is the last one
      if (is_skip(i)) { set_error("a skip at the last trunk layer is not supported"); return PLNERF_E_UNSUPPORTED; }
      if (d->use_viewdirs) L.flags = FLAG_ALPHA;
      else { L.epi = EPI_RELU_HEAD; L.flags = FLAG_OUTHEAD; }
    }
  }
  if (d->use_viewdirs) {
    LayerPlan& F = P.L[nl++];
    F.n_pe_ks = 0; F.n_h_ks = 16; F.n_halves = 2; F.epi = EPI_LINEAR_A; F.flags = 0;
    F.bias_off = foff; foff += 256;
    F.stash_idx = (int16_t)(1 + d->D); F.mask_idx = -1;
    F.W = p ? p->feature_w : nullptr; F.ldw = 256; F.pe_col0 = -1; F.h_col0 = 0;
    LayerPlan& V = P.L[nl++];
    V.n_pe_ks = 0; V.n_h_ks = 16; V.n_halves = 1; V.epi = EPI_VIEWS; V.flags = 0;
    V.bias_off = foff; P.views_b_off = foff; foff += 128;
    V.stash_idx = (int16_t)(2 + d->D); V.mask_idx = (int16_t)d->D;
    V.W = p ? p->views_w : nullptr; V.ldw = 256 + d->input_ch_views; V.pe_col0 = -1; V.h_col0 = 0;
    P.alpha_w_off = foff; foff += 256;          // every block starts 16-byte aligned (float4 loads)
    P.alpha_b_off = foff; foff += 4;
    P.rgb_w_off = foff; foff += 3 * 128;
    P.rgb_b_off = foff; foff += 4;
    foff = (foff + 3) & ~3;
    P.const_floats = foff;
    P.dirw_off = foff; foff += 128 * d->input_ch_views;
  } else {
    P.out_w_off = foff; foff += d->output_ch * 256;
    P.out_b_off = foff; foff += (d->output_ch + 3) & ~3;
    foff = (foff + 3) & ~3;
    P.const_floats = foff;
  }
  P.tail_floats = (foff + 3) & ~3;
  P.n_layers = nl;
  if (P.const_floats > MAX_CONST_FLOATS) { set_error("const block too large"); return PLNERF_E_UNSUPPORTED; }
  // issue_tile relies on it: hidden K-steps come in full stages of KS_PER_STAGE, PE K-steps fit one stage
  for (int l = 0; l < nl; ++l)
    if (P.L[l].n_h_ks % KS_PER_STAGE != 0 || P.L[l].n_pe_ks > MAX_PE_KS) { set_error("unsupported layer shape"); return PLNERF_E_UNSUPPORTED; }
  int64_t wb = 0;
  const int nsplit = (precision == PLNERF_PREC_BF16X3) ? 2 : 1;
  for (int l = 0; l < nl; ++l) wb += (int64_t)P.L[l].n_halves * (P.L[l].n_pe_ks + P.L[l].n_h_ks) * KS_BYTES * nsplit;
  P.weight_bytes = wb;
  return PLNERF_OK;
}

// Training stash layout for a (viewdirs) network.
int build_train_layout(const plnerf_net_desc* d, TrainLayout* out) {
  if (d->D + 4 > MAX_STASH) { set_error("network too deep for the training stash"); return PLNERF_E_UNSUPPORTED; }
  if (!d->use_viewdirs && (d->output_ch < 4 || d->output_ch > MAX_OUT_CH)) { set_error("training needs output_ch in [4,%d] (got %d)", MAX_OUT_CH, d->output_ch); return PLNERF_E_UNSUPPORTED; }
  TrainLayout& T = *out;
  memset(&T, 0, sizeof(T));
  const int pe_w = ((d->input_ch + 15) / 16) * 16;
  int n = 0, off = 0;
  auto add_in = [&](int w) { T.in_width[n] = w; T.in_off[n] = off; off += w * 256; return n++; };
  T.idx_pe = add_in(pe_w);
  T.idx_h0 = n;
  for (int l = 0; l < d->D; ++l) add_in(256);
  T.idx_feat = T.idx_hv = T.idx_dir = -1;          // networks without view directions end at h_{D-1} (output_linear head)
  if (d->use_viewdirs) {
    T.idx_feat = add_in(256);
    T.idx_hv = add_in(128);
    T.idx_dir = add_in(32);
  }
  T.n_in = n; T.in_tile_bytes = off;
  n = 0; off = 0;
  auto add_dy = [&](int w) { T.dy_width[n] = w; T.dy_off[n] = off; off += w * 256; return n++; };
  T.dy_h0 = 0;
  for (int l = 0; l < d->D; ++l) add_dy(256);
  T.dy_feat = T.dy_views = -1;
  if (d->use_viewdirs) {
    T.dy_feat = add_dy(256);
    T.dy_views = add_dy(128);
  }
  T.dy_head = add_dy(16);        // [g_rgb(3), g_alpha, 0 ...]: N operand of the alpha / rgb (or output_linear) head weight gradients
  T.n_dy = n; T.dy_tile_bytes = off;
  n = 0; off = 0;
  for (int l = 0; l < d->D; ++l) { T.mask_words[n] = 8; T.mask_off[n] = off; off += 8 * 128; ++n; }
  T.mask_views = -1;
  if (d->use_viewdirs) { T.mask_views = n; T.mask_words[n] = 4; T.mask_off[n] = off; off += 4 * 128; ++n; }
  T.n_mask = n; T.mask_tile_words = off;
  return PLNERF_OK;
}

// Input-gradient chain as a plan for the same fused kernel: activations = gradients (bf16, TMEM),
// weights = W^T.  Layer order: views (128->256, d_feature), feature (256->256, +alpha rank-1, mask D-1),
// trunk D-1 .. 1 (mask l-1).  Trunk layer 0 needs no input gradient.
int build_dgrad_plan(const plnerf_net_desc* d, const plnerf_net_params* p, NetPlan* out) {
  NetPlan fwd;
  int rc = build_plan(d, PLNERF_PREC_BF16, nullptr, &fwd);
  if (rc) return rc;
  NetPlan& P = *out;
  memset(&P, 0, sizeof(P));
  P.input_ch = d->input_ch; P.input_ch_views = d->use_viewdirs ? d->input_ch_views : 0; P.out_ch = fwd.out_ch;
  P.use_viewdirs = d->use_viewdirs;
  P.pe_ks = 0; P.precision = PLNERF_PREC_BF16; P.is_dgrad = 1; P.D = d->D;
  auto is_skip = [&](int i) { for (int k = 0; k < d->n_skips; ++k) if (d->skips[k] == i) return true; return false; };
  int nl = 0;
  // (without view directions the chain starts at dh_{D-1} = (g_out . W_out) * relu'(h_{D-1}), formed by the prologue)
  if (d->use_viewdirs) {  // views: d_feature = d_hv . W_views[:, :256]
    LayerPlan& L = P.L[nl++];
    L.n_pe_ks = 0; L.n_h_ks = 8; L.n_halves = 2; L.epi = EPI_LINEAR_A; L.flags = 0; L.bias_off = 0;
    L.stash_idx = (int16_t)(d->D); L.mask_idx = -1;                 // dY tensors: l = dY_l, D = dY_feat, D+1 = dY_views
    L.W = p ? p->views_w : nullptr; L.ldw = 256 + d->input_ch_views; L.pe_col0 = -1; L.h_col0 = 0; L.transposed = 1;
  }
  if (d->use_viewdirs) {  // feature: dh_{D-1} = d_feature . W_feat + g_alpha * w_alpha, masked by relu(D-1)
    LayerPlan& L = P.L[nl++];
    L.n_pe_ks = 0; L.n_h_ks = 16; L.n_halves = 2; L.epi = EPI_RELU_A; L.flags = FLAG_ALPHA; L.bias_off = 0;
    L.stash_idx = (int16_t)(d->D - 1); L.mask_idx = (int16_t)(d->D - 1);
    L.W = p ? p->feature_w : nullptr; L.ldw = 256; L.pe_col0 = -1; L.h_col0 = 0; L.transposed = 1;
  }
  for (int l = d->D - 1; l >= 1; --l) {  // trunk layer l: dh_{l-1} = dY_l . W_l[:, hidden part], masked by relu(l-1)
    LayerPlan& L = P.L[nl++];
    const bool skip_in = is_skip(l - 1);
    L.n_pe_ks = 0; L.n_h_ks = 16; L.n_halves = 2; L.epi = EPI_RELU_A; L.flags = 0; L.bias_off = 0;
    L.stash_idx = (int16_t)(l - 1); L.mask_idx = (int16_t)(l - 1);
    L.W = p ? p->pts_w[l] : nullptr; L.ldw = skip_in ? 256 + d->input_ch : 256; L.pe_col0 = -1;
    L.h_col0 = skip_in ? d->input_ch : 0; L.transposed = 1;
  }
  P.n_layers = nl;
  // const block: rgb_w [3][128] and alpha_w [256] (same offsets as the forward plan so the tail is shared)
  P.alpha_w_off = fwd.alpha_w_off; P.alpha_b_off = fwd.alpha_b_off; P.rgb_w_off = fwd.rgb_w_off; P.rgb_b_off = fwd.rgb_b_off;
  P.out_w_off = fwd.out_w_off; P.out_b_off = fwd.out_b_off;
  P.const_floats = fwd.const_floats; P.tail_floats = fwd.tail_floats;
  if (nl == 0) { set_error("training a one-layer network without view directions is not supported"); return PLNERF_E_UNSUPPORTED; }
  int64_t wb = 0;
  for (int l = 0; l < nl; ++l) wb += (int64_t)P.L[l].n_halves * P.L[l].n_h_ks * KS_BYTES;
  P.weight_bytes = wb;
  return PLNERF_OK;
}

// =============================================================================================
// weight packing
// =============================================================================================
struct PackArgs {
  NetPlan plan;
  uint8_t* dst;   // bf16 stream
  float* tail;    // fp32 tail
  plnerf_net_params prm;
};

// one block (256 threads) per 4 KB K-step block; thread u -> 16-byte unit (panel = u/128, row = u%128)
__device__ __forceinline__ void pack_weights_block(const PackArgs& a, const int block) {
  const NetPlan& P = a.plan;
  const int nsplit = (P.precision == PLNERF_PREC_BF16X3) ? 2 : 1;
  int blk = block;
  // locate (layer, half, stage, rep, j)
  int l = 0, h = 0;
  for (; l < P.n_layers; ++l) {
    const int per_half = (P.L[l].n_pe_ks + P.L[l].n_h_ks) * nsplit;
    const int per_layer = per_half * P.L[l].n_halves;
    if (blk < per_layer) { h = blk / per_half; blk -= h * per_half; break; }
    blk -= per_layer;
  }
  if (l >= P.n_layers) return;
  const LayerPlan& L = P.L[l];
  // within a half: stages (stage_info), each stage = [hi blocks][lo blocks]
  int rep = 0, ks = 0;
  for (int st = 0;; ++st) {
    const StageInfo si = stage_info(L.n_pe_ks, L.n_h_ks, st);
    if (blk < si.nks * nsplit) {
      rep = blk / si.nks;
      ks = (si.is_pe ? 0 : L.n_pe_ks) + si.k0 + (blk - rep * si.nks);
      break;
    }
    blk -= si.nks * nsplit;
  }
  const int u = threadIdx.x;
  const int panel = u >> 7, row = u & 127;
  const int n = h * 128 + row;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    int col = -1;
    if (ks < L.n_pe_ks) {
      const int c = ks * 16 + panel * 8 + e;
      if (c < P.input_ch) col = L.pe_col0 + c;
    } else {
      col = L.h_col0 + (ks - L.n_pe_ks) * 16 + panel * 8 + e;
    }
    if (L.transposed) v[e] = L.W[(int64_t)((ks - L.n_pe_ks) * 16 + panel * 8 + e) * L.ldw + L.h_col0 + n];   // B'[n][k] = W[k][col0+n]
    else v[e] = (col >= 0) ? L.W[(int64_t)n * L.ldw + col] : 0.0f;
    if (rep == 1) v[e] = v[e] - ptx::bf16_round(v[e]);   // lo part
  }
  uint4 q;
  q.x = ptx::pack_bf16(v[0], v[1]); q.y = ptx::pack_bf16(v[2], v[3]);
  q.z = ptx::pack_bf16(v[4], v[5]); q.w = ptx::pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(a.dst + (int64_t)block * KS_BYTES + (size_t)u * 16) = q;
}
__global__ void __launch_bounds__(256) k_pack_weights(const __grid_constant__ PackArgs a) { pack_weights_block(a, blockIdx.x); }

__device__ __forceinline__ void pack_tail_elem(const PackArgs& a, const int i) {
  const NetPlan& P = a.plan;
  if (i >= P.tail_floats) return;
  float v = 0.f;
  // biases
  int trunk_layers = P.use_viewdirs ? P.n_layers - 2 : P.n_layers;
  bool done = false;
  for (int l = 0; l < P.n_layers && !done; ++l) {
    const int nb = P.L[l].n_halves * 128;
    const int o = P.L[l].bias_off;
    if (i >= o && i < o + nb) {
      const float* b = (l < trunk_layers) ? a.prm.pts_b[l] : (l == trunk_layers ? a.prm.feature_b : a.prm.views_b);
      v = b[i - o]; done = true;
    }
  }
  if (!done) {
    if (P.use_viewdirs) {
      if (i >= P.alpha_w_off && i < P.alpha_w_off + 256) v = a.prm.alpha_w[i - P.alpha_w_off];
      else if (i == P.alpha_b_off) v = a.prm.alpha_b[0];
      else if (i >= P.rgb_w_off && i < P.rgb_w_off + 384) v = a.prm.rgb_w[i - P.rgb_w_off];
      else if (i >= P.rgb_b_off && i < P.rgb_b_off + 3) v = a.prm.rgb_b[i - P.rgb_b_off];
      else if (i >= P.dirw_off && i < P.dirw_off + 128 * P.input_ch_views) {
        const int k = i - P.dirw_off;
        const int n = k / P.input_ch_views, jv = k - n * P.input_ch_views;
        v = a.prm.views_w[(int64_t)n * (256 + P.input_ch_views) + 256 + jv];
      }
    } else {
      if (i >= P.out_w_off && i < P.out_w_off + P.out_ch * 256) v = a.prm.output_w[i - P.out_w_off];
      else if (i >= P.out_b_off && i < P.out_b_off + P.out_ch) v = a.prm.output_b[i - P.out_b_off];
    }
  }
  a.tail[i] = v;
}
__global__ void __launch_bounds__(256) k_pack_tail(const __grid_constant__ PackArgs a) { pack_tail_elem(a, blockIdx.x * 256 + threadIdx.x); }

// Every packed copy a training step needs, in ONE launch (six separate launches cost ~45 us of a 1.8 ms iteration): per
// network the forward stream, its fp32 tail and the transposed stream of the input-gradient chain.  Blocks are dealt to jobs
// by ranges; the (up to four) PackArgs ride in the kernel's parameter space (sm_70+: 32 KB).
constexpr int MAX_PACK_ARGS = 4, MAX_PACK_JOBS = 6;
struct PackTrainArgs {
  int32_t n_jobs;
  int32_t blk0[MAX_PACK_JOBS + 1];   // first block of job j (blk0[n_jobs] = grid size)
  int8_t kind[MAX_PACK_JOBS];        // 0: weight stream (one block per K-step block), 1: tail (256 floats per block)
  int8_t arg[MAX_PACK_JOBS];
  PackArgs args[MAX_PACK_ARGS];
};
__global__ void __launch_bounds__(256) k_pack_train(const __grid_constant__ PackTrainArgs t) {
  int j = 0;
  while (j + 1 < t.n_jobs && (int)blockIdx.x >= t.blk0[j + 1]) ++j;
  const int b = (int)blockIdx.x - t.blk0[j];
  const PackArgs& a = t.args[t.arg[j]];
  if (t.kind[j] == 0) pack_weights_block(a, b);
  else pack_tail_elem(a, b * 256 + (int)threadIdx.x);
}

// =============================================================================================
// per-ray view bias:  vb[r][n] = views_b[n] + sum_j Wv[n][256+j] * gamma_dir(viewdir_r)_j   (fp32)
// (the viewdir columns of views_linears[0], run_nerf_helpers.py:117-121, are constant per ray)
// =============================================================================================
constexpr int VB_RAYS = 4;   // rays per block iteration
__global__ void __launch_bounds__(128) k_viewbias(const float* __restrict__ tail, int views_b_off, int dirw_off,
                                                  int icv, int multires_views, const float* __restrict__ rays,
                                                  int stride, const float* __restrict__ x_emb, int x_ld, int x_col0,
                                                  int64_t n, float* __restrict__ vb, float* __restrict__ dirpe) {
  // thread t owns output neuron t: its icv weights live in registers for the whole (grid-strided) ray loop
  __shared__ float emb[VB_RAYS][64];
  const int t = threadIdx.x;
  float w[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) w[j] = (j < icv) ? tail[dirw_off + t * icv + j] : 0.f;
  const float b = tail[views_b_off + t];
  for (int64_t r0 = (int64_t)blockIdx.x * VB_RAYS; r0 < n; r0 += (int64_t)gridDim.x * VB_RAYS) {
    // the rays' direction encodings: thread (rr, j) = (t / 32, t % 32) and (t / 32, 32 + t % 32)
    for (int e = t; e < VB_RAYS * 64; e += 128) {
      const int rr = e >> 6, j = e & 63;
      const int64_t r = r0 + rr;
      float v = 0.f;
      if (r < n && j < icv) {
        if (x_emb) {
          v = x_emb[r * (int64_t)x_ld + x_col0 + j];
        } else {
          const float* vd = rays + r * (int64_t)stride + (stride - 3);
          if (multires_views < 0 || j < 3) v = vd[j % 3];
          else {
            const int k = (j - 3) / 6, rem = (j - 3) % 6;
            const float a = vd[rem % 3] * exp2f((float)k);
            v = (rem < 3) ? sinf(a) : cosf(a);
          }
        }
      }
      emb[rr][j] = v;
      if (dirpe && r < n && j < 32) dirpe[r * 32 + j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < VB_RAYS; ++rr) {
      const int64_t r = r0 + rr;
      if (r < n) {
        float acc = b;
#pragma unroll
        for (int j = 0; j < 64; ++j) if (j < icv) acc = fmaf(w[j], emb[rr][j], acc);
        vb[r * 128 + t] = acc;
      }
    }
    __syncthreads();
  }
}

// Ray setup of render_rays in ONE launch: the stratified depths (run_plnerf.py:683-705, same arithmetic as k_stratified_z)
// and the per-ray view bias of BOTH networks (coarse and fine differ only in their weights).  Block = 128 threads = the 128
// neurons of views_linears[0]; each neuron's view-direction weights of both networks stay in registers over a grid-strided
// loop of VB_RAYS rays.
struct RaySetupArgs {
  const float *tail_c, *tail_f;          // fp32 tails of the two packed networks (tail_f may equal tail_c)
  int views_b_off, dirw_off, icv, multires_views;
  const float* rays; int stride; int64_t n;
  float *vb_c, *vb_f;                    // [n,128] each (vb_f null: one network serves both passes)
  float* dirpe;                          // [n,32] or null: the (zero-padded) direction encoding itself (training stash input)
  // depths
  int Ns, lindisp, perturb; const float* t_rand; uint64_t seed, ray0; float* z;
  int rpi;                               // rays per block iteration (<= RS_RAYS)
};
__device__ __forceinline__ float setup_linspace01(int i, int n, float step) {
  return (i < n / 2) ? step * (float)i : fmaf(-step, (float)(n - 1 - i), 1.0f);   // torch.linspace(0,1,n)
}
__device__ __forceinline__ float setup_base_z(float near, float far, float t, int lindisp) {
  const float omt = __fsub_rn(1.0f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
  return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(__fdiv_rn(1.0f, near), omt), __fmul_rn(__fdiv_rn(1.0f, far), t)));
}
constexpr int RS_RAYS = 16;  // most rays per block iteration (two block-wide barriers per iteration; the weights stay in registers)
__global__ void __launch_bounds__(128) k_ray_setup(const __grid_constant__ RaySetupArgs a) {
  __shared__ __align__(16) float emb[RS_RAYS][32];
  const int t = threadIdx.x;
  const int icv = a.icv;                   // <= 27 on this path (multires_views <= 4)
  float wc[27], wf[27];
#pragma unroll
  for (int j = 0; j < 27; ++j) {
    wc[j] = (j < icv) ? a.tail_c[a.dirw_off + t * icv + j] : 0.f;
    wf[j] = (a.vb_f && j < icv) ? a.tail_f[a.dirw_off + t * icv + j] : 0.f;
  }
  const float bc = a.tail_c[a.views_b_off + t], bf = a.vb_f ? a.tail_f[a.views_b_off + t] : 0.f;
  const float step = (a.Ns > 1) ? __fdiv_rn(1.0f, (float)(a.Ns - 1)) : 0.0f;
  const int gpr = (a.Ns + 3) >> 2;         // groups of four consecutive depths per ray (one Philox block each)
  const int rpi = a.rpi;
  for (int64_t r0 = (int64_t)blockIdx.x * rpi; r0 < a.n; r0 += (int64_t)gridDim.x * rpi) {
    for (int e = t; e < rpi * 32; e += 128) {
      const int rr = e >> 5, j = e & 31;
      const int64_t r = r0 + rr;
      float v = 0.f;
      if (r < a.n && j < icv) {
        const float* vd = a.rays + r * (int64_t)a.stride + (a.stride - 3);
        if (a.multires_views < 0 || j < 3) v = vd[j % 3];
        else {
          const int k = (j - 3) / 6, rem = (j - 3) % 6;
          const float ang = vd[rem % 3] * exp2f((float)k);
          v = (rem < 3) ? sinf(ang) : cosf(ang);
        }
      }
      emb[rr][j] = v;
      if (a.dirpe && r < a.n) a.dirpe[r * 32 + j] = v;
    }
    // this block's depths: RS_RAYS rays x Ns samples (k_stratified_z's arithmetic, one rounding per reference op), four
    // consecutive samples per thread: their draws are the four words of one Philox block, the neighbouring base depths are shared
    for (int e = t; e < rpi * gpr; e += 128) {
      const int rr = e / gpr, g = e - rr * gpr;
      const int64_t r = r0 + rr;
      if (r >= a.n) continue;
      const float near = a.rays[r * a.stride + 6], far = a.rays[r * a.stride + 7];
      const int i0 = 4 * g;
      float zb[6];                           // base depths of samples i0 - 1 .. i0 + 4
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const int i = i0 - 1 + q;
        zb[q] = (i >= 0 && i < a.Ns) ? setup_base_z(near, far, setup_linspace01(i, a.Ns, step), a.lindisp) : 0.f;
      }
      uint32_t w4[4] = {0u, 0u, 0u, 0u};
      if (a.perturb && !a.t_rand) philox4x32(a.seed, a.ray0 + (uint64_t)r, RNG_STREAM_TRAND, (uint32_t)g, w4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + q;
        if (i >= a.Ns) break;
        const float zi = zb[q + 1];
        float z = zi;
        if (a.perturb) {
          float lower = zi, upper = zi;
          if (i > 0) lower = __fmul_rn(0.5f, __fadd_rn(zi, zb[q]));
          if (i < a.Ns - 1) upper = __fmul_rn(0.5f, __fadd_rn(zb[q + 2], zi));
          const float tr = a.t_rand ? a.t_rand[r * (int64_t)a.Ns + i] : (float)(w4[q] >> 8) * (1.0f / 16777216.0f);
          z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), tr));
        }
        a.z[r * (int64_t)a.Ns + i] = z;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < rpi; ++rr) {
      const int64_t r = r0 + rr;
      if (r < a.n) {
        float acc_c = bc, acc_f = bf;
        const float4* e4 = reinterpret_cast<const float4*>(emb[rr]);     // (broadcast reads, 16 bytes at a time; zero beyond icv)
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const float4 ev = e4[q];
          const float e[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = 4 * q + i;
            if (j < 27) { acc_c = fmaf(wc[j], e[i], acc_c); acc_f = fmaf(wf[j], e[i], acc_f); }
          }
        }
        a.vb_c[r * 128 + t] = acc_c;
        if (a.vb_f) a.vb_f[r * 128 + t] = acc_f;
      }
    }
    __syncthreads();
  }
}

// =============================================================================================
// the fused MLP kernel
// =============================================================================================
struct MlpArgs {
  NetPlan plan;
  const uint8_t* w;       // packed bf16 stream
  const float* tail;      // fp32 consts
  // fused-query inputs
  const float* rays; int stride; const float* z; int S; int multires;
  // embedded-input mode
  const float* x_emb; int x_ld;
  const float* viewbias;  // [rows/vb_div, 128]
  int vb_div;
  int64_t M;              // total rows
  int64_t n_tiles;
  float* out; int out_stride;
  // training (MODE 2 = forward + stash, MODE 3 = input-gradient chain)
  TrainLayout tl;
  uint8_t* in_stash;       // [tiles][tl.in_tile_bytes]   forward activations
  uint8_t* dy_stash;       // [tiles][tl.dy_tile_bytes]   output gradients
  uint32_t* masks;         // [tiles][tl.mask_tile_words] relu masks
  const float* dirpe;      // [rays][32] fp32 dir encoding (padded), MODE 2
  const float* g_raw; int g_stride;   // MODE 3: upstream gradient of the network output [rows, >=4]
  float *d_rgb_b, *d_alpha_b;         // MODE 3: bias gradients of the two heads (column sums of g_raw), accumulated
  int n_stages;
  // k_mlp3: whole units (rays) per CTA; fused quadrature (composite.cuh) of every completed ray inside the kernel
  int64_t cta_units;
  int64_t unit_rows;       // rows per unit: S (a ray) with the fused quadrature, else one tile pair
  int fuse_comp, comp_mode;
  int skip_out;            // k_mlp3 with fuse_comp: the rows' outputs are not written to `out` (no retraw)
  CompositeArgs comp;
  long long* trace;  // debug timeline buffer (null in production)
  int debug_flags;   // developer library only (PLNERF_DBG): bring-up experiments selected by PLNERF_DEBUG_FLAGS
};

// Bring-up switches exist only in the developer library (-DPLNERF_DEBUG): in the product build they fold to `false`.
#ifdef PLNERF_DEBUG
#define PLNERF_DBG(bit) ((A.debug_flags & (bit)) != 0)
#else
#define PLNERF_DBG(bit) false
#endif

struct SmemLayout {
  uint32_t pe_hi, pe_lo, ring, consts, xch, prog, prog2, vb, bars;  // byte offsets
  uint32_t total;
};
constexpr int VB_SMEM_RAYS = 3;
__host__ __device__ inline SmemLayout smem_layout(int n_stages) {
  SmemLayout s;
  s.pe_hi = 0;
  s.pe_lo = PE_TILE_BYTES;
  s.ring = 2 * PE_TILE_BYTES;
  s.consts = s.ring + (uint32_t)n_stages * (uint32_t)STAGE_BYTES;
  s.xch = s.consts + MAX_CONST_FLOATS * 4;
  s.prog = s.xch + (NGRP - 1) * TILE_M * (MAX_OUT_CH + 1) * 4;
  s.prog2 = s.prog + 8 * 128;  // flattened MMA stage program (<= 128 entries)
  s.vb = s.prog2 + 16 * 128;   // (prog2: the same program, pre-decoded for the asm issue loop, 16 bytes per stage)
  s.bars = s.vb + VB_SMEM_RAYS * 128 * 4;   // vb: the tile's per-ray view-bias rows (<= 3 rays per 128 samples)
  s.total = s.bars + 512;
  return s;
}

// ---- packed epilogue math -----------------------------------------------------------------------
// (x0,x1) += (b0,b1) as ONE instruction (Blackwell packed fp32 add)
__device__ __forceinline__ void add2(float& x0, float& x1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%0, %1};\n\t"
      "mov.b64 rb, {%2, %3};\n\t"
      "add.rn.f32x2 rc, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(x0), "+f"(x1)
      : "f"(b0), "f"(b1));
}
// {lo, hi} -> bf16x2 with ReLU fused into the conversion
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// ---- tight MMA issue: 8 (or 4) consecutive K=16 steps of one 128-neuron half in ONE asm block ------
// A from TMEM: step j reads columns a + 8j; B from the ring slot: descriptor + j*(4096>>4).
#define PLNERF_TS_STEP(PRED) \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], [ad], bd, %3, " PRED ";\n\t" \
  "add.u64 bd, bd, 256;\n\tadd.u32 ad, ad, 8;\n\t"
#define PLNERF_TS_STEP_X3(PRED) \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], [ad], bd, %3, " PRED ";\n\t" \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], [al], bd, %3, pt;\n\t" \
  "add.u64 bd, bd, 256;\n\tadd.u32 ad, ad, 8;\n\tadd.u32 al, al, 8;\n\t"
#define PLNERF_TS_PROLOG \
  "{\n\t.reg .pred p, pt;\n\t.reg .b64 bd;\n\t.reg .b32 ad, al;\n\t" \
  "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 pt, %4, %4;\n\tmov.b64 bd, %2;\n\tmov.b32 ad, %1;\n\tmov.b32 al, %5;\n\t"

template <bool X3PAIR>
__device__ __forceinline__ void issue_ts8(uint32_t d, uint32_t a, uint32_t a_lo, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (X3PAIR) {
    asm volatile(PLNERF_TS_PROLOG
                 PLNERF_TS_STEP_X3("p") PLNERF_TS_STEP_X3("pt") PLNERF_TS_STEP_X3("pt") PLNERF_TS_STEP_X3("pt")
                 PLNERF_TS_STEP_X3("pt") PLNERF_TS_STEP_X3("pt") PLNERF_TS_STEP_X3("pt") PLNERF_TS_STEP_X3("pt") "}"
                 ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(a_lo) : "memory");
  } else {
    asm volatile(PLNERF_TS_PROLOG
                 PLNERF_TS_STEP("p") PLNERF_TS_STEP("pt") PLNERF_TS_STEP("pt") PLNERF_TS_STEP("pt")
                 PLNERF_TS_STEP("pt") PLNERF_TS_STEP("pt") PLNERF_TS_STEP("pt") PLNERF_TS_STEP("pt") "}"
                 ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(a_lo) : "memory");
  }
}

// ---- a whole tile's stage program from ONE asm block --------------------------------------------------------------
// Measured on this kernel: the MMA-issuing warp retires one dependent instruction per ~6.5 cycles and its instruction
// stream -- not the tensor pipe, not the epilogue -- bounded the tile time (adding ~40 instructions per batch cost 13%).
// So a tile's whole stage program -- batch barriers (weights, activations), program entry fetch, B descriptor of the
// ring slot, 8 TMEM-operand or <= 4 shared-operand MMAs, the slot release, the accumulator-full commit, the ring cursor
// -- runs as ONE PTX loop whose registers ptxas keeps on the uniform datapath; C++ makes one call per tile.
// Program entry (16 bytes): {accumulator tmem address, A operand (tmem address | low word of the shared-memory
// descriptor of its first K-step), flags (bit 0 first-of-accumulator, bit 1 shared-memory A, bit 2 commit, bits 8-15
// K-steps), accumulator-full barrier}.
#define PLNERF_B_MMA_TS(PRED) \
  "tcgen05.mma.cta_group::1.kind::f16 [ed], [ea], bd, %11, " PRED ";\n\t" \
  "add.u64 bd, bd, 256;\n\tadd.u32 ea, ea, 8;\n\t"
// Ring / dependency state of the issuing thread, carried across tiles.
struct IssueState { uint32_t slot, batch, uses0, uses1, waited0, waited1, pend0, pend1; };
// Flags of a program entry: 1 first-of-accumulator, 2 shared-memory A, 4 commit accumulator-full, 8 / 16 batch waits for
// a_ready[a] / a_ready[b], 32 first entry of a batch, 64 / 128 the batch completes accumulator half a / b, bits 8-15 K-steps.
__device__ __forceinline__ void issue_tile(IssueState& st, uint32_t prog_addr, uint32_t n_entries, uint32_t idesc, uint64_t ring_desc,
                                           uint32_t desc_hi, uint32_t wempty0, uint32_t n_stages, uint32_t bfull0, uint32_t aready0) {
  asm volatile(
      "{\n\t.reg .pred p, p2, pacc, pt, ppe, pl;\n\t"
      ".reg .b32 sl, n, pa, ed, ea, ef, eb, t, k, wb, bt, u0, u1, w0, w1, q0, q1;\n\t.reg .b64 bd, so, ad;\n\t"
      "mov.b32 sl, %0;\n\tmov.b32 bt, %1;\n\tmov.b32 u0, %2;\n\tmov.b32 u1, %3;\n\tmov.b32 w0, %4;\n\tmov.b32 w1, %5;\n\t"
      "mov.b32 q0, %6;\n\tmov.b32 q1, %7;\n\t"
      "mov.b32 n, %9;\n\tmov.b32 pa, %8;\n\tsetp.eq.b32 pt, sl, sl;\n\t"
      "LOOP:\n\t"
      "ld.shared.v4.u32 {ed, ea, ef, eb}, [pa];\n\t"
      "and.b32 t, ef, 32;\n\tsetp.eq.b32 p, t, 0;\n\t@p bra NOBATCH;\n\t"
      // ---- batch start: account the previous batch's accumulator commits, wait for this batch's weights ...
      "add.u32 u0, u0, q0;\n\tadd.u32 u1, u1, q1;\n\t"
      "shr.u32 q0, ef, 6;\n\tand.b32 q0, q0, 1;\n\tshr.u32 q1, ef, 7;\n\tand.b32 q1, q1, 1;\n\t"
      "and.b32 t, bt, 7;\n\tshl.b32 t, t, 3;\n\tadd.u32 wb, t, %15;\n\tshr.u32 t, bt, 3;\n\tand.b32 t, t, 1;\n\t"
      "WB:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [wb], t;\n\t@!p bra WB;\n\t"
      "add.u32 bt, bt, 1;\n\t"
      // ---- ... and for the activations it reads (every outstanding phase of a_ready[a] / a_ready[b])
      "and.b32 t, ef, 8;\n\tsetp.eq.b32 p, t, 0;\n\t@p bra NA0;\n\t"
      "LA0:\n\tsetp.ge.u32 p2, w0, u0;\n\t@p2 bra NA0;\n\tand.b32 t, w0, 1;\n\t"
      "WA0:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%16], t;\n\t@!p bra WA0;\n\tadd.u32 w0, w0, 1;\n\tbra LA0;\n\t"
      "NA0:\n\t"
      "and.b32 t, ef, 16;\n\tsetp.eq.b32 p, t, 0;\n\t@p bra NA1;\n\t"
      "LA1:\n\tsetp.ge.u32 p2, w1, u1;\n\t@p2 bra NA1;\n\tand.b32 t, w1, 1;\n\t"
      "WA1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%16+8], t;\n\t@!p bra WA1;\n\tadd.u32 w1, w1, 1;\n\tbra LA1;\n\t"
      "NA1:\n\t"
      "tcgen05.fence::after_thread_sync;\n\t"
      "NOBATCH:\n\t"
      "mul.wide.u32 so, sl, 2048;\n\tadd.u64 bd, so, %10;\n\t"
      "and.b32 t, ef, 1;\n\tsetp.eq.b32 pacc, t, 0;\n\t"
      "and.b32 t, ef, 2;\n\tsetp.ne.b32 ppe, t, 0;\n\t"
      "@ppe bra PE;\n\t"
      PLNERF_B_MMA_TS("pacc") PLNERF_B_MMA_TS("pt") PLNERF_B_MMA_TS("pt") PLNERF_B_MMA_TS("pt")
      PLNERF_B_MMA_TS("pt") PLNERF_B_MMA_TS("pt") PLNERF_B_MMA_TS("pt") PLNERF_B_MMA_TS("pt")
      "bra COMMIT;\n\t"
      "PE:\n\t"
      "mov.b64 ad, {ea, %12};\n\tshr.u32 k, ef, 8;\n\tand.b32 k, k, 255;\n\t"
      "PEL:\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [ed], ad, bd, %11, pacc;\n\t"
      "setp.eq.b32 pacc, sl, sl;\n\tadd.u64 ad, ad, 256;\n\tadd.u64 bd, bd, 256;\n\t"
      "sub.u32 k, k, 1;\n\tsetp.ne.b32 p, k, 0;\n\t@p bra PEL;\n\t"
      "COMMIT:\n\t"
      "shl.b32 wb, sl, 3;\n\tadd.u32 wb, wb, %13;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [wb];\n\t"
      "and.b32 t, ef, 4;\n\tsetp.ne.b32 pl, t, 0;\n\t"
      "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [eb];\n\t"
      "add.u32 sl, sl, 1;\n\tsetp.eq.u32 p, sl, %14;\n\t@p mov.b32 sl, 0;\n\t"
      "add.u32 pa, pa, 16;\n\tsub.u32 n, n, 1;\n\tsetp.ne.b32 p, n, 0;\n\t@p bra LOOP;\n\t"
      "mov.b32 %0, sl;\n\tmov.b32 %1, bt;\n\tmov.b32 %2, u0;\n\tmov.b32 %3, u1;\n\tmov.b32 %4, w0;\n\tmov.b32 %5, w1;\n\t"
      "mov.b32 %6, q0;\n\tmov.b32 %7, q1;\n\t}"
      : "+r"(st.slot), "+r"(st.batch), "+r"(st.uses0), "+r"(st.uses1), "+r"(st.waited0), "+r"(st.waited1), "+r"(st.pend0), "+r"(st.pend1)
      : "r"(prog_addr), "r"(n_entries), "l"(ring_desc), "r"(idesc), "r"(desc_hi), "r"(wempty0), "r"(n_stages), "r"(bfull0), "r"(aready0)
      : "memory");
}

// The same interpreter for the hi+lo split mode (bf16x3): a stage occupies TWO ring slots (hi weights, then lo weights),
// each with its own full barrier; slot 1 issues A_hi*B_hi + A_lo*B_hi per K-step, slot 2 A_hi*B_lo.  A_lo sits 128 TMEM
// columns (or PE_TILE_BYTES of shared memory) behind A_hi.  No batch barriers in this mode.
struct IssueStateX3 { uint32_t slot, phase, uses0, uses1, waited0, waited1, pend0, pend1; };
#define PLNERF_X3_TS2(PRED) \
  "tcgen05.mma.cta_group::1.kind::f16 [ed], [ea], bd, %11, " PRED ";\n\t" \
  "tcgen05.mma.cta_group::1.kind::f16 [ed], [el], bd, %11, pt;\n\t" \
  "add.u64 bd, bd, 256;\n\tadd.u32 ea, ea, 8;\n\tadd.u32 el, el, 8;\n\t"
#define PLNERF_X3_TS1 \
  "tcgen05.mma.cta_group::1.kind::f16 [ed], [ea], bd, %11, pt;\n\t" \
  "add.u64 bd, bd, 256;\n\tadd.u32 ea, ea, 8;\n\t"
#define PLNERF_X3_SLOT_WAIT(L) \
  "shl.b32 wb, sl, 3;\n\tadd.u32 wb, wb, %13;\n\t" \
  L ":\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [wb], ph;\n\t@!p bra " L ";\n\t" \
  "tcgen05.fence::after_thread_sync;\n\t" \
  "mul.wide.u32 so, sl, 2048;\n\tadd.u64 bd, so, %10;\n\t"
#define PLNERF_X3_SLOT_DONE \
  "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [wb+48];\n\t" \
  "add.u32 sl, sl, 1;\n\tsetp.eq.u32 p, sl, %14;\n\t@p mov.b32 sl, 0;\n\t@p xor.b32 ph, ph, 1;\n\t"
static_assert(MAX_STAGES == 6 && STAGE_BYTES == 32768 && PE_TILE_BYTES == 16384, "issue_tile_x3 hard-codes barrier / slot / PE strides");
__device__ __forceinline__ void issue_tile_x3(IssueStateX3& st, uint32_t prog_addr, uint32_t n_entries, uint32_t idesc, uint64_t ring_desc,
                                              uint32_t desc_hi, uint32_t wfull0, uint32_t n_stages, uint32_t aready0) {
  asm volatile(
      "{\n\t.reg .pred p, p2, pacc, pt, ppe, pl;\n\t"
      ".reg .b32 sl, ph, n, pa, ed, ea, el, e0, ef, eb, t, k, wb, u0, u1, w0, w1, q0, q1;\n\t.reg .b64 bd, so, ad, al;\n\t"
      "mov.b32 sl, %0;\n\tmov.b32 ph, %1;\n\tmov.b32 u0, %2;\n\tmov.b32 u1, %3;\n\tmov.b32 w0, %4;\n\tmov.b32 w1, %5;\n\t"
      "mov.b32 q0, %6;\n\tmov.b32 q1, %7;\n\t"
      "mov.b32 n, %9;\n\tmov.b32 pa, %8;\n\tsetp.eq.b32 pt, sl, sl;\n\t"
      "LOOP:\n\t"
      "ld.shared.v4.u32 {ed, e0, ef, eb}, [pa];\n\t"
      "and.b32 t, ef, 32;\n\tsetp.eq.b32 p, t, 0;\n\t@p bra NOBATCH;\n\t"
      "add.u32 u0, u0, q0;\n\tadd.u32 u1, u1, q1;\n\t"
      "shr.u32 q0, ef, 6;\n\tand.b32 q0, q0, 1;\n\tshr.u32 q1, ef, 7;\n\tand.b32 q1, q1, 1;\n\t"
      "and.b32 t, ef, 8;\n\tsetp.eq.b32 p, t, 0;\n\t@p bra NA0;\n\t"
      "LA0:\n\tsetp.ge.u32 p2, w0, u0;\n\t@p2 bra NA0;\n\tand.b32 t, w0, 1;\n\t"
      "WA0:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%15], t;\n\t@!p bra WA0;\n\tadd.u32 w0, w0, 1;\n\tbra LA0;\n\t"
      "NA0:\n\t"
      "and.b32 t, ef, 16;\n\tsetp.eq.b32 p, t, 0;\n\t@p bra NA1;\n\t"
      "LA1:\n\tsetp.ge.u32 p2, w1, u1;\n\t@p2 bra NA1;\n\tand.b32 t, w1, 1;\n\t"
      "WA1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%15+8], t;\n\t@!p bra WA1;\n\tadd.u32 w1, w1, 1;\n\tbra LA1;\n\t"
      "NA1:\n\t"
      "NOBATCH:\n\t"
      "and.b32 t, ef, 1;\n\tsetp.eq.b32 pacc, t, 0;\n\t"
      "and.b32 t, ef, 2;\n\tsetp.ne.b32 ppe, t, 0;\n\t"
      // ---------------- ring slot 1 of the stage: hi weights
      PLNERF_X3_SLOT_WAIT("S0")
      "@ppe bra PE0;\n\t"
      "mov.b32 ea, e0;\n\tadd.u32 el, e0, 128;\n\t"
      PLNERF_X3_TS2("pacc") PLNERF_X3_TS2("pt") PLNERF_X3_TS2("pt") PLNERF_X3_TS2("pt")
      PLNERF_X3_TS2("pt") PLNERF_X3_TS2("pt") PLNERF_X3_TS2("pt") PLNERF_X3_TS2("pt")
      "bra C0;\n\t"
      "PE0:\n\t"
      "mov.b64 ad, {e0, %12};\n\tadd.u64 al, ad, 1024;\n\tshr.u32 k, ef, 8;\n\tand.b32 k, k, 255;\n\t"
      "PEL0:\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [ed], ad, bd, %11, pacc;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [ed], al, bd, %11, pt;\n\t"
      "setp.eq.b32 pacc, sl, sl;\n\tadd.u64 ad, ad, 256;\n\tadd.u64 al, al, 256;\n\tadd.u64 bd, bd, 256;\n\t"
      "sub.u32 k, k, 1;\n\tsetp.ne.b32 p, k, 0;\n\t@p bra PEL0;\n\t"
      "C0:\n\t"
      PLNERF_X3_SLOT_DONE
      // ---------------- ring slot 2 of the stage: lo weights
      PLNERF_X3_SLOT_WAIT("S1")
      "@ppe bra PE1;\n\t"
      "mov.b32 ea, e0;\n\t"
      PLNERF_X3_TS1 PLNERF_X3_TS1 PLNERF_X3_TS1 PLNERF_X3_TS1 PLNERF_X3_TS1 PLNERF_X3_TS1 PLNERF_X3_TS1 PLNERF_X3_TS1
      "bra C1;\n\t"
      "PE1:\n\t"
      "mov.b64 ad, {e0, %12};\n\tshr.u32 k, ef, 8;\n\tand.b32 k, k, 255;\n\t"
      "PEL1:\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [ed], ad, bd, %11, pt;\n\t"
      "add.u64 ad, ad, 256;\n\tadd.u64 bd, bd, 256;\n\t"
      "sub.u32 k, k, 1;\n\tsetp.ne.b32 p, k, 0;\n\t@p bra PEL1;\n\t"
      "C1:\n\t"
      "and.b32 t, ef, 4;\n\tsetp.ne.b32 pl, t, 0;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [wb+48];\n\t"
      "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [eb];\n\t"
      "add.u32 sl, sl, 1;\n\tsetp.eq.u32 p, sl, %14;\n\t@p mov.b32 sl, 0;\n\t@p xor.b32 ph, ph, 1;\n\t"
      "add.u32 pa, pa, 16;\n\tsub.u32 n, n, 1;\n\tsetp.ne.b32 p, n, 0;\n\t@p bra LOOP;\n\t"
      "mov.b32 %0, sl;\n\tmov.b32 %1, ph;\n\tmov.b32 %2, u0;\n\tmov.b32 %3, u1;\n\tmov.b32 %4, w0;\n\tmov.b32 %5, w1;\n\t"
      "mov.b32 %6, q0;\n\tmov.b32 %7, q1;\n\t}"
      : "+r"(st.slot), "+r"(st.phase), "+r"(st.uses0), "+r"(st.uses1), "+r"(st.waited0), "+r"(st.waited1), "+r"(st.pend0), "+r"(st.pend1)
      : "r"(prog_addr), "r"(n_entries), "l"(ring_desc), "r"(idesc), "r"(desc_hi), "r"(wfull0), "r"(n_stages), "r"(aready0)
      : "memory");
}

// Positional encoding of one row -> this thread's panels of the PE tile(s).
// store 8 consecutive bf16 columns (16 bytes) of row `row` into an MN-major stash tile
template <bool STREAM = true>
__device__ __forceinline__ void stash_store8(uint8_t* tile, int width, int row, int col8, uint4 v) {
  const int mh = row >> 6, m8 = (row & 63) >> 3, i = row & 7;
  uint8_t* p = tile + ((size_t)((mh * (width >> 3) + col8) * 8 + m8)) * 128 + i * 16;
  // STREAM (the forward's activation tiles): streaming, evict-first store -- written once and read two kernels later, after
  // far more traffic than the L2 holds, they should not displace the packed weights every CTA keeps re-reading.
  // !STREAM (the gradient chain's dY tiles): plain store -- k_wgrad reads them right after the chain, the tail is still in
  // L2 (same-box A/B: streaming dY stores cost +0.9% on the training step, streaming activation stores save 1.3%).
  if (STREAM) asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else *reinterpret_cast<uint4*>(p) = v;
}

// ReLU masks of the training stash: one 32-bit word per (row, 32 columns).  The bit order is chosen for the gradient chain,
// which applies the mask to PACKED bf16 pairs: column 2k sits at bit 7 + k, column 2k + 1 at bit (23 + k) % 32 (k = 0..15), so
// the AND mask of packed word k is one rotate + one byte permute (sign replication).  A bit is set iff the stored bf16
// activation is non-zero (= relu'(x) of the reference, which is 0 at x <= 0).
__device__ __forceinline__ uint32_t relu_mask_from_packed(const uint32_t* pk) {   // pk[16]: post-ReLU (non-negative) bf16 pairs
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t t = pk[k] + 0x7FFF7FFFu;               // bit 15 / bit 31 = (low / high half != 0); halves < 0x8000: no carry
    const int s = (8 - k) & 31;
    const uint32_t c = s ? ((0x80008000u >> s) | (0x80008000u << (32 - s))) : 0x80008000u;
    m |= __funnelshift_r(t, t, s) & c;
  }
  return m;
}
__device__ __forceinline__ uint32_t relu_mask_word(uint32_t m, int k) {           // 0xFFFF per kept half of packed word k
  const uint32_t t = __funnelshift_r(m, m, k);
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(t), "r"(0u), "r"(0xAA88u));   // bytes 0,1 <- sign of byte 0; bytes 2,3 <- sign of byte 2
  return r;
}

template <int MODE>
__device__ __forceinline__ void pe_prologue(const MlpArgs& A, uint8_t* smem, const SmemLayout& SL, int64_t tile, int row,
                                            int grp, const float* p_pre = nullptr) {
  constexpr bool X3 = (MODE == 1);
  constexpr bool STASH = (MODE == 2);
  if (PLNERF_DBG(8)) return;   // bring-up experiment: no encoding at all (garbage inputs)
  const NetPlan& P = A.plan;
  const int64_t g = tile * TILE_M + row;
  const int64_t gc = (g < A.M) ? g : (A.M - 1);
  const int n_panels = 2 * P.pe_ks;
  const int p_lo = grp * n_panels / NGRP, p_hi = (grp + 1) * n_panels / NGRP;   // this column group's panels
  float p[3] = {0.f, 0.f, 0.f};
  const float* xr = nullptr;
  if (A.x_emb) {
    xr = A.x_emb + gc * (int64_t)A.x_ld;
  } else if (p_pre) {
    p[0] = p_pre[0]; p[1] = p_pre[1]; p[2] = p_pre[2];     // sample position fetched earlier in the tile (see the epilogue loop)
  } else {
    const int64_t ray = gc / A.S;
    const float* rp = A.rays + ray * (int64_t)A.stride;
    const float zz = A.z[gc];
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = __fadd_rn(rp[c], __fmul_rn(rp[3 + c], zz));  // o + d*z, two roundings
  }
  const uint32_t turns[3] = {pe_turns(p[0]), pe_turns(p[1]), pe_turns(p[2])};
  // element idx of the encoding: [x,y,z, sin(2^0 p), cos(2^0 p), sin(2^1 p), ...] (run_nerf_helpers.py:45-48)
  auto elem = [&](int idx) -> float {
    if (idx >= P.input_ch) return 0.f;
    if (xr) return xr[idx];
    if (idx < 3) return p[idx];
    const int t = idx - 3, k = t / 6, r = t - 6 * k, c = (r >= 3) ? r - 3 : r;
    return (r >= 3) ? pe_cos(turns[c], k) : pe_sin(turns[c], k);   // angle p * 2^k as an exact shift of p's turn fraction
  };
  // fast path: 63-wide encoding computed in-kernel, 4 column groups of two panels each -> one fully unrolled 16-element
  // group whose values never leave registers (this kernel leaves ~3 KB of L1: a local-memory array costs L2 round trips)
  if (!X3 && (xr == nullptr) && (n_panels == 8) && (NGRP == 4)) {   // (the split mode measured slower with this path: register-bound)
    float v16[16];
    pe_group16<false>(grp, p, turns, P.input_ch, v16);
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      uint4 hi;
      hi.x = ptx::pack_bf16(v16[8 * h2 + 0], v16[8 * h2 + 1]); hi.y = ptx::pack_bf16(v16[8 * h2 + 2], v16[8 * h2 + 3]);
      hi.z = ptx::pack_bf16(v16[8 * h2 + 4], v16[8 * h2 + 5]); hi.w = ptx::pack_bf16(v16[8 * h2 + 6], v16[8 * h2 + 7]);
      *reinterpret_cast<uint4*>(smem + SL.pe_hi + (p_lo + h2) * 2048 + row * 16) = hi;
      if (STASH) stash_store8(A.in_stash + tile * (int64_t)A.tl.in_tile_bytes + A.tl.in_off[A.tl.idx_pe],
                              A.tl.in_width[A.tl.idx_pe], row, p_lo + h2, hi);
    }
  } else
  for (int pnl = p_lo; pnl < p_hi; ++pnl) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = elem(8 * pnl + e);
    uint4 hi;
    hi.x = ptx::pack_bf16(v[0], v[1]); hi.y = ptx::pack_bf16(v[2], v[3]);
    hi.z = ptx::pack_bf16(v[4], v[5]); hi.w = ptx::pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(smem + SL.pe_hi + pnl * 2048 + row * 16) = hi;
    if (STASH) stash_store8(A.in_stash + tile * (int64_t)A.tl.in_tile_bytes + A.tl.in_off[A.tl.idx_pe],
                            A.tl.in_width[A.tl.idx_pe], row, pnl, hi);
    if (X3) {
      uint4 l4;
      l4.x = ptx::pack_bf16(v[0] - ptx::bf16_round(v[0]), v[1] - ptx::bf16_round(v[1]));
      l4.y = ptx::pack_bf16(v[2] - ptx::bf16_round(v[2]), v[3] - ptx::bf16_round(v[3]));
      l4.z = ptx::pack_bf16(v[4] - ptx::bf16_round(v[4]), v[5] - ptx::bf16_round(v[5]));
      l4.w = ptx::pack_bf16(v[6] - ptx::bf16_round(v[6]), v[7] - ptx::bf16_round(v[7]));
      *reinterpret_cast<uint4*>(smem + SL.pe_lo + pnl * 2048 + row * 16) = l4;
    }
  }
  if (STASH && A.plan.use_viewdirs) {
    // the (padded, 32-wide) viewdir encoding of this row's ray as a weight-gradient operand tile:
    // column group g writes its share of the four 8-column blocks
    const float* dp = A.dirpe + (gc / A.vb_div) * 32;
    uint8_t* tile_dir = A.in_stash + tile * (int64_t)A.tl.in_tile_bytes + A.tl.in_off[A.tl.idx_dir];
    for (int c8 = grp * 4 / NGRP; c8 < (grp + 1) * 4 / NGRP; ++c8) {
      uint4 q;
      q.x = ptx::pack_bf16(dp[8 * c8 + 0], dp[8 * c8 + 1]); q.y = ptx::pack_bf16(dp[8 * c8 + 2], dp[8 * c8 + 3]);
      q.z = ptx::pack_bf16(dp[8 * c8 + 4], dp[8 * c8 + 5]); q.w = ptx::pack_bf16(dp[8 * c8 + 6], dp[8 * c8 + 7]);
      stash_store8(tile_dir, 32, row, c8, q);
    }
  }
}

// Input of the gradient chain: d_hv = (g_rgb . W_rgb) * relu'(views) for this thread's columns
// -> bf16 A operand in TMEM (cols COL_A1 + n/2) and the dY_views stash tile (128 columns over the column groups).
__device__ __forceinline__ void dgrad_prologue(const MlpArgs& A, const float* consts, uint32_t tmem_lane_a1, int64_t tile,
                                               int row, int grp, float g_alpha) {
  const NetPlan& P = A.plan;
  const int64_t g = tile * TILE_M + row;
  float gr = 0.f, gg = 0.f, gb = 0.f;
  if (g < A.M) { const float* q = A.g_raw + g * (int64_t)A.g_stride; gr = q[0]; gg = q[1]; gb = q[2]; }
  if (grp == 0) {
    // the heads' weight gradients are k_wgrad items against this 16-column tile [g_rgb, g_alpha, 0...] (rows past the end: 0);
    // their bias gradients are its column sums
    uint8_t* tile_h = A.dy_stash + tile * (int64_t)A.tl.dy_tile_bytes + A.tl.dy_off[A.tl.dy_head];
    stash_store8<false>(tile_h, 16, row, 0, make_uint4(ptx::pack_bf16(gr, gg), ptx::pack_bf16(gb, g_alpha), 0u, 0u));
    stash_store8<false>(tile_h, 16, row, 1, make_uint4(0u, 0u, 0u, 0u));
    float s0 = gr, s1 = gg, s2 = gb, s3 = g_alpha;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(A.d_rgb_b + 0, s0); atomicAdd(A.d_rgb_b + 1, s1); atomicAdd(A.d_rgb_b + 2, s2); atomicAdd(A.d_alpha_b, s3);
    }
  }
  if (!P.use_viewdirs) {
    // output_linear head (run_nerf_helpers.py:126): dh_{D-1} = (g_out . W_out) * relu'(h_{D-1}), 256 wide (g_out: the four
    // channels raw2outputs reads; further output channels get no gradient) -> A operand + the dY_{D-1} stash tile
    const float* ow = consts + P.out_w_off;
    const int oc = P.out_ch < 4 ? P.out_ch : 4;
    const float gv[4] = {gr, gg, gb, g_alpha};
    const uint32_t* mk = A.masks + tile * (int64_t)A.tl.mask_tile_words + A.tl.mask_off[P.D - 1];
    uint8_t* tile_dy = A.dy_stash + tile * (int64_t)A.tl.dy_tile_bytes + A.tl.dy_off[A.tl.dy_h0 + P.D - 1];
    for (int c = grp * CHUNKS_PER_GRP; c < 8; c += NGRP * CHUNKS_PER_GRP) {
      for (int cc = c; cc < c + CHUNKS_PER_GRP; ++cc) {
        const int n0 = 32 * cc;
        const uint32_t m = mk[cc * 128 + row];
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float v0 = 0.f, v1 = 0.f;
          for (int ch = 0; ch < oc; ++ch) { v0 = fmaf(gv[ch], ow[ch * 256 + n0 + i], v0); v1 = fmaf(gv[ch], ow[ch * 256 + n0 + i + 1], v1); }
          pk[i >> 1] = ptx::pack_bf16(v0, v1) & relu_mask_word(m, i >> 1);
        }
        ptx::tmem_st16(tmem_lane_a1 + (uint32_t)(n0 >> 1), pk);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          stash_store8<false>(tile_dy, 256, row, (n0 >> 3) + q4, make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]));
      }
    }
    ptx::tmem_st_wait();
    return;
  }
  const float* rw = consts + P.rgb_w_off;
  const uint32_t* mk = A.masks + tile * (int64_t)A.tl.mask_tile_words + A.tl.mask_off[A.tl.mask_views];
  uint8_t* tile_dy = A.dy_stash + tile * (int64_t)A.tl.dy_tile_bytes + A.tl.dy_off[A.tl.dy_views];
  for (int cc = grp * CHUNKS_PER_GRP; cc < (grp + 1) * CHUNKS_PER_GRP; ++cc) {
    const int n0 = 32 * cc;
    const uint32_t m = mk[(n0 >> 5) * 128 + row];
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float v0 = gr * rw[n0 + i] + gg * rw[128 + n0 + i] + gb * rw[256 + n0 + i];
      const float v1 = gr * rw[n0 + i + 1] + gg * rw[128 + n0 + i + 1] + gb * rw[256 + n0 + i + 1];
      pk[i >> 1] = ptx::pack_bf16(v0, v1) & relu_mask_word(m, i >> 1);
    }
    ptx::tmem_st16(tmem_lane_a1 + (uint32_t)(n0 >> 1), pk);
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4)
      stash_store8<false>(tile_dy, 128, row, (n0 >> 3) + q4, make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]));
  }
  ptx::tmem_st_wait();
}

// debug timeline: (clock, code) pairs for block 0, third tile; region r holds up to 256 events
#if defined(PLNERF_DEBUG) && defined(PLNERF_ENABLE_TRACE)
#define PLNERF_TRACE(region, cnt, code)                                                        \
  do {                                                                                          \
    if (A.trace && blockIdx.x == 0 && trace_on && (cnt) < 256) {                               \
      A.trace[((region) * 256 + (cnt)) * 2] = clock64();                                        \
      A.trace[((region) * 256 + (cnt)) * 2 + 1] = (code);                                       \
      ++(cnt);                                                                                  \
    }                                                                                           \
  } while (0)
#else
#define PLNERF_TRACE(region, cnt, code) do { (void)(cnt); (void)trace_on; } while (0)
#endif

// MODE 0: bf16 forward, 1: bf16x3 forward, 2: bf16 forward + training stash, 3: input-gradient chain
template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_mlp_fwd(const __grid_constant__ MlpArgs A) {
  constexpr bool X3 = (MODE == 1);
  constexpr bool STASH = (MODE == 2);
  constexpr bool DGRAD = (MODE == 3);
  extern __shared__ __align__(1024) uint8_t smem[];
  const NetPlan& P = A.plan;
  const SmemLayout SL = smem_layout(A.n_stages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int nsplit = X3 ? 2 : 1;

  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t s_pe_hi = sbase + SL.pe_hi, s_pe_lo = sbase + SL.pe_lo, s_ring = sbase + SL.ring;
  float* consts = reinterpret_cast<float*>(smem + SL.consts);
  float* xch = reinterpret_cast<float*>(smem + SL.xch);
  const uint32_t s_bars = sbase + SL.bars;
  // barrier map (8 bytes each)
  auto w_full = [&](int s) { return s_bars + 8u * s; };
  auto w_empty = [&](int s) { return s_bars + 8u * (MAX_STAGES + s); };
  const uint32_t d_full0 = s_bars + 8u * (2 * MAX_STAGES);      // [2]
  const uint32_t a_ready0 = s_bars + 8u * (2 * MAX_STAGES + 2);  // [2]
  const uint32_t pe_ready = s_bars + 8u * (2 * MAX_STAGES + 4);
  auto b_full = [&](uint32_t b) { return s_bars + 8u * (2 * MAX_STAGES + 8 + (b & 7u)); };   // batch-level weight barriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SL.bars + 8u * (2 * MAX_STAGES + 6));

  if (threadIdx.x == 0) {
    for (int s = 0; s < A.n_stages; ++s) { ptx::mbar_init(w_full(s), 1); ptx::mbar_init(w_empty(s), 1); }
    ptx::mbar_init(d_full0, 1); ptx::mbar_init(d_full0 + 8, 1);
    ptx::mbar_init(a_ready0, NUM_EPI_THREADS / 32); ptx::mbar_init(a_ready0 + 8, NUM_EPI_THREADS / 32);
    ptx::mbar_init(pe_ready, NUM_EPI_THREADS / 32);
    for (uint32_t b2 = 0; b2 < 8; ++b2) ptx::mbar_init(b_full(b2), 1);
    ptx::fence_mbar_init();
  }
  // per-layer facts of the epilogue, 16 bytes per layer behind the barriers {epi | halves << 8 | flags << 16, bias offset,
  // byte offset of the layer's stash tensor in a tile (forward: activations, gradient chain: dY), word offset of its ReLU
  // mask (-1: none)}: one shared-memory load per layer instead of chained register-indexed constant-bank loads
  static_assert(8 * (2 * MAX_STAGES + 16) + 16 * MAX_LAYERS <= 512, "layer table does not fit behind the barriers");
  const int4* ltab = reinterpret_cast<const int4*>(smem + SL.bars + 8 * (2 * MAX_STAGES + 16));
  if ((int)threadIdx.x < P.n_layers) {
    const LayerPlan& L = P.L[threadIdx.x];
    int4 e;
    e.x = (int)L.epi | ((int)L.n_halves << 8) | ((int)(uint8_t)L.flags << 16);
    e.y = L.bias_off;
    e.z = ((STASH || DGRAD) && L.stash_idx >= 0) ? (DGRAD ? A.tl.dy_off[L.stash_idx] : A.tl.in_off[L.stash_idx]) : 0;
    e.w = ((STASH || DGRAD) && L.mask_idx >= 0) ? A.tl.mask_off[L.mask_idx] : -1;
    const_cast<int4*>(ltab)[threadIdx.x] = e;
  }
  // ---- flattened per-tile stage program (shared by the TMA producer and the MMA issuer) ----------
  //   batch 1 of layer l: [(l,a) PE stage] (l,a) hidden K-steps 0-7      needs a_ready[a](l-1)
  //   batch 2 of layer l: (l,a) hidden 8-15, [(l,b) PE], (l,b) 0-7, 8-15  needs a_ready[b](l-1)
  // entry.x: bits 0-7 K-steps | F_* flags | bits 12-15 batch length (first entry of a batch only)
  // entry.y: TMEM column of the A operand (hidden stages) or first PE K-step (PE stages)
  enum : uint32_t { F_PE = 1u << 8, F_H = 1u << 9, F_FIRST = 1u << 10, F_LAST = 1u << 11, F_WAIT_A0 = 1u << 16,
                    F_WAIT_A1 = 1u << 17, F_INC0 = 1u << 18, F_INC1 = 1u << 19 };
  uint2* prog = reinterpret_cast<uint2*>(smem + SL.prog);
  int* prog_n = reinterpret_cast<int*>(smem + SL.prog + 8 * 127);
  if (threadIdx.x == 0) {
    int n_entries = 0;
    for (int l = 0; l < P.n_layers; ++l) {
      const int n_pe = P.L[l].n_pe_ks, n_h = P.L[l].n_h_ks, nh = P.L[l].n_halves;
      const uint32_t a_col = (l > 0 ? (X3 ? 256u : ((((l - 1) & 1) ? 384u : 256u))) : (DGRAD ? 384u : 256u));
      const int nst = stages_of(n_pe, n_h);
      int batch_first[2] = {-1, -1};
      uint32_t batch_len[2] = {0, 0}, batch_inc[2] = {0, 0};
      for (int h = 0; h < nh; ++h) {
        for (int st = 0; st < nst; ++st) {
          const StageInfo si = stage_info(n_pe, n_h, st);
          uint32_t w0 = (uint32_t)si.nks | (si.is_pe ? F_PE : 0u) | (h ? F_H : 0u) | (st == 0 ? F_FIRST : 0u) |
                        (st == nst - 1 ? F_LAST : 0u);
          const uint32_t w1 = si.is_pe ? (uint32_t)si.k0 : (a_col + 8u * si.k0);
          const int bsel = (h == 1 || (!si.is_pe && si.k0 + si.nks > 8)) ? 1 : 0;
          if (batch_first[bsel] < 0) { batch_first[bsel] = n_entries; w0 |= bsel ? F_WAIT_A1 : F_WAIT_A0; }
          ++batch_len[bsel];
          if (st == nst - 1) batch_inc[bsel] |= h ? F_INC1 : F_INC0;
          prog[n_entries++] = make_uint2(w0, w1);
        }
      }
      for (int bs = 0; bs < 2; ++bs)
        if (batch_first[bs] >= 0) prog[batch_first[bs]].x |= (batch_len[bs] << 12) | batch_inc[bs];
    }
    *prog_n = n_entries;
  }
  if (warp == WARP_TMA) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < P.const_floats; i += NUM_THREADS) consts[i] = A.tail[i];
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_entries = *prog_n;

  constexpr uint32_t COL_DA = 0, COL_A0 = 256, COL_A1 = 384;
  // activation buffer written by the epilogue of layer l (and read by layer l+1)
  auto a_out_col = [&](int l) -> uint32_t { return X3 ? COL_A0 : ((l & 1) ? COL_A1 : COL_A0); };

  if (warp == WARP_TMA) {
    // ===================== TMA producer: stream the packed weights through the ring ==============
    if (lane == 0) {
      uint32_t slot = 0, phase = 0, batch = 0;
      for (int64_t tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        const uint8_t* src = A.w;
        const bool copy = !(PLNERF_DBG(1) && tile != (int64_t)blockIdx.x);
        int i = 0;
        while (i < n_entries) {
          const int blen = (prog[i].x >> 12) & 15;
          if (!X3) {
            // one full-barrier per batch: armed with the batch's total bytes, every stage copy signals it.
            // 8 batch barriers > ring slots, so a barrier is never re-armed before its previous phase was consumed.
            uint32_t total = 0;
            for (int j = 0; j < blen; ++j) total += (prog[i + j].x & 255u) * KS_BYTES;
            const uint32_t bar = b_full(batch);
            if (copy) ptx::mbar_arrive_expect_tx(bar, total); else ptx::mbar_arrive(bar);
            for (int j = 0; j < blen; ++j) {
              const uint32_t bytes = (prog[i + j].x & 255u) * KS_BYTES;
              ptx::mbar_wait(w_empty(slot), phase ^ 1);
              if (copy) ptx::bulk_g2s_hint(s_ring + slot * (uint32_t)STAGE_BYTES, src, bytes, bar, ptx::l2_policy_evict_last());
              src += bytes;
              if (++slot == (uint32_t)A.n_stages) { slot = 0; phase ^= 1; }
            }
            ++batch;
          } else {
            for (int j = 0; j < blen; ++j) {
              const uint32_t bytes = (prog[i + j].x & 255u) * KS_BYTES;
              for (int rep = 0; rep < 2; ++rep) {
                ptx::mbar_wait(w_empty(slot), phase ^ 1);
                if (copy) { ptx::mbar_arrive_expect_tx(w_full(slot), bytes); ptx::bulk_g2s_hint(s_ring + slot * (uint32_t)STAGE_BYTES, src, bytes, w_full(slot), ptx::l2_policy_evict_last()); }
                else ptx::mbar_arrive(w_full(slot));
                src += bytes;
                if (++slot == (uint32_t)A.n_stages) { slot = 0; phase ^= 1; }
              }
            }
          }
          i += blen;
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer ===========================================================
    // The tensor pipe only buffers ~3-4 MMAs behind the issuing thread (measured), so everything the
    // thread does between two MMA batches must fit in ~200 cycles or the pipe idles.  The per-tile
    // control flow is therefore flattened once into a small "program" of stages in shared memory,
    // and stages that share their dependencies are issued as one batch:
    //   batch 1 of layer l: [(l,a) PE stage] (l,a) hidden K-steps 0-7     needs a_ready[a](l-1)
    //   batch 2 of layer l: (l,a) hidden 8-15, [(l,b) PE], (l,b) 0-7, 8-15 needs a_ready[b](l-1)
    // One elected lane issues; the whole warp follows the (warp-uniform) control flow.
    const uint32_t idesc = ptx::idesc_bf16_f32(128, 128);
    const uint64_t desc_base = ptx::smem_desc(0, 2048, 128);
    const uint32_t desc_hi = (uint32_t)(desc_base >> 32);
    const uint32_t desc_lo0 = (uint32_t)(desc_base & 0xFFFFFFFFu);  // LBO field, address bits zero
    auto mk_desc = [&](uint32_t lo) -> uint64_t { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
    auto lo_of = [&](uint32_t saddr) -> uint32_t { return desc_lo0 | ((saddr & 0x3FFFFu) >> 4); };
    constexpr uint32_t KS_DESC = KS_BYTES >> 4;   // descriptor address increment per K-step
    uint32_t slot = 0, phase = 0, batch = 0;
    uint32_t uses0 = 0, uses1 = 0, waited0 = 0, waited1 = 0;
    uint32_t tile_iter = 0;
    int tcnt = 0;
    // the weights of a batch always land long before its activations: their barrier is waited right after
    // the PREVIOUS batch was issued (while the tensor pipe is still busy), never on the critical path
    uint32_t bw0 = prog[0].x;
    uint4* prog2 = reinterpret_cast<uint4*>(smem + SL.prog2);
    for (int n = lane; n < n_entries; n += 32) {
      const uint2 e = prog[n];
      const uint32_t fl = ((e.x & F_FIRST) ? 1u : 0u) | ((e.x & F_PE) ? 2u : 0u) | ((e.x & F_LAST) ? 4u : 0u) | ((e.x & 255u) << 8) |
                          ((e.x & F_WAIT_A0) ? 8u : 0u) | ((e.x & F_WAIT_A1) ? 16u : 0u) | (((e.x >> 12) & 15u) ? 32u : 0u) |
                          ((e.x & F_INC0) ? 64u : 0u) | ((e.x & F_INC1) ? 128u : 0u);
      prog2[n] = make_uint4(tmem + COL_DA + ((e.x & F_H) ? 128u : 0u), (e.x & F_PE) ? lo_of(s_pe_hi + e.y * KS_BYTES) : tmem + e.y, fl,
                            d_full0 + ((e.x & F_H) ? 8u : 0u));
    }
    __syncwarp();
    const uint64_t ring_desc = mk_desc(lo_of(s_ring));
    if (!X3) {
      // production path: one asm call per tile (issue_tile); b_full(0) was waited above, the program re-waits it harmlessly
      // only through the batch counter, so start the counter's bookkeeping consistently: batch 0's wait happens in the asm
      IssueState st = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      for (int64_t tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tile_iter) {
        ptx::mbar_wait(pe_ready, tile_iter & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one())
          issue_tile(st, sbase + SL.prog2, (uint32_t)n_entries, idesc, ring_desc, desc_hi, w_empty(0), (uint32_t)A.n_stages, b_full(0), a_ready0);
        __syncwarp();
        // the state lives in the elected lane; elect.sync of a converged warp picks the same lane every time
      }
    } else if (!PLNERF_DBG(16)) {
      // hi+lo split mode: same one-call-per-tile structure (issue_tile_x3); debug flag 16 selects the C++ loop below
      IssueStateX3 st = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      for (int64_t tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tile_iter) {
        ptx::mbar_wait(pe_ready, tile_iter & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one())
          issue_tile_x3(st, sbase + SL.prog2, (uint32_t)n_entries, idesc, ring_desc, desc_hi, w_full(0), (uint32_t)A.n_stages, a_ready0);
        __syncwarp();
      }
    } else
    for (int64_t tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tile_iter) {
      const bool trace_on = (tile_iter == 2) && lane == 0;
      ptx::mbar_wait(pe_ready, tile_iter & 1);
      int i = 0;
      while (i < n_entries) {
        PLNERF_TRACE(0, tcnt, 1000 + i);                 // batch loop top
        const int blen = (bw0 >> 12) & 15;
        if (bw0 & F_WAIT_A0) { while (waited0 < uses0) { ptx::mbar_wait(a_ready0, waited0 & 1); ++waited0; } }
        if (bw0 & F_WAIT_A1) { while (waited1 < uses1) { ptx::mbar_wait(a_ready0 + 8u, waited1 & 1); ++waited1; } }
        ptx::tc_fence_after();
        PLNERF_TRACE(0, tcnt, 2000 + i);                 // dependencies + weights ready, issuing batch
        if (ptx::elect_one()) {
          uint32_t sl = slot, ph = phase;   // private ring cursor of the issuing lane
          for (int j = 0; j < blen; ++j) {
            const uint2 e = prog[i + j];
            const int nks = e.x & 255;
            const uint32_t d = tmem + COL_DA + ((e.x & F_H) ? 128u : 0u);
            uint32_t acc = (e.x & F_FIRST) ? 0u : 1u;
#pragma unroll
            for (int rep = 0; rep < nsplit; ++rep) {
              if (X3) { ptx::mbar_wait(w_full(sl), ph); ptx::tc_fence_after(); }
              const uint32_t b_lo = lo_of(s_ring + sl * (uint32_t)STAGE_BYTES);
              if (e.x & F_PE) {
                const uint32_t a_lo_hi = lo_of(s_pe_hi + e.y * KS_BYTES), a_lo_lo = lo_of(s_pe_lo + e.y * KS_BYTES);
                for (int k = 0; k < nks; ++k) {
                  const uint64_t bd = mk_desc(b_lo + k * KS_DESC);
                  if (rep == 0) {
                    ptx::mma_ss(d, mk_desc(a_lo_hi + k * KS_DESC), bd, idesc, (k > 0) ? 1u : acc);
                    if (X3) ptx::mma_ss(d, mk_desc(a_lo_lo + k * KS_DESC), bd, idesc, 1);
                  } else {
                    ptx::mma_ss(d, mk_desc(a_lo_hi + k * KS_DESC), bd, idesc, 1);
                  }
                }
              } else {
                const uint32_t at = tmem + e.y, at_lo = at + (COL_A1 - COL_A0);
                if (nks == KS_PER_STAGE) {
                  if (rep == 0) issue_ts8<X3>(d, at, at_lo, mk_desc(b_lo), idesc, acc);
                  else issue_ts8<false>(d, at, at_lo, mk_desc(b_lo), idesc, 1);
                } else {
                  for (int k = 0; k < nks; ++k) {
                    const uint64_t bd = mk_desc(b_lo + k * KS_DESC);
                    if (rep == 0) {
                      ptx::mma_ts(d, at + 8u * k, bd, idesc, (k > 0) ? 1u : acc);
                      if (X3) ptx::mma_ts(d, at_lo + 8u * k, bd, idesc, 1);
                    } else {
                      ptx::mma_ts(d, at + 8u * k, bd, idesc, 1);
                    }
                  }
                }
              }
              ptx::mma_commit(w_empty(sl));
              acc = 1;
              if (++sl == (uint32_t)A.n_stages) { sl = 0; ph ^= 1; }
            }
            if (e.x & F_LAST) ptx::mma_commit(d_full0 + ((e.x & F_H) ? 8u : 0u));
          }
        }
        __syncwarp();
        PLNERF_TRACE(0, tcnt, 5000 + i);                 // issue returned
        // every lane advances the (warp-uniform) ring cursor and use counters past this batch
        slot += (uint32_t)(blen * nsplit);
        while (slot >= (uint32_t)A.n_stages) { slot -= (uint32_t)A.n_stages; phase ^= 1; }
        uses0 += (bw0 >> 18) & 1u;
        uses1 += (bw0 >> 19) & 1u;
        i += blen;
        ++batch;
        {
          const bool more = (i < n_entries) || (tile + (int64_t)gridDim.x < A.n_tiles);
          if (more) {
            bw0 = prog[(i < n_entries) ? i : 0].x;
            if (!X3) ptx::mbar_wait(b_full(batch), (batch >> 3) & 1u);
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps (8): PE prologue, bias/ReLU/pack, heads, output ========
    // lane quarter q = warp % 4 (TMEM lanes 32q..32q+31, one row per thread), column group grp = (warp-2)/4
    // owns CHUNKS_PER_GRP 32-column chunks of every 128-column half.
    const int q = warp & 3, grp = warp >> 2;   // TMEM lane quarter = hardware warp id % 4; grp in [0, NGRP)
    const int row = q * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    uint32_t seen0 = 0, seen1 = 0;
    int l_pe_last = 0;
    for (int l = 0; l < P.n_layers; ++l) if (P.L[l].n_pe_ks > 0) l_pe_last = l;
    float pe_pos[3] = {0.f, 0.f, 0.f}, pre_od[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, pre_z = 0.f;

    if (!DGRAD && (int64_t)blockIdx.x < A.n_tiles) {
      pe_prologue<MODE>(A, smem, SL, blockIdx.x, row, grp);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(pe_ready);
    }
    int tcnt = 0;
    uint32_t tile_iter = 0;
    for (int64_t tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tile_iter) {
      const bool trace_on = (tile_iter == 2) && lane == 0 && q == 0;
      const int64_t g = tile * TILE_M + row;
      const bool valid = g < A.M;
      const int64_t gc = valid ? g : (A.M - 1);
      float alpha_acc = 0.f;
      float head[MAX_OUT_CH];
#pragma unroll
      for (int c = 0; c < MAX_OUT_CH; ++c) head[c] = 0.f;
      const float* vbrow = A.viewbias ? (A.viewbias + (gc / A.vb_div) * 128) : nullptr;
      // The views layer's epilogue is the tile's last link; with ~3 KB of L1 left its per-ray bias rows would come from L2
      // (~600 cycles on the critical path).  They are copied to shared memory asynchronously at the start of the tile.
      const bool vb_smem = !DGRAD && A.viewbias && A.vb_div >= 64;
      const int64_t vb_ray0 = (tile * TILE_M) / A.vb_div;
      if (vb_smem && threadIdx.x < VB_SMEM_RAYS * 128) {
        const int64_t last_row = (tile * TILE_M + TILE_M - 1 < A.M) ? tile * TILE_M + TILE_M - 1 : A.M - 1;
        int64_t ray = vb_ray0 + (threadIdx.x >> 7);
        if (ray > last_row / A.vb_div) ray = last_row / A.vb_div;
        const uint32_t dst = sbase + SL.vb + 4u * threadIdx.x;
        const float* srcp = A.viewbias + ray * 128 + (threadIdx.x & 127);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(srcp) : "memory");
      }
      if (vb_smem) asm volatile("cp.async.commit_group;" ::: "memory");
      // this row's staged bias row (hoisted: the 64-bit division must not sit in the views epilogue's inner loop)
      const float* vb_srow = reinterpret_cast<const float*>(smem + SL.vb) + (vb_smem ? (int)((gc / A.vb_div) - vb_ray0) * 128 : 0);
      float g_alpha = 0.f;
      if (DGRAD) {
        // the chain's input buffer (A1) is free again: the previous tile's last layer has been drained
        g_alpha = valid ? A.g_raw[g * (int64_t)A.g_stride + 3] : 0.f;
        dgrad_prologue(A, consts, tmem + lane_addr + COL_A1, tile, row, grp, g_alpha);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(pe_ready);
      }
      uint8_t* in_tile = (STASH) ? A.in_stash + tile * (int64_t)A.tl.in_tile_bytes : nullptr;
      uint8_t* dy_tile = (DGRAD) ? A.dy_stash + tile * (int64_t)A.tl.dy_tile_bytes : nullptr;
      uint32_t* mask_tile = (STASH || DGRAD) ? A.masks + tile * (int64_t)A.tl.mask_tile_words : nullptr;

      if (DGRAD) {
        // the NEXT tile's ReLU masks (written by the forward two kernels ago: a DRAM read) -> L2 a whole tile ahead, so
        // the half-layer-ahead register loads below find them there
        auto prefetch_masks = [&](int64_t t2) {
          if (t2 >= A.n_tiles) return;
          const char* mb = reinterpret_cast<const char*>(A.masks + t2 * (int64_t)A.tl.mask_tile_words);
          const int bytes = A.tl.mask_tile_words * 4;
          for (int o = (int)threadIdx.x * 128; o < bytes; o += NUM_EPI_THREADS * 128)
            asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(mb + o));
        };
        if (tile_iter == 0) prefetch_masks(tile);
        prefetch_masks(tile + gridDim.x);
      }
      uint32_t m_next[CHUNKS_PER_GRP] = {};
      auto load_masks = [&](int l2, int h2, uint32_t (&m)[CHUNKS_PER_GRP]) {
        const int mo = ltab[l2].w;
        if (mo < 0) return;
#pragma unroll
        for (int cc = 0; cc < CHUNKS_PER_GRP; ++cc)
          m[cc] = __ldg(mask_tile + mo + ((h2 * 128 + (CHUNKS_PER_GRP * grp + cc) * 32) >> 5) * 128 + row);
      };
      if (DGRAD) load_masks(0, 0, m_next);
      for (int l = 0; l < P.n_layers; ++l) {
        const int4 lt = ltab[l];
        const int epi = lt.x & 255, flags = (lt.x >> 16) & 255, n_halves = (lt.x >> 8) & 255, bias_off = lt.y;
        const int st_off = lt.z, m_off = lt.w;
        const uint32_t a_out = tmem + lane_addr + a_out_col(l);
        const uint32_t a_out_lo = tmem + lane_addr + COL_A1;
        if (vb_smem && epi == EPI_VIEWS) {
          asm volatile("cp.async.wait_all;" ::: "memory");
          asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_THREADS) : "memory");     // rows were fetched by other threads
        }
        for (int h = 0; h < n_halves; ++h) {
          // gradient chain: this thread's ReLU-mask word of the half (one per 32 columns) was requested one half-layer AHEAD
          // (a global load on first use was the top stall of the kernel: 17% of its samples); request the next one now
          uint32_t m_pre[CHUNKS_PER_GRP];
          if (DGRAD) {
#pragma unroll
            for (int cc = 0; cc < CHUNKS_PER_GRP; ++cc) m_pre[cc] = m_next[cc];
            const int ln = (h + 1 < n_halves) ? l : l + 1, hn = (h + 1 < n_halves) ? h + 1 : 0;
            if (ln < P.n_layers) load_masks(ln, hn, m_next);
          }
          if (h == 0) {
            ptx::mbar_wait(d_full0, seen0 & 1); ++seen0;
            if (X3 && n_halves == 2) {
              // bf16x3 keeps ONE activation buffer: half b's MMAs still read it, wait for them too
              ptx::mbar_wait(d_full0 + 8u, seen1 & 1); ++seen1;
            }
          } else if (!X3) {
            ptx::mbar_wait(d_full0 + 8u, seen1 & 1); ++seen1;
          }
          ptx::tc_fence_after();
          PLNERF_TRACE(1 + grp, tcnt, 3000 + l * 10 + h);     // accumulator half observed full
          // all 32-column chunks of this warp are requested before the single wait::ld
          uint32_t r2[CHUNKS_PER_GRP][32];
          if (!PLNERF_DBG(2)) {
#pragma unroll
            for (int cc = 0; cc < CHUNKS_PER_GRP; ++cc)
              ptx::tmem_ld32(tmem + lane_addr + COL_DA + 128u * h + 32u * (CHUNKS_PER_GRP * grp + cc), r2[cc]);
            ptx::tmem_ld_wait();
          }
          // a layer that writes no activations (views / last trunk layer of a no-viewdirs net) hands its accumulator half
          // back as soon as the values are in registers: its head math is then off the next tile's critical path
          const bool early_arrive = !DGRAD && (epi == EPI_VIEWS || epi == EPI_RELU_HEAD);
          if (early_arrive) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(a_ready0 + 8u * h);
          }
#pragma unroll
          for (int cc = 0; cc < CHUNKS_PER_GRP; ++cc) {
            if (PLNERF_DBG(2)) break;
            const int c = CHUNKS_PER_GRP * grp + cc;
            const int n0 = h * 128 + c * 32;
            float* val = reinterpret_cast<float*>(r2[cc]);
            if (DGRAD) {
              // gradient chain: (+ g_alpha * w_alpha) then relu'(.) of the forward activation, no bias
              if (flags & FLAG_ALPHA) {
                const float* aw = consts + P.alpha_w_off + n0;
#pragma unroll
                for (int i = 0; i < 32; ++i) val[i] = fmaf(g_alpha, aw[i], val[i]);
              }
              uint32_t pk[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = ptx::pack_bf16(val[2 * i], val[2 * i + 1]);
              if (m_off >= 0) {
                const uint32_t m = m_pre[cc];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] &= relu_mask_word(m, i);
              }
              ptx::tmem_st16(a_out + (uint32_t)(n0 >> 1), pk);
              uint8_t* t = dy_tile + st_off;
              if (!PLNERF_DBG(32))                 // (measurement: the gradient chain without its dY stores)
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                stash_store8<false>(t, 256, row, (n0 >> 3) + q4, make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]));
            } else if (epi == EPI_VIEWS) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>((vb_smem ? vb_srow : vbrow) + n0 + i);
                val[i] = fmaxf(val[i] + b4.x, 0.f); val[i + 1] = fmaxf(val[i + 1] + b4.y, 0.f);
                val[i + 2] = fmaxf(val[i + 2] + b4.z, 0.f); val[i + 3] = fmaxf(val[i + 3] + b4.w, 0.f);
              }
              if (STASH) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = ptx::pack_bf16(val[2 * i], val[2 * i + 1]);
                mask_tile[m_off + (n0 >> 5) * 128 + row] = relu_mask_from_packed(pk);
                uint8_t* t = in_tile + st_off;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                  stash_store8(t, 128, row, (n0 >> 3) + q4, make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]));
              }
              const float* rw = consts + P.rgb_w_off + n0;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
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
              {
                const float* bias = consts + bias_off + n0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(bias + i);
                  add2(val[i], val[i + 1], b4.x, b4.y);
                  add2(val[i + 2], val[i + 3], b4.z, b4.w);
                }
              }
              uint32_t pk[16];
              if (!X3 && !STASH && epi == EPI_RELU_A) {
                // hot path: ReLU fused into the bf16x2 conversion (a head flagged on this layer re-applies the ReLU to
                // the fp32 values after the slot has been handed back)
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = pack_bf16_relu(val[2 * i], val[2 * i + 1]);
                ptx::tmem_st16(a_out + (uint32_t)(n0 >> 1), pk);
              } else if (!X3 && !STASH && epi == EPI_LINEAR_A && flags == 0) {
                // linear layer (feature_linear): plain conversion
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = ptx::pack_bf16(val[2 * i], val[2 * i + 1]);
                ptx::tmem_st16(a_out + (uint32_t)(n0 >> 1), pk);
              } else {
                if (epi != EPI_LINEAR_A) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) val[i] = fmaxf(val[i], 0.f);
                }
                // (the alpha / output heads read these relu'd fp32 values AFTER the slot has been handed back, below)
                if (epi != EPI_RELU_HEAD) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) pk[i] = ptx::pack_bf16(val[2 * i], val[2 * i + 1]);
                  ptx::tmem_st16(a_out + (uint32_t)(n0 >> 1), pk);
                  if (STASH) {
                    if (m_off >= 0) mask_tile[m_off + (n0 >> 5) * 128 + row] = relu_mask_from_packed(pk);
                    uint8_t* t = in_tile + st_off;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                      stash_store8(t, 256, row, (n0 >> 3) + q4, make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]));
                  }
                  if (X3) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                      pk[i] = ptx::pack_bf16(val[2 * i] - ptx::bf16_round(val[2 * i]),
                                             val[2 * i + 1] - ptx::bf16_round(val[2 * i + 1]));
                    ptx::tmem_st16(a_out_lo + (uint32_t)(n0 >> 1), pk);
                  }
                } else if (STASH) {
                  // last trunk layer of a network without view directions: no next layer reads it, but the weight
                  // gradients of output_linear and the gradient chain's first mask do
#pragma unroll
                  for (int i = 0; i < 16; ++i) pk[i] = ptx::pack_bf16(val[2 * i], val[2 * i + 1]);
                  if (m_off >= 0) mask_tile[m_off + (n0 >> 5) * 128 + row] = relu_mask_from_packed(pk);
                  uint8_t* t = in_tile + st_off;
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4)
                    stash_store8(t, 256, row, (n0 >> 3) + q4, make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]));
                }
              }
            }
          }
          if (!early_arrive) {
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(a_ready0 + 8u * h);
          }
          PLNERF_TRACE(1 + grp, tcnt, 4000 + l * 10 + h);     // activations written, arrived
          // heads with tiny N, off the layer-to-layer critical chain: the fp32 values are still in registers (both heads
          // sit behind a ReLU layer; the hot path above did not apply it in place, so it is (re-)applied here)
          if (!DGRAD && !PLNERF_DBG(2) && epi != EPI_VIEWS && (flags & (FLAG_ALPHA | FLAG_OUTHEAD))) {
            static_assert(CHUNKS_PER_GRP == 1, "deferred heads assume one chunk per warp and half");
            const float* val = reinterpret_cast<const float*>(r2[0]);
            const int n0 = h * 128 + grp * 32;
            if (flags & FLAG_ALPHA) {
              const float* aw = consts + P.alpha_w_off + n0;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(aw + i);
                alpha_acc = fmaf(fmaxf(val[i], 0.f), w4.x, alpha_acc); alpha_acc = fmaf(fmaxf(val[i + 1], 0.f), w4.y, alpha_acc);
                alpha_acc = fmaf(fmaxf(val[i + 2], 0.f), w4.z, alpha_acc); alpha_acc = fmaf(fmaxf(val[i + 3], 0.f), w4.w, alpha_acc);
              }
            }
            if (flags & FLAG_OUTHEAD) {
#pragma unroll
              for (int ch = 0; ch < MAX_OUT_CH; ++ch) {
                if (ch < P.out_ch) {
                  const float* ow = consts + P.out_w_off + ch * 256 + n0;
#pragma unroll
                  for (int i = 0; i < 32; ++i) head[ch] = fmaf(fmaxf(val[i], 0.f), ow[i], head[ch]);
                }
              }
            }
          }
        }
        // The next tile's depths / ray rows are a cold DRAM read (~2000+ cycles): the loads are issued after layer 0's
        // epilogue, consumed (o + d*z) after layer 1's, and only the cheap encoding itself is left for l_pe_last --
        // otherwise that latency sits between two layers' epilogues and idles the tensor pipe.
        if (!DGRAD && !X3 && !A.x_emb && l_pe_last >= 2) {
          const int64_t nt = tile + gridDim.x;
          if (nt < A.n_tiles) {
            if (l == 0) {
              const int64_t gn = nt * TILE_M + row, gcn = (gn < A.M) ? gn : (A.M - 1);
              const float* rp = A.rays + (gcn / A.S) * (int64_t)A.stride;
              pre_z = A.z[gcn];
#pragma unroll
              for (int c = 0; c < 6; ++c) pre_od[c] = rp[c];
            } else if (l == 1) {
#pragma unroll
              for (int c = 0; c < 3; ++c) pe_pos[c] = __fadd_rn(pre_od[c], __fmul_rn(pre_od[3 + c], pre_z));  // o + d*z, two roundings
            }
          }
        }
        if (!DGRAD && l == l_pe_last) {
          // every MMA that reads the PE tile of this tile has completed (its d_full was waited):
          // encode the NEXT tile's rows now, overlapped with the remaining layers
          const int64_t nt = tile + gridDim.x;
          if (nt < A.n_tiles) {
            PLNERF_TRACE(1 + grp, tcnt, 7000);
            pe_prologue<MODE>(A, smem, SL, nt, row, grp, (!X3 && !A.x_emb && l_pe_last >= 2) ? pe_pos : nullptr);
            PLNERF_TRACE(1 + grp, tcnt, 7001);
            ptx::fence_proxy_async_smem();
            PLNERF_TRACE(1 + grp, tcnt, 7002);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(pe_ready);
            PLNERF_TRACE(1 + grp, tcnt, 7003);
          }
        }
      }
      if (DGRAD) continue;   // the gradient chain's outputs are the dY stash tiles
      // ---- combine the column groups' partial head sums and write the row
      if (grp > 0) {
        float* x = xch + ((grp - 1) * TILE_M + row) * (MAX_OUT_CH + 1);
        x[0] = alpha_acc;
#pragma unroll
        for (int c = 0; c < MAX_OUT_CH; ++c) x[1 + c] = head[c];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_THREADS) : "memory");
      if (grp == 0) {
#pragma unroll
        for (int g2 = 1; g2 < NGRP; ++g2) {
          const float* x = xch + ((g2 - 1) * TILE_M + row) * (MAX_OUT_CH + 1);
          alpha_acc += x[0];
#pragma unroll
          for (int c = 0; c < MAX_OUT_CH; ++c) head[c] += x[1 + c];
        }
        if (valid) {
          float* o = A.out + g * (int64_t)A.out_stride;
          if (P.use_viewdirs) {
            o[0] = head[0] + consts[P.rgb_b_off + 0];
            o[1] = head[1] + consts[P.rgb_b_off + 1];
            o[2] = head[2] + consts[P.rgb_b_off + 2];
            o[3] = alpha_acc + consts[P.alpha_b_off];
          } else {
#pragma unroll
            for (int ch = 0; ch < MAX_OUT_CH; ++ch)
              if (ch < P.out_ch) o[ch] = head[ch] + consts[P.out_b_off + ch];
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_THREADS) : "memory");   // xch is reused by the next tile
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == WARP_TMA) ptx::tmem_dealloc(tmem, 512);
}

#include "mlp_fwd3.cuh"

// =============================================================================================
// weight gradients:  dW[n, k] += sum_m dY[m, n] * In[m, k]   (contraction over the sample rows m)
// Both operands are the MN-major stash tiles written by the forward (In) and by the gradient chain
// (dY), loaded with plain bulk TMA; accumulators stay in tensor memory across all tiles of a CTA and
// are flushed once with fp32 atomics.  The bias gradient is the same GEMM against a ones operand.
// A work item = (dY tensor, 128-row half, In tensor); grid = items x splits (tiles strided by split).
// Pipeline: four stages of HALF a tile each (the 64 rows of one m-half: 16 KB of dY columns + up to 32 KB of activations,
// the two m-halves of an MN-major tile being separate contiguous regions), four MMAs per stage.  Against two whole-tile
// stages (the same 192 KB) the finer grain keeps the DRAM queue fuller: same-box A/B -2.6% on the training step.
// =============================================================================================
constexpr int MAX_WG_ITEMS = 40;
struct WgradItem {
  // M side (128 accumulator rows) = columns [128 half, +128) of the dY tensor dy_idx; N side = the In tensor in_idx.
  // swapped (the alpha / rgb heads): M side = that column half of the In tensor, N side = the dY tensor (dy_head, 16 wide).
  int32_t dy_idx, half, in_idx, swapped;
  int32_t row_stride, col_stride, c_first, ncols;
  float* dW;   // accumulator (r, c), c in [c_first, c_first + ncols) -> dW[(128 half + r) * row_stride + (c - c_first) * col_stride]
  float* db;   // db[128 half + r] (row sums of the M side against ones) or null
  int32_t cta0, splits;   // this item's CTAs: [cta0, cta0 + splits), each taking every splits-th tile
};
struct WgradArgs {
  WgradItem items[MAX_WG_ITEMS];
  int32_t n_items;
  TrainLayout tl;
  const uint8_t* in_stash;
  const uint8_t* dy_stash;
  int64_t n_tiles;
};
constexpr int WG_THREADS = 192;
constexpr int WG_STAGES = 4, WG_STAGE_BYTES = 16384 + 32768;   // a stage = 64 rows (one m-half): dY column half + activation tile half

__global__ void __launch_bounds__(WG_THREADS, 1) k_wgrad(const __grid_constant__ WgradArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  int item = 0;
  while (item + 1 < A.n_items && (int)blockIdx.x >= A.items[item + 1].cta0) ++item;
  const WgradItem& it = A.items[item];
  const int split = (int)blockIdx.x - it.cta0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint8_t* m_base = it.swapped ? A.in_stash + A.tl.in_off[it.in_idx] : A.dy_stash + A.tl.dy_off[it.dy_idx];
  const uint8_t* n_base = it.swapped ? A.dy_stash + A.tl.dy_off[it.dy_idx] : A.in_stash + A.tl.in_off[it.in_idx];
  const int64_t m_tile_bytes = it.swapped ? A.tl.in_tile_bytes : A.tl.dy_tile_bytes;
  const int64_t n_tile_bytes = it.swapped ? A.tl.dy_tile_bytes : A.tl.in_tile_bytes;
  const int Wd = it.swapped ? A.tl.dy_width[it.dy_idx] : A.tl.in_width[it.in_idx];     // N extent
  const int Wdy = it.swapped ? A.tl.in_width[it.in_idx] : A.tl.dy_width[it.dy_idx];    // width of the M-side tensor
  uint8_t* s_ones = smem + WG_STAGES * WG_STAGE_BYTES;
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t s_bars = sbase + WG_STAGES * WG_STAGE_BYTES + 1024;
  auto full = [&](int s) { return s_bars + 8u * s; };
  auto empty = [&](int s) { return s_bars + 8u * WG_STAGES + 8u * s; };
  const uint32_t done = s_bars + 16u * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + WG_STAGES * WG_STAGE_BYTES + 1024 + 16 * WG_STAGES + 16);
  if (threadIdx.x == 0) {
    for (int s2 = 0; s2 < WG_STAGES; ++s2) { ptx::mbar_init(full(s2), 1); ptx::mbar_init(empty(s2), 1); }
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < 256; i += WG_THREADS) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3F803F80u;  // bf16 1.0 pairs
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  int64_t my_tiles = 0;
  for (int64_t t = split; t < A.n_tiles; t += it.splits) ++my_tiles;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      const uint32_t in_half = (uint32_t)Wd * 128u;       // one m-half (64 rows) of the N-side tile
      for (int64_t t = split; t < A.n_tiles; t += it.splits) {
        const uint8_t* dy = m_base + t * m_tile_bytes;
        for (int mh = 0; mh < 2; ++mh) {
          ptx::mbar_wait(empty(st), ph ^ 1);
          ptx::mbar_arrive_expect_tx(full(st), 16384u + in_half);
          const uint32_t sdst = sbase + st * WG_STAGE_BYTES;
          ptx::bulk_g2s(sdst, dy + ((size_t)(mh * (Wdy >> 3) + 16 * it.half)) * 1024, 16384u, full(st));
          ptx::bulk_g2s(sdst + 16384, n_base + t * n_tile_bytes + (size_t)mh * in_half, in_half, full(st));
          if (++st == WG_STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::idesc_bf16_f32_mn(128, Wd);
    const uint32_t idesc_b = ptx::idesc_bf16_f32_mn(128, 16);
    const uint64_t ones_desc = ptx::smem_desc(ptx::smem_u32(s_ones), 128, 256);
    uint32_t st = 0, ph = 0, acc = 0;
    for (int64_t t = split; t < A.n_tiles; t += it.splits) {
      for (int mh = 0; mh < 2; ++mh) {
        ptx::mbar_wait(full(st), ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = sbase + st * WG_STAGE_BYTES, sb = sa + 16384;
          uint32_t a2 = acc;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t ad = ptx::smem_desc(sa + j * 256, 128, 1024);
            const uint64_t bd = ptx::smem_desc(sb + j * 256, 128, 1024);
            ptx::mma_ss(tmem, ad, bd, idesc, a2);
            if (it.db) ptx::mma_ss(tmem + 256u, ad, ones_desc, idesc_b, a2);
            a2 = 1;
          }
          ptx::mma_commit(empty(st));
        }
        __syncwarp();
        acc = 1;
        if (++st == WG_STAGES) { st = 0; ph ^= 1; }
      }
    }
    if (ptx::elect_one()) ptx::mma_commit(done);
    __syncwarp();
  } else {
    // flush warps 2..5: TMEM lane quarter = warp % 4
    if (my_tiles > 0) {
      ptx::mbar_wait(done, 0);
      ptx::tc_fence_after();
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
      float* drow = it.dW + (int64_t)(128 * it.half + r) * it.row_stride;
      for (int c0 = 0; c0 < Wd; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem + lane_addr + (uint32_t)c0, v);     // (columns past Wd: allocated, never written, filtered below)
        ptx::tmem_ld_wait();
        if (it.col_stride == 1) {
          // four columns per reduction (red.global.add.v4.f32) wherever the row's address is 16-byte aligned: a CTA's flush
          // is 32 768 fp32 reductions during which it reads nothing
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const int c = c0 + i - it.c_first;
            float* p = drow + c;
            if (c >= 0 && c + 3 < it.ncols && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                           "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3])) : "memory");
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (c + k >= 0 && c + k < it.ncols) atomicAdd(p + k, __uint_as_float(v[i + k]));
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = c0 + i - it.c_first;
            if (c >= 0 && c < it.ncols) atomicAdd(drow + (int64_t)c * it.col_stride, __uint_as_float(v[i]));
          }
        }
      }
      if (it.db) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem + lane_addr + 256u, v);
        ptx::tmem_ld_wait();
        atomicAdd(it.db + 128 * it.half + r, __uint_as_float(v[0]));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem, 512);
}

// Per-device facts and one-time kernel attributes (cudaFuncSetAttribute is per device): indexed by the current device,
// so one process may drive several GPUs.
constexpr int MAX_DEVICES = 64;
struct DeviceInfo { int sms, max_smem; bool ok, attrs_v1, attrs_v3, attrs_wgrad; };
DeviceInfo g_dev[MAX_DEVICES] = {};
std::mutex g_dev_mu;
int g_num_sms = 0, g_max_smem = 0;   // of the current device; refreshed by query_device() at every entry
DeviceInfo* g_cur = nullptr;
int query_device() {
  int dev = 0;
  PLNERF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEVICES) { set_error("device ordinal %d out of range", dev); return PLNERF_E_UNSUPPORTED; }
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DeviceInfo& D = g_dev[dev];
  if (!D.ok) {
    int sms = 0, smem = 0, major = 0;
    PLNERF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PLNERF_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    PLNERF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) { set_error("plnerf_b200 needs an sm_100a (B200) device, found compute capability %d.x", major); return PLNERF_E_UNSUPPORTED; }
    D.sms = sms; D.max_smem = smem; D.ok = true;
  }
  g_num_sms = D.sms; g_max_smem = D.max_smem; g_cur = &D;
  return PLNERF_OK;
}

long long* g_trace = nullptr;
struct ProfRec { cudaEvent_t e0, e1; int64_t rows; };
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_evt_pool;
bool g_prof_on = false;
std::mutex g_prof_mu;

cudaEvent_t get_event() {
  if (!g_evt_pool.empty()) { cudaEvent_t e = g_evt_pool.back(); g_evt_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}

// developer library only: integer knob from the environment (0 when unset); the product build has no knobs
inline int dbg_env(const char* name) {
#ifdef PLNERF_DEBUG
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
#else
  (void)name;
  return 0;
#endif
}

// k_mlp3 (two tiles in flight, TMEM-resident activations, shared weight stages): the bf16 inference kernel.
// Returns 1 when the plan does not fit it (very deep networks / not enough shared memory) so the caller falls back to
// k_mlp_fwd -- both are sm_100a tcgen05 kernels on the same packed weights.
int launch_mlp3(MlpArgs& a, cudaStream_t st, bool stash = false) {
  if (stash && !a.plan.use_viewdirs) return 1;
  int n_stage_tile = 0;
  for (int l = 0; l < a.plan.n_layers; ++l) n_stage_tile += a.plan.L[l].n_halves * stages_of(a.plan.L[l].n_pe_ks, a.plan.L[l].n_h_ks);
  if (n_stage_tile > 48) return 1;
  int n_slots = v3::MAX_SLOTS;
  auto total_of = [&](int ns) -> int {
    return a.plan.use_viewdirs ? (int)v3::smem3_layout<true>(ns, a.plan.const_floats).total : (int)v3::smem3_layout<false>(ns, a.plan.const_floats).total;
  };
  while (n_slots > 3 && total_of(n_slots) > g_max_smem) --n_slots;
  { const int force = dbg_env("PLNERF_STAGES"); if (force >= 3 && force < n_slots) n_slots = force; }
  const int smem_total = total_of(n_slots);
  if (smem_total > g_max_smem) return 1;
  if (!g_cur->attrs_v3) {
    PLNERF_CUDA(cudaFuncSetAttribute(v3::k_mlp3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    PLNERF_CUDA(cudaFuncSetAttribute(v3::k_mlp3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    PLNERF_CUDA((cudaFuncSetAttribute(v3::k_mlp3<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem)));
    g_cur->attrs_v3 = true;
  }
  a.n_stages = n_slots;
  a.trace = g_trace;
  a.debug_flags = dbg_env("PLNERF_DEBUG_FLAGS");
  a.n_tiles = ceil_div(a.M, TILE_M);
  // contiguous ranges of whole units (rays) per CTA; a CTA should have at least one full pair of tiles of work
  a.unit_rows = a.fuse_comp ? a.vb_div : 2 * TILE_M;
  const int64_t n_units = ceil_div(a.M, a.unit_rows);
  int64_t cta_units = ceil_div(n_units, g_num_sms);
  const int64_t min_units = ceil_div(2 * TILE_M, a.unit_rows);
  if (cta_units < min_units) cta_units = min_units;
  a.cta_units = cta_units;
  const unsigned grid = (unsigned)ceil_div(n_units, cta_units);
  ProfRec rec{nullptr, nullptr, a.M};
  if (g_prof_on) { rec.e0 = get_event(); rec.e1 = get_event(); cudaEventRecord(rec.e0, st); }
  if (stash) v3::k_mlp3<true, true><<<grid, v3::THREADS3, smem_total, st>>>(a);
  else if (a.plan.use_viewdirs) v3::k_mlp3<true><<<grid, v3::THREADS3, smem_total, st>>>(a);
  else v3::k_mlp3<false><<<grid, v3::THREADS3, smem_total, st>>>(a);
  if (g_prof_on) { cudaEventRecord(rec.e1, st); std::lock_guard<std::mutex> lk(g_prof_mu); g_prof.push_back(rec); }
  PLNERF_LAUNCH_CHECK("k_mlp3");
  return PLNERF_OK;
}

// mode: -1 = inference in the plan's precision (bf16 -> k_mlp3, bf16x3 -> k_mlp_fwd<1>), 2 = forward + training stash,
// 3 = input-gradient chain
int launch_mlp(MlpArgs& a, cudaStream_t st, int mode = -1) {
  int rc = query_device();
  if (rc) return rc;
  if (mode < 0) mode = (a.plan.precision == PLNERF_PREC_BF16X3) ? 1 : 0;
  if ((mode == 0 || mode == 2) && a.M < ((int64_t)1 << 40) && !dbg_env("PLNERF_MLP_V1")) {
    rc = launch_mlp3(a, st, mode == 2);
    if (rc <= 0) return rc;     // > 0: the plan does not fit k_mlp3, fall through to the single-tile kernel
  }
  a.fuse_comp = 0;              // only k_mlp3 composites in-kernel
  int n_stages = MAX_STAGES;
  while (n_stages > 2 && (int)smem_layout(n_stages).total > g_max_smem) --n_stages;
  { const int force = dbg_env("PLNERF_STAGES"); if (force >= 2 && force < n_stages) n_stages = force; }
  a.n_stages = n_stages;
  const SmemLayout SL = smem_layout(n_stages);
  if (!g_cur->attrs_v1) {
    PLNERF_CUDA(cudaFuncSetAttribute(k_mlp_fwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    PLNERF_CUDA(cudaFuncSetAttribute(k_mlp_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    PLNERF_CUDA(cudaFuncSetAttribute(k_mlp_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    PLNERF_CUDA(cudaFuncSetAttribute(k_mlp_fwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    g_cur->attrs_v1 = true;
  }
  a.debug_flags = dbg_env("PLNERF_DEBUG_FLAGS");
  a.trace = g_trace;
  a.n_tiles = ceil_div(a.M, TILE_M);
  const unsigned grid = (unsigned)((a.n_tiles < g_num_sms) ? a.n_tiles : g_num_sms);
  ProfRec rec{nullptr, nullptr, a.M};
  if (g_prof_on) { rec.e0 = get_event(); rec.e1 = get_event(); cudaEventRecord(rec.e0, st); }
  if (mode == 1) k_mlp_fwd<1><<<grid, NUM_THREADS, SL.total, st>>>(a);
  else if (mode == 2) k_mlp_fwd<2><<<grid, NUM_THREADS, SL.total, st>>>(a);
  else if (mode == 3) k_mlp_fwd<3><<<grid, NUM_THREADS, SL.total, st>>>(a);
  else k_mlp_fwd<0><<<grid, NUM_THREADS, SL.total, st>>>(a);
  if (g_prof_on) { cudaEventRecord(rec.e1, st); std::lock_guard<std::mutex> lk(g_prof_mu); g_prof.push_back(rec); }
  PLNERF_LAUNCH_CHECK("k_mlp_fwd");
  return PLNERF_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
size_t mlp_packed_bytes(const plnerf_net_desc* d, int precision) {
  NetPlan P;
  if (build_plan(d, precision, nullptr, &P)) return 0;
  return (size_t)P.weight_bytes + (size_t)P.tail_floats * 4;
}

int mlp_pack(const plnerf_net_desc* d, const plnerf_net_params* p, int precision, void* packed, cudaStream_t st) {
  PLNERF_CHECK_ARG(p && packed, "pack_weights: null argument");
  PLNERF_CHECK_ARG(((uintptr_t)packed & 15) == 0, "pack_weights: packed buffer must be 16-byte aligned");
  PackArgs a;
  int rc = build_plan(d, precision, p, &a.plan);
  if (rc) return rc;
  for (int i = 0; i < d->D; ++i) PLNERF_CHECK_ARG(p->pts_w[i] && p->pts_b[i], "pack_weights: pts_linears.%d missing", i);
  if (d->use_viewdirs) PLNERF_CHECK_ARG(p->views_w && p->views_b && p->feature_w && p->feature_b && p->alpha_w && p->alpha_b && p->rgb_w && p->rgb_b, "pack_weights: viewdirs head parameters missing");
  else PLNERF_CHECK_ARG(p->output_w && p->output_b, "pack_weights: output_linear missing");
  a.dst = static_cast<uint8_t*>(packed);
  a.tail = reinterpret_cast<float*>(a.dst + a.plan.weight_bytes);
  a.prm = *p;
  const unsigned nblk = (unsigned)(a.plan.weight_bytes / KS_BYTES);
  k_pack_weights<<<nblk, 256, 0, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_pack_weights");
  k_pack_tail<<<(unsigned)ceil_div(a.plan.tail_floats, 256), 256, 0, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_pack_tail");
  return PLNERF_OK;
}

size_t mlp_workspace_bytes(const plnerf_net_desc* d, int64_t n_rays) {
  if (!d || n_rays < 0) return 0;
  return d->use_viewdirs ? (size_t)n_rays * 128 * sizeof(float) + 256 : 256;
}

// One launch for what render_rays needs per ray before the first network query: stratified depths + the view bias of both
// networks.  Returns 1 when the configuration is not covered (no view directions / wide direction encodings): the caller then
// uses launch_stratified_z and lets mlp_query compute its own view bias.
int launch_ray_setup(const plnerf_net_desc* cd, const void* cpacked, const plnerf_net_desc* fd, const void* fpacked, int precision,
                     int multires_views, const float* rays, int64_t n, int stride, int Ns, int lindisp, int perturb,
                     const float* t_rand, uint64_t seed, uint64_t ray0, float* z, float* vb_c, float* vb_f, cudaStream_t st,
                     float* dirpe) {
  if (!cd->use_viewdirs || cd->input_ch_views > 27 || stride < 11) return 1;
  if (fd && (fd->input_ch_views != cd->input_ch_views || !fd->use_viewdirs)) return 1;
  int rc = query_device();
  if (rc) return rc;
  NetPlan pc, pf;
  rc = build_plan(cd, precision, nullptr, &pc);
  if (rc) return rc;
  if (fd) { rc = build_plan(fd, precision, nullptr, &pf); if (rc) return rc; }
  if (fd && (pf.dirw_off != pc.dirw_off || pf.views_b_off != pc.views_b_off)) return 1;
  RaySetupArgs a;
  a.tail_c = reinterpret_cast<const float*>(static_cast<const uint8_t*>(cpacked) + pc.weight_bytes);
  a.tail_f = fd ? reinterpret_cast<const float*>(static_cast<const uint8_t*>(fpacked) + pf.weight_bytes) : a.tail_c;
  a.views_b_off = pc.views_b_off; a.dirw_off = pc.dirw_off; a.icv = cd->input_ch_views; a.multires_views = multires_views;
  a.rays = rays; a.stride = stride; a.n = n; a.vb_c = vb_c; a.vb_f = fd ? vb_f : nullptr; a.dirpe = dirpe;
  a.Ns = Ns; a.lindisp = lindisp; a.perturb = perturb; a.t_rand = t_rand; a.seed = seed; a.ray0 = ray0; a.z = z;
  // 16 rays per block iteration for large batches, fewer (down to 4) while that keeps every SM busy
  int rpi = RS_RAYS;
  while (rpi > 4 && ceil_div(n, rpi) < 4 * g_num_sms) rpi >>= 1;
  a.rpi = rpi;
  const int64_t blocks = ceil_div(n, rpi);
  k_ray_setup<<<(unsigned)(blocks < 8 * g_num_sms ? blocks : 8 * g_num_sms), 128, 0, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_ray_setup");
  return PLNERF_OK;
}

static int run_mlp_common(const plnerf_net_desc* d, const void* packed, int precision, MlpArgs& a, int64_t vb_rows,
                          int multires_views, const float* rays, int stride, const float* x_emb, int x_ld, void* ws,
                          size_t ws_bytes, cudaStream_t st, int mode = -1, float* dirpe_out = nullptr,
                          const float* viewbias_pre = nullptr) {
  int rc = query_device();
  if (rc) return rc;
  rc = build_plan(d, precision, nullptr, &a.plan);
  if (rc) return rc;
  PLNERF_CHECK_ARG(packed && ((uintptr_t)packed & 15) == 0, "packed weights null or misaligned");
  a.w = static_cast<const uint8_t*>(packed);
  a.tail = reinterpret_cast<const float*>(a.w + a.plan.weight_bytes);
  a.viewbias = nullptr;
  if (d->use_viewdirs && viewbias_pre) {
    a.viewbias = viewbias_pre;          // computed by launch_ray_setup
  } else if (d->use_viewdirs) {
    const size_t need = (size_t)vb_rows * 128 * sizeof(float);
    if (!ws || ws_bytes < need) { set_error("workspace too small: need %zu bytes, got %zu", need, ws_bytes); return PLNERF_E_WORKSPACE; }
    float* vb = static_cast<float*>(ws);
    const int64_t vb_blocks = ceil_div(vb_rows, VB_RAYS);
    k_viewbias<<<(unsigned)(vb_blocks < 8 * g_num_sms ? vb_blocks : 8 * g_num_sms), 128, 0, st>>>(a.tail, a.plan.views_b_off, a.plan.dirw_off, d->input_ch_views, multires_views,
                                                  rays, stride, x_emb, x_ld, d->input_ch, vb_rows, vb, dirpe_out);
    PLNERF_LAUNCH_CHECK("k_viewbias");
    a.viewbias = vb;
  }
  return launch_mlp(a, st, mode);
}

int mlp_query(const plnerf_net_desc* d, const void* packed, int precision, int multires, int multires_views,
              const float* rays, int64_t n, int stride, const float* z, int S, float* raw, int raw_stride,
              void* ws, size_t ws_bytes, cudaStream_t st, const FusedComposite* fc, bool need_raw, bool* fused,
              const float* viewbias_pre) {
  if (fused) *fused = false;
  PLNERF_CHECK_ARG(d && rays && z && raw, "network_query: null argument");
  PLNERF_CHECK_ARG(n >= 0 && S > 0, "network_query: bad sizes");
  if (n == 0) return PLNERF_OK;
  const int want_ic = multires < 0 ? 3 : 3 + 6 * multires;
  PLNERF_CHECK_ARG(want_ic == d->input_ch && multires <= 10, "network_query: multires=%d does not match input_ch=%d", multires, d->input_ch);
  if (d->use_viewdirs) {
    const int want_icv = multires_views < 0 ? 3 : 3 + 6 * multires_views;
    PLNERF_CHECK_ARG(want_icv == d->input_ch_views, "network_query: multires_views=%d does not match input_ch_views=%d", multires_views, d->input_ch_views);
    PLNERF_CHECK_ARG(stride >= 11, "network_query: use_viewdirs needs rays with a viewdir (stride >= 11)");
  }
  PLNERF_CHECK_ARG(stride >= 8, "network_query: ray stride must be >= 8");
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  a.rays = rays; a.stride = stride; a.z = z; a.S = S; a.multires = multires;
  a.x_emb = nullptr; a.x_ld = 0; a.vb_div = S; a.M = n * S; a.out = raw; a.out_stride = raw_stride;
  if (fc && precision == PLNERF_PREC_BF16 && d->use_viewdirs && S <= 2 * TILE_M && !dbg_env("PLNERF_NO_FUSE")) {
    a.fuse_comp = 1;
    a.comp_mode = fc->mode;
    CompositeArgs& c = a.comp;
    c.raw = nullptr; c.raw_stride = 4; c.z = z; c.rays = rays; c.n = n; c.stride = stride; c.S = S;
    c.color_mode = fc->color_mode; c.white_bkgd = fc->white_bkgd; c.farcolorfix = fc->farcolorfix;
    c.noise = fc->noise; c.noise_std = fc->noise_std; c.seed = fc->seed; c.ray0 = fc->ray0; c.noise_stream = fc->noise_stream;
    c.rgb_map = fc->rgb_map; c.disp_map = fc->disp_map; c.acc_map = fc->acc_map; c.depth_map = fc->depth_map;
    c.weights = fc->weights; c.tau = fc->tau; c.T = fc->T;
    a.skip_out = need_raw ? 0 : 1;
  }
  const int rc = run_mlp_common(d, packed, precision, a, n, multires_views, rays, stride, nullptr, 0, ws, ws_bytes, st, -1, nullptr,
                                viewbias_pre);
  if (rc == PLNERF_OK && a.fuse_comp && fused) *fused = true;     // (launch_mlp clears fuse_comp when it falls back to k_mlp_fwd)
  return rc;
}

int mlp_forward_embedded(const plnerf_net_desc* d, const void* packed, int precision, const float* x, int64_t m,
                         float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  PLNERF_CHECK_ARG(d && x && out, "mlp_forward: null argument");
  PLNERF_CHECK_ARG(m >= 0, "mlp_forward: bad size");
  if (m == 0) return PLNERF_OK;
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  a.S = 1; a.multires = 0;
  a.x_emb = x; a.x_ld = d->input_ch + (d->use_viewdirs ? d->input_ch_views : 0);
  a.vb_div = 1; a.M = m; a.out = out; a.out_stride = d->use_viewdirs ? 4 : d->output_ch;
  return run_mlp_common(d, packed, precision, a, m, 0, nullptr, 0, x, a.x_ld, ws, ws_bytes, st);
}



int profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { g_evt_pool.push_back(r.e0); g_evt_pool.push_back(r.e1); }
  g_prof.clear();
  g_prof_on = on != 0;
  return PLNERF_OK;
}

int profile_read(double* ms_sum, int64_t* launches, int64_t* rows) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0; int64_t nr = 0;
  for (auto& r : g_prof) {
    PLNERF_CUDA(cudaEventSynchronize(r.e1));
    float t = 0;
    PLNERF_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t; nr += r.rows;
  }
  if (ms_sum) *ms_sum = ms;
  if (launches) *launches = (int64_t)g_prof.size();
  if (rows) *rows = nr;
  return PLNERF_OK;
}

// ---------------------------------------------------------------------------------------------
// training entry points
// ---------------------------------------------------------------------------------------------
struct StashPtrs { uint8_t* in; uint8_t* dy; uint32_t* masks; float* dirpe; size_t total; };
static StashPtrs carve_stash(const TrainLayout& tl, int64_t rows, int64_t n_rays, uint8_t* base) {
  const int64_t tiles = ceil_div(rows, TILE_M);
  auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  StashPtrs s;
  size_t off = 0;
  s.in = base + off; off += up((size_t)tiles * tl.in_tile_bytes);
  s.dy = base + off; off += up((size_t)tiles * tl.dy_tile_bytes);
  s.masks = reinterpret_cast<uint32_t*>(base + off); off += up((size_t)tiles * tl.mask_tile_words * 4);
  s.dirpe = reinterpret_cast<float*>(base + off); off += up((size_t)n_rays * 32 * 4);
  s.total = off;
  return s;
}

size_t mlp_train_stash_bytes(const plnerf_net_desc* d, int64_t n_rays, int S) {
  TrainLayout tl;
  if (!d || build_train_layout(d, &tl)) return 0;
  return carve_stash(tl, n_rays * S, n_rays, nullptr).total;
}

int mlp_query_train(const plnerf_net_desc* d, const void* packed, int multires, int multires_views, const float* rays,
                    int64_t n, int stride, const float* z, int S, float* raw, int raw_stride, void* stash,
                    size_t stash_bytes, void* ws, size_t ws_bytes, cudaStream_t st, const float* viewbias_pre,
                    const float* dirpe_pre) {
  PLNERF_CHECK_ARG(d && rays && z && raw && stash, "network_query_train: null argument");
  PLNERF_CHECK_ARG((viewbias_pre == nullptr) == (dirpe_pre == nullptr), "network_query_train: view bias and direction encoding go together");
  PLNERF_CHECK_ARG(n >= 0 && S > 0 && stride >= (d->use_viewdirs ? 11 : 8), "network_query_train: bad sizes (rays of a view-direction network need a viewdir)");
  if (n == 0) return PLNERF_OK;
  PLNERF_CHECK_ARG(((uintptr_t)stash & 1023) == 0, "network_query_train: stash must be 1024-byte aligned");
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  int rc = build_train_layout(d, &a.tl);
  if (rc) return rc;
  const StashPtrs sp = carve_stash(a.tl, n * S, n, static_cast<uint8_t*>(stash));
  if (stash_bytes < sp.total) { set_error("training stash too small: need %zu bytes, got %zu", sp.total, stash_bytes); return PLNERF_E_WORKSPACE; }
  const int want_ic = multires < 0 ? 3 : 3 + 6 * multires;
  const int want_icv = multires_views < 0 ? 3 : 3 + 6 * multires_views;
  PLNERF_CHECK_ARG(want_ic == d->input_ch && (!d->use_viewdirs || want_icv == d->input_ch_views) && multires <= 10,
                   "network_query_train: multires/multires_views do not match the network");
  a.rays = rays; a.stride = stride; a.z = z; a.S = S; a.multires = multires;
  a.vb_div = S; a.M = n * S; a.out = raw; a.out_stride = raw_stride;
  a.in_stash = sp.in; a.dy_stash = sp.dy; a.masks = sp.masks; a.dirpe = dirpe_pre ? dirpe_pre : sp.dirpe;
  return run_mlp_common(d, packed, PLNERF_PREC_BF16, a, n, multires_views, rays, stride, nullptr, 0, ws, ws_bytes, st, 2, sp.dirpe,
                        viewbias_pre);
}

size_t mlp_packed_bwd_bytes(const plnerf_net_desc* d) {
  NetPlan P;
  if (!d || build_dgrad_plan(d, nullptr, &P)) return 0;
  return (size_t)P.weight_bytes;
}

int mlp_pack_bwd(const plnerf_net_desc* d, const plnerf_net_params* p, void* packed, cudaStream_t st) {
  PLNERF_CHECK_ARG(d && p && packed, "pack_weights_bwd: null argument");
  PLNERF_CHECK_ARG(((uintptr_t)packed & 15) == 0, "pack_weights_bwd: buffer must be 16-byte aligned");
  PackArgs a;
  int rc = build_dgrad_plan(d, p, &a.plan);
  if (rc) return rc;
  a.dst = static_cast<uint8_t*>(packed);
  a.tail = nullptr;
  a.prm = *p;
  k_pack_weights<<<(unsigned)(a.plan.weight_bytes / KS_BYTES), 256, 0, st>>>(a);
  PLNERF_LAUNCH_CHECK("k_pack_weights(bwd)");
  return PLNERF_OK;
}

int mlp_pack_train(int n_nets, const plnerf_net_desc* const* descs, const plnerf_net_params* const* params, void* const* packed,
                   void* const* packed_bwd, cudaStream_t st) {
  PLNERF_CHECK_ARG(n_nets >= 1 && 2 * n_nets <= MAX_PACK_ARGS && descs && params && packed && packed_bwd, "pack_weights_train: need 1 or 2 networks");
  PackTrainArgs t;
  memset(&t, 0, sizeof(t));
  int nj = 0, na = 0, blk = 0;
  auto job = [&](int kind, int arg, int64_t blocks) { t.kind[nj] = (int8_t)kind; t.arg[nj] = (int8_t)arg; t.blk0[nj] = blk; blk += (int)blocks; ++nj; };
  for (int i = 0; i < n_nets; ++i) {
    const plnerf_net_desc* d = descs[i];
    const plnerf_net_params* p = params[i];
    PLNERF_CHECK_ARG(d && p && packed[i] && packed_bwd[i], "pack_weights_train: null argument (network %d)", i);
    PLNERF_CHECK_ARG((((uintptr_t)packed[i] | (uintptr_t)packed_bwd[i]) & 15) == 0, "pack_weights_train: buffers must be 16-byte aligned");
    for (int k = 0; k < d->D; ++k) PLNERF_CHECK_ARG(p->pts_w[k] && p->pts_b[k], "pack_weights_train: pts_linears.%d missing", k);
    if (d->use_viewdirs) PLNERF_CHECK_ARG(p->views_w && p->views_b && p->feature_w && p->feature_b && p->alpha_w && p->alpha_b && p->rgb_w && p->rgb_b, "pack_weights_train: viewdirs head parameters missing");
    else PLNERF_CHECK_ARG(p->output_w && p->output_b, "pack_weights_train: output_linear missing");
    PackArgs& f = t.args[na];
    int rc = build_plan(d, PLNERF_PREC_BF16, p, &f.plan);
    if (rc) return rc;
    f.dst = static_cast<uint8_t*>(packed[i]);
    f.tail = reinterpret_cast<float*>(f.dst + f.plan.weight_bytes);
    f.prm = *p;
    job(0, na, f.plan.weight_bytes / KS_BYTES);
    job(1, na, ceil_div(f.plan.tail_floats, 256));
    ++na;
    PackArgs& b = t.args[na];
    rc = build_dgrad_plan(d, p, &b.plan);
    if (rc) return rc;
    b.dst = static_cast<uint8_t*>(packed_bwd[i]);
    b.tail = nullptr;
    b.prm = *p;
    job(0, na, b.plan.weight_bytes / KS_BYTES);
    ++na;
  }
  t.n_jobs = nj;
  t.blk0[nj] = blk;
  k_pack_train<<<(unsigned)blk, 256, 0, st>>>(t);
  PLNERF_LAUNCH_CHECK("k_pack_train");
  return PLNERF_OK;
}

int mlp_query_bwd(const plnerf_net_desc* d, const void* packed_fwd, const void* packed_bwd, int64_t n, int S,
                  const float* g_raw, int g_stride, void* stash, size_t stash_bytes, const plnerf_net_grads* g,
                  cudaStream_t st) {
  PLNERF_CHECK_ARG(d && packed_fwd && packed_bwd && g_raw && stash && g, "network_query_bwd: null argument");
  PLNERF_CHECK_ARG(n >= 0 && S > 0 && g_stride >= 4, "network_query_bwd: bad sizes");
  if (n == 0) return PLNERF_OK;
  int rc = query_device();
  if (rc) return rc;
  NetPlan fwd;
  rc = build_plan(d, PLNERF_PREC_BF16, nullptr, &fwd);
  if (rc) return rc;
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  rc = build_train_layout(d, &a.tl);
  if (rc) return rc;
  rc = build_dgrad_plan(d, nullptr, &a.plan);
  if (rc) return rc;
  const StashPtrs sp = carve_stash(a.tl, n * S, n, static_cast<uint8_t*>(stash));
  if (stash_bytes < sp.total) { set_error("training stash too small: need %zu bytes, got %zu", sp.total, stash_bytes); return PLNERF_E_WORKSPACE; }
  for (int i = 0; i < d->D; ++i) PLNERF_CHECK_ARG(g->pts_w[i] && g->pts_b[i], "network_query_bwd: missing gradient buffer for pts_linears.%d", i);
  if (d->use_viewdirs)
    PLNERF_CHECK_ARG(g->views_w && g->views_b && g->feature_w && g->feature_b && g->alpha_w && g->alpha_b && g->rgb_w && g->rgb_b,
                     "network_query_bwd: missing head gradient buffers");
  else
    PLNERF_CHECK_ARG(g->output_w && g->output_b, "network_query_bwd: missing output_linear gradient buffers");
  // (1) input-gradient chain -> dY stash
  a.w = static_cast<const uint8_t*>(packed_bwd);
  a.tail = reinterpret_cast<const float*>(static_cast<const uint8_t*>(packed_fwd) + fwd.weight_bytes);
  a.S = S; a.vb_div = S; a.M = n * S;
  a.in_stash = sp.in; a.dy_stash = sp.dy; a.masks = sp.masks; a.dirpe = sp.dirpe;
  a.g_raw = g_raw; a.g_stride = g_stride;
  // bias gradients of the heads = column sums of g_raw (output_linear: channels 0-2 and 3 of its bias)
  a.d_rgb_b = d->use_viewdirs ? g->rgb_b : g->output_b;
  a.d_alpha_b = d->use_viewdirs ? g->alpha_b : g->output_b + 3;
  rc = launch_mlp(a, st, 3);
  if (rc) return rc;
  // (2) weight gradients
  WgradArgs w;
  int w_ctas = 0;
  memset(&w, 0, sizeof(w));
  w.tl = a.tl; w.in_stash = sp.in; w.dy_stash = sp.dy; w.n_tiles = ceil_div(n * S, TILE_M);
  int ni = 0;
  auto add = [&](int dy_idx, int halves, int in_idx, float* dW, int ldw, int col0, int ncols, float* db) {
    for (int h = 0; h < halves; ++h) {
      WgradItem& it = w.items[ni++];
      it.dy_idx = dy_idx; it.half = h; it.in_idx = in_idx; it.swapped = 0; it.dW = dW + col0; it.row_stride = ldw; it.col_stride = 1;
      it.c_first = 0; it.ncols = ncols; it.db = db;
    }
  };
  // heads (CUDA-core layers in the forward): d alpha_linear.weight[0, k] = sum_m g_alpha[m] h_{D-1}[m, k],
  // d rgb_linear.weight[c, k] = sum_m g_rgb[m, c] hv[m, k] -- the activation tensor on the M side against the 16-wide g tile
  auto add_head = [&](int in_idx, int halves, float* dW, int col_stride, int c_first, int ncols) {
    for (int h = 0; h < halves; ++h) {
      WgradItem& it = w.items[ni++];
      it.dy_idx = a.tl.dy_head; it.half = h; it.in_idx = in_idx; it.swapped = 1; it.dW = dW; it.row_stride = 1; it.col_stride = col_stride;
      it.c_first = c_first; it.ncols = ncols; it.db = nullptr;
    }
  };
  auto is_skip = [&](int i) { for (int k = 0; k < d->n_skips; ++k) if (d->skips[k] == i) return true; return false; };
  const TrainLayout& T = a.tl;
  for (int i = 0; i < d->D; ++i) {
    const bool first = (i == 0), skip_in = (i > 0) && is_skip(i - 1);
    if (first) add(T.dy_h0 + i, 2, T.idx_pe, g->pts_w[i], d->input_ch, 0, d->input_ch, g->pts_b[i]);
    else if (skip_in) {
      add(T.dy_h0 + i, 2, T.idx_pe, g->pts_w[i], 256 + d->input_ch, 0, d->input_ch, nullptr);
      add(T.dy_h0 + i, 2, T.idx_h0 + i - 1, g->pts_w[i], 256 + d->input_ch, d->input_ch, 256, g->pts_b[i]);
    } else add(T.dy_h0 + i, 2, T.idx_h0 + i - 1, g->pts_w[i], 256, 0, 256, g->pts_b[i]);
  }
  if (d->use_viewdirs) {
    add(T.dy_feat, 2, T.idx_h0 + d->D - 1, g->feature_w, 256, 0, 256, g->feature_b);
    add(T.dy_views, 1, T.idx_feat, g->views_w, 256 + d->input_ch_views, 0, 256, g->views_b);
    add(T.dy_views, 1, T.idx_dir, g->views_w, 256 + d->input_ch_views, 256, d->input_ch_views, nullptr);
    add_head(T.idx_h0 + d->D - 1, 2, g->alpha_w, 0, 3, 1);
    add_head(T.idx_hv, 1, g->rgb_w, 128, 0, 3);
  } else {
    // d output_linear.weight[c, k] = sum_m g_out[m, c] h_{D-1}[m, k], c < 4 (further channels: no gradient)
    add_head(T.idx_h0 + d->D - 1, 2, g->output_w, 256, 0, d->output_ch < 4 ? d->output_ch : 4);
  }
  w.n_items = ni;
  // CTAs per item.  A CTA's time is set by the number of tiles it walks (~1 us per tile whatever the
  // tile's bytes; measured -- byte-proportional counts were 40% slower), so every layer item gets the same count; the light
  // head items take what is left of the SMs.  At most one CTA per SM in total, never more CTAs than tiles.
  {
    int n_light = 0;
    for (int i = 0; i < ni; ++i) n_light += w.items[i].swapped;
    const int n_main = ni - n_light;
    int wg_sms = g_num_sms;
    { const int force = dbg_env("PLNERF_WGRAD_SMS"); if (force > 0 && force < wg_sms) wg_sms = force; }
    int s_main = (wg_sms - n_light) / (n_main > 0 ? n_main : 1);
    if (s_main < 1) s_main = 1;
    int s_light = n_light ? (wg_sms - s_main * n_main) / n_light : 1;
    if (s_light > s_main) s_light = s_main;
    if (s_light < 1) s_light = 1;
    int c0 = 0;
    for (int i = 0; i < ni; ++i) {
      int sp = w.items[i].swapped ? s_light : s_main;
      if ((int64_t)sp > w.n_tiles) sp = (int)w.n_tiles;
      w.items[i].splits = sp; w.items[i].cta0 = c0; c0 += sp;
    }
    w_ctas = c0;
  }
  const size_t wsmem = WG_STAGES * WG_STAGE_BYTES + 1024 + 16 * WG_STAGES + 32;
  if (!g_cur->attrs_wgrad) { PLNERF_CUDA(cudaFuncSetAttribute(k_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem)); g_cur->attrs_wgrad = true; }
  k_wgrad<<<(unsigned)w_ctas, WG_THREADS, wsmem, st>>>(w);
  PLNERF_LAUNCH_CHECK("k_wgrad");
  return PLNERF_OK;
}




#ifdef PLNERF_DEBUG
#include "debug_kernels.cuh"
#endif

}  // namespace plnerf
