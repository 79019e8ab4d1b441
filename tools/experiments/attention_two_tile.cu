// EXPERIMENT, NOT BUILT INTO THE LIBRARY (kernel 3b): one CTA per SM with TWO query tiles sharing every
// K / V tile (8 softmax warps, 512 TMEM columns).  Parity-green (tests/test_gpu_ops.py, tests/test_gpu_model.py) and
// free of barrier waits (S 98 clk, P buffer 75 clk per KV block), yet 189 us against 170 us for the shipped
// two-CTAs-per-SM kernel at B = 64, N = 589, 12 heads: a softmax warp's KV block is bound by its own serial
// instruction stream (~1500 clk).  Kept as the starting point for the next round (shorten that stream: max pass off
// the critical path, exp2 partly on the FMA pipe, P in TMEM).  To try it: copy over csrc/attention.cu and rebuild.
// Kernel 3b -- fused multi-head attention over T x 14 x 14 (+cls) tokens, head_dim 64.
//
// Replaces timm 1.0.20 Attention.forward (fused branch: F.scaled_dot_product_attention with
// scale head_dim**-0.5) as constructed by instageo/model/pritvhi.py:445-457.  Input is the
// qkv GEMM output in timm's own layout [B*N, 3*D] = (q | k | v) x (head, 64), so the
// reshape/permute of the reference costs nothing: Q, K and V tiles are 2-D TMA boxes of that
// matrix.  Output is written as [B*N, D] (head-major), the exact operand of the proj GEMM.
//
// One CTA = 128 query rows of one (batch, head); 2 CTAs per SM.  tcgen05 throughout:
//   S_j  = Q K_j^T  : UMMA 128x64x16, both operands K-major (128-byte swizzle), S double-buffered in TMEM
//   O   += P_j V_j  : UMMA 128x64x16, A = P_j (bf16, written to swizzled smem by the softmax warps),
//                     B = V_j used MN-major straight from its [kv, 64] tile (no transpose pass)
//   L   += P_j 1    : UMMA 128x8x16 against a tile of ones: the softmax denominator is accumulated by the
//                     tensor core from the SAME bf16-rounded probabilities as the numerator
// O and L stay in TMEM for the whole KV loop.  The softmax warps (thread <-> TMEM lane <-> query row)
// therefore do nothing per block but: pull the 64 scores, take their max, exponentiate against a
// reference max, and write P_j.  The reference max is LAZY: it is only raised when a block's max exceeds
// it by more than 2^8 in the exp2 domain (probabilities stay <= 256, harmless in f32/bf16), and only then
// is O/L rescaled in TMEM (tcgen05.ld -> multiply -> tcgen05.st, between PV_{j-1} and PV_j).  With
// attention logits of trained or random-init ViTs that happens in the first block or two; every other
// block costs 1 FFMA + 1 MUFU + 1/3 FMNMX3 + 1/2 F2F per score.  The first version folded every block's
// O into 64 register accumulators (64 FFMA + 64 FADD per row and block on top of the exponentials) and ran
// at 45 % issue utilisation, 2.9x above the MUFU floor (profiles/r01_ncu_attention_before.txt).
// In the last KV block only the 16-column groups that contain valid keys are exponentiated.
#include <stdlib.h>

#include "ig_ops.cuh"

namespace attn {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int NT = 2;                    // query tiles per CTA (share every K / V tile)
constexpr int Q_BYTES = BQ * HD * 2;    // 16384 per tile
constexpr int KV_BYTES = BKV * HD * 2;  // 8192
constexpr int K_STAGES = 4, V_STAGES = 4;
constexpr int OFF_Q = 0;                              // [NT x 128 x 64] bf16: one 256-row TMA box
constexpr int OFF_K = OFF_Q + NT * Q_BYTES;
constexpr int OFF_V = OFF_K + K_STAGES * KV_BYTES;
constexpr int OFF_P = OFF_V + V_STAGES * KV_BYTES;   // [NT][2] x [128 x 64] bf16 (one swizzle atom column each)
constexpr int P_BYTES = BQ * BKV * 2;                 // 16384
constexpr int OFF_ONES = OFF_P + NT * 2 * P_BYTES;    // [8 x 64] bf16 ones (K-major B operand of the L MMA)
constexpr int OFF_BAR = OFF_ONES + 1024;
constexpr int SMEM_TOTAL = 1024 + OFF_BAR + 512;
constexpr int THREADS = 64 + NT * 128;
constexpr int TMEM_COLS = 512;
// TMEM columns: S[t][buf] at (2t + buf) * 64, O[t] at 256 + 64 t, L[t] at 384 + 8 t
constexpr int COL_S = 0, COL_O = 256, COL_L = 384;
constexpr float RESCALE_LOG2 = 8.f;                 // raise the reference max only for jumps > 2^8
#ifndef STAGGER_NS
#define STAGGER_NS 400
#endif
static_assert(OFF_ONES % 1024 == 0 && OFF_K % 1024 == 0 && OFF_P % 1024 == 0, "UMMA operand tiles are 1024-byte aligned");
static_assert(SMEM_TOTAL <= 232448, "over the 227 KB shared-memory limit");
// Timing ablations (tools/attn_ablate.sh; never defined in the shipped library): 1 = no MUFU (exp2 -> identity),
// 2 = no P stores, 5 / 6 / 7 = one instead of four L / PV / QK UMMAs per KV block.  Results are wrong by construction.
#ifndef ATTN_ABLATE
#define ATTN_ABLATE 0
#endif
// In-kernel phase timing (-DATTN_PROFILE, tools/attn_ablate.sh prof): clock64 deltas of lane 0 of the first softmax
// warp of tile 0 and of the two issuing warps, summed over all CTAs into g_attn_prof; ig_attention_profile() reads it.
#ifdef ATTN_PROFILE
__device__ unsigned long long g_attn_prof[32];
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, t0, t1) do { if (lane == 0 && warp <= 2) atomicAdd(&g_attn_prof[slot], static_cast<unsigned long long>((t1) - (t0))); } while (0)
#else
#define PROF_T(var)
#define PROF_ADD(slot, t0, t1)
#endif
__device__ __forceinline__ float exp2_or_ablate(float x) {
#if ATTN_ABLATE == 1
  return x;
#else
  return ig::ex2(x);
#endif
}

// PERSISTENT kernel, ONE CTA per SM, TWO query tiles per CTA.  gridDim.x = min(work items, SMs); a work item is 256
// query rows (two 128-row tiles) of one (batch, head); items are dealt round-robin, so the query-tile pairs of a
// (batch, head) run on neighbouring CTAs at the same time and share K / V through L2.  Both tiles consume the same
// K / V tiles out of shared memory (half the TMA and L2 traffic per query), each has its own S double buffer, O and
// L accumulators in TMEM, its own P double buffer and its own four softmax warps, so the exponentials of one tile
// overlap the TMEM loads, maxima, fences and barrier round trips of the other on every scheduler.
// Roles (320 threads): warp 0 = K / Q loads + QK^T issue, warp 1 = V loads + PV issue, warps 2-5 = softmax of
// tile 0, warps 6-9 = softmax of tile 1 (thread = query row).  All rings and double buffers run on KV-block
// counters that keep counting across items, so loads and QK^T run two blocks ahead INTO THE NEXT ITEM.
// History (B = 64, N = 589, 12 heads): one short-lived CTA per 128-row tile, two per SM: 185 us, 28 % of every CTA's
// life outside the steady state; persistent two-per-SM: 170 us, but ~8 % of the K / V tile loads of one CTA landed
// > 10 000 clk late whenever a second CTA shared the SM (11 of 38 400 with one CTA per SM), and one CTA per SM with a
// single tile ran exactly as fast as two -- the second CTA only added interference (in-kernel probes).
__global__ void __launch_bounds__(THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmkv,
                 __nv_bfloat16* __restrict__ out, int N, int D, int items_per_bh, int heads, int total_items) {
  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the __shared__ array itself (not a round trip through uintptr_t) keeps the pointer in
  // the shared address space: LDS/STS with 32-bit addresses instead of generic LD/ST with 64-bit address math
  uint8_t* smem = smem_raw + ((1024u - (ig::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;    // [1]
  uint64_t* q_empty = bars + 1;   // [1]  every QK^T of the item has completed
  uint64_t* k_full = bars + 2;    // [4]
  uint64_t* k_empty = bars + 6;   // [4]
  uint64_t* v_full = bars + 10;   // [4]
  uint64_t* v_empty = bars + 14;  // [4]
  uint64_t* s_full = bars + 18;   // [NT][2]
  uint64_t* s_free = bars + 22;   // [NT][2]
  uint64_t* p_full = bars + 26;   // [NT][2]
  uint64_t* p_free = bars + 30;   // [NT][2]
  uint64_t* o_done = bars + 34;   // [NT][2]  PV of a block (and everything before it) has completed
  uint64_t* o_free = bars + 38;   // [NT]     the softmax warps have read O / L of the finished item
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 40);

  const int warp = ig::warp_idx_uniform(), lane = threadIdx.x & 31;
  PROF_T(cta0);
  const int nb = (N + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    ig::tma_prefetch_desc(&tmq);
    ig::tma_prefetch_desc(&tmkv);
    ig::mbar_init(q_full, 1);
    ig::mbar_init(q_empty, 1);
    for (int s = 0; s < NT * 2; ++s) {
      ig::mbar_init(&s_full[s], 1);
      ig::mbar_init(&s_free[s], 4);
      ig::mbar_init(&p_full[s], 4);
      ig::mbar_init(&p_free[s], 1);
      ig::mbar_init(&o_done[s], 1);
    }
    for (int s = 0; s < K_STAGES; ++s) {
      ig::mbar_init(&k_full[s], 1);
      ig::mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < V_STAGES; ++s) {
      ig::mbar_init(&v_full[s], 1);
      ig::mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < NT; ++t) ig::mbar_init(&o_free[t], 4);
    ig::fence_barrier_init();
  }
  if (warp == 1) {
    ig::tmem_alloc(tmem_ptr, TMEM_COLS);
    ig::tmem_relinquish();
  }
  if (warp >= 2) {  // the ones tile (bf16 1.0 everywhere: the swizzle does not matter)
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + OFF_ONES);
    for (int i = threadIdx.x - 64; i < 256; i += THREADS - 64) ones[i] = 0x3f803f80u;
    ig::fence_proxy_async_smem();
  }
  ig::tc_fence_before();
  __syncthreads();
  ig::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  // Work items are dealt ROUND-ROBIN (item n of this CTA is w_begin + n * w_step): DRAM reads equal the qkv matrix
  // once (174 MB at B = 64); contiguous runs per CTA evicted K / V between the query tiles of a (batch, head)
  // (308 MB, ncu).  Tile t of an item is skipped by every role when it starts past the last query row.
  const int w_begin = blockIdx.x, w_step = gridDim.x;
  const int my_items = (total_items - w_begin + w_step - 1) / w_step;
  const int total_blocks = my_items * nb;
  const uint32_t smem_base = ig::smem_u32(smem);
  // an item's tile 1 exists unless the item is the last of its (batch, head) and N leaves it empty
  const int last_qt = items_per_bh - 1;
  const bool last_has_t1 = last_qt * (NT * BQ) + BQ < N;
  auto decode = [&](int w, int& col_h, int& row0, int& qt) {
    const int bh = w / items_per_bh;
    qt = w - bh * items_per_bh;
    const int b = bh / heads, h = bh - b * heads;
    col_h = h * HD, row0 = b * N;
  };

  if (warp == 0) {
    // ===================== warp 0: K / Q loads + QK^T issuer =====================
    // The issuing warps are single serial instruction streams (one dependent instruction every ~6 clk): descriptors
    // are formed ADDITIVELY from base words computed once, all ring / parity / coordinate state is CARRIED, and an
    // item is decoded (integer divisions) once per item, not per block.
    const uint32_t idesc = ig::umma_idesc_bf16(BQ, BKV, 0, 0);    // S = Q K^T (N = 64 kv)
    const uint32_t q_lo = ig::umma_desc_lo(smem_base + OFF_Q);        // K-major tiles: LBO 16, SBO 1024
    const uint32_t k_lo = ig::umma_desc_lo(smem_base + OFF_K);
    // K stream
    int k_left = total_blocks, k_j = 0, k_w = w_begin, k_st = 0, k_col = 0, k_row = 0;
    uint32_t k_par = 1;  // parity to wait for on k_empty (fresh barrier: passes)
    if (total_blocks > 0) {
      int ch, r0, qt;
      decode(k_w, ch, r0, qt);
      k_col = D + ch, k_row = r0;
      if (ig::elect_one()) {
        ig::mbar_expect_tx(q_full, NT * Q_BYTES);
        ig::tma_load_2d(smem + OFF_Q, &tmq, q_full, ch, r0 + qt * (NT * BQ));
      }
      __syncwarp();
    }
    auto emit_k = [&]() {
      ig::mbar_wait(&k_empty[k_st], k_par);
      if (ig::elect_one()) {
        ig::mbar_expect_tx(&k_full[k_st], KV_BYTES);
        ig::tma_load_2d(smem + OFF_K + k_st * KV_BYTES, &tmkv, &k_full[k_st], k_col, k_row);
      }
      __syncwarp();
      --k_left;
      k_row += BKV;
      if (++k_st == K_STAGES) k_st = 0, k_par ^= 1;
      if (++k_j == nb) {
        k_j = 0, k_w += w_step;
        if (k_left > 0) {
          int ch, r0, qt;
          decode(k_w, ch, r0, qt);
          k_col = D + ch, k_row = r0;
        }
      }
    };
    // QK^T stream (per-tile S buffer state: a skipped tile does not advance its buffers)
    int q_left = total_blocks, q_j = 0, q_n = 0, q_st = 0;
    int q_sb[NT] = {0, 0};
    uint32_t q_kpar = 0, q_spar[NT] = {1, 1};  // parities to wait for on k_full / s_free
    bool q_t1 = true;                          // tile 1 of the current item exists
    {
      int ch, r0, qt;
      if (total_blocks > 0) {
        decode(w_begin, ch, r0, qt);
        q_t1 = qt != last_qt || last_has_t1;
      }
    }
    auto issue_qk = [&]() {
      PROF_T(w0);
      if (q_j == 0) ig::mbar_wait(q_full, q_n & 1);
      ig::mbar_wait(&k_full[q_st], q_kpar);
      PROF_T(w1);
      const uint32_t dk = k_lo + q_st * (KV_BYTES >> 4);
      const bool last = q_j == nb - 1;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        if (t == 1 && !q_t1) break;
        const int sb = q_sb[t];
        ig::mbar_wait(&s_free[t * 2 + sb], q_spar[t]);
        ig::tc_fence_after();
        const uint32_t dq = q_lo + t * (Q_BYTES >> 4);
        const uint32_t d_s = tmem_base + COL_S + (t * 2 + sb) * BKV;
        if (ig::elect_one()) {
#pragma unroll
          for (int k = 0; k < (ATTN_ABLATE == 7 ? 1 : HD / 16); ++k)
            ig::umma_bf16(d_s, ig::umma_desc_pack(dq + 2 * k), ig::umma_desc_pack(dk + 2 * k), idesc, k > 0);
          ig::umma_commit(&s_full[t * 2 + sb]);
        }
        __syncwarp();
        q_sb[t] = sb ^ 1;
        if (sb == 1) q_spar[t] ^= 1;
      }
      if (ig::elect_one()) {
        ig::umma_commit(&k_empty[q_st]);
        if (last) ig::umma_commit(q_empty);
      }
      __syncwarp();
      PROF_T(w3);
      PROF_ADD(11, w0, w1);  // QK: wait Q / K
      PROF_ADD(8, w1, w3);   // QK: S-buffer waits + issue, both tiles
      --q_left;
      if (++q_st == K_STAGES) q_st = 0, q_kpar ^= 1;
      if (++q_j == nb) {
        // The item's last QK^T is in flight: as soon as it has completed, the Q tiles are reloaded for the next
        // item -- two blocks before that item's first QK^T is issued.
        if (q_left > 0) {
          int ch, r0, qt;
          decode(w_begin + (q_n + 1) * w_step, ch, r0, qt);
          q_t1 = qt != last_qt || last_has_t1;
          ig::mbar_wait(q_empty, q_n & 1);
          if (ig::elect_one()) {
            ig::mbar_expect_tx(q_full, NT * Q_BYTES);
            ig::tma_load_2d(smem + OFF_Q, &tmq, q_full, ch, r0 + qt * (NT * BQ));
          }
          __syncwarp();
        }
        q_j = 0, ++q_n;
      }
    };
    // Per block g:  K_{g+4} load (buffer released by QK_g)  ->  S_{g+2} = Q K_{g+2}^T for both tiles (S buffers handed
    // back by the softmax warps as soon as S_g is in their registers).
    for (int i = 0; i < K_STAGES && i < total_blocks; ++i) emit_k();
    if (total_blocks > 0) issue_qk();
    if (total_blocks > 1) issue_qk();
    for (int g = 0; g < total_blocks; ++g) {
      if (k_left > 0) emit_k();
      if (q_left > 0) issue_qk();
    }
  } else if (warp == 1) {
    // ===================== warp 1: V loads + PV issuer (whole warp walks the loop, one elected lane issues) =====
    const uint32_t idesc_o = ig::umma_idesc_bf16(BQ, HD, 0, 1);   // O += P V  (B = V, MN-major)
    const uint32_t idesc_l = ig::umma_idesc_bf16(BQ, 8, 0, 0);    // L += P 1  (B = ones, K-major, N = 8)
    const uint32_t p_lo = ig::umma_desc_lo(smem_base + OFF_P);
    const uint32_t one_lo = ig::umma_desc_lo(smem_base + OFF_ONES);
    // V is consumed MN-major straight from its [kv, 64] tile: 16 kv rows of 128 bytes per K step, 8-row groups
    // 1024 B apart (LBO = SBO = 1024)
    const uint32_t v_lo = (((smem_base + OFF_V) & 0x3FFFF) >> 4) | ((1024u >> 4) << 16);
    // V stream (its buffers are released by this warp's own PV commits)
    int v_left = total_blocks, v_j = 0, v_w = w_begin, v_st = 0, v_col = 0, v_row = 0;
    uint32_t v_par = 1;
    auto decode_v = [&]() {
      int ch, r0, qt;
      decode(v_w, ch, r0, qt);
      v_col = 2 * D + ch, v_row = r0;
    };
    if (total_blocks > 0) decode_v();
    auto emit_v = [&]() {
      ig::mbar_wait(&v_empty[v_st], v_par);
      if (ig::elect_one()) {
        ig::mbar_expect_tx(&v_full[v_st], KV_BYTES);
        ig::tma_load_2d(smem + OFF_V + v_st * KV_BYTES, &tmkv, &v_full[v_st], v_col, v_row);
      }
      __syncwarp();
      --v_left;
      v_row += BKV;
      if (++v_st == V_STAGES) v_st = 0, v_par ^= 1;
      if (++v_j == nb) {
        v_j = 0, v_w += w_step;
        if (v_left > 0) decode_v();
      }
    };
    for (int i = 0; i < V_STAGES - 1 && i < total_blocks; ++i) emit_v();
    int n = 0, j = 0, sv = 0;
    int pb[NT] = {0, 0}, n_t[NT] = {0, 0};  // per tile: P buffer, items finished (o_free parity)
    uint32_t vf_par = 0, pf_par[NT] = {0, 0};
    bool t1 = true;
    {
      int ch, r0, qt;
      if (total_blocks > 0) {
        decode(w_begin, ch, r0, qt);
        t1 = qt != last_qt || last_has_t1;
      }
    }
    for (int g = 0; g < total_blocks; ++g) {
      PROF_T(m1);
      ig::mbar_wait(&v_full[sv], vf_par);
      PROF_T(m1b);
      PROF_ADD(12, m1, m1b);  // wait V_g
      const uint32_t dv = v_lo + sv * (KV_BYTES >> 4);
      const uint32_t acc0 = j > 0 ? 1u : 0u;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        if (t == 1 && !t1) break;
        const int b = pb[t];
        PROF_T(p0);
        ig::mbar_wait(&p_full[t * 2 + b], pf_par[t]);  // P_g of tile t is in smem, any rescale of O / L finished
        // first block of an item overwrites O / L: the tile's previous epilogue must have read them
        if (j == 0 && n_t[t] > 0) ig::mbar_wait(&o_free[t], (n_t[t] - 1) & 1);
        PROF_T(p1);
        ig::tc_fence_after();
        const uint32_t dp = p_lo + (t * 2 + b) * (P_BYTES >> 4);
        if (ig::elect_one()) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            // A = P_g: 128 rows x 64 kv (one swizzle atom column), 32 bytes per K step
            const uint64_t da = ig::umma_desc_pack(dp + 2 * k);
            if (ATTN_ABLATE != 6 || k == 0)
              ig::umma_bf16(tmem_base + COL_O + t * HD, da, ig::umma_desc_pack(dv + k * (2048 >> 4)), idesc_o,
                            k > 0 ? 1u : acc0);
            if (ATTN_ABLATE != 5 || k == 0)
              ig::umma_bf16(tmem_base + COL_L + t * 8, da, ig::umma_desc_pack(one_lo + 2 * k), idesc_l, k > 0 ? 1u : acc0);
          }
          ig::umma_commit(&o_done[t * 2 + b]);
          ig::umma_commit(&p_free[t * 2 + b]);
        }
        __syncwarp();
        PROF_T(p2);
        if (t == 0) { PROF_ADD(9, p0, p1); PROF_ADD(10, p1, p2); }  // tile 0: wait P_g (, O free); issue PV_g
        pb[t] = b ^ 1;
        if (b == 1) pf_par[t] ^= 1;
      }
      if (ig::elect_one()) ig::umma_commit(&v_empty[sv]);
      __syncwarp();
      if (++sv == V_STAGES) sv = 0, vf_par ^= 1;
      if (++j == nb) {
        j = 0, ++n;
        ++n_t[0];
        if (t1) ++n_t[1];
        if (g + 1 < total_blocks) {
          int ch, r0, qt;
          decode(w_begin + n * w_step, ch, r0, qt);
          t1 = qt != last_qt || last_has_t1;
        }
      }
      // V_{g+3} goes into the buffer PV_{g-1} released (that commit was issued one iteration ago)
      if (v_left > 0) emit_v();
    }
  } else {
    // ===================== softmax / output warps (one thread per query row) =====================
    const int t = (warp - 2) >> 2;  // query tile of this warp
    const int quad = warp & 3;      // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_s = t_lane + COL_S + t * 2 * BKV, t_o = t_lane + COL_O + t * HD, t_l = t_lane + COL_L + t * 8;
    uint64_t* const sf = s_full + t * 2;
    uint64_t* const sfr = s_free + t * 2;
    uint64_t* const pf = p_full + t * 2;
    uint64_t* const pfr = p_free + t * 2;
    uint64_t* const od = o_done + t * 2;
    const float sl2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
    const float jump = RESCALE_LOG2 / sl2;           // the same threshold in raw score units
    const uint32_t NEG_INF = 0xff800000u;
    uint32_t sc[64];
    uint32_t(&sa)[32] = *reinterpret_cast<uint32_t(*)[32]>(sc);
    uint32_t(&sb2)[32] = *reinterpret_cast<uint32_t(*)[32]>(sc + 32);
    int g = 0;  // KV blocks consumed so far by THIS tile (all items): buffer index and barrier parity
    // The two tiles' softmax warps share the MUFU of every scheduler.  Started together they run in lockstep: both
    // in their exponential phase (2 x 490 clk of MUFU time) and then both out of it, the MUFU idle.  Tile 1 starts
    // half a block period late so that its exponentials fall into tile 0's load / max / fence phases.
    if (t == 1) __nanosleep(STAGGER_NS);
    for (int w = w_begin; w < total_items; w += w_step) {
      int col_h, row0, qt;
      decode(w, col_h, row0, qt);
      const int q0 = qt * (NT * BQ) + t * BQ;
      if (q0 >= N) continue;  // this tile of the item is empty: every role skips it
      float m_ref = -INFINITY;
      for (int j = 0; j < nb; ++j, ++g) {
        const int kv0 = j * BKV, sb = g & 1;
        const uint32_t par = (g >> 1) & 1;
        const int nvalid = min(BKV, N - kv0);  // warp-uniform
        PROF_T(c0);
        ig::mbar_wait(&sf[sb], par);
        PROF_T(c1);
        ig::tc_fence_after();
        ig::tmem_ld32(t_s + sb * BKV, sa);
        ig::tmem_ld32(t_s + sb * BKV + 32, sb2);
        ig::tmem_ld_wait();
        // the scores are in registers: hand the S buffer back so QK^T of block g+2 can start
        ig::tc_fence_before();
        __syncwarp();
        if (lane == 0) ig::mbar_arrive(&sfr[sb]);
        PROF_T(c2);
        if (nvalid < BKV) {  // last block: masked columns become -inf (=> exp 0, ignored by the max)
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= nvalid) sc[i] = NEG_INF;
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sc[i]), __uint_as_float(sc[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sc[i + 2]), __uint_as_float(sc[i + 3])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sc[i + 4]), __uint_as_float(sc[i + 5])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sc[i + 6]), __uint_as_float(sc[i + 7])));
        }
        const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (j == 0) {
          m_ref = m_blk;  // PV_0 overwrites O / L (accumulate = 0): nothing to rescale
        } else {
          const bool need = m_blk > m_ref + jump;
          if (__any_sync(0xffffffffu, need)) {
            // ---- lazy rescale: O and L of this warp's 32 rows are multiplied by 2^(old - new) in TMEM.
            // PV_{g-1} must have completed; PV_g cannot start before this warp arrives on p_full below.
            const float m_new = need ? m_blk : m_ref;
            const float alpha = ig::ex2((m_ref - m_new) * sl2);  // 1 for rows that keep their reference
            m_ref = m_new;
            ig::mbar_wait(&od[(g - 1) & 1], ((g - 1) >> 1) & 1);
            ig::tc_fence_after();
            uint32_t tt[32];
#pragma unroll
            for (int c = 0; c < HD; c += 32) {
              ig::tmem_ld32(t_o + c, tt);
              ig::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) tt[i] = __float_as_uint(__uint_as_float(tt[i]) * alpha);
              ig::tmem_st32(t_o + c, tt);
            }
            const uint32_t lv = ig::tmem_ld1(t_l);
            ig::tmem_ld_wait();
            ig::tmem_st1(t_l, __float_as_uint(__uint_as_float(lv) * alpha));
            ig::tmem_st_wait();
          }
        }
        const float mc = m_ref * sl2;
        // ---- exponentials -> P_g (bf16, swizzled smem); fully masked 16-column groups are written as zeros
        PROF_T(c3);
        ig::mbar_wait(&pfr[sb], par ^ 1);
        PROF_T(c4);
        uint8_t* prow = smem + OFF_P + (t * 2 + sb) * P_BYTES + row * 128;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          uint32_t pk[8];
          if (gq * 16 < nvalid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float p0 = exp2_or_ablate(fmaf(__uint_as_float(sc[gq * 16 + 2 * i]), sl2, -mc));
              const float p1 = exp2_or_ablate(fmaf(__uint_as_float(sc[gq * 16 + 2 * i + 1]), sl2, -mc));
              pk[i] = ig::pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = 0u;
          }
#if ATTN_ABLATE == 2
          if (pk[0] == 0x12345678u)
#endif
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int chunk = gq * 2 + q;  // 16-byte chunk inside the 128-byte row
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
        PROF_T(c5);
        ig::fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the UMMA async proxy
        ig::tc_fence_before();         // orders a rescale's tcgen05.st before the MMA warp's PV_g
        __syncwarp();
        if (lane == 0) ig::mbar_arrive(&pf[sb]);
        PROF_T(c6);
        PROF_ADD(0, c0, c1);  // wait S
        PROF_ADD(1, c1, c2);  // TMEM load of S
        PROF_ADD(2, c2, c3);  // mask + max (+ rescale)
        PROF_ADD(3, c3, c4);  // wait P buffer free
        PROF_ADD(4, c4, c5);  // exponentials + P stores
        PROF_ADD(5, c5, c6);  // fences + arrive
      }
      // ---- item epilogue: O / L out of TMEM, then the accumulators are free for the tile's next PV_0
      PROF_T(e0);
      ig::mbar_wait(&od[(g - 1) & 1], ((g - 1) >> 1) & 1);
      PROF_T(e1);
      ig::tc_fence_after();
      ig::tmem_ld32(t_o, sa);
      ig::tmem_ld32(t_o + 32, sb2);
      const uint32_t lv = ig::tmem_ld1(t_l);
      ig::tmem_ld_wait();
      ig::tc_fence_before();
      __syncwarp();
      if (lane == 0) ig::mbar_arrive(&o_free[t]);
      const float inv = 1.f / __uint_as_float(lv);
      const int qrow = q0 + row;
      if (qrow < N) {
        __nv_bfloat16* orow = out + (static_cast<int64_t>(row0) + qrow) * D + col_h;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 o;
          o.x = ig::pack_bf16(__uint_as_float(sc[8 * q + 0]) * inv, __uint_as_float(sc[8 * q + 1]) * inv);
          o.y = ig::pack_bf16(__uint_as_float(sc[8 * q + 2]) * inv, __uint_as_float(sc[8 * q + 3]) * inv);
          o.z = ig::pack_bf16(__uint_as_float(sc[8 * q + 4]) * inv, __uint_as_float(sc[8 * q + 5]) * inv);
          o.w = ig::pack_bf16(__uint_as_float(sc[8 * q + 6]) * inv, __uint_as_float(sc[8 * q + 7]) * inv);
          reinterpret_cast<uint4*>(orow)[q] = o;
        }
      }
      PROF_T(e2);
      PROF_ADD(7, e0, e1);   // wait for the item's last PV
      PROF_ADD(13, e1, e2);  // O load, normalise, store
    }
  }

  ig::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ig::tc_fence_after();
    ig::tmem_dealloc(tmem_base, TMEM_COLS);
  }
#ifdef ATTN_PROFILE
  {
    PROF_T(cta1);
    if (threadIdx.x == 64) {
      atomicAdd(&g_attn_prof[14], static_cast<unsigned long long>(cta1 - cta0));
      atomicAdd(&g_attn_prof[15], 1ull);
    }
  }
#endif
}

}  // namespace attn

#ifdef ATTN_PROFILE
extern "C" int ig_attention_profile(unsigned long long* out16) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, attn::g_attn_prof, sizeof(unsigned long long) * 32);
  unsigned long long z[32] = {0};
  cudaMemcpyToSymbol(attn::g_attn_prof, z, sizeof(z));
  return 0;
}
#endif

namespace ops {
int attention(const void* qkv, void* out, int B, int N, int heads, cudaStream_t st) {
  IG_REQUIRE(B >= 1 && N >= 1 && heads >= 1, IG_ESHAPE, "attention: bad shape B=%d N=%d heads=%d", B, N, heads);
  const int D = heads * attn::HD;
  static bool configured = false;
  if (!configured) {
    IG_CUDA_OK(cudaFuncSetAttribute(attn::attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    attn::SMEM_TOTAL));
    configured = true;
  }
  CUtensorMap tmq, tmkv;
  IG_TRY(ig_make_tmap_bf16(&tmq, qkv, static_cast<uint64_t>(B) * N, 3 * D, 3 * D, attn::NT * attn::BQ, 64));
  IG_TRY(ig_make_tmap_bf16(&tmkv, qkv, static_cast<uint64_t>(B) * N, 3 * D, 3 * D, attn::BKV, 64));
  const int items_per_bh = (N + attn::NT * attn::BQ - 1) / (attn::NT * attn::BQ);
  IG_REQUIRE(static_cast<int64_t>(items_per_bh) * heads * B < (1ll << 31), IG_ESHAPE, "attention: too many work items");
  const int total_items = items_per_bh * heads * B;
  const int grid = total_items < ig_num_sms() ? total_items : ig_num_sms();
  ig::ProfScope prof(ig::PROF_ATTENTION, st);
  attn::attention_kernel<<<grid, attn::THREADS, attn::SMEM_TOTAL, st>>>(tmq, tmkv, static_cast<__nv_bfloat16*>(out), N, D,
                                                                      items_per_bh, heads, total_items);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}
}  // namespace ops

extern "C" int ig_attention(const void* qkv, void* out, int B, int N, int heads, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(qkv && out, IG_EINVAL, "ig_attention: null pointer");
  return ops::attention(qkv, out, B, N, heads, static_cast<cudaStream_t>(stream));
}

#ifdef ATTN_PROFILE
// TMA latency probe: one thread loads [64 x 64] bf16 tiles of the qkv matrix (the K / V box of the attention kernel)
// one at a time and records issue -> mbarrier completion in clock cycles.  reps 0..n-1 touch new tiles (cold: DRAM or
// whatever L2 holds), reps n..2n-1 touch the same tiles again (L2 hits).
namespace attn {
__global__ void tma_latency_kernel(const __grid_constant__ CUtensorMap tmkv, long long* out, int n, int row_step, int col0) {
  __shared__ __align__(1024) uint8_t tile[KV_BYTES];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    ig::tma_prefetch_desc(&tmkv);
    ig::mbar_init(&bar, 1);
    ig::fence_barrier_init();
    for (int i = 0; i < 2 * n; ++i) {
      const int r = (i % n) * row_step + blockIdx.x * 64;
      const long long t0 = clock64();
      ig::mbar_expect_tx(&bar, KV_BYTES);
      ig::tma_load_2d(tile, &tmkv, &bar, col0, r);
      ig::mbar_wait(&bar, i & 1);
      out[blockIdx.x * 2 * n + i] = clock64() - t0;
    }
  }
}
}  // namespace attn
extern "C" int ig_debug_tma_latency(const void* qkv, int rows, int D3, long long* out_dev, int n, int row_step, int col0, int ctas) {
  CUtensorMap tmkv;
  IG_TRY(ig_make_tmap_bf16(&tmkv, qkv, rows, D3, D3, attn::BKV, 64));
  attn::tma_latency_kernel<<<ctas, 32>>>(tmkv, out_dev, n, row_step, col0);
  IG_CUDA_OK(cudaDeviceSynchronize());
  return IG_OK;
}
#endif
