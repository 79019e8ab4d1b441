// Kernel 3b -- fused multi-head attention over T x 14 x 14 (+cls) tokens, head_dim 64.
//
// Replaces timm 1.0.20 Attention.forward (fused branch: F.scaled_dot_product_attention with
// scale head_dim**-0.5) as constructed by instageo/model/pritvhi.py:445-457.  Input is the
// qkv GEMM output in timm's own layout [B*N, 3*D] = (q | k | v) x (head, 64), so the
// reshape/permute of the reference costs nothing: Q, K and V tiles are 2-D TMA boxes of that
// matrix.  Output is written as [B*N, D] (head-major), the exact operand of the proj GEMM.
//
// PERSISTENT: 2 x (number of SMs) CTAs, each walking work items (one item = 128 query rows of one (batch, head))
// with its K / V / Q loads and QK^T issue running two KV blocks ahead, across item boundaries (see the comment on
// the kernel).  tcgen05 throughout:
//   S_j  = Q K_j^T  : UMMA 128x64x16, both operands K-major (128-byte swizzle), S double-buffered in TMEM
//   O   += P_j V_j  : UMMA 128x64x16, A = P_j read from TENSOR MEMORY (bf16 pairs written by the softmax warps with
//                     tcgen05.st over the columns of S_j they have just pulled into registers: no shared-memory
//                     round trip, no fence.proxy.async, and the 32 KB of P buffers per CTA went to deeper K / V rings),
//                     B = V_j used MN-major straight from its [kv, 64] tile (no transpose pass)
//   L   += P_j 1    : UMMA 128x8x16 against a tile of ones: the softmax denominator is accumulated by the
//                     tensor core from the SAME bf16-rounded probabilities as the numerator (summing it in the
//                     softmax threads' registers instead cost 12 % of the kernel: 64 dependent-on-MUFU adds per block)
// O and L stay in TMEM for the whole KV loop.  The softmax warps (thread <-> TMEM lane <-> query row)
// therefore do nothing per block but: pull the 64 scores, take their max, exponentiate against a
// reference max, and write P_j.  The reference max is LAZY: it is only raised when a block's max exceeds
// it by more than 2^8 in the exp2 domain (probabilities stay <= 256, harmless in f32/bf16), and only then
// is O/L rescaled in TMEM (tcgen05.ld -> multiply -> tcgen05.st, between PV_{j-1} and PV_j).  With
// attention logits of trained or random-init ViTs that happens in the first block or two; every other
// block costs 1 FFMA + 1 MUFU + 1/3 FMNMX3 + 1/2 F2F per score.
// In the last KV block only the 16-column groups that contain valid keys are exponentiated.
// Measurement tooling: tools/attn_ablate.sh (timing ablations, -DATTN_PROFILE in-kernel clock64 phase profile and
// event trace of CTA 0), tools/attn_profile.py, tools/attn_trace.py, tools/tma_latency.py; findings in DESIGN.md.
#include "ig_ops.cuh"

namespace attn {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int Q_BYTES = BQ * HD * 2;    // 16384
constexpr int KV_BYTES = BKV * HD * 2;  // 8192
constexpr int K_STAGES = 5, V_STAGES = 6;
constexpr int OFF_Q = 0;                              // [128 x 64] bf16
constexpr int OFF_K = OFF_Q + Q_BYTES;
constexpr int OFF_V = OFF_K + K_STAGES * KV_BYTES;
constexpr int OFF_ONES = OFF_V + V_STAGES * KV_BYTES;  // [8 x 64] bf16 ones (K-major B operand of the L MMA)
constexpr int OFF_BAR = OFF_ONES + 1024;
constexpr int SMEM_TOTAL = 1024 + OFF_BAR + 512;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;
// TMEM columns (256 = the whole allocation of a CTA, two CTAs per SM): S buffers at 0 / 64 (128 x 64 f32), ONE P
// buffer at 128 (128 x 64 bf16 = 32 columns of bf16 pairs), O at 160..223, L at 224..231.  P has its OWN columns:
// with P_j written over S_j the buffer could only be handed back to QK^T_{j+2} after PV_j, and that chain (publish
// P_j -> PV_j -> free -> QK^T_{j+2} -> S_{j+2}) was longer than the softmax of block j+1 it has to hide under.  One
// buffer is enough: PV_{j-1}, issued when block j-1 was published, completes long before block j has been
// exponentiated.
constexpr int COL_S = 0, COL_P = 128, COL_O = 160, COL_L = 224;
// ATTN_P_ALIAS = 1: P_j is written over columns [0, 32) of S_j instead (every softmax warp overwrites only the lanes
// whose scores it holds in registers) and the S buffer returns to QK^T_{j+2} with PV_j's commit.  Measured on B200
// (64 x 589 x 12 heads, CUDA events, same process): aliased 127-131 us, own P columns 138-142 us -- the shorter
// softmax stream (no second barrier pair) wins over the shorter dependency chain, so the alias form ships.
#ifndef ATTN_P_ALIAS
#define ATTN_P_ALIAS 1
#endif
constexpr float RESCALE_LOG2 = 8.f;                 // raise the reference max only for jumps > 2^8
static_assert(OFF_ONES % 1024 == 0 && OFF_V % 1024 == 0 && OFF_K % 1024 == 0, "UMMA operand tiles are 1024-byte aligned");
static_assert(2 * SMEM_TOTAL <= 232448 - 2048, "two CTAs per SM");
// Timing ablations (tools/attn_ablate.sh; never defined in the shipped library): 1 = no MUFU (exp2 -> identity),
// 2 = no P stores, 5 / 6 / 7 = one instead of four L / PV / QK UMMAs per KV block.  Results are wrong by construction.
#ifndef ATTN_ABLATE
#define ATTN_ABLATE 0
#endif
// In-kernel phase timing (-DATTN_PROFILE, tools/attn_ablate.sh): clock64 deltas of lane 0 of softmax warp 2 and of
// the MMA warp, summed over all CTAs into g_attn_prof[16]; read back through ig_attention_profile().
#ifdef ATTN_PROFILE
__device__ unsigned long long g_attn_prof[32];
// event trace of CTA 0 (ATTN_PROFILE == 3): [role 0 producer | 1 mma | 2 softmax warp 2][4096] x {event, block, clock}
__device__ long long g_attn_trace[3][4096][3];
__device__ int g_attn_trace_n[3];
__device__ __forceinline__ void attn_trace(int role, int ev, int blk) {
#if ATTN_PROFILE == 3
  if (blockIdx.x == 0) {
    const int i = atomicAdd(&g_attn_trace_n[role], 1);
    if (i < 4096) {
      g_attn_trace[role][i][0] = ev;
      g_attn_trace[role][i][1] = blk;
      g_attn_trace[role][i][2] = clock64();
    }
  }
#endif
}
#define TRACE(role, ev, blk) do { if (lane == 0 || (role) == 0) attn_trace(role, ev, blk); } while (0)
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, t0, t1) do { if (lane == 0 && warp <= 2) atomicAdd(&g_attn_prof[slot], static_cast<unsigned long long>((t1) - (t0))); } while (0)
#else
#define PROF_T(var)
#define PROF_ADD(slot, t0, t1)
#define TRACE(role, ev, blk)
#endif
#ifndef ATTN_POLY
#define ATTN_POLY 1
#endif
#ifndef ATTN_STALE_MAX
#define ATTN_STALE_MAX 1
#endif
// 2^x on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial of 2^f on [-0.5, 0.5], relative error
// 7.7e-5 -- far below the bf16 rounding of P): round-to-nearest integer part through the 1.5 * 2^23 magic add, whose
// low mantissa bits are then shifted into the exponent field.  x is clamped to >= -125 (such probabilities round to
// zero against a row's maximum anyway).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  const float p = fmaf(f, fmaf(f, fmaf(f, 0.05508868396282196f, 0.24260404706001282f), 0.6932762265205383f),
                       0.9999289512634277f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
__device__ __forceinline__ float exp2_or_ablate(float x) {
#if ATTN_ABLATE == 1
  return x;
#else
  return ig::ex2(x);
#endif
}

// PERSISTENT kernel: gridDim.x = min(work items, 2 x SMs) CTAs, each walks work items w = blockIdx.x, +gridDim.x, ...
// (one item = 128 query rows of one (batch, head), query tile fastest).  All rings and double buffers run on a
// KV-block counter that keeps counting across items, so the TMA producer and the QK^T issue run ahead INTO THE
// NEXT ITEM while the softmax warps finish the current one.  The first version launched one short-lived CTA per
// item and spent 28 % of every CTA's life outside the steady state (in-kernel clock64 profile, per CTA of ~33.6 k
// clk: 5.1 k waiting for the first S tile behind the Q / K_0 loads, 1.4 k for the last PV, 2.0 k in the output
// epilogue, 0.85 k of set-up), with TMEM allocated and freed 3840 times per launch.
__global__ void __launch_bounds__(THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmkv,
                 __nv_bfloat16* __restrict__ out, int N, int D, int nqt, int heads, int total_items) {
  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the __shared__ array itself (not a round trip through uintptr_t) keeps the pointer in
  // the shared address space: LDS/STS with 32-bit addresses instead of generic LD/ST with 64-bit address math
  uint8_t* smem = smem_raw + ((1024u - (ig::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;    // [1]
  uint64_t* q_empty = bars + 1;   // [1]  every QK^T of the item has completed
  uint64_t* k_full = bars + 2;                  // [K_STAGES]
  uint64_t* k_empty = k_full + K_STAGES;        // [K_STAGES]
  uint64_t* v_full = k_empty + K_STAGES;        // [V_STAGES]
  uint64_t* v_empty = v_full + V_STAGES;        // [V_STAGES]
  uint64_t* s_full = v_empty + V_STAGES;        // [2]
  uint64_t* s_free = s_full + 2;                // [2]  the softmax warps hold S_j in registers
  uint64_t* p_full = s_free + 2;                // [2] (own P columns: only [0], one P buffer)
  uint64_t* p_free = p_full + 2;                // [2] (own P columns only, [0]): PV_j has consumed P_j
  uint64_t* o_done = p_free + 2;                // [2]  PV of a block (and everything before it) has completed
  uint64_t* o_free = o_done + 2;                // [1]  the softmax warps have read O / L of the finished item
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 1);
  static_assert((2 + 2 * K_STAGES + 2 * V_STAGES + 11) * 8 + 4 <= 512, "barrier block");

  const int warp = ig::warp_idx_uniform(), lane = threadIdx.x & 31;
  PROF_T(cta0);
  const int nb = (N + BKV - 1) / BKV;
  const int items_per_bh = nqt;

  if (warp == 0 && lane == 0) {
    ig::tma_prefetch_desc(&tmq);
    ig::tma_prefetch_desc(&tmkv);
    ig::mbar_init(q_full, 1);
    ig::mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) {
      ig::mbar_init(&s_full[s], 1);
      ig::mbar_init(&s_free[s], ATTN_P_ALIAS ? 1 : 4);
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
    ig::mbar_init(o_free, 4);
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
  ig::pdl_launch_dependents();  // programmatic dependent launch: see ig::launch
  ig::pdl_wait();

  // Both issuing warps need these (uniform values; descriptors are formed ADDITIVELY from base words computed
  // once: start address >> 4 in the low word, a stage / buffer / K-step is a constant added to it, so a UMMA costs
  // one or two uniform-datapath adds instead of a shift-mask-or chain per operand).
  // Work items are dealt ROUND-ROBIN (item n of this CTA is w_begin + n * w_step): the five query tiles of one
  // (batch, head) then run on five CTAs at the same time and share their K / V tiles through L2 -- DRAM reads equal
  // the qkv matrix once (174 MB at B = 64).  Contiguous runs per CTA were measured too: same time, but the K / V
  // tiles of a (batch, head) were evicted between its query tiles and DRAM reads rose to 308 MB (ncu).
  // What still delays PV is a heavy tail of the K / V tile loads -- ~8 % of them land > 10 000 clk after their issue
  // (in-kernel probe, tools/attn_profile.py), against 440 clk (L2 hit) / 1100 clk (DRAM) for the same box on an
  // idle SM (tools/tma_latency.py), spread over all block indices of an item; open question for the next round.
  const int w_begin = blockIdx.x, w_step = gridDim.x;
  const int my_items = (total_items - w_begin + w_step - 1) / w_step;
  const int total_blocks = my_items * nb;
  const uint32_t smem_base = ig::smem_u32(smem);

  if (warp == 0) {
    // ===================== warp 0: TMA producer + QK^T issuer =====================
    // A single issuing warp is a serial instruction stream: with one warp issuing QK^T and PV (12 UMMAs, 5
    // commits and 4 barrier waits per KV block, ~1300 clk) it, not the softmax warps, set the pace of the CTA
    // (event trace of CTA 0, tools/attn_trace.py).  QK^T issue therefore lives here, next to the two TMA loads
    // per block, and warp 1 issues only PV.
    // Order per block g:  V_g load (buffer released by PV_{g-2})  ->  K_{g+3} load (+ the next item's Q in
    // front of its first block; released by QK_g)  ->  S_{g+2} = Q K_{g+2}^T (S buffer handed back by the
    // softmax warps as soon as S_g is in their registers).  K runs three blocks ahead of V, QK^T two blocks
    // ahead of PV, across item boundaries.
    const uint32_t idesc = ig::umma_idesc_bf16(BQ, BKV, 0, 0);    // S = Q K^T (N = 64 kv)
    const uint32_t q_lo = ig::umma_desc_lo(smem_base + OFF_Q);        // K-major tiles: LBO 16, SBO 1024
    const uint32_t k_lo = ig::umma_desc_lo(smem_base + OFF_K);
    // The issuing warps are single serial instruction streams (one dependent instruction every ~6 clk): all
    // ring / parity / coordinate state is CARRIED and updated incrementally, an item is decoded (three integer
    // divisions) once per item and not per block.  With divisions and modulos per block this warp needed ~3000 clk
    // per KV block and was the slowest actor of the CTA (event trace, tools/attn_trace.py).
    // K stream
    int k_left = total_blocks, k_j = 0, k_w = w_begin, k_st = 0, k_col = 0, k_row = 0;
    uint32_t k_par = 1;  // parity to wait for on k_empty (fresh barrier: passes)
    auto decode = [&](int w, int& col_h, int& row0, int& qrow) {
      const int bh = w / items_per_bh, qt = w - bh * items_per_bh;
      const int b = bh / heads, h = bh - b * heads;
      col_h = h * HD, row0 = b * N, qrow = b * N + qt * BQ;
    };
    int dummy_q;
    if (total_blocks > 0) {
      int ch, r0;
      decode(k_w, ch, r0, dummy_q);
      k_col = D + ch, k_row = r0;
      if (ig::elect_one()) {
        ig::mbar_expect_tx(q_full, Q_BYTES);
        ig::tma_load_2d(smem + OFF_Q, &tmq, q_full, ch, dummy_q);
      }
      __syncwarp();
    }
    auto emit_k = [&]() {
      ig::mbar_wait(&k_empty[k_st], k_par);
      if (ig::elect_one()) {
        TRACE(0, 1, k_left);  // K load issued
        ig::mbar_expect_tx(&k_full[k_st], KV_BYTES);
        ig::tma_load_2d(smem + OFF_K + k_st * KV_BYTES, &tmkv, &k_full[k_st], k_col, k_row);
      }
      __syncwarp();
      --k_left;
#if ATTN_ABLATE != 11
      k_row += BKV;
#endif
      if (++k_st == K_STAGES) k_st = 0, k_par ^= 1;
      if (++k_j == nb) {
        k_j = 0, k_w += w_step;
        if (k_left > 0) {
          int ch, r0;
          decode(k_w, ch, r0, dummy_q);
          k_col = D + ch, k_row = r0;
        }
      }
    };
    // QK^T stream
    int q_left = total_blocks, q_j = 0, q_n = 0, q_st = 0, q_sb = 0;
    uint32_t q_kpar = 0, q_spar = 1;  // parities to wait for on k_full / s_free
    auto issue_qk = [&]() {
      PROF_T(w0);
      if (q_j == 0) ig::mbar_wait(q_full, q_n & 1);
      ig::mbar_wait(&k_full[q_st], q_kpar);
      PROF_T(w1);
      ig::mbar_wait(&s_free[q_sb], q_spar);
      PROF_T(w2);
      TRACE(1, 10, q_left);    // QK waits satisfied
      ig::tc_fence_after();
      const uint32_t dk = k_lo + q_st * (KV_BYTES >> 4);
      const uint32_t d_s = tmem_base + COL_S + q_sb * BKV;
      const bool last = q_j == nb - 1;
      if (ig::elect_one()) {
#pragma unroll
        for (int k = 0; k < (ATTN_ABLATE == 7 ? 1 : HD / 16); ++k)
          ig::umma_bf16(d_s, ig::umma_desc_pack(q_lo + 2 * k), ig::umma_desc_pack(dk + 2 * k), idesc, k > 0);
        ig::umma_commit(&s_full[q_sb]);
        ig::umma_commit(&k_empty[q_st]);
        if (last) ig::umma_commit(q_empty);
      }
      __syncwarp();
      PROF_T(w3);
      PROF_ADD(11, w0, w1);  // QK: wait Q / K
      PROF_ADD(6, w1, w2);   // QK: wait S buffer free
      PROF_ADD(8, w2, w3);   // QK: issue
      TRACE(1, 11, q_left);  // QK issued
      --q_left;
      if (++q_st == K_STAGES) q_st = 0, q_kpar ^= 1;
      q_sb ^= 1;
      if (q_sb == 0) q_spar ^= 1;
      if (++q_j == nb) {
        // The item's last QK^T is in flight: as soon as it has completed, the (single) Q tile is reloaded for
        // the next item -- two blocks before that item's first QK^T is issued.
        if (q_left > 0) {
          int ch, r0, qrow;
          decode(w_begin + (q_n + 1) * w_step, ch, r0, qrow);
          ig::mbar_wait(q_empty, q_n & 1);
          if (ig::elect_one()) {
            ig::mbar_expect_tx(q_full, Q_BYTES);
            ig::tma_load_2d(smem + OFF_Q, &tmq, q_full, ch, qrow);
          }
          __syncwarp();
        }
        q_j = 0, ++q_n;
      }
    };
    // Per block g:  K_{g+4} load (buffer released by QK_g)  ->  S_{g+2} = Q K_{g+2}^T (S buffer handed back by the
    // softmax warps as soon as S_g is in their registers).  Every stream runs two blocks ahead of its consumer,
    // across item boundaries: inside a tensor-busy SM a tile lands ~2500 clk after its load is issued (440 clk on
    // an otherwise idle SM with the tile in L2: tools/tma_latency.py), more than one block period.
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
    const uint32_t one_lo = ig::umma_desc_lo(smem_base + OFF_ONES);
    // V is consumed MN-major straight from its [kv, 64] tile: 16 kv rows of 128 bytes per K step, 8-row groups
    // 1024 B apart (LBO = SBO = 1024)
    const uint32_t v_lo = (((smem_base + OFF_V) & 0x3FFFF) >> 4) | ((1024u >> 4) << 16);
    // V stream (its buffers are released by this warp's own PV commits)
    int v_left = total_blocks, v_j = 0, v_w = w_begin, v_st = 0, v_col = 0, v_row = 0;
    uint32_t v_par = 1;
    auto decode_v = [&]() {
      const int bh = v_w / items_per_bh;
      const int b = bh / heads, h = bh - b * heads;
      v_col = 2 * D + h * HD, v_row = b * N;
    };
    if (total_blocks > 0) decode_v();
#ifdef ATTN_PROFILE
    long long v_issue_t[V_STAGES] = {0, 0, 0};
#endif
    auto emit_v = [&]() {
      ig::mbar_wait(&v_empty[v_st], v_par);
#ifdef ATTN_PROFILE
      v_issue_t[v_st] = clock64();
#endif
      if (ig::elect_one()) {
        TRACE(0, 2, v_left);  // V load issued
        ig::mbar_expect_tx(&v_full[v_st], KV_BYTES);
        ig::tma_load_2d(smem + OFF_V + v_st * KV_BYTES, &tmkv, &v_full[v_st], v_col, v_row);
      }
      __syncwarp();
      --v_left;
#if ATTN_ABLATE != 11
      v_row += BKV;
#endif
      if (++v_st == V_STAGES) v_st = 0, v_par ^= 1;
      if (++v_j == nb) {
        v_j = 0, v_w += w_step;
        if (v_left > 0) decode_v();
      }
    };
    for (int i = 0; i < V_STAGES - 1 && i < total_blocks; ++i) emit_v();
    int n = 0, j = 0, sv = 0, pb = 0;
    uint32_t vf_par = 0, pf_par = 0;  // parities to wait for on v_full / p_full
    for (int g = 0; g < total_blocks; ++g) {
      PROF_T(m1);
#ifdef ATTN_PROFILE
      const bool v_was_ready = ig::mbar_try_wait(&v_full[sv], vf_par);
#endif
      ig::mbar_wait(&v_full[sv], vf_par);
      PROF_T(m1b);
#ifdef ATTN_PROFILE
      if (!v_was_ready) { PROF_ADD(19, v_issue_t[sv], m1b); PROF_ADD(20, 0, 1); PROF_ADD(22 + (j < 9 ? j : 9), 0, 1); }
      PROF_ADD(21, 0, 1);
#endif
      TRACE(1, 14, g);  // V ready
      PROF_ADD(12, m1, m1b);  // wait V_g
      // P_g is in tensor memory and any rescale of O / L is finished.  Aliased P: the softmax warps may publish P_{g+1}
      // (other S buffer) before this warp has seen P_g, so the two buffers have a barrier each.
      ig::mbar_wait(&p_full[ATTN_P_ALIAS ? pb : 0], pf_par);
      // first block of an item overwrites O / L: the previous item's epilogue must have read them
      if (j == 0 && n > 0) ig::mbar_wait(o_free, (n - 1) & 1);
      PROF_T(m2);
      TRACE(1, 12, g);  // V and P ready
      ig::tc_fence_after();
      const uint32_t tp = ATTN_P_ALIAS ? tmem_base + COL_S + pb * BKV : tmem_base + COL_P;   // P_g: 128 lanes x 32 columns
      const uint32_t dv = v_lo + sv * (KV_BYTES >> 4);
      const uint32_t acc0 = j > 0 ? 1u : 0u;
      if (ig::elect_one()) {
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // A = P_g from tensor memory: 16 kv (one K step) = 8 columns of bf16 pairs
          if (ATTN_ABLATE != 6 || k == 0)
            ig::umma_bf16_ts(tmem_base + COL_O, tp + 8 * k, ig::umma_desc_pack(dv + k * (2048 >> 4)), idesc_o, k > 0 ? 1u : acc0);
          if (ATTN_ABLATE != 5 || k == 0)
            ig::umma_bf16_ts(tmem_base + COL_L, tp + 8 * k, ig::umma_desc_pack(one_lo + 2 * k), idesc_l, k > 0 ? 1u : acc0);
        }
        ig::umma_commit(&o_done[pb]);
        ig::umma_commit(&v_empty[sv]);
        if (ATTN_P_ALIAS) ig::umma_commit(&s_free[pb]);   // the S / P buffer may be overwritten by QK^T of block g + 2
        else ig::umma_commit(&p_free[0]);
      }
      __syncwarp();
      PROF_T(m3);
      TRACE(1, 13, g);  // PV issued
      PROF_ADD(9, m1b, m2);  // wait P_g (, O free)
      PROF_ADD(10, m2, m3);  // issue PV_g
      if (++j == nb) j = 0, ++n;
      if (++sv == V_STAGES) sv = 0, vf_par ^= 1;
      pb ^= 1;
      if (!ATTN_P_ALIAS || pb == 0) pf_par ^= 1;
      // V_{g+2} goes into the buffer PV_{g-1} released (that commit was issued one iteration ago)
      if (v_left > 0) emit_v();
    }
  } else {
    // ===================== softmax / output warps (one thread per query row) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_s = t_lane + COL_S, t_o = t_lane + COL_O, t_p = t_lane + COL_P, t_l = t_lane + COL_L;
    const float sl2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
    const float jump = RESCALE_LOG2 / sl2;           // the same threshold in raw score units
    const uint32_t NEG_INF = 0xff800000u;
    uint32_t sc[64];
    uint32_t(&sa)[32] = *reinterpret_cast<uint32_t(*)[32]>(sc);
    uint32_t(&sb2)[32] = *reinterpret_cast<uint32_t(*)[32]>(sc + 32);
    int g = 0;  // KV blocks consumed so far (all items): buffer index and barrier parity
    for (int w = w_begin; w < total_items; w += w_step) {
      const int qt = w % items_per_bh, bh = w / items_per_bh;
      const int h = bh % heads, b = bh / heads;
      const int row0 = b * N, q0 = qt * BQ;
      float m_ref = -INFINITY;
      float m_seen = -INFINITY;   // running maximum of the row over the blocks seen so far (ATTN_STALE_MAX)
      for (int j = 0; j < nb; ++j, ++g) {
        const int kv0 = j * BKV, sb = g & 1;
        const uint32_t par = (g >> 1) & 1;
        const int nvalid = min(BKV, N - kv0);  // warp-uniform
        PROF_T(c0);
        ig::mbar_wait(&s_full[sb], par);
        PROF_T(c1);
        TRACE(2, 20 + 100 * warp, g);  // S ready (every softmax warp)
        ig::tc_fence_after();
        ig::tmem_ld32(t_s + sb * BKV, sa);
        ig::tmem_ld32(t_s + sb * BKV + 32, sb2);
        ig::tmem_ld_wait();
#if !ATTN_P_ALIAS
        // the scores are in registers: hand the S buffer back so QK^T of block g+2 can start
        ig::tc_fence_before();
        __syncwarp();
        if (lane == 0) ig::mbar_arrive(&s_free[sb]);
#endif
        PROF_T(c2);
        if (nvalid < BKV) {  // last block: masked columns become -inf (=> exp 0, ignored by the max)
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= nvalid) sc[i] = NEG_INF;
        }
        // block maximum: 8 independent chains of 3-input maxima (4 dependent steps each) and a 3-level tree (the
        // 4-chain version was bound by the latency of its 8-deep chains: ~220 clk per block in the phase profile)
        auto block_max = [&]() {
          float mx[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) mx[c] = fmaxf(__uint_as_float(sc[c]), __uint_as_float(sc[c + 8]));
#pragma unroll
          for (int i = 16; i < 64; i += 16)
#pragma unroll
            for (int c = 0; c < 8; ++c)
              mx[c] = fmaxf(mx[c], fmaxf(__uint_as_float(sc[i + c]), __uint_as_float(sc[i + c + 8])));
          return fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
        };
        // ---- lazy rescale: O and L of this warp's 32 rows are multiplied by 2^(old - new) in TMEM (rows that keep
        // their reference by 1).  PV_{g-1} must have completed; PV_g cannot start before this warp arrives on p_full.
        auto rescale_to = [&](float m_new) {
          const float alpha = ig::ex2((m_ref - m_new) * sl2);
          m_ref = m_new;
          ig::mbar_wait(&o_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
          ig::tc_fence_after();
          uint32_t t[32];
#pragma unroll
          for (int c = 0; c < HD; c += 32) {
            ig::tmem_ld32(t_o + c, t);
            ig::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            ig::tmem_st32(t_o + c, t);
          }
          const uint32_t lv = ig::tmem_ld1(t_l);
          ig::tmem_ld_wait();
          ig::tmem_st1(t_l, __float_as_uint(__uint_as_float(lv) * alpha));
          ig::tmem_st_wait();
        };
        // ---- exponentials -> P_g (bf16 pairs); fully masked 16-column groups become zeros.  A share of the scores
        // takes the FMA-pipe polynomial instead of MUFU.EX2 (the MUFU issues one warp instruction per 8 clk and is
        // shared by the softmax warps of both resident CTAs): ATTN_POLY = 1 one in four, 2 one in two.
        uint32_t pk[32];
        auto exps = [&](float mc) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            if (gq * 16 < nvalid) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float a0 = fmaf(__uint_as_float(sc[gq * 16 + 2 * i]), sl2, -mc);
                const float a1 = fmaf(__uint_as_float(sc[gq * 16 + 2 * i + 1]), sl2, -mc);
                const float p0 = exp2_or_ablate(a0);
                const float p1 = (ATTN_POLY == 2 || (ATTN_POLY == 1 && (i & 1))) ? ex2_poly(a1) : exp2_or_ablate(a1);
                pk[gq * 8 + i] = ig::pack_bf16(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) pk[gq * 8 + i] = 0u;
            }
          }
        };
#if ATTN_STALE_MAX
        // The reference maximum of a row trails by one block: block j is exponentiated against the reference the
        // EARLIER blocks established (raised, with the O / L rescale, when they exceeded it by more than 2^8), so
        // its own maximum is not on the critical path -- the FMNMX chains overlap the MUFU-bound exponentials.  Any
        // common reference gives the same softmax; the only constraint is the float exponent range, checked
        // afterwards: a score more than 2^64 above the reference (never seen) takes the slow exact path.
        if (j == 0) {
          m_ref = block_max();   // PV_0 overwrites O / L (accumulate = 0): nothing to rescale
          m_seen = m_ref;
        } else {
          const bool need = m_seen > m_ref + jump;
          if (__any_sync(0xffffffffu, need)) rescale_to(need ? m_seen : m_ref);
        }
        PROF_T(c3);
        PROF_T(c4);
        TRACE(2, 21 + 100 * warp, g);  // exps start
        exps(m_ref * sl2);
        if (j > 0) {
          const float m_blk = block_max();
          const bool over = m_blk > m_ref + 8.f * jump;
          if (__any_sync(0xffffffffu, over)) {
            rescale_to(over ? m_blk : m_ref);
            exps(m_ref * sl2);
          }
          m_seen = fmaxf(m_seen, m_blk);
        }
#else
        const float m_blk = block_max();
        if (j == 0) {
          m_ref = m_blk;  // PV_0 overwrites O / L (accumulate = 0): nothing to rescale
        } else {
          const bool need = m_blk > m_ref + jump;
          if (__any_sync(0xffffffffu, need)) rescale_to(need ? m_blk : m_ref);
        }
        PROF_T(c3);
        PROF_T(c4);
        TRACE(2, 21 + 100 * warp, g);  // exps start
        exps(m_ref * sl2);
#endif
#if ATTN_ABLATE == 2
        if (pk[0] == 0x12345678u)
#endif
#if ATTN_P_ALIAS
        ig::tmem_st32(t_s + sb * BKV, pk);
#else
        ig::mbar_wait(&p_free[0], (g & 1) ^ 1);   // PV_{g-1} has consumed the previous tenant of the P buffer
        ig::tmem_st32(t_p, pk);
#endif
        PROF_T(c5);
        ig::tmem_st_wait();
        ig::tc_fence_before();         // orders the P store (and a rescale's tcgen05.st) before the MMA warp's PV_g
        __syncwarp();
        if (lane == 0) ig::mbar_arrive(&p_full[ATTN_P_ALIAS ? sb : 0]);
        PROF_T(c6);
        TRACE(2, 22 + 100 * warp, g);  // P_g published (every softmax warp)
        PROF_ADD(0, c0, c1);  // wait S
        PROF_ADD(1, c1, c2);  // TMEM load of S
        PROF_ADD(2, c2, c3);  // mask + max (+ rescale)
        PROF_ADD(3, c3, c4);  // wait P buffer free
        PROF_ADD(4, c4, c5);  // exponentials + P stores
        PROF_ADD(5, c5, c6);  // fences + arrive
      }
      // ---- item epilogue: O / L out of TMEM, then the accumulators are free for the next item's PV_0
      PROF_T(e0);
      ig::mbar_wait(&o_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
      PROF_T(e1);
      if (warp == 2) TRACE(2, 23, g);  // last PV done
      ig::tc_fence_after();
      ig::tmem_ld32(t_o, sa);
      ig::tmem_ld32(t_o + 32, sb2);
      const uint32_t lv = ig::tmem_ld1(t_l);
      ig::tmem_ld_wait();
      ig::tc_fence_before();
      __syncwarp();
      if (lane == 0) ig::mbar_arrive(o_free);
      const float inv = 1.f / __uint_as_float(lv);
      const int qrow = q0 + row;
      if (qrow < N) {
        __nv_bfloat16* orow = out + (static_cast<int64_t>(row0) + qrow) * D + h * HD;
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
      if (warp == 2) TRACE(2, 24, g);  // item stored
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
extern "C" int ig_attention_trace(long long* out, int* counts) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, attn::g_attn_trace, sizeof(long long) * 3 * 4096 * 3);
  cudaMemcpyFromSymbol(counts, attn::g_attn_trace_n, sizeof(int) * 3);
  int z[3] = {0, 0, 0};
  cudaMemcpyToSymbol(attn::g_attn_trace_n, z, sizeof(z));
  return 0;
}
extern "C" int ig_attention_profile(unsigned long long* out16) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, attn::g_attn_prof, sizeof(unsigned long long) * 32);
  unsigned long long z[32] = {0};
  cudaMemcpyToSymbol(attn::g_attn_prof, z, sizeof(z));
  return 0;
}
#endif

namespace ops {
// The two TMA descriptors of a launch depend only on (qkv pointer, B, N, heads): the model engine encodes them once
// per (batch, workspace) plan instead of once per block per forward.
int attention_maps(const void* qkv, int B, int N, int heads, CUtensorMap* tmq, CUtensorMap* tmkv) {
  IG_REQUIRE(B >= 1 && N >= 1 && heads >= 1, IG_ESHAPE, "attention: bad shape B=%d N=%d heads=%d", B, N, heads);
  const int D = heads * attn::HD;
  IG_TRY(ig_make_tmap_bf16(tmq, qkv, static_cast<uint64_t>(B) * N, 3 * D, 3 * D, attn::BQ, 64));
  IG_TRY(ig_make_tmap_bf16(tmkv, qkv, static_cast<uint64_t>(B) * N, 3 * D, 3 * D, attn::BKV, 64));
  return IG_OK;
}
int attention_planned(const CUtensorMap& tmq, const CUtensorMap& tmkv, void* out, int B, int N, int heads,
                      cudaStream_t st) {
  const int D = heads * attn::HD;
  static IgPerDevice configured = {};
  if (!configured.get()) {
    IG_CUDA_OK(cudaFuncSetAttribute(attn::attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    attn::SMEM_TOTAL));
    configured.set(1);
  }
  const int nqt = (N + attn::BQ - 1) / attn::BQ;
  IG_REQUIRE(static_cast<int64_t>(nqt) * heads * B < (1ll << 31), IG_ESHAPE, "attention: too many work items");
  const int total_items = nqt * heads * B;
  int grid = total_items < 2 * ig_num_sms() ? total_items : 2 * ig_num_sms();
  size_t smem = attn::SMEM_TOTAL;
  if (getenv("IG_ATTN_ONE_CTA")) {  // measurement aid: one CTA per SM (the second is kept out by the shared-memory request)
    smem = 150 * 1024;
    grid = total_items < ig_num_sms() ? total_items : ig_num_sms();
    cudaFuncSetAttribute(attn::attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  }
  ig::ProfScope prof(ig::PROF_ATTENTION, st);
  IG_CUDA_OK(ig::launch(attn::attention_kernel, dim3(grid), dim3(attn::THREADS), smem, st, true, tmq, tmkv,
                        static_cast<__nv_bfloat16*>(out), N, D, nqt, heads, total_items));
  return IG_OK;
}
int attention(const void* qkv, void* out, int B, int N, int heads, cudaStream_t st) {
  CUtensorMap tmq, tmkv;
  IG_TRY(attention_maps(qkv, B, N, heads, &tmq, &tmkv));
  return attention_planned(tmq, tmkv, out, B, N, heads, st);
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
