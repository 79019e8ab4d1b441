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
#include "ig_ops.cuh"

namespace attn {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int Q_BYTES = BQ * HD * 2;    // 16384
constexpr int KV_BYTES = BKV * HD * 2;  // 8192
constexpr int KV_STAGES = 3;
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
constexpr int OFF_P = OFF_V + KV_STAGES * KV_BYTES;  // 2 x [128 x 64] bf16 (one swizzle atom column each)
constexpr int P_BYTES = BQ * BKV * 2;                 // 16384
constexpr int OFF_ONES = OFF_P + 2 * P_BYTES;         // [8 x 64] bf16 ones (K-major B operand of the L MMA)
constexpr int OFF_BAR = OFF_ONES + 1024;
constexpr int SMEM_TOTAL = 1024 + OFF_BAR + 256;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0, COL_O = 128, COL_L = 192;  // S buffers at 0 / 64, O at 128..191, L at 192..199
constexpr float RESCALE_LOG2 = 8.f;                 // raise the reference max only for jumps > 2^8
// Timing ablations (tools/attn_ablate.sh; never defined in the shipped library): 1 = no MUFU (exp2 -> identity),
// 2 = no P stores, 3 = no TMEM loads of S, 4 = no max pass, 5 / 6 / 7 = one instead of four L / PV / QK UMMAs per
// KV block.  Results are wrong by construction.
#ifndef ATTN_ABLATE
#define ATTN_ABLATE 0
#endif
// In-kernel phase timing (-DATTN_PROFILE, tools/attn_ablate.sh): clock64 deltas of lane 0 of softmax warp 2 and of
// the MMA warp, summed over all CTAs into g_attn_prof[16]; read back through ig_attention_profile().
#ifdef ATTN_PROFILE
__device__ unsigned long long g_attn_prof[16];
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, t0, t1) do { if (lane == 0 && (warp == 2 || warp == 1)) atomicAdd(&g_attn_prof[slot], static_cast<unsigned long long>((t1) - (t0))); } while (0)
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
static_assert(OFF_ONES % 1024 == 0, "UMMA operand tiles are 1024-byte aligned");

__global__ void __launch_bounds__(THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmkv,
                 __nv_bfloat16* __restrict__ out, int N, int D) {
  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the __shared__ array itself (not a round trip through uintptr_t) keeps the pointer in
  // the shared address space: LDS/STS with 32-bit addresses instead of generic LD/ST with 64-bit address math
  uint8_t* smem = smem_raw + ((1024u - (ig::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [3]
  uint64_t* v_empty = bars + 10;  // [3]
  uint64_t* s_full = bars + 13;   // [2]
  uint64_t* s_free = bars + 15;   // [2]
  uint64_t* p_full = bars + 17;   // [2]
  uint64_t* p_free = bars + 19;   // [2]
  uint64_t* o_done = bars + 21;   // [2]  PV_j (and everything before it) has completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = ig::warp_idx_uniform(), lane = threadIdx.x & 31;
  PROF_T(cta0);
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nb = (N + BKV - 1) / BKV;
  const int row0 = b * N;  // first token row of this batch element in the qkv matrix

  if (warp == 0 && lane == 0) {
    ig::tma_prefetch_desc(&tmq);
    ig::tma_prefetch_desc(&tmkv);
    ig::mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      ig::mbar_init(&k_full[s], 1);
      ig::mbar_init(&k_empty[s], 1);
      ig::mbar_init(&v_full[s], 1);
      ig::mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ig::mbar_init(&s_full[s], 1);
      ig::mbar_init(&s_free[s], 4);
      ig::mbar_init(&p_full[s], 4);
      ig::mbar_init(&p_free[s], 1);
      ig::mbar_init(&o_done[s], 1);
    }
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

  if (warp == 0) {
    if (ig::elect_one()) {
      // ===================== TMA producer =====================
      ig::mbar_expect_tx(q_full, Q_BYTES);
      ig::tma_load_2d(smem + OFF_Q, &tmq, q_full, h * HD, row0 + q0);
      for (int j = 0; j < nb; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t par = ((j / KV_STAGES) & 1) ^ 1;
        ig::mbar_wait(&k_empty[st], par);
        ig::mbar_expect_tx(&k_full[st], KV_BYTES);
        ig::tma_load_2d(smem + OFF_K + st * KV_BYTES, &tmkv, &k_full[st], D + h * HD, row0 + j * BKV);
        ig::mbar_wait(&v_empty[st], par);
        ig::mbar_expect_tx(&v_full[st], KV_BYTES);
        ig::tma_load_2d(smem + OFF_V + st * KV_BYTES, &tmkv, &v_full[st], 2 * D + h * HD, row0 + j * BKV);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====
    // Descriptors are formed ADDITIVELY from five base words computed once (start address >> 4 in the low
    // word; a stage / buffer / K-step is a constant added to it), so a UMMA costs one or two uniform-datapath
    // adds instead of a shift-mask-or chain per operand: the single issuing warp is a serial instruction stream
    // and was the slowest actor of the CTA (in-kernel clock64 profile: ~700 clk to issue PV_j, ~900 for QK_{j+2},
    // while the softmax warps waited ~1200 clk per block for S_j).
    const uint32_t idesc = ig::umma_idesc_bf16(BQ, BKV, 0, 0);    // S = Q K^T (N = 64 kv)
    const uint32_t idesc_o = ig::umma_idesc_bf16(BQ, HD, 0, 1);   // O += P V  (B = V, MN-major)
    const uint32_t idesc_l = ig::umma_idesc_bf16(BQ, 8, 0, 0);    // L += P 1  (B = ones, K-major, N = 8)
    const uint32_t smem_base = ig::smem_u32(smem);
    const uint32_t q_lo = ig::umma_desc_lo(smem_base + OFF_Q);        // K-major tiles: LBO 16, SBO 1024
    const uint32_t k_lo = ig::umma_desc_lo(smem_base + OFF_K);
    const uint32_t p_lo = ig::umma_desc_lo(smem_base + OFF_P);
    const uint32_t one_lo = ig::umma_desc_lo(smem_base + OFF_ONES);
    // V is consumed MN-major straight from its [kv, 64] tile: 16 kv rows of 128 bytes per K step, 8-row groups
    // 1024 B apart (LBO = SBO = 1024)
    const uint32_t v_lo = (((smem_base + OFF_V) & 0x3FFFF) >> 4) | ((1024u >> 4) << 16);
    auto issue_qk = [&](int i) {
      const int st = i % KV_STAGES, sb = i & 1;
      ig::mbar_wait(&k_full[st], (i / KV_STAGES) & 1);
      ig::mbar_wait(&s_free[sb], ((i >> 1) & 1) ^ 1);
      ig::tc_fence_after();
      const uint32_t dk = k_lo + st * (KV_BYTES >> 4);
      const uint32_t d_s = tmem_base + COL_S + sb * BKV;
      if (ig::elect_one()) {
#pragma unroll
        for (int k = 0; k < (ATTN_ABLATE == 7 ? 1 : HD / 16); ++k)
          ig::umma_bf16(d_s, ig::umma_desc_pack(q_lo + 2 * k), ig::umma_desc_pack(dk + 2 * k), idesc, k > 0);
        ig::umma_commit(&s_full[sb]);
        ig::umma_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    ig::mbar_wait(q_full, 0);
    issue_qk(0);
    if (nb > 1) issue_qk(1);
    for (int j = 0; j < nb; ++j) {
      const int st = j % KV_STAGES, pb = j & 1;
      // S_{j+2} = Q K_{j+2}^T goes first: it needs only the S buffer that the softmax warps hand back as soon as
      // S_j is in their registers, not P_j -- so the scores of the next two blocks are always ready ahead of the
      // softmax warps and the QK issue latency is off the P_j -> PV_j critical path.
      PROF_T(m0);
      if (j + 2 < nb) issue_qk(j + 2);
      PROF_T(m1);
      ig::mbar_wait(&v_full[st], (j / KV_STAGES) & 1);
      ig::mbar_wait(&p_full[pb], (j >> 1) & 1);  // P_j is in smem and any rescale of O / L is finished
      PROF_T(m2);
      ig::tc_fence_after();
      const uint32_t dp = p_lo + pb * (P_BYTES >> 4);
      const uint32_t dv = v_lo + st * (KV_BYTES >> 4);
      const uint32_t acc0 = j > 0 ? 1u : 0u;
      if (ig::elect_one()) {
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // A = P_j: 128 rows x 64 kv (one swizzle atom column), 32 bytes per K step
          const uint64_t da = ig::umma_desc_pack(dp + 2 * k);
          if (ATTN_ABLATE != 6 || k == 0)
            ig::umma_bf16(tmem_base + COL_O, da, ig::umma_desc_pack(dv + k * (2048 >> 4)), idesc_o, k > 0 ? 1u : acc0);
          if (ATTN_ABLATE != 5 || k == 0)
            ig::umma_bf16(tmem_base + COL_L, da, ig::umma_desc_pack(one_lo + 2 * k), idesc_l, k > 0 ? 1u : acc0);
        }
        ig::umma_commit(&o_done[pb]);
        ig::umma_commit(&v_empty[st]);
        ig::umma_commit(&p_free[pb]);
      }
      __syncwarp();
      PROF_T(m3);
      PROF_ADD(8, m0, m1);   // issue QK_{j+2} (incl. waits K, S free)
      PROF_ADD(9, m1, m2);   // wait V_j, P_j
      PROF_ADD(10, m2, m3);  // issue PV_j
    }
  } else {
    // ===================== softmax / output warps (one thread per query row) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_s = t_lane + COL_S, t_o = t_lane + COL_O, t_l = t_lane + COL_L;
    const float sl2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
    const float jump = RESCALE_LOG2 / sl2;           // the same threshold in raw score units
    const uint32_t NEG_INF = 0xff800000u;
    float m_ref = -INFINITY;
    uint32_t sc[64];
    uint32_t(&sa)[32] = *reinterpret_cast<uint32_t(*)[32]>(sc);
    uint32_t(&sb2)[32] = *reinterpret_cast<uint32_t(*)[32]>(sc + 32);

    for (int j = 0; j < nb; ++j) {
      const int kv0 = j * BKV, sb = j & 1;
      const int nvalid = min(BKV, N - kv0);  // warp-uniform
      PROF_T(c0);
      ig::mbar_wait(&s_full[sb], (j >> 1) & 1);
      PROF_T(c1);
      ig::tc_fence_after();
#if ATTN_ABLATE == 3
#pragma unroll
      for (int i = 0; i < 64; ++i) sc[i] = __float_as_uint(0.01f * (i + j + lane));
#else
      ig::tmem_ld32(t_s + sb * BKV, sa);
      ig::tmem_ld32(t_s + sb * BKV + 32, sb2);
      ig::tmem_ld_wait();
#endif
      // the scores are in registers: hand the S buffer back so QK^T of block j+2 can start
      ig::tc_fence_before();
      __syncwarp();
      if (lane == 0) ig::mbar_arrive(&s_free[sb]);
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
#if ATTN_ABLATE == 4
      const float m_blk = __uint_as_float(sc[lane & 63]);
#else
      const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
#endif
      if (j == 0) {
        m_ref = m_blk;
      } else {
        const bool need = m_blk > m_ref + jump;
        if (__any_sync(0xffffffffu, need)) {
          // ---- lazy rescale: O and L of this warp's 32 rows are multiplied by 2^(old - new) in TMEM.
          // PV_{j-1} must have completed; PV_j cannot start before this warp arrives on p_full below.
          const float m_new = need ? m_blk : m_ref;
          const float alpha = ig::ex2((m_ref - m_new) * sl2);  // 1 for rows that keep their reference
          m_ref = m_new;
          ig::mbar_wait(&o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
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
        }
      }
      const float mc = m_ref * sl2;
      // ---- exponentials -> P_j (bf16, swizzled smem); fully masked 16-column groups are written as zeros
      PROF_T(c3);
      ig::mbar_wait(&p_free[sb], ((j >> 1) & 1) ^ 1);
      PROF_T(c4);
      uint8_t* prow = smem + OFF_P + sb * P_BYTES + row * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t pk[8];
        if (g * 16 < nvalid) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float p0 = exp2_or_ablate(fmaf(__uint_as_float(sc[g * 16 + 2 * i]), sl2, -mc));
            const float p1 = exp2_or_ablate(fmaf(__uint_as_float(sc[g * 16 + 2 * i + 1]), sl2, -mc));
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
          const int chunk = g * 2 + q;  // 16-byte chunk inside the 128-byte row
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      PROF_T(c5);
      ig::fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the UMMA async proxy
      ig::tc_fence_before();         // orders a rescale's tcgen05.st before the MMA warp's PV_j
      __syncwarp();
      if (lane == 0) ig::mbar_arrive(&p_full[sb]);
      PROF_T(c6);
      if (j == 0) { PROF_ADD(6, c0, c1); PROF_ADD(12, cta0, c0); }  // first block: wait S_0; kernel start -> softmax loop
      PROF_ADD(0, c0, c1);  // wait S_j
      PROF_ADD(1, c1, c2);  // TMEM load of S_j
      PROF_ADD(2, c2, c3);  // mask + max (+ rescale)
      PROF_ADD(3, c3, c4);  // wait P buffer free
      PROF_ADD(4, c4, c5);  // exponentials + P stores
      PROF_ADD(5, c5, c6);  // fences + arrive
    }
    // ---- epilogue: O / L
    PROF_T(e0);
    ig::mbar_wait(&o_done[(nb - 1) & 1], ((nb - 1) >> 1) & 1);
    PROF_T(e1);
    PROF_ADD(7, e0, e1);  // wait for the last PV
    ig::tc_fence_after();
    ig::tmem_ld32(t_o, sa);
    ig::tmem_ld32(t_o + 32, sb2);
    const uint32_t lv = ig::tmem_ld1(t_l);
    ig::tmem_ld_wait();
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
#ifdef ATTN_PROFILE
    { PROF_T(e2); PROF_ADD(13, e1, e2); }  // O load, normalise, store
#endif
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
  cudaMemcpyFromSymbol(out16, attn::g_attn_prof, sizeof(unsigned long long) * 16);
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(attn::g_attn_prof, z, sizeof(z));
  return 0;
}
#endif

namespace ops {
int attention(const void* qkv, void* out, int B, int N, int heads, cudaStream_t st) {
  IG_REQUIRE(B >= 1 && N >= 1 && heads >= 1, IG_ESHAPE, "attention: bad shape B=%d N=%d heads=%d", B, N, heads);
  IG_REQUIRE(B <= 65535 && heads <= 65535, IG_ESHAPE, "attention: grid too large");
  const int D = heads * attn::HD;
  static bool configured = false;
  if (!configured) {
    IG_CUDA_OK(cudaFuncSetAttribute(attn::attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    attn::SMEM_TOTAL));
    configured = true;
  }
  CUtensorMap tmq, tmkv;
  IG_TRY(ig_make_tmap_bf16(&tmq, qkv, static_cast<uint64_t>(B) * N, 3 * D, 3 * D, attn::BQ, 64));
  IG_TRY(ig_make_tmap_bf16(&tmkv, qkv, static_cast<uint64_t>(B) * N, 3 * D, 3 * D, attn::BKV, 64));
  dim3 grid((N + attn::BQ - 1) / attn::BQ, heads, B);
  ig::ProfScope prof(ig::PROF_ATTENTION, st);
  attn::attention_kernel<<<grid, attn::THREADS, attn::SMEM_TOTAL, st>>>(tmq, tmkv, static_cast<__nv_bfloat16*>(out), N, D);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}
}  // namespace ops

extern "C" int ig_attention(const void* qkv, void* out, int B, int N, int heads, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(qkv && out, IG_EINVAL, "ig_attention: null pointer");
  return ops::attention(qkv, out, B, N, heads, static_cast<cudaStream_t>(stream));
}
