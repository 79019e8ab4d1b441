// Kernel 3b -- fused multi-head attention over T x 14 x 14 (+cls) tokens, head_dim 64.
//
// Replaces timm 1.0.20 Attention.forward (fused branch: F.scaled_dot_product_attention with
// scale head_dim**-0.5) as constructed by instageo/model/pritvhi.py:445-457.  Input is the
// qkv GEMM output in timm's own layout [B*N, 3*D] = (q | k | v) x (head, 64), so the
// reshape/permute of the reference costs nothing: Q, K and V tiles are 2-D TMA boxes of that
// matrix.  Output is written as [B*N, D] (head-major), the exact operand of the proj GEMM.
//
// One CTA = 128 query rows of one (batch, head); 2 CTAs per SM.  tcgen05 throughout:
//   S   = Q K_j^T : UMMA 128x128x16, both operands K-major (128-byte swizzle), S in TMEM
//   O_j = P_j V_j : UMMA 128x64x16, A = P_j (bf16, written to swizzled smem by the softmax warps),
//                   B = V_j used MN-major straight from its [kv, 64] tile (no transpose pass);
//                   NOT accumulated in TMEM: each block's product lands in its own TMEM buffer
// Softmax (4 warps, thread <-> TMEM lane <-> query row): single pass over the KV blocks with a
// running max.  Per block: read S once for the row max, read it again to exponentiate against the
// new max (TMEM reads are cheap, a second QK^T is not needed), write P_j; then fold the PREVIOUS
// block's O_{j-1} into register accumulators: acc = acc * exp(m_{j-2} - m_{j-1}) + O_{j-1}.
// Keeping the running output in registers means no TMEM read-modify-write rescale and no ordering
// hazard between the rescale and the next P V MMA: the tensor core only ever writes fresh tiles.
// TMEM loads are software-pipelined (the next 32-column chunk is in flight while the current one
// is processed).  The exp count (N^2 per head) on the 16-lane/clk MUFU is the floor of this kernel.
#include "ig_ops.cuh"

namespace attn {

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16384
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;       // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;   // 1 stage
constexpr int OFF_P = OFF_V + TILE_BYTES;       // 128 x 128 bf16 = two 128x64 swizzle atoms
constexpr int OFF_BAR = OFF_P + 2 * TILE_BYTES;
constexpr int SMEM_TOTAL = 1024 + OFF_BAR + 256;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0, COL_O = 128;           // O buffers at 128 and 192

// 32 columns of one TMEM lane quadrant -> registers (asynchronous until tmem_ld_wait)
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) { ig::tmem_ld32(taddr, v); }

__global__ void __launch_bounds__(THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* __restrict__ out, int N, int D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* s_free = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* p_free = bars + 10;
  uint64_t* o_full = bars + 11;  // [2]
  uint64_t* o_free = bars + 13;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = ig::warp_idx_uniform(), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nb = (N + BKV - 1) / BKV;
  const int row0 = b * N;  // first token row of this batch element in the qkv matrix

  if (warp == 0 && lane == 0) {
    ig::tma_prefetch_desc(&tm);
    ig::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ig::mbar_init(&k_full[s], 1);
      ig::mbar_init(&k_empty[s], 1);
      ig::mbar_init(&o_full[s], 1);
      ig::mbar_init(&o_free[s], 4);
    }
    ig::mbar_init(v_full, 1);
    ig::mbar_init(v_empty, 1);
    ig::mbar_init(s_full, 1);
    ig::mbar_init(s_free, 4);
    ig::mbar_init(p_full, 4);
    ig::mbar_init(p_free, 1);
    ig::fence_barrier_init();
  }
  if (warp == 1) {
    ig::tmem_alloc(tmem_ptr, TMEM_COLS);
    ig::tmem_relinquish();
  }
  ig::tc_fence_before();
  __syncthreads();
  ig::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    if (ig::elect_one()) {
      // ===================== TMA producer =====================
      ig::mbar_expect_tx(q_full, TILE_BYTES);
      ig::tma_load_2d(smem + OFF_Q, &tm, q_full, h * HD, row0 + q0);
      for (int j = 0; j < nb; ++j) {
        const int st = j & 1;
        ig::mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
        ig::mbar_expect_tx(&k_full[st], TILE_BYTES);
        ig::tma_load_2d(smem + OFF_K + st * TILE_BYTES, &tm, &k_full[st], D + h * HD, row0 + j * BKV);
        ig::mbar_wait(v_empty, (j & 1) ^ 1);
        ig::mbar_expect_tx(v_full, TILE_BYTES);
        ig::tma_load_2d(smem + OFF_V, &tm, v_full, 2 * D + h * HD, row0 + j * BKV);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====
    const uint32_t idesc_s = ig::umma_idesc_bf16(BQ, BKV, 0, 0);
    const uint32_t idesc_o = ig::umma_idesc_bf16(BQ, HD, 0, 1);  // B = V, MN-major
    const uint32_t sq = ig::smem_u32(smem + OFF_Q);
    const uint32_t sp = ig::smem_u32(smem + OFF_P);
    const uint32_t sv = ig::smem_u32(smem + OFF_V);
    ig::mbar_wait(q_full, 0);
    for (int j = 0; j <= nb; ++j) {
      if (j < nb) {
        const int st = j & 1;
        ig::mbar_wait(&k_full[st], (j >> 1) & 1);
        ig::mbar_wait(s_free, (j & 1) ^ 1);
        ig::tc_fence_after();
        if (ig::elect_one()) {
          const uint64_t dq = ig::umma_desc_sw128(sq, 1024, 16);
          const uint64_t dk = ig::umma_desc_sw128(ig::smem_u32(smem + OFF_K + st * TILE_BYTES), 1024, 16);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            ig::umma_bf16(tmem_base + COL_S, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
          ig::umma_commit(s_full);
          ig::umma_commit(&k_empty[st]);
        }
        __syncwarp();
      }
      if (j > 0) {
        const int jj = j - 1, ob = jj & 1;  // O_jj = P_jj V_jj into its own buffer
        ig::mbar_wait(v_full, jj & 1);
        ig::mbar_wait(p_full, jj & 1);
        ig::mbar_wait(&o_free[ob], ((jj >> 1) & 1) ^ 1);
        ig::tc_fence_after();
        if (ig::elect_one()) {
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            // A = P: atom (k / 4) of 128 rows x 64 kv, 32 bytes per K step inside the atom
            const uint64_t dp = ig::umma_desc_sw128(sp + (k >> 2) * TILE_BYTES + (k & 3) * 32, 1024, 16);
            // B = V (MN-major): 16 kv rows of 128 bytes per K step, 8-row groups 1024 B apart
            const uint64_t dv = ig::umma_desc_sw128(sv + k * 2048, 1024, 1024);
            ig::umma_bf16(tmem_base + COL_O + ob * HD, dp, dv, idesc_o, k > 0);
          }
          ig::umma_commit(&o_full[ob]);
          ig::umma_commit(v_empty);
          ig::umma_commit(p_free);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax / output warps (one thread per query row) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + COL_S;
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + COL_O;
    const float sl2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
    uint8_t* prow = smem + OFF_P + row * 128;
    float acc[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) acc[i] = 0.f;
    float m_ref = -INFINITY, l = 0.f, alpha_pend = 0.f;
    uint32_t va[32], vb[32];

    for (int j = 0; j < nb; ++j) {
      const int kv0 = j * BKV;
      const bool tail = kv0 + BKV > N;  // block has masked columns
      ig::mbar_wait(s_full, j & 1);
      ig::tc_fence_after();
      // ---- pass A: row max (next chunk in flight while the current one is reduced)
      float mb = -INFINITY;
      ld32(t_s, va);
      ig::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t(&cur)[32] = (c & 1) ? vb : va;
        uint32_t(&nxt)[32] = (c & 1) ? va : vb;
        if (c < 3) ld32(t_s + (c + 1) * 32, nxt);
        else ld32(t_s, nxt);  // first chunk of pass B
        const int col0 = kv0 + c * 32;
        if (!tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mb = fmaxf(mb, __uint_as_float(cur[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i < N) mb = fmaxf(mb, __uint_as_float(cur[i]));
        }
        ig::tmem_ld_wait();
      }
      const float m_new = fmaxf(m_ref, mb);
      const float alpha = ig::ex2((m_ref - m_new) * sl2);  // 0 on the first block
      const float mc = m_new * sl2;
      m_ref = m_new;
      l *= alpha;
      // ---- pass B: exponentials -> P (bf16, swizzled smem); chunk 0 is already in `va`
      ig::mbar_wait(p_free, (j & 1) ^ 1);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t(&cur)[32] = (c & 1) ? vb : va;
        uint32_t(&nxt)[32] = (c & 1) ? va : vb;
        if (c < 3) ld32(t_s + (c + 1) * 32, nxt);
        const int col0 = kv0 + c * 32;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ig::ex2(fmaf(__uint_as_float(cur[2 * i]), sl2, -mc));
          float p1 = ig::ex2(fmaf(__uint_as_float(cur[2 * i + 1]), sl2, -mc));
          if (tail) {
            p0 = (col0 + 2 * i < N) ? p0 : 0.f;
            p1 = (col0 + 2 * i + 1 < N) ? p1 : 0.f;
          }
          l += p0 + p1;
          pk[i] = ig::pack_bf16(p0, p1);
        }
        uint8_t* atom = prow + (c >> 1) * TILE_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c & 1) * 4 + q;  // 16-byte chunk inside the 128-byte row
          *reinterpret_cast<uint4*>(atom + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        if (c < 3) ig::tmem_ld_wait();
      }
      ig::fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the UMMA async proxy
      ig::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ig::mbar_arrive(p_full);
        ig::mbar_arrive(s_free);
      }
      // ---- fold the previous block's O into the register accumulators
      if (j > 0) {
        const int jj = j - 1, ob = jj & 1;
        ig::mbar_wait(&o_full[ob], (jj >> 1) & 1);
        ig::tc_fence_after();
        ld32(t_o + ob * HD, va);
        ld32(t_o + ob * HD + 32, vb);
        ig::tmem_ld_wait();
        ig::tc_fence_before();
        __syncwarp();
        if (lane == 0) ig::mbar_arrive(&o_free[ob]);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          acc[i] = fmaf(acc[i], alpha_pend, __uint_as_float(va[i]));
          acc[32 + i] = fmaf(acc[32 + i], alpha_pend, __uint_as_float(vb[i]));
        }
      }
      alpha_pend = alpha;
    }
    {
      const int jj = nb - 1, ob = jj & 1;
      ig::mbar_wait(&o_full[ob], (jj >> 1) & 1);
      ig::tc_fence_after();
      ld32(t_o + ob * HD, va);
      ld32(t_o + ob * HD + 32, vb);
      ig::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        acc[i] = fmaf(acc[i], alpha_pend, __uint_as_float(va[i]));
        acc[32 + i] = fmaf(acc[32 + i], alpha_pend, __uint_as_float(vb[i]));
      }
    }
    const float inv = 1.f / l;
    const int qrow = q0 + row;
    if (qrow < N) {
      __nv_bfloat16* orow = out + (static_cast<int64_t>(row0) + qrow) * D + h * HD;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 o;
        o.x = ig::pack_bf16(acc[8 * q + 0] * inv, acc[8 * q + 1] * inv);
        o.y = ig::pack_bf16(acc[8 * q + 2] * inv, acc[8 * q + 3] * inv);
        o.z = ig::pack_bf16(acc[8 * q + 4] * inv, acc[8 * q + 5] * inv);
        o.w = ig::pack_bf16(acc[8 * q + 6] * inv, acc[8 * q + 7] * inv);
        reinterpret_cast<uint4*>(orow)[q] = o;
      }
    }
  }

  ig::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ig::tc_fence_after();
    ig::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace attn

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
  CUtensorMap tm;
  IG_TRY(ig_make_tmap_bf16(&tm, qkv, static_cast<uint64_t>(B) * N, 3 * D, 3 * D, 128, 64));
  dim3 grid((N + attn::BQ - 1) / attn::BQ, heads, B);
  ig::ProfScope prof(ig::PROF_ATTENTION, st);
  attn::attention_kernel<<<grid, attn::THREADS, attn::SMEM_TOTAL, st>>>(tm, static_cast<__nv_bfloat16*>(out), N, D);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}
}  // namespace ops

extern "C" int ig_attention(const void* qkv, void* out, int B, int N, int heads, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(qkv && out, IG_EINVAL, "ig_attention: null pointer");
  return ops::attention(qkv, out, B, N, heads, static_cast<cudaStream_t>(stream));
}
