// Kernel 3b -- fused multi-head attention over T x 14 x 14 (+cls) tokens, head_dim 64.
//
// Replaces timm 1.0.20 Attention.forward (fused branch: F.scaled_dot_product_attention with
// scale head_dim**-0.5) as constructed by instageo/model/pritvhi.py:445-457.  Input is the
// qkv GEMM output in timm's own layout [B*N, 3*D] = (q | k | v) x (head, 64), so the
// reshape/permute of the reference costs nothing: Q, K and V tiles are 2-D TMA boxes of that
// matrix.  Output is written as [B*N, D] (head-major), the exact operand of the proj GEMM.
//
// One CTA = 128 query rows of one (batch, head).  tcgen05 throughout:
//   S = Q K^T   : UMMA 128x128x16, both operands K-major (128-byte swizzle), S in TMEM
//   O += P V    : UMMA 128x64x16, A = P (bf16, written to smem by the softmax warps),
//                 B = V used MN-major straight from its [kv, 64] tile (no transpose pass)
// Softmax: thread <-> TMEM lane <-> query row, so row max / row sum are thread-local.
// Two passes over the KV blocks: pass 0 finds the exact row max (QK^T only), pass 1
// recomputes S, exponentiates against the final max and accumulates P V in TMEM -- no
// accumulator rescaling, at the price of issuing the (cheap) QK^T MMAs twice.
// 96 KB smem + 256 TMEM columns per CTA -> 2 CTAs per SM overlap softmax with MMA.
#include "ig_ops.cuh"

namespace attn {

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16384
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;       // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;   // 1 stage
constexpr int OFF_P = OFF_V + TILE_BYTES;       // 128 x 128 bf16 = two 128x64 swizzle atoms
constexpr int OFF_BAR = OFF_P + 2 * TILE_BYTES;
constexpr int SMEM_TOTAL = 1024 + OFF_BAR + 128;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0, COL_O = 128;

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
  uint64_t* s_empty = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* p_empty = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nb = (N + BKV - 1) / BKV;
  const int row0 = b * N;  // first token row of this batch element in the qkv matrix

  if (warp == 0 && lane == 0) {
    ig::tma_prefetch_desc(&tm);
    ig::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ig::mbar_init(&k_full[s], 1);
      ig::mbar_init(&k_empty[s], 1);
    }
    ig::mbar_init(v_full, 1);
    ig::mbar_init(v_empty, 1);
    ig::mbar_init(s_full, 1);
    ig::mbar_init(s_empty, 4);
    ig::mbar_init(p_full, 4);
    ig::mbar_init(p_empty, 1);
    ig::mbar_init(o_full, 1);
    ig::fence_barrier_init();
  }
  if (warp == 1) {
    ig::tmem_alloc(tmem_ptr, TMEM_COLS);
    ig::tmem_relinquish();
  }
  ig::tc_fence_before();
  __syncthreads();
  ig::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      ig::mbar_expect_tx(q_full, TILE_BYTES);
      ig::tma_load_2d(smem + OFF_Q, &tm, q_full, h * HD, row0 + q0);
      int kit = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < nb; ++j, ++kit) {
          const int st = kit & 1;
          ig::mbar_wait(&k_empty[st], ((kit >> 1) & 1) ^ 1);
          ig::mbar_expect_tx(&k_full[st], TILE_BYTES);
          ig::tma_load_2d(smem + OFF_K + st * TILE_BYTES, &tm, &k_full[st], D + h * HD, row0 + j * BKV);
          if (pass == 1) {
            ig::mbar_wait(v_empty, (j & 1) ^ 1);
            ig::mbar_expect_tx(v_full, TILE_BYTES);
            ig::tma_load_2d(smem + OFF_V, &tm, v_full, 2 * D + h * HD, row0 + j * BKV);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc_s = ig::umma_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t idesc_o = ig::umma_idesc_bf16(BQ, HD, 0, 1);  // B = V, MN-major
      const uint32_t sq = ig::smem_u32(smem + OFF_Q);
      const uint32_t sp = ig::smem_u32(smem + OFF_P);
      const uint32_t sv = ig::smem_u32(smem + OFF_V);
      ig::mbar_wait(q_full, 0);
      int kit = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < nb; ++j, ++kit) {
          const int st = kit & 1;
          ig::mbar_wait(&k_full[st], (kit >> 1) & 1);
          ig::mbar_wait(s_empty, (kit & 1) ^ 1);
          ig::tc_fence_after();
          const uint64_t dq = ig::umma_desc_sw128(sq, 1024, 16);
          const uint64_t dk = ig::umma_desc_sw128(ig::smem_u32(smem + OFF_K + st * TILE_BYTES), 1024, 16);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            ig::umma_bf16(tmem_base + COL_S, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
          ig::umma_commit(s_full);
          ig::umma_commit(&k_empty[st]);
          if (pass == 1) {
            ig::mbar_wait(v_full, j & 1);
            ig::mbar_wait(p_full, j & 1);
            ig::tc_fence_after();
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k) {
              // A = P: atom (k / 4) of 128 rows x 64 kv, 32 bytes per K step inside the atom
              const uint64_t dp = ig::umma_desc_sw128(sp + (k >> 2) * TILE_BYTES + (k & 3) * 32, 1024, 16);
              // B = V (MN-major): 16 kv rows of 128 bytes per K step, 8-row groups 1024 B apart
              const uint64_t dv = ig::umma_desc_sw128(sv + k * 2048, 1024, 1024);
              ig::umma_bf16(tmem_base + COL_O, dp, dv, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
            }
            ig::umma_commit(v_empty);
            ig::umma_commit(p_empty);
          }
        }
      }
      ig::umma_commit(o_full);
    }
  } else {
    // ===================== softmax / output warps (one thread per query row) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + COL_S;
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + COL_O;
    const float sl2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
    int sit = 0;
    float m = -INFINITY;
    for (int j = 0; j < nb; ++j, ++sit) {
      ig::mbar_wait(s_full, sit & 1);
      ig::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        ig::tmem_ld32(t_s + c * 32, v);
        ig::tmem_ld_wait();
        const int col0 = j * BKV + c * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < N) m = fmaxf(m, __uint_as_float(v[i]));
      }
      ig::tc_fence_before();
      __syncwarp();
      if (lane == 0) ig::mbar_arrive(s_empty);
    }
    const float mc = (m == -INFINITY) ? 0.f : m * sl2;
    float l = 0.f;
    uint8_t* prow = smem + OFF_P + row * 128;
    for (int j = 0; j < nb; ++j, ++sit) {
      ig::mbar_wait(s_full, sit & 1);
      ig::tc_fence_after();
      ig::mbar_wait(p_empty, (j & 1) ^ 1);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        ig::tmem_ld32(t_s + c * 32, v);
        ig::tmem_ld_wait();
        const int col0 = j * BKV + c * 32;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = (col0 + 2 * i < N) ? ig::ex2(fmaf(__uint_as_float(v[2 * i]), sl2, -mc)) : 0.f;
          const float p1 = (col0 + 2 * i + 1 < N) ? ig::ex2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -mc)) : 0.f;
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
      }
      ig::fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the UMMA async proxy
      ig::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ig::mbar_arrive(p_full);
        ig::mbar_arrive(s_empty);
      }
    }
    ig::mbar_wait(o_full, 0);
    ig::tc_fence_after();
    const float inv = 1.f / l;
    const int qrow = q0 + row;
    __nv_bfloat16* orow = out + (static_cast<int64_t>(row0) + qrow) * D + h * HD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      ig::tmem_ld32(t_o + c * 32, v);
      ig::tmem_ld_wait();
      if (qrow < N) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = ig::pack_bf16(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
          o.y = ig::pack_bf16(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
          o.z = ig::pack_bf16(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
          o.w = ig::pack_bf16(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
          reinterpret_cast<uint4*>(orow + c * 32)[q] = o;
        }
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
