// Kernels 2-4 core: persistent, warp-specialised tcgen05 GEMM on CTA PAIRS (cta_group::2).
//
//   D[256 x block_n] (f32, 128 TMEM lanes in each CTA of the pair)
//       += A[256 x 64] (bf16, 128 rows in each CTA's smem) * B[block_n x 64]^T (half in each CTA)
//
// Why pairs: a 128 x 256 single-CTA tile needs 48 KB of operands per 64-wide K step, i.e. ~96 B/clk/SM
// or >14 KB/clk chip-wide from L2 -- more than the L2 delivers, and the measured 1-CTA kernel sat at
// ~45 % of cuBLAS for exactly that reason.  With cta_group::2 each CTA loads its own 128 A rows and only
// HALF of the B tile (the UMMA reads both CTAs' shared memories), 32 KB per step for the same math.
//
// Roles per CTA (320 threads, 1 CTA per SM, cluster = 2 CTAs = one TPC):
//   warp 0      TMA producer   : 5-stage ring of {A 16 KB, B-half <= 16 KB}, 128-byte swizzle; both CTAs'
//                                loads complete on the LEADER's full barrier
//   warp 1      MMA issuer     : leader CTA only; one lane issues tcgen05.mma.cta_group::2 (UMMA
//                                256 x block_n x 16); commits are multicast to both CTAs' barriers
//   warps 2-9   epilogue       : tcgen05.ld the CTA's 128 accumulator rows (2 warps per TMEM lane
//                                quadrant), fused epilogue in registers (thread = row), then a swizzled
//                                per-warp smem transpose so that every global access of the epilogue
//                                (bf16/f32 store, f32 residual read, pos-embed read) is a full 128-byte
//                                line per row instead of 32 scattered 16-byte pieces
// Two accumulator stages (2 x 256 TMEM columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// The MMA warp is one serial instruction stream and sets the pace of every short tile (tools/gemm_trace.py): the tile
// decode has no divisions, the accumulate flag of a UMMA is a literal or a warp-uniform value (no R2UR chain per
// UTCHMMA), the last K block of a K that is not a multiple of 64 travels as a 16- / 32-column box (SWIZZLE_32B / 64B),
// and where a unit is a single K block of a narrow tile the whole tile's units share one ring slot.
// The in-place f32 residual (x += A W^T + b) leaves as bulk tensor REDUCTIONS of the per-warp staging tiles
// (cp.reduce.async.bulk.tensor .add.f32: the L2 adds, the SM neither loads x nor stores the sum).
//
// The same kernel runs the segmentation head as an implicit GEMM with NO im2col: activations
// live in a zero-bordered "padded-flat" NHWC layout [B*(H+2)*(W+2), C], so a 3x3 tap (dy,dx)
// is the SAME 2-D matrix shifted by dy*(W+2)+dx rows -- one TMA coordinate offset per tap
// (Taps::a_off).  ConvTranspose2d(k3,s2,p1,op1) is four output-parity phases with 1/2/2/4
// taps that scatter rows to (2y+a, 2x+b).
//
// Reference arithmetic being replaced: nn.Linear / nn.Conv3d / nn.ConvTranspose2d /
// nn.Conv2d / nn.BatchNorm2d(eval) / nn.ReLU / nn.GELU / torch.argmax in
// instageo/model/pritvhi.py:243-268, :526-527 (timm Block), instageo/model/model.py:349-390,
// instageo/model/infer_utils.py:96-101.
#include "ig_gemm.cuh"

namespace gemm {

constexpr int NUM_EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int TMEM_COLS = 512;
constexpr int STG_BYTES = 32 * 128;               // per-warp transpose tile: 32 rows x 128 B
constexpr int SMEM_STAGING = NUM_EPI_WARPS * STG_BYTES;  // 32768 (EPI_FINAL: cross-half logit exchange)
constexpr int SMEM_W1 = MAX_BN * NCP * 4;         // 16384 (EPI_FINAL 1x1 weights)
constexpr int SMEM_BARS = 256;                    // barriers + tmem ptr
constexpr int PC_COLS = MAX_BN / 2;               // columns one epilogue warp touches per tile
constexpr int SMEM_PCACHE = NUM_EPI_WARPS * 2 * PC_COLS * 4;  // per-warp bias/scale + shift copies
constexpr int SMEM_TOTAL = 1024 + SMEM_MAIN + SMEM_STAGING + SMEM_W1 + SMEM_BARS + SMEM_PCACHE;
static_assert(2 * BM * NCP * 4 <= SMEM_STAGING, "logit exchange must fit the staging area");
static_assert(SMEM_TOTAL <= 232448, "over the 227 KB shared-memory limit");

struct RowInfo {
  bool valid;      // row < M
  bool interior;   // conv modes: not a border pixel
  int img, yy, xx; // conv modes: padded coordinates (EPI_PATCH: xx = token index)
  int64_t orow;    // output row (mode dependent)
};

template <int EPI>
__device__ __forceinline__ RowInfo make_row(const Args& a, int r, int phase) {
  RowInfo ri;
  ri.valid = r < a.M;
  ri.interior = false;
  ri.img = ri.yy = ri.xx = 0;
  ri.orow = r;
  if (EPI == EPI_PATCH) {
    const int b = r / a.tok_per_img, tok = r - b * a.tok_per_img;
    ri.orow = static_cast<int64_t>(b) * a.ntok + 1 + tok;
    ri.xx = tok;
  } else if (EPI == EPI_CONV || EPI == EPI_CONVT || EPI == EPI_FINAL) {
    // two divisions per row per tile, as multiply-shift: the narrow last head stages are bound by the epilogue's
    // instruction count (ncu: 12 % tensor pipe, 37 % issue slots on the T = 1 final stage), not by the UMMAs
    const int hw = a.Hp * a.Wp;
    ri.img = static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(r)) * a.hw_magic) >> 48);
    const int rem = r - ri.img * hw;
    ri.yy = static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(rem)) * a.wp_magic) >> 48);
    ri.xx = rem - ri.yy * a.Wp;
    ri.interior = ri.valid && ri.yy >= 1 && ri.yy <= a.Hp - 2 && ri.xx >= 1 && ri.xx <= a.Wp - 2;
    if (EPI == EPI_CONV) {
      ri.orow = static_cast<int64_t>(a.out_guard) + r;
    } else if (EPI == EPI_CONVT) {
      const int Hp2 = 2 * (a.Hp - 2) + 2, Wp2 = 2 * (a.Wp - 2) + 2;
      ri.orow = static_cast<int64_t>(a.out_guard) + static_cast<int64_t>(ri.img) * Hp2 * Wp2 +
                static_cast<int64_t>(2 * (ri.yy - 1) + a.phase_a[phase] + 1) * Wp2 +
                (2 * (ri.xx - 1) + a.phase_b[phase] + 1);
    }
  }
  return ri;
}

// Tile order of the persistent loop: output-parity phase fastest, then N tile, then M tile.  The CTA pairs
// that run at the same time therefore work on the SAME rows of A (all phases and N tiles of one M tile),
// which they find in L2.  With the phase outermost the transposed convolutions streamed their whole input
// once per phase: 0.95 GB of DRAM reads per launch on average (2 GB at the 112 -> 224 stage, whose input
// alone is 479 MB) instead of one pass.  The phase is rotated by the M tile so that a pair's tiles cycle
// through the 1-, 2-, 2- and 4-tap phases instead of always getting the same two.
__device__ __forceinline__ void decode_tile(const Args& a, int tile, int& phase, int& mt, int& nt) {
  // no hardware division: every role decodes every tile, and the four generic divisions this used to be (~600 clk of
  // dependent MUFU.RCP / IMAD.HI chains) sat on the MMA warp's critical path between two tiles -- a third of the
  // T = 1 final stage's tile period (tools/gemm_trace.py)
  const unsigned int pm = static_cast<unsigned int>(a.num_phases - 1);
  const unsigned int ph0 = static_cast<unsigned int>(tile) & pm;
  const unsigned int rest = static_cast<unsigned int>(tile) >> a.phase_shift;
  const unsigned int m = a.num_n_tiles == 1 ? rest : __umulhi(rest, a.nt_magic);
  mt = static_cast<int>(m);
  nt = static_cast<int>(rest - m * static_cast<unsigned int>(a.num_n_tiles));
  phase = static_cast<int>((ph0 + m) & pm);
}

// 16-byte chunk `chunk` of staging row `row`, XOR-swizzled so that both the row-per-thread writes and
// the 4-rows-per-instruction read-back are bank-conflict free.
__device__ __forceinline__ uint4* stg_slot(uint8_t* stg, int row, int chunk) {
  return reinterpret_cast<uint4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4));
}

// EPI_FINAL inner math for one 16-channel accumulator chunk of a pixel: BatchNorm(eval) + ReLU, then the 1x1 class
// convolution as f32 FMAs, classes in groups of four.  NC4 (= ceil(nc / 4)) is a COMPILE-TIME count: with a run-time
// bound inside the unrolled loop ptxas predicated the FMAs off instead of skipping them, and a 2-class head still
// issued all 16 per channel (ncu source page: 38 % of the T = 1 final stage's 990 M warp instructions were FFMA).
template <int NC4>
__device__ __forceinline__ void final_chunk(const uint32_t (&vr)[16], const float* pc0, const float* pc1, const float* w1row,
                                            float (&logit)[NCP]) {
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 sc = *reinterpret_cast<const float4*>(pc0 + 4 * j4);
    const float4 sh = *reinterpret_cast<const float4*>(pc1 + 4 * j4);
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 4 * j4 + jj;
      const float act = fmaxf(fmaf(__uint_as_float(vr[j]), scv[jj], shv[jj]), 0.f);
      const float4* w4 = reinterpret_cast<const float4*>(w1row + j * NCP);
#pragma unroll
      for (int k4 = 0; k4 < NC4; ++k4) {
        const float4 w = w4[k4];
        logit[4 * k4 + 0] = fmaf(act, w.x, logit[4 * k4 + 0]);
        logit[4 * k4 + 1] = fmaf(act, w.y, logit[4 * k4 + 1]);
        logit[4 * k4 + 2] = fmaf(act, w.z, logit[4 * k4 + 2]);
        logit[4 * k4 + 3] = fmaf(act, w.w, logit[4 * k4 + 3]);
      }
    }
  }
}

// all accumulator chunks of this warp's column share (chunks half, half + 2, ...) for one pixel row
template <int NC4>
__device__ __forceinline__ void final_row(uint32_t taddr0, int half, int nchunks, int n0, const float* pc0, const float* pc1,
                                          const float* w1s, float (&logit)[NCP]) {
  for (int ch = half; ch < nchunks; ch += 2) {
    uint32_t vr[16];
    ig::tmem_ld16(taddr0 + ch * 16, vr);
    ig::tmem_ld_wait();
    const int pci = (ch >> 1) * 16;
    final_chunk<NC4>(vr, pc0 + pci, pc1 + pci, w1s + (n0 + ch * 16) * NCP, logit);
  }
}

#ifndef IG_GEMM_ABLATE
#define IG_GEMM_ABLATE 0
#endif

// Debug build only (-DIG_GEMM_TRACE=<epi>, tools/gemm_trace.py): SM-clock timestamps of the hand-offs between the three
// warp roles of the leader CTA of pair 0, for the kernel instantiation EPI == IG_GEMM_TRACE.
#ifdef IG_GEMM_TRACE
__device__ unsigned long long g_trace[3][1 << 13];   // one region per role: plain stores, no atomics on the traced path
#define IG_TRACE_DECL unsigned int trace_i = 0
#define IG_TRACE(tag, tile)                                                                               \
  do {                                                                                                    \
    if (EPI == IG_GEMM_TRACE && blockIdx.x == 0 && trace_i < (1u << 13))                                  \
      g_trace[((tag) >> 4) - 1][trace_i++] = (static_cast<unsigned long long>(clock64()) << 24) |         \
                                             (static_cast<unsigned long long>((tile) & 0xffff) << 8) | (tag); \
  } while (0)
#else
#define IG_TRACE_DECL do { } while (0)
#define IG_TRACE(tag, tile) do { } while (0)
#endif

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmAr, const __grid_constant__ CUtensorMap tmBr,
            const __grid_constant__ CUtensorMap tmO, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the __shared__ array itself (not a round trip through uintptr_t) keeps the pointer in
  // the shared address space: LDS/STS with 32-bit addresses instead of generic LD/ST with 64-bit address math
  uint8_t* smem = smem_raw + ((1024u - (ig::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stg_base = smem + SMEM_MAIN;
  float* w1s = reinterpret_cast<float*>(smem + SMEM_MAIN + SMEM_STAGING);  // [N][NCP]
  float* exch = reinterpret_cast<float*>(stg_base);                         // [2][BM][NCP] (EPI_FINAL)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SMEM_MAIN + SMEM_STAGING + SMEM_W1);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* tfull = empty + MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* pcache = reinterpret_cast<float*>(smem + SMEM_MAIN + SMEM_STAGING + SMEM_W1 + SMEM_BARS);

  const int warp = ig::warp_idx_uniform(), lane = threadIdx.x & 31;
  IG_TRACE_DECL;
  const uint32_t rank = ig::cluster_ctarank();       // 0 = leader of the pair
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tiles_per_phase = a.num_m_tiles * a.num_n_tiles;
  const int total_tiles = tiles_per_phase * a.num_phases;
  const int kblocks_per_tap = (a.kc + BK - 1) / BK;
  const int half_n = a.block_n >> 1;

  if (warp == 0 && lane == 0) {
    ig::tma_prefetch_desc(&tmA);
    ig::tma_prefetch_desc(&tmB);
    if (a.rem_cols) {
      ig::tma_prefetch_desc(&tmAr);
      ig::tma_prefetch_desc(&tmBr);
    }
    for (int s = 0; s < MAX_STAGES; ++s) {
      ig::mbar_init(&full[s], 1);    // leader: its own expect_tx arrive; bytes from both CTAs
      ig::mbar_init(&empty[s], 1);   // one multicast commit per use
    }
    for (int s = 0; s < 2; ++s) {
      ig::mbar_init(&tfull[s], 1);
      ig::mbar_init(&tempty[s], 2 * NUM_EPI_WARPS);  // leader: epilogue warps of both CTAs
    }
    ig::fence_barrier_init();
  }
  if (warp == 1) {
    ig::tmem_alloc_cg2(tmem_ptr, TMEM_COLS);
    ig::tmem_relinquish_cg2();
  }
  if (EPI == EPI_FINAL) {
    for (int i = threadIdx.x; i < a.N * NCP; i += THREADS) w1s[i] = a.w1[i];
  }
  ig::tc_fence_before();
  __syncthreads();
  ig::cluster_sync();  // peer barriers are initialised before anyone signals them
  ig::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  // everything above overlapped the tail of the previous kernel of the stream (programmatic dependent launch);
  // nothing below may run before that kernel has completed and its writes are visible
  ig::pdl_launch_dependents();
  ig::pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (ig::elect_one()) {
      int stage = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        int phase, mt, nt;
        decode_tile(a, tile, phase, mt, nt);
        const int m0 = mt * PAIR_M + static_cast<int>(rank) * BM;
        const int nb0 = nt * a.block_n + static_cast<int>(rank) * half_n;
        const Taps& tp = a.taps[phase];
        if (a.units_per_stage > 1) {
          // the tile's (tap group t, K block kb) units in order, units_per_stage of them per ring slot
          const int units = tp.n * kblocks_per_tap;
          int t = 0, kb = 0;
          for (int u0 = 0; u0 < units; u0 += a.units_per_stage) {
            const int nu = (units - u0) < a.units_per_stage ? (units - u0) : a.units_per_stage;
            // the last K block of a K that is not a multiple of 64 travels as 16- / 32-column boxes (a full box would
            // move 4x / 2x the bytes its one or two UMMAs consume)
            uint32_t tx_bytes = 0;
            for (int j = 0, tt = t, kk = kb; j < nu; ++j) {
              const int cols = (a.rem_cols && kk == kblocks_per_tap - 1) ? a.rem_cols : BK;
#if IG_GEMM_ABLATE == 1   // timing ablation (wrong results): no B loads
              tx_bytes += 2u * (a.a_box_rows * cols * 2);
#elif IG_GEMM_ABLATE == 2  // no A loads
              tx_bytes += 2u * (tp.g[tt].nsub * half_n * cols * 2);
#else
              tx_bytes += 2u * (a.a_box_rows * cols * 2 + tp.g[tt].nsub * half_n * cols * 2);  // both CTAs' boxes
#endif
              if (++kk == kblocks_per_tap) {
                kk = 0;
                ++tt;
              }
            }
            ig::mbar_wait(&empty[stage], ph ^ 1);
            IG_TRACE(0x10, tile);   // producer: stage free, loads issued
            if (rank == 0) ig::mbar_expect_tx(&full[stage], tx_bytes);
            const uint32_t bar = ig::mapa_u32(&full[stage], 0);
            for (int j = 0; j < nu; ++j) {
              const TapGroup& tg = tp.g[t];
              const bool narrow = a.rem_cols && kb == kblocks_per_tap - 1;
              uint8_t* sa = smem + stage * a.stage_bytes + j * a.unit_bytes;
#if IG_GEMM_ABLATE != 2
              ig::tma_load_2d_cg2(sa, narrow ? &tmAr : &tmA, bar, kb * BK, a.a_row_base + m0 + tg.a_off);
#endif
#if IG_GEMM_ABLATE != 1
              for (int sb = 0; sb < tg.nsub; ++sb)
                ig::tma_load_2d_cg2(sa + a.a_bytes + sb * a.b_tap_bytes, narrow ? &tmBr : &tmB, bar, tg.b_off[sb] + kb * BK, nb0);
#endif
              if (++kb == kblocks_per_tap) {
                kb = 0;
                ++t;
              }
            }
            if (++stage == a.num_stages) {
              stage = 0;
              ph ^= 1;
            }
          }
          continue;
        }
        for (int t = 0; t < tp.n; ++t) {
          const TapGroup& tg = tp.g[t];
          for (int kb = 0; kb < kblocks_per_tap; ++kb) {
            // the last K block of a K that is not a multiple of 64 travels as 16- / 32-column boxes (a full box would
            // move 4x / 2x the bytes its one or two UMMAs consume: the ring then starves the tensor pipe)
            const bool narrow = a.rem_cols && kb == kblocks_per_tap - 1;
            const int cols = narrow ? a.rem_cols : BK;
#if IG_GEMM_ABLATE == 1   // timing ablation (wrong results): no B loads
            const uint32_t tx_bytes = 2u * (a.a_box_rows * cols * 2);
#elif IG_GEMM_ABLATE == 2  // no A loads
            const uint32_t tx_bytes = 2u * (tg.nsub * half_n * cols * 2);
#else
            const uint32_t tx_bytes = 2u * (a.a_box_rows * cols * 2 + tg.nsub * half_n * cols * 2);  // both CTAs' boxes
#endif
            ig::mbar_wait(&empty[stage], ph ^ 1);
            IG_TRACE(0x10, tile);   // producer: stage free, loads issued
            if (rank == 0) ig::mbar_expect_tx(&full[stage], tx_bytes);
            const uint32_t bar = ig::mapa_u32(&full[stage], 0);
            uint8_t* sa = smem + stage * a.stage_bytes;
#if IG_GEMM_ABLATE != 2
            ig::tma_load_2d_cg2(sa, narrow ? &tmAr : &tmA, bar, kb * BK, a.a_row_base + m0 + tg.a_off);
#endif
#if IG_GEMM_ABLATE != 1
            for (int sb = 0; sb < tg.nsub; ++sb)
              ig::tma_load_2d_cg2(sa + a.a_bytes + sb * a.b_tap_bytes, narrow ? &tmBr : &tmB, bar, tg.b_off[sb] + kb * BK, nb0);
#endif
            if (++stage == a.num_stages) {
              stage = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp walks the (uniform) loop and waits on the barriers; one elected lane issues.
    // Everything the UMMAs consume is derived from uniform values, so the descriptors live in
    // uniform registers: a single-lane loop made ptxas wrap every UTCHMMA in an ELECT / R2UR
    // "waterfall" (~25 instructions per MMA) and the ISSUE rate, not the tensor pipe or L2,
    // bounded the kernel at ~50 % tensor-pipe activity.
    if (rank == 0) {
      const uint32_t idesc = ig::umma_idesc_bf16(PAIR_M, a.block_n, 0, 0);
      const uint32_t smem_base = ig::smem_u32(smem);
      int stage = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        int phase, mt, nt;
        decode_tile(a, tile, phase, mt, nt);
        const Taps& tp = a.taps[phase];
        ig::mbar_wait(&tempty[acc], acc_ph ^ 1);
        ig::tc_fence_after();
        if (lane == 0) IG_TRACE(0x20, tile);   // MMA: accumulator slot free
        const uint32_t d_tmem = tmem_base + acc * MAX_BN;
        // 0 for the tile's first UMMA only.  Warp-uniform and never written inside the elected region: as a per-lane
        // variable it lived in a vector register and every UTCHMMA waited for an R2UR + UISETP chain of its own
        uint32_t acc_first = 0;
        if (a.units_per_stage > 1) {
          const int units = tp.n * kblocks_per_tap;
          int t = 0, kb = 0;
          for (int u0 = 0; u0 < units; u0 += a.units_per_stage) {
            const int nu = (units - u0) < a.units_per_stage ? (units - u0) : a.units_per_stage;
            ig::mbar_wait(&full[stage], ph);
            ig::tc_fence_after();
            if (lane == 0) IG_TRACE(0x21, tile);   // MMA: stage data landed
            if (ig::elect_one()) {
              int tt = t, kk = kb;   // lane-private walk; the warp's (t, kb) advance below, outside the elected region
              for (int j = 0; j < nu; ++j) {
                const uint32_t sa = smem_base + stage * a.stage_bytes + j * a.unit_bytes;
                const int krem = a.kc - kk * BK;
                const int nmma = krem >= BK ? BK / 16 : krem / 16;
                // narrow last K block: rows of rem_cols bf16 (SWIZZLE_32B / 64B tiles), same K-major descriptor otherwise
                const bool narrow = a.rem_cols && kk == kblocks_per_tap - 1;
                const uint32_t row_bytes = narrow ? 2u * a.rem_cols : 128u;
                const uint32_t hi = !narrow ? ig::UMMA_DESC_HI_SW128 : (a.rem_cols == 16 ? ig::UMMA_DESC_HI_SW32 : ig::UMMA_DESC_HI_SW64);
                const TapGroup& tg = tp.g[tt];
                auto issue_sub = [&](int sb, uint32_t acc0) {
                  // row-shifted view of the shared A tile (start inside the swizzle atom, see ig_common.cuh)
                  const uint32_t a_lo = ig::umma_desc_lo(sa + tg.shift[sb] * row_bytes);
                  const uint32_t b_lo = ig::umma_desc_lo(sa + a.a_bytes + sb * a.b_tap_bytes);
  #pragma unroll
                  for (int k = 0; k < BK / 16; ++k) {
#if IG_GEMM_ABLATE == 3   // one UMMA per unit
                    if (k == 0 && sb == 0) {
#else
                    if (k < nmma) {
#endif
                      ig::umma_bf16_cg2(d_tmem, ig::umma_desc_pack_hi(a_lo + 2 * k, hi), ig::umma_desc_pack_hi(b_lo + 2 * k, hi),
                                        idesc, k == 0 ? acc0 : 1u);
                    }
                  }
                };
                issue_sub(0, j == 0 ? acc_first : 1u);
                for (int sb = 1; sb < tg.nsub; ++sb) issue_sub(sb, 1u);
                if (++kk == kblocks_per_tap) {
                  kk = 0;
                  ++tt;
                }
              }
              ig::umma_commit_cg2(&empty[stage], 3);  // frees the stage in both CTAs
            }
            __syncwarp();
            if (lane == 0) IG_TRACE(0x23, tile);   // MMA: stage's UMMAs issued + committed
            kb += nu;
            while (kb >= kblocks_per_tap) {
              kb -= kblocks_per_tap;
              ++t;
            }
            acc_first = 1;
            if (++stage == a.num_stages) {
              stage = 0;
              ph ^= 1;
            }
          }
        } else {
          for (int t = 0; t < tp.n; ++t) {
            const int nsub = tp.g[t].nsub;
            for (int kb = 0; kb < kblocks_per_tap; ++kb) {
              ig::mbar_wait(&full[stage], ph);
              ig::tc_fence_after();
              if (lane == 0) IG_TRACE(0x21, tile);   // MMA: stage data landed
              const uint32_t sa = smem_base + stage * a.stage_bytes;
              const int krem = a.kc - kb * BK;
              const int nmma = krem >= BK ? BK / 16 : krem / 16;
              // narrow last K block: rows of rem_cols bf16 (SWIZZLE_32B / 64B tiles), same K-major descriptor otherwise
              const bool narrow = a.rem_cols && kb == kblocks_per_tap - 1;
              const uint32_t row_bytes = narrow ? 2u * a.rem_cols : 128u;
              const uint32_t hi = !narrow ? ig::UMMA_DESC_HI_SW128 : (a.rem_cols == 16 ? ig::UMMA_DESC_HI_SW32 : ig::UMMA_DESC_HI_SW64);
              if (ig::elect_one()) {
                auto issue_sub = [&](int sb, uint32_t acc0) {
                  // row-shifted view of the shared A tile (start inside the swizzle atom, see ig_common.cuh)
                  const uint32_t a_lo = ig::umma_desc_lo(sa + tp.g[t].shift[sb] * row_bytes);
                  const uint32_t b_lo = ig::umma_desc_lo(sa + a.a_bytes + sb * a.b_tap_bytes);
  #pragma unroll
                  for (int k = 0; k < BK / 16; ++k) {
#if IG_GEMM_ABLATE == 3   // one UMMA per stage
                    if (k == 0 && sb == 0) {
#else
                    if (k < nmma) {
#endif
                      ig::umma_bf16_cg2(d_tmem, ig::umma_desc_pack_hi(a_lo + 2 * k, hi), ig::umma_desc_pack_hi(b_lo + 2 * k, hi),
                                        idesc, k == 0 ? acc0 : 1u);
                    }
                  }
                };
                issue_sub(0, acc_first);
                for (int sb = 1; sb < nsub; ++sb) issue_sub(sb, 1u);
                ig::umma_commit_cg2(&empty[stage], 3);  // frees the stage in both CTAs
              }
              __syncwarp();
              acc_first = 1;
              if (lane == 0) IG_TRACE(0x23, tile);   // MMA: stage's UMMAs issued + committed
              if (++stage == a.num_stages) {
                stage = 0;
                ph ^= 1;
              }
            }
          }
        }
        if (ig::elect_one()) ig::umma_commit_cg2(&tfull[acc], 3);  // accumulator ready in both CTAs
        __syncwarp();
        if (lane == 0) IG_TRACE(0x22, tile);   // MMA: tile issued
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;      // TMEM lane quadrant this warp may read
    const int half = ew >> 2;       // which interleaved set of column groups
    uint8_t* stg = stg_base + ew * STG_BYTES;
    constexpr bool OUT_F32 = (EPI == EPI_F32 || EPI == EPI_RESID || EPI == EPI_PATCH);
    constexpr int ESZ = OUT_F32 ? 4 : 2;
    constexpr int GC = OUT_F32 ? 32 : 64;   // columns per staged group = 128 bytes per row
    constexpr int EPC = 16 / ESZ;           // elements per 16-byte chunk
    constexpr int GCX = (EPI == EPI_FINAL) ? 16 : GC;  // column-group width this warp interleaves by
    constexpr bool HAS_SHIFT = (EPI == EPI_CONV || EPI == EPI_FINAL);
    const int ngroups = (a.block_n + GC - 1) / GC;
    // Per-warp copy of the per-column epilogue parameters (bias | BN scale, BN shift) of the tile's
    // columns this warp owns.  Filled with coalesced loads BEFORE waiting for the accumulator, read
    // back as broadcast LDS.128: a global load per use sat on the critical path of every column
    // group (L1 keeps ~14 KB next to 214 KB of shared memory and the streaming stores evict it).
    float* pc0 = pcache + ew * 2 * PC_COLS;
    float* pc1 = pc0 + PC_COLS;
    const uint32_t tempty_leader[2] = {ig::mapa_u32(&tempty[0], 0), ig::mapa_u32(&tempty[1], 0)};
    int acc = 0;
    uint32_t acc_ph = 0;
    int tile_par = 0;
    auto fill_params = [&](int n0) {
      __syncwarp();
      for (int idx = lane; idx < PC_COLS; idx += 32) {
        const int c = (half + 2 * (idx / GCX)) * GCX + (idx % GCX);
        if (c < a.block_n) {
          const int cb = (EPI == EPI_CONVT && a.stack_cout) ? (n0 + c) % a.stack_cout : n0 + c;
          pc0[idx] = a.bias ? __ldg(a.bias + cb) : 0.f;
          if (HAS_SHIFT) pc1[idx] = __ldg(a.shift + n0 + c);
        }
      }
      __syncwarp();
    };
    const bool one_n_tile = a.num_n_tiles == 1;   // the tile's columns never change: fill the parameter cache once
    if (one_n_tile) fill_params(0);
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      int phase, mt, nt;
      decode_tile(a, tile, phase, mt, nt);
      const int m0 = mt * PAIR_M + static_cast<int>(rank) * BM;
      const int n0 = nt * a.block_n;
      const int row_in_tile = quad * 32 + lane;
      const int r = m0 + row_in_tile;
      const RowInfo ri = make_row<EPI>(a, r, phase);

      if (!one_n_tile) fill_params(n0);
      // element offset of this thread's output row (+ n0), -1 = row is not stored
      const bool store_row = (EPI == EPI_CONVT) ? ri.interior : ri.valid;
      const long long obase = store_row ? static_cast<long long>(ri.orow) * a.ldo + n0 : -1ll;
      // The residual tile does not depend on the MMAs.  Its pieces for the warp's NEXT column group are always
      // in flight: the first group's are requested here, before the accumulator wait (and the later groups'
      // rows are pulled into L2), group g+2's right after group g has been written out -- so their latency
      // hides under the tensor core and under the TMEM load / bias / staging of the group in between,
      // instead of being exposed once per group (4x per tile).
      float4 extra[8];
      auto load_resid = [&](int g) {
        const int col0 = g * GC;
        const int gcols = (a.block_n - col0) < GC ? (a.block_n - col0) : GC;
        const int nchunk = gcols / EPC, ch = lane & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const long long ob = __shfl_sync(0xffffffffu, obase, it * 4 + (lane >> 3));
          extra[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ob >= 0 && ch < nchunk) extra[it] = *reinterpret_cast<const float4*>(a.resid + ob + col0 + ch * EPC);
        }
      };
      // in-place residual through bulk reductions: nothing to load (see the store section)
      const bool out_tma = EPI == EPI_RESID && a.out_tma;
      if (EPI == EPI_RESID && !out_tma) {
        if (half < ngroups) load_resid(half);
        if (ri.valid) {
          const float* rrow = a.resid + static_cast<long long>(ri.orow) * a.ldo + n0;
          for (int g = half + 2; g < ngroups; g += 2)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rrow + g * GC));
        }
      }
      if (ew == 0 && lane == 0) IG_TRACE(0x30, tile);   // epilogue warp 0: ready for the accumulator
      ig::mbar_wait(&tfull[acc], acc_ph);
      ig::tc_fence_after();
      if (ew == 0 && lane == 0) IG_TRACE(0x31, tile);   // accumulator complete
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * MAX_BN;

      if (EPI != EPI_FINAL) {
        bool released = false;
        for (int g = half; g < ngroups; g += 2) {
          const int col0 = g * GC;
          const int gcols = (a.block_n - col0) < GC ? (a.block_n - col0) : GC;
          uint32_t vr[GC];
#pragma unroll
          for (int u = 0; u < GC / 16; ++u)
            if (u * 16 < gcols) ig::tmem_ld16p(taddr0 + col0 + u * 16, vr + u * 16);
          ig::tmem_ld_wait();
          if (g + 2 >= ngroups) {  // last TMEM read of this warp: hand the accumulator stage back early
            ig::tc_fence_before();
            __syncwarp();
            if (lane == 0) ig::mbar_arrive_cluster(tempty_leader[acc]);
            released = true;
          }
          if (out_tma) {   // the previous group's reduction must have read the staging tile before it is rewritten
            if (lane == 0) ig::bulk_wait_read_all();
            __syncwarp();
          }
#pragma unroll
          for (int u = 0; u < GC / 16; ++u) {
            if (u * 16 < gcols) {
              float v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(vr[u * 16 + j]);
              const int pci = (g >> 1) * GC + u * 16;  // this unit inside the warp's parameter cache
              if (EPI == EPI_CONV) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 sc = *reinterpret_cast<const float4*>(pc0 + pci + 4 * j4);
                  const float4 sh = *reinterpret_cast<const float4*>(pc1 + pci + 4 * j4);
                  v[4 * j4 + 0] = fmaxf(fmaf(v[4 * j4 + 0], sc.x, sh.x), 0.f);
                  v[4 * j4 + 1] = fmaxf(fmaf(v[4 * j4 + 1], sc.y, sh.y), 0.f);
                  v[4 * j4 + 2] = fmaxf(fmaf(v[4 * j4 + 2], sc.z, sh.z), 0.f);
                  v[4 * j4 + 3] = fmaxf(fmaf(v[4 * j4 + 3], sc.w, sh.w), 0.f);
                }
                if (!ri.interior) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
              } else {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 b = *reinterpret_cast<const float4*>(pc0 + pci + 4 * j4);
                  v[4 * j4 + 0] += b.x;
                  v[4 * j4 + 1] += b.y;
                  v[4 * j4 + 2] += b.z;
                  v[4 * j4 + 3] += b.w;
                }
              }
              if (EPI == EPI_BF16 && a.act) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = ig::gelu_tanh3(v[j]);
              }
              if (OUT_F32) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4)
                  *stg_slot(stg, lane, 4 * u + j4) =
                      make_uint4(__float_as_uint(v[4 * j4]), __float_as_uint(v[4 * j4 + 1]),
                                 __float_as_uint(v[4 * j4 + 2]), __float_as_uint(v[4 * j4 + 3]));
              } else {
                *stg_slot(stg, lane, 2 * u) =
                    make_uint4(ig::pack_bf16(v[0], v[1]), ig::pack_bf16(v[2], v[3]),
                               ig::pack_bf16(v[4], v[5]), ig::pack_bf16(v[6], v[7]));
                *stg_slot(stg, lane, 2 * u + 1) =
                    make_uint4(ig::pack_bf16(v[8], v[9]), ig::pack_bf16(v[10], v[11]),
                               ig::pack_bf16(v[12], v[13]), ig::pack_bf16(v[14], v[15]));
              }
            }
          }
          if (out_tma) {
            // x += staging tile [32 rows x 32 f32] (the XOR layout of stg_slot IS the 128-byte TMA swizzle): one bulk
            // tensor reduction per column group.  The L2 performs the adds; the SM neither loads the residual (the
            // old epilogue moved 256 KB per tile through registers and took 4x the K = 768 projection's UMMA time,
            // ncu: 32 % tensor pipe) nor stores the sum.  Each element is touched by exactly one reduction per launch,
            // so the result is order-independent; rows >= M are clipped by the tensor map.
            ig::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              ig::tma_reduce_add_2d(&tmO, stg, n0 + col0, m0 + quad * 32);
              ig::bulk_commit_group();
            }
            continue;
          }
          __syncwarp();
          // coalesced write-out: 8 lanes cover one 128-byte row piece, 4 rows per instruction.
          // All residual / pos-embed loads are issued before the first store: `out` may alias
          // `resid`, so the compiler cannot hoist a later load above an earlier store itself, and
          // a load -> add -> store chain per row would expose one DRAM latency per iteration.
          const int nchunk = gcols / EPC;
          const int ch = lane & 7;
          // phase-stacked transposed convolution: this lane's 16-byte chunk (8 channels of ONE output-parity
          // phase, stack_cout % 8 == 0) goes to output pixel (2y + a, 2x + b), a row of stack_cout channels
          long long stack_off = 0;
          if (EPI == EPI_CONVT && a.stack_cout) {
            const int c = col0 + ch * EPC;
            const int ph = c / a.stack_cout;
            const int Wp2 = 2 * (a.Wp - 2) + 2;
            stack_off = static_cast<long long>((ph >> 1) * Wp2 + (ph & 1)) * a.stack_cout + (c - ph * a.stack_cout) - c;
          }
          long long offs[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = it * 4 + (lane >> 3);
            const long long ob = __shfl_sync(0xffffffffu, obase, row);
            const int tok = (EPI == EPI_PATCH) ? __shfl_sync(0xffffffffu, ri.xx, row) : 0;
            const bool ok = ob >= 0 && ch < nchunk;
            offs[it] = ok ? ob + col0 + ch * EPC + stack_off : -1ll;
            if (EPI == EPI_PATCH) {
              extra[it] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ok)
                extra[it] = __ldg(reinterpret_cast<const float4*>(
                    a.pos + static_cast<int64_t>(1 + tok) * a.N + n0 + col0 + ch * EPC));
            }
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = it * 4 + (lane >> 3);
            if (offs[it] >= 0) {
              uint4 q = *stg_slot(stg, row, ch);
              if (EPI == EPI_RESID || EPI == EPI_PATCH) {
                q.x = __float_as_uint(__uint_as_float(q.x) + extra[it].x);
                q.y = __float_as_uint(__uint_as_float(q.y) + extra[it].y);
                q.z = __float_as_uint(__uint_as_float(q.z) + extra[it].z);
                q.w = __float_as_uint(__uint_as_float(q.w) + extra[it].w);
              }
              if (OUT_F32)
                *reinterpret_cast<uint4*>(static_cast<float*>(a.out) + offs[it]) = q;
              else
                *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.out) + offs[it]) = q;
            }
          }
          if (EPI == EPI_RESID && !out_tma && g + 2 < ngroups) load_resid(g + 2);  // other columns than the stores above
          __syncwarp();
        }
        if (!released) {  // warp had no column group in this tile (ngroups == 1, half == 1)
          ig::tc_fence_before();
          __syncwarp();
          if (lane == 0) ig::mbar_arrive_cluster(tempty_leader[acc]);
        }
      } else {
        // ---- EPI_FINAL: conv3x3 + BN + ReLU, then the 1x1 class conv and argmax in registers
        float logit[NCP];
#pragma unroll
        for (int k = 0; k < NCP; ++k) logit[k] = 0.f;
        // classes in groups of four: a 2-class (flood) or 1-channel (regression) head does a quarter of the 1x1
        // convolution's FMAs and weight loads of the padded NCP = 16 (this loop, not the UMMAs, bounded the T = 1
        // head: ~5400 clk per 256-pixel tile against ~650 clk of tensor work)
#ifdef IG_FINAL_ABLATE   // timing ablation (wrong logits): how fast is the stage with a quarter / none of the class FMAs?
        const int nc4 = IG_FINAL_ABLATE;
#else
        const int nc4 = (a.nc + 3) >> 2;
#endif
        const int nchunks = a.block_n / 16;
        switch (nc4) {   // kernel-uniform: one dispatch per tile
          case 0: final_row<0>(taddr0, half, nchunks, n0, pc0, pc1, w1s, logit); break;
          case 1: final_row<1>(taddr0, half, nchunks, n0, pc0, pc1, w1s, logit); break;
          case 2: final_row<2>(taddr0, half, nchunks, n0, pc0, pc1, w1s, logit); break;
          case 3: final_row<3>(taddr0, half, nchunks, n0, pc0, pc1, w1s, logit); break;
          default: final_row<4>(taddr0, half, nchunks, n0, pc0, pc1, w1s, logit); break;
        }
        // accumulator fully read -> hand the TMEM stage back to the MMA warp
        ig::tc_fence_before();
        __syncwarp();
        if (lane == 0) ig::mbar_arrive_cluster(tempty_leader[acc]);
        if (ew == 0 && lane == 0) IG_TRACE(0x32, tile);   // slot handed back

        float* ex = exch + (tile_par * BM + row_in_tile) * NCP;
        if (half == 1) {
#pragma unroll
          for (int k4 = 0; k4 < NCP / 4; ++k4)
            if (k4 < nc4)
              reinterpret_cast<float4*>(ex)[k4] =
                  make_float4(logit[4 * k4], logit[4 * k4 + 1], logit[4 * k4 + 2], logit[4 * k4 + 3]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
        if (half == 0 && ri.interior) {
          const int H = a.Hp - 2, W = a.Wp - 2;
          const int64_t px = static_cast<int64_t>(ri.yy - 1) * W + (ri.xx - 1);
          float best = 0.f;
          int bi = 0;
#pragma unroll
          for (int k = 0; k < NCP; ++k) {
            if (k < a.nc) {
              const float lv = logit[k] + ex[k] + __ldg(a.b1 + k);
              logit[k] = lv;
              if (a.logits) a.logits[(static_cast<int64_t>(ri.img) * a.nc + k) * H * W + px] = lv;
              if (k == 0 || lv > best) {
                best = lv;
                bi = k;
              }
            }
          }
          if (a.argmax) a.argmax[static_cast<int64_t>(ri.img) * H * W + px] = static_cast<int8_t>(bi);
          if (a.prob1) {  // predict_step: softmax over classes, channel 1 (float32, exp(x - max) / sum)
            float den = 0.f;
#pragma unroll
            for (int k = 0; k < NCP; ++k)
              if (k < a.nc) den += expf(logit[k] - best);
            a.prob1[static_cast<int64_t>(ri.img) * H * W + px] = __fdiv_rn(expf(logit[1] - best), den);
          }
        }
        tile_par ^= 1;
      }
      if (ew == 0 && lane == 0) IG_TRACE(0x33, tile);   // tile stored
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
  }

  if (EPI == EPI_RESID && a.out_tma && warp >= 2 && lane == 0) ig::bulk_wait_read_all();  // staging tiles still being read
  // No CTA of the pair may exit (or free TMEM) while its peer can still read its shared memory,
  // signal its barriers or write its accumulators.
  ig::tc_fence_before();
  __syncthreads();
  ig::cluster_sync();
  if (warp == 1) {
    ig::tc_fence_after();
    ig::tmem_dealloc_cg2(tmem_base, TMEM_COLS);
  }
}

template <int EPI>
static int launch_epi(const Plan& p, cudaStream_t stream) {
  static IgPerDevice configured = {};
  if (!configured.get()) {
    IG_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    SMEM_TOTAL));
    configured.set(1);
  }
  const int total = p.args.num_m_tiles * p.args.num_n_tiles * p.args.num_phases;
  if (total <= 0) return IG_OK;
  const int max_pairs = ig_num_sms() / 2;
  const int pairs = total < max_pairs ? total : max_pairs;
  ig::ProfScope prof((EPI == EPI_CONV || EPI == EPI_CONVT || EPI == EPI_FINAL) ? ig::PROF_GEMM_CONV : ig::PROF_GEMM_LINEAR, stream);
  IG_CUDA_OK(ig::launch(gemm_kernel<EPI>, dim3(2 * pairs), dim3(THREADS), SMEM_TOTAL, stream, true, p.tmA, p.tmB, p.tmAr,
                        p.tmBr, p.tmO, p.args));
  return IG_OK;
}

int launch(const Plan& p, cudaStream_t stream) {
  const Args& a = p.args;
  IG_REQUIRE(a.block_n >= 16 && a.block_n <= MAX_BN && a.block_n % 16 == 0, IG_ESHAPE,
             "gemm: block_n=%d must be a multiple of 16 in [16,256]", a.block_n);
  IG_REQUIRE(a.kc >= 16 && a.kc % 16 == 0, IG_ESHAPE, "gemm: K per tap %d must be a multiple of 16", a.kc);
  IG_REQUIRE(a.N % a.block_n == 0, IG_ESHAPE, "gemm: N=%d not a multiple of block_n=%d", a.N, a.block_n);
  IG_REQUIRE(a.num_stages >= 2 && a.num_stages * a.stage_bytes <= SMEM_MAIN, IG_ESHAPE,
             "gemm: operand ring does not fit (stage %d B x %d)", a.stage_bytes, a.num_stages);
  switch (p.epi) {
    case EPI_BF16: return launch_epi<EPI_BF16>(p, stream);
    case EPI_F32: return launch_epi<EPI_F32>(p, stream);
    case EPI_RESID: return launch_epi<EPI_RESID>(p, stream);
    case EPI_PATCH: return launch_epi<EPI_PATCH>(p, stream);
    case EPI_CONV: return launch_epi<EPI_CONV>(p, stream);
    case EPI_CONVT:
      IG_REQUIRE(a.stack_cout == 0 || (a.num_n_tiles == 1 && a.num_phases == 1 && a.stack_cout % 8 == 0 &&
                                       a.block_n == 4 * a.stack_cout),
                 IG_ESHAPE, "gemm: phase-stacked transposed convolution needs one N tile of 4 x %d columns", a.stack_cout);
      return launch_epi<EPI_CONVT>(p, stream);
    case EPI_FINAL:
      IG_REQUIRE(a.num_n_tiles == 1 && a.nc <= NCP, IG_ESHAPE, "gemm: fused head needs one N tile and nc <= %d", NCP);
      return launch_epi<EPI_FINAL>(p, stream);
    default: break;
  }
  ig_set_error("gemm: unknown epilogue %d", p.epi);
  return IG_EINVAL;
}

void finish_geometry(Args* a) {
  int maxsub = 1;
  for (int ph = 0; ph < a->num_phases; ++ph)
    for (int t = 0; t < a->taps[ph].n; ++t)
      if (a->taps[ph].g[t].nsub > maxsub) maxsub = a->taps[ph].g[t].nsub;
  if (a->Hp > 0 && a->Wp > 0) {
    const unsigned long long one = 1ull << 48;
    const unsigned long long hw = static_cast<unsigned long long>(a->Hp) * a->Wp;
    a->hw_magic = (one + hw - 1) / hw;
    a->wp_magic = (one + a->Wp - 1) / a->Wp;
  }
  a->phase_shift = a->num_phases == 4 ? 2 : 0;
  a->nt_magic = a->num_n_tiles > 1 ? static_cast<unsigned int>(((1ull << 32) + a->num_n_tiles - 1) / a->num_n_tiles) : 0u;
  const int rem = a->kc % BK;
  a->rem_cols = ((rem == 16 || rem == 32) && getenv("IG_NO_NARROW_K") == nullptr) ? rem : 0;
  a->a_bytes = (a->a_box_rows * BK * 2 + 1023) / 1024 * 1024;
  a->b_tap_bytes = ((a->block_n / 2) * BK * 2 + 1023) / 1024 * 1024;
  a->unit_bytes = a->a_bytes + maxsub * a->b_tap_bytes;
  // Units per ring slot.  Each slot costs the MMA warp one full-barrier wait, one commit and ~400 clk of its own
  // instruction latency (tools/gemm_trace.py), during which the tensor pipe drains its short queue.  Where a unit is a
  // single K block of a narrow tile (the T = 1 final stage: K = 48, N = 48 -- nine ~60-clk UMMAs per unit) the whole
  // tile's units go into one slot.  Everything else keeps one unit per slot: the generic packed loop costs the wide
  // GEMMs ~8 % (same-box A/B), and packing gained nothing on stages with two or more K blocks per tap.
  int max_units = 1;
  for (int ph = 0; ph < a->num_phases; ++ph)
    if (a->taps[ph].n > max_units) max_units = a->taps[ph].n;
  a->units_per_stage = 1;
  if (a->block_n <= 64 && a->kc <= BK) {
    int u = SMEM_MAIN / (2 * a->unit_bytes);
    if (u > max_units) u = max_units;
    if (u > 1) a->units_per_stage = u;
  }
  if (const char* e = getenv("IG_GEMM_UNITS")) {  // measurement aid
    const int u = atoi(e);
    if (u >= 1 && SMEM_MAIN / (u * a->unit_bytes) >= 2) a->units_per_stage = u;
  }
  a->stage_bytes = a->units_per_stage * a->unit_bytes;
  a->num_stages = SMEM_MAIN / a->stage_bytes;
  if (a->num_stages > MAX_STAGES) a->num_stages = MAX_STAGES;
  if (const char* e = getenv("IG_GEMM_STAGES")) {  // measurement aid: cap the operand ring depth
    const int cap = atoi(e);
    if (cap >= 2 && cap < a->num_stages) a->num_stages = cap;
  }
}

int make_maps(Plan* p, const void* A, uint64_t a_rows, uint64_t lda, const void* W, uint64_t b_cols) {
  const Args& a = p->args;
  IG_TRY(ig_make_tmap_bf16(&p->tmA, A, a_rows, a.kc, lda, a.a_box_rows, BK));
  IG_TRY(ig_make_tmap_bf16(&p->tmB, W, a.N, b_cols, b_cols, a.block_n / 2, BK));
  p->tmAr = p->tmA;
  p->tmBr = p->tmB;
  p->tmO = p->tmA;
  if (a.rem_cols) {
    IG_TRY(ig_make_tmap_bf16(&p->tmAr, A, a_rows, a.kc, lda, a.a_box_rows, a.rem_cols));
    IG_TRY(ig_make_tmap_bf16(&p->tmBr, W, a.N, b_cols, b_cols, a.block_n / 2, a.rem_cols));
  }
  return IG_OK;
}

#ifdef IG_GEMM_TRACE
}  // namespace gemm
extern "C" int ig_debug_gemm_trace(unsigned long long* out, int cap, int reset) {
  // out: [3][1 << 13] (zero entries = unused); reset clears the device buffer
  const size_t bytes = sizeof(unsigned long long) * 3 * (1 << 13);
  if (out && cap >= 3 * (1 << 13) && cudaMemcpyFromSymbol(out, gemm::g_trace, bytes) != cudaSuccess) return -1;
  if (reset) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, gemm::g_trace) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess) return -1;
  }
  return 3 * (1 << 13);
}
namespace gemm {
#endif

int set_residual_inplace(Plan* p, float* x) {
  Args& a = p->args;
  IG_REQUIRE(p->epi == EPI_RESID && x != nullptr, IG_EINVAL, "gemm: in-place residual needs an EPI_RESID plan");
  a.resid = x;
  a.out = x;
  a.out_tma = 0;
  if (a.block_n % 32 == 0 && a.N % 32 == 0 && a.ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      getenv("IG_NO_RESID_TMA") == nullptr) {
    IG_TRY(ig_make_tmap_f32_tile(&p->tmO, x, static_cast<uint64_t>(a.M), static_cast<uint64_t>(a.N),
                                 static_cast<uint64_t>(a.ldo), 32));
    a.out_tma = 1;
  }
  return IG_OK;
}

int pick_block_n(int N) {
  // largest multiple of 16 that divides N and fits one UMMA (<= 256)
  for (int bn = MAX_BN; bn >= 16; bn -= 16)
    if (N % bn == 0) return bn;
  return 0;
}

int plan_linear(Plan* p, int epi, const void* A, int64_t lda, const void* W, int M, int N, int K) {
  IG_REQUIRE(M >= 1 && N >= 16 && K >= 16 && K % 16 == 0 && N % 16 == 0, IG_ESHAPE,
             "linear: unsupported shape M=%d N=%d K=%d (N, K multiples of 16)", M, N, K);
  IG_REQUIRE(lda % 8 == 0 && K % 8 == 0, IG_ESHAPE, "linear: row pitch must be a multiple of 8 elements");
  Args& a = p->args;
  a = Args{};
  a.M = M;
  a.N = N;
  a.block_n = pick_block_n(N);
  a.kc = K;
  a.num_m_tiles = (M + PAIR_M - 1) / PAIR_M;
  a.num_n_tiles = N / a.block_n;
  a.num_phases = 1;
  a.a_row_base = 0;
  a.a_box_rows = BM;
  a.taps[0].n = 1;
  a.taps[0].g[0].a_off = 0;
  a.taps[0].g[0].nsub = 1;
  a.taps[0].g[0].shift[0] = 0;
  a.taps[0].g[0].b_off[0] = 0;
  a.ldo = N;
  finish_geometry(&a);
  p->epi = epi;
  IG_TRY(make_maps(p, A, M, lda, W, K));
  return IG_OK;
}

}  // namespace gemm

extern "C" int ig_linear(const void* A, const void* W, const float* bias, const float* resid,
                         void* out, int out_dtype, int M, int N, int K, int act, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(A && W && out, IG_EINVAL, "ig_linear: null pointer");
  IG_REQUIRE(out_dtype == IG_BF16 || out_dtype == IG_F32, IG_EINVAL, "ig_linear: bad out_dtype");
  IG_REQUIRE(!(resid && out_dtype != IG_F32), IG_EINVAL, "ig_linear: residual needs f32 output");
  IG_REQUIRE(!(act && out_dtype != IG_BF16), IG_EINVAL, "ig_linear: activation needs bf16 output");
  gemm::Plan p;
  const int epi = out_dtype == IG_BF16 ? gemm::EPI_BF16 : (resid ? gemm::EPI_RESID : gemm::EPI_F32);
  IG_TRY(gemm::plan_linear(&p, epi, A, K, W, M, N, K));
  p.args.bias = bias;
  p.args.resid = resid;
  p.args.act = act;
  p.args.out = out;
  if (resid && resid == out) IG_TRY(gemm::set_residual_inplace(&p, static_cast<float*>(out)));
  return gemm::launch(p, static_cast<cudaStream_t>(stream));
}
