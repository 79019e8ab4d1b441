// Host-side plan of one tcgen05 GEMM / implicit-conv launch (see gemm_tc.cu).
#pragma once
#include "ig_common.cuh"

namespace gemm {

constexpr int BM = 128;      // rows per CTA = TMEM lanes
constexpr int PAIR_M = 256;  // rows per CTA pair = one cta_group::2 UMMA tile
constexpr int BK = 64;       // K per pipeline stage = one 128-byte swizzle row of bf16
constexpr int MAX_STAGES = 8; // pipeline depth = SMEM_MAIN / stage bytes, at most this
constexpr int SMEM_MAIN = 160 * 1024;  // operand ring
constexpr int MAX_BN = 256;  // UMMA N limit
constexpr int MAX_GROUPS = 3;  // A tiles per accumulation phase (vertical taps dy)
constexpr int MAX_SUB = 3;     // taps that share one A tile (horizontal taps dx = row shifts 0..2)
constexpr int A_HALO = 8;      // extra rows loaded after the 128 so shifted taps stay inside the tile
constexpr int NCP = 16;      // padded class count of the fused 1x1 head

enum Epi {
  EPI_BF16 = 0,   // out bf16 [M,N] = act(acc + bias)
  EPI_F32 = 1,    // out f32  [M,N] = acc + bias
  EPI_RESID = 2,  // out f32  [M,N] = resid + acc + bias            (transformer residual)
  EPI_PATCH = 3,  // tubelet rows -> token rows (+ bias + pos-embed), f32
  EPI_CONV = 4,   // padded-flat conv3x3: bf16 = relu(acc*scale + shift), border rows = 0
  EPI_CONVT = 5,  // padded-flat conv-transpose phase: bf16 = acc + bias scattered to (2y+a, 2x+b)
  EPI_FINAL = 6,  // conv3x3 + BN + ReLU + 1x1 conv + argmax, activations never stored
  EPI_COUNT = 7
};

// One accumulation phase = up to MAX_GROUPS shared A tiles; each is loaded ONCE per K block
// (rows a_off .. a_off + 128 + A_HALO relative to the output tile) and used by `nsub` taps that
// read it at row shift `shift[s]` (UMMA descriptor start + shift*128 B) against B columns b_off[s].
// A 3x3 convolution is 3 groups (dy) x 3 sub-taps (dx): A traffic drops 3x versus one load per tap.
struct TapGroup {
  int a_off;
  int nsub;
  int shift[MAX_SUB];
  int b_off[MAX_SUB];
};
struct Taps {
  int n;
  TapGroup g[MAX_GROUPS];
};

struct Args {
  int M, N, block_n, kc;
  int num_m_tiles, num_n_tiles, num_phases;  // m tiles are PAIR_M rows
  int a_row_base;  // guard rows in front of A (keeps shifted TMA coordinates >= 0)
  int a_box_rows;  // rows per A TMA box: 128, or 128 + A_HALO when taps are row-shifted
  int a_bytes, b_tap_bytes, stage_bytes, num_stages;  // operand ring geometry (host-computed)
  int out_tma;     // EPI_RESID with out == resid: the epilogue adds (acc + bias) INTO the f32 stream with bulk tensor
                   // reductions (cp.reduce.async.bulk.tensor .add) from its staging tiles instead of loading the residual
  int unit_bytes, units_per_stage;  // a ring slot (stage) carries up to units_per_stage (tap group, K block) units of
                                    // unit_bytes = A box + max-nsub B boxes each: one full / empty hand-off for all of them
  int rem_cols;    // 16 | 32: K per tap is not a multiple of 64 and the LAST K block is loaded with boxes of that many
                   // columns (SWIZZLE_32B / SWIZZLE_64B tiles through tmAr / tmBr) instead of a full 64-column box of which
                   // a quarter / half is used; 0: every K block is a 64-column box
  Taps taps[4];
  const float* bias;   // bias, or BatchNorm scale for EPI_CONV / EPI_FINAL
  const float* shift;  // BatchNorm shift (with conv bias folded)
  int act;             // EPI_BF16: 1 = GELU(erf)
  void* out;
  int ldo;
  const float* resid;
  int tok_per_img, ntok;  // EPI_PATCH
  const float* pos;
  int Hp, Wp;             // padded input geometry of conv modes (H+2, W+2)
  unsigned int nt_magic;   // ceil(2^32 / num_n_tiles) (num_n_tiles > 1): the tile decode's division as __umulhi
  int phase_shift;         // log2(num_phases) (1 or 4 phases)
  unsigned long long hw_magic, wp_magic;  // ceil(2^48 / (Hp*Wp)), ceil(2^48 / Wp): exact division of a row index by
                                          // multiply-shift (n * d < 2^48), set by finish_geometry
  int out_guard;          // guard rows in front of the output buffer
  int phase_a[4], phase_b[4];  // EPI_CONVT output parity of each phase
  int stack_cout;         // EPI_CONVT, phase-stacked form (> 0): N = 4 * stack_cout, column block ph = (a, b) of a row is the
                          // output pixel (2y + a, 2x + b); one phase, one N tile (see plan_conv)
  const float* w1;        // EPI_FINAL: [N][NCP] f32 (class fastest), b1 [NCP]
  const float* b1;
  int nc;
  float* logits;          // [B, nc, H, W] or null
  int8_t* argmax;         // [B, H, W] or null
  float* prob1;           // [B, H, W] or null: softmax(logits, dim=1)[:, 1] (segmentation.py:211-213)
};

struct Plan {
  CUtensorMap tmA, tmB;
  CUtensorMap tmAr, tmBr;  // narrow boxes of the last K block (copies of tmA / tmB when args.rem_cols == 0)
  CUtensorMap tmO;         // f32 output as [32-row, 32-column] tiles (args.out_tma; a copy of tmA otherwise)
  Args args;
  int epi;
};

int launch(const Plan& p, cudaStream_t stream);
// tensor maps of a plan (call after finish_geometry): A [a_rows, kc] pitch lda, B [N, b_cols] pitch b_cols
// EPI_RESID in place (x += A W^T + bias): sets resid = out = x and the output tensor map.  Falls back to the load /
// add / store epilogue (out_tma = 0) where the tile shape does not allow the bulk reduction.
int set_residual_inplace(Plan* p, float* x);
int make_maps(Plan* p, const void* A, uint64_t a_rows, uint64_t lda, const void* W, uint64_t b_cols);

// Plain linear: A [M,K] bf16 row-major (lda elements), W [N,K] bf16 row-major.
int plan_linear(Plan* p, int epi, const void* A, int64_t lda, const void* W, int M, int N, int K);
int pick_block_n(int N);
// fills a_bytes / b_tap_bytes / stage_bytes / num_stages from block_n, a_box_rows and the tap lists
void finish_geometry(Args* a);

}  // namespace gemm
