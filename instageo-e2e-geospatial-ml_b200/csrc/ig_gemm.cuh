// Host-side plan of one tcgen05 GEMM / implicit-conv launch (see gemm_tc.cu).
#pragma once
#include "ig_common.cuh"

namespace gemm {

constexpr int BM = 128;      // rows per CTA = TMEM lanes
constexpr int PAIR_M = 256;  // rows per CTA pair = one cta_group::2 UMMA tile
constexpr int BK = 64;       // K per pipeline stage = one 128-byte swizzle row of bf16
constexpr int STAGES = 5;
constexpr int MAX_BN = 256;  // UMMA N limit
constexpr int MAX_TAPS = 9;
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

// One accumulation phase = a list of (A row shift, B column offset) taps, each `kc` wide.
struct Taps {
  int n;
  int a_off[MAX_TAPS];
  int b_off[MAX_TAPS];
};

struct Args {
  int M, N, block_n, kc;
  int num_m_tiles, num_n_tiles, num_phases;  // m tiles are PAIR_M rows
  int a_row_base;  // guard rows in front of A (keeps shifted TMA coordinates >= 0)
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
  int out_guard;          // guard rows in front of the output buffer
  int phase_a[4], phase_b[4];  // EPI_CONVT output parity of each phase
  const float* w1;        // EPI_FINAL: [N][NCP] f32 (class fastest), b1 [NCP]
  const float* b1;
  int nc;
  float* logits;          // [B, nc, H, W] or null
  int8_t* argmax;         // [B, H, W] or null
};

struct Plan {
  CUtensorMap tmA, tmB;
  Args args;
  int epi;
};

int launch(const Plan& p, cudaStream_t stream);

// Plain linear: A [M,K] bf16 row-major (lda elements), W [N,K] bf16 row-major.
int plan_linear(Plan* p, int epi, const void* A, int64_t lda, const void* W, int M, int N, int K);
int pick_block_n(int N);

}  // namespace gemm
