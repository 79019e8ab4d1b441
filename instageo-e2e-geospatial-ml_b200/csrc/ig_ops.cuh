// Internal launchers shared by model.cu and the C-ABI test entry points.
#pragma once
#include "ig_common.cuh"

namespace ops {
int layernorm(const float* x, const float* gamma, const float* beta, void* out, int M, int D, int mode,
              int ntok, int T, int g, int guard, cudaStream_t st);
int patchify(const float* x, void* out, int B, int C, int T, int S, cudaStream_t st);
int init_cls(float* x, const float* cls, const float* pos, int B, int ntok, int D, cudaStream_t st);
int zero_ring(void* buf, int B, int Hp, int Wp, int C, cudaStream_t st);
int cvt_bf16(const float* s, void* d, int64_t n, cudaStream_t st);
int repack_conv_weight(const float* s, void* d, int Cin, int Cout, int transposed, int permT, cudaStream_t st);
// phase-stacked ConvTranspose2d(k3, s2, p1, op1) weights: src = repacked [Cout][9*Cin] (tap-major), dst [4*Cout][4*Cin]
int stack_convt_weight(const void* src, void* dst, int Cin, int Cout, cudaStream_t st);
int bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, const float* cbias,
            float* scale, float* shift, int C, cudaStream_t st);
int repack_head1x1(const float* w, const float* b, float* wd, float* bd, int nc, int C, int ncp, cudaStream_t st);
int unpad_to_nchw(const void* buf, float* dst, int B, int Hp, int Wp, int C, int guard, int permT, cudaStream_t st);
// fused attention over qkv bf16 [B*N, 3*D] -> out bf16 [B*N, D] (attention.cu)
int attention(const void* qkv, void* out, int B, int N, int heads, cudaStream_t st);
int attention_maps(const void* qkv, int B, int N, int heads, CUtensorMap* tmq, CUtensorMap* tmkv);
int attention_planned(const CUtensorMap& tmq, const CUtensorMap& tmkv, void* out, int B, int N, int heads,
                      cudaStream_t st);
// bf16 -> f32 copy (parity taps)
int cvt_f32(const void* s, float* d, int64_t n, cudaStream_t st);
}  // namespace ops
