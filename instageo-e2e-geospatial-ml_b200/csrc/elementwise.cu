// Row / elementwise kernels around the tensor-core GEMMs: LayerNorm (f32 -> bf16),
// tubelet patchify of an f32 chip batch, cls-token rows, zero borders of the padded-flat
// head buffers, weight repacking and parity taps.  All are HBM/L2-bound streaming kernels:
// one warp per token row, 16-byte vector accesses, no shared memory.
//
// Reference arithmetic: nn.LayerNorm(D, eps=1e-5) inside timm Block and PrithviViT.norm
// (instageo/model/pritvhi.py:445-459, :529); PatchEmbed's Conv3d window extraction (:266);
// cls token + pos-embed row 0 (:520-522); token -> image reshape (instageo/model/model.py:405-413).
#include "ig_ops.cuh"

namespace ops {

// ------------------------------------------------------------------------------ LayerNorm
// mode 0: out[row * D + d]
// mode 1: token rows [B, 1+T*g*g, D] -> padded-flat head input [guard + B*(g+2)^2, T*D],
//         channel = t*D + d (the ConvTranspose weights are permuted to match), cls dropped.
struct LnArgs {
  const float* x;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* out;
  int M, D, mode;
  int ntok, T, g, guard;
};

__global__ void __launch_bounds__(256) layernorm_kernel(const LnArgs a) {
  ig::pdl_launch_dependents();  // programmatic dependent launch: see ig::launch
  ig::pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= a.M) return;
  const int nvec = a.D >> 7;  // float4 per lane
  const float4* xr = reinterpret_cast<const float4*>(a.x + static_cast<int64_t>(warp) * a.D);
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nvec) {
      v[i] = xr[lane + 32 * i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / a.D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nvec) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / a.D + 1e-5f);

  __nv_bfloat16* orow;
  if (a.mode == 0) {
    orow = a.out + static_cast<int64_t>(warp) * a.D;
  } else {
    const int b = warp / a.ntok, n = warp - b * a.ntok;
    if (n == 0) return;  // cls token is dropped (model.py:406)
    const int tok = n - 1, gg = a.g * a.g;
    const int t = tok / gg, p = tok - t * gg;
    const int ph = p / a.g, pw = p - ph * a.g;
    const int gp = a.g + 2;
    const int64_t row = static_cast<int64_t>(a.guard) + static_cast<int64_t>(b) * gp * gp + (ph + 1) * gp + (pw + 1);
    orow = a.out + row * (static_cast<int64_t>(a.T) * a.D) + static_cast<int64_t>(t) * a.D;
  }
  const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
  const float4* b4 = reinterpret_cast<const float4*>(a.beta);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nvec) {
      const float4 g = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
      uint2 o;
      o.x = ig::pack_bf16((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
      o.y = ig::pack_bf16((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
      reinterpret_cast<uint2*>(orow)[lane + 32 * i] = o;
    }
}

int layernorm(const float* x, const float* gamma, const float* beta, void* out, int M, int D, int mode,
              int ntok, int T, int g, int guard, cudaStream_t st) {
  IG_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, IG_ESHAPE, "layernorm: D=%d unsupported (128..1024, %%128)", D);
  if (M <= 0) return IG_OK;
  LnArgs a{x, gamma, beta, static_cast<__nv_bfloat16*>(out), M, D, mode, ntok, T, g, guard};
  const int wpb = 8;
  ig::ProfScope prof(ig::PROF_LAYERNORM, st);
  IG_CUDA_OK(ig::launch(layernorm_kernel, dim3((M + wpb - 1) / wpb), dim3(wpb * 32), 0, st, true, a));
  return IG_OK;
}

// ------------------------------------------------------------------------------ patchify
// x f32 [B, C, T, S, S] -> tubelet rows bf16 [B*T*g*g, C*256], col = c*256 + (y%16)*16 + x%16
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                       int B, int C, int T, int S) {
  const int g = S >> 4, gpr = S >> 3;
  const int64_t total = static_cast<int64_t>(B) * C * T * S * gpr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xg = static_cast<int>(i % gpr);
    int64_t r = i / gpr;
    const int y = static_cast<int>(r % S);
    r /= S;
    const int t = static_cast<int>(r % T);
    r /= T;
    const int c = static_cast<int>(r % C);
    const int b = static_cast<int>(r / C);
    const float4* src = reinterpret_cast<const float4*>(x + i * 8);
    const float4 u = __ldcs(src), w = __ldcs(src + 1);
    const int xx = xg * 8;
    const int64_t prow = (static_cast<int64_t>(b) * T + t) * g * g + (y >> 4) * g + (xx >> 4);
    uint4 q;
    q.x = ig::pack_bf16(u.x, u.y);
    q.y = ig::pack_bf16(u.z, u.w);
    q.z = ig::pack_bf16(w.x, w.y);
    q.w = ig::pack_bf16(w.z, w.w);
    *reinterpret_cast<uint4*>(out + prow * (C * 256) + c * 256 + (y & 15) * 16 + (xx & 15)) = q;
  }
}

int patchify(const float* x, void* out, int B, int C, int T, int S, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(B) * C * T * S * (S / 8);
  if (total <= 0) return IG_OK;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(ig_num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  ig::ProfScope prof(ig::PROF_MISC, st);
  patchify_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, static_cast<__nv_bfloat16*>(out), B, C, T, S);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

// x[b*ntok + 0, :] = cls + pos[0, :]
__global__ void cls_kernel(float* x, const float* cls, const float* pos, int B, int ntok, int D) {
  ig::pdl_launch_dependents();
  ig::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  x[static_cast<int64_t>(b) * ntok * D + d] = cls[d] + pos[d];
}
int init_cls(float* x, const float* cls, const float* pos, int B, int ntok, int D, cudaStream_t st) {
  ig::ProfScope prof(ig::PROF_MISC, st);
  IG_CUDA_OK(ig::launch(cls_kernel, dim3((B * D + 255) / 256), dim3(256), 0, st, true, x, cls, pos, B, ntok, D));
  return IG_OK;
}

// Zero the one-pixel border of a padded-flat NHWC buffer [B, Hp, Wp, C] (bf16), 16 bytes per thread.
__global__ void __launch_bounds__(256) ring_kernel(__nv_bfloat16* buf, int B, int Hp, int Wp, int C) {
  const int ring = 2 * Wp + 2 * (Hp - 2);
  const int cv = C >> 3;
  const int64_t total = static_cast<int64_t>(B) * ring * cv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % cv);
    const int64_t r = i / cv;
    const int k = static_cast<int>(r % ring);
    const int b = static_cast<int>(r / ring);
    int yy, xx;
    if (k < Wp) { yy = 0; xx = k; }
    else if (k < 2 * Wp) { yy = Hp - 1; xx = k - Wp; }
    else { const int j = k - 2 * Wp; yy = 1 + (j >> 1); xx = (j & 1) ? Wp - 1 : 0; }
    const int64_t row = (static_cast<int64_t>(b) * Hp + yy) * Wp + xx;
    reinterpret_cast<uint4*>(buf + row * C)[c8] = make_uint4(0, 0, 0, 0);
  }
}
int zero_ring(void* buf, int B, int Hp, int Wp, int C, cudaStream_t st) {
  IG_REQUIRE(C % 8 == 0, IG_ESHAPE, "zero_ring: C=%d not a multiple of 8", C);
  const int64_t total = static_cast<int64_t>(B) * (2 * Wp + 2 * (Hp - 2)) * (C / 8);
  if (total <= 0) return IG_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ig::ProfScope prof(ig::PROF_MISC, st);
  ring_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<__nv_bfloat16*>(buf), B, Hp, Wp, C);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

// ------------------------------------------------------------------------------ weight repack
__global__ void cvt_bf16_kernel(const float* s, __nv_bfloat16* d, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    d[i] = __float2bfloat16_rn(s[i]);
}
int cvt_bf16(const float* s, void* d, int64_t n, cudaStream_t st) {
  if (n <= 0) return IG_OK;
  cvt_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, st>>>(
      s, static_cast<__nv_bfloat16*>(d), n);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

__global__ void cvt_f32_kernel(const __nv_bfloat16* s, float* d, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    d[i] = __bfloat162float(s[i]);
}
int cvt_f32(const void* s, float* d, int64_t n, cudaStream_t st) {
  if (n <= 0) return IG_OK;
  cvt_f32_kernel<<<static_cast<unsigned>((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(s), d, n);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

// transposed = 1: src [Cin][Cout][3][3] (ConvTranspose2d); 0: src [Cout][Cin][3][3] (Conv2d)
// dst [Cout][9*Cin], k = (ky*3+kx)*Cin + perm(cin); permT > 1: cin = d*permT + t -> t*(Cin/permT) + d
__global__ void conv_w_kernel(const float* s, __nv_bfloat16* d, int Cin, int Cout, int transposed, int permT) {
  const int64_t n = static_cast<int64_t>(Cin) * Cout * 9;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % 9);
    const int64_t r = i / 9;
    int ci, co;
    if (transposed) { co = static_cast<int>(r % Cout); ci = static_cast<int>(r / Cout); }
    else { ci = static_cast<int>(r % Cin); co = static_cast<int>(r / Cin); }
    int cp = ci;
    if (permT > 1) { const int dd = ci / permT, t = ci - dd * permT; cp = t * (Cin / permT) + dd; }
    d[static_cast<int64_t>(co) * 9 * Cin + static_cast<int64_t>(tap) * Cin + cp] = __float2bfloat16_rn(s[i]);
  }
}
int repack_conv_weight(const float* s, void* d, int Cin, int Cout, int transposed, int permT, cudaStream_t st) {
  conv_w_kernel<<<1024, 256, 0, st>>>(s, static_cast<__nv_bfloat16*>(d), Cin, Cout, transposed, permT);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

// Phase-stacked transposed convolution: output pixel (2y + pa, 2x + pb) = sum over the 2 x 2 input neighbourhood
// (y + iy, x + ix) of W[ky][kx] with ky = (pa == 0 ? 1 : iy == 0 ? 2 : 0) -- used iff iy <= pa -- and the same along x
// (SURVEY.md Appendix A.5).  Row (pa*2 + pb)*Cout + co, column (iy*2 + ix)*Cin + ci; unused blocks are zero.
__global__ void stack_convt_kernel(const __nv_bfloat16* __restrict__ s, __nv_bfloat16* __restrict__ d, int Cin, int Cout) {
  const int64_t n = static_cast<int64_t>(16) * Cin * Cout;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % (4 * Cin));
    const int row = static_cast<int>(i / (4 * Cin));
    const int ph = row / Cout, co = row - ph * Cout;
    const int tp = col / Cin, ci = col - tp * Cin;
    const int pa = ph >> 1, pb = ph & 1, iy = tp >> 1, ix = tp & 1;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (iy <= pa && ix <= pb) {
      const int ky = pa == 0 ? 1 : (iy == 0 ? 2 : 0);
      const int kx = pb == 0 ? 1 : (ix == 0 ? 2 : 0);
      v = s[static_cast<int64_t>(co) * 9 * Cin + static_cast<int64_t>(ky * 3 + kx) * Cin + ci];
    }
    d[i] = v;
  }
}
int stack_convt_weight(const void* src, void* dst, int Cin, int Cout, cudaStream_t st) {
  stack_convt_kernel<<<256, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), Cin, Cout);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

// scale = gamma / sqrt(var + 1e-5); shift = (conv_bias - mean) * scale + beta   (model.py:376, eval BN)
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var,
                               const float* cbias, float* scale, float* shift, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const float s = gamma[i] / sqrtf(var[i] + 1e-5f);
  scale[i] = s;
  shift[i] = (cbias[i] - mean[i]) * s + beta[i];
}
int bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, const float* cbias,
            float* scale, float* shift, int C, cudaStream_t st) {
  bn_fold_kernel<<<(C + 255) / 256, 256, 0, st>>>(gamma, beta, mean, var, cbias, scale, shift, C);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

// w1 [nc][C] -> [C][NCP] zero padded; b1 [nc] -> [NCP]
__global__ void head1x1_kernel(const float* w, const float* b, float* wd, float* bd, int nc, int C, int ncp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * ncp) {
    const int c = i / ncp, k = i - c * ncp;
    wd[i] = k < nc ? w[k * C + c] : 0.f;
  }
  if (i < ncp) bd[i] = i < nc ? b[i] : 0.f;
}
int repack_head1x1(const float* w, const float* b, float* wd, float* bd, int nc, int C, int ncp, cudaStream_t st) {
  head1x1_kernel<<<(C * ncp + 255) / 256, 256, 0, st>>>(w, b, wd, bd, nc, C, ncp);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

// ------------------------------------------------------------------------------ taps
// padded-flat NHWC bf16 [guard + B*Hp*Wp, C] -> NCHW f32 [B, C, Hp-2, Wp-2];
// permT > 1 undoes the head-input channel permutation (stored t*D + d -> reference d*T + t).
__global__ void unpad_kernel(const __nv_bfloat16* buf, float* dst, int B, int Hp, int Wp, int C, int guard, int permT) {
  const int H = Hp - 2, W = Wp - 2;
  const int64_t n = static_cast<int64_t>(B) * C * H * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    int64_t r = i / W;
    const int y = static_cast<int>(r % H);
    r /= H;
    const int c = static_cast<int>(r % C);
    const int b = static_cast<int>(r / C);
    int cs = c;
    if (permT > 1) { const int dd = c / permT, t = c - dd * permT; cs = t * (C / permT) + dd; }
    const int64_t row = static_cast<int64_t>(guard) + (static_cast<int64_t>(b) * Hp + y + 1) * Wp + x + 1;
    dst[i] = __bfloat162float(buf[row * C + cs]);
  }
}
int unpad_to_nchw(const void* buf, float* dst, int B, int Hp, int Wp, int C, int guard, int permT, cudaStream_t st) {
  ig::ProfScope prof(ig::PROF_MISC, st);
  unpad_kernel<<<2048, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(buf), dst, B, Hp, Wp, C, guard, permT);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

}  // namespace ops

extern "C" int ig_layernorm(const float* x, const float* gamma, const float* beta, void* out, int M, int D,
                            void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(x && gamma && beta && out, IG_EINVAL, "ig_layernorm: null pointer");
  return ops::layernorm(x, gamma, beta, out, M, D, 0, 0, 0, 0, 0, static_cast<cudaStream_t>(stream));
}
