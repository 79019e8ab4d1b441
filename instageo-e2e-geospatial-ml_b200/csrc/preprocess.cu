// Kernel 1 -- fused raster -> chip preprocessing (HBM-bound, bit-exact integer/f32 work).
//
// Reference behaviour restated (paths relative to the reference repo):
//   band gather              instageo/model/dataloader.py:700-703
//   raw * constant_multiplier (float64 product)            :741
//   nodata mask AFTER the multiply                          :899
//   float64 -> float32 (PIL mode "F"), Normalize = f32 subtract then f32 TRUE divide,
//   [T*C,H,W] -> [C,T,H,W]                                  :515-521
//   window crops in row-major order                         :655-664
//   Fmask bit decode + each/any masking     instageo/data/hls_utils.py:77-86,
//                                           instageo/data/data_pipeline.py:229-267
//
// One thread owns 8 consecutive pixels of one window row and walks all T*C bands, so the
// per-pixel "any band" mask needs no cross-thread reduction and every global access is a
// 16-byte vector (aligned case) that a warp coalesces into full 128-byte lines.
// Algorithmic bytes per element: 2 (int16 in) + 4 (f32 out) + 1 (mask) [parity mode], or
// 2 + 2 (bf16 tubelet row) + 1/TC (pixel mask) [production mode].
#include "ig_common.cuh"

namespace {

constexpr int VEC = 8;
constexpr int MAX_TC = 64;

struct PreArgs {
  const void* raw;
  int64_t img_stride, band_stride, row_stride;
  const int32_t* band_idx;
  const int32_t* win_yx;
  int n_win, win, T, C, H, W;
  double cm, nodata;
  int has_nodata, cm_is_one;
  const float* mean;
  const float* std;
  const uint8_t* fmask;
  uint32_t fmask_bits;
  int mask_any, fill_raw;
  float* out_f32;
  __nv_bfloat16* out_patch;
  uint8_t* mask_elem;
  uint8_t* mask_px;
};

template <typename RawT>
__device__ __forceinline__ void load8(const RawT* p, RawT (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = __ldg(p + i);
}

template <typename RawT>
__device__ __forceinline__ void load8(const RawT* p, int (&v)[VEC]) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if ((a & 15) == 0) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = static_cast<int>(static_cast<RawT>(w[i] & 0xffffu));
      v[2 * i + 1] = static_cast<int>(static_cast<RawT>(w[i] >> 16));
    }
  } else if ((a & 7) == 0) {
    const uint2 q0 = __ldg(reinterpret_cast<const uint2*>(p));
    const uint2 q1 = __ldg(reinterpret_cast<const uint2*>(p) + 1);
    const uint32_t w[4] = {q0.x, q0.y, q1.x, q1.y};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = static_cast<int>(static_cast<RawT>(w[i] & 0xffffu));
      v[2 * i + 1] = static_cast<int>(static_cast<RawT>(w[i] >> 16));
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = static_cast<int>(__ldg(p + i));
  }
}

// Fmask bit test by the reference's floor-division arithmetic (hls_utils.py:84-86):
// q = v // 2^p ; bit = q - (q // 2) * 2.  For non-negative v this is (v >> p) & 1.
__device__ __forceinline__ bool fmask_hit(uint32_t v, uint32_t bits) {
  bool hit = false;
#pragma unroll
  for (int p = 1; p < 8; ++p)  // position 0 is skipped by the reference's `if pos:` test
    if (bits & (1u << p)) {
      const uint32_t q = v / (1u << p);
      hit |= (q - (q / 2u) * 2u) != 0u;
    }
  return hit;
}

template <typename RawT> struct RawVal { typedef int type; };
template <> struct RawVal<float> { typedef float type; };
template <> struct RawVal<double> { typedef double type; };

template <typename RawT>
__global__ void __launch_bounds__(256) preprocess_kernel(const PreArgs a) {
  const int groups_per_row = a.win / VEC;
  const int64_t total = static_cast<int64_t>(a.n_win) * a.win * groups_per_row;
  const int TC = a.T * a.C;
  const int gp = a.win / 16;  // tubelet grid side
  const RawT* raw = static_cast<const RawT*>(a.raw);

  for (int64_t item = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; item < total;
       item += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(item % groups_per_row);
    const int64_t r = item / groups_per_row;
    const int y = static_cast<int>(r % a.win);
    const int w = static_cast<int>(r / a.win);
    int img = w, top = 0, left = 0;
    if (a.win_yx) {
      img = __ldg(a.win_yx + 3 * w);
      top = __ldg(a.win_yx + 3 * w + 1);
      left = __ldg(a.win_yx + 3 * w + 2);
    }
    const int x = g * VEC;
    const RawT* src0 = raw + img * a.img_stride + (top + y) * a.row_stride + (left + x);

    // cloud / water mask of this pixel group, per timestep
    uint32_t cloud_t[VEC];  // bit t set => masked at timestep t
#pragma unroll
    for (int i = 0; i < VEC; ++i) cloud_t[i] = 0;
    if (a.fmask) {
      for (int t = 0; t < a.T; ++t) {
        const uint8_t* fm = a.fmask + ((static_cast<int64_t>(img) * a.T + t) * a.H + (top + y)) *
                                          static_cast<int64_t>(a.W) + (left + x);
#pragma unroll
        for (int i = 0; i < VEC; ++i)
          if (fmask_hit(__ldg(fm + i), a.fmask_bits)) cloud_t[i] |= (1u << t);
      }
      if (a.mask_any) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) cloud_t[i] = cloud_t[i] ? 0xffffffffu : 0u;
      }
    }

    uint32_t any_nodata = 0;  // bit i => pixel i is nodata in some band
    for (int t = 0; t < a.T; ++t) {
      for (int c = 0; c < a.C; ++c) {
        const int tc = t * a.C + c;
        const int sb = __ldg(a.band_idx + tc);
        typename RawVal<RawT>::type v[VEC];
        load8<RawT>(src0 + sb * a.band_stride, v);
        const float mean = __ldg(a.mean + c), sd = __ldg(a.std + c);
        float o[VEC];
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          typename RawVal<RawT>::type rv = v[i];
          if (cloud_t[i] & (1u << t)) rv = a.fill_raw;
          float f;
          bool nd;
          if (a.cm_is_one) {  // int16/uint16 -> f32 is exact; float64 -> f32 rounds like PIL mode "F"
            f = static_cast<float>(rv);
            nd = static_cast<double>(rv) == a.nodata;
          } else {
            const double d = __dmul_rn(static_cast<double>(rv), a.cm);
            f = __double2float_rn(d);
            nd = d == a.nodata;
          }
          if (a.has_nodata && nd) m |= (1u << i);
          o[i] = __fdiv_rn(__fsub_rn(f, mean), sd);
        }
        any_nodata |= m;
        if (a.out_f32) {
          float* dst = a.out_f32 +
                       (((static_cast<int64_t>(w) * a.C + c) * a.T + t) * a.win + y) * a.win + x;
          __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
          __stcs(reinterpret_cast<float4*>(dst) + 1, make_float4(o[4], o[5], o[6], o[7]));
        }
        if (a.out_patch) {
          // tubelet row (w, t, y/16, x/16), column c*256 + (y%16)*16 + x%16
          const int64_t prow = (static_cast<int64_t>(w) * a.T + t) * gp * gp + (y >> 4) * gp + (x >> 4);
          __nv_bfloat16* dst = a.out_patch + prow * (a.C * 256) + c * 256 + (y & 15) * 16 + (x & 15);
          uint4 q;
          q.x = ig::pack_bf16(o[0], o[1]);
          q.y = ig::pack_bf16(o[2], o[3]);
          q.z = ig::pack_bf16(o[4], o[5]);
          q.w = ig::pack_bf16(o[6], o[7]);
          *reinterpret_cast<uint4*>(dst) = q;
        }
        if (a.mask_elem) {
          uint8_t* dst = a.mask_elem + ((static_cast<int64_t>(w) * TC + tc) * a.win + y) * a.win + x;
          uint2 q;
          q.x = ((m >> 0) & 1u) | (((m >> 1) & 1u) << 8) | (((m >> 2) & 1u) << 16) | (((m >> 3) & 1u) << 24);
          q.y = ((m >> 4) & 1u) | (((m >> 5) & 1u) << 8) | (((m >> 6) & 1u) << 16) | (((m >> 7) & 1u) << 24);
          __stcs(reinterpret_cast<uint2*>(dst), q);
        }
      }
    }
    if (a.mask_px) {
      uint8_t* dst = a.mask_px + (static_cast<int64_t>(w) * a.win + y) * a.win + x;
      const uint32_t m = any_nodata;
      uint2 q;
      q.x = ((m >> 0) & 1u) | (((m >> 1) & 1u) << 8) | (((m >> 2) & 1u) << 16) | (((m >> 3) & 1u) << 24);
      q.y = ((m >> 4) & 1u) | (((m >> 5) & 1u) << 8) | (((m >> 6) & 1u) << 16) | (((m >> 7) & 1u) << 24);
      *reinterpret_cast<uint2*>(dst) = q;
    }
  }
}

}  // namespace

extern "C" int ig_preprocess(const void* raw, int raw_dtype, int n_img, int n_src_bands, int H,
                             int W, int64_t img_stride, int64_t band_stride, int64_t row_stride,
                             const int32_t* band_idx, int T, int C, const int32_t* win_yx,
                             int n_win, int win, double constant_multiplier, const float* mean,
                             const float* std, int has_nodata, double no_data_value,
                             const uint8_t* fmask, uint32_t fmask_bits, int masking_strategy,
                             float* out_f32, void* out_patch, uint8_t* mask_elem, uint8_t* mask_px,
                             void* stream) {
  IG_TRY(ig_check_device());
  if (n_win == 0 && n_img >= 0) return IG_OK;  // empty batch: torch hands out null pointers for 0-element tensors
  IG_REQUIRE(raw && band_idx && mean && std, IG_EINVAL, "ig_preprocess: null input pointer");
  IG_REQUIRE(raw_dtype == IG_I16 || raw_dtype == IG_U16 || raw_dtype == IG_F32 || raw_dtype == IG_F64, IG_EINVAL,
             "ig_preprocess: raw_dtype must be IG_I16, IG_U16, IG_F32 or IG_F64");
  IG_REQUIRE(T >= 1 && C >= 1 && T * C <= MAX_TC && T <= 32, IG_ESHAPE,
             "ig_preprocess: unsupported T=%d C=%d", T, C);
  IG_REQUIRE(win >= 16 && win % 16 == 0, IG_ESHAPE, "ig_preprocess: window %d not a multiple of 16", win);
  IG_REQUIRE(n_img >= 0 && n_win >= 0 && n_src_bands >= 1 && H >= win && W >= win, IG_ESHAPE,
             "ig_preprocess: bad geometry n_img=%d n_win=%d H=%d W=%d win=%d", n_img, n_win, H, W, win);
  if (!win_yx)
    IG_REQUIRE(n_win == n_img && H == win && W == win, IG_ESHAPE,
               "ig_preprocess: without a window list images must be win x win and n_win == n_img");
  IG_REQUIRE(masking_strategy == IG_MASK_EACH || masking_strategy == IG_MASK_ANY, IG_EINVAL,
             "ig_preprocess: bad masking strategy");
  IG_REQUIRE(out_f32 || out_patch || mask_elem || mask_px, IG_EINVAL, "ig_preprocess: no output requested");
  if (n_win == 0) return IG_OK;

  PreArgs a;
  a.raw = raw;
  a.img_stride = img_stride;
  a.band_stride = band_stride;
  a.row_stride = row_stride;
  a.band_idx = band_idx;
  a.win_yx = win_yx;
  a.n_win = n_win;
  a.win = win;
  a.T = T;
  a.C = C;
  a.H = H;
  a.W = W;
  a.cm = constant_multiplier;
  a.nodata = no_data_value;
  a.has_nodata = has_nodata;
  a.cm_is_one = constant_multiplier == 1.0;
  a.mean = mean;
  a.std = std;
  a.fmask = fmask_bits ? fmask : nullptr;
  a.fmask_bits = fmask_bits;
  a.mask_any = masking_strategy == IG_MASK_ANY;
  a.fill_raw = has_nodata ? static_cast<int>(no_data_value) : 0;
  a.out_f32 = out_f32;
  a.out_patch = static_cast<__nv_bfloat16*>(out_patch);
  a.mask_elem = mask_elem;
  a.mask_px = mask_px;

  const int64_t total = static_cast<int64_t>(n_win) * win * (win / VEC);
  const int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(ig_num_sms()) * 8 * 4;  // 8 resident CTAs/SM, 4 waves
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ig::ProfScope prof(ig::PROF_PREPROCESS, st);
  if (raw_dtype == IG_I16)
    preprocess_kernel<int16_t><<<static_cast<unsigned>(blocks), threads, 0, st>>>(a);
  else if (raw_dtype == IG_U16)
    preprocess_kernel<uint16_t><<<static_cast<unsigned>(blocks), threads, 0, st>>>(a);
  else if (raw_dtype == IG_F32)
    preprocess_kernel<float><<<static_cast<unsigned>(blocks), threads, 0, st>>>(a);
  else
    preprocess_kernel<double><<<static_cast<unsigned>(blocks), threads, 0, st>>>(a);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}
