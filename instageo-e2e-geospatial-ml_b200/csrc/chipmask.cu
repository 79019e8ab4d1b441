// Chip-creation masking at tile scale -- one HBM-bound pass.
//
// Reference, per chip and on the host through xarray (instageo/data/hls_utils.py:359-403):
//   apply_mask(chip, fmask, no_data_value=0, decode_fmask_value, "HLS", mask_types, strategy)
//                                             instageo/data/data_pipeline.py:229-267, hls_utils.py:77-86
//   chip.clip(min=0, max=10000)               hls_utils.py:371, :386
//   chip.where(chip != 0).count() == 0  ->  skip (all cloud)          hls_utils.py:389
//   mask_segmentation_map(chip, seg_map, 0, strategy)                  data_pipeline.py:66-98
//   seg_map.where(seg_map != -1).count() == 0  ->  skip (empty label)  hls_utils.py:395
//   astype(uint16) / astype(int8)             hls_utils.py:398-401
// i.e. five full passes with float64 / NaN intermediates.  Here: every band of 8 consecutive pixels is
// read once (16-byte loads), the Fmask bytes of the T mask steps once, and the uint16 chip, the int8
// label map and the two "anything left?" counts are produced in the same pass.  The planes of a
// [bands, H, W] raster are contiguous, and the operation is pixel-wise, so pixels are addressed flat
// (H*W), which keeps 16-byte alignment even when W*2 is not a multiple of 16 (W = 3660).
#include "ig_common.cuh"

namespace {

constexpr int THREADS = 256;

struct MaskArgs {
  const void* chip;          // [n_bands, P] int16 | uint16
  int chip_is_signed;
  int n_bands, n_steps;      // n_steps = mask timesteps; band b belongs to step b / (n_bands / n_steps)
  long long P;               // pixels per plane
  const uint8_t* fmask;      // [n_steps, P] or nullptr
  uint32_t bits;             // bit p set => Fmask bit p masks the pixel
  int strategy;              // IG_MASK_EACH | IG_MASK_ANY
  int no_data, clip_lo, clip_hi;
  uint16_t* out;             // [n_bands, P]
  const int8_t* seg_in;      // [P] or nullptr
  int seg_strategy, seg_no_data;
  int8_t* seg_out;           // [P]
  unsigned long long* counts;  // [0] chip elements != no_data after clipping, [1] label pixels != seg_no_data
};

// decode_fmask_value(v, p) = (v // 2**p) - ((v // 2**p) // 2) * 2 on a uint8 is bit p of v
__device__ __forceinline__ bool hit(uint32_t v, uint32_t bits) { return (v & bits) != 0; }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <int VEC>
__global__ void __launch_bounds__(THREADS) chip_mask_kernel(const MaskArgs a) {
  const long long nvec = (a.P + VEC - 1) / VEC;
  const int per_step = a.n_bands / a.n_steps;
  unsigned long long kept = 0, lab_kept = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * THREADS) {
    const long long p0 = i * VEC;
    // cloud bits of the pixel's mask steps: bit t of cloud[j] = step t masks pixel j
    uint32_t cloud[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) cloud[j] = 0;
    if (a.fmask) {
      for (int t = 0; t < a.n_steps; ++t) {
        uint8_t fb[VEC];
        if (VEC == 8) {
          *reinterpret_cast<uint2*>(fb) = __ldcs(reinterpret_cast<const uint2*>(a.fmask + t * a.P + p0));
        } else {
          fb[0] = a.fmask[t * a.P + p0];
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (hit(fb[j], a.bits)) cloud[j] |= 1u << t;
      }
      if (a.strategy == IG_MASK_ANY) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) cloud[j] = cloud[j] ? 0xffffffffu : 0u;
      }
    }
    uint32_t any_valid = 0, all_valid = (1u << VEC) - 1;  // bit j: over bands of (chip != no_data)
    // bands in groups of BG: all loads of a group are requested before the first one is consumed
    constexpr int BG = 6;
    for (int b0 = 0; b0 < a.n_bands; b0 += BG) {
      uint16_t raw[BG][VEC];
#pragma unroll
      for (int u = 0; u < BG; ++u) {
        if (b0 + u < a.n_bands) {
          const uint16_t* src = static_cast<const uint16_t*>(a.chip) + (b0 + u) * a.P + p0;
          if (VEC == 8) *reinterpret_cast<uint4*>(raw[u]) = __ldcs(reinterpret_cast<const uint4*>(src));
          else raw[u][0] = *src;
        }
      }
#pragma unroll
      for (int u = 0; u < BG; ++u) {
        if (b0 + u < a.n_bands) {
          const int b = b0 + u, t = b / per_step;
          uint16_t o[VEC];
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            int v = a.chip_is_signed ? static_cast<int>(static_cast<int16_t>(raw[u][j])) : static_cast<int>(raw[u][j]);
            if ((cloud[j] >> t) & 1u) v = a.no_data;
            v = clampi(v, a.clip_lo, a.clip_hi);
            const bool valid = v != a.no_data;
            kept += valid;
            any_valid |= static_cast<uint32_t>(valid) << j;
            if (!valid) all_valid &= ~(1u << j);
            o[j] = static_cast<uint16_t>(v);
          }
          if (VEC == 8) __stcs(reinterpret_cast<uint4*>(a.out + b * a.P + p0), *reinterpret_cast<const uint4*>(o));
          else a.out[b * a.P + p0] = o[0];
        }
      }
    }
    if (a.seg_in) {
      const uint32_t ok = a.seg_strategy == IG_MASK_EACH ? any_valid : all_valid;
      int8_t s[VEC];
      if (VEC == 8) {
        *reinterpret_cast<uint2*>(s) = __ldcs(reinterpret_cast<const uint2*>(a.seg_in + p0));
      } else {
        s[0] = a.seg_in[p0];
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        if (!((ok >> j) & 1u)) s[j] = static_cast<int8_t>(a.seg_no_data);
        lab_kept += s[j] != a.seg_no_data;
      }
      if (VEC == 8) {
        __stcs(reinterpret_cast<uint2*>(a.seg_out + p0), *reinterpret_cast<const uint2*>(s));
      } else {
        a.seg_out[p0] = s[0];
      }
    }
  }
  if (a.counts) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kept += __shfl_xor_sync(0xffffffffu, kept, o);
      lab_kept += __shfl_xor_sync(0xffffffffu, lab_kept, o);
    }
    __shared__ unsigned long long red[2][THREADS / 32];
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = kept, red[1][threadIdx.x >> 5] = lab_kept;
    __syncthreads();
    if (threadIdx.x < 2) {
      unsigned long long t = 0;
      for (int w = 0; w < THREADS / 32; ++w) t += red[threadIdx.x][w];
      if (t) atomicAdd(a.counts + threadIdx.x, t);
    }
  }
}

// Vector kernel: one thread = 8 consecutive pixels = one 16-byte vector per band, arithmetic on packed
// int16x2 words (VIMNMX.S16x2 / .U16x2 are native on sm_100): per pixel PAIR one LOP3 to substitute the
// fill value under the cloud mask, a packed min and max for the clip, a packed compare against the fill
// value whose lane masks feed the any/all-band label rule (OR / AND) and the valid-element count (POPC).
// The cloud bits of the <= 8 mask steps live in one 64-bit register (byte t = the 8 pixels of step t) and
// are expanded to lane masks only when the band's step changes (a warp-uniform branch).
template <bool SIGNED>
__global__ void __launch_bounds__(THREADS) chip_mask_vec_kernel(const MaskArgs a, uint32_t nd2, uint32_t lo2,
                                                                uint32_t hi2, int nd_reachable) {
  const long long nvec = a.P >> 3;
  const int per_step = a.n_bands / a.n_steps;
  uint32_t kept_bits = 0;  // 16 per valid element
  unsigned long long kept = 0, lab_kept = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * THREADS) {
    const long long p0 = i << 3;
    unsigned long long cbits = 0;
    if (a.fmask) {
      uint32_t any8 = 0;
      for (int t = 0; t < a.n_steps; ++t) {
        const uint2 f = __ldcs(reinterpret_cast<const uint2*>(a.fmask + t * a.P + p0));
        // byte j of (f & bits) non-zero -> bit j
        const uint32_t bx = f.x & (a.bits * 0x01010101u), by = f.y & (a.bits * 0x01010101u);
        uint32_t m8 = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          m8 |= ((bx >> (8 * j)) & 0xffu) ? (1u << j) : 0u;
          m8 |= ((by >> (8 * j)) & 0xffu) ? (16u << j) : 0u;
        }
        cbits |= static_cast<unsigned long long>(m8) << (8 * t);
        any8 |= m8;
      }
      if (a.strategy == IG_MASK_ANY) cbits = any8 * 0x0101010101010101ull;
    }
    uint32_t cm[4] = {0, 0, 0, 0};
    int t_cur = -1;
    uint32_t any_valid[4] = {0, 0, 0, 0}, all_valid[4] = {~0u, ~0u, ~0u, ~0u};
    constexpr int BG = 6;
    for (int b0 = 0; b0 < a.n_bands; b0 += BG) {
      uint4 raw[BG];
#pragma unroll
      for (int u = 0; u < BG; ++u)
        if (b0 + u < a.n_bands)
          raw[u] = __ldcs(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(a.chip) + (b0 + u) * a.P + p0));
#pragma unroll
      for (int u = 0; u < BG; ++u) {
        if (b0 + u < a.n_bands) {
          const int b = b0 + u, t = b / per_step;
          if (t != t_cur) {
            t_cur = t;
            const uint32_t m8 = static_cast<uint32_t>(cbits >> (8 * t)) & 0xffu;
#pragma unroll
            for (int w = 0; w < 4; ++w)
              cm[w] = ((m8 >> (2 * w)) & 1u ? 0x0000ffffu : 0u) | ((m8 >> (2 * w + 1)) & 1u ? 0xffff0000u : 0u);
          }
          uint32_t x[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            uint32_t v = (x[w] & ~cm[w]) | (nd2 & cm[w]);
            v = SIGNED ? __vmaxs2(__vmins2(v, hi2), lo2) : __vmaxu2(__vminu2(v, hi2), lo2);
            const uint32_t ne = nd_reachable ? __vcmpne2(v, nd2) : 0xffffffffu;
            kept_bits += __popc(ne);
            any_valid[w] |= ne;
            all_valid[w] &= ne;
            x[w] = v;
          }
          __stcs(reinterpret_cast<uint4*>(a.out + b * a.P + p0), make_uint4(x[0], x[1], x[2], x[3]));
        }
      }
    }
    kept += kept_bits >> 4;
    kept_bits = 0;
    if (a.seg_in) {
      const uint2 sv = __ldcs(reinterpret_cast<const uint2*>(a.seg_in + p0));
      uint32_t sw[2] = {sv.x, sv.y};
      const uint32_t fill = static_cast<uint32_t>(static_cast<uint8_t>(a.seg_no_data)) * 0x01010101u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // lane masks of pixels (4h .. 4h+3) -> byte masks
        const uint32_t* src = a.seg_strategy == IG_MASK_EACH ? any_valid : all_valid;
        const uint32_t l0 = src[2 * h], l1 = src[2 * h + 1];
        const uint32_t ok = (l0 & 0x000000ffu) | ((l0 >> 8) & 0x0000ff00u) | ((l1 << 16) & 0x00ff0000u) | (l1 & 0xff000000u);
        sw[h] = (sw[h] & ok) | (fill & ~ok);
        lab_kept += __popc(__vcmpne4(sw[h], fill)) >> 3;
      }
      __stcs(reinterpret_cast<uint2*>(a.seg_out + p0), make_uint2(sw[0], sw[1]));
    }
  }
  if (a.counts) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kept += __shfl_xor_sync(0xffffffffu, kept, o);
      lab_kept += __shfl_xor_sync(0xffffffffu, lab_kept, o);
    }
    __shared__ unsigned long long red[2][THREADS / 32];
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = kept, red[1][threadIdx.x >> 5] = lab_kept;
    __syncthreads();
    if (threadIdx.x < 2) {
      unsigned long long t = 0;
      for (int w = 0; w < THREADS / 32; ++w) t += red[threadIdx.x][w];
      if (t) atomicAdd(a.counts + threadIdx.x, t);
    }
  }
}

}  // namespace

extern "C" int ig_chip_mask(const void* chip, int chip_dtype, int n_bands, int64_t height, int64_t width,
                            const uint8_t* fmask, int n_mask_steps, uint32_t fmask_bits, int masking_strategy,
                            int no_data_value, int clip_min, int clip_max, uint16_t* out, const int8_t* seg_map,
                            int seg_masking_strategy, int seg_no_data_value, int8_t* seg_out,
                            unsigned long long* counts, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(chip_dtype == IG_I16 || chip_dtype == IG_U16, IG_EINVAL, "ig_chip_mask: chip must be int16 or uint16");
  IG_REQUIRE(n_bands >= 1 && height >= 0 && width >= 0, IG_EINVAL, "ig_chip_mask: bad shape");
  // clip_min > clip_max: no clipping, the output keeps the input's element type (apply_mask alone)
  IG_REQUIRE(clip_min > clip_max || (clip_min >= 0 && clip_max <= 65535), IG_EINVAL,
             "ig_chip_mask: clip range [%d, %d] must lie inside uint16", clip_min, clip_max);
  if (clip_min > clip_max) clip_min = chip_dtype == IG_I16 ? -32768 : 0, clip_max = chip_dtype == IG_I16 ? 32767 : 65535;
  IG_REQUIRE(masking_strategy == IG_MASK_EACH || masking_strategy == IG_MASK_ANY, IG_EINVAL,
             "ig_chip_mask: masking strategy %d", masking_strategy);
  IG_REQUIRE(seg_masking_strategy == IG_MASK_EACH || seg_masking_strategy == IG_MASK_ANY, IG_EINVAL,
             "ig_chip_mask: label masking strategy %d", seg_masking_strategy);
  IG_REQUIRE(seg_no_data_value >= -128 && seg_no_data_value <= 127, IG_EINVAL, "ig_chip_mask: label fill outside int8");
  const bool use_fmask = fmask != nullptr && fmask_bits != 0;
  if (use_fmask)
    IG_REQUIRE(n_mask_steps >= 1 && n_mask_steps <= 32 && n_bands % n_mask_steps == 0, IG_ESHAPE,
               "ig_chip_mask: %d bands cannot be split over %d mask steps", n_bands, n_mask_steps);
  IG_REQUIRE((seg_map == nullptr) == (seg_out == nullptr), IG_EINVAL, "ig_chip_mask: seg_map and seg_out go together");
  const long long P = static_cast<long long>(height) * width;
  if (P == 0) return IG_OK;
  IG_REQUIRE(chip && out, IG_EINVAL, "ig_chip_mask: null pointer");
  MaskArgs a{};
  a.chip = chip, a.chip_is_signed = chip_dtype == IG_I16;
  a.n_bands = n_bands, a.n_steps = use_fmask ? n_mask_steps : 1, a.P = P;
  a.fmask = use_fmask ? fmask : nullptr, a.bits = fmask_bits & 0xffu, a.strategy = masking_strategy;
  a.no_data = no_data_value, a.clip_lo = clip_min, a.clip_hi = clip_max;
  a.out = out, a.seg_in = seg_map, a.seg_strategy = seg_masking_strategy, a.seg_no_data = seg_no_data_value;
  a.seg_out = seg_out, a.counts = counts;
  auto al = [](const void* p, int n) { return (reinterpret_cast<uintptr_t>(p) & (n - 1)) == 0; };
  const bool vec = P % 8 == 0 && al(chip, 16) && al(out, 16) && (!use_fmask || (al(fmask, 8) && n_mask_steps <= 8)) &&
                   (!seg_map || (al(seg_map, 8) && al(seg_out, 8)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = vec ? P / 8 : P;
  long long blocks = (work + THREADS - 1) / THREADS;
  const long long cap = 8ll * ig_num_sms();
  if (blocks > cap) blocks = cap;
  ig::ProfScope prof(ig::PROF_PREPROCESS, st);
  if (vec) {
    // the fill value as the kernel's 16-bit lanes see it: substituted BEFORE the clip, compared AFTER it
    const int type_lo = a.chip_is_signed ? -32768 : 0, type_hi = a.chip_is_signed ? 32767 : 65535;
    const int nd_in = no_data_value < type_lo ? type_lo : (no_data_value > type_hi ? type_hi : no_data_value);
    const int nd_after = nd_in < clip_min ? clip_min : (nd_in > clip_max ? clip_max : nd_in);
    // a fill value outside the input type saturates; it then behaves like the scalar path only if the clip maps
    // both to the same number, which holds whenever it is representable or lies outside the clip range
    const bool same = (no_data_value < clip_min ? clip_min : (no_data_value > clip_max ? clip_max : no_data_value)) == nd_after;
    // clip bounds as 16-bit lanes of the input's type (same result for every representable input)
    const int eff_lo = clip_min < type_lo ? type_lo : clip_min, eff_hi = clip_max > type_hi ? type_hi : clip_max;
    if (same && eff_lo <= eff_hi) {
      clip_min = eff_lo, clip_max = eff_hi;
      auto pk = [](int v) { return (static_cast<uint32_t>(v) & 0xffffu) * 0x00010001u; };
      const int reachable = no_data_value >= eff_lo && no_data_value <= eff_hi;  // else every element is valid
      if (a.chip_is_signed)
        chip_mask_vec_kernel<true><<<static_cast<unsigned>(blocks), THREADS, 0, st>>>(a, pk(nd_in), pk(clip_min), pk(clip_max), reachable);
      else
        chip_mask_vec_kernel<false><<<static_cast<unsigned>(blocks), THREADS, 0, st>>>(a, pk(nd_in), pk(clip_min), pk(clip_max), reachable);
    } else {
      chip_mask_kernel<8><<<static_cast<unsigned>(blocks), THREADS, 0, st>>>(a);
    }
  } else
    chip_mask_kernel<1><<<static_cast<unsigned>(blocks), THREADS, 0, st>>>(a);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}
