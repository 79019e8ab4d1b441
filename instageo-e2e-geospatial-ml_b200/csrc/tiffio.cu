// §8(f) row 2, device side: what sits between zlib and kernel 1 when a raw tile / chip is read, and between the
// class map and zlib when predictions are written.
//
// The reference reads a raster with rasterio (GDAL): `src.read()` in instageo/model/dataloader.py:672-704, and
// writes predictions band by band in instageo/model/infer_utils.py:37-54.  Inside GDAL that is, per TIFF block:
// inflate -> undo the horizontal predictor (TIFF 6.0 section 14: every sample is the difference to its left
// neighbour of the same band; a wrapping prefix sum along x) -> byte swap -> de-interleave chunky (pixel-interleaved)
// samples into band planes.  Inflate is a serial bit stream and stays on host threads (zlib into pinned memory);
// everything after it is byte / integer work over the whole raster and runs here, writing the planar
// [bands, H, W] int16 / uint16 raster that ig_preprocess (kernel 1) reads in place.
//
// ig_tiff_unpack16: one CTA per row of a TIFF block (strip or tile).  The row (bw * spp 16-bit samples) is staged in
// shared memory with 16-byte loads, each thread scans a contiguous run of pixels per sample channel, the per-thread
// totals are combined by a warp-shuffle + cross-warp exclusive scan, and the row leaves as band planes with the widest
// store the destination alignment allows.  HBM-bound: 2 B read + 2 B written per sample.
// ig_tiff_predict: forward differencing of rows (writer side; 1- or 2-byte samples), elementwise.
#include "ig_common.cuh"

namespace {

constexpr int TIFF_THREADS = 256;
constexpr int MAX_SPP = 8;   // samples per pixel scanned in registers (chunky blocks); planar blocks have 1

struct UnpackArgs {
  const uint16_t* src;   // inflated blocks, block-major: ((plane * nby + by) * nbx + bx) * bh * bw * cs samples
  uint16_t* dst;         // [spp, H, W]
  int W, H, spp;
  int bw, bh, nbx, nby;  // block geometry (strips: bw = W, nbx = 1)
  int planes, cs;        // planar: planes = spp, cs = 1; chunky: planes = 1, cs = spp
  int predictor, swap;
};

__device__ __forceinline__ uint32_t bswap16x2(uint32_t v) { return __byte_perm(v, 0, 0x2301); }

__global__ void __launch_bounds__(TIFF_THREADS) tiff_unpack16_kernel(const UnpackArgs a) {
  extern __shared__ __align__(16) uint16_t row[];   // bw * cs samples (+ padding to 16 bytes)
  __shared__ uint32_t wtot[TIFF_THREADS / 32][MAX_SPP];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // row -> (plane, by, bx, r)
  int64_t rid = blockIdx.x;
  const int r = static_cast<int>(rid % a.bh);
  rid /= a.bh;
  const int bx = static_cast<int>(rid % a.nbx);
  rid /= a.nbx;
  const int by = static_cast<int>(rid % a.nby);
  const int plane = static_cast<int>(rid / a.nby);
  const int y = by * a.bh + r;
  if (y >= a.H) return;   // rows of the last strip / tile row beyond the image
  const int n = a.bw * a.cs;
  const uint16_t* src = a.src + (((static_cast<int64_t>(plane) * a.nby + by) * a.nbx + bx) * a.bh + r) * n;

  // ---- stage the row (byte-swapped if the file is big endian)
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int nv = n >> 3;
    for (int i = tid; i < nv; i += TIFF_THREADS) {
      uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + i);
      if (a.swap) { v.x = bswap16x2(v.x); v.y = bswap16x2(v.y); v.z = bswap16x2(v.z); v.w = bswap16x2(v.w); }
      reinterpret_cast<uint4*>(row)[i] = v;
    }
    for (int i = (nv << 3) + tid; i < n; i += TIFF_THREADS) {
      const uint16_t v = src[i];
      row[i] = a.swap ? static_cast<uint16_t>((v << 8) | (v >> 8)) : v;
    }
  } else {
    for (int i = tid; i < n; i += TIFF_THREADS) {
      const uint16_t v = src[i];
      row[i] = a.swap ? static_cast<uint16_t>((v << 8) | (v >> 8)) : v;
    }
  }
  __syncthreads();

  // ---- horizontal predictor: inclusive wrapping prefix sum along x, per sample channel
  if (a.predictor == 2) {
    const int seg = (a.bw + TIFF_THREADS - 1) / TIFF_THREADS;
    const int x_lo = min(a.bw, tid * seg), x_hi = min(a.bw, x_lo + seg);
    uint32_t run[MAX_SPP];
#pragma unroll
    for (int c = 0; c < MAX_SPP; ++c) run[c] = 0;
    for (int x = x_lo; x < x_hi; ++x) {
#pragma unroll
      for (int c = 0; c < MAX_SPP; ++c)
        if (c < a.cs) {
          run[c] += row[x * a.cs + c];
          row[x * a.cs + c] = static_cast<uint16_t>(run[c]);
        }
    }
    // exclusive scan of the per-thread totals over the block (mod 2^16 arithmetic carried in 32 bits)
    uint32_t excl[MAX_SPP];
#pragma unroll
    for (int c = 0; c < MAX_SPP; ++c) {
      excl[c] = 0;
      if (c < a.cs) {
        uint32_t v = run[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
          if (lane >= o) v += u;
        }
        if (lane == 31) wtot[wid][c] = v;
        excl[c] = v - run[c];
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < MAX_SPP; ++c)
      if (c < a.cs) {
        uint32_t base = 0;
        for (int w = 0; w < wid; ++w) base += wtot[w][c];
        excl[c] += base;
      }
    for (int x = x_lo; x < x_hi; ++x) {
#pragma unroll
      for (int c = 0; c < MAX_SPP; ++c)
        if (c < a.cs) row[x * a.cs + c] = static_cast<uint16_t>(row[x * a.cs + c] + excl[c]);
    }
    __syncthreads();
  }

  // ---- band planes out: dst[band][y][x0 + x], the widest store the row's alignment allows
  const int x0 = bx * a.bw;
  const int wv = min(a.bw, a.W - x0);   // valid pixels of this block row
  for (int c = 0; c < a.cs; ++c) {
    const int band = a.planes > 1 ? plane : c;
    uint16_t* drow = a.dst + (static_cast<int64_t>(band) * a.H + y) * a.W + x0;
    const uintptr_t ad = reinterpret_cast<uintptr_t>(drow);
    if ((ad & 15) == 0) {
      const int nv = wv >> 3;
      for (int i = tid; i < nv; i += TIFF_THREADS) {
        uint32_t p[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          p[q] = static_cast<uint32_t>(row[(8 * i + 2 * q) * a.cs + c]) |
                 (static_cast<uint32_t>(row[(8 * i + 2 * q + 1) * a.cs + c]) << 16);
        __stcs(reinterpret_cast<uint4*>(drow) + i, make_uint4(p[0], p[1], p[2], p[3]));
      }
      for (int x = (nv << 3) + tid; x < wv; x += TIFF_THREADS) drow[x] = row[x * a.cs + c];
    } else if ((ad & 7) == 0) {
      const int nv = wv >> 2;
      for (int i = tid; i < nv; i += TIFF_THREADS) {
        uint32_t p[2];
#pragma unroll
        for (int q = 0; q < 2; ++q)
          p[q] = static_cast<uint32_t>(row[(4 * i + 2 * q) * a.cs + c]) |
                 (static_cast<uint32_t>(row[(4 * i + 2 * q + 1) * a.cs + c]) << 16);
        __stcs(reinterpret_cast<uint2*>(drow) + i, make_uint2(p[0], p[1]));
      }
      for (int x = (nv << 2) + tid; x < wv; x += TIFF_THREADS) drow[x] = row[x * a.cs + c];
    } else {
      for (int x = tid; x < wv; x += TIFF_THREADS) drow[x] = row[x * a.cs + c];
    }
  }
}

// forward horizontal differencing of `rows` rows of W samples (T = uint8_t | uint16_t), one band plane
template <typename T>
__global__ void __launch_bounds__(256) tiff_predict_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t rows, int W) {
  const int64_t n = rows * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const T v = src[i];
    dst[i] = x ? static_cast<T>(v - src[i - 1]) : v;
  }
}

}  // namespace

extern "C" int ig_tiff_unpack16(const void* blocks, void* dst, int W, int H, int spp, int block_w, int block_h,
                                int planar, int predictor, int byteswap, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(blocks && dst, IG_EINVAL, "ig_tiff_unpack16: null pointer");
  IG_REQUIRE(W >= 1 && H >= 1 && spp >= 1 && block_w >= 1 && block_h >= 1, IG_ESHAPE, "ig_tiff_unpack16: bad geometry");
  IG_REQUIRE(predictor == 1 || predictor == 2, IG_EINVAL, "ig_tiff_unpack16: predictor %d (1 = none, 2 = horizontal)", predictor);
  IG_REQUIRE((reinterpret_cast<uintptr_t>(blocks) & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 1) == 0, IG_EINVAL,
             "ig_tiff_unpack16: buffers must be 2-byte aligned");
  UnpackArgs a;
  a.src = static_cast<const uint16_t*>(blocks);
  a.dst = static_cast<uint16_t*>(dst);
  a.W = W, a.H = H, a.spp = spp;
  a.bw = block_w, a.bh = block_h;
  a.nbx = (W + block_w - 1) / block_w, a.nby = (H + block_h - 1) / block_h;
  a.planes = planar ? spp : 1;
  a.cs = planar ? 1 : spp;
  a.predictor = predictor, a.swap = byteswap ? 1 : 0;
  IG_REQUIRE(a.cs <= MAX_SPP, IG_ESHAPE, "ig_tiff_unpack16: %d interleaved samples per pixel (at most %d)", a.cs, MAX_SPP);
  const size_t smem = (static_cast<size_t>(a.bw) * a.cs * 2 + 15) / 16 * 16;
  IG_REQUIRE(smem <= 200 * 1024, IG_ESHAPE, "ig_tiff_unpack16: a block row of %zu bytes does not fit shared memory", smem);
  static IgPerDevice configured = {};
  if (static_cast<int>(smem) > configured.get() && smem > 48 * 1024) {
    IG_CUDA_OK(cudaFuncSetAttribute(tiff_unpack16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured.set(static_cast<int>(smem));
  }
  const int64_t nrows = static_cast<int64_t>(a.planes) * a.nby * a.nbx * a.bh;
  IG_REQUIRE(nrows < (1ll << 31), IG_ESHAPE, "ig_tiff_unpack16: too many block rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ig::ProfScope prof(ig::PROF_PREPROCESS, st);
  tiff_unpack16_kernel<<<static_cast<unsigned>(nrows), TIFF_THREADS, smem, st>>>(a);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

extern "C" int ig_tiff_predict(const void* src, void* dst, int sample_bytes, int64_t rows, int W, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(src && dst && src != dst, IG_EINVAL, "ig_tiff_predict: null or aliased pointers");
  IG_REQUIRE(sample_bytes == 1 || sample_bytes == 2, IG_EINVAL, "ig_tiff_predict: 1- or 2-byte samples");
  IG_REQUIRE(rows >= 0 && W >= 1, IG_ESHAPE, "ig_tiff_predict: bad geometry");
  if (rows == 0) return IG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n = rows * W;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(ig_num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  ig::ProfScope prof(ig::PROF_MISC, st);
  if (sample_bytes == 1)
    tiff_predict_kernel<uint8_t><<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), rows, W);
  else
    tiff_predict_kernel<uint16_t><<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), rows, W);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}
