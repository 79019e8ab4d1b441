// Kernel 5 -- overlap-averaging stitch of sliding-window logits (HBM-bound, bit-exact).
//
// Specification: SURVEY.md Appendix A.6 (frozen in oracle/stitch.py).  The reference snapshot
// has no stitch; it contributes the window order (instageo/model/dataloader.py:655-664, top
// outer / left inner), the first-max-wins argmax -> int8 (instageo/model/infer_utils.py:96-101)
// and the nodata comparison (instageo/model/dataloader.py:899).
//
// Gather form.  A warp owns a 128-pixel segment of RPI consecutive output rows; a thread owns 4
// consecutive pixels of it and walks the <= ceil(win/stride)^2 windows covering them in row-major
// window order, so the float32 sum of every pixel is formed in exactly the order of the scatter-form
// oracle (0 + l1 + l2 ...) and the result is bit-identical for any row-stripe sharding.
//
// What the first version got wrong (ncu, profiles/r01_ncu_stitch_before.txt): 120 thread-instructions
// per pixel at nc = 2 and 128 registers at nc = 13 -- the kernel was issue- and latency-bound at 16 %
// of DRAM throughput.  This version
//   * finds a thread's covering windows with ONE binary search per axis (upper bound, then a short
//     downward scan), once per RPI rows, and advances the y cover incrementally;
//   * when the 4 pixels sit 16-byte aligned inside every covering window (always true for strides and
//     origins that are multiples of 4) reads one float4 per (window, class, row) and batches R rows x CH
//     classes of loads before the first add: R*CH*16 bytes in flight per thread, a warp reads 512
//     contiguous bytes per instruction;
//   * walks the classes in chunks of CH <= 8 with a running first-max argmax, so 13 classes cost the
//     registers of 7 and any nc <= 32 takes the same code path;
//   * hoists the cover count: every pixel of a thread-row shares it, so "is the divisor a power of two"
//     (=> exact scaling by the reciprocal instead of the IEEE division sequence) is one uniform branch;
//   * counts the class histogram in a separate pass over the int8 map (13 MB, L2-resident) with byte
//     compares instead of warp match + shared atomics per pixel.
// Algorithmic bytes per tile: n_win*nc*win^2*4 (logits read once) + H*W (map) [+ H*W nodata].
#include <stdlib.h>

#include <algorithm>

#include "ig_common.cuh"

namespace {

constexpr int MAX_AX = 256;   // window origins per axis held in shared memory
constexpr int MAX_NC = 32;
constexpr int PX = 4;         // pixels per thread along x
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int SEG = 32 * PX;  // pixels per warp segment

struct StitchArgs {
  const float* logits;
  int n_win, win_base, nc, win;
  const int32_t* ys;
  const int32_t* xs;
  int ny, nx, H, W, y0, y1;
  const uint8_t* nodata_px;
  int nodata_class;
  float* avg;
  int8_t* cls;
  int ptr_ok;   // every buffer 16-byte aligned: vector path allowed
  int rpi;      // rows per warp item
  int nseg;     // 128-pixel segments per row
  int nitems;   // nseg * ceil(rows / rpi)
};

// windows covering coordinate p: origins o (sorted) with o <= p < o + win  ->  index range [lo, hi)
__device__ __forceinline__ void cover_range(const int* org, int n, int win, int p, int& lo, int& hi) {
  int a = 0, b = n;  // first index with org[i] > p
  while (a < b) {
    const int m = (a + b) >> 1;
    if (org[m] > p) b = m; else a = m + 1;
  }
  hi = a;
  while (a > 0 && org[a - 1] + win > p) --a;
  lo = a;
}

__device__ __forceinline__ uint32_t pack4(const int (&c)[PX]) {
  return (static_cast<uint32_t>(c[0]) & 0xffu) | ((static_cast<uint32_t>(c[1]) & 0xffu) << 8) |
         ((static_cast<uint32_t>(c[2]) & 0xffu) << 16) | ((static_cast<uint32_t>(c[3]) & 0xffu) << 24);
}

// R rows x 4 aligned pixels, every row with the same covering windows [ylo,yhi) x [xlo,xhi).
template <int CH, int R>
__device__ __forceinline__ void rows_vec(const StitchArgs& a, const int* ysrc, const int* s_xs, int y, int ylo,
                                         int yhi, int xlo, int xhi, int x0) {
  const int64_t plane = static_cast<int64_t>(a.win) * a.win;
  const int rows = a.y1 - a.y0;
  int cnt = 0;
  for (int iy = ylo; iy < yhi; ++iy)
    for (int ix = xlo; ix < xhi; ++ix) {
      const int wi = iy * a.nx + ix - a.win_base;
      cnt += (wi >= 0 && wi < a.n_win) ? 1 : 0;
    }
  const float cntf = static_cast<float>(cnt);
  // 1, 2, 4, ... covering windows: the true division is an exact scaling, x * (1/cnt) is bit-identical
  const bool pow2 = (cnt & (cnt - 1)) == 0;
  const float inv = cnt > 0 ? __frcp_rn(cntf) : 0.f;  // cnt == 0: sums are 0, 0 * 0 = 0 = the oracle's avg
  float best[R][PX];
  int bi[R][PX];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int q = 0; q < PX; ++q) {
      best[r][q] = 0.f;
      bi[r][q] = 0;
    }
  for (int c0 = 0; c0 < a.nc; c0 += CH) {
    float4 acc[R][CH];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < CH; ++k) acc[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = ylo; iy < yhi; ++iy) {
      const int ly = y - ysrc[iy];
      for (int ix = xlo; ix < xhi; ++ix) {
        const int wi = iy * a.nx + ix - a.win_base;
        if (wi < 0 || wi >= a.n_win) continue;
        const float* p = a.logits + (static_cast<int64_t>(wi) * a.nc + c0) * plane + ly * a.win + (x0 - s_xs[ix]);
        float4 v[R][CH];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < CH; ++k)
            v[r][k] = (c0 + k < a.nc) ? __ldcs(reinterpret_cast<const float4*>(p + r * a.win + k * plane))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < CH; ++k) {
            acc[r][k].x = __fadd_rn(acc[r][k].x, v[r][k].x);
            acc[r][k].y = __fadd_rn(acc[r][k].y, v[r][k].y);
            acc[r][k].z = __fadd_rn(acc[r][k].z, v[r][k].z);
            acc[r][k].w = __fadd_rn(acc[r][k].w, v[r][k].w);
          }
      }
    }
    if (pow2) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          acc[r][k].x = __fmul_rn(acc[r][k].x, inv);
          acc[r][k].y = __fmul_rn(acc[r][k].y, inv);
          acc[r][k].z = __fmul_rn(acc[r][k].z, inv);
          acc[r][k].w = __fmul_rn(acc[r][k].w, inv);
        }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          acc[r][k].x = __fdiv_rn(acc[r][k].x, cntf);
          acc[r][k].y = __fdiv_rn(acc[r][k].y, cntf);
          acc[r][k].z = __fdiv_rn(acc[r][k].z, cntf);
          acc[r][k].w = __fdiv_rn(acc[r][k].w, cntf);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        if (c0 + k < a.nc) {
          const float vv[PX] = {acc[r][k].x, acc[r][k].y, acc[r][k].z, acc[r][k].w};
#pragma unroll
          for (int q = 0; q < PX; ++q)
            if (c0 + k == 0 || vv[q] > best[r][q]) {  // strict > : first maximum wins (torch.argmax)
              best[r][q] = vv[q];
              bi[r][q] = c0 + k;
            }
          if (a.avg)
            __stcs(reinterpret_cast<float4*>(a.avg + (static_cast<int64_t>(c0 + k) * rows + (y + r - a.y0)) * a.W + x0),
                   acc[r][k]);
        }
      }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int c4[PX];
#pragma unroll
    for (int q = 0; q < PX; ++q) c4[q] = cnt > 0 ? bi[r][q] : a.nodata_class;
    if (a.nodata_px) {
      const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(a.nodata_px + static_cast<int64_t>(y + r) * a.W + x0));
#pragma unroll
      for (int q = 0; q < PX; ++q)
        if ((m >> (8 * q)) & 0xffu) c4[q] = a.nodata_class;
    }
    *reinterpret_cast<uint32_t*>(a.cls + static_cast<int64_t>(y + r - a.y0) * a.W + x0) = pack4(c4);
  }
}

// General path: one row, up to 4 pixels, each with its own cover set and scalar loads.
template <int CH>
__device__ __noinline__ void row_scalar(const StitchArgs& a, const int* ysrc, const int* s_xs, int y, int ylo,
                                        int yhi, int x0) {
  const int64_t plane = static_cast<int64_t>(a.win) * a.win;
  const int rows = a.y1 - a.y0;
  for (int q = 0; q < PX; ++q) {
    const int x = x0 + q;
    if (x >= a.W) break;
    int lo, hi;
    cover_range(s_xs, a.nx, a.win, x, lo, hi);
    int cnt = 0;
    for (int iy = ylo; iy < yhi; ++iy)
      for (int ix = lo; ix < hi; ++ix) {
        const int wi = iy * a.nx + ix - a.win_base;
        cnt += (wi >= 0 && wi < a.n_win) ? 1 : 0;
      }
    const float cntf = static_cast<float>(cnt);
    float best = 0.f;
    int bi = 0;
    for (int c0 = 0; c0 < a.nc; c0 += CH) {
      float acc[CH];
#pragma unroll
      for (int k = 0; k < CH; ++k) acc[k] = 0.f;
      for (int iy = ylo; iy < yhi; ++iy) {
        const int ly = y - ysrc[iy];
        for (int ix = lo; ix < hi; ++ix) {
          const int wi = iy * a.nx + ix - a.win_base;
          if (wi < 0 || wi >= a.n_win) continue;
          const float* p = a.logits + (static_cast<int64_t>(wi) * a.nc + c0) * plane + ly * a.win + (x - s_xs[ix]);
#pragma unroll
          for (int k = 0; k < CH; ++k)
            if (c0 + k < a.nc) acc[k] = __fadd_rn(acc[k], __ldcs(p + k * plane));
        }
      }
#pragma unroll
      for (int k = 0; k < CH; ++k)
        if (c0 + k < a.nc) {
          const float v = cnt > 0 ? __fdiv_rn(acc[k], cntf) : 0.f;
          if (c0 + k == 0 || v > best) {
            best = v;
            bi = c0 + k;
          }
          if (a.avg) a.avg[(static_cast<int64_t>(c0 + k) * rows + (y - a.y0)) * a.W + x] = v;
        }
    }
    int c = cnt > 0 ? bi : a.nodata_class;
    if (a.nodata_px && a.nodata_px[static_cast<int64_t>(y) * a.W + x]) c = a.nodata_class;
    a.cls[static_cast<int64_t>(y - a.y0) * a.W + x] = static_cast<int8_t>(c);
  }
}

template <int CH, int R>
__global__ void __launch_bounds__(THREADS) stitch_kernel(const StitchArgs a) {
  __shared__ int s_xs[MAX_AX];
  __shared__ int s_ys[MAX_AX];
  for (int i = threadIdx.x; i < a.nx; i += blockDim.x) s_xs[i] = a.xs[i];
  const int ny_s = a.ny < MAX_AX ? a.ny : MAX_AX;
  for (int i = threadIdx.x; i < ny_s; i += blockDim.x) s_ys[i] = a.ys[i];
  __syncthreads();
  const int* ysrc = a.ny <= MAX_AX ? s_ys : a.ys;  // tall grids fall back to global (L1-cached) origins
  const int item = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (item >= a.nitems) return;
  const int rg = item / a.nseg, seg = item - rg * a.nseg;
  const int x0 = seg * SEG + (threadIdx.x & 31) * PX;
  if (x0 >= a.W) return;
  const int y_begin = a.y0 + rg * a.rpi;
  const int y_end = min(a.y1, y_begin + a.rpi);

  int xlo, xhi;
  cover_range(s_xs, a.nx, a.win, x0, xlo, xhi);
  // same cover set for all 4 pixels, and 16-byte aligned inside every covering window?
  bool vec = a.ptr_ok && (x0 + PX <= a.W) && ((a.win & 3) == 0) && ((a.W & 3) == 0);
  vec = vec && (xhi == a.nx || s_xs[xhi] > x0 + PX - 1) && (xlo == xhi || s_xs[xlo] + a.win > x0 + PX - 1);
  for (int ix = xlo; ix < xhi; ++ix) vec = vec && (((x0 - s_xs[ix]) & 3) == 0);

  int ylo, yhi;
  cover_range(ysrc, a.ny, a.win, y_begin, ylo, yhi);
  int y = y_begin;
  while (y < y_end) {
    // origins are sorted: the cover range only moves forward from one row to the next
    while (ylo < a.ny && ysrc[ylo] + a.win <= y) ++ylo;
    while (yhi < a.ny && ysrc[yhi] <= y) ++yhi;
    if (!vec) {
      row_scalar<CH>(a, ysrc, s_xs, y, ylo, yhi, x0);
      ++y;
      continue;
    }
    const int yl = y + R - 1;  // do R rows at once when they share the cover set
    if (R > 1 && yl < y_end && (yhi == a.ny || ysrc[yhi] > yl) && (ylo == yhi || ysrc[ylo] + a.win > yl)) {
      rows_vec<CH, R>(a, ysrc, s_xs, y, ylo, yhi, xlo, xhi, x0);
      y += R;
    } else {
      rows_vec<CH, 1>(a, ysrc, s_xs, y, ylo, yhi, xlo, xhi, x0);
      ++y;
    }
  }
}

// Class histogram of the finished int8 map: 16 pixels per load, one byte-compare + popc per class and
// word, per-thread counters, one warp reduction and one global atomic per class per block.
template <int NB>
__global__ void __launch_bounds__(256) class_hist_kernel(const int8_t* __restrict__ cls, int64_t n, int nc,
                                                         unsigned long long* __restrict__ hist) {
  __shared__ unsigned int s_hist[MAX_NC + 1];
  if (threadIdx.x <= MAX_NC) s_hist[threadIdx.x] = 0;
  __syncthreads();
  unsigned int cnt[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) cnt[k] = 0;
  unsigned int total = 0;
  const int64_t nvec = ((reinterpret_cast<uintptr_t>(cls) & 15) == 0) ? n / 16 : 0;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = tid; i < nvec; i += nthr) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(cls) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < NB; ++k)
      if (k < nc) {
        const uint32_t pat = 0x01010101u * static_cast<uint32_t>(k);
#pragma unroll
        for (int j = 0; j < 4; ++j) cnt[k] += __popc(__vcmpeq4(w[j], pat)) >> 3;
      }
    total += 16;
  }
  for (int64_t i = nvec * 16 + tid; i < n; i += nthr) {
    const int c = cls[i];
#pragma unroll
    for (int k = 0; k < NB; ++k)
      if (k < nc) cnt[k] += (c == k) ? 1u : 0u;
    total += 1;
  }
  unsigned int in_range = 0;
#pragma unroll
  for (int k = 0; k < NB; ++k)
    if (k < nc) {
      in_range += cnt[k];
      const unsigned int s = __reduce_add_sync(0xffffffffu, cnt[k]);
      if ((threadIdx.x & 31) == 0 && s) atomicAdd(&s_hist[k], s);
    }
  const unsigned int nd = __reduce_add_sync(0xffffffffu, total - in_range);  // anything outside [0, nc)
  if ((threadIdx.x & 31) == 0 && nd) atomicAdd(&s_hist[nc], nd);
  __syncthreads();
  if (threadIdx.x <= nc && s_hist[threadIdx.x])
    atomicAdd(hist + threadIdx.x, static_cast<unsigned long long>(s_hist[threadIdx.x]));
}

template <int CH, int R>
void launch_stitch(const StitchArgs& a, int blocks, cudaStream_t st) {
  stitch_kernel<CH, R><<<blocks, THREADS, 0, st>>>(a);
}

}  // namespace

extern "C" int ig_stitch(const float* win_logits, int n_win, int win_base, int nc, int win,
                         const int32_t* ys, int ny, const int32_t* xs, int nx, int H, int W,
                         int y0, int y1, const uint8_t* nodata_px, int nodata_class, float* avg,
                         int8_t* class_map, unsigned long long* hist, void* stream) {
  IG_TRY(ig_check_device());
  if (y1 == y0 && y0 >= 0 && y0 <= H) return IG_OK;  // empty stripe (torch passes null for 0-element tensors)
  IG_REQUIRE(win_logits && ys && xs && class_map, IG_EINVAL, "ig_stitch: null pointer");
  IG_REQUIRE(nc >= 1 && nc <= MAX_NC, IG_ESHAPE, "ig_stitch: nc=%d unsupported (1..%d)", nc, MAX_NC);
  IG_REQUIRE(nx >= 1 && nx <= MAX_AX && ny >= 1 && ny <= 65535, IG_ESHAPE,
             "ig_stitch: window grid %dx%d unsupported", ny, nx);
  IG_REQUIRE(0 <= y0 && y0 <= y1 && y1 <= H && W >= 1 && win >= 1, IG_ESHAPE,
             "ig_stitch: bad stripe [%d,%d) of H=%d", y0, y1, H);
  if (y1 == y0) return IG_OK;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  StitchArgs a{win_logits, n_win, win_base, nc, win, ys, xs, ny, nx, H, W, y0, y1,
               nodata_px, nodata_class, avg, class_map,
               al16(win_logits) && al16(class_map) && al16(avg) && al16(nodata_px), 0, 0, 0};
  // classes per pass: measured on B200 (tools/stitch_sweep.sh) 4 classes x 2 rows of float4 loads in flight per
  // thread beat wider passes (7-8 classes need > 128 registers and halve the resident warps)
  static const int env_ch = getenv("IG_STITCH_CH") ? atoi(getenv("IG_STITCH_CH")) : 0;  // tuning aid
  const int ch = env_ch > 0 ? std::min(env_ch, 8) : std::min(nc, 4);
  // rows per warp item: amortise the per-thread cover search, but keep >= ~4 waves of warps on the chip
  const int rows = y1 - y0;
  a.nseg = (W + SEG - 1) / SEG;
  const long long warp_slots = static_cast<long long>(ig_num_sms()) * 64;
  static const int env_rpi = getenv("IG_STITCH_RPI") ? atoi(getenv("IG_STITCH_RPI")) : 0;  // tuning aid
  int rpi = ch <= 2 ? 8 : 4;
  while (rpi > 4 && static_cast<long long>(a.nseg) * ((rows + rpi - 1) / rpi) < 4 * warp_slots) rpi >>= 1;
  if (env_rpi > 0) rpi = env_rpi;
  a.rpi = rpi;
  const long long nitems = static_cast<long long>(a.nseg) * ((rows + rpi - 1) / rpi);
  IG_REQUIRE(nitems < (1ll << 31) - WARPS, IG_ESHAPE, "ig_stitch: stripe too large");
  a.nitems = static_cast<int>(nitems);
  const int blocks = (a.nitems + WARPS - 1) / WARPS;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ig::ProfScope prof(ig::PROF_STITCH, st);
    static const int env_r = getenv("IG_STITCH_R") ? atoi(getenv("IG_STITCH_R")) : 0;  // tuning aid
    const int r = env_r > 0 ? env_r : (ch <= 2 ? 4 : (ch <= 4 ? 2 : 1));
#define IG_ST(CH_) \
    case CH_: \
      if (r >= 4) launch_stitch<CH_, 4>(a, blocks, st); \
      else if (r >= 2) launch_stitch<CH_, 2>(a, blocks, st); \
      else launch_stitch<CH_, 1>(a, blocks, st); \
      break;
    switch (ch) {
      IG_ST(1) IG_ST(2) IG_ST(3) IG_ST(4)
      case 5: launch_stitch<5, 1>(a, blocks, st); break;
      case 6: launch_stitch<6, 1>(a, blocks, st); break;
      case 7: launch_stitch<7, 1>(a, blocks, st); break;
      default: launch_stitch<8, 1>(a, blocks, st); break;
    }
#undef IG_ST
    IG_CUDA_OK(cudaGetLastError());
  }
  if (hist) {
    const int64_t n = static_cast<int64_t>(rows) * W;
    const int hb = static_cast<int>(std::min<int64_t>((n / 16 + 255) / 256 + 1, static_cast<int64_t>(ig_num_sms()) * 8));
    if (nc <= 4) class_hist_kernel<4><<<hb, 256, 0, st>>>(class_map, n, nc, hist);
    else if (nc <= 16) class_hist_kernel<16><<<hb, 256, 0, st>>>(class_map, n, nc, hist);
    else class_hist_kernel<MAX_NC><<<hb, 256, 0, st>>>(class_map, n, nc, hist);
    IG_CUDA_OK(cudaGetLastError());
  }
  return IG_OK;
}
