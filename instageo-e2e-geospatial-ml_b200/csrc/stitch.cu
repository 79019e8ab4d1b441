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
#include <string.h>

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
      const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(a.nodata_px + static_cast<int64_t>(y + r - a.y0) * a.W + x0));
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
    if (a.nodata_px && a.nodata_px[static_cast<int64_t>(y - a.y0) * a.W + x]) c = a.nodata_class;
    a.cls[static_cast<int64_t>(y - a.y0) * a.W + x] = static_cast<int8_t>(c);
  }
}

template <int CH, int R>
__global__ void __launch_bounds__(THREADS) stitch_kernel(const __grid_constant__ StitchArgs a) {
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

// ------------------------------------------------------------------------------------------------
// TMA-staged tile path (the fast path).  Output tiles are 128 pixels x TR rows.  A producer warp walks
// the windows that INTERSECT a tile in row-major window order and pulls, per window, one
// [CH classes][TR rows][132 px] float32 box of its logits into a shared-memory ring with a single 4-D
// cp.async.bulk.tensor: the box is positioned at (tile - window origin), so whatever part of it falls
// outside the window (or past the last class) is zero-filled by the TMA unit and never fetched from
// DRAM.  Adding +0.0f is exact and the accumulators start at +0.0f, so the sums are still formed in
// the oracle's order, bit for bit.  Four consumer warps (thread = 4 px x RT rows) add the boxes up out
// of shared memory with conflict-free LDS.128.  Bytes in flight are bounded by shared memory, not by
// registers, and the load side costs one instruction per box instead of one per 512 bytes.
// One tile per CTA, launched in row-major tile order, so the CTAs resident at any moment sweep a band
// of adjacent tiles and the window rows they read are contiguous in DRAM.  (A persistent grid with the
// ring running across tile boundaries was measured and is slower -- 750 vs 470 us at nc = 13, stride
// 112: its loop-carried state costs 40 registers and a resident CTA per SM -- as is a column-major split.)
// TMA wants the box start 16-byte aligned in global memory: the box is 4 pixels wider than the tile
// and starts at the aligned pixel below; a window whose origin is not a multiple of 4 (the
// edge-aligned last window when W % 4 != 0) is read back with a 1-3 element shift.
// The per-pixel cover count is separable: cnt(x, y) = cx(x) * cy(y).
constexpr int T_CONS = 128;             // consumer threads
constexpr int T_THREADS = T_CONS + 32;  // + producer warp
constexpr int T_BARS = 16;              // ring slots supported by the barrier array
#ifndef IG_STITCH_PITCH
#define IG_STITCH_PITCH (SEG + 4)
#endif
constexpr int PITCH = IG_STITCH_PITCH;  // floats per staged row
constexpr int TILE_ROWS = 8;            // rows per tile (2 per consumer thread)

struct TileArgs {
  StitchArgs s;
  int vy0, vy1;     // window rows actually present in win_logits (win_base / n_win are whole rows)
  int stages;
  int vst;          // outputs / nodata may be accessed as 4-pixel vectors
  int nseg;         // 128-pixel column strips
  int ntiles;       // strips * row tiles, row-major (strip fastest)
};

// windows [lo, hi) of a sorted origin list that intersect the pixel range [p0, p1]
__device__ __forceinline__ void intersect_range(const int* org, int n, int win, int p0, int p1, int& lo, int& hi) {
  int dummy;
  cover_range(org, n, win, p0, lo, dummy);  // first origin with o + win > p0
  cover_range(org, n, win, p1, dummy, hi);  // first origin beyond p1
}

template <int CH, int RT>
__global__ void __launch_bounds__(T_THREADS) stitch_tma_kernel(const __grid_constant__ CUtensorMap tm,
                                                               const __grid_constant__ TileArgs ta) {
  constexpr int TR = 4 * RT;
  constexpr int STAGE_FLOATS = CH * TR * PITCH;
  extern __shared__ uint8_t smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw + ((128u - (ig::smem_u32(smem_raw) & 127u)) & 127u));
  __shared__ int s_xs[MAX_AX];
  __shared__ int s_ys[MAX_AX];
  __shared__ uint8_t s_cx[SEG];   // windows covering each pixel column of the tile
  __shared__ uint8_t s_cy[TR];    // windows (present in win_logits) covering each row of the tile
  __shared__ int s_rng[4];        // window index ranges that intersect the tile
  __shared__ __align__(8) uint64_t s_full[T_BARS];
  __shared__ __align__(8) uint64_t s_empty[T_BARS];
  const StitchArgs& a = ta.s;
  const int tid = threadIdx.x, warp = ig::warp_idx_uniform(), lane = tid & 31;
  const int ty = blockIdx.x / ta.nseg, seg = blockIdx.x - ty * ta.nseg;
  const int tile_x0 = seg * SEG;
  const int tile_y0 = a.y0 + ty * TR, tile_y1 = min(a.y1, tile_y0 + TR);

  for (int i = tid; i < a.nx; i += T_THREADS) s_xs[i] = a.xs[i];
  for (int i = tid; i < a.ny; i += T_THREADS) s_ys[i] = a.ys[i];
  if (tid == 0) {
    ig::tma_prefetch_desc(&tm);
    for (int s = 0; s < ta.stages; ++s) {
      ig::mbar_init(&s_full[s], 1);
      ig::mbar_init(&s_empty[s], 4);
    }
    ig::fence_barrier_init();
  }
  __syncthreads();
  if (tid < SEG) {
    const int x = tile_x0 + tid;
    int lo = 0, hi = 0;
    if (x < a.W) cover_range(s_xs, a.nx, a.win, x, lo, hi);
    s_cx[tid] = static_cast<uint8_t>(hi - lo);
  } else if (tid < SEG + TR) {
    const int y = tile_y0 + tid - SEG;
    int lo = 0, hi = 0;
    if (y < tile_y1) cover_range(s_ys, a.ny, a.win, y, lo, hi);
    lo = max(lo, ta.vy0);
    hi = min(hi, ta.vy1);
    s_cy[tid - SEG] = static_cast<uint8_t>(hi > lo ? hi - lo : 0);
  } else if (tid == SEG + TR) {
    int lo, hi;
    intersect_range(s_xs, a.nx, a.win, tile_x0, min(tile_x0 + SEG, a.W) - 1, lo, hi);
    s_rng[0] = lo;
    s_rng[1] = hi;
    intersect_range(s_ys, a.ny, a.win, tile_y0, tile_y1 - 1, lo, hi);
    s_rng[2] = max(lo, ta.vy0);
    s_rng[3] = min(hi, ta.vy1);
  }
  __syncthreads();
  const int txlo = s_rng[0], txhi = s_rng[1], tylo = s_rng[2], tyhi = s_rng[3];

  if (warp == 4) {
    // ===================== producer: one 4-D TMA box per (class pass, window) =====================
    if (ig::elect_one()) {
      int st = 0;
      uint32_t ph = 1;  // parity of a never-completed phase: the first pass over the ring does not wait
      for (int c0 = 0; c0 < a.nc; c0 += CH)
        for (int iy = tylo; iy < tyhi; ++iy)
          for (int ix = txlo; ix < txhi; ++ix) {
            ig::mbar_wait(&s_empty[st], ph);
            ig::mbar_expect_tx(&s_full[st], STAGE_FLOATS * 4);
            ig::tma_load_4d(ring + static_cast<size_t>(st) * STAGE_FLOATS, &tm, &s_full[st],
                            (tile_x0 - s_xs[ix]) & ~3, tile_y0 - s_ys[iy], c0, iy * a.nx + ix - a.win_base);
            if (++st == ta.stages) {
              st = 0;
              ph ^= 1;
            }
          }
    }
    return;
  }

  // ===================== consumers: thread = 4 px x RT rows =====================
  const int x0 = tile_x0 + lane * PX;
  const int r0 = warp * RT;  // first tile row of this thread
  const int rows = a.y1 - a.y0;
  int cx[PX];
#pragma unroll
  for (int q = 0; q < PX; ++q) cx[q] = s_cx[lane * PX + q];
  const bool uni = cx[0] == cx[1] && cx[1] == cx[2] && cx[2] == cx[3];
  uint32_t ndw[RT];  // nodata bytes of the thread's pixels, fetched before the ring is consumed
#pragma unroll
  for (int j = 0; j < RT; ++j) {
    const int y = tile_y0 + r0 + j;
    ndw[j] = 0;
    if (a.nodata_px && y < tile_y1 && x0 < a.W) {
      const uint8_t* nd = a.nodata_px + static_cast<int64_t>(y - a.y0) * a.W + x0;
      if (ta.vst) ndw[j] = __ldg(reinterpret_cast<const uint32_t*>(nd));
      else
        for (int q = 0; q < PX; ++q)
          if (x0 + q < a.W && nd[q]) ndw[j] |= 0xffu << (8 * q);
    }
  }
  float best[RT][PX];
  int bi[RT][PX];
#pragma unroll
  for (int j = 0; j < RT; ++j)
#pragma unroll
    for (int q = 0; q < PX; ++q) {
      best[j][q] = 0.f;
      bi[j][q] = 0;
    }
  int st = 0;
  uint32_t ph = 0;
  for (int c0 = 0; c0 < a.nc; c0 += CH) {
    float4 acc[RT][CH];
#pragma unroll
    for (int j = 0; j < RT; ++j)
#pragma unroll
      for (int k = 0; k < CH; ++k) acc[j][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = tylo; iy < tyhi; ++iy)
      for (int ix = txlo; ix < txhi; ++ix) {
        ig::mbar_wait(&s_full[st], ph);
        const int sh = (tile_x0 - s_xs[ix]) & 3;  // 0 unless the window origin is not a multiple of 4
        const float* sp = ring + static_cast<size_t>(st) * STAGE_FLOATS + r0 * PITCH + lane * PX;
        if (sh == 0) {
#pragma unroll
          for (int j = 0; j < RT; ++j)
#pragma unroll
            for (int k = 0; k < CH; ++k) {
              const float4 v = *reinterpret_cast<const float4*>(sp + (k * TR + j) * PITCH);
              acc[j][k].x = __fadd_rn(acc[j][k].x, v.x);
              acc[j][k].y = __fadd_rn(acc[j][k].y, v.y);
              acc[j][k].z = __fadd_rn(acc[j][k].z, v.z);
              acc[j][k].w = __fadd_rn(acc[j][k].w, v.w);
            }
        } else {
#pragma unroll
          for (int j = 0; j < RT; ++j)
#pragma unroll
            for (int k = 0; k < CH; ++k) {
              const float* p = sp + (k * TR + j) * PITCH + sh;
              acc[j][k].x = __fadd_rn(acc[j][k].x, p[0]);
              acc[j][k].y = __fadd_rn(acc[j][k].y, p[1]);
              acc[j][k].z = __fadd_rn(acc[j][k].z, p[2]);
              acc[j][k].w = __fadd_rn(acc[j][k].w, p[3]);
            }
        }
        __syncwarp();
        if (lane == 0) ig::mbar_arrive(&s_empty[st]);  // the slot has been read: hand it back to the producer
        if (++st == ta.stages) {
          st = 0;
          ph ^= 1;
        }
      }
    // ---- this pass's classes: divide by the cover count, running first-max argmax, optional avg
#pragma unroll
    for (int j = 0; j < RT; ++j) {
      const int y = tile_y0 + r0 + j;
      const int cy = s_cy[r0 + j];
      const int cnt0 = cx[0] * cy;
      if (uni && (cnt0 & (cnt0 - 1)) == 0) {
        // 1, 2, 4, ... covering windows (or none: 0 * 0): exact scaling by the reciprocal
        const float inv = cnt0 > 0 ? __frcp_rn(static_cast<float>(cnt0)) : 0.f;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          acc[j][k].x = __fmul_rn(acc[j][k].x, inv);
          acc[j][k].y = __fmul_rn(acc[j][k].y, inv);
          acc[j][k].z = __fmul_rn(acc[j][k].z, inv);
          acc[j][k].w = __fmul_rn(acc[j][k].w, inv);
        }
      } else {
        float cf[PX];
#pragma unroll
        for (int q = 0; q < PX; ++q) cf[q] = static_cast<float>(cx[q] * cy);
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          acc[j][k].x = cf[0] > 0.f ? __fdiv_rn(acc[j][k].x, cf[0]) : 0.f;
          acc[j][k].y = cf[1] > 0.f ? __fdiv_rn(acc[j][k].y, cf[1]) : 0.f;
          acc[j][k].z = cf[2] > 0.f ? __fdiv_rn(acc[j][k].z, cf[2]) : 0.f;
          acc[j][k].w = cf[3] > 0.f ? __fdiv_rn(acc[j][k].w, cf[3]) : 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        if (c0 + k < a.nc) {
          const float vv[PX] = {acc[j][k].x, acc[j][k].y, acc[j][k].z, acc[j][k].w};
#pragma unroll
          for (int q = 0; q < PX; ++q)
            if (c0 + k == 0 || vv[q] > best[j][q]) {  // strict > : first maximum wins (torch.argmax)
              best[j][q] = vv[q];
              bi[j][q] = c0 + k;
            }
          if (a.avg && y < tile_y1 && x0 < a.W) {
            float* ap = a.avg + (static_cast<int64_t>(c0 + k) * rows + (y - a.y0)) * a.W + x0;
            if (ta.vst) __stcs(reinterpret_cast<float4*>(ap), acc[j][k]);
            else {
#pragma unroll
              for (int q = 0; q < PX; ++q)
                if (x0 + q < a.W) ap[q] = vv[q];
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < RT; ++j) {
    const int y = tile_y0 + r0 + j;
    if (y >= tile_y1 || x0 >= a.W) continue;
    const int cy = s_cy[r0 + j];
    int c4[PX];
#pragma unroll
    for (int q = 0; q < PX; ++q) {
      c4[q] = cx[q] * cy > 0 ? bi[j][q] : a.nodata_class;
      if ((ndw[j] >> (8 * q)) & 0xffu) c4[q] = a.nodata_class;
    }
    int8_t* cp = a.cls + static_cast<int64_t>(y - a.y0) * a.W + x0;
    if (ta.vst) *reinterpret_cast<uint32_t*>(cp) = pack4(c4);
    else {
#pragma unroll
      for (int q = 0; q < PX; ++q)
        if (x0 + q < a.W) cp[q] = static_cast<int8_t>(c4[q]);
    }
  }
}

template <int CH, int RT>
int launch_tma(const CUtensorMap& tm, TileArgs ta, cudaStream_t st) {
  constexpr int TR = 4 * RT;
  const size_t stage = static_cast<size_t>(CH) * TR * PITCH * 4;
  const size_t smem = 128 + stage * ta.stages;
  static IgPerDevice configured = {};
  if (static_cast<int>(smem) > configured.get()) {
    IG_CUDA_OK(cudaFuncSetAttribute(stitch_tma_kernel<CH, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured.set(static_cast<int>(smem));
  }
  ta.nseg = (ta.s.W + SEG - 1) / SEG;
  const long long ntiles = static_cast<long long>(ta.nseg) * ((ta.s.y1 - ta.s.y0 + TR - 1) / TR);
  IG_REQUIRE(ntiles < (1ll << 31), IG_ESHAPE, "ig_stitch: stripe too large");
  ta.ntiles = static_cast<int>(ntiles);
  const int grid = ta.ntiles;  // one tile per CTA, row-major: the resident CTAs sweep a band of adjacent tiles
  stitch_tma_kernel<CH, RT><<<grid, T_THREADS, smem, st>>>(tm, ta);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
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
  const int rows = y1 - y0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // IG_STITCH_PATH=tma|direct forces one path (the tests run both on the same inputs); default: by nc.
  const char* env_path = getenv("IG_STITCH_PATH");
  const bool force_tma = env_path && !strcmp(env_path, "tma"), force_direct = env_path && !strcmp(env_path, "direct");

  // ---- TMA-staged tiles need 16-byte aligned window rows and win_logits holding whole window rows of the
  // grid (what row-stripe sharding passes).  Measured on the 3660^2 tile (B200, profiles/r01_stitch_sweep.txt):
  // with >= 3 classes the boxes are big enough to run at 3.4-5.7 TB/s; at nc <= 2 a tile carries ~20 KB and
  // the per-CTA setup dominates, the direct-load kernel is faster there (59 vs 80 us).
  const bool tma_ok = (win % 4 == 0) && al16(win_logits) && ny <= MAX_AX && (win_base % nx == 0) &&
                      (n_win % nx == 0) && n_win > 0;
  if (tma_ok && !force_direct && (nc > 2 || force_tma)) {
    TileArgs ta;
    ta.s = a;
    ta.vy0 = win_base / nx;
    ta.vy1 = ta.vy0 + n_win / nx;
    ta.vst = (W % 4 == 0) && al16(class_map) && al16(avg) && al16(nodata_px);
    const int passes = (nc + 7) / 8;           // as few class passes as possible, <= 8 classes each
    const int ch = (nc + passes - 1) / passes;
    const size_t stage = static_cast<size_t>(ch) * TILE_ROWS * PITCH * 4;
    ta.stages = std::max(2, std::min(static_cast<int>((72 * 1024) / stage), T_BARS));  // <= 72 KB ring: 3 CTAs/SM
    CUtensorMap tm;
    const uint64_t dims[4] = {static_cast<uint64_t>(win), static_cast<uint64_t>(win), static_cast<uint64_t>(nc),
                              static_cast<uint64_t>(n_win)};
    const uint64_t plane_b = static_cast<uint64_t>(win) * win * 4;
    const uint64_t strides[3] = {static_cast<uint64_t>(win) * 4, plane_b, plane_b * nc};
    const uint32_t box[4] = {static_cast<uint32_t>(PITCH), static_cast<uint32_t>(TILE_ROWS), static_cast<uint32_t>(ch), 1u};
    IG_TRY(ig_make_tmap_nd(&tm, IG_F32, win_logits, 4, dims, strides, box));
    ig::ProfScope prof(ig::PROF_STITCH, st);
    int rc = IG_OK;
    switch (ch) {
      case 1: rc = launch_tma<1, TILE_ROWS / 4>(tm, ta, st); break;
      case 2: rc = launch_tma<2, TILE_ROWS / 4>(tm, ta, st); break;
      case 3: rc = launch_tma<3, TILE_ROWS / 4>(tm, ta, st); break;
      case 4: rc = launch_tma<4, TILE_ROWS / 4>(tm, ta, st); break;
      case 5: rc = launch_tma<5, TILE_ROWS / 4>(tm, ta, st); break;
      case 6: rc = launch_tma<6, TILE_ROWS / 4>(tm, ta, st); break;
      case 7: rc = launch_tma<7, TILE_ROWS / 4>(tm, ta, st); break;
      default: rc = launch_tma<8, TILE_ROWS / 4>(tm, ta, st); break;
    }
    IG_TRY(rc);
  } else {
    // ---- direct loads: unaligned windows, partial window rows, very tall grids, and nc <= 2.
    // 4 classes x 2 rows (or 2 classes x 4 rows) of float4 loads in flight per thread; wider passes need
    // > 128 registers and halve the resident warps (tools/stitch_sweep.sh).
    const int ch = std::min(nc, 4);
    // rows per warp item: amortise the per-thread cover search, but keep at least one full wave of warps
    a.nseg = (W + SEG - 1) / SEG;
    const long long warp_slots = static_cast<long long>(ig_num_sms()) * 64;
    int rpi = ch <= 2 ? 8 : 4;
    while (rpi > 4 && static_cast<long long>(a.nseg) * ((rows + rpi - 1) / rpi) < warp_slots) rpi >>= 1;
    a.rpi = rpi;
    const long long nitems = static_cast<long long>(a.nseg) * ((rows + rpi - 1) / rpi);
    IG_REQUIRE(nitems < (1ll << 31) - WARPS, IG_ESHAPE, "ig_stitch: stripe too large");
    a.nitems = static_cast<int>(nitems);
    const int blocks = (a.nitems + WARPS - 1) / WARPS;
    ig::ProfScope prof(ig::PROF_STITCH, st);
    switch (ch) {
      case 1: launch_stitch<1, 4>(a, blocks, st); break;
      case 2: launch_stitch<2, 4>(a, blocks, st); break;
      case 3: launch_stitch<3, 2>(a, blocks, st); break;
      default: launch_stitch<4, 2>(a, blocks, st); break;
    }
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
