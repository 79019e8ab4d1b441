// Kernel 5 -- overlap-averaging stitch of sliding-window logits (HBM-bound, bit-exact).
//
// Specification: SURVEY.md Appendix A.6 (frozen in oracle/stitch.py).  The reference snapshot
// has no stitch; it contributes the window order (instageo/model/dataloader.py:655-664, top
// outer / left inner), the first-max-wins argmax -> int8 (instageo/model/infer_utils.py:96-101)
// and the nodata comparison (instageo/model/dataloader.py:899).
//
// Gather form: one thread owns 4 consecutive output pixels of ROWS consecutive rows and walks the
// <= ceil(win/stride)^2 windows covering them in row-major window order, so the float32 sum of every
// pixel is formed in exactly the order of the scatter-form oracle (0 + l1 + l2 ...) and the result is
// bit-identical.  The covering window ranges come from two binary searches over the sorted origins
// (done once per thread for x, once per row for y) instead of a scan of all origins per pixel, and
// when the 4 pixels sit 16-byte aligned inside every covering window (always true for strides and
// origins that are multiples of 4) each (window, class) read is one float4 -- a warp reads 512
// contiguous bytes per instruction and keeps nc x windows of them in flight.
// Class histogram: warp match + popc, one shared atomic per distinct class per warp, one global
// atomic per class per block.
// Algorithmic bytes per tile: n_win*nc*win^2*4 (logits read once) + H*W (map) [+ H*W nodata].
#include "ig_common.cuh"

namespace {

constexpr int MAX_AX = 256;   // window origins per axis held in shared memory
constexpr int MAX_NC = 32;
constexpr int PX = 4;         // pixels per thread along x
constexpr int ROWS = 8;       // rows per block
constexpr int THREADS = 256;

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
  unsigned long long* hist;
  int ptr_ok;  // every buffer 16-byte aligned: vector path allowed
};

// windows covering coordinate p: origins o with o <= p < o + win  ->  index range [lo, hi)
__device__ __forceinline__ void cover_range(const int* org, int n, int win, int p, int& lo, int& hi) {
  int a = 0, b = n;  // first index with org[i] + win > p
  while (a < b) {
    const int m = (a + b) >> 1;
    if (org[m] + win > p) b = m; else a = m + 1;
  }
  lo = a;
  a = lo; b = n;     // first index with org[i] > p
  while (a < b) {
    const int m = (a + b) >> 1;
    if (org[m] > p) b = m; else a = m + 1;
  }
  hi = a;
}

template <int NCM>
__global__ void __launch_bounds__(THREADS) stitch_kernel(const StitchArgs a) {
  __shared__ int s_xs[MAX_AX];
  __shared__ int s_ys[MAX_AX];
  __shared__ unsigned int s_hist[MAX_NC + 1];
  for (int i = threadIdx.x; i < a.nx; i += blockDim.x) s_xs[i] = a.xs[i];
  const int ny_s = a.ny < MAX_AX ? a.ny : MAX_AX;
  for (int i = threadIdx.x; i < ny_s; i += blockDim.x) s_ys[i] = a.ys[i];
  if (threadIdx.x <= a.nc && threadIdx.x <= MAX_NC) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int* ysrc = a.ny <= MAX_AX ? s_ys : a.ys;  // tall grids fall back to global (L1-cached) origins
  const int64_t plane = static_cast<int64_t>(a.win) * a.win;
  const int rows = a.y1 - a.y0;
  const int x0 = (blockIdx.x * THREADS + threadIdx.x) * PX;
  const bool live = x0 < a.W;

  int xlo = 0, xhi = 0;
  bool vec = false;
  if (live) {
    cover_range(s_xs, a.nx, a.win, x0, xlo, xhi);
    vec = a.ptr_ok && (x0 + PX <= a.W) && ((a.win & 3) == 0) && ((a.W & 3) == 0);
    if (vec) {  // same cover set for all 4 pixels, and 16-byte aligned inside every window?
      int l3, h3;
      cover_range(s_xs, a.nx, a.win, x0 + PX - 1, l3, h3);
      vec = (l3 == xlo) && (h3 == xhi);
      for (int ix = xlo; ix < xhi; ++ix) vec = vec && (((x0 - s_xs[ix]) & 3) == 0);
    }
  }

  int ylo = 0, yhi = 0;
  cover_range(ysrc, a.ny, a.win, a.y0 + blockIdx.y * ROWS, ylo, yhi);
  for (int ry = 0; ry < ROWS; ++ry) {
    const int y = a.y0 + blockIdx.y * ROWS + ry;
    if (y >= a.y1) break;
    // origins are sorted: the cover range only moves forward from one row to the next
    while (ylo < a.ny && ysrc[ylo] + a.win <= y) ++ylo;
    while (yhi < a.ny && ysrc[yhi] <= y) ++yhi;
    int cls4[PX];
#pragma unroll
    for (int p = 0; p < PX; ++p) cls4[p] = a.nodata_class;
    if (live) {
      float acc[PX][NCM];
      float cnt[PX];
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        cnt[p] = 0.f;
#pragma unroll
        for (int k = 0; k < NCM; ++k) acc[p][k] = 0.f;
      }
      if (vec) {
        for (int iy = ylo; iy < yhi; ++iy) {
          const int ly = y - ysrc[iy];
          for (int ix = xlo; ix < xhi; ++ix) {
            const int wi = iy * a.nx + ix - a.win_base;
            if (wi < 0 || wi >= a.n_win) continue;
            const float* p = a.logits + static_cast<int64_t>(wi) * a.nc * plane + ly * a.win + (x0 - s_xs[ix]);
#pragma unroll
            for (int k = 0; k < NCM; ++k)
              if (k < a.nc) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(p + k * plane));
                acc[0][k] = __fadd_rn(acc[0][k], v.x);
                acc[1][k] = __fadd_rn(acc[1][k], v.y);
                acc[2][k] = __fadd_rn(acc[2][k], v.z);
                acc[3][k] = __fadd_rn(acc[3][k], v.w);
              }
#pragma unroll
            for (int q = 0; q < PX; ++q) cnt[q] += 1.f;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < PX; ++q) {
          const int x = x0 + q;
          if (x >= a.W) continue;
          int lo, hi;
          cover_range(s_xs, a.nx, a.win, x, lo, hi);
          for (int iy = ylo; iy < yhi; ++iy) {
            const int ly = y - ysrc[iy];
            for (int ix = lo; ix < hi; ++ix) {
              const int wi = iy * a.nx + ix - a.win_base;
              if (wi < 0 || wi >= a.n_win) continue;
              const float* p = a.logits + static_cast<int64_t>(wi) * a.nc * plane + ly * a.win + (x - s_xs[ix]);
#pragma unroll
              for (int k = 0; k < NCM; ++k)
                if (k < a.nc) acc[q][k] = __fadd_rn(acc[q][k], __ldcs(p + k * plane));
              cnt[q] += 1.f;
            }
          }
        }
      }
      const int64_t orow = static_cast<int64_t>(y - a.y0) * a.W;
#pragma unroll
      for (int q = 0; q < PX; ++q) {
        const int x = x0 + q;
        if (x >= a.W) continue;
        const bool covered = cnt[q] > 0.f;
        // 1, 2, 4, ... covering windows: the true division is an exact scaling, x * (1/cnt) is
        // bit-identical and skips the IEEE division sequence (all interior pixels for win = k * stride)
        const int ci = static_cast<int>(cnt[q]);
        const bool pow2 = (ci & (ci - 1)) == 0;
        const float inv = pow2 && covered ? __frcp_rn(cnt[q]) : 0.f;
        float best = 0.f;
        int bi = 0;
#pragma unroll
        for (int k = 0; k < NCM; ++k)
          if (k < a.nc) {
            const float v = covered ? (pow2 ? __fmul_rn(acc[q][k], inv) : __fdiv_rn(acc[q][k], cnt[q])) : 0.f;
            acc[q][k] = v;
            if (k == 0 || v > best) {  // strict > : first maximum wins (torch.argmax)
              best = v;
              bi = k;
            }
          }
        cls4[q] = covered ? bi : a.nodata_class;
      }
      if (a.nodata_px) {
        const uint8_t* nd = a.nodata_px + static_cast<int64_t>(y) * a.W + x0;
        if (vec) {
          const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(nd));
#pragma unroll
          for (int q = 0; q < PX; ++q)
            if ((m >> (8 * q)) & 0xffu) cls4[q] = a.nodata_class;
        } else {
#pragma unroll
          for (int q = 0; q < PX; ++q)
            if (x0 + q < a.W && nd[q]) cls4[q] = a.nodata_class;
        }
      }
      if (vec) {
        const uint32_t packed = (static_cast<uint32_t>(cls4[0]) & 0xffu) | ((static_cast<uint32_t>(cls4[1]) & 0xffu) << 8) |
                                ((static_cast<uint32_t>(cls4[2]) & 0xffu) << 16) | ((static_cast<uint32_t>(cls4[3]) & 0xffu) << 24);
        *reinterpret_cast<uint32_t*>(a.cls + orow + x0) = packed;
        if (a.avg) {
#pragma unroll
          for (int k = 0; k < NCM; ++k)
            if (k < a.nc)
              *reinterpret_cast<float4*>(a.avg + (static_cast<int64_t>(k) * rows + (y - a.y0)) * a.W + x0) =
                  make_float4(acc[0][k], acc[1][k], acc[2][k], acc[3][k]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < PX; ++q) {
          const int x = x0 + q;
          if (x >= a.W) continue;
          a.cls[orow + x] = static_cast<int8_t>(cls4[q]);
          if (a.avg) {
#pragma unroll
            for (int k = 0; k < NCM; ++k)
              if (k < a.nc) a.avg[(static_cast<int64_t>(k) * rows + (y - a.y0)) * a.W + x] = acc[q][k];
          }
        }
      }
    }
    if (a.hist) {
#pragma unroll
      for (int q = 0; q < PX; ++q) {
        const bool mine = live && (x0 + q < a.W);
        int bin = cls4[q];
        if (bin < 0 || bin >= a.nc) bin = a.nc;  // nodata bin
        const unsigned active = __ballot_sync(0xffffffffu, mine);
        if (mine) {
          const unsigned peers = __match_any_sync(active, bin);
          if ((threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&s_hist[bin], __popc(peers));
        }
      }
    }
  }
  if (a.hist) {
    __syncthreads();
    if (threadIdx.x <= a.nc && s_hist[threadIdx.x])
      atomicAdd(a.hist + threadIdx.x, static_cast<unsigned long long>(s_hist[threadIdx.x]));
  }
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
  IG_REQUIRE((y1 - y0 + ROWS - 1) / ROWS <= 65535, IG_ESHAPE, "ig_stitch: stripe too tall");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  StitchArgs a{win_logits, n_win, win_base, nc, win, ys, xs, ny, nx, H, W, y0, y1,
               nodata_px, nodata_class, avg, class_map, hist,
               al16(win_logits) && al16(class_map) && al16(avg) && al16(nodata_px)};
  const int threads = THREADS;
  dim3 grid((W + threads * PX - 1) / (threads * PX), (y1 - y0 + ROWS - 1) / ROWS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ig::ProfScope prof(ig::PROF_STITCH, st);
  if (nc <= 2) stitch_kernel<2><<<grid, threads, 0, st>>>(a);
  else if (nc <= 4) stitch_kernel<4><<<grid, threads, 0, st>>>(a);
  else if (nc <= 16) stitch_kernel<16><<<grid, threads, 0, st>>>(a);
  else stitch_kernel<MAX_NC><<<grid, threads, 0, st>>>(a);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}
