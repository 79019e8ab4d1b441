// Kernel 5 -- overlap-averaging stitch of sliding-window logits (HBM-bound, bit-exact).
//
// Specification: SURVEY.md Appendix A.6 (frozen in oracle/stitch.py).  The reference snapshot
// has no stitch; it contributes the window order (instageo/model/dataloader.py:655-664, top
// outer / left inner), the first-max-wins argmax -> int8 (instageo/model/infer_utils.py:96-101)
// and the nodata comparison (instageo/model/dataloader.py:899).
//
// Gather form: one thread per output pixel walks the <= ceil(win/stride)^2 windows covering
// it in row-major window order, so the float32 sum is formed in exactly the order of the
// scatter-form oracle (0 + l1 + l2 ...) and the result is bit-identical.  Consecutive
// threads are consecutive x, so every (window, class) read is a coalesced 128-byte line.
// Class histogram: per-class warp ballots + popc, one shared atomic per warp and one global
// atomic per block (warp-level reduction instead of per-pixel atomics).
// Algorithmic bytes per tile: n_win*nc*win^2*4 (logits read once) + H*W (map) [+ H*W nodata].
#include "ig_common.cuh"

namespace {

constexpr int MAX_AX = 256;   // window origins per axis
constexpr int MAX_NC = 32;

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
};

template <int NCM>
__global__ void __launch_bounds__(256) stitch_kernel(const StitchArgs a) {
  __shared__ int s_xs[MAX_AX];
  __shared__ int s_iy0, s_iy1;
  __shared__ unsigned int s_hist[MAX_NC + 1];
  const int y = a.y0 + blockIdx.y;
  for (int i = threadIdx.x; i < a.nx; i += blockDim.x) s_xs[i] = a.xs[i];
  if (threadIdx.x <= a.nc && threadIdx.x <= MAX_NC) s_hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    s_iy0 = a.ny;
    s_iy1 = 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.ny; i += blockDim.x) {  // parallel cover test, no serial global loads
    const int o = a.ys[i];
    if (o <= y && y < o + a.win) {
      atomicMin(&s_iy0, i);
      atomicMax(&s_iy1, i + 1);
    }
  }
  __syncthreads();
  const int iy0 = s_iy0, iy1 = s_iy1;
  const int64_t plane = static_cast<int64_t>(a.win) * a.win;
  const int rows = a.y1 - a.y0;

  for (int xb = blockIdx.x * blockDim.x; xb < a.W; xb += gridDim.x * blockDim.x) {
    const int x = xb + threadIdx.x;
    int cls = a.nodata_class;
    if (x < a.W) {
      float acc[NCM];
#pragma unroll
      for (int k = 0; k < NCM; ++k) acc[k] = 0.f;
      float cnt = 0.f;
      for (int iy = iy0; iy < iy1; ++iy) {
        const int ly = y - a.ys[iy];
        if (ly < 0 || ly >= a.win) continue;
        for (int ix = 0; ix < a.nx; ++ix) {
          const int lx = x - s_xs[ix];
          if (lx < 0 || lx >= a.win) continue;
          const int wi = iy * a.nx + ix - a.win_base;
          if (wi < 0 || wi >= a.n_win) continue;
          const float* p = a.logits + static_cast<int64_t>(wi) * a.nc * plane + ly * a.win + lx;
#pragma unroll
          for (int k = 0; k < NCM; ++k)
            if (k < a.nc) acc[k] = __fadd_rn(acc[k], __ldcs(p + k * plane));
          cnt += 1.f;
        }
      }
      const bool covered = cnt > 0.f;
      float best = 0.f;
      int bi = 0;
#pragma unroll
      for (int k = 0; k < NCM; ++k)
        if (k < a.nc) {
          const float v = covered ? __fdiv_rn(acc[k], cnt) : 0.f;
          if (a.avg) a.avg[(static_cast<int64_t>(k) * rows + (y - a.y0)) * a.W + x] = v;
          if (k == 0 || v > best) {  // strict > : first maximum wins (torch.argmax)
            best = v;
            bi = k;
          }
        }
      cls = covered ? bi : a.nodata_class;
      if (a.nodata_px && a.nodata_px[static_cast<int64_t>(y) * a.W + x]) cls = a.nodata_class;
      a.cls[static_cast<int64_t>(y - a.y0) * a.W + x] = static_cast<int8_t>(cls);
    }
    if (a.hist) {
      const bool live = x < a.W;
      for (int k = 0; k <= a.nc; ++k) {
        const bool mine = live && (k < a.nc ? (cls == k) : (cls < 0 || cls >= a.nc));
        const unsigned b = __ballot_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_hist[k], __popc(b));
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
  IG_REQUIRE(win_logits && ys && xs && class_map, IG_EINVAL, "ig_stitch: null pointer");
  IG_REQUIRE(nc >= 1 && nc <= MAX_NC, IG_ESHAPE, "ig_stitch: nc=%d unsupported (1..%d)", nc, MAX_NC);
  IG_REQUIRE(nx >= 1 && nx <= MAX_AX && ny >= 1 && ny <= 65535, IG_ESHAPE,
             "ig_stitch: window grid %dx%d unsupported", ny, nx);
  IG_REQUIRE(0 <= y0 && y0 <= y1 && y1 <= H && W >= 1 && win >= 1, IG_ESHAPE,
             "ig_stitch: bad stripe [%d,%d) of H=%d", y0, y1, H);
  if (y1 == y0) return IG_OK;
  IG_REQUIRE(y1 - y0 <= 65535, IG_ESHAPE, "ig_stitch: stripe too tall");
  StitchArgs a{win_logits, n_win, win_base, nc, win, ys, xs, ny, nx, H, W, y0, y1,
               nodata_px, nodata_class, avg, class_map, hist};
  const int threads = 256;
  dim3 grid((W + threads - 1) / threads, y1 - y0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ig::ProfScope prof(ig::PROF_STITCH, st);
  if (nc <= 2) stitch_kernel<2><<<grid, threads, 0, st>>>(a);
  else if (nc <= 4) stitch_kernel<4><<<grid, threads, 0, st>>>(a);
  else if (nc <= 16) stitch_kernel<16><<<grid, threads, 0, st>>>(a);
  else stitch_kernel<MAX_NC><<<grid, threads, 0, st>>>(a);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}
