// Eval-mode streaming metrics on the device -- HBM-bound counting kernels.
//
// The reference updates its metric objects on the HOST every step: logits -> argmax / softmax on the
// device, boolean gather of the non-ignored pixels, three D2H copies (labels, preds, [n, nc] probabilities),
// then np.bincount(y_true * k + y_pred) and a Python loop over classes with np.add.at
// (instageo/model/segmentation.py:117-156, instageo/model/metrics.py:86-108, :214-244, :330-356).
// Here one pass over the logits (or over the int8 class map the head epilogue already wrote) accumulates
//   * the k x k confusion matrix                          (RunningConfusionMatrix.update)
//   * the one-vs-rest positive / negative score histograms (RunningAUC.update, softmax fused)
// into device-resident uint64 counters; nothing but k*k + 2*k*n_bins integers ever leaves the GPU
// (n_pos / n_neg of the reference are the row sums of the two histograms).
//
// Layout: logits f32 [n_img, nc, hw] (NCHW), labels [n_img, hw] (int64 as the reference's labels.long(),
// or int32 / uint8 / int8), one thread = 4 consecutive pixels, float4 loads per class plane, all nc
// planes requested before the first use.  Counters live in shared memory (u32) per block and are
// flushed with one 64-bit global atomic per non-zero entry.  Softmax scores pile up in bin 0 (negatives)
// and bin n_bins-1 (positives): those two bins per class, and the sample counts, are counted in registers;
// confusion-matrix cells (few keys, always contended) are merged across the warp with match.any.
#include "ig_common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int MAX_K = 16;  // = gemm::NCP, the widest class head the model kernels support

struct SegArgs {
  const float* logits;    // [n_img, nc, hw] or nullptr (then pred is given)
  const int8_t* pred;     // [n_img * hw]
  const void* labels;
  long long n_img, hw;
  int nc, k, label_dtype;
  int has_ignore;
  long long ignore;
  unsigned long long* matrix;    // [k*k] or nullptr
  unsigned long long* counters;  // [0] valid samples, [1] labels / predictions outside [0, k)
  int n_bins;
  float lo, hi;
  unsigned long long *pos_hist, *neg_hist;  // AUC [nc, n_bins] each, or pos_hist == nullptr
};

__device__ __forceinline__ long long load_label(const void* p, int dtype, long long i) {
  switch (dtype) {
    case IG_I64: return static_cast<const long long*>(p)[i];
    case IG_I32: return static_cast<const int*>(p)[i];
    case IG_U8: return static_cast<const uint8_t*>(p)[i];
    default: return static_cast<const int8_t*>(p)[i];
  }
}

// four consecutive labels, vectorised when the element type allows it (i is a multiple of 4)
__device__ __forceinline__ void load_label4(const void* p, int dtype, long long i, long long* out) {
  if (dtype == IG_I64) {
    const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(static_cast<const long long*>(p) + i));
    const longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(static_cast<const long long*>(p) + i) + 1);
    out[0] = a.x, out[1] = a.y, out[2] = b.x, out[3] = b.y;
  } else if (dtype == IG_I32) {
    const int4 a = __ldcs(reinterpret_cast<const int4*>(static_cast<const int*>(p) + i));
    out[0] = a.x, out[1] = a.y, out[2] = a.z, out[3] = a.w;
  } else {
    const uint32_t w = __ldcs(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p) + i));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t b = (w >> (8 * j)) & 0xffu;
      out[j] = dtype == IG_U8 ? static_cast<long long>(b) : static_cast<long long>(static_cast<int8_t>(b));
    }
  }
}

// add 1 to smem[key] for every active lane whose `on` is true; lanes with the same key are merged
__device__ __forceinline__ void warp_count(uint32_t* smem, uint32_t key, bool on) {
  const uint32_t active = __ballot_sync(0xffffffffu, on);
  if (!on) return;
  const uint32_t peers = __match_any_sync(active, key);
  if ((__ffs(peers) - 1) == static_cast<int>(threadIdx.x & 31)) atomicAdd(smem + key, __popc(peers));
}

// RunningAUC._bin in float32 (numpy >= 2 scalar rules, see oracle/metrics.py:auc_bins): clamp, then
// trunc((s - lo) / (hi - lo) * (n_bins - 1)) with IEEE sub / div / mul; NaN -> 0.
__device__ __forceinline__ int score_bin(float s, float lo, float hi, float range, float nbm1, int n_bins) {
  if (!(s > lo)) return 0;
  if (s >= hi) return n_bins - 1;
  // x / 1.0f == x: skip the division for the default [0, 1] range (a denormal score would take its slow path)
  const float num = __fsub_rn(s, lo);
  const float x = __fmul_rn(range == 1.0f ? num : __fdiv_rn(num, range), nbm1);
  return static_cast<int>(x);
}

// NC = compile-time class count of the logits (0: no logits, predictions are given)
template <int NC>
__global__ void __launch_bounds__(THREADS) seg_metrics_kernel(const SegArgs a) {
  extern __shared__ uint32_t sm[];
  const bool want_auc = NC > 0 && a.pos_hist != nullptr;
  const int kk = a.k * a.k;
  // [kk confusion][2 valid/invalid][NC * n_bins pos][NC * n_bins neg]
  uint32_t* s_conf = sm;
  uint32_t* s_cnt = sm + kk;
  uint32_t* s_pos = s_cnt + 2;
  uint32_t* s_neg = s_pos + (want_auc ? NC * a.n_bins : 0);
  const int n_sm = kk + 2 + (want_auc ? 2 * NC * a.n_bins : 0);
  for (int i = threadIdx.x; i < n_sm; i += THREADS) sm[i] = 0;
  __syncthreads();

  const long long quads_per_img = a.hw >> 2;
  const long long total_quads = a.n_img * quads_per_img;
  const float range = a.hi - a.lo, nbm1 = static_cast<float>(a.n_bins - 1);
  // a block's shared counters are u32: flush before any of them can wrap (every pixel adds at most 1 to a bin)
  // ... and the 16-bit per-thread fields hold 4 pixels per iteration
  const long long flush_every = 8192;
  long long iters = 0;
  // Per-thread counters for what nearly every pixel hits: the valid / out-of-range sample counts and, for
  // the ROC histograms, the first and the last bin of every class (softmax scores pile up at 0 and 1), kept
  // as two 16-bit fields per register and warp-reduced into shared memory at flush time.  Every other bin
  // takes a plain shared atomic (rarely contended: ~1 lane per clock per SM is all ATOMS delivers).
  uint32_t n_ok = 0, n_bad = 0;
  uint32_t hot_pos[NC > 0 ? NC : 1], hot_neg[NC > 0 ? NC : 1];  // [15:0] first bin, [31:16] last bin
#pragma unroll
  for (int c = 0; c < (NC > 0 ? NC : 1); ++c) hot_pos[c] = hot_neg[c] = 0;
  auto flush = [&]() {
    {
      const int lane = threadIdx.x & 31;
      uint32_t t0 = n_ok, t1 = n_bad;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t0 += __shfl_xor_sync(0xffffffffu, t0, o), t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      if (lane == 0 && t0) atomicAdd(s_cnt, t0);
      if (lane == 0 && t1) atomicAdd(s_cnt + 1, t1);
      n_ok = n_bad = 0;
      if (want_auc) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t u[4] = {hot_pos[c] & 0xffffu, hot_pos[c] >> 16, hot_neg[c] & 0xffffu, hot_neg[c] >> 16};
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int f = 0; f < 4; ++f) u[f] += __shfl_xor_sync(0xffffffffu, u[f], o);
          if (lane == 0) {
            if (u[0]) atomicAdd(s_pos + c * a.n_bins, u[0]);
            if (u[1]) atomicAdd(s_pos + c * a.n_bins + (a.n_bins - 1), u[1]);
            if (u[2]) atomicAdd(s_neg + c * a.n_bins, u[2]);
            if (u[3]) atomicAdd(s_neg + c * a.n_bins + (a.n_bins - 1), u[3]);
          }
          hot_pos[c] = hot_neg[c] = 0;
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_sm; i += THREADS) {
      const uint32_t v = sm[i];
      if (v == 0) continue;
      sm[i] = 0;
      if (i < kk) {
        if (a.matrix) atomicAdd(a.matrix + i, static_cast<unsigned long long>(v));
      } else if (i < kk + 2) {
        atomicAdd(a.counters + (i - kk), static_cast<unsigned long long>(v));
      } else {
        const int j = i - kk - 2;
        const bool is_pos = j < NC * a.n_bins;
        const int jj = is_pos ? j : j - NC * a.n_bins;
        atomicAdd((is_pos ? a.pos_hist : a.neg_hist) + jj, static_cast<unsigned long long>(v));
      }
    }
    __syncthreads();
  };

  for (long long q0 = static_cast<long long>(blockIdx.x) * THREADS; q0 < total_quads;
       q0 += static_cast<long long>(gridDim.x) * THREADS) {
    const long long q = q0 + threadIdx.x;
    const bool live = q < total_quads;
    long long lab[4] = {0, 0, 0, 0};
    int prd[4] = {0, 0, 0, 0};
    float v[NC > 0 ? NC : 1][4];
    if (live) {
      const long long img = q / quads_per_img, px = (q - img * quads_per_img) << 2;
      load_label4(a.labels, a.label_dtype, img * a.hw + px, lab);
      if (NC > 0) {
        const float* base = a.logits + (img * NC) * a.hw + px;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float4 t = __ldcs(reinterpret_cast<const float4*>(base + c * a.hw));
          v[c][0] = t.x, v[c][1] = t.y, v[c][2] = t.z, v[c][3] = t.w;
        }
      } else {
        const uint32_t w = __ldcs(reinterpret_cast<const uint32_t*>(a.pred + img * a.hw + px));
#pragma unroll
        for (int j = 0; j < 4; ++j) prd[j] = static_cast<int8_t>((w >> (8 * j)) & 0xffu);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool use = live && !(a.has_ignore && lab[j] == a.ignore);
      float prob[NC > 0 ? NC : 1];
      if (NC > 0) {
        // torch.argmax (first maximum wins) and torch.softmax in float32: exp(x - max) / sum, class order
        float m = v[0][j];
        int bi = 0;
#pragma unroll
        for (int c = 1; c < NC; ++c)
          if (v[c][j] > m) m = v[c][j], bi = c;
        prd[j] = bi;
        if (want_auc) {
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            prob[c] = expf(v[c][j] - m);
            s += prob[c];
          }
          // e / s through the reciprocal and one FMA correction step (s is in [1, NC]): within 1 ulp of the
          // true quotient like the exponentials feeding it, and free of __fdiv_rn's slow path, which every
          // denormal numerator (confident models: exp(-90)) would take
          const float r = __frcp_rn(s);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float q = prob[c] * r;
            prob[c] = fmaf(fmaf(-q, s, prob[c]), r, q);
          }
        }
      }
      const bool in_range = lab[j] >= 0 && lab[j] < a.k && prd[j] >= 0 && prd[j] < a.k;
      warp_count(s_conf, static_cast<uint32_t>(in_range ? lab[j] * a.k + prd[j] : 0), use && in_range);
      n_ok += use && in_range;
      n_bad += use && !in_range;
      if (want_auc) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int b = score_bin(prob[c], a.lo, a.hi, range, nbm1, a.n_bins);
          const bool is_pos = lab[j] == c;
          const uint32_t inc = !use ? 0u : (b == 0 ? 1u : (b == a.n_bins - 1 ? 0x10000u : 0u));
          hot_pos[c] += is_pos ? inc : 0u;
          hot_neg[c] += is_pos ? 0u : inc;
          if (use && inc == 0u) atomicAdd((is_pos ? s_pos : s_neg) + c * a.n_bins + b, 1u);
        }
      }
    }
    if (++iters == flush_every) {
      flush();
      iters = 0;
    }
  }
  flush();
}

// scalar tail / generic path for hw % 4 != 0 is not needed: callers pad nothing -- the Python mirror and
// the C ABI require hw % 4 == 0 for the fused path and fall back to n_img = 1, hw = n (any n % 4 == 0) or
// the element kernel below for ragged sizes.
__global__ void __launch_bounds__(THREADS) confusion_elem_kernel(const SegArgs a, long long start, long long n) {
  extern __shared__ uint32_t sm[];
  const int kk = a.k * a.k;
  for (int i = threadIdx.x; i < kk + 2; i += THREADS) sm[i] = 0;
  __syncthreads();
  for (long long i0 = static_cast<long long>(blockIdx.x) * THREADS; i0 < n;
       i0 += static_cast<long long>(gridDim.x) * THREADS) {
    const long long i = i0 + threadIdx.x;
    const bool live = i < n;
    const long long lab = live ? load_label(a.labels, a.label_dtype, start + i) : 0;
    const int prd = live ? a.pred[start + i] : 0;
    const bool use = live && !(a.has_ignore && lab == a.ignore);
    const bool in_range = lab >= 0 && lab < a.k && prd >= 0 && prd < a.k;
    warp_count(sm, static_cast<uint32_t>(in_range ? lab * a.k + prd : 0), use && in_range);
    if (use) atomicAdd(sm + kk + (in_range ? 0 : 1), 1u);  // ragged tails only: a few elements
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kk + 2; i += THREADS) {
    const uint32_t v = sm[i];
    if (v == 0) continue;
    if (i < kk) atomicAdd(a.matrix + i, static_cast<unsigned long long>(v));
    else atomicAdd(a.counters + (i - kk), static_cast<unsigned long long>(v));
  }
}

// RunningAUC.update on given probabilities, y_score [n, k] row-major (float32 or float64 -- the reference
// bins in the scores' own precision), one thread per sample, global histograms through shared memory.
template <typename T>
__global__ void __launch_bounds__(THREADS) auc_scores_kernel(const T* __restrict__ scores, const void* labels,
                                                             int label_dtype, long long n, int k, int n_bins, T lo,
                                                             T hi, unsigned long long* pos_hist,
                                                             unsigned long long* neg_hist, int use_smem) {
  extern __shared__ uint32_t sm[];
  const int nh = 2 * k * n_bins;
  if (use_smem) {
    for (int i = threadIdx.x; i < nh; i += THREADS) sm[i] = 0;
    __syncthreads();
  }
  const T range = hi - lo, nbm1 = static_cast<T>(n_bins - 1);
  for (long long i0 = static_cast<long long>(blockIdx.x) * THREADS; i0 < n;
       i0 += static_cast<long long>(gridDim.x) * THREADS) {
    const long long i = i0 + threadIdx.x;
    const bool live = i < n;
    const long long lab = live ? load_label(labels, label_dtype, i) : 0;
    for (int c = 0; c < k; ++c) {
      int b = 0;
      if (live) {
        const T sc = scores[i * k + c];
        if (!(sc > lo)) b = 0;
        else if (sc >= hi) b = n_bins - 1;
        else b = static_cast<int>((sc - lo) / range * nbm1);  // no fast-math, no contraction possible here
      }
      const uint32_t key = static_cast<uint32_t>((lab == c ? 0 : k * n_bins) + c * n_bins + b);
      if (use_smem) {
        warp_count(sm, key, live);
      } else if (live) {
        atomicAdd((lab == c ? pos_hist : neg_hist) + c * n_bins + b, 1ull);
      }
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < nh; i += THREADS) {
      const uint32_t v = sm[i];
      if (v) atomicAdd((i < k * n_bins ? pos_hist : neg_hist) + (i < k * n_bins ? i : i - k * n_bins),
                       static_cast<unsigned long long>(v));
    }
  }
}

template <int NC>
int launch_seg(const SegArgs& a, size_t smem, cudaStream_t st) {
  static IgPerDevice configured = {};
  if (!configured.get()) {
    IG_CUDA_OK(cudaFuncSetAttribute(seg_metrics_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.set(1);
  }
  const long long quads = a.n_img * (a.hw >> 2);
  if (quads == 0) return IG_OK;
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4);
  long long blocks = (quads + THREADS - 1) / THREADS;
  const long long cap = static_cast<long long>(ig_num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  ig::ProfScope prof(ig::PROF_MISC, st);
  seg_metrics_kernel<NC><<<static_cast<unsigned>(blocks), THREADS, smem, st>>>(a);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

int dispatch_seg(const SegArgs& a, size_t smem, cudaStream_t st) {
  switch (a.logits ? a.nc : 0) {
#define IG_CASE(n) case n: return launch_seg<n>(a, smem, st);
    IG_CASE(0) IG_CASE(1) IG_CASE(2) IG_CASE(3) IG_CASE(4) IG_CASE(5) IG_CASE(6) IG_CASE(7) IG_CASE(8)
    IG_CASE(9) IG_CASE(10) IG_CASE(11) IG_CASE(12) IG_CASE(13) IG_CASE(14) IG_CASE(15) IG_CASE(16)
#undef IG_CASE
    default: break;
  }
  ig_set_error("seg metrics: %d classes unsupported (1..%d)", a.nc, MAX_K);
  return IG_ESHAPE;
}

bool label_dtype_ok(int d) { return d == IG_I64 || d == IG_I32 || d == IG_U8 || d == IG_I8; }
int label_size(int d) { return d == IG_I64 ? 8 : (d == IG_I32 ? 4 : 1); }

// ------------------------------------------------------------------ regression sums (metrics.py:330-356)
struct RegArgs {
  const float *y_true, *y_pred;
  long long n;
  int has_ignore;
  float ignore;
  float ee_bias, ee_coef;
  double* sums;                  // x, y, xy, x2, y2, |e|, e2
  unsigned long long* counts;    // n, within expected error
};

__global__ void __launch_bounds__(THREADS) regression_kernel(const RegArgs a) {
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  unsigned long long cnt = 0, ee = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < a.n;
       i += static_cast<long long>(gridDim.x) * THREADS) {
    const float xf = __ldcs(a.y_true + i), yf = __ldcs(a.y_pred + i);
    if (a.has_ignore && xf == a.ignore) continue;
    const double x = xf, y = yf;
    const float ef = fabsf(__fsub_rn(yf, xf));  // the reference's float32 |y_pred - y_true|
    const double e = fabs(y - x);
    s[0] += x, s[1] += y, s[2] += x * y, s[3] += x * x, s[4] += y * y, s[5] += e, s[6] += e * e;
    ++cnt;
    ee += ef <= __fadd_rn(a.ee_bias, __fmul_rn(a.ee_coef, xf)) ? 1 : 0;
  }
  __shared__ double red[7][THREADS / 32];
  __shared__ unsigned long long redc[2][THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int j = 0; j < 7; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    ee += __shfl_xor_sync(0xffffffffu, ee, o);
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 7; ++j) red[j][warp] = s[j];
    redc[0][warp] = cnt, redc[1][warp] = ee;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    if (threadIdx.x < 7) {
      double t = 0;
      for (int w = 0; w < THREADS / 32; ++w) t += red[threadIdx.x][w];
      atomicAdd(a.sums + threadIdx.x, t);
    } else {
      unsigned long long t = 0;
      for (int w = 0; w < THREADS / 32; ++w) t += redc[threadIdx.x - 7][w];
      atomicAdd(a.counts + (threadIdx.x - 7), t);
    }
  }
}

}  // namespace

extern "C" int ig_confusion_update(const int8_t* pred, const void* labels, int label_dtype, int64_t n,
                                   int num_classes, int has_ignore, int64_t ignore_index,
                                   unsigned long long* matrix, unsigned long long* counters, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(matrix && counters, IG_EINVAL, "ig_confusion_update: null output");
  IG_REQUIRE(n >= 0 && (n == 0 || (pred && labels)), IG_EINVAL, "ig_confusion_update: null input");
  IG_REQUIRE(num_classes >= 1 && num_classes <= 64, IG_ESHAPE, "ig_confusion_update: num_classes %d (1..64)",
             num_classes);
  IG_REQUIRE(label_dtype_ok(label_dtype), IG_EINVAL, "ig_confusion_update: label dtype %d", label_dtype);
  if (n == 0) return IG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SegArgs a{};
  a.pred = pred, a.labels = labels, a.label_dtype = label_dtype;
  a.k = a.nc = num_classes;
  a.has_ignore = has_ignore, a.ignore = ignore_index;
  a.matrix = matrix, a.counters = counters;
  const size_t smem = (static_cast<size_t>(num_classes) * num_classes + 2) * 4;
  const bool aligned = (reinterpret_cast<uintptr_t>(pred) & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(labels) & (label_size(label_dtype) >= 4 ? 15 : 3)) == 0;
  const int64_t n4 = aligned ? (n & ~static_cast<int64_t>(3)) : 0;
  if (n4 > 0) {
    a.n_img = 1, a.hw = n4;
    IG_TRY(dispatch_seg(a, smem, st));
  }
  if (n > n4) {
    ig::ProfScope prof(ig::PROF_MISC, st);
    const int64_t rest = n - n4;
    long long blocks = (rest + THREADS - 1) / THREADS;
    if (blocks > 4ll * ig_num_sms()) blocks = 4ll * ig_num_sms();
    confusion_elem_kernel<<<static_cast<unsigned>(blocks), THREADS, smem, st>>>(a, n4, rest);
    IG_CUDA_OK(cudaGetLastError());
  }
  return IG_OK;
}

extern "C" int ig_seg_metrics_update(const float* logits, int64_t n_img, int num_classes, int64_t hw,
                                     const void* labels, int label_dtype, int has_ignore, int64_t ignore_index,
                                     unsigned long long* matrix, unsigned long long* counters, int n_bins,
                                     float min_score, float max_score, unsigned long long* pos_hist,
                                     unsigned long long* neg_hist, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(counters, IG_EINVAL, "ig_seg_metrics_update: null counters");
  IG_REQUIRE(n_img >= 0 && hw >= 0, IG_EINVAL, "ig_seg_metrics_update: negative size");
  IG_REQUIRE(num_classes >= 1 && num_classes <= MAX_K, IG_ESHAPE, "ig_seg_metrics_update: num_classes %d (1..%d)",
             num_classes, MAX_K);
  IG_REQUIRE(label_dtype_ok(label_dtype), IG_EINVAL, "ig_seg_metrics_update: label dtype %d", label_dtype);
  IG_REQUIRE(hw % 4 == 0, IG_ESHAPE, "ig_seg_metrics_update: pixels per image %lld must be a multiple of 4",
             static_cast<long long>(hw));
  if (n_img == 0 || hw == 0) return IG_OK;
  IG_REQUIRE(logits && labels, IG_EINVAL, "ig_seg_metrics_update: null input");
  const int lab_align = label_size(label_dtype) >= 4 ? 16 : 4;  // four labels per vector load
  IG_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(labels) & (lab_align - 1)) == 0,
             IG_EINVAL, "ig_seg_metrics_update: logits must be 16-byte and labels %d-byte aligned", lab_align);
  const bool auc = pos_hist != nullptr;
  if (auc) {
    IG_REQUIRE(neg_hist, IG_EINVAL, "ig_seg_metrics_update: AUC needs both histograms");
    IG_REQUIRE(n_bins >= 2 && max_score > min_score, IG_EINVAL, "ig_seg_metrics_update: bad histogram range");
  }
  SegArgs a{};
  a.logits = logits, a.labels = labels, a.label_dtype = label_dtype;
  a.n_img = n_img, a.hw = hw, a.nc = a.k = num_classes;
  a.has_ignore = has_ignore, a.ignore = ignore_index;
  a.matrix = matrix, a.counters = counters;
  a.n_bins = n_bins, a.lo = min_score, a.hi = max_score;
  a.pos_hist = pos_hist, a.neg_hist = neg_hist;
  const size_t smem = (static_cast<size_t>(num_classes) * num_classes + 2 +
                       (auc ? 2ull * num_classes * n_bins : 0)) * 4;
  IG_REQUIRE(smem <= 200 * 1024, IG_ESHAPE,
             "ig_seg_metrics_update: num_classes * n_bins = %d x %d does not fit the shared-memory histograms",
             num_classes, n_bins);
  return dispatch_seg(a, smem, static_cast<cudaStream_t>(stream));
}

extern "C" int ig_auc_update(const void* scores, int score_dtype, int64_t n, int num_classes, const void* labels,
                             int label_dtype, int n_bins, double min_score, double max_score,
                             unsigned long long* pos_hist, unsigned long long* neg_hist, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(pos_hist && neg_hist, IG_EINVAL, "ig_auc_update: null output");
  IG_REQUIRE(score_dtype == IG_F32 || score_dtype == IG_F64, IG_EINVAL, "ig_auc_update: scores must be f32 or f64");
  IG_REQUIRE(label_dtype_ok(label_dtype), IG_EINVAL, "ig_auc_update: label dtype %d", label_dtype);
  IG_REQUIRE(n >= 0 && num_classes >= 1 && n_bins >= 2 && max_score > min_score, IG_EINVAL,
             "ig_auc_update: bad sizes or histogram range");
  IG_REQUIRE(static_cast<long long>(num_classes) * n_bins <= (1 << 24), IG_ESHAPE, "ig_auc_update: histogram too large");
  if (n == 0) return IG_OK;
  IG_REQUIRE(scores && labels, IG_EINVAL, "ig_auc_update: null input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = 2ull * num_classes * n_bins * 4;
  const int use_smem = smem <= 96 * 1024;
  static IgPerDevice configured = {};
  if (!configured.get()) {
    IG_CUDA_OK(cudaFuncSetAttribute(auc_scores_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    IG_CUDA_OK(cudaFuncSetAttribute(auc_scores_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured.set(1);
  }
  long long blocks = (n + THREADS - 1) / THREADS;
  if (blocks > 2ll * ig_num_sms()) blocks = 2ll * ig_num_sms();
  ig::ProfScope prof(ig::PROF_MISC, st);
  if (score_dtype == IG_F32)
    auc_scores_kernel<float><<<static_cast<unsigned>(blocks), THREADS, use_smem ? smem : 0, st>>>(
        static_cast<const float*>(scores), labels, label_dtype, n, num_classes, n_bins,
        static_cast<float>(min_score), static_cast<float>(max_score), pos_hist, neg_hist, use_smem);
  else
    auc_scores_kernel<double><<<static_cast<unsigned>(blocks), THREADS, use_smem ? smem : 0, st>>>(
        static_cast<const double*>(scores), labels, label_dtype, n, num_classes, n_bins, min_score, max_score,
        pos_hist, neg_hist, use_smem);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}

extern "C" int ig_regression_update(const float* y_true, const float* y_pred, int64_t n, int has_ignore,
                                    float ignore_value, float ee_bias, float ee_coef, double* sums,
                                    unsigned long long* counts, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(sums && counts, IG_EINVAL, "ig_regression_update: null output");
  IG_REQUIRE(n >= 0, IG_EINVAL, "ig_regression_update: negative size");
  if (n == 0) return IG_OK;
  IG_REQUIRE(y_true && y_pred, IG_EINVAL, "ig_regression_update: null input");
  RegArgs a{y_true, y_pred, n, has_ignore, ignore_value, ee_bias, ee_coef, sums, counts};
  long long blocks = (n + THREADS * 8 - 1) / (THREADS * 8);
  if (blocks > 4ll * ig_num_sms()) blocks = 4ll * ig_num_sms();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ig::ProfScope prof(ig::PROF_MISC, st);
  regression_kernel<<<static_cast<unsigned>(blocks), THREADS, 0, st>>>(a);
  IG_CUDA_OK(cudaGetLastError());
  return IG_OK;
}
