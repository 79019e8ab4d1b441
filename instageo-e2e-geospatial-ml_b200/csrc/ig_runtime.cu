// Host runtime glue: thread-local error string, device gate, TMA descriptor encoding.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "ig_common.cuh"

static thread_local char g_err[512] = "";

void ig_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* ig_last_error(void) { return g_err; }
extern "C" int ig_version(void) { return 100; }  // 0.1.0

int ig_check_device() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = IG_OK;
  int dev = 0;
  IG_CUDA_OK(cudaGetDevice(&dev));
  if (dev == cached_dev) {
    if (cached_rc != IG_OK) ig_set_error("device %d is not an sm_100 (B200) GPU", dev);
    return cached_rc;
  }
  int major = 0, minor = 0;
  IG_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  IG_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  cached_dev = dev;
  cached_rc = (major == 10 && minor == 0) ? IG_OK : IG_EARCH;
  if (cached_rc != IG_OK)
    ig_set_error("device %d has compute capability %d.%d; this library is sm_100a only", dev, major, minor);
  return cached_rc;
}

int ig_num_sms() {
  static thread_local int dev_cached = -1, sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != dev_cached) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    dev_cached = dev;
  }
  return sms;
}

bool ig_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("IG_NO_PDL");
    v = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return v == 1;
}

int IgPerDevice::get() const {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  return __atomic_load_n(&v[dev], __ATOMIC_ACQUIRE);
}
void IgPerDevice::set(int value) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  __atomic_store_n(&v[dev], value, __ATOMIC_RELEASE);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int ig_make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_pitch, uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode();
  IG_REQUIRE(enc != nullptr, IG_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  IG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, IG_EINVAL, "TMA base %p not 16-byte aligned", base);
  IG_REQUIRE((row_pitch * 2) % 16 == 0, IG_ESHAPE, "TMA row pitch %llu elements not 16-byte aligned",
             static_cast<unsigned long long>(row_pitch));
  IG_REQUIRE(box_rows >= 1 && box_rows <= 256 && (box_cols == 64 || box_cols == 32 || box_cols == 16), IG_ESHAPE,
             "TMA box %ux%u unsupported (rows of 128, 64 or 32 bytes)", box_rows, box_cols);
  // the swizzle span equals the box row: 64 columns = SWIZZLE_128B (every full K block), 32 / 16 columns = SWIZZLE_64B /
  // SWIZZLE_32B (the narrow last K block of a K that is not a multiple of 64, see gemm::Args::rem_cols)
  const CUtensorMapSwizzle swz = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {row_pitch * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IG_REQUIRE(r == CUDA_SUCCESS, IG_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu)",
             static_cast<int>(r), static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols));
  return IG_OK;
}

int ig_make_tmap_f32_tile(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch,
                          uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  IG_REQUIRE(enc != nullptr, IG_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  IG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, IG_EINVAL, "TMA base %p not 16-byte aligned", base);
  IG_REQUIRE((row_pitch * 4) % 16 == 0 && cols % 32 == 0 && box_rows >= 1 && box_rows <= 256, IG_ESHAPE,
             "f32 TMA tile: pitch %llu / cols %llu / box rows %u unsupported", static_cast<unsigned long long>(row_pitch),
             static_cast<unsigned long long>(cols), box_rows);
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {row_pitch * 4};
  const cuuint32_t box[2] = {32, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IG_REQUIRE(r == CUDA_SUCCESS, IG_ECUDA, "cuTensorMapEncodeTiled (f32 tile) failed with CUresult %d", static_cast<int>(r));
  return IG_OK;
}

int ig_make_tmap_nd(CUtensorMap* map, int dtype, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  IG_REQUIRE(enc != nullptr, IG_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  IG_REQUIRE(rank >= 1 && rank <= 5, IG_EINVAL, "TMA rank %d unsupported", rank);
  IG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, IG_EINVAL, "TMA base %p not 16-byte aligned", base);
  CUtensorMapDataType dt;
  switch (dtype) {
    case IG_F32: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    case IG_BF16: dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; break;
    case IG_I16: case IG_U16: dt = CU_TENSOR_MAP_DATA_TYPE_UINT16; break;
    default: ig_set_error("TMA dtype %d unsupported", dtype); return IG_EINVAL;
  }
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
    IG_REQUIRE(box[i] >= 1 && box[i] <= 256, IG_ESHAPE, "TMA box dim %d = %u unsupported", i, box[i]);
    if (i > 0) {
      st[i - 1] = strides_bytes[i - 1];
      IG_REQUIRE(strides_bytes[i - 1] % 16 == 0, IG_ESHAPE, "TMA stride %d = %llu bytes not a multiple of 16", i,
                 static_cast<unsigned long long>(strides_bytes[i - 1]));
    }
  }
  const CUresult r = enc(map, dt, rank, const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);  // out-of-bounds elements read as zero
  IG_REQUIRE(r == CUDA_SUCCESS, IG_ECUDA, "cuTensorMapEncodeTiled (rank %d) failed with CUresult %d", rank,
             static_cast<int>(r));
  return IG_OK;
}

// ----------------------------------------------------------------------------- launch profiling
namespace {
struct ProfRec { int cat; cudaEvent_t e0, e1; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_open[ig::PROF_COUNT];
}  // namespace

bool ig::prof_enabled() { return g_prof_on; }
void ig::prof_begin(int cat, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_open[cat].push_back(e);
}
void ig::prof_end(int cat, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_open[cat].empty()) return;
  cudaEvent_t e0 = g_open[cat].back();
  g_open[cat].pop_back();
  cudaEvent_t e1;
  if (cudaEventCreate(&e1) != cudaSuccess) return;
  cudaEventRecord(e1, st);
  g_prof.push_back({cat, e0, e1});
}

extern "C" int ig_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  return IG_OK;
}

// Per-launch form of the report (in launch order): ms[i], cat[i] for the first `cap` launches since the last report;
// returns the number of launches recorded (may exceed cap) and clears the records.
extern "C" int ig_profile_report_launches(double* ms, int* cat, int cap) {
  IG_REQUIRE(ms && cat && cap >= 0, IG_EINVAL, "ig_profile_report_launches: null pointer");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (auto& r : g_prof) {
    float t = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess && n < cap) {
      ms[n] = t;
      cat[n] = r.cat;
    }
    ++n;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return n;
}

// ms[c] = summed device time of family c since the last report, launches[c] = launch count.
extern "C" int ig_profile_report(double* ms, int* launches, int ncat) {
  IG_REQUIRE(ms && launches && ncat >= ig::PROF_COUNT, IG_EINVAL, "ig_profile_report: need %d categories", ig::PROF_COUNT);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < ncat; ++i) { ms[i] = 0; launches[i] = 0; }
  for (auto& r : g_prof) {
    float t = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
      ms[r.cat] += t;
      launches[r.cat] += 1;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return IG_OK;
}
