// Shared host/device helpers for libinstageo_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/instageo_b200.h"

// ----------------------------------------------------------------------------- host errors
void ig_set_error(const char* fmt, ...);
int ig_check_device();  // IG_OK if current device is sm_100, else IG_EARCH (+message)

#define IG_CUDA_OK(expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ig_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,      \
                   __LINE__);                                                             \
      return IG_ECUDA;                                                                    \
    }                                                                                     \
  } while (0)

#define IG_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ig_set_error(__VA_ARGS__);     \
      return (code);                 \
    }                                \
  } while (0)

#define IG_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != IG_OK) return _r; \
  } while (0)

int ig_num_sms();
// Programmatic dependent launch of the forward's kernel chain (on unless IG_NO_PDL=1): see ig::launch below.
bool ig_pdl_enabled();

// cudaFuncSetAttribute is a PER-DEVICE setting: a process that drives several GPUs (one model per device, or the
// test suite's cuda:1 cases) must configure every kernel once on each of them.  get() = the largest setting made so
// far on the current device for this call site (0 = none); the caller records a new one with set().
struct IgPerDevice {
  int v[64];
  int get() const;
  void set(int value);
};

// TMA descriptor for a row-major 2-D bf16 matrix [rows, cols] (cols contiguous), box
// [box_rows, box_cols], box_cols = 64 (128-byte swizzle), 32 (64-byte) or 16 (32-byte).  row_pitch in elements.
// f32 [rows, cols] row-major, box [box_rows, 32 columns] = 128-byte rows, 128-byte swizzle (epilogue staging tiles)
int ig_make_tmap_f32_tile(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch,
                          uint32_t box_rows);
int ig_make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_pitch, uint32_t box_rows, uint32_t box_cols);

// General tiled TMA descriptor (no swizzle, zero fill out of bounds): dims/box innermost first,
// strides_bytes[i] = byte stride of dimension i+1.  dtype = IG_F32 | IG_BF16 | IG_I16 | IG_U16.
int ig_make_tmap_nd(CUtensorMap* map, int dtype, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box);

// Optional per-launch CUDA-event timing (ig_profile_enable): brackets a launch with two events on
// the launching stream; ig_profile_report sums elapsed time per kernel family.
namespace ig {
enum ProfCat { PROF_PREPROCESS = 0, PROF_STITCH, PROF_GEMM_LINEAR, PROF_GEMM_CONV, PROF_ATTENTION,
               PROF_LAYERNORM, PROF_MISC, PROF_COUNT };
bool prof_enabled();
void prof_begin(int cat, cudaStream_t st);
void prof_end(int cat, cudaStream_t st);
struct ProfScope {
  int cat;
  cudaStream_t st;
  ProfScope(int c, cudaStream_t s) : cat(c), st(s) { prof_begin(cat, st); }
  ~ProfScope() { prof_end(cat, st); }
};
}  // namespace ig

// ----------------------------------------------------------------------------- device PTX
#ifdef __CUDACC__
namespace ig {

// Launch with (optionally) programmatic stream serialisation: the kernel may become resident and run its prologue
// (shared-memory carve-up, mbarrier init, TMEM allocation, tensor-map prefetch) while its predecessor in the stream is
// still draining; it must execute pdl_wait() before it touches global memory.  Every kernel of the forward chain calls
// pdl_launch_dependents() + pdl_wait() right after its prologue (both are no-ops for a plain launch).  The dependents
// are released once ALL CTAs of the primary have issued the trigger (or exited), i.e. during its last wave, so they
// can never take SM resources from CTAs of the primary that have not started.  Captured into the forward's CUDA graph
// these launches become programmatic dependency edges.
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  const bool on = pdl && ig_pdl_enabled();
  cfg.attrs = on ? at : nullptr;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware parks the warp until the phase completes or ~`ns` elapse,
// instead of returning after its short default limit.  A waiting role warp (TMA producer, MMA issuer) that polls
// without the hint issues a try_wait + branch every few cycles and takes issue slots from the compute warps of
// its scheduler: in the attention kernel a third of all executed instructions were such polls, and the two
// softmax warps that share a scheduler with the producer / MMA warps fell thousands of cycles behind the others.
__device__ __forceinline__ bool mbar_try_wait_parked(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (-> CUDA error on the host) instead of hanging
// the GPU.  ~4 s budget, far above any kernel in this library.  The spin loop lives out of line
// so the hot path (barrier already complete) is a single try_wait.
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  unsigned long long t0 = 0;
  uint32_t spins = 0;
#ifndef IG_WAIT_MODE
#define IG_WAIT_MODE 0
#endif
#if IG_WAIT_MODE == 0
  while (!mbar_try_wait_parked(bar, parity, 20000u)) {
#elif IG_WAIT_MODE == 1
  while (!mbar_try_wait(bar, parity)) {
#else
  while (!mbar_try_wait_parked(bar, parity, 200u)) {
#endif
    if ((++spins & 0x3f) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("ig: mbarrier timeout block (%d,%d,%d) thread %d parity %u\n", blockIdx.x,
               blockIdx.y, blockIdx.z, threadIdx.x, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}
// One lane of a converged warp (elect.sync): the issuing lane of TMA / tcgen05 instructions.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// Warp index as a value the compiler KNOWS is warp-uniform (shuffle broadcast), so role branches are
// uniform branches and descriptor arithmetic stays on the uniform datapath.
__device__ __forceinline__ int warp_idx_uniform() {
  return __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
}

// ---- proxies / fences
// ---- bulk tensor reduction shared -> global (f32 add performed by the L2, element by element; no read into the SM)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's committed bulk groups have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// L2 prefetch of a 2-D box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// ---- UMMA (tcgen05.mma kind::f16, bf16 x bf16 -> f32 in TMEM)
// Shared-memory matrix descriptor, 128-byte swizzle (sm_100 "version 1" format):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
//   [46,48) version = 1, [49,52) base offset, [61,64) layout type (2 = SWIZZLE_128B).
// K-major operand: rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO); LBO unused.
// MN-major operand (64 MN elements wide): rows are K, 8-row K groups 1024 B apart (SBO).
// The base-offset field [49,52) stays 0 even when the operand starts INSIDE a 1024-byte swizzle atom
// (row-shifted taps of the implicit-GEMM convolution start at tile + shift*128 B): measured on B200,
// the 128-byte swizzle XOR is applied to absolute shared-memory address bits, so a shifted start reads
// exactly the rows TMA wrote; setting base_offset = (addr >> 7) & 7 produced wrong operands.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes,
                                                    uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// Instruction descriptor: c=f32 (bits 4-5 = 1), a=b=bf16 (bits 7-9, 10-12 = 1),
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 at 17, M>>4 at 24.
// Same descriptor as (lo, hi) words: only `lo` changes inside a K loop (start address >> 4, +2 per
// 16 bf16 of K), `hi` is the constant SBO=1024 | version 1 | SWIZZLE_128B word.
constexpr uint32_t UMMA_DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
// K-major tiles whose rows are 64 / 32 bytes (32 / 16 bf16: the narrow last K block): 8-row groups 512 / 256 B apart,
// layout type 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
constexpr uint32_t UMMA_DESC_HI_SW64 = (512u >> 4) | (1u << 14) | (4u << 29);
constexpr uint32_t UMMA_DESC_HI_SW32 = (256u >> 4) | (1u << 14) | (6u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16);
}
__device__ __forceinline__ uint64_t umma_desc_pack(uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(UMMA_DESC_HI_SW128));
  return d;
}
__device__ __forceinline__ uint64_t umma_desc_pack_hi(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn,
                                                    uint32_t b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand in TENSOR MEMORY (M = 128 rows on the lanes, K along the columns, two bf16 per 32-bit
// column: 16 K elements = 8 columns): D[128 x N] += A_tmem[128 x 16] * B_smem[N x 16]^T.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued tcgen05.mma of this thread arrive on `bar` when complete
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}

// TMEM -> registers: 32 lanes x 32 bit, 16 / 32 consecutive columns per thread.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM, 32 lanes x 32 consecutive columns (the read-modify-write of a lazy softmax rescale)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
      "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
      "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster drive one 256-row UMMA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// Arrive on an mbarrier of a (possibly remote) CTA of the cluster.  Used to hand a TMEM accumulator stage back to the
// pair leader's MMA warp: the tcgen05.ld's are ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync, so the
// arrive itself needs no memory ordering.  The first version asked for .release.cluster, which ptxas turns into
// MEMBAR.ALL.GPU (+ ERRBAR / CGAERRBAR): every epilogue warp then waited, once per tile, until the global stores of its
// previous tile had been acknowledged -- the ~4-5 k clk of per-tile epilogue latency that bounded every short-K /
// narrow-N tile.  IG_ARRIVE_SEM: 0 = .relaxed.cluster, 1 = the default .release.cta (shipped; what CUTLASS emits), 2 = .release.cluster (old).
#ifndef IG_ARRIVE_SEM
#define IG_ARRIVE_SEM 1
#endif
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
#if IG_ARRIVE_SEM == 0
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#elif IG_ARRIVE_SEM == 1
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
// TMA load issued by either CTA of a pair; completes on an mbarrier of the pair's leader
// (`bar_cluster_addr` is a shared::cluster address).
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m,
                                                uint32_t bar_cluster_addr, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[256 x N] (128 TMEM lanes in each CTA of the pair) += A[256 x 16] * B[N x 16]^T; A rows and
// B rows are split across the two CTAs' shared memories at identical offsets.
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of a pair: arrives on the barrier at the same offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// TMEM -> registers, 16 columns into v[0..16) (pointer form, for runtime-sized column groups)
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// erf by the float32 rational approximation used by Eigen / XLA (|err| ~ 1e-7 abs),
// ~14 FMA + one reciprocal; cheap enough to hide in a GEMM epilogue.
__device__ __forceinline__ float erf_fast(float x) {
  x = fminf(fmaxf(x, -4.f), 4.f);
  const float x2 = x * x;
  float p = -2.72614225801306e-10f;
  p = fmaf(p, x2, 2.77068142495902e-08f);
  p = fmaf(p, x2, -2.10102402082508e-06f);
  p = fmaf(p, x2, -5.69250639462346e-05f);
  p = fmaf(p, x2, -7.34990630326855e-04f);
  p = fmaf(p, x2, -2.95459980854025e-03f);
  p = fmaf(p, x2, -1.60960333262415e-02f);
  float q = -1.45660718464996e-05f;
  q = fmaf(q, x2, -2.13374055278905e-04f);
  q = fmaf(q, x2, -1.68282697438203e-03f);
  q = fmaf(q, x2, -7.37332916720468e-03f);
  q = fmaf(q, x2, -1.42647390514189e-02f);
  return __fdividef(x * p, q);
}
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + erf_fast(x * 0.70710678118654752f));
}
// GELU(erf) through one MUFU.TANH: erf(x/sqrt2) ~ tanh(x*(c0 + c1 x^2 + c2 x^4)), coefficients
// least-squares fitted on [-8, 8]; |gelu_tanh3 - gelu_erf| <= 3.0e-5 (below bf16 rounding of the
// stored activation), ~9 instructions instead of ~30, so the fc1 epilogue stays under the MMA time.
__device__ __forceinline__ float gelu_tanh3(float x) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f);
  const float x2 = xc * xc;
  const float u = xc * fmaf(x2, fmaf(x2, -3.58732361e-04f, 3.70503451e-02f), 7.97458471e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

}  // namespace ig
#endif  // __CUDACC__
