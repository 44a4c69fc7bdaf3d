// Shared device/host helpers for the mpgcn_b200 CUDA library (sm_100a only).
//
// PTX wrappers for the Blackwell primitives the engine uses: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences).  Nothing here is
// reference code; the reference (underdoc-wang/MPGCN) has no native code at all.
#pragma once

#include <cuda.h>            // CUtensorMap (types only; libcuda is resolved at run time)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && !(defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM103_ALL))
#error "mpgcn_b200 must be compiled for sm_100a (-gencode arch=compute_100a,code=sm_100a)"
#endif

namespace mpgcn {

// ----------------------------------------------------------------------------------------
// host-side error plumbing (C-ABI functions return int, message via mpgcn_last_error())
// ----------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define MPGCN_CHECK(cond, ...)                                                     \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      ::mpgcn::set_error(__VA_ARGS__);                                             \
      return 1;                                                                    \
    }                                                                              \
  } while (0)
#define MPGCN_CUDA(call)                                                           \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      ::mpgcn::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ----------------------------------------------------------------------------------------
// device: shared-memory addresses, mbarrier
// ----------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (-> cudaErrorLaunchFailure), never hang the GPU.  The slow path is kept out
// of line so that the many call sites do not bloat the kernels past the instruction cache.
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0xFFFu) == 0 && (clock64() - t0) > 6000000000LL) {   // ~3 s
      printf("mpgcn_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}

// ----------------------------------------------------------------------------------------
// device: TMA (bulk tensor copies global -> shared, completion on an mbarrier)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// 2-CTA (cta_group::2) variant: executed by both CTAs of a pair; the transaction bytes are credited to the
// mbarrier of the pair's leader CTA (peer bit 24 of the shared::cluster address cleared).
__device__ __forceinline__ void tma_load_4d_2cta(void* smem_dst, const CUtensorMap* map, uint64_t* leader_bar,
                                                 int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(leader_bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ----------------------------------------------------------------------------------------
// device: thread-block clusters
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(rank)
      : "memory");
}

// ----------------------------------------------------------------------------------------
// device: tcgen05 (TMEM allocation, UMMA issue, commit, TMEM loads, fences)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {   // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; single thread issues on behalf of the CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// pair-wide MMA: M = 256 (128 rows per CTA), issued by the leader CTA only
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit, arriving on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
// 32 lanes x 32 columns of 32-bit accumulators -> 32 registers per thread (thread = lane).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 (Blackwell) | [61,64) layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t umma_desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) | (1ull << 14) | (static_cast<uint64_t>(layout_type) << 29)) << 32;
}
__device__ __forceinline__ uint64_t umma_desc(uint64_t hi, uint32_t smem_addr, uint32_t lbo_bytes) {
  return hi | static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
// Instruction descriptor for kind::f16, fp16 x fp16 -> fp32 (cute::UMMA::InstrDescriptor):
//   [4,6) D fmt = 1 (F32) | [7,10) A fmt = 0 (F16) | [10,13) B fmt = 0 (F16) | [15] A major (1 = MN)
//   [16] B major (1 = MN) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
#endif  // __CUDACC__

// ----------------------------------------------------------------------------------------
// host: tensor-map encoding through the driver entry point (no link-time libcuda dependency)
// ----------------------------------------------------------------------------------------
enum TmapSwizzle { TMAP_SW64 = 0, TMAP_SW128 = 1 };
// fp16 tensor, up to 4 dims (innermost first); strides_bytes[i] is the stride of dim i+1.
int make_tmap_f16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, TmapSwizzle swz);

int device_sm_count();

// Function attributes (the > 48 KB dynamic shared-memory opt-in) belong to a device / context, not to the process: a
// kernel that already ran on cuda:0 still needs the opt-in on cuda:1.  One cache per kernel, indexed by device ordinal.
struct DynSmemAttr { int bytes[64]; };
int ensure_dyn_smem_impl(const void* kernel, int bytes, DynSmemAttr& cache);
template <class Kernel>
inline int ensure_dyn_smem(Kernel kernel, int bytes, DynSmemAttr& cache) {
  return ensure_dyn_smem_impl(reinterpret_cast<const void*>(kernel), bytes, cache);
}

// ----------------------------------------------------------------------------------------
// host: launch accounting / per-launch CUDA-event timing (bench.py's roofline evidence)
// ----------------------------------------------------------------------------------------
enum ProfTag {
  PROF_FWD_A = 0, PROF_FWD_MIX, PROF_FWD_B, PROF_BWD_V, PROF_BWD_DW, PROF_BWD_MIX, PROF_BWD_DX,   // tcgen05 contractions
  PROF_SIMT_GEMM, PROF_ELEMENTWISE, PROF_LSTM_FWD, PROF_LSTM_BWD,
  // regions, not kernels: a whole C-ABI call (every kernel of it, the gaps between them included); they count calls, not launches
  PROF_LAYER_FWD, PROF_LAYER_BWD, PROF_HEAD,
  PROF_EXCHANGE,      // kernels: the peer-memory exchange steps of the row shard (rows_reduce_bias_act, relu_backward_scatter[_f16])
  PROF_NUM_TAGS
};
constexpr int PROF_FIRST_REGION_TAG = PROF_LAYER_FWD;
// All of these may be called from several host threads (one stream each): the counters sit behind a mutex, the
// "next launch" annotation and the open begin/end bracket are thread-local.
void prof_set_next(int tag, double flops);             // annotate the next contraction launch (this thread's)
void prof_take_next(int* tag, double* flops);          // fetch and clear this thread's annotation
void prof_count(int tag);                              // count one launch of our own kernels
void prof_begin(int tag, double flops, cudaStream_t s);   // event before launch (no-op unless enabled)
void prof_end(cudaStream_t s);                            // event after launch
// whole-call bracket (nests around the per-launch brackets): event pair recorded only while profiling is enabled
struct ProfRegion {
  void* a = nullptr;
  int tag;
  cudaStream_t s;
  ProfRegion(int tag, double flops, cudaStream_t s);
  ~ProfRegion();
};

}  // namespace mpgcn
