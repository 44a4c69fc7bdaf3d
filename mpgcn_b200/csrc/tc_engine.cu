// Host side of the tcgen05 contraction engine: kernel instantiations, launch, tensor maps.
#include "tc_engine.cuh"

#include <mutex>
#include <stdarg.h>
#include <string.h>

namespace mpgcn {

// ---------------------------------------------------------------------------------------
// error string (thread local), device attributes
// ---------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---------------------------------------------------------------------------------------
// tensor maps: cuTensorMapEncodeTiled resolved through the runtime (no -lcuda at link time,
// so the library loads -- and its symbols can be checked -- on a machine without a driver)
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_f16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, TmapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  MPGCN_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  MPGCN_CHECK(rank >= 1 && rank <= 4, "tensor map rank %d unsupported", rank);
  cuuint64_t gdim[4] = {1, 1, 1, 1};
  cuuint64_t gstr[3] = {0, 0, 0};
  cuuint32_t bx[4] = {1, 1, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  // always encode rank 4 (the kernel issues 4-d copies); pad with unit dims whose stride
  // continues the outermost real stride
  uint64_t last = (rank >= 2) ? strides_bytes[rank - 2] * dims[rank - 1] : dims[0] * 2;
  for (int i = 0; i < 3; ++i) {
    if (i < rank - 1) gstr[i] = strides_bytes[i];
    else { gstr[i] = align_up(last, 16); }
  }
  for (int i = 0; i < 3; ++i)
    MPGCN_CHECK(gstr[i] % 16 == 0 && gstr[i] < (1ull << 40), "tensor map stride %llu of dim %d is not a multiple of 16",
                (unsigned long long)gstr[i], i + 1);
  MPGCN_CHECK((reinterpret_cast<uintptr_t>(gptr) & 15) == 0, "tensor map base pointer must be 16-byte aligned");
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(gptr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz == TMAP_SW128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MPGCN_CHECK(r == CUDA_SUCCESS,
              "cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) box=(%u,%u,%u,%u)", (int)r,
              (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2], (unsigned long long)gdim[3],
              (unsigned long long)gstr[0], (unsigned long long)gstr[1], (unsigned long long)gstr[2], bx[0], bx[1], bx[2], bx[3]);
  return 0;
}

namespace tc {

static const int kMaxSmem = 232448;   // 227 KB opt-in limit per CTA on sm_100

template <int AK, int BK>
static int launch_impl(GemmParams& p, cudaStream_t stream) {
  using C = Cfg<AK, BK>;
  static bool attr_done = false;
  if (!attr_done) {
    MPGCN_CUDA(cudaFuncSetAttribute(contract_kernel<AK, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_done = true;
  }
  MPGCN_CHECK(p.R >= 1 && p.R <= 8, "R=%d out of range", p.R);
  // N of the UMMA must be a multiple of 16 for M=128: R odd -> N=32R is still a multiple of 32. ok.
  const size_t stage_bytes = (size_t)C::A_STAGE + (size_t)p.R * BK * 64;
  int stages = (int)((kMaxSmem - 1024 - 512) / stage_bytes);
  if (stages > 8) stages = 8;
  MPGCN_CHECK(stages >= 2, "tile does not fit in shared memory");
  p.stages = stages;
  // always request the full opt-in budget: exactly one CTA per SM, so the 512-column TMEM allocation never contends
  const size_t smem = kMaxSmem;
  MPGCN_CHECK(smem_bytes(C::A_STAGE, p.R, BK, stages) <= smem, "internal: smem budget");
  const long long tiles = (long long)p.MT * p.NT * p.Z;
  MPGCN_CHECK(tiles > 0 && tiles < (1ll << 31), "bad tile count %lld", tiles);
  MPGCN_CHECK(p.kb_total > 0 && p.kb_per_seg > 0, "empty contraction");
  int grid = (int)(tiles < device_sm_count() ? tiles : device_sm_count());
  contract_kernel<AK, BK><<<grid, kThreads, smem, stream>>>(p);
  MPGCN_CUDA(cudaGetLastError());
  return 0;
}

int launch_contract(int ak, int bk, GemmParams& p, cudaStream_t stream) {
  if (ak == A_MN128 && bk == 64) return launch_impl<A_MN128, 64>(p, stream);
  if (ak == A_K128 && bk == 64) return launch_impl<A_K128, 64>(p, stream);
  if (ak == A_K64 && bk == 32) return launch_impl<A_K64, 32>(p, stream);
  if (ak == A_MN64 && bk == 64) return launch_impl<A_MN64, 64>(p, stream);
  set_error("no contraction kernel for A kind %d, BK %d", ak, bk);
  return 1;
}

}  // namespace tc
}  // namespace mpgcn
