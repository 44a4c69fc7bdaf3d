// Host side of the tcgen05 contraction engine: kernel instantiations, launch, tensor maps.
#include "tc_engine.cuh"

#include <mutex>
#include <vector>
#include <stdarg.h>
#include <stdlib.h>
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

int ensure_dyn_smem_impl(const void* kernel, int bytes, DynSmemAttr& cache) {
  static std::mutex mu;
  int dev = 0;
  MPGCN_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  const bool known = dev >= 0 && dev < 64;
  if (!known || cache.bytes[dev] < bytes) {
    MPGCN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (known) cache.bytes[dev] = bytes;
  }
  return 0;
}

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
// launch accounting and optional per-launch event timing
// ---------------------------------------------------------------------------------------
namespace {
struct ProfState {
  bool enabled = false;
  long long launches[PROF_NUM_TAGS] = {0};
  double flops[PROF_NUM_TAGS] = {0};
  double ms[PROF_NUM_TAGS] = {0};
  struct Pending { cudaEvent_t a, b; int tag; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  std::mutex mu;
} g_prof;
thread_local int t_cur_tag = -1;
thread_local cudaEvent_t t_cur_a = nullptr;
thread_local int t_next_tag = -1;
thread_local double t_next_flops = 0;
cudaEvent_t prof_event() {          // caller holds g_prof.mu
  if (!g_prof.pool.empty()) { cudaEvent_t e = g_prof.pool.back(); g_prof.pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void prof_set_next(int tag, double flops) { t_next_tag = tag; t_next_flops = flops; }
void prof_take_next(int* tag, double* flops) {
  *tag = t_next_tag;
  *flops = t_next_flops;
  t_next_tag = -1;
  t_next_flops = 0;
}
void prof_count(int tag) {
  std::lock_guard<std::mutex> lock(g_prof.mu);
  if (tag >= 0 && tag < PROF_NUM_TAGS) g_prof.launches[tag]++;
}
void prof_begin(int tag, double flops, cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_prof.mu);
  if (tag >= 0 && tag < PROF_NUM_TAGS) { g_prof.launches[tag]++; g_prof.flops[tag] += flops; }
  if (!g_prof.enabled) return;
  t_cur_tag = tag;
  t_cur_a = prof_event();
  cudaEventRecord(t_cur_a, s);
}
void prof_end(cudaStream_t s) {
  if (t_cur_a == nullptr) return;
  std::lock_guard<std::mutex> lock(g_prof.mu);
  cudaEvent_t b = prof_event();
  cudaEventRecord(b, s);
  g_prof.pending.push_back({t_cur_a, b, t_cur_tag});
  t_cur_a = nullptr;
}
ProfRegion::ProfRegion(int tag_, double flops, cudaStream_t s_) : tag(tag_), s(s_) {
  std::lock_guard<std::mutex> lock(g_prof.mu);
  if (tag >= 0 && tag < PROF_NUM_TAGS) { g_prof.launches[tag]++; g_prof.flops[tag] += flops; }
  if (!g_prof.enabled) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, s);
  a = e;
}
ProfRegion::~ProfRegion() {
  if (a == nullptr) return;
  std::lock_guard<std::mutex> lock(g_prof.mu);
  cudaEvent_t b = prof_event();
  cudaEventRecord(b, s);
  g_prof.pending.push_back({static_cast<cudaEvent_t>(a), b, tag});
}
void prof_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof.mu);
  g_prof.enabled = on != 0;
}
void prof_reset() {
  std::lock_guard<std::mutex> lock(g_prof.mu);
  for (int i = 0; i < PROF_NUM_TAGS; ++i) { g_prof.launches[i] = 0; g_prof.flops[i] = 0; g_prof.ms[i] = 0; }
  for (auto& p : g_prof.pending) { g_prof.pool.push_back(p.a); g_prof.pool.push_back(p.b); }
  g_prof.pending.clear();
}
// resolves pending event pairs (caller must have synchronised the stream)
int prof_read(int tag, long long* launches, double* flops, double* ms) {
  std::lock_guard<std::mutex> lock(g_prof.mu);
  for (auto& p : g_prof.pending) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess && p.tag >= 0 && p.tag < PROF_NUM_TAGS) g_prof.ms[p.tag] += t;
    g_prof.pool.push_back(p.a);
    g_prof.pool.push_back(p.b);
  }
  g_prof.pending.clear();
  if (tag < 0 || tag >= PROF_NUM_TAGS) return 1;
  *launches = g_prof.launches[tag];
  *flops = g_prof.flops[tag];
  *ms = g_prof.ms[tag];
  return 0;
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
  static DynSmemAttr attr = {};
  if (int e = ensure_dyn_smem(contract_kernel<AK, BK>, kMaxSmem, attr)) return e;
  MPGCN_CHECK(p.R >= 1 && p.R <= 8, "R=%d out of range", p.R);
  // N of the UMMA must be a multiple of 16 for M=128: R odd -> N=32R is still a multiple of 32. ok.
  const size_t b_stage = (size_t)p.R * BK * 64;
  const int nres = p.b_res_reps * p.kb_total;
  int stages;
  if (nres) {
    MPGCN_CHECK(p.NT == 1 && p.kb_per_seg == 1 && !p.split_k && !p.b_flat && p.bm.z_mul == 0, "resident B needs a tile-independent B operand");
    MPGCN_CHECK((size_t)nres * b_stage + 2 * (size_t)C::A_STAGE + 1536 <= (size_t)kMaxSmem, "resident B operand does not fit in shared memory");
    stages = (int)((kMaxSmem - 1024 - 512 - (size_t)nres * b_stage) / (size_t)C::A_STAGE);
  } else {
    stages = (int)((kMaxSmem - 1024 - 512) / ((size_t)C::A_STAGE + b_stage));
  }
  // small stages (channel mixes, 8 KB of A per k-block) are HBM-latency bound: keep >= 128 KB of loads in flight per SM
  const int max_stages = ((size_t)C::A_STAGE + (nres ? 0 : b_stage) <= 16384) ? 16 : 8;
  if (stages > max_stages) stages = max_stages;
  MPGCN_CHECK(stages >= 2, "tile does not fit in shared memory");
  p.stages = stages;
  // always request the full opt-in budget: exactly one CTA per SM, so the 512-column TMEM allocation never contends
  const size_t smem = kMaxSmem;
  MPGCN_CHECK(smem_bytes(C::A_STAGE, p.R, BK, stages, nres) <= smem, "internal: smem budget");
  const long long tiles = (long long)p.MT * p.NT * p.Z;
  MPGCN_CHECK(tiles > 0 && tiles < (1ll << 31), "bad tile count %lld", tiles);
  MPGCN_CHECK(p.kb_total > 0 && p.kb_per_seg > 0, "empty contraction");
  int grid = (int)(tiles < device_sm_count() ? tiles : device_sm_count());
  int tag = -1;
  double fl = 0;
  prof_take_next(&tag, &fl);
  prof_begin(tag, fl, stream);
  contract_kernel<AK, BK><<<grid, kThreads1, smem, stream>>>(p);
  prof_end(stream);
  MPGCN_CUDA(cudaGetLastError());
  return 0;
}

template <int AK>
static int launch2_impl(GemmParams& p, cudaStream_t stream) {
  static DynSmemAttr attr = {};
  if (int e = ensure_dyn_smem(contract2_kernel<AK>, kMaxSmem, attr)) return e;
  MPGCN_CHECK(p.R == 8 && !p.split_k, "2-CTA kernel needs R = 8 and no split-K");
  const size_t stage_bytes = 32768;
  int stages = (int)((kMaxSmem - 1024 - 512) / stage_bytes);
  if (stages > 8) stages = 8;
  static int knob_stages = -1, knob_order = -1;
  if (knob_stages < 0) {
    const char* e = getenv("MPGCN_B200_STAGES");
    knob_stages = e ? atoi(e) : 0;
    const char* o = getenv("MPGCN_B200_TILE_ORDER");
    knob_order = o ? atoi(o) : 0;
  }
  if (knob_stages >= 2 && knob_stages < stages) stages = knob_stages;
  p.stages = stages;
  p.nt_fastest = knob_order;
  const long long tiles = (long long)p.MT * p.NT * p.Z;
  MPGCN_CHECK(tiles > 0 && tiles < (1ll << 31), "bad tile count %lld", tiles);
  MPGCN_CHECK(p.kb_total > 0 && p.kb_per_seg > 0, "empty contraction");
  long long pairs = device_sm_count() / 2;
  if (pairs > tiles) pairs = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kMaxSmem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int tag = -1;
  double fl = 0;
  prof_take_next(&tag, &fl);
  prof_begin(tag, fl, stream);
  cudaError_t e = cudaLaunchKernelEx(&cfg, contract2_kernel<AK>, p);
  prof_end(stream);
  if (e != cudaSuccess) {
    set_error("cudaLaunchKernelEx(contract2_kernel) failed: %s", cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

int launch_contract_2cta(int ak, GemmParams& p, cudaStream_t stream) {
  if (ak == A_MN128) return launch2_impl<A_MN128>(p, stream);
  if (ak == A_K128) return launch2_impl<A_K128>(p, stream);
  set_error("no 2-CTA contraction kernel for A kind %d", ak);
  return 1;
}

// the pair tile is 256 rows: worthwhile once a second 128-row tile exists (env MPGCN_B200_NO_2CTA=1 disables it)
bool use_2cta(int n_rows) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("MPGCN_B200_NO_2CTA");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  return !disabled && n_rows > 128;
}

int launch_contract(int ak, int bk, GemmParams& p, cudaStream_t stream) {
  if (ak == A_MN128 && bk == 64) return launch_impl<A_MN128, 64>(p, stream);
  if (ak == A_K128 && bk == 64) return launch_impl<A_K128, 64>(p, stream);
  if (ak == A_K64 && bk == 32) return launch_impl<A_K64, 32>(p, stream);
  if (ak == A_K64 && bk == 64) return launch_impl<A_K64, 64>(p, stream);
  if (ak == A_K64 && bk == 96) return launch_impl<A_K64, 96>(p, stream);
  if (ak == A_MN64 && bk == 64) return launch_impl<A_MN64, 64>(p, stream);
  set_error("no contraction kernel for A kind %d, BK %d", ak, bk);
  return 1;
}

}  // namespace tc
}  // namespace mpgcn
