// Per-OD-cell LSTM on tcgen05 tensor cores (hidden size 32), forward and BPTT backward.
//
// Reference semantics: nn.LSTM(1, 32, 1, batch_first=True) over B*N*N independent cells with zero
// initial state, last hidden state only (/root/reference/MPGCN.py:69,80-87,100-104); gate order i,f,g,o.
//
// Forward, per tile of 128 cells (thread = cell, TMEM lane = cell):
//     gates[128 x 128] = H_{t-1}[128 x 32] (fp16, smem, K-major SW64)  x  W_hh^T (fp16, smem)   -> TMEM fp32
//     epilogue: + b + w_ih*x_t, sigmoid/tanh (ex2/rcp on the SFU), c in registers (fp32), h -> fp16 -> smem
// Two tiles are in flight per CTA (two epilogue warp-groups, one MMA warp each), the step chain of one tile
// hiding behind the other's.  The recurrence rounds h to fp16 only as the MMA operand; c, the gate
// pre-activations and the returned h_T stay fp32.
//
// Backward, per tile: (1) recompute the forward, stashing i,f,g,o,c,h (fp16) per step in an L2-resident
// scratch; (2) walk back in time: da (fp16, power-of-two scaled) is written once to shared memory and read by
// two MMAs through two different descriptors over the same bytes:
//     dh_{t-1}[128 x 32]  = da[128 cells x 128 gates] (K-major)  x  W_hh[128 x 32]          -> TMEM, read back
//     dWext[128 x 64]    += da^T (MN-major) x [h_{t-1} | x_t | 1 | 0..][128 cells x 64]     -> TMEM, accumulated
// over every step and every tile of the CTA; columns 0..31 of dWext are dW_hh, column 32 dW_ih, column 33 db.
#include "kernels.h"

namespace mpgcn {
namespace lstm_tc {

constexpr int C = 32;
constexpr int G4 = 128;
constexpr int CELLS = 128;
constexpr int STASH = 6 * C;   // halves per cell-step: i f g o c h

__device__ __forceinline__ uint32_t sw64_off(int row, int chunk) { return (uint32_t)row * 64u + (uint32_t)((chunk ^ ((row >> 1) & 3)) << 4); }
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ float sigm(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) { return 2.f * __fdividef(1.f, 1.f + __expf(-2.f * x)) - 1.f; }

__device__ __forceinline__ size_t x_index(long long cell, int t, int T, long long NN) {
  const long long b = cell / NN;
  return (size_t)((b * T + t) * NN + (cell - b * NN));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}

// W_hh (fp32 [128][32]) -> fp16 smem tile [128 rows j][64 B], SWIZZLE_64B (serves as K-major B with N=j and as
// MN-major B with K=j); bias = b_ih + b_hh; wih.
__device__ void load_weights(uint8_t* sW, float* s_bias, float* s_wih, const float* w_ih, const float* w_hh, const float* b_ih,
                             const float* b_hh) {
  for (int e = threadIdx.x; e < G4 * 4; e += blockDim.x) {
    const int j = e >> 2, ch = e & 3;
    const float* src = w_hh + j * C + ch * 8;
    st_shared_v4(sW + sw64_off(j, ch), pack2(src[0], src[1]), pack2(src[2], src[3]), pack2(src[4], src[5]), pack2(src[6], src[7]));
  }
  for (int j = threadIdx.x; j < G4; j += blockDim.x) {
    s_bias[j] = b_ih[j] + b_hh[j];
    s_wih[j] = w_ih[j];
  }
}

// One LSTM step for one cell given the gate pre-activation accumulators in TMEM (or zero when !has_mma).
// Updates c[], returns h[]; optionally emits the post-activation gates, c and h (fp16, 16-byte stores).
// Stash layout (per CTA): [t][24 chunks = 6 blocks (i f g o c h) x 4][128 cells][8 halves]: the 32 lanes of a warp
// touch 32 consecutive 16-byte chunks, i.e. every stash load / store is fully coalesced.
constexpr int STASH_CHUNKS = 24;
__device__ __forceinline__ __half* stash_at(__half* base, int t, int chunk) {
  return base + ((size_t)(t * STASH_CHUNKS + chunk) * CELLS) * 8;
}
__device__ __forceinline__ void st_half8(__half* dst, const float* v) {
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}

template <bool STASH_OUT>
__device__ __forceinline__ void cell_step(uint32_t t_row, bool has_mma, float xv, const float* s_bias, const float* s_wih, float (&c)[C],
                                          float (&h)[C], __half* stash, int t_stash) {
  uint32_t r[32];
  float ig[C];
  // input gate
  if (has_mma) { tmem_ld_32x32(t_row + 0 * C, r); tmem_ld_wait(); }
#pragma unroll
  for (int u = 0; u < C; ++u) ig[u] = sigm((has_mma ? __uint_as_float(r[u]) : 0.f) + fmaf(s_wih[u], xv, s_bias[u]));
  if (STASH_OUT) {
#pragma unroll
    for (int q = 0; q < 4; ++q) st_half8(stash_at(stash, t_stash, 0 * 4 + q), ig + 8 * q);
  }
  // cell candidate
  if (has_mma) { tmem_ld_32x32(t_row + 2 * C, r); tmem_ld_wait(); }
  {
    float g[C];
#pragma unroll
    for (int u = 0; u < C; ++u) g[u] = tanh_((has_mma ? __uint_as_float(r[u]) : 0.f) + fmaf(s_wih[2 * C + u], xv, s_bias[2 * C + u]));
    if (STASH_OUT) {
#pragma unroll
      for (int q = 0; q < 4; ++q) st_half8(stash_at(stash, t_stash, 2 * 4 + q), g + 8 * q);
    }
#pragma unroll
    for (int u = 0; u < C; ++u) ig[u] *= g[u];
  }
  // forget gate
  if (has_mma) { tmem_ld_32x32(t_row + 1 * C, r); tmem_ld_wait(); }
  {
    float f[C];
#pragma unroll
    for (int u = 0; u < C; ++u) f[u] = sigm((has_mma ? __uint_as_float(r[u]) : 0.f) + fmaf(s_wih[C + u], xv, s_bias[C + u]));
    if (STASH_OUT) {
#pragma unroll
      for (int q = 0; q < 4; ++q) st_half8(stash_at(stash, t_stash, 1 * 4 + q), f + 8 * q);
    }
#pragma unroll
    for (int u = 0; u < C; ++u) c[u] = fmaf(f[u], c[u], ig[u]);
  }
  if (STASH_OUT) {
#pragma unroll
    for (int q = 0; q < 4; ++q) st_half8(stash_at(stash, t_stash, 4 * 4 + q), c + 8 * q);
  }
  // output gate
  if (has_mma) { tmem_ld_32x32(t_row + 3 * C, r); tmem_ld_wait(); }
  {
    float o[C];
#pragma unroll
    for (int u = 0; u < C; ++u) o[u] = sigm((has_mma ? __uint_as_float(r[u]) : 0.f) + fmaf(s_wih[3 * C + u], xv, s_bias[3 * C + u]));
    if (STASH_OUT) {
#pragma unroll
      for (int q = 0; q < 4; ++q) st_half8(stash_at(stash, t_stash, 3 * 4 + q), o + 8 * q);
    }
#pragma unroll
    for (int u = 0; u < C; ++u) h[u] = o[u] * tanh_(c[u]);
  }
  if (STASH_OUT) {
#pragma unroll
    for (int q = 0; q < 4; ++q) st_half8(stash_at(stash, t_stash, 5 * 4 + q), h + 8 * q);
  }
}

__device__ __forceinline__ void write_h_tile(uint8_t* sH, int row, const float (&h)[C]) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch)
    st_shared_v4(sH + sw64_off(row, ch), pack2(h[8 * ch], h[8 * ch + 1]), pack2(h[8 * ch + 2], h[8 * ch + 3]),
                 pack2(h[8 * ch + 4], h[8 * ch + 5]), pack2(h[8 * ch + 6], h[8 * ch + 7]));
}

// ---------------------------------------------------------------------------------------
// forward: 2 tile groups x 4 epilogue warps + 2 MMA warps
// ---------------------------------------------------------------------------------------
constexpr int FWD_THREADS = 320;

__global__ void __launch_bounds__(FWD_THREADS, 1)
lstm_fwd_tc_kernel(const float* __restrict__ x_seq, const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                   const float* __restrict__ b_ih, const float* __restrict__ b_hh, float* __restrict__ hT, long long cells, int T,
                   long long NN) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                      // 8 KB
  uint8_t* sH = smem + 8192;               // 2 x 8 KB
  float* s_bias = reinterpret_cast<float*>(smem + 3 * 8192);
  float* s_wih = s_bias + G4;
  uint64_t* h_ready = reinterpret_cast<uint64_t*>(s_wih + G4);   // [2]
  uint64_t* g_ready = h_ready + 2;                                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_ready + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_weights(sW, s_bias, s_wih, w_ih, w_hh, b_ih, b_hh);
  if (threadIdx.x == 0) {
    for (int g = 0; g < 2; ++g) { mbar_init(&h_ready[g], CELLS); mbar_init(&g_ready[g], 1); }
    fence_barrier_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();       // weight tile was written with generic stores, will be read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long tiles = (cells + CELLS - 1) / CELLS;

  if (warp >= 8) {
    // ---------------- MMA warp of group g ----------------
    const int g = warp - 8;
    const uint32_t idesc = umma_idesc_f16(128, G4, 0, 0);
    const uint64_t hi = umma_desc_hi(512, 4u);
    const uint32_t a_addr = smem_u32(sH + g * 8192), b_addr = smem_u32(sW);
    uint32_t ph = 0;
    for (long long tile = (long long)blockIdx.x * 2 + g; tile < tiles; tile += (long long)gridDim.x * 2) {
      for (int t = 1; t < T; ++t) {
        mbar_wait(&h_ready[g], ph);
        ph ^= 1u;
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_f16(tmem_base + g * G4, umma_desc(hi, a_addr + k * 32, 16), umma_desc(hi, b_addr + k * 32, 16), idesc, k > 0 ? 1u : 0u);
          umma_commit(&g_ready[g]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue group g: thread = cell ----------------
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)g * G4;
    uint8_t* myH = sH + g * 8192;
    uint32_t ph = 0;
    for (long long tile = (long long)blockIdx.x * 2 + g; tile < tiles; tile += (long long)gridDim.x * 2) {
      const long long cell = tile * CELLS + row;
      const bool live = cell < cells;
      float c[C], h[C];
#pragma unroll
      for (int u = 0; u < C; ++u) c[u] = 0.f;
      float xv = live ? x_seq[x_index(cell, 0, T, NN)] : 0.f;
      for (int t = 0; t < T; ++t) {
        const float xn = (live && t + 1 < T) ? x_seq[x_index(cell, t + 1, T, NN)] : 0.f;   // prefetch next step's input
        if (t > 0) {
          mbar_wait(&g_ready[g], ph);
          ph ^= 1u;
          tc_fence_after();
        }
        cell_step<false>(t_row, t > 0, xv, s_bias, s_wih, c, h, nullptr, 0);
        if (t + 1 < T) {
          write_h_tile(myH, row, h);
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(&h_ready[g]);
        }
        xv = xn;
      }
      if (live) {
        float4* dst = reinterpret_cast<float4*>(hT + (size_t)cell * C);
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ---------------------------------------------------------------------------------------
// backward: 4 epilogue warps + 1 MMA warp per CTA, two CTAs per SM (the per-tile chain is latency bound:
// stash loads -> math -> MMA round trip; a second resident CTA fills the gaps)
// ---------------------------------------------------------------------------------------
constexpr int BWD_THREADS = 160;
constexpr int DA_BYTES = 32768;     // [128 cells][128 gates] fp16 as two [128][64] SW128 sub-tiles
constexpr int HX_BYTES = 16384;     // [128 cells][64] fp16, SW128

__global__ void __launch_bounds__(BWD_THREADS, 2)
lstm_bwd_tc_kernel(const float* __restrict__ x_seq, const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                   const float* __restrict__ b_ih, const float* __restrict__ b_hh, const float* __restrict__ d_hT,
                   float* __restrict__ d_w_ih, float* __restrict__ d_w_hh, float* __restrict__ d_b, float* __restrict__ d_x,
                   __half* __restrict__ scratch, const float* __restrict__ scale2, long long cells, int T, long long NN) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sDA = smem;                               // 32 KB (single buffer: MMA2 of step t retires long before step t-1's math ends)
  uint8_t* sHX = smem + DA_BYTES;                    // 16 KB
  uint8_t* sW = sHX + HX_BYTES;                      // 8 KB
  uint8_t* sH = sW + 8192;                           // 8 KB
  float* s_bias = reinterpret_cast<float*>(sH + 8192);
  float* s_wih = s_bias + G4;
  uint64_t* h_ready = reinterpret_cast<uint64_t*>(s_wih + G4);
  uint64_t* g_ready = h_ready + 1;
  uint64_t* da_ready = g_ready + 1;   // [2]
  uint64_t* da_free = da_ready + 2;   // [2]
  uint64_t* dh_ready = da_free + 2;
  uint64_t* acc_done = dh_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_weights(sW, s_bias, s_wih, w_ih, w_hh, b_ih, b_hh);
  if (threadIdx.x == 0) {
    mbar_init(h_ready, CELLS); mbar_init(g_ready, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&da_ready[b], CELLS); mbar_init(&da_free[b], 1); }
    mbar_init(dh_ready, 1); mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t TM_GATES = tmem_base, TM_DH = tmem_base + 128, TM_DW = tmem_base + 160;
  const long long tiles = (cells + CELLS - 1) / CELLS;
  const float S = scale2[0], invS = scale2[1];

  if (warp == 4) {
    // ---------------- MMA warp ----------------
    const uint32_t id_gates = umma_idesc_f16(128, G4, 0, 0);
    const uint32_t id_dh = umma_idesc_f16(128, 32, 0, 1);      // A = da K-major, B = W_hh MN-major
    const uint32_t id_dw = umma_idesc_f16(128, 64, 1, 1);      // A = da MN-major, B = [h|x|1] MN-major
    const uint64_t hi64 = umma_desc_hi(512, 4u), hi128 = umma_desc_hi(1024, 2u);
    const uint32_t w_addr = smem_u32(sW), h_addr = smem_u32(sH);
    uint32_t ph_h = 0, ph_da[2] = {0, 0};
    bool first_dw = true;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int t = 1; t < T; ++t) {           // forward recompute
        mbar_wait(h_ready, ph_h);
        ph_h ^= 1u;
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_f16(TM_GATES, umma_desc(hi64, h_addr + k * 32, 16), umma_desc(hi64, w_addr + k * 32, 16), id_gates, k > 0 ? 1u : 0u);
          umma_commit(g_ready);
        }
        __syncwarp();
      }
      for (int t = T - 1; t >= 0; --t) {      // backward through time
        const int buf = 0;
        mbar_wait(&da_ready[buf], ph_da[buf]);
        ph_da[buf] ^= 1u;
        tc_fence_after();
        if (lane == 0) {
          const uint32_t da_addr = smem_u32(sDA + buf * DA_BYTES), hx_addr = smem_u32(sHX + buf * HX_BYTES);
          if (t > 0) {
            // dh_{t-1} = da (K-major, two 64-gate sub-tiles of [128 cells][128 B]) x W_hh (rows j, K step = 16 rows x 64 B)
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_f16(TM_DH, umma_desc(hi128, da_addr + (k >> 2) * 16384 + (k & 3) * 32, 16), umma_desc(hi64, w_addr + k * 1024, 2048),
                       id_dh, k > 0 ? 1u : 0u);
            umma_commit(dh_ready);
          }
          // dWext += da^T (MN-major: k rows = cells, 2 m-chunks of 64 gates 16 KB apart) x [h|x|1] (k rows = cells, 128 B)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16(TM_DW, umma_desc(hi128, da_addr + k * 2048, 16384), umma_desc(hi128, hx_addr + k * 2048, 16384), id_dw,
                     (first_dw && k == 0) ? 0u : 1u);
          first_dw = false;
          umma_commit(&da_free[buf]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) umma_commit(acc_done);
    __syncwarp();
  } else {
    // ---------------- epilogue warps: thread = cell (row) / gate row j for the final flush ----------------
    const int row = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    __half* my_stash = scratch + (size_t)blockIdx.x * T * CELLS * STASH + (size_t)row * 8;   // see stash_at()
    uint32_t ph_g = 0, ph_dh = 0, ph_free[2] = {0, 0};
    int free_uses[2] = {0, 0};
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const long long cell = tile * CELLS + row;
      const bool live = cell < cells;
      // ---- (1) recompute forward, stash per step ----
      {
        float c[C], h[C];
#pragma unroll
        for (int u = 0; u < C; ++u) c[u] = 0.f;
        for (int t = 0; t < T; ++t) {
          const float xv = live ? x_seq[x_index(cell, t, T, NN)] : 0.f;
          if (t > 0) {
            mbar_wait(g_ready, ph_g);
            ph_g ^= 1u;
            tc_fence_after();
          }
          cell_step<true>(TM_GATES + lane_base, t > 0, xv, s_bias, s_wih, c, h, my_stash, t);
          if (t + 1 < T) {
            write_h_tile(sH, row, h);
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(h_ready);
          }
        }
      }
      // ---- (2) backward through time ----
      float dh[C], dc[C];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) v = *reinterpret_cast<const float4*>(d_hT + (size_t)cell * C + 4 * q);
        dh[4 * q] = v.x * S; dh[4 * q + 1] = v.y * S; dh[4 * q + 2] = v.z * S; dh[4 * q + 3] = v.w * S;
      }
#pragma unroll
      for (int u = 0; u < C; ++u) dc[u] = 0.f;
      for (int t = T - 1; t >= 0; --t) {
        const int buf = 0;
        if (free_uses[buf] > 0) {          // the MMAs that read this buffer two steps ago must have retired
          mbar_wait(&da_free[buf], ph_free[buf]);
          ph_free[buf] ^= 1u;
        }
        free_uses[buf]++;
        const float xv = live ? x_seq[x_index(cell, t, T, NN)] : 0.f;
        uint8_t* da_t = sDA + buf * DA_BYTES;
        uint8_t* hx_t = sHX + buf * HX_BYTES;
        float dx_acc = 0.f;
        // 8 hidden units at a time: one 16-byte chunk of each gate block
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float di[8], df[8], dg[8], d_o[8];
          const uint4 vi = *reinterpret_cast<const uint4*>(stash_at(my_stash, t, 0 * 4 + q));
          const uint4 vf = *reinterpret_cast<const uint4*>(stash_at(my_stash, t, 1 * 4 + q));
          const uint4 vg = *reinterpret_cast<const uint4*>(stash_at(my_stash, t, 2 * 4 + q));
          const uint4 vo = *reinterpret_cast<const uint4*>(stash_at(my_stash, t, 3 * 4 + q));
          const uint4 vc = *reinterpret_cast<const uint4*>(stash_at(my_stash, t, 4 * 4 + q));
          uint4 vcp = make_uint4(0, 0, 0, 0);
          if (t > 0) vcp = *reinterpret_cast<const uint4*>(stash_at(my_stash, t - 1, 4 * 4 + q));
          const __half* hi_ = reinterpret_cast<const __half*>(&vi);
          const __half* hf = reinterpret_cast<const __half*>(&vf);
          const __half* hg = reinterpret_cast<const __half*>(&vg);
          const __half* ho = reinterpret_cast<const __half*>(&vo);
          const __half* hc = reinterpret_cast<const __half*>(&vc);
          const __half* hcp = reinterpret_cast<const __half*>(&vcp);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int u = 8 * q + e;
            const float ig = __half2float(hi_[e]), fg = __half2float(hf[e]), gg = __half2float(hg[e]), og = __half2float(ho[e]);
            const float tcv = tanh_(__half2float(hc[e]));
            const float cp = __half2float(hcp[e]);
            const float dhv = dh[u];
            const float dcv = fmaf(dhv * og, 1.f - tcv * tcv, dc[u]);
            d_o[e] = dhv * tcv * og * (1.f - og);
            di[e] = dcv * gg * ig * (1.f - ig);
            df[e] = dcv * cp * fg * (1.f - fg);
            dg[e] = dcv * ig * (1.f - gg * gg);
            dc[u] = dcv * fg;
            if (d_x != nullptr)
              dx_acc += di[e] * s_wih[u] + df[e] * s_wih[C + u] + dg[e] * s_wih[2 * C + u] + d_o[e] * s_wih[3 * C + u];
          }
          // gate j = blk*32 + u  ->  sub-tile (j >> 6), 16-byte chunk ((j & 63) >> 3)
#define MPGCN_ST_DA(blk, arr)                                                                                      \
  st_shared_v4(da_t + (((blk) * 32 + 8 * q) >> 6) * 16384 + sw128_off(row, (((blk) * 32 + 8 * q) & 63) >> 3),     \
               pack2(arr[0], arr[1]), pack2(arr[2], arr[3]), pack2(arr[4], arr[5]), pack2(arr[6], arr[7]))
          MPGCN_ST_DA(0, di);
          MPGCN_ST_DA(1, df);
          MPGCN_ST_DA(2, dg);
          MPGCN_ST_DA(3, d_o);
#undef MPGCN_ST_DA
          // h_{t-1} chunk q of the [h | x | 1 | 0] row
          uint4 vh = make_uint4(0, 0, 0, 0);
          if (t > 0) vh = *reinterpret_cast<const uint4*>(stash_at(my_stash, t - 1, 5 * 4 + q));
          *reinterpret_cast<uint4*>(hx_t + sw128_off(row, q)) = vh;
        }
        st_shared_v4(hx_t + sw128_off(row, 4), pack2(xv, 1.f), 0u, 0u, 0u);
        st_shared_v4(hx_t + sw128_off(row, 5), 0u, 0u, 0u, 0u);
        st_shared_v4(hx_t + sw128_off(row, 6), 0u, 0u, 0u, 0u);
        st_shared_v4(hx_t + sw128_off(row, 7), 0u, 0u, 0u, 0u);
        if (d_x != nullptr && live) d_x[x_index(cell, t, T, NN)] = dx_acc * invS;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&da_ready[buf]);
        if (t > 0) {
          mbar_wait(dh_ready, ph_dh);
          ph_dh ^= 1u;
          tc_fence_after();
          uint32_t r[32];
          tmem_ld_32x32(TM_DH + lane_base, r);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < C; ++u) dh[u] = __uint_as_float(r[u]);
          tc_fence_before();
        }
      }
    }
    // ---- flush the weight-gradient accumulator: lane = gate row j ----
    mbar_wait(acc_done, 0);
    tc_fence_after();
    if (tiles > (long long)blockIdx.x) {
      uint32_t r[32];
      tmem_ld_32x32(TM_DW + lane_base, r);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < C; ++k) atomicAdd(&d_w_hh[row * C + k], __uint_as_float(r[k]) * invS);
      tmem_ld_32x32(TM_DW + lane_base + 32, r);
      tmem_ld_wait();
      atomicAdd(&d_w_ih[row], __uint_as_float(r[0]) * invS);
      atomicAdd(&d_b[row], __uint_as_float(r[1]) * invS);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

__global__ void copy_vec_kernel(const float* src, float* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace lstm_tc

// ---------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------
bool lstm_tc_supported(int T, int C) { return C == 32 && T >= 1 && T <= 256; }

static int lstm_bwd_grid(long long cells) {
  const long long tiles = (cells + lstm_tc::CELLS - 1) / lstm_tc::CELLS;
  long long g = 2LL * device_sm_count();
  return (int)(g < tiles ? g : tiles);
}

size_t lstm_tc_bwd_workspace_bytes(int B, int T, long long NN) {
  const long long cells = (long long)B * NN;
  return 1024 + (size_t)lstm_bwd_grid(cells) * T * lstm_tc::CELLS * lstm_tc::STASH * sizeof(__half);
}

int lstm_last_forward_tc(const float* x_seq, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* hT,
                         int B, int T, long long NN, cudaStream_t st) {
  using namespace lstm_tc;
  const long long cells = (long long)B * NN;
  const size_t smem = 1024 + 3 * 8192 + 2 * G4 * sizeof(float) + 128;
  static bool attr = false;
  if (!attr) {
    MPGCN_CUDA(cudaFuncSetAttribute(lstm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  const long long tiles = (cells + CELLS - 1) / CELLS;
  long long grid = device_sm_count();
  if (grid > (tiles + 1) / 2) grid = (tiles + 1) / 2;
  prof_begin(PROF_LSTM_FWD, 8.0 * C * (C + 1) * (double)cells * T, st);
  // 160 KB of dynamic smem requested on purpose: one CTA per SM, so its 256-column TMEM allocation never waits
  lstm_fwd_tc_kernel<<<(unsigned)grid, FWD_THREADS, smem > 160 * 1024 ? smem : 160 * 1024, st>>>(x_seq, w_ih, w_hh, b_ih, b_hh, hT, cells,
                                                                                                  T, NN);
  prof_end(st);
  MPGCN_CUDA(cudaGetLastError());
  return 0;
}

int lstm_last_backward_tc(const float* x_seq, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                          const float* d_hT, float* d_w_ih, float* d_w_hh, float* d_b_ih, float* d_b_hh, float* d_x, int B, int T,
                          long long NN, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace lstm_tc;
  const long long cells = (long long)B * NN;
  MPGCN_CHECK(ws != nullptr && ws_bytes >= lstm_tc_bwd_workspace_bytes(B, T, NN), "lstm backward: workspace too small (%zu < %zu)",
              ws_bytes, lstm_tc_bwd_workspace_bytes(B, T, NN));
  MPGCN_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "lstm backward: workspace must be 256-byte aligned");
  float* scale2 = static_cast<float*>(ws);
  __half* scratch = reinterpret_cast<__half*>(static_cast<uint8_t*>(ws) + 1024);
  if (int e = grad_scale_prepare(d_hT, (size_t)cells * C, scale2, st)) return e;
  MPGCN_CUDA(cudaMemsetAsync(d_w_ih, 0, sizeof(float) * G4, st));
  MPGCN_CUDA(cudaMemsetAsync(d_w_hh, 0, sizeof(float) * G4 * C, st));
  MPGCN_CUDA(cudaMemsetAsync(d_b_ih, 0, sizeof(float) * G4, st));
  const size_t smem = 1024 + DA_BYTES + HX_BYTES + 2 * 8192 + 2 * G4 * sizeof(float) + 256;
  const int kBwdSmem = 100 * 1024;      // exactly two CTAs per SM (2 x 256 TMEM columns)
  static bool attr = false;
  if (!attr) {
    MPGCN_CUDA(cudaFuncSetAttribute(lstm_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    attr = true;
  }
  MPGCN_CHECK(smem <= (size_t)kBwdSmem, "internal: lstm backward smem");
  const int grid = lstm_bwd_grid(cells);
  prof_begin(PROF_LSTM_BWD, 16.0 * C * (C + 1) * (double)cells * T, st);
  lstm_bwd_tc_kernel<<<grid, BWD_THREADS, kBwdSmem, st>>>(x_seq, w_ih, w_hh, b_ih, b_hh, d_hT, d_w_ih, d_w_hh, d_b_ih, d_x, scratch, scale2,
                                                            cells, T, NN);
  prof_end(st);
  MPGCN_CUDA(cudaGetLastError());
  prof_count(PROF_ELEMENTWISE);
  copy_vec_kernel<<<1, G4, 0, st>>>(d_b_ih, d_b_hh, G4);
  MPGCN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mpgcn
