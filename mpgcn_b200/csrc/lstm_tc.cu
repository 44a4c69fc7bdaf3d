// Per-OD-cell LSTM on tcgen05 tensor cores (hidden size 32), forward and BPTT backward.
//
// Reference semantics: nn.LSTM(1, 32, 1, batch_first=True) over B*N*N independent cells with zero
// initial state, last hidden state only (/root/reference/MPGCN.py:69,80-87,100-104); gate order i,f,g,o.
//
// Two kernels (details at each):
//   lstm_fwd_tc_kernel<SAVE>       forward; SAVE additionally stores c_t, h_t (fp16) of every step for training
//   lstm_bwd_saved_tc_kernel<2>    backward from that saved state: one reverse walk, gate MMA prefetched a step ahead
// (a backward call that comes without saved state first re-runs the SAVE forward into its workspace)
// Tile = 128 cells = the 128 TMEM lanes.  The gate GEMM of a step, affine part and ex2 scaling included, is ONE MMA
// [128 cells x 48] . [48 x 128 gates] whose fp32 accumulator is the exponent argument of the activation; activations
// cost 7 SFU operations per hidden unit and step.  The recurrence rounds h to fp16 only as MMA operand; c, the gate
// arguments and the returned h_T stay fp32.  Both kernels use 8-warp CTAs without a dedicated MMA warp so that
// two CTAs stay resident per SM at 128 registers per thread (see lstm_bwd_saved_tc_kernel).
#include "kernels.h"

#include <stdlib.h>

namespace mpgcn {
namespace lstm_tc {

constexpr int C = 32;
constexpr int UN = 16;          // hidden units per thread
constexpr int G4 = 128;
constexpr int CELLS = 128;
constexpr int DA_BYTES = 32768;     // [128 cells][128 gates] fp16 as two [128][64] SW128 sub-tiles
constexpr int HX_BYTES = 16384;     // [128 cells][64] fp16, SW128
constexpr int WX_BYTES = 16384;     // [128 gates][64] fp16, SW128

__device__ __forceinline__ uint32_t sw64_off(int row, int chunk) { return (uint32_t)row * 64u + (uint32_t)((chunk ^ ((row >> 1) & 3)) << 4); }
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

// SFU primitives (2 ulp each); ex2 saturates to 0 / +inf and rcp(inf) = 0, which are the limits the activations need.
__device__ __forceinline__ float ex2_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 2^x WITHOUT the SFU, for x <= 40: round-to-nearest split x = n + f through the 1.5 * 2^23 trick, degree-6 polynomial of 2^f on
// [-0.5, 0.5] (Cephes exp2f coefficients; measured max relative error 1.0e-7, the SFU's ex2.approx is 2 ulp = 2.4e-7), 2^n added into
// the exponent field.  11 FMA / ALU-pipe instructions instead of one of the 7 SFU operations per unit and step: the forward kernel
// is SFU-bound (MIO throttle, 72 % XU) with half of its issue slots free (profiles/ncu_lstm_r2.txt).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = 1.535336188319500e-4f;
  p = fmaf(p, f, 1.339887440266574e-3f);
  p = fmaf(p, f, 9.618437357674640e-3f);
  p = fmaf(p, f, 5.550332471162809e-2f);
  p = fmaf(p, f, 2.402264791363012e-1f);
  p = fmaf(p, f, 6.931472028550421e-1f);
  p = fmaf(p, f, 1.f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot for two IEEE operations).  The activation math of two
// neighbouring hidden units is identical, so everything on the FMA pipe is done on pairs; the SFU operations (ex2, rcp) and the
// clamps stay scalar.  ~21 FMA-pipe instructions per unit and step become ~10.5: the kernels are co-limited by issue slots
// (51 % / 61 % used) and the SFU queue (profiles/ncu_lstm_r2.txt).
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 f2_mk(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_un(f2 x, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
// 1 + 2^min(a, 40) on a pair (two clamps, two SFU ops, one packed add)
__device__ __forceinline__ f2 f2_one_plus_ex2(float a0, float a1, f2 one) {
  return f2_add(f2_mk(ex2_(fminf(a0, 40.f)), ex2_(fminf(a1, 40.f))), one);
}
// 2^x for a pair on the FMA pipe (see ex2_poly): inputs already clamped to <= 40
__device__ __forceinline__ f2 f2_ex2_poly(float x0, float x1) {
  const f2 x = f2_mk(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const f2 magic = f2_mk(12582912.f, 12582912.f), nmagic = f2_mk(-12582912.f, -12582912.f), m1 = f2_mk(-1.f, -1.f);
  const f2 t = f2_add(x, magic);
  const f2 f = f2_fma(f2_add(t, nmagic), m1, x);                 // x - round(x)  in [-0.5, 0.5]
  f2 p = f2_mk(1.535336188319500e-4f, 1.535336188319500e-4f);
  p = f2_fma(p, f, f2_mk(1.339887440266574e-3f, 1.339887440266574e-3f));
  p = f2_fma(p, f, f2_mk(9.618437357674640e-3f, 9.618437357674640e-3f));
  p = f2_fma(p, f, f2_mk(5.550332471162809e-2f, 5.550332471162809e-2f));
  p = f2_fma(p, f, f2_mk(2.402264791363012e-1f, 2.402264791363012e-1f));
  p = f2_fma(p, f, f2_mk(6.931472028550421e-1f, 6.931472028550421e-1f));
  p = f2_fma(p, f, f2_mk(1.f, 1.f));
  float p0, p1, t0, t1;
  f2_un(p, p0, p1);
  f2_un(t, t0, t1);
  return f2_mk(__int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23)), __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23)));
}
template <bool USE_POLY>
__device__ __forceinline__ f2 f2_one_plus_ex2_sel(float a0, float a1, f2 one) {
  if (USE_POLY) return f2_add(f2_ex2_poly(fminf(a0, 40.f), fminf(a1, 40.f)), one);
  return f2_one_plus_ex2(a0, a1, one);
}
__device__ __forceinline__ f2 f2_rcp(f2 x) {
  float a, b;
  f2_un(x, a, b);
  return f2_mk(rcp_(a), rcp_(b));
}

// POLY = how many of the five exponentials per unit and step take the polynomial (0: none; 1: tanh(c); 2: tanh(c) and the o gate)
template <int POLY, int WHICH> __device__ __forceinline__ float ex2_sel(float x) { return (WHICH < POLY) ? ex2_poly(x) : ex2_(x); }

// x enters the gate MMA as fp16 hi + lo (exact to ~22 bits for |x| < 65504); both parts saturate instead of overflowing to inf,
// so larger inputs give finite (saturated-gate) results rather than NaN
__device__ __forceinline__ float x_split_hi(float x) { return __half2float(__float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f))); }
__device__ __forceinline__ float x_split_lo(float x, float hi) { return fminf(fmaxf(x - hi, -65504.f), 65504.f); }

// x_seq is [B][T][NN]: element (cell, t) = x_base(cell) + t * NN; the 64-bit division is done once per tile
__device__ __forceinline__ size_t x_base(long long cell, int T, long long NN) {
  const long long b = cell / NN;
  return (size_t)(b * T * NN + (cell - b * NN));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __half22float2(h2[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}

// 32 lanes x 16 columns of TMEM -> 16 registers
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// W_hh (fp32 [128][32]) -> fp16 smem tile [128 rows j][64 B], SWIZZLE_64B (serves as K-major B with N=j and as
// MN-major B with K=j); bias = b_ih + b_hh; wih.
__device__ void load_weights(uint8_t* sW, float* s_bias, float* s_wih, const float* w_ih, const float* w_hh, const float* b_ih,
                             const float* b_hh) {
  for (int e = threadIdx.x; e < G4 * 4; e += blockDim.x) {
    const int j = e >> 2, ch = e & 3;
    *reinterpret_cast<uint4*>(sW + sw64_off(j, ch)) = pack8(w_hh + j * C + ch * 8);
  }
  for (int j = threadIdx.x; j < G4; j += blockDim.x) {
    s_bias[j] = b_ih[j] + b_hh[j];
    s_wih[j] = w_ih[j];
  }
}

// Training state written by the forward kernel and read by lstm_bwd_saved_tc_kernel, per 128-cell tile:
// [t][8 chunks = c (4) then h (4)][128 cells][8 halves], again one 16-byte chunk per lane.  A thread's base pointer already
// includes its row and its unit half, so chunk q (0/1) addresses its c units and chunk 4+q its h units.
constexpr int SAVE_CHUNKS = 8;
__device__ __forceinline__ uint4* save_at(__half* base, int t, int chunk) {
  return reinterpret_cast<uint4*>(base + (size_t)t * (SAVE_CHUNKS * CELLS * 8)) + chunk * CELLS;
}
__device__ __forceinline__ const uint4* save_at(const __half* base, int t, int chunk) {
  return reinterpret_cast<const uint4*>(base + (size_t)t * (SAVE_CHUNKS * CELLS * 8)) + chunk * CELLS;
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
// Per 128-cell tile and step ONE MMA produces the ex2 arguments of all four gates, affine part included:
//     hx_t[cell] = [ h_{t-1} (32) | x_hi  1  x_lo  x_hi  1  0 0 0 | 0 .. ]                    (fp16, 64 columns, SWIZZLE_128B)
//     Wx[j]      = s_j * [ W_hh[j,:] | wih_hi  b_hi  wih_hi  wih_lo  b_lo  0 0 0 | 0 .. ]      s_j = -log2 e (i, f, o), -2 log2 e (g)
//     acc[cell][j] = hx_t . Wx[j] = s_j * (W_hh h_{t-1} + w_ih x_t + b)_j
// x, w_ih and b are split into fp16 hi + lo parts, so the affine part keeps ~22 bits; h_{t-1} is rounded to fp16 (the only
// reduced-precision operand).  Eight warps, two threads per cell (16 hidden units each); no dedicated MMA warp (see the
// register-file note at lstm_bwd_saved_tc_kernel): a step ends in one CTA barrier, then lane 0 of warp 0 issues the MMA.

__device__ void load_weights_ext(uint8_t* sWx, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh) {
  for (int e = threadIdx.x; e < G4 * 8; e += blockDim.x) {
    const int j = e >> 3, ch = e & 7;
    const float sc = ((j >> 5) == 2) ? -2.8853900817779268f : -1.4426950408889634f;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ch < 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = sc * w_hh[j * C + ch * 8 + i];
    } else if (ch == 4) {
      const float wi = sc * w_ih[j], bb = sc * (b_ih[j] + b_hh[j]);
      const float wi_hi = __half2float(__float2half_rn(wi)), b_hi = __half2float(__float2half_rn(bb));
      v[0] = wi_hi; v[1] = b_hi; v[2] = wi_hi; v[3] = wi - wi_hi; v[4] = bb - b_hi;
    }
    *reinterpret_cast<uint4*>(sWx + sw128_off(j, ch)) = pack8(v);
  }
}


constexpr int FWD_THREADS = 256;

template <bool SAVE, int POLY, bool PACK>
__global__ void __launch_bounds__(FWD_THREADS, 2)
lstm_fwd_tc_kernel(const float* __restrict__ x_seq, const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                   const float* __restrict__ b_ih, const float* __restrict__ b_hh, float* __restrict__ hT, __half* __restrict__ saved,
                   long long cells, int T, long long NN) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sWx = smem;                     // 16 KB
  uint8_t* sHX = smem + WX_BYTES;          // 16 KB
  uint64_t* g_ready = reinterpret_cast<uint64_t*>(sHX + HX_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_weights_ext(sWx, w_ih, w_hh, b_ih, b_hh);
  for (int e = threadIdx.x; e < CELLS * 3; e += blockDim.x)      // constant zero columns 40..63 (chunks 5, 6, 7)
    *reinterpret_cast<uint4*>(sHX + sw128_off(e / 3, 5 + e % 3)) = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    mbar_init(g_ready, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long tiles = (cells + CELLS - 1) / CELLS;
  const uint32_t idesc = umma_idesc_f16(128, G4, 0, 0);
  const uint64_t hi128 = umma_desc_hi(1024, 2u);
  const uint32_t a_addr = smem_u32(sHX), b_addr = smem_u32(sWx);

  const int hh = warp >> 2;
  const int row = (warp & 3) * 32 + lane;
  const int u0 = UN * hh;
  const uint32_t t_col = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)u0;
  uint32_t ph = 0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long cell = tile * CELLS + row;
    const bool live = cell < cells;
    float c[UN], h[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) { c[u] = 0.f; h[u] = 0.f; }
    const size_t xb = live ? x_base(cell, T, NN) : 0;
    float xv = live ? x_seq[xb] : 0.f;
    __half* my_save = SAVE ? saved + (size_t)tile * T * (SAVE_CHUNKS * CELLS * 8) + (size_t)(2 * hh) * CELLS * 8 + (size_t)row * 8 : nullptr;
    for (int t = 0; t < T; ++t) {
      // hx_t = [h_{t-1} | x_t ...]: the tile is free (the MMA of step t-1 was waited for before h_{t-1} was computed)
#pragma unroll
      for (int q = 0; q < 2; ++q) *reinterpret_cast<uint4*>(sHX + sw128_off(row, 2 * hh + q)) = pack8(h + 8 * q);
      if (hh == 0) {
        const float x_hi = x_split_hi(xv);
        *reinterpret_cast<uint4*>(sHX + sw128_off(row, 4)) = make_uint4(pack2(x_hi, 1.f), pack2(x_split_lo(xv, x_hi), x_hi), pack2(1.f, 0.f), 0u);
      }
      fence_proxy_async_smem();
      tc_fence_before();                 // also orders this thread's TMEM reads of step t-1 before the MMA that overwrites them
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k)
            umma_f16(tmem_base, umma_desc(hi128, a_addr + k * 32, 16), umma_desc(hi128, b_addr + k * 32, 16), idesc, k > 0 ? 1u : 0u);
          umma_commit(g_ready);
        }
        __syncwarp();
      }
      xv = (live && t + 1 < T) ? x_seq[xb + (size_t)(t + 1) * NN] : 0.f;     // next step's input, requested under the MMA
      mbar_wait(g_ready, ph);
      ph ^= 1u;
      tc_fence_after();
      {
        // Seven SFU operations per hidden unit and step (5 ex2 + 2 rcp): i, g, f share one reciprocal of the product of their
        // three (1 + 2^arg) terms, o and tanh(c) share another.  The accumulators are -log2e * pre (i, f, o) and -2 log2e * pre
        // (g); arguments are clamped from above at 40 (ex2(-big) = 0 is fine), so a triple product stays below 1.4e36; the
        // clamp moves sigmoid / tanh by < 1e-12.
        uint32_t ra[UN], rb[UN];
        tmem_ld_32x16(t_col + 0 * C, ra);
        tmem_ld_32x16(t_col + 2 * C, rb);
        tmem_ld_wait();
        if constexpr (PACK) {             // the same arithmetic on pairs of units (FFMA2 / FMUL2 / FADD2)
          const f2 one = f2_mk(1.f, 1.f), m1 = f2_mk(-1.f, -1.f), k2 = f2_mk(-2.8853900817779268f, -2.8853900817779268f);
          f2 AI[UN / 2], AG[UN / 2], P[UN / 2];
#pragma unroll
          for (int k = 0; k < UN / 2; ++k) {
            AI[k] = f2_one_plus_ex2(__uint_as_float(ra[2 * k]), __uint_as_float(ra[2 * k + 1]), one);
            AG[k] = f2_one_plus_ex2(__uint_as_float(rb[2 * k]), __uint_as_float(rb[2 * k + 1]), one);
          }
          tmem_ld_32x16(t_col + 1 * C, ra);
          tmem_ld_32x16(t_col + 3 * C, rb);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < UN / 2; ++k) {
            const f2 AF = f2_one_plus_ex2_sel<(POLY > 2)>(__uint_as_float(ra[2 * k]), __uint_as_float(ra[2 * k + 1]), one);
            const f2 PIG = f2_mul(AI[k], AG[k]);
            const f2 R = f2_rcp(f2_mul(PIG, AF));
            const f2 GI = f2_mul(R, f2_mul(AG[k], AF));                         // sigmoid(i)
            const f2 GG = f2_fma(f2_add(R, R), f2_mul(AI[k], AF), m1);          // tanh(g)
            const f2 GF = f2_mul(R, PIG);                                       // sigmoid(f)
            f2 Cc = f2_mk(c[2 * k], c[2 * k + 1]);
            Cc = f2_fma(GF, Cc, f2_mul(GI, GG));
            f2_un(Cc, c[2 * k], c[2 * k + 1]);
            P[k] = f2_one_plus_ex2_sel<(POLY > 1)>(__uint_as_float(rb[2 * k]), __uint_as_float(rb[2 * k + 1]), one);
          }
#pragma unroll
          for (int k = 0; k < UN / 2; ++k) {
            float x0, x1;
            f2_un(f2_mul(f2_mk(c[2 * k], c[2 * k + 1]), k2), x0, x1);
            const f2 AC = f2_one_plus_ex2_sel<(POLY > 0)>(x0, x1, one);
            const f2 R = f2_rcp(f2_mul(P[k], AC));
            f2_un(f2_mul(f2_mul(R, AC), f2_fma(f2_add(R, R), P[k], m1)), h[2 * k], h[2 * k + 1]);      // sigmoid(o) * tanh(c)
          }
        } else {
        float p[UN];                      // a_i * a_g, later o's (1 + 2^arg) term
        float ai[UN], ag[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          ai[u] = 1.f + ex2_(fminf(__uint_as_float(ra[u]), 40.f));
          ag[u] = 1.f + ex2_(fminf(__uint_as_float(rb[u]), 40.f));
        }
        tmem_ld_32x16(t_col + 1 * C, ra);
        tmem_ld_32x16(t_col + 3 * C, rb);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const float af = 1.f + ex2_(fminf(__uint_as_float(ra[u]), 40.f));
          const float pig = ai[u] * ag[u];
          const float r = rcp_(pig * af);
          const float gi = r * (ag[u] * af);                    // sigmoid(i)
          const float gg = fmaf(r + r, ai[u] * af, -1.f);       // tanh(g)
          const float gf = r * pig;                             // sigmoid(f)
          c[u] = fmaf(gf, c[u], gi * gg);
          p[u] = 1.f + ex2_sel<POLY, 1>(fminf(__uint_as_float(rb[u]), 40.f));
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const float ac = 1.f + ex2_sel<POLY, 0>(fminf(-2.8853900817779268f * c[u], 40.f));
          const float r = rcp_(p[u] * ac);
          h[u] = (r * ac) * fmaf(r + r, p[u], -1.f);            // sigmoid(o) * tanh(c)
        }
        }
      }
      if (SAVE) {                     // training: c_t and h_t (fp16) for the backward kernel, see save_at()
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          *save_at(my_save, t, q) = pack8(c + 8 * q);
          *save_at(my_save, t, 4 + q) = pack8(h + 8 * q);
        }
      }
    }
    if (live && hT != nullptr) {
      float4* dst = reinterpret_cast<float4*>(hT + (size_t)cell * C + u0);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// ---------------------------------------------------------------------------------------
// backward from the forward kernel's saved c_t / h_t (training path)
// ---------------------------------------------------------------------------------------
// One reverse walk, no forward recompute and no serial forward chain: the gate pre-activations of step t depend only on
// the SAVED h_{t-1} and on x_t, so their MMA is issued one step ahead and overlaps the gradient math of step t+1.  The
// affine part rides in the same MMA: the operand row of a cell is
//     hx_t[cell] = [ h_{t-1} (32) | x_hi  1  x_lo  x_hi  1  0 0 0 | 0 .. ]                    (fp16, 64 columns, SWIZZLE_128B)
// and the weight tile is  Wx[j] = s_j * [ W_hh[j,:] | wih_hi  b_hi  wih_hi  wih_lo  b_lo 0 0 0 | 0 .. ]  with s_j = -log2(e)
// for the sigmoid gates and -2 log2(e) for the tanh gate (x, w_ih, b split into fp16 hi + lo, so the affine part keeps ~22
// bits): the accumulator IS the ex2 argument of the activation.  The same hx_t bytes are the B operand (MN-major) of the
// weight-gradient MMA.  Per step and tile the tensor core runs
//     MMA0: ex2arg_{t-1} = hx_{t-1} (K-major) x Wx^T              (prefetch for the next iteration)
//     MMA1: dh_{t-1}     = da_t (K-major) x W_hh
//     MMA2: dWext       += da_t^T (MN-major) x hx_t              (cols 0..31 dW_hh, 32 + 34 dW_ih, 33 db)
// TMEM: gates 0..127 | dh 128..159 | dWext 160..223 | dc 224..255.
// x4 TMEM load / store (32 lanes x 4 columns)
__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
template <int W> __device__ __forceinline__ void tmem_ld_w(uint32_t taddr, uint32_t (&r)[W]);
template <> __device__ __forceinline__ void tmem_ld_w<8>(uint32_t taddr, uint32_t (&r)[8]) { tmem_ld_32x8(taddr, r); }
template <> __device__ __forceinline__ void tmem_ld_w<4>(uint32_t taddr, uint32_t (&r)[4]) { tmem_ld_32x4(taddr, r); }
template <int W> __device__ __forceinline__ void tmem_st_w(uint32_t taddr, const uint32_t (&r)[W]);
template <> __device__ __forceinline__ void tmem_st_w<8>(uint32_t taddr, const uint32_t (&r)[8]) { tmem_st_32x8(taddr, r); }
template <> __device__ __forceinline__ void tmem_st_w<4>(uint32_t taddr, const uint32_t (&r)[4]) { tmem_st_32x4(taddr, r); }

// TPC = threads per cell: 2 (8 warps, 16 hidden units per thread, <= 128 registers) is what is launched; 4 (16 warps, 8 units per
// thread, <= 64 registers) was measured 12-18 % slower on B200 (5.4 vs 6.0-6.4 ms) and is not instantiated.  Two CTAs are
// resident per SM and a thread makes two passes of PW = UNT / 2 units per step.
// No dedicated MMA warp: with 9 warps per CTA the register file only holds ONE CTA per SM at > 102 registers per thread
// (18 warps -> 5 on one scheduler partition).  Every step ends in one CTA-wide barrier, after which lane 0 of warp 0 issues
// the step's three MMA groups while everybody moves on.
template <int TPC, int POLY, bool PACK>
__global__ void __launch_bounds__(128 * TPC, 2)
lstm_bwd_saved_tc_kernel(const float* __restrict__ x_seq, const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                         const float* __restrict__ b_ih, const float* __restrict__ b_hh, const float* __restrict__ d_hT,
                         float* __restrict__ d_w_ih, float* __restrict__ d_w_hh, float* __restrict__ d_b, float* __restrict__ d_x,
                         const __half* __restrict__ saved, const float* __restrict__ scale2, long long cells, int T, long long NN) {
  constexpr int UNT = C / TPC;          // hidden units per thread
  constexpr int PW = UNT / 2;           // units per pass
  constexpr int NCH = UNT / 8;          // 16-byte chunks of c / h per thread and step
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sDA = smem;                               // 32 KB
  uint8_t* sHX = smem + DA_BYTES;                    // 3 x 16 KB: hx_t in buffer t % 3 (staged while the MMAs of step t+2 may still read theirs)
  uint8_t* sWx = sHX + 3 * HX_BYTES;                 // 16 KB
  uint8_t* sW = sWx + WX_BYTES;                      // 8 KB
  float* s_bias = reinterpret_cast<float*>(sW + 8192);
  float* s_wih = s_bias + G4;
  uint64_t* mma_ready = reinterpret_cast<uint64_t*>(s_wih + G4);     // dh_t and the ex2 arguments of step t are in TMEM
  uint64_t* mma_free = mma_ready + 1;                                  // the weight-gradient MMA has retired: sDA / hx reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_weights(sW, s_bias, s_wih, w_ih, w_hh, b_ih, b_hh);
  load_weights_ext(sWx, w_ih, w_hh, b_ih, b_hh);
  // constant zero columns 40..63 of the hx buffers (chunks 5, 6, 7); chunks 0..4 are rewritten every step
  for (int e = threadIdx.x; e < 3 * CELLS * 3; e += blockDim.x) {
    const int buf = e / (CELLS * 3), r = (e / 3) % CELLS, ch = 5 + e % 3;
    *reinterpret_cast<uint4*>(sHX + buf * HX_BYTES + sw128_off(r, ch)) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    mbar_init(mma_ready, 1);
    mbar_init(mma_free, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t TM_GATES = tmem_base, TM_DH = tmem_base + 128, TM_DW = tmem_base + 160, TM_DC = tmem_base + 224;
  const long long tiles = (cells + CELLS - 1) / CELLS;
  const float S = scale2[0], invS = scale2[1];

  const uint32_t id_gates = umma_idesc_f16(128, G4, 0, 0);
  const uint32_t id_dh = umma_idesc_f16(128, 32, 0, 1);
  const uint32_t id_dw = umma_idesc_f16(128, 64, 1, 1);
  const uint64_t hi64 = umma_desc_hi(512, 4u), hi128 = umma_desc_hi(1024, 2u);
  const uint32_t w_addr = smem_u32(sW), wx_addr = smem_u32(sWx);
  const uint32_t da_addr = smem_u32(sDA), hx_addr = smem_u32(sHX);
  bool first_dw = true;               // meaningful in the issuing thread only

  const int us = warp >> 2;           // unit slice of this thread: units u0 .. u0 + UNT - 1
  const int row = (warp & 3) * 32 + lane;
  const int u0 = UNT * us;
  const int ch0 = u0 >> 3;            // first 16-byte chunk of this thread's units
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t t_g = TM_GATES + lane_base + u0, t_dh = TM_DH + lane_base + u0, t_dc = TM_DC + lane_base + u0;
  uint32_t ph_ready = 0, ph_free = 0;
  bool mma2_pending = false;          // a weight-gradient MMA that reads sDA / an hx buffer may still be in flight

  // stage hx_t: this thread's units of h_{t-1} (or zeros at t = 0) and, from unit slice 0, the x / 1 columns
  auto stage_hx = [&](const __half* my_save, int t, float xt) {
    uint8_t* buf = sHX + (t % 3) * HX_BYTES;
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
      uint8_t* dst = buf + sw128_off(row, ch0 + q);
      if (t > 0) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(save_at(my_save, t - 1, 4 + q)) : "memory");
      } else {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    if (us == 0) {
      const float x_hi = x_split_hi(xt);
      *reinterpret_cast<uint4*>(buf + sw128_off(row, 4)) = make_uint4(pack2(x_hi, 1.f), pack2(x_split_lo(xt, x_hi), x_hi), pack2(1.f, 0.f), 0u);
    }
  };
  // ex2 arguments of step t from hx_t (48 of its 64 columns are live)
  auto issue_gates = [&](int t) {
    const uint32_t a = hx_addr + (uint32_t)(t % 3) * HX_BYTES;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      umma_f16(TM_GATES, umma_desc(hi128, a + k * 32, 16), umma_desc(hi128, wx_addr + k * 32, 16), id_gates, k > 0 ? 1u : 0u);
  };

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long cell = tile * CELLS + row;
    const bool live = cell < cells;
    const size_t xb = live ? x_base(cell, T, NN) : 0;
    // save_at(my_save, t, q) = chunk q of this thread's c_t units, save_at(my_save, t, 4 + q) = of its h_t units
    const __half* my_save = saved + (size_t)tile * T * (SAVE_CHUNKS * CELLS * 8) + (size_t)ch0 * CELLS * 8 + (size_t)row * 8;
    if (mma2_pending) {                // last step of the previous tile
      mbar_wait(mma_free, ph_free);
      ph_free ^= 1u;
      mma2_pending = false;
    }
    stage_hx(my_save, T - 1, live ? x_seq[xb + (size_t)(T - 1) * NN] : 0.f);
    float x_stage = (live && T > 1) ? x_seq[xb + (size_t)(T - 2) * NN] : 0.f;      // x of the hx tile the next step stages
    // seed dh (scaled d_hT) and dc (0) in TMEM
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t r[PW];
#pragma unroll
      for (int e = 0; e < PW; e += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) v = *reinterpret_cast<const float4*>(d_hT + (size_t)cell * C + u0 + PW * q + e);
        r[e] = __float_as_uint(v.x * S); r[e + 1] = __float_as_uint(v.y * S);
        r[e + 2] = __float_as_uint(v.z * S); r[e + 3] = __float_as_uint(v.w * S);
      }
      tmem_st_w<PW>(t_dh + PW * q, r);
#pragma unroll
      for (int e = 0; e < PW; ++e) r[e] = 0u;
      tmem_st_w<PW>(t_dc + PW * q, r);
    }
    tmem_st_wait();
    asm volatile("cp.async.wait_all;" ::: "memory");
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (lane == 0) {
        issue_gates(T - 1);
        umma_commit(mma_ready);
      }
      __syncwarp();
    }
    uint4 vc[NCH];
#pragma unroll
    for (int q = 0; q < NCH; ++q) vc[q] = *save_at(my_save, T - 1, q);
    for (int t = T - 1; t >= 0; --t) {
      uint4 vcp[NCH];
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        vcp[q] = make_uint4(0, 0, 0, 0);
        if (t > 0) vcp[q] = *save_at(my_save, t - 1, q);
      }
      const float x_cur = x_stage;
      x_stage = (live && t > 1) ? x_seq[xb + (size_t)(t - 2) * NN] : 0.f;
      mbar_wait(mma_ready, ph_ready);    // ex2 arguments of this step and (t < T-1) dh_t
      ph_ready ^= 1u;
      tc_fence_after();
      if (t > 0) stage_hx(my_save, t - 1, x_cur);      // buffer (t-1) % 3: last read by the MMAs of step t+2, long retired
      float dx_acc = 0.f;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t ri[PW], rf[PW], rg[PW], ro[PW], rdh[PW], rdc[PW];
        tmem_ld_w<PW>(t_g + 0 * C + PW * q, ri);
        tmem_ld_w<PW>(t_g + 1 * C + PW * q, rf);
        tmem_ld_w<PW>(t_g + 2 * C + PW * q, rg);
        tmem_ld_w<PW>(t_g + 3 * C + PW * q, ro);
        tmem_ld_w<PW>(t_dh + PW * q, rdh);
        tmem_ld_w<PW>(t_dc + PW * q, rdc);
        tmem_ld_wait();
        float fc[PW], fcp[PW];
        if (PW == 8) {
          unpack8(vc[q % NCH], fc); unpack8(vcp[q % NCH], fcp);
        } else {                         // one chunk holds both passes: halves 4q .. 4q+3
          const uint32_t c0 = q ? vc[0].z : vc[0].x, c1 = q ? vc[0].w : vc[0].y;
          const uint32_t p0 = q ? vcp[0].z : vcp[0].x, p1 = q ? vcp[0].w : vcp[0].y;
          const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&c0)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&c1));
          const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&p0)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&p1));
          fc[0] = a0.x; fc[1] = a0.y; fc[2] = a1.x; fc[3] = a1.y;
          fcp[0] = b0.x; fcp[1] = b0.y; fcp[2] = b1.x; fcp[3] = b1.y;
        }
        float di[PW], df[PW], dg[PW], d_o[PW];
        if constexpr (PACK) {            // the same arithmetic on pairs of units (FFMA2 / FMUL2 / FADD2); d_x is accumulated below
          const f2 one = f2_mk(1.f, 1.f), m1 = f2_mk(-1.f, -1.f), k2 = f2_mk(-2.8853900817779268f, -2.8853900817779268f);
#pragma unroll
          for (int e = 0; e < PW; e += 2) {
            const f2 AI = f2_one_plus_ex2(__uint_as_float(ri[e]), __uint_as_float(ri[e + 1]), one);
            const f2 AG = f2_one_plus_ex2(__uint_as_float(rg[e]), __uint_as_float(rg[e + 1]), one);
            const f2 AF = f2_one_plus_ex2_sel<(POLY > 2)>(__uint_as_float(rf[e]), __uint_as_float(rf[e + 1]), one);
            const f2 AO = f2_one_plus_ex2_sel<(POLY > 1)>(__uint_as_float(ro[e]), __uint_as_float(ro[e + 1]), one);
            float x0, x1;
            f2_un(f2_mul(f2_mk(fc[e], fc[e + 1]), k2), x0, x1);
            const f2 AC = f2_one_plus_ex2_sel<(POLY > 0)>(x0, x1, one);
            const f2 PIG = f2_mul(AI, AG);
            const f2 R1 = f2_rcp(f2_mul(PIG, AF)), R2 = f2_rcp(f2_mul(AO, AC));
            const f2 GI = f2_mul(R1, f2_mul(AG, AF)), GG = f2_fma(f2_add(R1, R1), f2_mul(AI, AF), m1), GF = f2_mul(R1, PIG);
            const f2 GO = f2_mul(R2, AC), TC = f2_fma(f2_add(R2, R2), AO, m1);
            const f2 DH = f2_mk(__uint_as_float(rdh[e]), __uint_as_float(rdh[e + 1]));
            const f2 nTC = f2_mul(TC, m1), nGO = f2_mul(GO, m1), nGI = f2_mul(GI, m1), nGF = f2_mul(GF, m1), nGG = f2_mul(GG, m1);
            const f2 DC = f2_fma(f2_mul(DH, GO), f2_fma(nTC, TC, one), f2_mk(__uint_as_float(rdc[e]), __uint_as_float(rdc[e + 1])));
            f2_un(f2_mul(f2_mul(DH, TC), f2_fma(nGO, GO, GO)), d_o[e], d_o[e + 1]);
            f2_un(f2_mul(f2_mul(DC, GG), f2_fma(nGI, GI, GI)), di[e], di[e + 1]);
            f2_un(f2_mul(f2_mul(DC, f2_mk(fcp[e], fcp[e + 1])), f2_fma(nGF, GF, GF)), df[e], df[e + 1]);
            f2_un(f2_mul(f2_mul(DC, GI), f2_fma(nGG, GG, one)), dg[e], dg[e + 1]);
            float c0, c1;
            f2_un(f2_mul(DC, GF), c0, c1);
            rdc[e] = __float_as_uint(c0);
            rdc[e + 1] = __float_as_uint(c1);
          }
          if (d_x != nullptr) {
#pragma unroll
            for (int e = 0; e < PW; ++e) {
              const int u = u0 + PW * q + e;
              dx_acc += di[e] * s_wih[u] + df[e] * s_wih[C + u] + dg[e] * s_wih[2 * C + u] + d_o[e] * s_wih[3 * C + u];
            }
          }
        } else {
#pragma unroll
        for (int e = 0; e < PW; ++e) {
          // accumulators are -log2e * pre (i, f, o) and -2 log2e * pre (g); clamp from above only (ex2(-big) = 0 is fine),
          // which keeps the product of three (1 + 2^arg) terms below 1.4e36
          const float ai = 1.f + ex2_(fminf(__uint_as_float(ri[e]), 40.f));
          const float ag = 1.f + ex2_(fminf(__uint_as_float(rg[e]), 40.f));
          const float af = 1.f + ex2_(fminf(__uint_as_float(rf[e]), 40.f));
          const float ao = 1.f + ex2_sel<POLY, 1>(fminf(__uint_as_float(ro[e]), 40.f));
          const float ac = 1.f + ex2_sel<POLY, 0>(fminf(-2.8853900817779268f * fc[e], 40.f));
          const float pig = ai * ag;
          const float r1 = rcp_(pig * af), r2 = rcp_(ao * ac);      // 7 SFU ops per unit: see the forward kernel
          const float gi = r1 * (ag * af), gg = fmaf(r1 + r1, ai * af, -1.f), gf = r1 * pig;
          const float go = r2 * ac, tcv = fmaf(r2 + r2, ao, -1.f);
          const float dhv = __uint_as_float(rdh[e]);
          const float dcv = fmaf(dhv * go, fmaf(-tcv, tcv, 1.f), __uint_as_float(rdc[e]));
          d_o[e] = (dhv * tcv) * fmaf(-go, go, go);
          di[e] = (dcv * gg) * fmaf(-gi, gi, gi);
          df[e] = (dcv * fcp[e]) * fmaf(-gf, gf, gf);
          dg[e] = (dcv * gi) * fmaf(-gg, gg, 1.f);
          rdc[e] = __float_as_uint(dcv * gf);
          if (d_x != nullptr) {
            const int u = u0 + PW * q + e;
            dx_acc += di[e] * s_wih[u] + df[e] * s_wih[C + u] + dg[e] * s_wih[2 * C + u] + d_o[e] * s_wih[3 * C + u];
          }
        }
        }
        if (q == 0 && mma2_pending) {    // MMA2 of step t+1 must have retired before sDA is overwritten; by now it has
          mbar_wait(mma_free, ph_free);
          ph_free ^= 1u;
          mma2_pending = false;
        }
        // gate j = blk*32 + u0 + PW*q .. : sub-tile (j >> 6), 16-byte chunk ((j & 63) >> 3), byte (j & 7) * 2 inside the chunk
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
          const float* arr = blk == 0 ? di : blk == 1 ? df : blk == 2 ? dg : d_o;
          const int j0 = blk * 32 + u0 + PW * q;
          uint8_t* dst = sDA + (j0 >> 6) * 16384 + sw128_off(row, (j0 & 63) >> 3) + (j0 & 7) * 2;
          if (PW == 8) *reinterpret_cast<uint4*>(dst) = pack8(arr);
          else *reinterpret_cast<uint2*>(dst) = make_uint2(pack2(arr[0], arr[1]), pack2(arr[2], arr[3]));
        }
        tmem_st_w<PW>(t_dc + PW * q, rdc);
      }
      if (d_x != nullptr && live) atomicAdd(&d_x[xb + (size_t)t * NN], dx_acc * invS);
#pragma unroll
      for (int q = 0; q < NCH; ++q) vc[q] = vcp[q];
      tmem_st_wait();
      asm volatile("cp.async.wait_all;" ::: "memory");
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();                   // da_t, hx_{t-1} and dc are in place; every read of this step's gates / dh is done
      if (warp == 0) {
        tc_fence_after();
        if (lane == 0) {
          const uint32_t hx = hx_addr + (uint32_t)(t % 3) * HX_BYTES;
          if (t > 0) {
            // dh_{t-1} = da (K-major, two 64-gate sub-tiles of [128 cells][128 B]) x W_hh (rows j, K step = 16 rows x 64 B)
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_f16(TM_DH, umma_desc(hi128, da_addr + (k >> 2) * 16384 + (k & 3) * 32, 16), umma_desc(hi64, w_addr + k * 1024, 2048),
                       id_dh, k > 0 ? 1u : 0u);
            issue_gates(t - 1);
            umma_commit(mma_ready);
          }
          // dWext += da^T (MN-major: k rows = cells, 2 m-chunks of 64 gates 16 KB apart) x hx_t (k rows = cells, 128 B)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16(TM_DW, umma_desc(hi128, da_addr + k * 2048, 16384), umma_desc(hi128, hx + k * 2048, 16384), id_dw,
                     (first_dw && k == 0) ? 0u : 1u);
          first_dw = false;
          umma_commit(mma_free);
        }
        __syncwarp();
      }
      mma2_pending = true;
    }
  }
  // ---- flush the weight-gradient accumulator: TMEM lane = gate row j, this thread's UNT columns ----
  if (mma2_pending) {
    mbar_wait(mma_free, ph_free);
    tc_fence_after();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t r[PW];
      tmem_ld_w<PW>(TM_DW + lane_base + u0 + PW * q, r);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < PW; ++k) atomicAdd(&d_w_hh[row * C + u0 + PW * q + k], __uint_as_float(r[k]) * invS);
    }
    if (us == 0) {
      uint32_t r[4];
      tmem_ld_32x4(TM_DW + lane_base + 32, r);
      tmem_ld_wait();
      atomicAdd(&d_w_ih[row], (__uint_as_float(r[0]) + __uint_as_float(r[2])) * invS);
      atomicAdd(&d_b[row], __uint_as_float(r[1]) * invS);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
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

// MPGCN_B200_LSTM_POLY = 0 | 1 | 2 | 3: exponentials per unit and step evaluated by ex2_poly (FMA pipe) instead of the SFU.
// Measured (B200, batch 8, N = 1000, T = 12; forward / backward ms per launch): scalar arithmetic 6.88 / 10.03 with POLY = 0 and
// slower with POLY = 1, 2 (7.26 / 10.22, 7.69 / 11.08: the polynomial costs more issue slots than the SFU slots it frees);
// packed f32x2 arithmetic 6.69 / 9.98 (POLY 0), **6.21 / 9.67 (POLY 1, the default)**, 7.06 / 10.01 (POLY 2), 6.82 / 10.70 (POLY 3).
static int lstm_poly_knob() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MPGCN_B200_LSTM_POLY");
    v = e ? atoi(e) : 1;          // measured best with packed arithmetic: tanh(c)'s exponential on the FMA pipe (profiles/lstm_poly_r2.jsonl)
    if (v < 0 || v > 3) v = 1;
  }
  return v;
}

// MPGCN_B200_LSTM_PACK = 0 | 1: FMA-pipe arithmetic of the LSTM kernels on packed fp32 pairs (FFMA2 / FMUL2 / FADD2)
static int lstm_pack_knob() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MPGCN_B200_LSTM_PACK");
    v = e ? (atoi(e) != 0) : 1;   // default on
  }
  return v;
}

static int lstm_grid(long long cells) {
  const long long tiles = (cells + lstm_tc::CELLS - 1) / lstm_tc::CELLS;
  static int per_sm = 0;
  if (per_sm == 0) {
    const char* e = getenv("MPGCN_B200_LSTM_CTAS_PER_SM");      // tuning knob: 1 or 2 (default) resident CTAs per SM
    per_sm = (e && e[0] == '1') ? 1 : 2;
  }
  long long g = (long long)per_sm * device_sm_count();
  return (int)(g < tiles ? g : tiles);
}

// dynamic shared memory requests: just what the kernels carve (registers already limit residency to two CTAs per SM,
// whose TMEM allocations -- 2 x 128 / 2 x 256 columns -- always fit); the rest of the 228 KB stays L1
static const int kLstmFwdSmem = 34 * 1024;
static const int kLstmSavedSmem = 107 * 1024;     // da tile + three hx buffers + both weight tiles

size_t lstm_tc_saved_bytes(int B, int T, long long NN) {
  const long long tiles = ((long long)B * NN + lstm_tc::CELLS - 1) / lstm_tc::CELLS;
  return (size_t)tiles * T * lstm_tc::SAVE_CHUNKS * lstm_tc::CELLS * 16;
}

// without a saved buffer from the forward, the backward first re-runs the (training) forward into its workspace
size_t lstm_tc_bwd_workspace_bytes(int B, int T, long long NN) { return 1024 + align_up(lstm_tc_saved_bytes(B, T, NN), 256); }

int lstm_last_forward_tc(const float* x_seq, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* hT,
                         void* saved, int B, int T, long long NN, cudaStream_t st) {
  using namespace lstm_tc;
  const long long cells = (long long)B * NN;
  MPGCN_CHECK(saved == nullptr || (reinterpret_cast<uintptr_t>(saved) & 15) == 0, "lstm forward: saved buffer must be 16-byte aligned");
  static int fwd_smem = 0;
  if (fwd_smem == 0) {
    const char* e = getenv("MPGCN_B200_LSTM_FWD_SMEM_KB");     // tuning knob: dynamic smem request (L1 carve-out)
    fwd_smem = e ? atoi(e) * 1024 : kLstmFwdSmem;
    if (fwd_smem < kLstmFwdSmem) fwd_smem = kLstmFwdSmem;
  }
  const int poly = lstm_poly_knob();
  using Kern = void (*)(const float*, const float*, const float*, const float*, const float*, float*, __half*, long long, int, long long);
  // variants 0..2: scalar arithmetic with POLY polynomial exponentials; 3..6: packed f32x2 arithmetic with POLY = 0..3
  static const Kern kerns[2][7] = {{lstm_fwd_tc_kernel<false, 0, false>, lstm_fwd_tc_kernel<false, 1, false>, lstm_fwd_tc_kernel<false, 2, false>,
                                    lstm_fwd_tc_kernel<false, 0, true>, lstm_fwd_tc_kernel<false, 1, true>, lstm_fwd_tc_kernel<false, 2, true>,
                                    lstm_fwd_tc_kernel<false, 3, true>},
                                   {lstm_fwd_tc_kernel<true, 0, false>, lstm_fwd_tc_kernel<true, 1, false>, lstm_fwd_tc_kernel<true, 2, false>,
                                    lstm_fwd_tc_kernel<true, 0, true>, lstm_fwd_tc_kernel<true, 1, true>, lstm_fwd_tc_kernel<true, 2, true>,
                                    lstm_fwd_tc_kernel<true, 3, true>}};
  static DynSmemAttr attrs[2][7] = {};
  const int sv = saved ? 1 : 0;
  const int var = lstm_pack_knob() ? 3 + poly : (poly > 2 ? 2 : poly);
  if (int e = ensure_dyn_smem(kerns[sv][var], fwd_smem, attrs[sv][var])) return e;
  prof_begin(PROF_LSTM_FWD, 8.0 * C * (C + 1) * (double)cells * T, st);
  kerns[sv][var]<<<lstm_grid(cells), FWD_THREADS, fwd_smem, st>>>(x_seq, w_ih, w_hh, b_ih, b_hh, hT, static_cast<__half*>(saved), cells, T, NN);
  prof_end(st);
  MPGCN_CUDA(cudaGetLastError());
  return 0;
}

int lstm_last_backward_tc(const float* x_seq, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                          const float* d_hT, float* d_w_ih, float* d_w_hh, float* d_b_ih, float* d_b_hh, float* d_x, const void* saved,
                          int B, int T, long long NN, void* ws, size_t ws_bytes, const float* d_hT_absmax, cudaStream_t st) {
  using namespace lstm_tc;
  const long long cells = (long long)B * NN;
  const size_t need = saved ? 1024 : lstm_tc_bwd_workspace_bytes(B, T, NN);
  MPGCN_CHECK(ws != nullptr && ws_bytes >= need, "lstm backward: workspace too small (%zu < %zu)", ws_bytes, need);
  MPGCN_CHECK(saved == nullptr || (reinterpret_cast<uintptr_t>(saved) & 15) == 0, "lstm backward: saved buffer must be 16-byte aligned");
  MPGCN_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "lstm backward: workspace must be 256-byte aligned");
  float* scale2 = static_cast<float*>(ws);
  if (saved == nullptr) {          // the caller kept no forward state: rebuild it (same kernel, same bits as the training forward)
    void* tmp = static_cast<uint8_t*>(ws) + 1024;
    if (int e = lstm_last_forward_tc(x_seq, w_ih, w_hh, b_ih, b_hh, nullptr, tmp, B, T, NN, st)) return e;
    saved = tmp;
  }
  if (int e = grad_scale_prepare(d_hT, (size_t)cells * C, scale2, d_hT_absmax, st)) return e;
  MPGCN_CUDA(cudaMemsetAsync(d_w_ih, 0, sizeof(float) * G4, st));
  MPGCN_CUDA(cudaMemsetAsync(d_w_hh, 0, sizeof(float) * G4 * C, st));
  MPGCN_CUDA(cudaMemsetAsync(d_b_ih, 0, sizeof(float) * G4, st));
  if (d_x) MPGCN_CUDA(cudaMemsetAsync(d_x, 0, sizeof(float) * (size_t)cells * T, st));
  const int poly = lstm_poly_knob();
  using KernB = void (*)(const float*, const float*, const float*, const float*, const float*, const float*, float*, float*, float*, float*,
                         const __half*, const float*, long long, int, long long);
  static const KernB kernb[7] = {lstm_bwd_saved_tc_kernel<2, 0, false>, lstm_bwd_saved_tc_kernel<2, 1, false>, lstm_bwd_saved_tc_kernel<2, 2, false>,
                                 lstm_bwd_saved_tc_kernel<2, 0, true>, lstm_bwd_saved_tc_kernel<2, 1, true>, lstm_bwd_saved_tc_kernel<2, 2, true>,
                                 lstm_bwd_saved_tc_kernel<2, 3, true>};
  static DynSmemAttr attr_b[7] = {};
  const int varb = lstm_pack_knob() ? 3 + poly : (poly > 2 ? 2 : poly);        // as in the forward
  if (int e = ensure_dyn_smem(kernb[varb], kLstmSavedSmem, attr_b[varb])) return e;
  static_assert(1024 + DA_BYTES + 3 * HX_BYTES + WX_BYTES + 8192 + 2 * G4 * sizeof(float) + 256 <= (size_t)kLstmSavedSmem, "smem");
  prof_begin(PROF_LSTM_BWD, 12.0 * C * (C + 1) * (double)cells * T, st);
  kernb[varb]<<<lstm_grid(cells), 256, kLstmSavedSmem, st>>>(x_seq, w_ih, w_hh, b_ih, b_hh, d_hT, d_w_ih, d_w_hh, d_b_ih, d_x,
                                                            static_cast<const __half*>(saved), scale2, cells, T, NN);
  prof_end(st);
  MPGCN_CUDA(cudaGetLastError());
  prof_count(PROF_ELEMENTWISE);
  copy_vec_kernel<<<1, G4, 0, st>>>(d_b_ih, d_b_hh, G4);
  MPGCN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mpgcn
