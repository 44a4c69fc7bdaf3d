// tcgen05 / TMA contraction engine: one persistent, warp-specialised kernel template that
// evaluates every tensor-core contraction of the BDGCN layer (forward and backward).
//
//   D[z][i][(r,ch)] = alpha * sum_k A_z[k or i major] * B_z[k][(r,ch)]   (+ bias[ch], ReLU)
//
// * fp16 operands (kind::f16), fp32 accumulation in TMEM, M = 128 rows per tile,
//   N = 32*R columns per tile (R <= 8 "channel chunks" of 32 = one 64-byte swizzle row).
// * operands are staged by TMA into a multi-stage shared-memory ring in exactly the
//   canonical UMMA layouts, so no thread ever touches operand bytes:
//     A_MN128 : A is [k][m], m contiguous      (G_d for  Z = X x2 G_d ; G_o flat for  pre = sum G_o^T U)
//     A_K128  : A is [m][k], k contiguous      (G_o for  V = G_o x1 dPre ; G_d for dX)
//     A_K64   : A is [m][32], one 32-wide k block per plane (channel mixes Z->U, V->Y)
//     A_MN64  : A is [k][chunk][32]            (Z for the weight gradient dW = Z^T V)
//     B       : always [k][chunk r][32 ch], 64-byte rows, SWIZZLE_64B, MN-major
// * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one lane),
//   warps 2..9 = epilogue (TMEM -> registers -> global; 2..5 in the 2-CTA kernel).  Two TMEM accumulators
//   (2 x 256 columns) let the epilogue of tile t overlap the MMAs of tile t+1.
//
// Reference math being evaluated: BDGCN.forward, /root/reference/MPGCN.py:24-50, in the
// factored order of SURVEY.md section 7.1.
#pragma once

#include "common.cuh"

namespace mpgcn {
namespace tc {

enum AKind : int { A_MN128 = 0, A_K128 = 1, A_K64 = 2, A_MN64 = 3 };

// tile/k-block -> TMA coordinate map of one operand
struct OperandMap {
  int z_div, z_mod, z_mul;   // z' = ((z / z_div) % z_mod) * z_mul + (seg % seg_mod) * seg_mul + (seg / seg_mod) * seg_hi_mul
  int seg_mod, seg_mul, seg_hi_mul;
  int k_seg;                 // k coordinate (elements) += (seg % seg_mod) * k_seg
};

struct Epilogue {
  void* out;                 // float* or __half*
  __half* out16;             // optional fp16 shadow of a float output (same strides), may be null
  long long sZ, sI, sR;      // element strides of z, row i, chunk r (channel stride is 1)
  int m_valid, r_valid;      // bounds on i and r
  int out_f16;               // 1: out is __half
  int relu;
  const float* bias;         // [32] or null
  float alpha;
  const float* alpha_dev;    // optional device scalar multiplied into alpha (gradient un-scaling), may be null
  float* absmax_out;         // optional device scalar (pre-zeroed): receives max |stored value| (bit pattern, atomicMax)
  // optional diagonal correction (forward only): out[z][i][r][:] += sum_seg delta[(zA*corr_nseg + seg)*m_valid + i] *
  // corr_src[zB*cZ + seg*cSeg + i*cI + r*cR + :], the exact remainder of the support diagonal lost by its fp16 rounding
  const __half* corr_src;
  const float* corr_delta;
  long long cZ, cI, cR, cSeg;
  int corr_nseg;             // <= 8
  // optional PUSH of a partial result into peer memory (origin-row shard, FWD_B): row i belongs to owner i / peer_rows and is
  // stored at peer_out[owner] + peer_slot + z * peer_sZ + (i % peer_rows) * sI (+ r * sR + channel) -- the owner's staging slot
  // for THIS rank, over NVLink when the owner is another GPU.  The transfer rides in the epilogue, tile by tile, under the MMAs
  // of the next tile; peer_g == 0: plain store to `out`.
  float* peer_out[8];
  int peer_g, peer_rows;
  long long peer_slot, peer_sZ;
};

struct alignas(64) GemmParams {
  CUtensorMap a_map;
  CUtensorMap b_map;
  OperandMap am, bm;
  int b_flat;                // 1: B tile = R separate (32 x BK) boxes of a flat [rows][cols] tensor; 2 (pair kernel): 64-column
                             // SWIZZLE_128B boxes (two chunks per box, 128-byte requests)
  int MT, NT, Z;             // tile grid: tile id = (z * NT + nt) * MT + mt
  int R;                     // 32-column chunks per tile
  int kb_total, kb_per_seg;  // k-blocks over all segments / per segment
  int split_k, kb_per_slice; // split-K: z is a k-slice [z*kb_per_slice, ...)
  int stages;
  // resident B (channel mixes): the whole B operand -- b_res_reps * kb_total tiles, index rep * kb_total + kb -- is loaded
  // once per CTA and every A k-block is multiplied by its b_res_reps tiles (the fp16 hi / lo halves of W); 0 = B streams
  // through the ring with A.  Needs NT == 1, kb_per_seg == 1, no split-K, a z-independent B map.
  int b_res_reps;
  int nt_fastest;            // tile order: 0 = m-tiles vary fastest (default), 1 = n-tiles vary fastest
  int z_inner;               // > 1: z = zo * z_inner + zi and zi varies right after the m-tiles (before the n-tiles): the z_inner
                             // contractions that share one B block (the K supports applied to one X16 / dP16 row block) run back to
                             // back, so that block is fetched from HBM once and then served from L2
  Epilogue ep;
};

constexpr int kThreads = 192;        // 2-CTA kernel: producer, MMA, 4 epilogue warps
constexpr int kEpiWarps1 = 8;        // 1-CTA kernel: 8 epilogue warps (two per TMEM lane quarter, alternating chunks): with one warp per
                                     // scheduler the TMEM -> convert -> store chain of the channel mixes (short MMAs, 3 chunks per
                                     // tile) ran at ~0.3 IPC and bounded the kernel at half the HBM rate
constexpr int kThreads1 = 64 + 32 * kEpiWarps1;
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;

template <int AK, int BK>
struct Cfg {
  // A_K64 with BK = 32 P: P planes per k-block, one [128 m][64 B] slab each (a single TMA box over the plane dimension)
  static constexpr int A_STAGE = (AK == A_MN128) ? BK * 256 : (AK == A_K128) ? 128 * 128 : (AK == A_K64) ? 128 * 64 * (BK / 32) : 4 * BK * 64;
  static constexpr bool A_MN = (AK == A_MN128) || (AK == A_MN64);
  // byte advance of the A descriptor start address per UMMA (K = 16)
  static constexpr uint32_t A_KSTEP = (AK == A_MN128) ? 16 * 128 : (AK == A_MN64) ? 16 * 64 : 32;
  static constexpr uint32_t A_LBO = (AK == A_MN128) ? BK * 128 : (AK == A_MN64) ? BK * 64 : 16;
  static constexpr uint32_t A_SBO = (AK == A_MN128 || AK == A_K128) ? 1024 : 512;
  static constexpr uint32_t A_LAYOUT = (AK == A_MN128 || AK == A_K128) ? 2u : 4u;   // SW128 : SW64
  static constexpr uint32_t B_KSTEP = 16 * 64;
  static constexpr uint32_t B_LBO = BK * 64;
  static constexpr uint32_t B_SBO = 512;
  static_assert(AK != A_K128 || BK == 64, "K-major SW128 rows hold exactly 64 halves");
  static_assert(AK != A_K64 || BK % 32 == 0, "K-major SW64 rows hold exactly 32 halves: BK counts whole planes");
  // byte offset of the A descriptor for the k-th UMMA (K = 16) of a k-block
  __host__ __device__ static constexpr uint32_t a_koff(int k) {
    return (AK == A_K64) ? (uint32_t)(k >> 1) * 8192u + (uint32_t)(k & 1) * 32u : (uint32_t)k * A_KSTEP;
  }
  static_assert(BK % 16 == 0, "UMMA K is 16 for fp16");
};

__host__ __device__ inline size_t smem_bytes(int a_stage, int R, int BK, int stages, int b_resident_tiles = 0) {
  const size_t b_stage = (size_t)R * BK * 64;
  return 1024 /*align slack*/ + (size_t)stages * a_stage + (size_t)(b_resident_tiles ? b_resident_tiles : stages) * b_stage +
         512 /*barriers, tmem slot, bias*/;
}

#ifdef __CUDACC__
// tile id -> (m tile, n tile, z); see GemmParams::nt_fastest / z_inner
__device__ __forceinline__ void decode_tile(const GemmParams& p, int t, int& mt, int& nt, int& z) {
  if (p.z_inner > 1) {
    mt = t % p.MT;
    int rest = t / p.MT;
    const int zi = rest % p.z_inner;
    rest /= p.z_inner;
    nt = rest % p.NT;
    z = (rest / p.NT) * p.z_inner + zi;
  } else if (p.nt_fastest) {
    nt = t % p.NT;
    const int rest = t / p.NT;
    mt = rest % p.MT;
    z = rest / p.MT;
  } else {
    mt = t % p.MT;
    const int rest = t / p.MT;
    nt = rest % p.NT;
    z = rest / p.NT;
  }
}

// one 32-byte (full L2 sector) store per lane: sm_100 has 256-bit global stores (STG.256); with 16-byte stores every lane of the
// scattered epilogue patterns (rows 64 B .. 128 KB apart) sent two half-filled sector requests instead of one full one
__device__ __forceinline__ void st_global_256(void* ptr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                              uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5),
               "r"(a6), "r"(a7) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void store_chunk(const Epilogue& ep, void* out_ptr, float alpha, const float* sbias, long long off,
                                            uint32_t (&acc)[32], float& amax) {
  float v[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    float x = __uint_as_float(acc[c]) * alpha;
    if (ep.bias) x += sbias[c];
    if (ep.relu) x = fmaxf(x, 0.f);
    v[c] = x;
  }
  if (ep.absmax_out) {
#pragma unroll
    for (int c = 0; c < 32; ++c) amax = fmaxf(amax, fabsf(v[c]));
  }
  if (ep.out_f16) {        // 64 bytes per row and chunk: two 32-byte stores
    __half* dst = reinterpret_cast<__half*>(out_ptr) + off;
#pragma unroll
    for (int q = 0; q < 2; ++q)
      st_global_256(dst + 16 * q, pack_h2(v[16 * q + 0], v[16 * q + 1]), pack_h2(v[16 * q + 2], v[16 * q + 3]),
                    pack_h2(v[16 * q + 4], v[16 * q + 5]), pack_h2(v[16 * q + 6], v[16 * q + 7]), pack_h2(v[16 * q + 8], v[16 * q + 9]),
                    pack_h2(v[16 * q + 10], v[16 * q + 11]), pack_h2(v[16 * q + 12], v[16 * q + 13]), pack_h2(v[16 * q + 14], v[16 * q + 15]));
  } else {                 // 128 bytes per row and chunk: four 32-byte stores
    float* dst = reinterpret_cast<float*>(out_ptr) + off;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      st_global_256(dst + 8 * q, __float_as_uint(v[8 * q + 0]), __float_as_uint(v[8 * q + 1]), __float_as_uint(v[8 * q + 2]),
                    __float_as_uint(v[8 * q + 3]), __float_as_uint(v[8 * q + 4]), __float_as_uint(v[8 * q + 5]),
                    __float_as_uint(v[8 * q + 6]), __float_as_uint(v[8 * q + 7]));
    if (ep.out16) {
      __half* d16 = ep.out16 + off;
#pragma unroll
      for (int q = 0; q < 2; ++q)
        st_global_256(d16 + 16 * q, pack_h2(v[16 * q + 0], v[16 * q + 1]), pack_h2(v[16 * q + 2], v[16 * q + 3]),
                      pack_h2(v[16 * q + 4], v[16 * q + 5]), pack_h2(v[16 * q + 6], v[16 * q + 7]), pack_h2(v[16 * q + 8], v[16 * q + 9]),
                      pack_h2(v[16 * q + 10], v[16 * q + 11]), pack_h2(v[16 * q + 12], v[16 * q + 13]), pack_h2(v[16 * q + 14], v[16 * q + 15]));
    }
  }
}

template <int AK, int BK>
__global__ void __launch_bounds__(kThreads1, 1) contract_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<AK, BK>;
  constexpr int A_STAGE = C::A_STAGE;
  const int R = p.R;
  const int B_STAGE = R * BK * 64;
  const int S = p.stages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NRES = p.b_res_reps * p.kb_total;       // resident B tiles (0: B goes through the ring)
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)S * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + (size_t)(NRES ? NRES : S) * B_STAGE);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 2;
  uint64_t* bres_full = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_full + 1);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map);
    tma_prefetch_desc(&p.b_map);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kEpiWarps1);
    }
    mbar_init(bres_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (warp == 2) sbias[lane] = p.ep.bias ? p.ep.bias[lane] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.MT * p.NT * p.Z;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (NRES && (int)blockIdx.x < num_tiles) {      // the whole B operand, once
        mbar_arrive_expect_tx(bres_full, (uint32_t)(NRES * B_STAGE));
        for (int sg = 0; sg < NRES; ++sg) {
          const int lo = sg % p.bm.seg_mod, hi = sg / p.bm.seg_mod;
          tma_load_4d(sB + (size_t)sg * B_STAGE, &p.b_map, bres_full, 0, lo * p.bm.k_seg, 0, lo * p.bm.seg_mul + hi * p.bm.seg_hi_mul);
        }
      }
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int mt, nt, z;
        decode_tile(p, t, mt, nt, z);
        const int kb0 = p.split_k ? z * p.kb_per_slice : 0;
        const int kb1 = p.split_k ? min(kb0 + p.kb_per_slice, p.kb_total) : p.kb_total;
        const int zA0 = ((z / p.am.z_div) % p.am.z_mod) * p.am.z_mul;
        const int zB0 = ((z / p.bm.z_div) % p.bm.z_mod) * p.bm.z_mul;
        // segment / k-block counters are advanced incrementally: the channel-mix contractions spend only ~100 clk of MMA
        // per k-block, so integer divisions in this single-thread loop would bound the whole kernel
        int seg = kb0 / p.kb_per_seg;
        int kk = kb0 - seg * p.kb_per_seg;
        int sa_lo = seg % p.am.seg_mod, sa_hi = seg / p.am.seg_mod;
        int sb_lo = seg % p.bm.seg_mod, sb_hi = seg / p.bm.seg_mod;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], (uint32_t)(A_STAGE + (NRES ? 0 : B_STAGE)));
          const int zA = zA0 + sa_lo * p.am.seg_mul + sa_hi * p.am.seg_hi_mul;
          const int zB = zB0 + sb_lo * p.bm.seg_mul + sb_hi * p.bm.seg_hi_mul;
          const int kA = kk * BK + sa_lo * p.am.k_seg;
          const int kB = kk * BK + sb_lo * p.bm.k_seg;
          uint8_t* a_dst = sA + (size_t)stage * A_STAGE;
          uint8_t* b_dst = sB + (size_t)stage * B_STAGE;
          if (AK == A_MN128) {          // dims (m, k, z, 1), two 64-wide m boxes
            tma_load_4d(a_dst, &p.a_map, &full[stage], mt * 128, kA, zA, 0);
            tma_load_4d(a_dst + BK * 128, &p.a_map, &full[stage], mt * 128 + 64, kA, zA, 0);
          } else if (AK == A_K128 || AK == A_K64) {   // dims (k, m, z, 1)
            tma_load_4d(a_dst, &p.a_map, &full[stage], kA, mt * 128, zA, 0);
          } else {                      // A_MN64: dims (ch, k, chunk, z), 4 chunks per tile
            tma_load_4d(a_dst, &p.a_map, &full[stage], 0, kA, mt * 4, zA);
          }
          if (NRES) {
            // B is resident
          } else if (!p.b_flat) {       // dims (ch, k, r, z)
            tma_load_4d(b_dst, &p.b_map, &full[stage], 0, kB, nt * R, zB);
          } else {                      // dims (col, k, z, 1): one 32-column box per chunk
            for (int j = 0; j < R; ++j)
              tma_load_4d(b_dst + (size_t)j * BK * 64, &p.b_map, &full[stage], (nt * R + j) * 32, kB, zB, 0);
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
          if (++kk == p.kb_per_seg) {
            kk = 0;
            if (++sa_lo == p.am.seg_mod) { sa_lo = 0; ++sa_hi; }
            if (++sb_lo == p.bm.seg_mod) { sb_lo = 0; ++sb_hi; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t idesc = umma_idesc_f16(128, 32 * R, C::A_MN ? 1 : 0, 1);
    const uint64_t a_hi = umma_desc_hi(C::A_SBO, C::A_LAYOUT);
    const uint64_t b_hi = umma_desc_hi(C::B_SBO, 4u);
    if (NRES && (int)blockIdx.x < num_tiles) {
      mbar_wait(bres_full, 0);
      tc_fence_after();
    }
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int mt_, nt_, z;
      decode_tile(p, t, mt_, nt_, z);
      const int kb0 = p.split_k ? z * p.kb_per_slice : 0;
      const int kb1 = p.split_k ? min(kb0 + p.kb_per_slice, p.kb_total) : p.kb_total;
      mbar_wait(&tempty[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(sA + (size_t)stage * A_STAGE);
          const int reps = NRES ? p.b_res_reps : 1;
          for (int rep = 0; rep < reps; ++rep) {
            const uint32_t b_addr = smem_u32(sB + (size_t)(NRES ? rep * p.kb_total + kb : stage) * B_STAGE);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = umma_desc(a_hi, a_addr + C::a_koff(k), C::A_LBO);
              const uint64_t bd = umma_desc(b_hi, b_addr + k * C::B_KSTEP, C::B_LBO);
              umma_f16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0 || rep > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);                 // frees this smem stage when the MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ------------------------------ epilogue (warps 2..9) ------------------------------
    const int quarter = warp & 3;     // TMEM lane quarter this warp may access
    const int jpar = (warp - 2) >> 2;  // which of the two warps of this quarter: takes chunks jpar, jpar + 2, ...
    const float alpha = p.ep.alpha_dev ? p.ep.alpha * __ldg(p.ep.alpha_dev) : p.ep.alpha;
    float amax = 0.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int mt, nt, z;
      decode_tile(p, t, mt, nt, z);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int i = mt * 128 + quarter * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * kAccCols;
      long long base = (long long)z * p.ep.sZ + (long long)i * p.ep.sI;
      void* out_ptr = p.ep.out;
      if (p.ep.peer_g > 0 && i < p.ep.m_valid) {      // push this row of the partial result into its owner's staging slot
        const int owner = i / p.ep.peer_rows;
        out_ptr = p.ep.peer_out[owner];
        base = p.ep.peer_slot + (long long)z * p.ep.peer_sZ + (long long)(i - owner * p.ep.peer_rows) * p.ep.sI;
      }
      float dl[8];
      long long cbase = 0;
      bool any_corr = false;
      if (p.ep.corr_src != nullptr) {
        const int zA = ((z / p.am.z_div) % p.am.z_mod) * p.am.z_mul;
        const int zB = ((z / p.bm.z_div) % p.bm.z_mod) * p.bm.z_mul;
        cbase = (long long)zB * p.ep.cZ + (long long)i * p.ep.cI;
#pragma unroll
        for (int sgi = 0; sgi < 8; ++sgi) {
          dl[sgi] = (sgi < p.ep.corr_nseg && i < p.ep.m_valid)
                        ? __ldg(p.ep.corr_delta + ((long long)zA * p.ep.corr_nseg + sgi) * p.ep.m_valid + i) : 0.f;
          any_corr |= (dl[sgi] != 0.f);
        }
      }
      for (int j = jpar; j < R; j += kEpiWarps1 / 4) {
        uint32_t regs[32];
        tmem_ld_32x32(t_row + (uint32_t)j * 32, regs);
        tmem_ld_wait();
        const int r = nt * R + j;
        if (i < p.ep.m_valid && r < p.ep.r_valid) {
          if (any_corr) {
#pragma unroll
            for (int sgi = 0; sgi < 8; ++sgi) {
              if (sgi < p.ep.corr_nseg && dl[sgi] != 0.f) {
                const uint4* src = reinterpret_cast<const uint4*>(p.ep.corr_src + cbase + (long long)sgi * p.ep.cSeg + (long long)r * p.ep.cR);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 v = __ldg(src + q);
                  const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(h2[e]);
                    regs[8 * q + 2 * e] = __float_as_uint(fmaf(dl[sgi], f.x, __uint_as_float(regs[8 * q + 2 * e])));
                    regs[8 * q + 2 * e + 1] = __float_as_uint(fmaf(dl[sgi], f.y, __uint_as_float(regs[8 * q + 2 * e + 1])));
                  }
                }
              }
            }
          }
          store_chunk(p.ep, out_ptr, alpha, sbias, base + (long long)r * p.ep.sR, regs, amax);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (p.ep.absmax_out) {
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(p.ep.absmax_out), __float_as_uint(amax));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2) for the N^3 contractions: a CTA pair owns a 256 x 256 tile.  Each CTA stages its
// own 128 rows of A and HALF of the B tile (4 of the 8 channel chunks); the pair-wide UMMA reads both halves, so
// per CTA the shared-memory traffic per k-block drops from 48+48 KB (TMA write + MMA read) to 32+32 KB -- the
// 1-CTA kernel is shared-memory-bandwidth bound (192 B/clk needed vs 128 B/clk available at full tensor rate).
// Roles per CTA: warp 0 producer (own A + own B half, bytes credited to the leader's barrier), warp 1 TMEM
// alloc + (leader only) MMA issue with multicast commits, warps 2..5 epilogue of the CTA's own 128 rows.
// ---------------------------------------------------------------------------------------------------------
template <int AK>
__global__ void __launch_bounds__(kThreads, 1) contract2_kernel(const __grid_constant__ GemmParams p) {
  static_assert(AK == A_MN128 || AK == A_K128, "2-CTA kernel serves the N^3 contractions only");
  constexpr int BK = 64;
  using C = Cfg<AK, BK>;
  constexpr int A_STAGE = C::A_STAGE;          // 16 KB: this CTA's 128 rows
  constexpr int B_STAGE = 4 * BK * 64;         // 16 KB: this CTA's 4 chunks (128 of the 256 columns)
  const int S = p.stages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NRES = p.b_res_reps * p.kb_total;       // resident B tiles (0: B goes through the ring)
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)S * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + (size_t)(NRES ? NRES : S) * B_STAGE);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 2;
  uint64_t* bres_full = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_full + 1);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();     // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map);
    tma_prefetch_desc(&p.b_map);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);       // leader's is the live one: its producer's arrive.expect_tx (both CTAs' bytes)
      mbar_init(&empty[s], 1);      // multicast commit of the leader's MMA warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);      // multicast commit
      mbar_init(&tempty[a], 8);     // leader's: 4 epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, kTmemCols);
    tmem_relinquish_2cta();
  }
  if (warp == 2) sbias[lane] = p.ep.bias ? p.ep.bias[lane] : 0.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();               // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.MT * p.NT * p.Z;     // MT counts 256-row tiles here

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        int mt, nt, z;
        decode_tile(p, t, mt, nt, z);
        const int m0 = mt * 256 + (int)rank * 128;
        const int zA0 = ((z / p.am.z_div) % p.am.z_mod) * p.am.z_mul;
        const int zB0 = ((z / p.bm.z_div) % p.bm.z_mod) * p.bm.z_mul;
        int kk = 0, sa_lo = 0, sa_hi = 0, sb_lo = 0, sb_hi = 0;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], (uint32_t)(2 * (A_STAGE + B_STAGE)));
          const int zA = zA0 + sa_lo * p.am.seg_mul + sa_hi * p.am.seg_hi_mul;
          const int zB = zB0 + sb_lo * p.bm.seg_mul + sb_hi * p.bm.seg_hi_mul;
          const int kA = kk * BK + sa_lo * p.am.k_seg;
          const int kB = kk * BK + sb_lo * p.bm.k_seg;
          uint8_t* a_dst = sA + (size_t)stage * A_STAGE;
          uint8_t* b_dst = sB + (size_t)stage * B_STAGE;
          if (AK == A_MN128) {
            tma_load_4d_2cta(a_dst, &p.a_map, &full[stage], m0, kA, zA, 0);
            tma_load_4d_2cta(a_dst + BK * 128, &p.a_map, &full[stage], m0 + 64, kA, zA, 0);
          } else {
            tma_load_4d_2cta(a_dst, &p.a_map, &full[stage], kA, m0, zA, 0);
          }
          const int r0 = nt * 8 + (int)rank * 4;      // this CTA's chunks of the 8-chunk tile
          if (!p.b_flat) {              // dims (ch, k, r, z), box (32, 64, 4)
            tma_load_4d_2cta(b_dst, &p.b_map, &full[stage], 0, kB, r0, zB);
          } else if (p.b_flat == 2) {   // flat [k][cols] tensor, two SWIZZLE_128B atoms of [64 k][64 cols = 128 B] (2 chunks each)
            for (int j = 0; j < 2; ++j)
              tma_load_4d_2cta(b_dst + (size_t)j * BK * 128, &p.b_map, &full[stage], (r0 + 2 * j) * 32, kB, zB, 0);
          } else {
            for (int j = 0; j < 4; ++j)
              tma_load_4d_2cta(b_dst + (size_t)j * BK * 64, &p.b_map, &full[stage], (r0 + j) * 32, kB, zB, 0);
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
          if (++kk == p.kb_per_seg) {
            kk = 0;
            if (++sa_lo == p.am.seg_mod) { sa_lo = 0; ++sa_hi; }
            if (++sb_lo == p.bm.seg_mod) { sb_lo = 0; ++sb_hi; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_f16(256, 256, C::A_MN ? 1 : 0, 1);
      const uint64_t a_hi = umma_desc_hi(C::A_SBO, C::A_LAYOUT);
      // B tile: four SWIZZLE_64B chunks of [64 k][32 cols], or (b_flat == 2) two SWIZZLE_128B atoms of [64 k][64 cols]
      const bool b128 = p.b_flat == 2;
      const uint64_t b_hi = b128 ? umma_desc_hi(1024, 2u) : umma_desc_hi(C::B_SBO, 4u);
      const uint32_t b_kstep = b128 ? 16u * 128u : C::B_KSTEP, b_lbo = b128 ? (uint32_t)BK * 128u : C::B_LBO;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_addr = smem_u32(sA + (size_t)stage * A_STAGE);
            const uint32_t b_addr = smem_u32(sB + (size_t)stage * B_STAGE);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = umma_desc(a_hi, a_addr + k * C::A_KSTEP, C::A_LBO);
              const uint64_t bd = umma_desc(b_hi, b_addr + k * b_kstep, b_lbo);
              umma_f16_2cta(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2cta(&empty[stage]);
            if (kb == p.kb_total - 1) umma_commit_2cta(&tfull[acc]);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..5 of both CTAs) ------------------------------
    const int quarter = warp & 3;
    const float alpha = p.ep.alpha_dev ? p.ep.alpha * __ldg(p.ep.alpha_dev) : p.ep.alpha;
    float amax = 0.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      int mt, nt, z;
      decode_tile(p, t, mt, nt, z);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int i = mt * 256 + (int)rank * 128 + quarter * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * kAccCols;
      long long base = (long long)z * p.ep.sZ + (long long)i * p.ep.sI;
      void* out_ptr = p.ep.out;
      if (p.ep.peer_g > 0 && i < p.ep.m_valid) {      // push this row of the partial result into its owner's staging slot
        const int owner = i / p.ep.peer_rows;
        out_ptr = p.ep.peer_out[owner];
        base = p.ep.peer_slot + (long long)z * p.ep.peer_sZ + (long long)(i - owner * p.ep.peer_rows) * p.ep.sI;
      }
      float dl[8];
      long long cbase = 0;
      bool any_corr = false;
      if (p.ep.corr_src != nullptr) {
        const int zA = ((z / p.am.z_div) % p.am.z_mod) * p.am.z_mul;
        const int zB = ((z / p.bm.z_div) % p.bm.z_mod) * p.bm.z_mul;
        cbase = (long long)zB * p.ep.cZ + (long long)i * p.ep.cI;
#pragma unroll
        for (int sgi = 0; sgi < 8; ++sgi) {
          dl[sgi] = (sgi < p.ep.corr_nseg && i < p.ep.m_valid)
                        ? __ldg(p.ep.corr_delta + ((long long)zA * p.ep.corr_nseg + sgi) * p.ep.m_valid + i) : 0.f;
          any_corr |= (dl[sgi] != 0.f);
        }
      }
      for (int j = 0; j < 8; ++j) {
        uint32_t regs[32];
        tmem_ld_32x32(t_row + (uint32_t)j * 32, regs);
        tmem_ld_wait();
        const int r = nt * 8 + j;
        if (i < p.ep.m_valid && r < p.ep.r_valid) {
          if (any_corr) {
#pragma unroll
            for (int sgi = 0; sgi < 8; ++sgi) {
              if (sgi < p.ep.corr_nseg && dl[sgi] != 0.f) {
                const uint4* src = reinterpret_cast<const uint4*>(p.ep.corr_src + cbase + (long long)sgi * p.ep.cSeg + (long long)r * p.ep.cR);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 v = __ldg(src + q);
                  const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(h2[e]);
                    regs[8 * q + 2 * e] = __float_as_uint(fmaf(dl[sgi], f.x, __uint_as_float(regs[8 * q + 2 * e])));
                    regs[8 * q + 2 * e + 1] = __float_as_uint(fmaf(dl[sgi], f.y, __uint_as_float(regs[8 * q + 2 * e + 1])));
                  }
                }
              }
            }
          }
          store_chunk(p.ep, out_ptr, alpha, sbias, base + (long long)r * p.ep.sR, regs, amax);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[acc], 0);    // tell the leader's MMA warp this accumulator is drained
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (p.ep.absmax_out) {
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(p.ep.absmax_out), __float_as_uint(amax));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();               // nobody leaves (or frees TMEM) while the peer may still touch this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}
#endif  // __CUDACC__

// Launch one contraction (implemented in tc_engine.cu).  ak/bk select the instantiation.
int launch_contract(int ak, int bk, GemmParams& p, cudaStream_t stream);
// 2-CTA (256 x 256 pair tile) launch for A_MN128 / A_K128 with R = 8; p.MT must count 256-row tiles and the B tensor
// map must have box_r = 4 (chunk mode).
int launch_contract_2cta(int ak, GemmParams& p, cudaStream_t stream);
bool use_2cta(int n_rows);

}  // namespace tc
}  // namespace mpgcn
