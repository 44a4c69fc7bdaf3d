// BDGCN layer on the tcgen05 contraction engine (precision 1: fp16 operands, fp32 accumulate).
//
// Same factored algebra as bdgcn_simt.cu (reference: /root/reference/MPGCN.py:24-50; factoring:
// SURVEY.md section 7.1); each contraction is one launch of tc::contract_kernel with operands
// described by TMA tensor maps over fp16 copies living in the caller's workspace:
//
//   forward   X16 [B][n][c][l]      G16 [zg][K][N][Np]  (Np = N rounded up to 8, rows padded)
//             Z16 [B][d][n][e][l]   (saved for backward)   U16 [B][o][n][e][h]
//     FWD_A   Z = X x2 G_d          A_MN128 (G_d [c][e])     B = X16  (ch, c, n, b)
//     FWD_MIX U = sum_d Z_d W[o,d]  A_K64   (Z plane)        B = W16  (h, (d,l), o)
//     FWD_B   out = act(sum_o G_o^T x1 U_o + b)   A_MN128 (G flat [(o,n)][m])  B = U16 flat
//   backward  dP16 [B][m][e][h]   V16 [B][o][n][e][h]   Y16 [B][d][n][e][l]   Wq16 [d][o][h][l]
//     BWD_V   V = G_o x1 dPre       A_K128  (G_o [n][m])     B = dP16 flat
//     BWD_DW  dW = Z^T V            A_MN64  (Z)              B = V16  (h, row, o, b)   split-K
//     BWD_MIX Y = sum_o V_o W[o,d]^T  A_K64 (V plane)        B = Wq16 (l, (o,h), d)
//     BWD_DX  dX = sum_d Y_d x2 G_d^T A_K128 (G_d [c][e])    B = Y16  (l, e, n, (b,d))
#include "kernels.h"
#include "tc_engine.cuh"

#include <stdlib.h>
#include <string.h>

namespace mpgcn {

using tc::GemmParams;

namespace {
inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
inline int pad8(int n) { return (n + 7) / 8 * 8; }
const int kBig = 1 << 30;

tc::OperandMap omap(int z_div, int z_mod, int z_mul, int seg_mul, int k_seg, int seg_mod = 1 << 30, int seg_hi_mul = 0) {
  tc::OperandMap m;
  m.z_div = z_div; m.z_mod = z_mod; m.z_mul = z_mul; m.seg_mul = seg_mul; m.k_seg = k_seg;
  m.seg_mod = seg_mod; m.seg_hi_mul = seg_hi_mul;
  return m;
}

void init_params(GemmParams& p) {
  memset(&p, 0, sizeof(p));
  p.am = omap(1, 1, 0, 0, 0);
  p.bm = omap(1, 1, 0, 0, 0);
  p.kb_per_slice = 1;
  p.ep.alpha = 1.f;
}

// support stack G16 [zg*K][N rows][Np]: as A_MN128 (dims m, k-rows, z) or A_K128 (dims k, m-rows, z)
int map_support_mn(CUtensorMap* m, const __half* g16, int N, int Np, long long rows_per_z, long long nz) {
  const uint64_t dims[4] = {(uint64_t)N, (uint64_t)rows_per_z, (uint64_t)nz, 1};
  const uint64_t str[3] = {(uint64_t)Np * 2, (uint64_t)rows_per_z * Np * 2, (uint64_t)rows_per_z * Np * 2 * (uint64_t)nz};
  const uint32_t box[4] = {64, 64, 1, 1};
  return make_tmap_f16(m, g16, 4, dims, str, box, TMAP_SW128);
}
int map_support_k(CUtensorMap* m, const __half* g16, int N, int Np, long long nz) {
  const uint64_t dims[4] = {(uint64_t)N, (uint64_t)N, (uint64_t)nz, 1};
  const uint64_t str[3] = {(uint64_t)Np * 2, (uint64_t)N * Np * 2, (uint64_t)N * Np * 2 * (uint64_t)nz};
  const uint32_t box[4] = {64, 128, 1, 1};
  return make_tmap_f16(m, g16, 4, dims, str, box, TMAP_SW128);
}
// channel-chunk tensor T[z][r][k][32]: dims (ch, k, r, z)
int map_chunks(CUtensorMap* m, const __half* t, long long k_rows, long long k_stride_el, long long r_count, long long r_stride_el,
               long long z_count, long long z_stride_el, int box_k, int box_r) {
  const uint64_t dims[4] = {32, (uint64_t)k_rows, (uint64_t)r_count, (uint64_t)z_count};
  const uint64_t str[3] = {(uint64_t)k_stride_el * 2, (uint64_t)r_stride_el * 2, (uint64_t)z_stride_el * 2};
  const uint32_t box[4] = {32, (uint32_t)box_k, (uint32_t)box_r, 1};
  return make_tmap_f16(m, t, 4, dims, str, box, TMAP_SW64);
}
// flat tensor T[z][k][cols]: dims (col, k, z, 1), box (32, box_k)
int map_flat(CUtensorMap* m, const __half* t, long long cols, long long k_rows, long long z_count, int box_k, int box_cols = 32) {
  const uint64_t dims[4] = {(uint64_t)cols, (uint64_t)k_rows, (uint64_t)z_count, 1};
  const uint64_t str[3] = {(uint64_t)cols * 2, (uint64_t)cols * k_rows * 2, (uint64_t)cols * k_rows * 2 * (uint64_t)z_count};
  const uint32_t box[4] = {(uint32_t)box_cols, (uint32_t)box_k, 1, 1};
  return make_tmap_f16(m, t, 4, dims, str, box, box_cols == 64 ? TMAP_SW128 : TMAP_SW64);
}
// K-major plane tensor T[plane][rows][32]: dims (k=32, row, plane, 1), box (32, 128)
int map_planes(CUtensorMap* m, const __half* t, long long rows, long long planes, int box_planes = 1) {
  const uint64_t dims[4] = {32, (uint64_t)rows, (uint64_t)planes, 1};
  const uint64_t str[3] = {64, (uint64_t)rows * 64, (uint64_t)rows * 64 * (uint64_t)planes};
  const uint32_t box[4] = {32, 128, (uint32_t)box_planes, 1};
  return make_tmap_f16(m, t, 4, dims, str, box, TMAP_SW64);
}
}  // namespace

// How the flat B operands (U16 of FWD_B, dP16 of BWD_V: k rows 64 KB apart, the chunks of a row contiguous) are fetched.
// MPGCN_B200_FLAT_B = 2 (default, pair kernel): 64-column SWIZZLE_128B boxes, i.e. 128-byte requests, two chunks per box --
// measured FWD_B 0.76 -> 0.70 ms, BWD_V 0.83 -> 0.74 ms against 0 = one permuted SWIZZLE_64B box (64-byte requests that
// revisit every 128-byte line in a second pass); 1 = one 32-column box per chunk.
static int flat_b_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MPGCN_B200_FLAT_B");
    v = e ? atoi(e) : 2;
  }
  return v;
}
static bool flat_b_boxes() { return flat_b_mode() == 1; }

// launch an N^3 contraction on the 2-CTA kernel when the M extent allows it (p prepared for the 1-CTA kernel)
static int launch_big(int ak, GemmParams& p, int m_rows, cudaStream_t st) {
  if (tc::use_2cta(m_rows)) {
    p.MT = ceil_div(m_rows, 256);
    return tc::launch_contract_2cta(ak, p, st);
  }
  return tc::launch_contract(ak, 64, p, st);
}

bool tc_supported(const BdgcnShape& s) {
  return s.C == 32 && s.H == 32 && s.Ko >= 1 && s.Ko <= 8 && s.Kd >= 1 && s.Kd <= 8 && s.N >= 1 && s.B >= 1 && s.R >= 1 && s.row0 >= 0 &&
         s.row0 + s.R <= s.N;
}

// cells of an activation slab: R origin rows x N destinations
static size_t rn(const BdgcnShape& s) { return (size_t)s.R * s.N; }
static size_t g16_elems(const BdgcnShape& s, int K) { return (size_t)(s.dynamic ? s.B : 1) * K * s.N * pad8(s.N); }
static size_t g_planes(const BdgcnShape& s, int K) { return (size_t)(s.dynamic ? s.B : 1) * K; }

size_t tc_saved_bytes(const BdgcnShape& s) { return (size_t)s.B * s.Kd * rn(s) * s.C * sizeof(__half); }

// workspace layouts (byte offsets); also served to tests by mpgcn_debug_tc_workspace_offset()
struct FwdLayout { size_t x16, gd16, go16, w16, u16, z16, dd, dgo, dgo_masked, total; };
struct BwdLayout { size_t dp16, gd16, go16, v16, y16, wq16, partials, scale, total; };
static size_t take(size_t& off, size_t bytes) {
  off = align_up(off, 1024);
  const size_t r = off;
  off += bytes;
  return r;
}
static FwdLayout fwd_layout(const BdgcnShape& s) {
  FwdLayout L;
  size_t off = 0;
  L.x16 = take(off, (size_t)s.B * rn(s) * 32 * 2);
  L.gd16 = take(off, g16_elems(s, s.Kd) * 2);
  L.go16 = take(off, g16_elems(s, s.Ko) * 2);
  L.w16 = take(off, (size_t)2 * s.Ko * s.Kd * 32 * 32 * 2);      // [hi | lo]
  L.u16 = take(off, (size_t)s.B * s.Ko * rn(s) * 32 * 2);
  L.dd = take(off, g_planes(s, s.Kd) * s.N * 4);   // support-diagonal fp16 remainders (destination / origin)
  L.dgo = take(off, g_planes(s, s.Ko) * s.N * 4);
  L.dgo_masked = take(off, g_planes(s, s.Ko) * s.N * 4);   // origin remainders restricted to the row slab
  L.z16 = take(off, tc_saved_bytes(s));            // used only when the caller passes no `saved` buffer
  L.total = align_up(off, 1024);
  return L;
}
size_t tc_fwd_ws_bytes(const BdgcnShape& s) { return fwd_layout(s).total; }
static int dw_slices(const BdgcnShape& s, int* kb_per_slice, int* kb_total) {
  const int kbps = ceil_div((long long)rn(s), 64);
  const int total = s.B * kbps;
  const int MT = ceil_div(s.Kd, 4);
  int want = device_sm_count() / MT;
  if (want < 1) want = 1;
  int per = ceil_div(total, want);
  if (per < 1) per = 1;
  *kb_per_slice = per;
  *kb_total = total;
  return ceil_div(total, per);
}
static BwdLayout bwd_layout(const BdgcnShape& s) {
  BwdLayout L;
  size_t off = 0;
  L.dp16 = take(off, (size_t)s.B * s.N * s.N * 32 * 2);        // dPre: every origin row m, always
  L.gd16 = take(off, g16_elems(s, s.Kd) * 2);
  L.go16 = take(off, g16_elems(s, s.Ko) * 2);
  L.v16 = take(off, (size_t)s.B * s.Ko * rn(s) * 32 * 2);
  L.y16 = take(off, (size_t)s.B * s.Kd * rn(s) * 32 * 2);
  L.wq16 = take(off, (size_t)s.Ko * s.Kd * 32 * 32 * 2);
  int per = 1, total = 1;
  const int slices = dw_slices(s, &per, &total);
  L.partials = take(off, (size_t)slices * ceil_div(s.Kd, 4) * 128 * s.Ko * 32 * 4);
  L.scale = take(off, 64);
  L.total = align_up(off, 1024);
  return L;
}
size_t tc_bwd_ws_bytes(const BdgcnShape& s) { return bwd_layout(s).total; }

// which: 0 x16, 1 gd16, 2 go16, 3 w16, 4 u16 (forward); 10 dp16, 11 gd16, 12 go16, 13 v16, 14 y16, 15 wq16, 16 partials,
// 17 number of dW slices (not an offset)
long long tc_debug_offset(const BdgcnShape& s, int which) {
  const FwdLayout F = fwd_layout(s);
  const BwdLayout Bw = bwd_layout(s);
  int per = 1, total = 1;
  switch (which) {
    case 0: return (long long)F.x16;
    case 1: return (long long)F.gd16;
    case 2: return (long long)F.go16;
    case 3: return (long long)F.w16;
    case 4: return (long long)F.u16;
    case 5: return (long long)F.dd;
    case 6: return (long long)F.dgo;
    case 10: return (long long)Bw.dp16;
    case 11: return (long long)Bw.gd16;
    case 12: return (long long)Bw.go16;
    case 13: return (long long)Bw.v16;
    case 14: return (long long)Bw.y16;
    case 15: return (long long)Bw.wq16;
    case 16: return (long long)Bw.partials;
    case 17: return dw_slices(s, &per, &total);
    case 18: return (long long)Bw.scale;
    default: return -1;
  }
}

// ---------------------------------------------------------------------------------------
// individual contractions.  Activations are [B][*][R rows n][N][32] slabs (R = N for a whole layer); supports G_d has Kd
// planes per sample, G_o has Ko.
// ---------------------------------------------------------------------------------------
// FWD_A:  Z16[b][d][n][e][l] = sum_c G_d[c][e] X16[b][n][c][l]
static int run_fwd_a(const BdgcnShape& s, const __half* gd16, const __half* x16, __half* z16, const float* delta_d, cudaStream_t st) {
  const int N = s.N, R = s.R, K = s.Kd, Np = pad8(N);
  GemmParams p;
  init_params(p);
  if (int e = map_support_mn(&p.a_map, gd16, N, Np, N, (long long)g_planes(s, K))) return e;
  if (int e = map_chunks(&p.b_map, x16, N, 32, R, (long long)N * 32, s.B, (long long)R * N * 32, 64, tc::use_2cta(N) ? 4 : 8)) return e;
  p.am = omap(1, s.dynamic ? kBig : K, 1, 0, 0);        // z = b*K + d -> support index
  p.bm = omap(K, kBig, 1, 0, 0);                        // -> b
  p.MT = ceil_div(N, 128); p.NT = ceil_div(R, 8); p.Z = s.B * K; p.R = 8;
  p.z_inner = K;                                        // the K supports of one (sample, row block) run back to back: X16 block from L2
  p.kb_total = p.kb_per_seg = ceil_div(N, 64);
  p.ep.out = z16; p.ep.out_f16 = 1;
  p.ep.sZ = (long long)R * N * 32; p.ep.sI = 32; p.ep.sR = (long long)N * 32;
  p.ep.m_valid = N; p.ep.r_valid = R;
  // Z[b,d,n,e,:] += (G_d[e,e] - fp16(G_d[e,e])) * X16[b,n,e,:]
  p.ep.corr_src = x16; p.ep.corr_delta = delta_d; p.ep.corr_nseg = 1;
  p.ep.cZ = (long long)R * N * 32; p.ep.cI = 32; p.ep.cR = (long long)N * 32; p.ep.cSeg = 0;
  prof_set_next(PROF_FWD_A, 2.0 * s.B * K * (double)R * N * N * 32);
  return launch_big(tc::A_MN128, p, N, st);
}

// MIX: D16[b][r][row][32] = sum_{seg < Kin} A16[b][seg][row][32] * Wm16[r][(seg,32)][32], r < Kout   (both channel mixes)
static int run_mix(const BdgcnShape& s, const __half* a16, const __half* w16, int w_halves, __half* d16, int tag, int Kin, int Kout,
                   cudaStream_t st) {
  const long long NN = (long long)rn(s);
  GemmParams p;
  init_params(p);
  static int mode = -1;     // A/B knob: 0 default, 1 = W streams with A (one plane per k-block), 2 = W resident, one plane per k-block
  if (mode < 0) { const char* e = getenv("MPGCN_B200_MIX_MODE"); mode = e ? atoi(e) : 0; }
  const size_t w_bytes = (size_t)Kout * w_halves * Kin * 32 * 64;      // Kin * w_halves tiles of Kout chunks x [32 k][64 B]
  int bk = 32;
  if (mode == 0 && (Kin == 2 || Kin == 3)) {
    // One k-block per tile: a single TMA box brings the Kin planes of a 128-cell tile (the single-thread producer / MMA loops
    // cost ~0.3 us per k-block, which bounded the per-plane version at a third of the HBM rate), W resident in shared memory
    bk = 32 * Kin;
    if (int e = map_planes(&p.a_map, a16, NN, (long long)s.B * Kin, Kin)) return e;
    if (int e = map_chunks(&p.b_map, w16, (long long)Kin * 32, 32, Kout, (long long)Kin * 32 * 32, w_halves, (long long)Kout * Kin * 32 * 32, bk, Kout)) return e;
    p.am = omap(1, kBig, Kin, 0, 0);               // z = b -> first plane b*Kin
    p.bm = omap(1, 1, 0, 1, 0);                    // resident tile index = weight half
    p.kb_total = 1; p.kb_per_seg = 1; p.b_res_reps = w_halves;
  } else {
    if (int e = map_planes(&p.a_map, a16, NN, (long long)s.B * Kin)) return e;
    if (int e = map_chunks(&p.b_map, w16, (long long)Kin * 32, 32, Kout, (long long)Kin * 32 * 32, w_halves, (long long)Kout * Kin * 32 * 32, 32, Kout)) return e;
    // segment s: plane = b*Kin + (s % Kin); weight rows (s % Kin)*32 of half s / Kin  (half 0 = fp16(W), half 1 = fp16(W - half 0))
    p.am = omap(1, kBig, Kin, 1, 0, Kin, 0);
    p.bm = omap(1, 1, 0, 0, 32, Kin, 1);
    if (mode == 1 || w_bytes > 160 * 1024) { p.kb_total = Kin * w_halves; p.kb_per_seg = 1; }      // large K: W streams with A
    else { p.kb_total = Kin; p.kb_per_seg = 1; p.b_res_reps = w_halves; }                           // W resident, A plane by plane
  }
  p.MT = ceil_div(NN, 128); p.NT = 1; p.Z = s.B; p.R = Kout;
  p.ep.out = d16; p.ep.out_f16 = 1;
  p.ep.sZ = (long long)Kout * NN * 32; p.ep.sI = 32; p.ep.sR = NN * 32;
  p.ep.m_valid = (int)NN; p.ep.r_valid = Kout;
  prof_set_next(tag, 2.0 * s.B * (double)Kin * Kout * NN * 32 * 32);   // algorithmic flops (the fp16 hi/lo weight split doubles the executed MMAs)
  return tc::launch_contract(tc::A_K64, bk, p, st);
}

// FWD_B: out[b][m][e][h] = act( sum_{(o,n)} G_o[row0 + n][m] U16[b][o][n][e][h] + bias[h] )      n < R, every m < N
static int run_fwd_b(const BdgcnShape& s, const __half* go16, const __half* u16, const float* bias, float* out, __half* out16,
                     const float* delta_o, cudaStream_t st) {
  const int N = s.N, R = s.R, K = s.Ko, Np = pad8(N);
  const bool slab = !(R == N && s.row0 == 0);
  GemmParams p;
  init_params(p);
  if (!slab) {
    // whole layer: the contraction index (o, n) is a plain row index of the flat [K*N][N] support stack and of U16
    if (int e = map_support_mn(&p.a_map, go16, N, Np, (long long)K * N, s.dynamic ? s.B : 1)) return e;
    // U16 [b][(o,n)][e][h] read as (h, k = (o,n) rows, r = e, b): dims listed with non-monotonic strides (the r stride,
    // 64 B, is smaller than the k stride) so that ONE box (32 ch, 64 k, 4|8 r) lands in the canonical [r][k][64 B] layout
    if (flat_b_mode() == 2 && tc::use_2cta(N)) {
      if (int e = map_flat(&p.b_map, u16, (long long)N * 32, (long long)K * N, s.B, 64, 64)) return e;
      p.b_flat = 2;
    } else if (flat_b_boxes()) {
      if (int e = map_flat(&p.b_map, u16, (long long)N * 32, (long long)K * N, s.B, 64)) return e;
      p.b_flat = 1;
    } else {
      if (int e = map_chunks(&p.b_map, u16, (long long)K * N, (long long)N * 32, N, 32, s.B, (long long)K * N * N * 32, 64, tc::use_2cta(N) ? 4 : 8)) return e;
    }
    p.am = omap(1, s.dynamic ? kBig : 1, 1, 0, 0);
    p.bm = omap(1, kBig, 1, 0, 0);
    p.kb_total = p.kb_per_seg = ceil_div((long long)K * N, 64);
  } else {
    // origin-row slab: one k-segment per support o.  A = rows [row0, row0 + R) of G_o (k coordinate o*N + kk inside the flat
    // stack, base pointer moved to row0); B = U16 [b][o][n < R]: TMA zero-fills rows >= R, which also cancels the rows of the
    // NEXT slab that the last k-block of a segment reads from G.
    const long long gplanes = s.dynamic ? s.B : 1;
    {
      const uint64_t dims[4] = {(uint64_t)N, (uint64_t)((long long)K * N - s.row0), (uint64_t)gplanes, 1};
      const uint64_t str[3] = {(uint64_t)Np * 2, (uint64_t)K * N * Np * 2, (uint64_t)K * N * Np * 2 * (uint64_t)gplanes};
      const uint32_t box[4] = {64, 64, 1, 1};
      if (int e = make_tmap_f16(&p.a_map, go16 + (size_t)s.row0 * Np, 4, dims, str, box, TMAP_SW128)) return e;
    }
    if (flat_b_mode() == 2 && tc::use_2cta(N)) {
      if (int e = map_flat(&p.b_map, u16, (long long)N * 32, R, (long long)s.B * K, 64, 64)) return e;
      p.b_flat = 2;
    } else if (flat_b_boxes()) {
      if (int e = map_flat(&p.b_map, u16, (long long)N * 32, R, (long long)s.B * K, 64)) return e;
      p.b_flat = 1;
    } else {
      if (int e = map_chunks(&p.b_map, u16, R, (long long)N * 32, N, 32, (long long)s.B * K, (long long)R * N * 32, 64, tc::use_2cta(N) ? 4 : 8)) return e;
    }
    p.am = omap(1, s.dynamic ? kBig : 1, 1, 0, N);          // k coordinate += o * N
    p.bm = omap(1, kBig, K, 1, 0);                          // plane = b*K + o
    p.kb_per_seg = ceil_div(R, 64);
    p.kb_total = K * p.kb_per_seg;
  }
  p.MT = ceil_div(N, 128); p.NT = ceil_div(N, 8); p.Z = s.B; p.R = 8;
  p.ep.out = out; p.ep.out_f16 = 0; p.ep.out16 = out16;
  p.ep.sZ = (long long)N * N * 32; p.ep.sI = (long long)N * 32; p.ep.sR = 32;
  p.ep.m_valid = N; p.ep.r_valid = N;
  p.ep.bias = s.partial ? nullptr : bias; p.ep.relu = s.partial ? 0 : s.act;
  if (s.partial && s.peer_g > 0) {        // push every row of the partial into its owner's staging slot for this rank (NVLink P2P stores)
    p.ep.peer_g = s.peer_g; p.ep.peer_rows = N / s.peer_g;
    p.ep.peer_sZ = (long long)p.ep.peer_rows * N * 32;
    p.ep.peer_slot = (long long)s.peer_rank * s.B * p.ep.peer_sZ;
    for (int j = 0; j < s.peer_g; ++j) p.ep.peer_out[j] = s.peer_out[j];
    p.ep.out16 = nullptr;
  }
  // pre[b,m,e,:] += sum_o (G_o[m,m] - fp16(G_o[m,m])) * U16[b,o,m,e,:]   (delta_o is zero outside the slab: the row n = m of U
  // exists only for row0 <= m < row0 + R; corr_src is moved so that row index m addresses slab row m - row0)
  p.ep.corr_src = u16 - (long long)s.row0 * N * 32; p.ep.corr_delta = delta_o; p.ep.corr_nseg = K;
  // (the epilogue forms the sample offset as zB * cZ with zB = b in the flat mode and b*K in the slab mode)
  p.ep.cZ = slab ? (long long)R * N * 32 : (long long)K * R * N * 32;
  p.ep.cSeg = (long long)R * N * 32; p.ep.cI = (long long)N * 32; p.ep.cR = 32;
  prof_set_next(PROF_FWD_B, 2.0 * s.B * K * (double)R * N * N * 32);
  return launch_big(tc::A_MN128, p, N, st);
}

// BWD_V: V16[b][o][n][e][h] = sum_m G_o[row0 + n][m] dP16[b][m][e][h]       n < R
static int run_bwd_v(const BdgcnShape& s, const __half* go16, const __half* dp16, __half* v16, cudaStream_t st) {
  const int N = s.N, R = s.R, K = s.Ko, Np = pad8(N);
  GemmParams p;
  init_params(p);
  {   // rows [row0, row0 + R) of every support plane, K-major
    const long long nz = (long long)g_planes(s, K);
    const uint64_t dims[4] = {(uint64_t)N, (uint64_t)R, (uint64_t)nz, 1};
    const uint64_t str[3] = {(uint64_t)Np * 2, (uint64_t)N * Np * 2, (uint64_t)N * Np * 2 * (uint64_t)nz};
    const uint32_t box[4] = {64, 128, 1, 1};
    if (int e = make_tmap_f16(&p.a_map, go16 + (size_t)s.row0 * Np, 4, dims, str, box, TMAP_SW128)) return e;
  }
  if (flat_b_mode() == 2 && tc::use_2cta(R)) {
    if (int e = map_flat(&p.b_map, dp16, (long long)N * 32, N, s.B, 64, 64)) return e;
    p.b_flat = 2;
  } else if (flat_b_boxes()) {
    if (int e = map_flat(&p.b_map, dp16, (long long)N * 32, N, s.B, 64)) return e;
    p.b_flat = 1;
  } else {   // dP16 [b][m][e][h] read as (h, k = m, r = e, b), see run_fwd_b
    if (int e = map_chunks(&p.b_map, dp16, N, (long long)N * 32, N, 32, s.B, (long long)N * N * 32, 64, tc::use_2cta(R) ? 4 : 8)) return e;
  }
  p.am = omap(1, s.dynamic ? kBig : K, 1, 0, 0);        // z = b*K + o
  p.bm = omap(K, kBig, 1, 0, 0);
  p.MT = ceil_div(R, 128); p.NT = ceil_div(N, 8); p.Z = s.B * K; p.R = 8;
  p.z_inner = K;                                        // dP16 block of one (sample, e block) serves the K supports back to back
  p.kb_total = p.kb_per_seg = ceil_div(N, 64);
  p.ep.out = v16; p.ep.out_f16 = 1;
  p.ep.sZ = (long long)R * N * 32; p.ep.sI = (long long)N * 32; p.ep.sR = 32;
  p.ep.m_valid = R; p.ep.r_valid = N;
  prof_set_next(PROF_BWD_V, 2.0 * s.B * K * (double)R * N * N * 32);
  return launch_big(tc::A_K128, p, R, st);
}

// BWD_DW: P[slice][mt][(d%4)*32+l][o][h] = sum over the slice's (b,row) range of Z16[b][d][row][l] V16[b][o][row][h]
static int run_bwd_dw(const BdgcnShape& s, const __half* z16, const __half* v16, float* partials, int* slices_out, int* mt_out,
                      cudaStream_t st) {
  const int Ko = s.Ko, Kd = s.Kd;
  const long long NN = (long long)rn(s);
  GemmParams p;
  init_params(p);
  if (int e = map_chunks(&p.a_map, z16, NN, 32, Kd, NN * 32, s.B, (long long)Kd * NN * 32, 64, 4)) return e;
  if (int e = map_chunks(&p.b_map, v16, NN, 32, Ko, NN * 32, s.B, (long long)Ko * NN * 32, 64, Ko)) return e;
  p.am = omap(1, 1, 0, 1, 0);           // z (slice) ignored; batch element = segment
  p.bm = omap(1, 1, 0, 1, 0);
  int per = 1, total = 1;
  const int slices = dw_slices(s, &per, &total);
  p.MT = ceil_div(Kd, 4); p.NT = 1; p.Z = slices; p.R = Ko;
  p.kb_total = total; p.kb_per_seg = ceil_div(NN, 64);
  p.split_k = 1; p.kb_per_slice = per;
  p.ep.out = partials; p.ep.out_f16 = 0;
  p.ep.sZ = (long long)p.MT * 128 * Ko * 32; p.ep.sI = (long long)Ko * 32; p.ep.sR = 32;
  p.ep.m_valid = p.MT * 128; p.ep.r_valid = Ko;
  *slices_out = slices;
  *mt_out = p.MT;
  prof_set_next(PROF_BWD_DW, 2.0 * s.B * (double)Ko * Kd * NN * 32 * 32);
  return tc::launch_contract(tc::A_MN64, 64, p, st);
}

// BWD_DX: dX[b][n][c][l] = sum_{d,e} G_d[c][e] Y16[b][d][n][e][l]       n < R
static int run_bwd_dx(const BdgcnShape& s, const __half* gd16, const __half* y16, float* dX, const float* inv_scale, float* dx_absmax,
                      cudaStream_t st) {
  const int N = s.N, R = s.R, K = s.Kd, Np = pad8(N);
  GemmParams p;
  init_params(p);
  if (int e = map_support_k(&p.a_map, gd16, N, Np, (long long)g_planes(s, K))) return e;
  if (int e = map_chunks(&p.b_map, y16, N, 32, R, (long long)N * 32, (long long)s.B * K, (long long)R * N * 32, 64, tc::use_2cta(N) ? 4 : 8)) return e;
  p.am = omap(1, s.dynamic ? kBig : 1, s.dynamic ? K : 0, 1, 0);   // support index = (b*K) + d
  p.bm = omap(1, kBig, K, 1, 0);                                    // plane = b*K + d
  p.MT = ceil_div(N, 128); p.NT = ceil_div(R, 8); p.Z = s.B; p.R = 8;
  p.kb_per_seg = ceil_div(N, 64); p.kb_total = K * p.kb_per_seg;
  p.ep.out = dX; p.ep.out_f16 = 0; p.ep.alpha_dev = inv_scale; p.ep.absmax_out = dx_absmax;
  p.ep.sZ = (long long)R * N * 32; p.ep.sI = 32; p.ep.sR = (long long)N * 32;
  p.ep.m_valid = N; p.ep.r_valid = R;
  prof_set_next(PROF_BWD_DX, 2.0 * s.B * K * (double)R * N * N * 32);
  return launch_big(tc::A_K128, p, N, st);
}

// Prepared supports: [planes][N][Np] fp16 (padding zeroed) followed, 256-byte aligned, by the [planes][N] diagonal remainders.
// A caller that uses the same supports for several layers / for forward and backward converts them once.
static size_t prep_g16_bytes(long long planes, int N) { return align_up((size_t)planes * N * pad8(N) * sizeof(__half), 256); }
size_t bdgcn_supports_prepared_bytes(long long planes, int N) { return prep_g16_bytes(planes, N) + align_up((size_t)planes * N * sizeof(float), 256); }
int bdgcn_prepare_supports(const float* G, void* prepared, long long planes, int N, cudaStream_t st) {
  MPGCN_CHECK((reinterpret_cast<uintptr_t>(prepared) & 255) == 0, "prepared-supports buffer must be 256-byte aligned");
  __half* g16 = static_cast<__half*>(prepared);
  float* delta = reinterpret_cast<float*>(static_cast<uint8_t*>(prepared) + prep_g16_bytes(planes, N));
  if (int e = cvt_f32_to_f16_padded(G, g16, (size_t)planes * N, N, pad8(N), st)) return e;
  return support_diag_delta(G, delta, (size_t)planes, N, st);
}

// resolves the fp16 supports (and, for the forward, the diagonal remainders) of one side: prepared by the caller, shared
// with the other side (static graph: Go == Gd), or converted here into the workspace
struct SideG { const __half* g16; const float* delta; };
static int resolve_side(const BdgcnShape& s, int K, const float* G, const void* prepared, __half* ws16, float* ws_delta, bool want_delta,
                        SideG* out, cudaStream_t st) {
  const long long planes = (long long)g_planes(s, K);
  if (prepared) {
    MPGCN_CHECK((reinterpret_cast<uintptr_t>(prepared) & 255) == 0, "prepared-supports buffer must be 256-byte aligned");
    out->g16 = static_cast<const __half*>(prepared);
    out->delta = reinterpret_cast<const float*>(static_cast<const uint8_t*>(prepared) + prep_g16_bytes(planes, s.N));
    return 0;
  }
  if (int e = cvt_f32_to_f16_padded(G, ws16, (size_t)planes * s.N, s.N, pad8(s.N), st)) return e;
  if (want_delta) { if (int e = support_diag_delta(G, ws_delta, (size_t)planes, s.N, st)) return e; }
  out->g16 = ws16;
  out->delta = ws_delta;
  return 0;
}

// ---------------------------------------------------------------------------------------
// layer forward / backward
// ---------------------------------------------------------------------------------------
int bdgcn_forward_tc(const BdgcnShape& s, const float* X, const float* Go, const float* Gd, const float* W, const float* bias,
                     float* out, void* saved, void* ws, size_t ws_bytes, const BdgcnExtras& ex, cudaStream_t st) {
  MPGCN_CHECK(tc_supported(s), "tensor-core path needs C = H = 32 and K <= 8 (got C=%d H=%d Ko=%d Kd=%d)", s.C, s.H, s.Ko, s.Kd);
  const size_t NN = rn(s);
  const FwdLayout L = fwd_layout(s);
  MPGCN_CHECK(ws_bytes >= L.total, "bdgcn_forward: workspace too small (%zu < %zu bytes)", ws_bytes, L.total);
  MPGCN_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
  uint8_t* wb = static_cast<uint8_t*>(ws);
  __half* x16_ws = reinterpret_cast<__half*>(wb + L.x16);
  __half* w16 = reinterpret_cast<__half*>(wb + L.w16);
  __half* u16 = reinterpret_cast<__half*>(wb + L.u16);
  __half* z16 = saved ? static_cast<__half*>(saved) : reinterpret_cast<__half*>(wb + L.z16);
  MPGCN_CHECK((reinterpret_cast<uintptr_t>(z16) & 63) == 0, "`saved` buffer must be 64-byte aligned");
  MPGCN_CHECK(((reinterpret_cast<uintptr_t>(ex.x_f16) | reinterpret_cast<uintptr_t>(ex.out_f16) | reinterpret_cast<uintptr_t>(out)) & 31) == 0,
              "bdgcn_forward: `out` and the fp16 side buffers must be 32-byte aligned (256-bit stores)");
  MPGCN_CHECK((reinterpret_cast<uintptr_t>(X) & 15) == 0, "bdgcn_forward: X must be 16-byte aligned");

  const __half* x16 = static_cast<const __half*>(ex.x_f16);
  if (x16 == nullptr) {
    if (int e = cvt_f32_to_f16(X, x16_ws, (size_t)s.B * NN * 32, st)) return e;
    x16 = x16_ws;
  }
  SideG gd{}, go{};
  if (int e = resolve_side(s, s.Kd, Gd, ex.gd_prepared, reinterpret_cast<__half*>(wb + L.gd16), reinterpret_cast<float*>(wb + L.dd), true, &gd, st)) return e;
  if (Go == Gd && ex.go_prepared == nullptr && s.Ko == s.Kd) go = gd;
  else if (int e = resolve_side(s, s.Ko, Go, ex.go_prepared, reinterpret_cast<__half*>(wb + L.go16), reinterpret_cast<float*>(wb + L.dgo), true, &go, st)) return e;
  const float* delta_o = go.delta;
  if (!(s.R == s.N && s.row0 == 0)) {     // the origin-side remainder applies to the rows n = m of this slab only
    float* masked = reinterpret_cast<float*>(wb + L.dgo_masked);
    if (int e = mask_delta_rows(go.delta, masked, g_planes(s, s.Ko), s.N, s.row0, s.R, st)) return e;
    delta_o = masked;
  }
  const size_t wn = (size_t)s.Ko * s.Kd * 32 * 32;
  if (int e = cvt_f32_to_f16_hilo(W, w16, w16 + wn, wn, st)) return e;
  if (int e = run_fwd_a(s, gd.g16, x16, z16, gd.delta, st)) return e;
  if (int e = run_mix(s, z16, w16, 2, u16, PROF_FWD_MIX, s.Kd, s.Ko, st)) return e;
  if (int e = run_fwd_b(s, go.g16, u16, bias, out, static_cast<__half*>(ex.out_f16), delta_o, st)) return e;
  return 0;
}

int bdgcn_backward_tc(const BdgcnShape& s, const float* d_out, const float* out, const float* Go, const float* Gd, const float* W,
                      const void* saved, float* dX, float* dW, float* db, void* ws, size_t ws_bytes, const BdgcnExtras& ex,
                      cudaStream_t st) {
  MPGCN_CHECK(tc_supported(s), "tensor-core path needs C = H = 32 and K <= 8 (got C=%d H=%d Ko=%d Kd=%d)", s.C, s.H, s.Ko, s.Kd);
  MPGCN_CHECK(saved != nullptr, "bdgcn_backward: forward was run without a `saved` buffer");
  const int act = s.partial ? 0 : s.act;       // a partial call receives dPre: the mask was applied by the caller, after the exchange
  MPGCN_CHECK(out != nullptr || ex.out_f16 != nullptr || !act, "bdgcn_backward: the ReLU mask needs `out` or its fp16 copy");
  MPGCN_CHECK((reinterpret_cast<uintptr_t>(dX) & 31) == 0, "bdgcn_backward: dX must be 32-byte aligned (256-bit stores)");
  MPGCN_CHECK(((reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "bdgcn_backward: d_out / out must be 16-byte aligned");
  const size_t NNfull = (size_t)s.N * s.N;
  const __half* z16 = static_cast<const __half*>(saved);
  const BwdLayout L = bwd_layout(s);
  MPGCN_CHECK(ws_bytes >= L.total, "bdgcn_backward: workspace too small (%zu < %zu bytes)", ws_bytes, L.total);
  MPGCN_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
  uint8_t* wb = static_cast<uint8_t*>(ws);
  const __half* dp16 = reinterpret_cast<__half*>(wb + L.dp16);
  __half* v16 = reinterpret_cast<__half*>(wb + L.v16);
  __half* y16 = reinterpret_cast<__half*>(wb + L.y16);
  __half* wq16 = reinterpret_cast<__half*>(wb + L.wq16);
  float* partials = reinterpret_cast<float*>(wb + L.partials);
  const float* scale2 = reinterpret_cast<float*>(wb + L.scale);   // [S, 1/S]: power-of-two gradient scale (fp16 range)

  if (ex.d_pre_f16 != nullptr) {
    // dPre arrives masked, scaled and cast (mpgcn_relu_backward_scatter_f16 wrote it into every rank's buffer): no pass over it here
    MPGCN_CHECK(s.partial && ex.d_pre_scale2 != nullptr, "bdgcn_backward: a prepared fp16 dPre needs a part call and its scale pair");
    MPGCN_CHECK((reinterpret_cast<uintptr_t>(ex.d_pre_f16) & 15) == 0, "bdgcn_backward: prepared fp16 dPre must be 16-byte aligned");
    dp16 = static_cast<const __half*>(ex.d_pre_f16);
    scale2 = ex.d_pre_scale2;
  } else {
    __half* dp16_ws = reinterpret_cast<__half*>(wb + L.dp16);
    float* scale2_ws = reinterpret_cast<float*>(wb + L.scale);
    if (db) MPGCN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * 32, st));
    if (int e = grad_scale_prepare(d_out, (size_t)s.B * NNfull * 32, scale2_ws, ex.d_out_absmax, st)) return e;
    if (ex.out_f16 != nullptr && act) {
      if (int e = relu_bwd_prep_f16mask(d_out, static_cast<const __half*>(ex.out_f16), act, dp16_ws, db, (size_t)s.B * NNfull * 32, 32, scale2_ws, st)) return e;
    } else {
      if (int e = relu_bwd_prep(d_out, out, act, dp16_ws, nullptr, db, (size_t)s.B * NNfull * 32, 32, scale2_ws, st)) return e;
    }
  }
  SideG gd{}, go{};
  const bool shared = Go == Gd && s.Ko == s.Kd;
  if (dX || shared) {
    if (int e = resolve_side(s, s.Kd, Gd, ex.gd_prepared, reinterpret_cast<__half*>(wb + L.gd16), nullptr, false, &gd, st)) return e;
  }
  if (shared && ex.go_prepared == nullptr) go = gd;
  else if (int e = resolve_side(s, s.Ko, Go, ex.go_prepared, reinterpret_cast<__half*>(wb + L.go16), nullptr, false, &go, st)) return e;
  if (int e = run_bwd_v(s, go.g16, dp16, v16, st)) return e;
  int slices = 0, mt = 0;
  if (int e = run_bwd_dw(s, z16, v16, partials, &slices, &mt, st)) return e;
  if (int e = reduce_dw_partials(partials, dW, slices, mt, s.Ko, s.Kd, scale2 + 1, st)) return e;
  if (dX) {
    if (ex.dx_absmax) MPGCN_CUDA(cudaMemsetAsync(ex.dx_absmax, 0, sizeof(float), st));
    if (int e = permute_w_bwd(W, wq16, nullptr, s.Ko, s.Kd, 32, 32, st)) return e;
    if (int e = run_mix(s, v16, wq16, 1, y16, PROF_BWD_MIX, s.Ko, s.Kd, st)) return e;
    if (int e = run_bwd_dx(s, gd.g16, y16, dX, scale2 + 1, ex.dx_absmax, st)) return e;
  }
  return 0;
}

}  // namespace mpgcn
