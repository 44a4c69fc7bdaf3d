// BDGCN layer, exact fp32 mode (precision 0): the factored evaluation order of
// SURVEY.md section 7.1 expressed as calls of the strided SIMT SGEMM.
//
//   forward  (reference: /root/reference/MPGCN.py:24-50)
//     Z[b,d,n,e,l] = sum_c X[b,n,c,l] G_d[c,e]
//     U[b,o,n,e,h] = sum_{d,l} Z[b,d,n,e,l] W[o,d,l,h]
//     out[b,m,e,h] = act( sum_{o,n} G_o[n,m] U[b,o,n,e,h] + bias[h] )
//   backward (what autograd derives from the same lines; supports never need grad)
//     dPre = dOut * act'(out) ; db = sum dPre
//     V[b,o,n,e,h] = sum_m G_o[n,m] dPre[b,m,e,h]
//     dW[o,d,l,h]  = sum_{b,n,e} Z[b,d,n,e,l] V[b,o,n,e,h]
//     Y[b,d,n,e,l] = sum_{o,h} V[b,o,n,e,h] W[o,d,l,h]
//     dX[b,n,c,l]  = sum_{d,e} Y[b,d,n,e,l] G_d[c,e]
#include "kernels.h"

namespace mpgcn {

namespace {
struct Carver {
  uint8_t* base;
  size_t off, cap;
  Carver(void* p, size_t bytes) : base(static_cast<uint8_t*>(p)), off(0), cap(bytes) {}
  template <class T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};
}  // namespace

// cells of an activation slab: R origin rows x N destinations (R = N for a whole layer)
static size_t rn(const BdgcnShape& s) { return (size_t)s.R * s.N; }

size_t simt_saved_bytes(const BdgcnShape& s) { return (size_t)s.B * s.Kd * rn(s) * s.C * sizeof(float); }
size_t simt_fwd_ws_bytes(const BdgcnShape& s) {
  return 256 + align_up((size_t)s.B * s.Ko * rn(s) * s.H * sizeof(float), 256) + align_up(simt_saved_bytes(s), 256);
}
size_t simt_bwd_ws_bytes(const BdgcnShape& s) {
  return 1024 + align_up((size_t)s.B * s.N * s.N * s.H * 4, 256) + align_up((size_t)s.B * s.Ko * rn(s) * s.H * 4, 256) +
         align_up((size_t)s.B * s.Kd * rn(s) * s.C * 4, 256) + align_up((size_t)s.Ko * s.Kd * s.C * s.H * 4, 256);
}

static void zero3(long long (&a)[3]) { a[0] = a[1] = a[2] = 0; }

// Activations are [B][*][R rows n][N][ch] slabs (rows [row0, row0 + R) of the N origins), G_d has Kd planes, G_o has Ko,
// W is the [Ko][Kd][C][H] slice; `partial`: the forward stops at the raw partial pre-activation, the backward starts from dPre.
int bdgcn_forward_simt(const BdgcnShape& s, const float* X, const float* Go, const float* Gd, const float* W, const float* bias,
                       float* out, void* saved, void* ws, size_t ws_bytes, cudaStream_t st) {
  const long long N = s.N, R = s.R, RN = R * N, NN = N * N, C = s.C, H = s.H, Ko = s.Ko, Kd = s.Kd;
  Carver cv(ws, ws_bytes);
  float* U = cv.take<float>((size_t)s.B * Ko * RN * H);
  float* Z = saved ? static_cast<float*>(saved) : cv.take<float>((size_t)s.B * Kd * RN * C);
  MPGCN_CHECK(cv.ok(), "bdgcn_forward: workspace too small (%zu < %zu bytes)", ws_bytes, cv.off);
  const long long go_sb = s.dynamic ? Ko * NN : 0, gd_sb = s.dynamic ? Kd * NN : 0;   // support batch strides

  {  // Z[b,d,n] (e x l) = G_d^T (e x c) * X[b,n] (c x l)
    SgemmParams p{};
    p.A = Gd; p.B = X; p.D = Z;
    p.M = (int)N; p.N = (int)C; p.K = (int)N;
    p.a_si = 1; p.a_sk = N; p.b_sk = C; p.b_sj = 1; p.d_si = C;
    p.nseg = 1; p.Z0 = s.B; p.Z1 = (int)Kd; p.Z2 = (int)R;
    p.a_sz[0] = gd_sb; p.a_sz[1] = NN; p.a_sz[2] = 0;
    p.b_sz[0] = RN * C; p.b_sz[1] = 0; p.b_sz[2] = N * C;
    p.d_sz[0] = Kd * RN * C; p.d_sz[1] = RN * C; p.d_sz[2] = N * C;
    p.ksplit = 1; p.alpha = 1.f;
    if (int e = simt_sgemm(p, st)) return e;
  }
  {  // U[b,o] (rows x h) = sum_d Z[b,d] (rows x l) * W[o,d] (l x h)
    SgemmParams p{};
    p.A = Z; p.B = W; p.D = U;
    p.M = (int)RN; p.N = (int)H; p.K = (int)C;
    p.a_si = C; p.a_sk = 1; p.b_sk = H; p.b_sj = 1; p.d_si = H;
    p.nseg = (int)Kd; p.a_sseg = RN * C; p.b_sseg = C * H;
    p.Z0 = s.B; p.Z1 = (int)Ko; p.Z2 = 1;
    zero3(p.a_sz); zero3(p.b_sz); zero3(p.d_sz);
    p.a_sz[0] = Kd * RN * C;
    p.b_sz[1] = Kd * C * H;
    p.d_sz[0] = Ko * RN * H; p.d_sz[1] = RN * H;
    p.ksplit = 1; p.alpha = 1.f;
    if (int e = simt_sgemm(p, st)) return e;
  }
  {  // out[b] (m x (e,h)) = act( sum_o G_o[row0.., :]^T (m x n) * U[b,o] (n x (e,h)) + bias[h] ): one k-segment per support
    SgemmParams p{};
    p.A = Go + (long long)s.row0 * N; p.B = U; p.D = out;
    p.M = (int)N; p.N = (int)(N * H); p.K = (int)R;
    p.a_si = 1; p.a_sk = N; p.b_sk = N * H; p.b_sj = 1; p.d_si = N * H;
    p.nseg = (int)Ko; p.a_sseg = NN; p.b_sseg = RN * H;
    p.Z0 = s.B; p.Z1 = 1; p.Z2 = 1;
    zero3(p.a_sz); zero3(p.b_sz); zero3(p.d_sz);
    p.a_sz[0] = go_sb; p.b_sz[0] = Ko * RN * H; p.d_sz[0] = NN * H;
    p.ksplit = 1; p.alpha = 1.f;
    if (!s.partial) { p.bias = bias; p.bias_mod = (int)H; p.relu = s.act; }
    if (int e = simt_sgemm(p, st)) return e;
  }
  return 0;
}

int bdgcn_backward_simt(const BdgcnShape& s, const float* d_out, const float* out, const float* Go, const float* Gd, const float* W,
                        const void* saved, float* dX, float* dW, float* db, void* ws, size_t ws_bytes, cudaStream_t st) {
  const long long N = s.N, R = s.R, RN = R * N, NN = N * N, C = s.C, H = s.H, Ko = s.Ko, Kd = s.Kd;
  const float* Z = static_cast<const float*>(saved);
  MPGCN_CHECK(Z != nullptr, "bdgcn_backward: forward was run without a `saved` buffer");
  Carver cv(ws, ws_bytes);
  float* dPre = cv.take<float>((size_t)s.B * NN * H);
  float* V = cv.take<float>((size_t)s.B * Ko * RN * H);
  float* Y = cv.take<float>((size_t)s.B * Kd * RN * C);
  float* Wq = cv.take<float>((size_t)Ko * Kd * H * C);
  MPGCN_CHECK(cv.ok(), "bdgcn_backward: workspace too small (%zu < %zu bytes)", ws_bytes, cv.off);
  const long long go_sb = s.dynamic ? Ko * NN : 0, gd_sb = s.dynamic ? Kd * NN : 0;

  const float* dP = d_out;         // a partial call receives dPre itself (every origin row m, already masked)
  if (!s.partial) {
    if (db) MPGCN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * H, st));
    if (int e = relu_bwd_prep(d_out, out, s.act, nullptr, dPre, db, (size_t)s.B * NN * H, (int)H, nullptr, st)) return e;
    dP = dPre;
  }

  {  // V[b,o] (n x (e,h)) = G_o[row0.., :] (n x m) * dPre[b] (m x (e,h))
    SgemmParams p{};
    p.A = Go + (long long)s.row0 * N; p.B = dP; p.D = V;
    p.M = (int)R; p.N = (int)(N * H); p.K = (int)N;
    p.a_si = N; p.a_sk = 1; p.b_sk = N * H; p.b_sj = 1; p.d_si = N * H;
    p.nseg = 1; p.Z0 = s.B; p.Z1 = (int)Ko; p.Z2 = 1;
    zero3(p.a_sz); zero3(p.b_sz); zero3(p.d_sz);
    p.a_sz[0] = go_sb; p.a_sz[1] = NN;
    p.b_sz[0] = NN * H;
    p.d_sz[0] = Ko * RN * H; p.d_sz[1] = RN * H;
    p.ksplit = 1; p.alpha = 1.f;
    if (int e = simt_sgemm(p, st)) return e;
  }
  {  // dW[o,d] (l x h) = sum_b Z[b,d]^T (l x rows) * V[b,o] (rows x h)
    MPGCN_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * Ko * Kd * C * H, st));
    SgemmParams p{};
    p.A = Z; p.B = V; p.D = dW;
    p.M = (int)C; p.N = (int)H; p.K = (int)RN;
    p.a_si = 1; p.a_sk = C; p.b_sk = H; p.b_sj = 1; p.d_si = H;
    p.nseg = s.B; p.a_sseg = Kd * RN * C; p.b_sseg = Ko * RN * H;
    p.Z0 = (int)Ko; p.Z1 = (int)Kd; p.Z2 = 1;      // z0 = o, z1 = d
    zero3(p.a_sz); zero3(p.b_sz); zero3(p.d_sz);
    p.a_sz[1] = RN * C;
    p.b_sz[0] = RN * H;
    p.d_sz[0] = Kd * C * H; p.d_sz[1] = C * H;
    long long ks = RN / 2048;
    if (ks < 1) ks = 1;
    if (ks > 256) ks = 256;
    p.ksplit = (int)ks; p.alpha = 1.f;
    if (int e = simt_sgemm(p, st)) return e;
  }
  if (dX) {
    if (int e = permute_w_bwd(W, nullptr, Wq, (int)Ko, (int)Kd, (int)C, (int)H, st)) return e;
    {  // Y[b,d] (rows x l) = sum_o V[b,o] (rows x h) * Wq[d,o] (h x l)
      SgemmParams p{};
      p.A = V; p.B = Wq; p.D = Y;
      p.M = (int)RN; p.N = (int)C; p.K = (int)H;
      p.a_si = H; p.a_sk = 1; p.b_sk = C; p.b_sj = 1; p.d_si = C;
      p.nseg = (int)Ko; p.a_sseg = RN * H; p.b_sseg = H * C;
      p.Z0 = s.B; p.Z1 = (int)Kd; p.Z2 = 1;
      zero3(p.a_sz); zero3(p.b_sz); zero3(p.d_sz);
      p.a_sz[0] = Ko * RN * H;
      p.b_sz[1] = Ko * H * C;
      p.d_sz[0] = Kd * RN * C; p.d_sz[1] = RN * C;
      p.ksplit = 1; p.alpha = 1.f;
      if (int e = simt_sgemm(p, st)) return e;
    }
    {  // dX[b,n] (c x l) = sum_d G_d (c x e) * Y[b,d,n] (e x l)
      SgemmParams p{};
      p.A = Gd; p.B = Y; p.D = dX;
      p.M = (int)N; p.N = (int)C; p.K = (int)N;
      p.a_si = N; p.a_sk = 1; p.b_sk = C; p.b_sj = 1; p.d_si = C;
      p.nseg = (int)Kd; p.a_sseg = NN; p.b_sseg = RN * C;
      p.Z0 = s.B; p.Z1 = (int)R; p.Z2 = 1;
      zero3(p.a_sz); zero3(p.b_sz); zero3(p.d_sz);
      p.a_sz[0] = gd_sb;
      p.b_sz[0] = Kd * RN * C; p.b_sz[1] = N * C;
      p.d_sz[0] = RN * C; p.d_sz[1] = N * C;
      p.ksplit = 1; p.alpha = 1.f;
      if (int e = simt_sgemm(p, st)) return e;
    }
  }
  return 0;
}

}  // namespace mpgcn
