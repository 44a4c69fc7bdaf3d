// Internal host-side declarations shared by the translation units of libmpgcn_b200.
#pragma once

#include "common.cuh"

namespace mpgcn {

const char* last_error();
void prof_enable(int on);
void prof_reset();
int prof_read(int tag, long long* launches, double* flops, double* ms);

// ---- generic fp32 SIMT strided/batched contraction (simt_kernels.cu) ---------------------
//   D[z](i,j) (+)= alpha * sum_seg sum_k A[z](i,k;seg) * B[z](k,j;seg)  (+ bias[j % bias_mod], ReLU)
struct SgemmParams {
  const float* A;
  const float* B;
  float* D;
  int M, N, K;                 // K per segment
  long long a_si, a_sk;        // element strides of A(i,k)
  long long b_sk, b_sj;        // element strides of B(k,j)
  long long d_si;              // D(i,j): j stride 1
  int nseg;
  long long a_sseg, b_sseg;
  int Z0, Z1, Z2;              // batch z = (z0*Z1 + z1)*Z2 + z2
  long long a_sz[3], b_sz[3], d_sz[3];
  int ksplit;                  // > 1: split each segment's K into slices, atomically add into D (D pre-zeroed)
  float alpha;
  const float* bias;
  int bias_mod;
  int relu;
  const float* Cin;            // optional: D = alpha*A*B + beta*Cin (Cin indexed like D); ksplit must be 1
  float beta;
  long long c_sz[3];
};
int simt_sgemm(const SgemmParams& p, cudaStream_t stream);

// ---- elementwise / layout helpers (prep_kernels.cu) ---------------------------------------
int cvt_f32_to_f16(const float* src, __half* dst, size_t n, cudaStream_t s);
// hi = fp16(x), lo = fp16(x - hi): x == hi + lo to ~22 bits
int cvt_f32_to_f16_hilo(const float* src, __half* hi, __half* lo, size_t n, cudaStream_t s);
// delta[p*N + i] = G[p][i][i] - float(fp16(G[p][i][i])) for `planes` N x N matrices
int support_diag_delta(const float* G, float* delta, size_t planes, int N, cudaStream_t s);
// [rows][cols] fp32 -> [rows][ld] fp16 (ld >= cols, padding zeroed)
int cvt_f32_to_f16_padded(const float* src, __half* dst, size_t rows, int cols, int ld, cudaStream_t s);
// d_pre = d_out * (out > 0) (relu) or d_out; fp16 and/or fp32 output; db[h] += sum (db may be null; pre-zeroed)
int relu_bwd_prep(const float* d_out, const float* out, int relu, __half* d_pre16, float* d_pre32, float* db, size_t n, int H,
                  const float* scale, cudaStream_t s);
// the same with the ReLU mask taken from an fp16 copy of the forward output (tensor-core path: fp16 d_pre only, H % 4 == 0)
int relu_bwd_prep_f16mask(const float* d_out, const __half* out16, int relu, __half* d_pre16, float* db, size_t n, int H,
                          const float* scale, cudaStream_t s);
// scale2[0] = S = 2^k with S*max|d_out| in [16,32), scale2[1] = 1/S (S = 1 for an all-zero or non-finite input).
// fp16 has 5 exponent bits: realistic gradients (MSE mean over B*N*N cells ~ 1e-7) must be rescaled before the cast.
// absmax_hint: optional device scalar already holding max|d_out| (produced by the epilogue that wrote d_out): skips the pass
int grad_scale_prepare(const float* d_out, size_t n, float* scale2, const float* absmax_hint, cudaStream_t s);
// W[o][d][l][h] fp32 (o < Ko, d < Kd) -> Wq[d][o][h][l] (fp16 and/or fp32)
int permute_w_bwd(const float* W, __half* wq16, float* wq32, int Ko, int Kd, int C, int H, cudaStream_t s);
// dW[o][d][l][h] = sum_slices P[slice][mt][(d%4)*32 + l][o][h]   (C = H = 32; o < Ko, d < Kd)
int reduce_dw_partials(const float* P, float* dW, int slices, int MT, int Ko, int Kd, const float* inv_scale, cudaStream_t s);
// out[p][i] = (row0 <= i < row0 + rows) ? delta[p][i] : 0   (diagonal remainders restricted to an origin-row slab)
int mask_delta_rows(const float* delta, float* out, size_t planes, int N, int row0, int rows, cudaStream_t s);
// x[i] = act(x[i] + bias[i % H]) in place (bias nullable; act 0 none / 1 ReLU): the epilogue a partial layer call leaves out
int bias_act_inplace(float* x, const float* bias, int act, size_t n, int H, cudaStream_t s);
// exchange steps of the origin-row shard fused into elementwise kernels over peer memory (parts / dsts: HOST arrays of g <= 8 DEVICE
// pointers to [B][N][N][H] buffers, the rank's own and its peers' NVLink-mapped ones)
// part_rows: origin rows held by each part buffer -- N (whole [B][N][N][H] partials, the rank's rows start at row0) or `rows`
// (staging slots [B][rows][N][H] that already hold only the rank's rows)
int rows_reduce_bias_act(float* out, const float* const* parts, int g, const float* bias, int act, int B, int N, int row0, int rows, int part_rows,
                         int H, cudaStream_t s);
// the same with the fp16 cast of the tensor-core path folded in: scale2 = [S, 1/S] from the GLOBAL max|d_out| (device scalar), values
// stored as fp16(S * d_pre) -- half the bytes on the wire and no cast / absmax pass over the gathered tensor on any rank
int relu_backward_scatter_f16(const float* d_out, const float* out, int act, __half* const* dsts, int g, float* db, const float* absmax,
                              float* scale2, int B, int N, int row0, int rows, int H, cudaStream_t s);
int absmax_f32(const float* x, size_t n, float* out, cudaStream_t s);
int relu_backward_scatter(const float* d_out, const float* out, int act, float* const* dsts, int g, float* db, int B, int N, int row0, int rows,
                          int H, cudaStream_t s);

// ---- per-cell LSTM, last hidden state (lstm_kernels.cu) ------------------------------------
int lstm_last_forward(const float* x_seq, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* hT,
                      int B, int T, long long NN, int C, cudaStream_t s);
int lstm_last_backward(const float* x_seq, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                       const float* d_hT, float* d_w_ih, float* d_w_hh, float* d_b_ih, float* d_b_hh, float* d_x, int B, int T,
                       long long NN, int C, cudaStream_t s);

// FC head + branch mean (head_kernels.cu); g / dg are HOST arrays of M device pointers
int head_forward(const float* const* g, const float* w, const float* bias, float* y, float* pre, long long cells, int C, int M,
                 cudaStream_t st);
int head_backward(const float* const* g, const float* w, const float* pre, const float* dy, float* const* dg, float* dw, float* db,
                  float* dg_absmax /*[M] or null*/, long long cells, int C, int M, cudaStream_t st);

// support-matrix builder (adj_kernels.cu): reference GCN.Adj_Processor.process
enum AdjKernel { ADJ_LOCALPOOL = 0, ADJ_CHEBYSHEV = 1, ADJ_RANDOM_WALK = 2, ADJ_DUAL_RANDOM_WALK = 3 };
int adj_num_supports(int kernel_type, int K);
size_t adj_workspace_bytes(int B, int N, int kernel_type, int K);
int adj_process(const float* flow, float* supports, int B, int N, int kernel_type, int K, void* ws, size_t ws_bytes, cudaStream_t st);

// dynamic O / D graphs from the OD history (dyn_graph_kernels.cu)
size_t dyn_graph_workspace_bytes(int P, int N);
int dyn_graph_build(const float* od_hist, int periods, float* o_g, float* d_g, int P, int N, void* ws, size_t ws_bytes, cudaStream_t st);

// tcgen05 LSTM (lstm_tc.cu): hidden size 32 only
bool lstm_tc_supported(int T, int C);
size_t lstm_tc_bwd_workspace_bytes(int B, int T, long long NN);
size_t lstm_tc_saved_bytes(int B, int T, long long NN);
// saved (nullable): training state c_t, h_t written by the forward; the backward walks it instead of recomputing the forward
int lstm_last_forward_tc(const float* x_seq, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* hT,
                         void* saved, int B, int T, long long NN, cudaStream_t s);
int lstm_last_backward_tc(const float* x_seq, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                          const float* d_hT, float* d_w_ih, float* d_w_hh, float* d_b_ih, float* d_b_hh, float* d_x, const void* saved,
                          int B, int T, long long NN, void* ws, size_t ws_bytes, const float* d_hT_absmax, cudaStream_t s);

// ---- BDGCN layer orchestration ---------------------------------------------------------------
struct BdgcnShape {
  int B, N, K, C, H;
  int dynamic;      // supports are per-sample [B,K,N,N] pairs
  int act;          // 0 none, 1 relu
  // Which PART of the layer this call evaluates (multi-GPU shards); a whole layer has
  // R = N, row0 = 0, Ko = Kd = K, partial = 0.
  int R, row0;      // origin rows n in [row0, row0 + R) are present: X / Z / U / V / Y / dX are [B, R, N, *] slabs
  int Ko, Kd;       // supports in G_o / in G_d; W is the [Ko*Kd*C, H] slice in (o, d, l) row order
  int partial;      // forward writes the raw partial pre-activation sum_{o, n in slab} ... (no bias, no activation);
                    // backward receives dPre (already masked) instead of dOut
  // forward of a part, optional: PUSH the partial into peer memory instead of writing it locally -- row m goes to
  // peer_out[m / (N / peer_g)] (that owner's staging buffer [peer_g slots][B][N / peer_g][N][H]), slot peer_rank
  int peer_g, peer_rank;
  float* peer_out[8];
  bool whole() const { return R == N && row0 == 0 && Ko == K && Kd == K && !partial; }
};
enum Precision { PREC_FP32_SIMT = 0, PREC_FP16_TC = 1 };

bool tc_supported(const BdgcnShape& s);
size_t bdgcn_saved_bytes(const BdgcnShape& s, int precision);
size_t bdgcn_fwd_workspace_bytes(const BdgcnShape& s, int precision);
size_t bdgcn_bwd_workspace_bytes(const BdgcnShape& s, int precision);

int bdgcn_forward_simt(const BdgcnShape& s, const float* X, const float* Go, const float* Gd, const float* W, const float* bias,
                       float* out, void* saved, void* ws, size_t ws_bytes, cudaStream_t st);
int bdgcn_backward_simt(const BdgcnShape& s, const float* d_out, const float* out, const float* Go, const float* Gd, const float* W,
                        const void* saved, float* dX, float* dW, float* db, void* ws, size_t ws_bytes, cudaStream_t st);
// optional side inputs / outputs of the tensor-core layer (mirror of mpgcn_bdgcn_extras in include/mpgcn_b200.h); all nullable
struct BdgcnExtras {
  const void* go_prepared = nullptr;    // supports already converted by bdgcn_prepare_supports (fp16 padded + diagonal remainders)
  const void* gd_prepared = nullptr;
  const void* x_f16 = nullptr;          // forward: fp16 copy of X (skips the conversion pass)
  void* out_f16 = nullptr;              // forward: receives an fp16 copy of out;  backward: that copy (ReLU mask source instead of out)
  const float* d_out_absmax = nullptr;  // backward: max|d_out| already known
  float* dx_absmax = nullptr;           // backward: receives max|dX|
  const void* d_pre_f16 = nullptr;      // backward of a PART: dPre [B,N,N,H] already masked, scaled by scale2[0] and cast to fp16
  const float* d_pre_scale2 = nullptr;  //   ... with its device [S, 1/S] pair (mpgcn_relu_backward_scatter_f16 produces both)
};
size_t bdgcn_supports_prepared_bytes(long long planes, int N);
int bdgcn_prepare_supports(const float* G, void* prepared, long long planes, int N, cudaStream_t st);
int bdgcn_forward_tc(const BdgcnShape& s, const float* X, const float* Go, const float* Gd, const float* W, const float* bias,
                     float* out, void* saved, void* ws, size_t ws_bytes, const BdgcnExtras& ex, cudaStream_t st);
int bdgcn_backward_tc(const BdgcnShape& s, const float* d_out, const float* out, const float* Go, const float* Gd, const float* W,
                      const void* saved, float* dX, float* dW, float* db, void* ws, size_t ws_bytes, const BdgcnExtras& ex,
                      cudaStream_t st);

}  // namespace mpgcn
