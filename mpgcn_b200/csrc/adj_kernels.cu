// Support-matrix builder on the GPU: the arithmetic of the reference's Adj_Processor.process
// (/root/reference/GCN.py:56-138), batched over the whole [B,N,N] flow tensor instead of a Python loop over the batch on
// the CPU (the reference calls it twice per training step, Model_Trainer.py:84,106).  SURVEY.md section 8(f) rank 1.
//
//   localpool                   I + D^-1/2 A D^-1/2                                   (GCN.py:69-72, 111-114)
//   chebyshev                   x = (2/lambda_max) (I - D^-1/2 A D^-1/2) - I with lambda_max = 2, the branch the reference
//                               always takes on torch >= 2 (torch.eig was removed; bare except, GCN.py:117-126)
//   random_walk_diffusion       x = (D^-1 A)^T, 1/0 -> 0                               (GCN.py:79-82, 103-108)
//   dual_random_walk_diffusion  forward series of (D^-1 A)^T, backward series of (D_T^-1 A^T)^T   (GCN.py:84-91)
//   series: T_0 = I, T_1 = x, T_k = 2 x T_{k-1} - T_{k-2}                              (GCN.py:128-138)
// All fp32: the N^3 recursion runs on the exact CUDA-core SGEMM with a fused "2 A B - C" epilogue.
#include "kernels.h"

namespace mpgcn {

int adj_num_supports(int kernel_type, int K) {
  switch (kernel_type) {
    case ADJ_LOCALPOOL: return 1;
    case ADJ_CHEBYSHEV:
    case ADJ_RANDOM_WALK: return K + 1;
    case ADJ_DUAL_RANDOM_WALK: return 2 * K + 1;
    default: return -1;
  }
}

size_t adj_workspace_bytes(int B, int N, int kernel_type, int K) {
  (void)kernel_type; (void)K;
  return 256 + 2 * align_up((size_t)B * N * sizeof(float), 256);      // row sums and column sums
}

// sums[b][i] = sum_j A[b][i][j] (by_col = 0) or sum_j A[b][j][i] (by_col = 1); one warp per (b, i)
__global__ void adj_sums_kernel(const float* __restrict__ A, float* __restrict__ sums, int B, int N, int by_col) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)B * N) return;
  const int b = (int)(w / N), i = (int)(w % N);
  const float* base = A + (size_t)b * N * N;
  float s = 0.f;
  for (int j = lane; j < N; j += 32) s += by_col ? base[(size_t)j * N + i] : base[(size_t)i * N + j];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) sums[w] = s;
}

// mode 0: out = I + sym_norm(A)                      (localpool)
// mode 1: out = (1 * (I - sym_norm(A))) - I            (chebyshev x, lambda_max = 2: same operation order as the reference)
// mode 2: out = (D^-1 A)^T                            (random walk, forward)         out[i][j] = A[j][i] / rowsum[j]
// mode 3: out = (D_T^-1 A^T)^T                        (random walk, backward)        out[i][j] = A[i][j] / colsum[j]
// `out` is support k_out of a [B][Ks][N][N] stack; optionally the identity is written to support 0.
__global__ void adj_build_kernel(const float* __restrict__ A, const float* __restrict__ rowsum, const float* __restrict__ colsum,
                                 float* __restrict__ sup, int B, int N, int Ks, int k_out, int mode, int write_identity) {
  const size_t total = (size_t)B * N * N;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int j = (int)(t % N);
    const int i = (int)((t / N) % N);
    const int b = (int)(t / ((size_t)N * N));
    const float* Ab = A + (size_t)b * N * N;
    float v;
    if (mode <= 1) {
      const float di = powf(rowsum[(size_t)b * N + i], -0.5f), dj = powf(rowsum[(size_t)b * N + j], -0.5f);
      const float an = (di * Ab[(size_t)i * N + j]) * dj;
      const float id = (i == j) ? 1.f : 0.f;
      v = (mode == 0) ? id + an : (1.0f * (id - an)) - id;
    } else if (mode == 2) {
      const float s = rowsum[(size_t)b * N + j];
      const float dinv = 1.f / s;
      v = (isinf(dinv) ? 0.f : dinv) * Ab[(size_t)j * N + i];
    } else {
      const float s = colsum[(size_t)b * N + j];
      const float dinv = 1.f / s;
      v = (isinf(dinv) ? 0.f : dinv) * Ab[(size_t)i * N + j];
    }
    float* Sb = sup + (size_t)b * Ks * N * N;
    Sb[(size_t)k_out * N * N + (size_t)i * N + j] = v;
    if (write_identity) Sb[(size_t)i * N + j] = (i == j) ? 1.f : 0.f;
  }
}

static unsigned adj_grid(size_t work, int threads) {
  size_t b = (work + threads - 1) / threads;
  const size_t cap = (size_t)device_sm_count() * 16;
  return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// T_k = 2 * x * T_{k-1} - T_{k-2} for every batch element; x = support k_x, T's are supports of the same stack
static int cheb_step(float* sup, int B, int N, int Ks, int k_x, int k_prev, int k_prev2, int k_out, cudaStream_t st) {
  const long long NN = (long long)N * N;
  SgemmParams p{};
  p.A = sup + k_x * NN; p.B = sup + k_prev * NN; p.D = sup + k_out * NN; p.Cin = sup + k_prev2 * NN;
  p.M = N; p.N = N; p.K = N;
  p.a_si = N; p.a_sk = 1; p.b_sk = N; p.b_sj = 1; p.d_si = N;
  p.nseg = 1; p.Z0 = B; p.Z1 = 1; p.Z2 = 1;
  for (int i = 0; i < 3; ++i) { p.a_sz[i] = 0; p.b_sz[i] = 0; p.d_sz[i] = 0; p.c_sz[i] = 0; }
  p.a_sz[0] = p.b_sz[0] = p.d_sz[0] = p.c_sz[0] = (long long)Ks * NN;
  p.ksplit = 1; p.alpha = 2.f; p.beta = -1.f;
  return simt_sgemm(p, st);
}

int adj_process(const float* flow, float* supports, int B, int N, int kernel_type, int K, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int Ks = adj_num_supports(kernel_type, K);
  MPGCN_CHECK(Ks >= 1, "Invalid kernel_type. Must be one of [chebyshev, localpool, random_walk_diffusion, dual_random_walk_diffusion].");
  MPGCN_CHECK(B >= 1 && N >= 1 && K >= 0, "adj_process: bad shape B=%d N=%d K=%d", B, N, K);
  MPGCN_CHECK(ws != nullptr && ws_bytes >= adj_workspace_bytes(B, N, kernel_type, K), "adj_process: workspace too small");
  float* rowsum = static_cast<float*>(ws);
  float* colsum = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + align_up((size_t)B * N * sizeof(float), 256));
  const size_t warps = (size_t)B * N;
  prof_count(PROF_ELEMENTWISE);
  adj_sums_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(flow, rowsum, B, N, 0);
  if (kernel_type == ADJ_DUAL_RANDOM_WALK) {
    prof_count(PROF_ELEMENTWISE);
    adj_sums_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(flow, colsum, B, N, 1);
  }
  const size_t total = (size_t)B * N * N;
  prof_count(PROF_ELEMENTWISE);
  if (kernel_type == ADJ_LOCALPOOL) {
    adj_build_kernel<<<adj_grid(total, 256), 256, 0, st>>>(flow, rowsum, colsum, supports, B, N, Ks, 0, 0, 0);
    MPGCN_CUDA(cudaGetLastError());
    return 0;
  }
  if (K == 0) {      // only T_0 = I: write x into a scratch-free way by building the identity alone
    adj_build_kernel<<<adj_grid(total, 256), 256, 0, st>>>(flow, rowsum, colsum, supports, B, N, Ks, 0, 2, 1);
    MPGCN_CUDA(cudaGetLastError());   // support 0 first receives x, then the identity (same thread, program order)
    return 0;
  }
  const int mode = (kernel_type == ADJ_CHEBYSHEV) ? 1 : 2;
  adj_build_kernel<<<adj_grid(total, 256), 256, 0, st>>>(flow, rowsum, colsum, supports, B, N, Ks, 1, mode, 1);   // T_0 = I, T_1 = x
  MPGCN_CUDA(cudaGetLastError());
  for (int k = 2; k <= K; ++k)
    if (int e = cheb_step(supports, B, N, Ks, 1, k - 1, k - 2, k, st)) return e;
  if (kernel_type == ADJ_DUAL_RANDOM_WALK) {     // backward series occupies supports K+1 .. 2K; its T_0 is the shared identity
    prof_count(PROF_ELEMENTWISE);
    adj_build_kernel<<<adj_grid(total, 256), 256, 0, st>>>(flow, rowsum, colsum, supports, B, N, Ks, K + 1, 3, 0);
    MPGCN_CUDA(cudaGetLastError());
    for (int k = 2; k <= K; ++k)
      if (int e = cheb_step(supports, B, N, Ks, K + 1, K + k - 1, k == 2 ? 0 : K + k - 2, K + k, st)) return e;
  }
  return 0;
}

}  // namespace mpgcn
