// fold.cu — the two constant-input column blocks of the MAVAE folded out of the dense layers.
//
// (1) Action embeddings (reference model.py:121,146,158-164).  The decoder input is [z | e_0 .. e_{A-1}] with
//     e_a = table_a[act_a]: only n_act distinct rows per agent.  Layer 0 of both decoders therefore sees
//         sum_a W0[:, cols_a] . table_a[act_a]  =  sum_a T_a[:, act_a],     T_a = W0[:, cols_a] . table_a^T   [2H x n_act]
//     so the K = A*C action columns of that GEMM become K = A*n_act one-hot columns against T (64 -> 5 per agent):
//         forward   H0 = relu([z | onehot(act)] . [W0_z | T]^T + b)                 (gemm second K segment, kernels.h)
//         backward  dT = dH0^T . onehot   (one more wgrad with N = A*n_act), then, exactly,
//                   dW0[:, cols_a] = dT_a . table_a          dtable_a = dT_a^T . W0[:, cols_a]
//     T is built from the fp32 master weights (fp32 accumulation) every step and rounded once to the operand type.
//
// (2) Agent-id embedding (model.py:113,142-143).  With the codebook index column (create_dataset, trainer.py:21: row b of
//     agent a carries index a) the first C_I input columns of encoder a are the constant emb[a], so
//         W0_a . [emb[a] | obs] + b0_a  =  W0_a[:, I:] . obs + (b0_a + W0_a[:, :I] . emb[a])
//     i.e. a per-agent bias; backward: d emb[a] = W0_a[:, :I]^T . db0_a and dW0_a[:, :I] = db0_a (x) emb[a], with db0_a
//     the bias gradient the step computes anyway.  An explicit per-row index column keeps the dense path.
#include <algorithm>

#include "kernels.h"

namespace mfvae {

constexpr int kFoldThreads = 256;

// one-hot action columns of the decoder input: zin[b][col0 + a*nmax + k] = (act[b][a] == k), zero elsewhere up to `width`
template <typename T>
__global__ void __launch_bounds__(kFoldThreads) onehot_kernel(const float* __restrict__ act, int act_ld, const int32_t* __restrict__ n_act,
                                                              T* __restrict__ zin, int64_t zin_ld, int col0, int A, int nmax, int width, int64_t B) {
  const int64_t total = B * width;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % width);
    const int64_t b = i / width;
    const int a = j / nmax, k = j - a * nmax;
    float v = 0.f;
    if (a < A) {
      int kk = static_cast<int>(__ldg(act + b * act_ld + a));
      kk = max(0, min(kk, n_act[a] - 1));             // same clamp as every other consumer of the action code
      v = (kk == k) ? 1.f : 0.f;
    }
    zin[b * zin_ld + col0 + j] = from_f<T>(v);
  }
}

int launch_onehot(const float* act, int act_ld, const int32_t* n_act, void* zin, int dtype, int64_t zin_ld, int col0, int A, int nmax,
                  int width, int64_t B, cudaStream_t s) {
  const int64_t total = B * width;
  const int grid = static_cast<int>(std::min<int64_t>((total + kFoldThreads - 1) / kFoldThreads, kNumSMs * 8));
  if (dtype == kBF16) onehot_kernel<__nv_bfloat16><<<grid, kFoldThreads, 0, s>>>(act, act_ld, n_act, static_cast<__nv_bfloat16*>(zin), zin_ld, col0, A, nmax, width, B);
  else                onehot_kernel<float><<<grid, kFoldThreads, 0, s>>>(act, act_ld, n_act, static_cast<float*>(zin), zin_ld, col0, A, nmax, width, B);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

// T[h][a*nmax + k] = sum_c W0[h][wcol0 + a*C + c] * table[a][k][c]     one warp per (h, a); columns >= A*nmax are zeroed
template <typename T>
__global__ void __launch_bounds__(kFoldThreads) act_fold_fwd_kernel(const float* __restrict__ W0, int64_t w_ld, int wcol0,
                                                                    const float* __restrict__ table, int64_t table_gs,
                                                                    T* __restrict__ Tt, int64_t t_ld, int rows, int A, int C, int nmax) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t total = static_cast<int64_t>(rows) * A;
  for (int64_t u = warp; u < total; u += nwarps) {
    const int a = static_cast<int>(u % A);
    const int64_t h = u / A;
    const float* w = W0 + h * w_ld + wcol0 + a * C;
    const float* tb = table + a * table_gs;
    for (int k = 0; k < nmax; ++k) {
      float acc = 0.f;
      for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(w + c), __ldg(tb + static_cast<int64_t>(k) * C + c), acc);
      acc = warp_sum(acc);
      if (lane == 0) Tt[h * t_ld + a * nmax + k] = from_f<T>(acc);
    }
    if (a == A - 1) for (int j = A * nmax + lane; j < t_ld; j += 32) Tt[h * t_ld + j] = from_f<T>(0.f);
  }
}

int launch_act_fold_fwd(const float* W0, int64_t w_ld, int wcol0, const float* table, int64_t table_gs, void* Tt, int dtype, int64_t t_ld,
                        int rows, int A, int C, int nmax, cudaStream_t s) {
  const int64_t warps = static_cast<int64_t>(rows) * A;
  const int grid = static_cast<int>(std::min<int64_t>((warps * 32 + kFoldThreads - 1) / kFoldThreads, kNumSMs * 8));
  if (dtype == kBF16) act_fold_fwd_kernel<__nv_bfloat16><<<grid, kFoldThreads, 0, s>>>(W0, w_ld, wcol0, table, table_gs, static_cast<__nv_bfloat16*>(Tt), t_ld, rows, A, C, nmax);
  else                act_fold_fwd_kernel<float><<<grid, kFoldThreads, 0, s>>>(W0, w_ld, wcol0, table, table_gs, static_cast<float*>(Tt), t_ld, rows, A, C, nmax);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

// From dT [rows][t_ld] (fp32, complete):
//   gW0[h][wcol0 + a*C + c]  = sum_k dT[h][a*nmax + k] * table[a][k][c]        (plain store: these columns have no other writer)
//   gtable[a][k][c]         += sum_h dT[h][a*nmax + k] * W0[h][wcol0 + a*C + c] (atomics into the zeroed gradient arena)
// grid = (row chunks, A); block = (C columns) x (256 / C row phases).  NMAX = compile-time bound on n_act.
template <int NMAX>
__global__ void __launch_bounds__(kFoldThreads) act_fold_bwd_kernel(const float* __restrict__ dT, int64_t t_ld, const float* __restrict__ W0,
                                                                    float* __restrict__ gW0, int64_t w_ld, int wcol0,
                                                                    const float* __restrict__ table, float* __restrict__ gtable, int64_t table_gs,
                                                                    int rows, int C, int nmax, int rows_per_cta) {
  __shared__ float fold[kFoldThreads];
  const int a = blockIdx.y;
  const int c = threadIdx.x % C, rphase = threadIdx.x / C, nph = blockDim.x / C;
  float tab[NMAX], acc[NMAX];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) { tab[k] = (k < nmax) ? __ldg(table + a * table_gs + static_cast<int64_t>(k) * C + c) : 0.f; acc[k] = 0.f; }
  const int h0 = blockIdx.x * rows_per_cta, h1 = min(rows, h0 + rows_per_cta);
  if (rphase < nph) {
    for (int h = h0 + rphase; h < h1; h += nph) {
      const float* d = dT + static_cast<int64_t>(h) * t_ld + a * nmax;
      const int64_t wi = static_cast<int64_t>(h) * w_ld + wcol0 + a * C + c;
      const float w = __ldg(W0 + wi);
      float o = 0.f;
#pragma unroll
      for (int k = 0; k < NMAX; ++k) {
        if (k < nmax) { const float dk = __ldg(d + k); o = fmaf(dk, tab[k], o); acc[k] = fmaf(dk, w, acc[k]); }
      }
      gW0[wi] = o;
    }
  }
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    if (k >= nmax) break;
    __syncthreads();
    fold[threadIdx.x] = (rphase < nph) ? acc[k] : 0.f;
    __syncthreads();
    if (rphase == 0) {
      float t = fold[c];
      for (int r = 1; r < nph; ++r) t += fold[r * C + c];
      atomicAdd(gtable + a * table_gs + static_cast<int64_t>(k) * C + c, t);
    }
  }
}

int launch_act_fold_bwd(const float* dT, int64_t t_ld, const float* W0, float* gW0, int64_t w_ld, int wcol0, const float* table, float* gtable,
                        int64_t table_gs, int rows, int A, int C, int nmax, cudaStream_t s) {
  MFVAE_CHECK(C <= kFoldThreads && nmax <= 16, "act fold: act_features <= 256 and n_act <= 16");
  const int rows_per_cta = 64;
  dim3 grid((rows + rows_per_cta - 1) / rows_per_cta, A);
  if (nmax <= 8) act_fold_bwd_kernel<8><<<grid, kFoldThreads, 0, s>>>(dT, t_ld, W0, gW0, w_ld, wcol0, table, gtable, table_gs, rows, C, nmax, rows_per_cta);
  else           act_fold_bwd_kernel<16><<<grid, kFoldThreads, 0, s>>>(dT, t_ld, W0, gW0, w_ld, wcol0, table, gtable, table_gs, rows, C, nmax, rows_per_cta);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

// eb[a][n] = b0[a][n] + sum_i W0[a][n][i] * emb[a][i]            one warp per (a, n)
__global__ void __launch_bounds__(kFoldThreads) enc_bias_fold_kernel(const float* __restrict__ W0, int64_t w_ld, const float* __restrict__ b0,
                                                                     const float* __restrict__ emb, float* __restrict__ eb, int A, int N, int I) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int u = warp; u < A * N; u += nwarps) {
    const int a = u / N;
    const float* w = W0 + static_cast<int64_t>(u) * w_ld;
    float acc = 0.f;
    for (int i = lane; i < I; i += 32) acc = fmaf(__ldg(w + i), __ldg(emb + static_cast<int64_t>(a) * I + i), acc);
    acc = warp_sum(acc);
    if (lane == 0) eb[u] = b0[u] + acc;
  }
}

int launch_enc_bias_fold(const float* W0, int64_t w_ld, const float* b0, const float* emb, float* eb, int A, int N, int I, cudaStream_t s) {
  const int grid = std::min((A * N * 32 + kFoldThreads - 1) / kFoldThreads, kNumSMs * 4);
  enc_bias_fold_kernel<<<grid, kFoldThreads, 0, s>>>(W0, w_ld, b0, emb, eb, A, N, I);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

// g_emb[a][i] += sum_n W0[a][n][i] * db0[a][n]        gW0[a][n][i] = db0[a][n] * emb[a][i]   (i < I)       one CTA per agent
__global__ void __launch_bounds__(kFoldThreads) emb_grad_fold_kernel(const float* __restrict__ W0, float* __restrict__ gW0, int64_t w_ld,
                                                                     const float* __restrict__ db0, const float* __restrict__ emb,
                                                                     float* __restrict__ g_emb, int N, int I) {
  const int a = blockIdx.x;
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    const float e = emb[static_cast<int64_t>(a) * I + i];
    float acc = 0.f;
    for (int n = 0; n < N; ++n) {
      const int64_t wi = (static_cast<int64_t>(a) * N + n) * w_ld + i;
      const float d = db0[a * N + n];
      acc = fmaf(__ldg(W0 + wi), d, acc);
      gW0[wi] = d * e;
    }
    atomicAdd(g_emb + static_cast<int64_t>(a) * I + i, acc);
  }
}

int launch_emb_grad_fold(const float* W0, float* gW0, int64_t w_ld, const float* db0, const float* emb, float* g_emb, int A, int N, int I,
                         cudaStream_t s) {
  emb_grad_fold_kernel<<<A, kFoldThreads, 0, s>>>(W0, gW0, w_ld, db0, emb, g_emb, N, I);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mfvae
