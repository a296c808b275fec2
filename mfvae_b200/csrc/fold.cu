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

// T[h][a*nmax + k] = sum_c W0[h][wcol0 + a*C + c] * table[a][k][c]; columns >= A*nmax are zeroed.
// One CTA per kRowsF rows of W0: the rows' action columns go through shared memory (coalesced 16-byte loads), thread j owns
// output column j = a*nmax + k and reads its table row through the read-only cache (the whole table is A*nmax*C floats).
constexpr int kRowsF = 4;
template <typename T>
__global__ void __launch_bounds__(kFoldThreads) act_fold_fwd_kernel(const float* __restrict__ W0, int64_t w_ld, int wcol0,
                                                                    const float* __restrict__ table, int64_t table_gs,
                                                                    T* __restrict__ Tt, int64_t t_ld, int rows, int A, int C, int nmax) {
  extern __shared__ float wrow[];                      // [kRowsF][A][C + 4]: the agent pitch is padded by one 16-byte chunk, so the
  const int AC = A * C, CP = C + 4, ACP = A * CP;      // lanes of a warp (5 k's of ~7 agents each) read distinct banks
  const int h0 = blockIdx.x * kRowsF;
  for (int i = threadIdx.x; i < kRowsF * (AC / 4); i += blockDim.x) {
    const int r = i / (AC / 4), q = i - r * (AC / 4);
    const int a = (q * 4) / C, c = q * 4 - a * C;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h0 + r < rows) v = ldg_stream4(W0 + static_cast<int64_t>(h0 + r) * w_ld + wcol0 + q * 4);
    *reinterpret_cast<float4*>(wrow + r * ACP + a * CP + c) = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < t_ld; j += blockDim.x) {
    float acc[kRowsF];
#pragma unroll
    for (int r = 0; r < kRowsF; ++r) acc[r] = 0.f;
    if (j < A * nmax) {
      const int a = j / nmax, k = j - a * nmax;
      const float4* tb = reinterpret_cast<const float4*>(table + a * table_gs + static_cast<int64_t>(k) * C);
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 t = __ldg(tb + c4);
#pragma unroll
        for (int r = 0; r < kRowsF; ++r) {
          const float4 w = reinterpret_cast<const float4*>(wrow + r * ACP + a * CP)[c4];
          acc[r] = fmaf(w.x, t.x, fmaf(w.y, t.y, fmaf(w.z, t.z, fmaf(w.w, t.w, acc[r]))));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsF; ++r)
      if (h0 + r < rows) Tt[static_cast<int64_t>(h0 + r) * t_ld + j] = from_f<T>(acc[r]);
  }
}

int launch_act_fold_fwd(const float* W0, int64_t w_ld, int wcol0, const float* table, int64_t table_gs, void* Tt, int dtype, int64_t t_ld,
                        int rows, int A, int C, int nmax, cudaStream_t s) {
  MFVAE_CHECK(C % 4 == 0 && w_ld % 4 == 0 && wcol0 % 4 == 0 && table_gs % 4 == 0, "act fold: widths must be multiples of 4");
  const size_t smem = static_cast<size_t>(kRowsF) * A * (C + 4) * sizeof(float);
  MFVAE_CHECK(smem <= 160 * 1024, "act fold: A * act_features too large for the shared-memory row buffer");
  const int grid = (rows + kRowsF - 1) / kRowsF;
  if (dtype == kBF16) {
    static bool set16 = false;
    if (!set16) { MFVAE_CUDA(cudaFuncSetAttribute(act_fold_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); set16 = true; }
    act_fold_fwd_kernel<__nv_bfloat16><<<grid, kFoldThreads, smem, s>>>(W0, w_ld, wcol0, table, table_gs, static_cast<__nv_bfloat16*>(Tt), t_ld, rows, A, C, nmax);
  } else {
    static bool set32 = false;
    if (!set32) { MFVAE_CUDA(cudaFuncSetAttribute(act_fold_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); set32 = true; }
    act_fold_fwd_kernel<float><<<grid, kFoldThreads, smem, s>>>(W0, w_ld, wcol0, table, table_gs, static_cast<float*>(Tt), t_ld, rows, A, C, nmax);
  }
  MFVAE_LAUNCH_CHECK();
  return 0;
}

// From dT [rows][t_ld] (fp32, complete):
//   gW0[h][wcol0 + a*C + c]  = sum_k dT[h][a*nmax + k] * table[a][k][c]        (plain store: these columns have no other writer)
//   gtable[a][k][c]         += sum_h dT[h][a*nmax + k] * W0[h][wcol0 + a*C + c] (atomics into the zeroed gradient arena)
// grid = (row chunks of kRowsB, A); a thread owns 4 consecutive columns c (16-byte accesses) and walks the chunk's rows.
template <int NMAX>
__global__ void __launch_bounds__(kFoldThreads) act_fold_bwd_kernel(const float* __restrict__ dT, int64_t t_ld, const float* __restrict__ W0,
                                                                    float* __restrict__ gW0, int64_t w_ld, int wcol0,
                                                                    const float* __restrict__ table, float* __restrict__ gtable, int64_t table_gs,
                                                                    int rows, int C, int nmax, int rows_per_cta) {
  __shared__ float4 fold[kFoldThreads];
  const int a = blockIdx.y;
  const int strips = C / 4;
  const int strip = threadIdx.x % strips, rphase = threadIdx.x / strips, nph = blockDim.x / strips;
  float4 tab[NMAX], acc[NMAX];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    tab[k] = (k < nmax) ? __ldg(reinterpret_cast<const float4*>(table + a * table_gs + static_cast<int64_t>(k) * C) + strip) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int h0 = blockIdx.x * rows_per_cta, h1 = min(rows, h0 + rows_per_cta);
  if (rphase < nph) {
    for (int h = h0 + rphase; h < h1; h += nph) {
      const float* d = dT + static_cast<int64_t>(h) * t_ld + a * nmax;
      const int64_t wi = static_cast<int64_t>(h) * w_ld + wcol0 + a * C + strip * 4;
      const float4 w = ldg_stream4(W0 + wi);
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < NMAX; ++k) {
        if (k < nmax) {
          const float dk = __ldg(d + k);
          o.x = fmaf(dk, tab[k].x, o.x); o.y = fmaf(dk, tab[k].y, o.y); o.z = fmaf(dk, tab[k].z, o.z); o.w = fmaf(dk, tab[k].w, o.w);
          acc[k].x = fmaf(dk, w.x, acc[k].x); acc[k].y = fmaf(dk, w.y, acc[k].y); acc[k].z = fmaf(dk, w.z, acc[k].z); acc[k].w = fmaf(dk, w.w, acc[k].w);
        }
      }
      *reinterpret_cast<float4*>(gW0 + wi) = o;
    }
  }
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    if (k >= nmax) break;
    __syncthreads();
    fold[threadIdx.x] = (rphase < nph) ? acc[k] : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (rphase == 0) {
      float4 t = fold[strip];
      for (int r = 1; r < nph; ++r) { const float4 u = fold[r * strips + strip]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
      float* o = gtable + a * table_gs + static_cast<int64_t>(k) * C + strip * 4;
      atomicAdd(o, t.x); atomicAdd(o + 1, t.y); atomicAdd(o + 2, t.z); atomicAdd(o + 3, t.w);
    }
  }
}

int launch_act_fold_bwd(const float* dT, int64_t t_ld, const float* W0, float* gW0, int64_t w_ld, int wcol0, const float* table, float* gtable,
                        int64_t table_gs, int rows, int A, int C, int nmax, cudaStream_t s) {
  MFVAE_CHECK(C / 4 <= kFoldThreads && C % 4 == 0 && nmax <= 16 && w_ld % 4 == 0 && wcol0 % 4 == 0 && table_gs % 4 == 0,
              "act fold: act_features % 4 == 0, <= 1024 and n_act <= 16");
  // one wave: (row chunks x agents) CTAs <= 148 SMs x 2 resident CTAs of this register footprint (96 registers x 256 threads)
  const int chunks = std::max(1, std::min(rows, (kNumSMs * 2) / std::max(A, 1)));
  const int rows_per_cta = (rows + chunks - 1) / chunks;
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

// g_emb[a][i] += sum_n W0[a][n][i] * db0[a][n]        gW0[a][n][i] = db0[a][n] * emb[a][i]   (i < I)
// grid = (N / kEmbRows, A): a CTA covers kEmbRows weight rows n of one agent; threads = (I columns) x (row phases)
constexpr int kEmbRows = 16;
__global__ void __launch_bounds__(kFoldThreads) emb_grad_fold_kernel(const float* __restrict__ W0, float* __restrict__ gW0, int64_t w_ld,
                                                                     const float* __restrict__ db0, const float* __restrict__ emb,
                                                                     float* __restrict__ g_emb, int N, int I) {
  __shared__ float fold[kFoldThreads];
  const int a = blockIdx.y;
  const int i = threadIdx.x % I, rphase = threadIdx.x / I, nph = blockDim.x / I;
  const int n0 = blockIdx.x * kEmbRows, n1 = min(N, n0 + kEmbRows);
  float acc = 0.f;
  if (rphase < nph) {
    const float e = emb[static_cast<int64_t>(a) * I + i];
    for (int n = n0 + rphase; n < n1; n += nph) {
      const int64_t wi = (static_cast<int64_t>(a) * N + n) * w_ld + i;
      const float d = db0[a * N + n];
      acc = fmaf(__ldg(W0 + wi), d, acc);
      gW0[wi] = d * e;
    }
  }
  fold[threadIdx.x] = acc;
  __syncthreads();
  if (rphase == 0) {
    float t = fold[i];
    for (int r = 1; r < nph; ++r) t += fold[r * I + i];
    atomicAdd(g_emb + static_cast<int64_t>(a) * I + i, t);
  }
}

int launch_emb_grad_fold(const float* W0, float* gW0, int64_t w_ld, const float* db0, const float* emb, float* g_emb, int A, int N, int I,
                         cudaStream_t s) {
  MFVAE_CHECK(I <= kFoldThreads, "emb fold: idx_features <= 256");
  dim3 grid((N + kEmbRows - 1) / kEmbRows, A);
  emb_grad_fold_kernel<<<grid, kFoldThreads, 0, s>>>(W0, gW0, w_ld, db0, emb, g_emb, N, I);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mfvae
