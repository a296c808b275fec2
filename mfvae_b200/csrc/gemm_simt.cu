// gemm_simt.cu — FFMA GEMM with fp32 accumulation and arbitrary element strides.
//
// Role: the fp32 precision mode of the step (north_star: 1e-5 relative against torch_ver needs true fp32
// products, which the bf16 tensor path cannot give) and the bring-up comparator for the tcgen05 kernels
// (same GemmOp, same epilogues).  C[g](m,n) = epi( sum_k A[g](m,k) * B[g](n,k) ).
//
// 64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread; operands staged through shared memory as fp32
// with the thread->element mapping chosen so that global reads follow whichever stride is 1.
#include "kernels.h"

namespace mfvae {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__device__ __forceinline__ void load_tile(float (*sm)[TM + 4], const T* __restrict__ base, int64_t rs, int64_t cs,
                                          int row0, int nrows, int k0, int k1, int tid) {
  // tile is TM(rows) x TK(k) = 1024 elements, 4 per thread
  if (cs == 1) {           // K-major: consecutive threads walk k
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (tid >> 4) + 16 * i, k = tid & 15;
      float v = 0.f;
      if (row0 + r < nrows && k0 + k < k1) v = to_f<T>(base[static_cast<int64_t>(row0 + r) * rs + (k0 + k)]);
      sm[k][r] = v;
    }
  } else {                 // MN-major (or generic): consecutive threads walk rows
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tid & 63, k = (tid >> 6) + 4 * i;
      float v = 0.f;
      if (row0 + r < nrows && k0 + k < k1) v = to_f<T>(base[static_cast<int64_t>(row0 + r) * rs + static_cast<int64_t>(k0 + k) * cs]);
      sm[k][r] = v;
    }
  }
}

template <typename T, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmOp op) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int g = blockIdx.z / op.split_k;
  const int ks = blockIdx.z - g * op.split_k;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;

  // k range of this split, in whole TK blocks
  const int kblocks = (op.K + TK - 1) / TK;
  const int per = (kblocks + op.split_k - 1) / op.split_k;
  const int kb0 = ks * per, kb1 = min(kblocks, kb0 + per);
  const int kbeg = kb0 * TK, kend = min(op.K, kb1 * TK);

  const T* A = static_cast<const T*>(op.A) + g * op.a_gs;
  const T* B = static_cast<const T*>(op.B) + g * op.b_gs;
  const T* B2 = op.B2 ? static_cast<const T*>(op.B2) + g * op.b2_gs : nullptr;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    load_tile<T>(As, A, op.a_rs, op.a_cs, m0, op.M, k0, kend, tid);
    if (B2 && k0 >= op.k_split) load_tile<T>(Bs, B2, op.b2_rs, 1, n0, op.N, k0 - op.k_split, min(kend - op.k_split, op.k2), tid);
    else if (B2)                load_tile<T>(Bs, B, op.b_rs, op.b_cs, n0, op.N, k0, min(kend, op.k1), tid);
    else                        load_tile<T>(Bs, B, op.b_rs, op.b_cs, n0, op.N, k0, kend, tid);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  TC* C = static_cast<TC*>(op.C) + g * op.c_gs;
  const float* bias = op.bias ? op.bias + g * op.bias_gs : nullptr;
  const T* aux = op.aux ? static_cast<const T*>(op.aux) + g * op.aux_gs : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= op.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= op.N) continue;
      float v = acc[i][j];
      if (op.epi == kEpiBias || op.epi == kEpiBiasRelu) v += bias[n];
      if (op.epi == kEpiBiasRelu) v = fmaxf(v, 0.f);
      if (op.epi == kEpiReluMask) v = (to_f<T>(aux[static_cast<int64_t>(m) * op.aux_ld + n]) > 0.f) ? v : 0.f;
      TC* dst = C + static_cast<int64_t>(m) * op.c_ld + n;
      if (op.epi == kEpiAccum) atomicAdd(reinterpret_cast<float*>(dst), v);
      else *dst = from_f<TC>(v);
    }
  }
}

int gemm_simt(const GemmOp& op, cudaStream_t s) {
  MFVAE_CHECK(op.M > 0 && op.N > 0 && op.K > 0 && op.G > 0, "gemm_simt: empty problem");
  MFVAE_CHECK(op.split_k >= 1, "gemm_simt: split_k >= 1");
  MFVAE_CHECK(!op.B2 || (op.k_split % 64 == 0 && op.k_split > 0 && op.k_split < op.K && op.b_cs == 1), "gemm_simt: second B segment needs a K-major B and k_split % 64 == 0");
  MFVAE_CHECK(!op.B2 || (op.k1 > 0 && op.k1 <= op.k_split && op.K == op.k_split + op.k2), "gemm_simt: second B segment: K = k_split + k2, k1 <= k_split");
  MFVAE_CHECK(op.split_k == 1 || op.epi == kEpiAccum, "gemm_simt: split-K needs the accumulate epilogue");
  MFVAE_CHECK(op.epi != kEpiAccum || op.c_dtype == kF32, "gemm_simt: accumulate epilogue needs fp32 C");
  MFVAE_CHECK(op.epi != kEpiReluMask || op.aux, "gemm_simt: relu-mask epilogue needs aux");
  MFVAE_CHECK((op.epi != kEpiBias && op.epi != kEpiBiasRelu) || op.bias, "gemm_simt: bias epilogue needs bias");
  dim3 grid((op.N + TN - 1) / TN, (op.M + TM - 1) / TM, op.G * op.split_k);
  MFVAE_CHECK(grid.y <= 65535 && grid.z <= 65535, "gemm_simt: grid too large");
  if (op.dtype == kBF16) {
    if (op.c_dtype == kBF16) gemm_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(op);
    else                     gemm_simt_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>(op);
  } else {
    MFVAE_CHECK(op.c_dtype == kF32, "gemm_simt: fp32 operands produce fp32 C");
    gemm_simt_kernel<float, float><<<grid, 256, 0, s>>>(op);
  }
  MFVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mfvae
