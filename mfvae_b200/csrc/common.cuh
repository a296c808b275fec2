// common.cuh — shared device/host helpers for the mfvae_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mfvae.h"

namespace mfvae {

// ---------------------------------------------------------------------------------------------
// error plumbing: nothing throws across the C ABI
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(const char* file, int line, const std::string& msg);

#define MFVAE_FAIL(msg) return ::mfvae::fail(__FILE__, __LINE__, (msg))
#define MFVAE_CHECK(cond, msg) do { if (!(cond)) MFVAE_FAIL(msg); } while (0)
#define MFVAE_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
    MFVAE_FAIL(std::string(#expr) + ": " + cudaGetErrorString(e__)); } while (0)
#define MFVAE_TRY(expr) do { int r__ = (expr); if (r__ != 0) return r__; } while (0)
extern unsigned long long g_launch_count;      // kernels launched by this library (bench.py's gpu_launches)
#define MFVAE_LAUNCH_CHECK() do { ++::mfvae::g_launch_count; MFVAE_CUDA(cudaGetLastError()); } while (0)

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// element type helpers (activations are float or bf16)
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T x);
template <> __device__ __forceinline__ float to_f<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// 4 consecutive elements <-> float4 (16-byte fp32 or 8-byte bf16 accesses)
template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// streaming (read-once) 128-bit load that does not allocate in L1
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) + Box-Muller.  Counter layout (must match
// oracle/mavae_oracle.py::philox_normal): ctr = (sample_lo, sample_hi, column/4, step), key = seed.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0; k.y += W1;
  }
  return c;
}

__device__ __forceinline__ float u01_open(uint32_t x) {   // ((x >> 9) + 0.5) * 2^-23, exact in fp32
  return (static_cast<float>(x >> 9) + 0.5f) * (1.0f / 8388608.0f);
}

// 4 standard normals for (global sample, column quad q) of step `step`
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t step, uint64_t sample, uint32_t q) {
  uint4 c = make_uint4(static_cast<uint32_t>(sample), static_cast<uint32_t>(sample >> 32), q,
                       static_cast<uint32_t>(step));
  uint2 k = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  uint4 r = philox4x32_10(c, k);
  float r0 = sqrtf(-2.0f * logf(u01_open(r.x)));
  float r1 = sqrtf(-2.0f * logf(u01_open(r.z)));
  float s0, c0, s1, c1;
  sincospif(2.0f * u01_open(r.y), &s0, &c0);
  sincospif(2.0f * u01_open(r.w), &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0.  `red` must hold >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = (lane < (blockDim.x + 31) / 32) ? red[lane] : 0.f;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;
}

constexpr int kMaxPartials = 2048;          // scratch[0..2047] partials, scratch[4095] ticket
constexpr int kTicketSlot = 4095;


// Sum per-CTA partials in a fixed order once every CTA has published; the last CTA to take a ticket
// does it, then re-arms the ticket for the next launch.
__device__ __forceinline__ void finish_scalar(float block_total, float* scratch, float scale, float* out,
                                              float* red) {
  __shared__ bool is_last;
  const unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) {
    scratch[bid] = block_total;
    __threadfence();
    unsigned int ticket = atomicAdd(reinterpret_cast<unsigned int*>(scratch + kTicketSlot), 1u);
    is_last = (ticket == nblocks - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float v = 0.f;
    for (int i = threadIdx.x; i < static_cast<int>(nblocks); i += blockDim.x) v += __ldcg(scratch + i);
    v = block_sum(v, red);
    if (threadIdx.x == 0) {
      out[0] = v * scale;
      *reinterpret_cast<unsigned int*>(scratch + kTicketSlot) = 0u;
    }
  }
}

}  // namespace mfvae
