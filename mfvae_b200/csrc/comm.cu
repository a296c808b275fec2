// comm.cu — the data-parallel exchange step as hand-written kernels over NVLink / NVSwitch peer memory (sm_100a).
//
// The reference has no collective at all (single process, torch_ver/main.py:34); data parallelism is this library's
// addition: every rank holds a replica, gradients of the optimised prefix are summed across ranks once
// per step.  Instead of an NCCL all-reduce (ring kernels with 16-32 CTAs that share SMs and HBM with backward for
// ~300 us per step at this payload), each gradient bucket is reduced by ONE small two-shot kernel on the caller's
// communication stream:
//
//   pack        this rank's fp32 gradients of the bucket -> the rank's window of a symmetric buffer (bf16 or fp32 payload)
//   all-reduce  barrier; rank r reduces slice r of the bucket across all windows and writes the sum back into every window;
//               barrier.  With a multicast mapping (NVSwitch, NVLS) that is `multimem.ld_reduce` (the switch adds, fp32
//               accumulation even for a bf16 payload) + `multimem.st` (the switch replicates): per rank 1/N of the bucket in
//               and 1/N out.  Without one, plain peer loads in fixed rank order and N peer stores.
//   consume     Adam reads the reduced gradient straight out of the window (adam_payload_kernel) and leaves the fp32 copy in
//               the gradient arena, so `.grad` still holds the reduced gradient.
//
// Cross-rank synchronisation is a flag exchange through per-rank signal pads (one 32-bit slot per (block, peer)): put =
// CAS 0 -> 1 on the peer's pad with release.sys, wait = CAS 1 -> 0 on the own pad with acquire.sys -- every barrier leaves
// the slots at 0, so launches need no epoch counter.  All ranks must launch the same sequence with the same grid sizes.
//
// The symmetric buffer, the peer / multicast pointers and the signal pads are handed in by the host (mfvae_comm_bind):
// the Python host gets them from torch.distributed._symmetric_memory (CUDA VMM + fabric handles); a C host would make
// the same cuMem* / cuMulticast* calls.  PyTorch is plumbing here, not the product.
#include <algorithm>

#include "kernels.h"

namespace mfvae {

constexpr int kCommThreads = 256;          // small CTAs: they have to find room beside the persistent GEMM CTAs of backward
constexpr int kCommMaxBlocks = 148;
constexpr int kCommSlotBase = 64;          // signal-pad slots [0, 64) are left to the host framework's own barriers

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// block b of every rank meets block b of every other rank
__device__ __forceinline__ void rank_barrier(uint32_t* const* pads, int rank, int world) {
  __syncthreads();
  if (threadIdx.x < world) {
    __threadfence_system();
    uint32_t* put = pads[threadIdx.x] + kCommSlotBase + blockIdx.x * world + rank;
    uint32_t spins = 0;
    while (cas_release_sys(put, 0u, 1u) != 0u) { if (++spins > (1u << 28)) __trap(); }
    uint32_t* get = pads[rank] + kCommSlotBase + blockIdx.x * world + threadIdx.x;
    spins = 0;
    while (cas_acquire_sys(get, 1u, 0u) != 1u) { if (++spins > (1u << 28)) __trap(); }
  }
  __syncthreads();
}

// 16-byte vectors: 4 fp32 or 8 bf16
template <typename TP> __device__ __forceinline__ uint4 mc_ld_reduce(const void* p);
template <> __device__ __forceinline__ uint4 mc_ld_reduce<float>(const void* p) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
template <> __device__ __forceinline__ uint4 mc_ld_reduce<__nv_bfloat16>(const void* p) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
template <typename TP> __device__ __forceinline__ void mc_st(void* p, uint4 v);
template <> __device__ __forceinline__ void mc_st<float>(void* p, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <> __device__ __forceinline__ void mc_st<__nv_bfloat16>(void* p, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint4 ld_sys_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// peer path: sum of the N windows' vectors in rank order 0..N-1 (the same order on every rank: identical bits everywhere)
template <typename TP> __device__ __forceinline__ uint4 peer_sum(void* const* peers, int world, int64_t byte_off);
template <> __device__ __forceinline__ uint4 peer_sum<float>(void* const* peers, int world, int64_t byte_off) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < world; ++r) {
    const uint4 v = ld_sys_v4(static_cast<const char*>(peers[r]) + byte_off);
    acc.x += __uint_as_float(v.x); acc.y += __uint_as_float(v.y); acc.z += __uint_as_float(v.z); acc.w += __uint_as_float(v.w);
  }
  return make_uint4(__float_as_uint(acc.x), __float_as_uint(acc.y), __float_as_uint(acc.z), __float_as_uint(acc.w));
}
template <> __device__ __forceinline__ uint4 peer_sum<__nv_bfloat16>(void* const* peers, int world, int64_t byte_off) {
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < world; ++r) {
    const uint4 v = ld_sys_v4(static_cast<const char*>(peers[r]) + byte_off);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
      acc[2 * k] += f.x; acc[2 * k + 1] += f.y;
    }
  }
  uint4 o;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(acc[0], acc[1]); o.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(acc[2], acc[3]); o.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(acc[4], acc[5]); o.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(acc[6], acc[7]); o.w = *reinterpret_cast<uint32_t*>(&t);
  return o;
}

struct CommArgs {
  int rank, world;
  void* const* peers;          // device array [world]: base of every rank's window (this rank's own included)
  char* mc;                    // multicast base of the windows, or nullptr
  uint32_t* const* pads;       // device array [world]: signal pads
};

// two-shot all-reduce of window bytes [byte0, byte1) (multiples of 16), in place in every window.  With `src` the kernel first
// packs this rank's fp32 gradients of the range into its window (no separate pack launch waiting for CTA slots).
// (Capping this kernel at 40 registers so that it fits beside three resident GEMM CTAs of backward's small layers was measured:
// no change at 2 GPUs -- the data-parallel overhead is not CTA eviction.)
template <typename TP, bool MC>
__global__ void __launch_bounds__(kCommThreads) allreduce_kernel(CommArgs c, int64_t byte0, int64_t byte1, const float* __restrict__ src,
                                                                 int barriers) {
  if (src) {
    TP* w = reinterpret_cast<TP*>(static_cast<char*>(c.peers[c.rank]) + byte0);
    const int64_t n4 = (byte1 - byte0) / (4 * static_cast<int64_t>(sizeof(TP)));
    constexpr int UP = 4;
    const int64_t pstride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += UP * pstride) {
      float4 v[UP];
#pragma unroll
      for (int u = 0; u < UP; ++u) if (i + u * pstride < n4) v[u] = ldg_stream4(src + (i + u * pstride) * 4);
#pragma unroll
      for (int u = 0; u < UP; ++u) if (i + u * pstride < n4) store4<TP>(w + (i + u * pstride) * 4, v[u]);
    }
  }
  if (barriers & 1) rank_barrier(c.pads, c.rank, c.world);     // every rank's pack of this range has landed
  const int64_t nvec = (byte1 - byte0) >> 4;
  const int64_t per = (nvec + c.world - 1) / c.world;
  const int64_t v0 = min(nvec, per * c.rank), v1 = min(nvec, v0 + per);
  constexpr int U = 8;                                         // 16-byte vectors in flight per thread
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = v0 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < v1; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = i + u * stride;
      if (j < v1) {
        const int64_t off = byte0 + (j << 4);
        v[u] = MC ? mc_ld_reduce<TP>(c.mc + off) : peer_sum<TP>(c.peers, c.world, off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = i + u * stride;
      if (j < v1) {
        const int64_t off = byte0 + (j << 4);
        if (MC) {
          mc_st<TP>(c.mc + off, v[u]);
        } else {
          for (int r = 0; r < c.world; ++r) *reinterpret_cast<uint4*>(static_cast<char*>(c.peers[r]) + off) = v[u];
        }
      }
    }
  }
  if (barriers & 2) rank_barrier(c.pads, c.rank, c.world);     // every slice has been written into every window
}

// the cross-rank meeting as its own one-warp kernel: between a wide pack and a wide reduce it keeps the waiting (rank skew:
// tens of microseconds) out of kernels whose resident CTAs would take slots from backward's GEMMs while they spin
__global__ void __launch_bounds__(32) rendezvous_kernel(CommArgs c) { rank_barrier(c.pads, c.rank, c.world); }

// one-shot all-reduce of a few fp32 scalars (the loss partials): every rank reduces all of them into its own `out`
template <bool MC>
__global__ void __launch_bounds__(32) allreduce_scalars_kernel(CommArgs c, int64_t byte0, int n, float* out) {
  rank_barrier(c.pads, c.rank, c.world);
  if (threadIdx.x * 4 < n) {
    const int64_t off = byte0 + threadIdx.x * 16;
    const uint4 v = MC ? mc_ld_reduce<float>(c.mc + off) : peer_sum<float>(c.peers, c.world, off);
    const float f[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
    for (int k = 0; k < 4 && threadIdx.x * 4 + k < n; ++k) out[threadIdx.x * 4 + k] = f[k];
  }
  rank_barrier(c.pads, c.rank, c.world);                       // nobody overwrites its window while a peer still reads it
}

template <typename TP>
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ g, TP* __restrict__ w, int64_t n4) {
  constexpr int U = 4;                        // 4 x 16-byte loads in flight per thread: few CTAs have to cover the HBM latency
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * stride < n4) v[u] = ldg_stream4(g + (i + u * stride) * 4);
#pragma unroll
    for (int u = 0; u < U; ++u) if (i + u * stride < n4) store4<TP>(w + (i + u * stride) * 4, v[u]);
  }
}
template <typename TP>
__global__ void __launch_bounds__(256) unpack_kernel(const TP* __restrict__ w, float* __restrict__ g, int64_t n4) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    *reinterpret_cast<float4*>(g + i * 4) = load4<TP>(w + i * 4);
}

// Adam with the gradient read from the window (bf16 or fp32) and written back as fp32 into the gradient arena
template <typename TP>
__global__ void __launch_bounds__(256) adam_payload_kernel(float* __restrict__ p, const TP* __restrict__ gw, float* __restrict__ g,
                                                           float* __restrict__ m, float* __restrict__ v, __nv_bfloat16* __restrict__ shadow,
                                                           int64_t n4, float one_minus_b1, float b2, float one_minus_b2, float step_size,
                                                           float bc2_sqrt, float eps) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = load4<TP>(gw + i * 4);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define MFVAE_ADAM_1(c)                                           \
    mm.c = mm.c + one_minus_b1 * (gg.c - mm.c);                   \
    vv.c = vv.c * b2 + one_minus_b2 * gg.c * gg.c;                \
    pp.c = pp.c - step_size * (mm.c / (sqrtf(vv.c) / bc2_sqrt + eps));
    MFVAE_ADAM_1(x) MFVAE_ADAM_1(y) MFVAE_ADAM_1(z) MFVAE_ADAM_1(w)
#undef MFVAE_ADAM_1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(g)[i] = gg;
    if (shadow) store4<__nv_bfloat16>(shadow + i * 4, pp);
  }
}

// Tiny ranges (the id-embedding bucket: a few thousand elements, final only at the very end of backward): ONE single-CTA kernel
// packs the range as fp32 into the window, meets the peers, reduces the whole range itself (one-shot: every rank reads every
// window, rank order) and applies Adam -- three launches and a stream hop less on the step's exposed tail.
__global__ void __launch_bounds__(1024) small_allreduce_adam_kernel(CommArgs c, int64_t byte0, int n4, float* __restrict__ p, float* __restrict__ g,
                                                                    float* __restrict__ m, float* __restrict__ v, __nv_bfloat16* __restrict__ shadow,
                                                                    int do_adam, float one_minus_b1, float b2, float one_minus_b2, float step_size,
                                                                    float bc2_sqrt, float eps) {
  float4* mine = reinterpret_cast<float4*>(static_cast<char*>(c.peers[c.rank]) + byte0);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) mine[i] = reinterpret_cast<const float4*>(g)[i];
  rank_barrier(c.pads, c.rank, c.world);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const uint4 r = peer_sum<float>(c.peers, c.world, byte0 + static_cast<int64_t>(i) * 16);
    const float4 gg = make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
    reinterpret_cast<float4*>(g)[i] = gg;
    if (do_adam) {
      float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
#define MFVAE_ADAM_1(c_)                                            \
      mm.c_ = mm.c_ + one_minus_b1 * (gg.c_ - mm.c_);               \
      vv.c_ = vv.c_ * b2 + one_minus_b2 * gg.c_ * gg.c_;            \
      pp.c_ = pp.c_ - step_size * (mm.c_ / (sqrtf(vv.c_) / bc2_sqrt + eps));
      MFVAE_ADAM_1(x) MFVAE_ADAM_1(y) MFVAE_ADAM_1(z) MFVAE_ADAM_1(w)
#undef MFVAE_ADAM_1
      reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
      if (shadow) store4<__nv_bfloat16>(shadow + i * 4, pp);
    }
  }
  rank_barrier(c.pads, c.rank, c.world);              // nobody re-packs its window while a peer still reads it
}

// `n` fp32 elements of the arena at `elem0` through the window's small-range scratch (fp32, at byte c.small_off)
int comm_small_allreduce_adam(const CommCtx& c, int64_t n, float* p, float* g, float* m, float* v, __nv_bfloat16* shadow, int do_adam,
                              float lr, float b1, float b2, float eps, int64_t t, cudaStream_t s) {
  MFVAE_CHECK(n % 4 == 0 && n * 4 <= c.small_bytes, "comm: small range does not fit the scratch region");
  const double bc1 = 1.0 - pow(static_cast<double>(b1), static_cast<double>(std::max<int64_t>(t, 1)));
  const double bc2 = 1.0 - pow(static_cast<double>(b2), static_cast<double>(std::max<int64_t>(t, 1)));
  CommArgs a{c.rank, c.world, c.d_peers, static_cast<char*>(c.mc), c.d_pads};
  small_allreduce_adam_kernel<<<1, 1024, 0, s>>>(a, c.small_off, static_cast<int>(n / 4), p, g, m, v, shadow, do_adam, 1.f - b1, b2, 1.f - b2,
                                                 static_cast<float>(static_cast<double>(lr) / bc1), static_cast<float>(sqrt(bc2)), eps);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

static inline int grid1d(int64_t items, int threads, int per_sm) {
  const int64_t blocks = (items + threads - 1) / threads;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(blocks, static_cast<int64_t>(kNumSMs) * per_sm)));
}

int comm_pack(const float* g, void* window, int dtype, int64_t begin, int64_t end, cudaStream_t s) {
  MFVAE_CHECK(begin % 8 == 0 && end % 8 == 0 && end >= begin, "comm: ranges must be 8-element aligned");
  if (end == begin) return 0;
  const int64_t n4 = (end - begin) / 4;
  if (dtype == kBF16) pack_kernel<__nv_bfloat16><<<grid1d(n4, 256, 2), 256, 0, s>>>(g + begin, static_cast<__nv_bfloat16*>(window) + begin, n4);
  else                pack_kernel<float><<<grid1d(n4, 256, 2), 256, 0, s>>>(g + begin, static_cast<float*>(window) + begin, n4);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

int comm_unpack(const void* window, int dtype, float* g, int64_t begin, int64_t end, cudaStream_t s) {
  if (end <= begin) return 0;
  const int64_t n4 = (end - begin) / 4;
  if (dtype == kBF16) unpack_kernel<__nv_bfloat16><<<grid1d(n4, 256, 2), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(window) + begin, g + begin, n4);
  else                unpack_kernel<float><<<grid1d(n4, 256, 2), 256, 0, s>>>(static_cast<const float*>(window) + begin, g + begin, n4);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

int comm_allreduce(const CommCtx& c, int64_t begin, int64_t end, cudaStream_t s, const float* pack_from) {
  if (c.split_sync && pack_from) {
    // pack (wide, no waiting) -> rendezvous (one warp) -> reduce (wide, no waiting) -> rendezvous (one warp)
    MFVAE_TRY(comm_pack(pack_from, c.local, c.dtype, begin, end, s));
    CommArgs a0{c.rank, c.world, c.d_peers, static_cast<char*>(c.mc), c.d_pads};
    rendezvous_kernel<<<1, 32, 0, s>>>(a0);
    MFVAE_LAUNCH_CHECK();
    pack_from = nullptr;
  }
  const int barriers = c.split_sync ? 0 : 3;
  MFVAE_CHECK(begin % 8 == 0 && end % 8 == 0 && end >= begin, "comm: ranges must be 8-element aligned");
  if (end == begin) return 0;
  const int64_t es = (c.dtype == kBF16) ? 2 : 4;
  const int64_t b0 = begin * es, b1 = end * es;
  CommArgs a{c.rank, c.world, c.d_peers, static_cast<char*>(c.mc), c.d_pads};
  const float* src = pack_from ? pack_from + begin : nullptr;
  // grid: enough CTAs for the pack of the whole range (every rank packs all of it), capped
  const int64_t nvec = (b1 - b0) / 16;
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(c.max_blocks, (nvec + kCommThreads * 8 - 1) / (kCommThreads * 8))));
  if (c.dtype == kBF16) {
    if (c.mc) allreduce_kernel<__nv_bfloat16, true><<<blocks, kCommThreads, 0, s>>>(a, b0, b1, src, barriers);
    else      allreduce_kernel<__nv_bfloat16, false><<<blocks, kCommThreads, 0, s>>>(a, b0, b1, src, barriers);
  } else {
    if (c.mc) allreduce_kernel<float, true><<<blocks, kCommThreads, 0, s>>>(a, b0, b1, src, barriers);
    else      allreduce_kernel<float, false><<<blocks, kCommThreads, 0, s>>>(a, b0, b1, src, barriers);
  }
  MFVAE_LAUNCH_CHECK();
  if (c.split_sync) { rendezvous_kernel<<<1, 32, 0, s>>>(a); MFVAE_LAUNCH_CHECK(); }
  return 0;
}

// loss scalars: `n` floats copied into this rank's window at byte `scalar_off`, summed over ranks into `out`
int comm_allreduce_scalars(const CommCtx& c, const float* src, int n, float* out, cudaStream_t s) {
  MFVAE_CHECK(n >= 1 && n <= 64, "comm: at most 64 scalars");
  MFVAE_CUDA(cudaMemcpyAsync(static_cast<char*>(c.local) + c.scalar_off, src, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CommArgs a{c.rank, c.world, c.d_peers, static_cast<char*>(c.mc), c.d_pads};
  if (c.mc) allreduce_scalars_kernel<true><<<1, 32, 0, s>>>(a, c.scalar_off, n, out);
  else      allreduce_scalars_kernel<false><<<1, 32, 0, s>>>(a, c.scalar_off, n, out);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

int launch_adam_payload(float* p, const void* gw, int dtype, float* g, float* m, float* v, __nv_bfloat16* shadow, int64_t n,
                        float lr, float b1, float b2, float eps, int64_t t, cudaStream_t s) {
  MFVAE_CHECK(n % 4 == 0, "adam: element count must be a multiple of 4");
  MFVAE_CHECK(t >= 1, "adam: step count starts at 1");
  const double bc1 = 1.0 - pow(static_cast<double>(b1), static_cast<double>(t));
  const double bc2 = 1.0 - pow(static_cast<double>(b2), static_cast<double>(t));
  const float step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  const float bc2_sqrt = static_cast<float>(sqrt(bc2));
  const int grid = grid1d(n / 4, 256, 4);       // beside the wgrad tail of backward: leave thread slots to it
  if (dtype == kBF16)
    adam_payload_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p, static_cast<const __nv_bfloat16*>(gw), g, m, v, shadow, n / 4, 1.f - b1, b2, 1.f - b2, step_size, bc2_sqrt, eps);
  else
    adam_payload_kernel<float><<<grid, 256, 0, s>>>(p, static_cast<const float*>(gw), g, m, v, shadow, n / 4, 1.f - b1, b2, 1.f - b2, step_size, bc2_sqrt, eps);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mfvae
