// gemm_tc.cu — bf16 tensor-core GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue.
//
//   C[g](m,n) = epi( sum_k A[g](m,k) * B[g](n,k) ),  fp32 accumulation in TMEM.
//
// One kernel template covers the three contractions of a Linear layer (reference nn.Linear at
// torch_ver/model.py:50,53,91,94 and its autograd backward):
//   forward  Y  = X  W^T      A = X  (K-major)   B = W (K-major)        epilogue: +bias [, relu]
//   dgrad    dX = dY W        A = dY (K-major)   B = W (MN-major)       epilogue: * (X > 0)
//   wgrad    dW = dY^T X      A = dY (MN-major)  B = X (MN-major)       epilogue: fp32 (+= with split-K)
// so no operand is ever transposed in HBM: the UMMA shared-memory descriptors take either major.
//
// CTA = 10 warps, persistent over a static round-robin tile schedule:
//   warp 0      TMA producer   (one lane): cp.async.bulk.tensor.3d into a kStages-deep 128B-swizzled ring
//   warp 1      MMA issuer     (one lane): tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 per instruction;
//                              tcgen05.commit releases ring slots and publishes finished accumulators
//   warps 2..9  epilogue       tcgen05.ld 32 lanes x 32 columns -> registers -> bias/relu/mask -> global
//                              (two warps per TMEM lane quarter, next chunk's TMEM load overlapped with the stores)
// TMEM holds two accumulator stages (2 x BN fp32 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "kernels.h"
#include "tc_ptx.cuh"

namespace mfvae {

// SMs left free for a concurrently running collective (data parallel): persistent GEMMs whose CTAs cannot all be resident
// at once run a second wave -- see mfvae_set_sm_reserve
int g_tc_sm_reserve = 0;

constexpr int BM = 128;
constexpr int BK = 64;                    // 64 bf16 = 128 bytes = one SWIZZLE_128B row
// Two launch shapes.  CPS = CTAs per SM:
//   CPS = 1  deep ring (up to 8 stages), 8 epilogue warps (two per TMEM lane quarter)  -> K-long, tensor-bound GEMMs
//   CPS = 3  3-stage ring of 24 KB (BN = 64), 4 epilogue warps, <= 112 registers        -> K <= 256 GEMMs, whose tiles
//            are 1-4 MMAs followed by an epilogue: three co-resident CTAs hide each other's TMEM / store latency
template <int CPS> struct TcShape {
  static constexpr int kEpiWarps = (CPS == 1) ? 8 : 4;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
};

struct TcParams {
  int G, M, N, K;
  int m_tiles, n_tiles, k_blocks, splits, kb_per_split;
  long long total_work;
  void* C; long long c_gs, c_ld; int c_dtype;
  const float* bias; long long bias_gs;
  int epi;
  const __nv_bfloat16* aux; long long aux_gs, aux_ld;
  int accumulate_atomic;                  // split-K: red.global.add.f32
  int c_v8, aux_v8;                       // C rows / aux rows are 32-byte aligned: 256-bit accesses, one sector per lane
  int kb_switch;                          // > 0: k-blocks >= kb_switch read the B operand through map_b2 (second K segment)
  // kEpiLossGrad: C = d loss / d (A B^T + bias) against `tgt`, loss value accumulated per epilogue warp
  const float* tgt; long long tgt_ld; float grad_scale; int huber; float* loss_partials;
};

template <int BN, int CPS> struct TcCfg {
  static constexpr int kABytes = BM * BK * 2;                  // 16 KB
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesDeep = (200 * 1024 / kStageBytes) > 8 ? 8 : (200 * 1024 / kStageBytes);
  static constexpr int kStages = (CPS == 1) ? kStagesDeep : 3;
  static constexpr int kTmemCols = 2 * BN;                     // 128 / 256 / 512: a power of two >= 32
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ------------------------------------------------------------------------------------------------
// epilogue for one 32-column chunk held by one thread (= one output row)
// ------------------------------------------------------------------------------------------------
// Every lane of the warp must call this (it shuffles); rows >= M only skip their loads / stores.
__device__ __forceinline__ void epilogue_chunk(const TcParams& p, int g, int row, int col0, const uint32_t (&v)[32], float& loss_acc) {
  const bool row_ok = row < p.M;
  const int ncols = min(32, p.N - col0);
  const int lane = threadIdx.x & 31;
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  if (p.epi == kEpiReluMask && p.c_dtype == kBF16 && ncols == 32 && p.c_v8 && p.aux_v8) {
    // dgrad fast path: dX = (dY W) * (X > 0) on packed bf16 -- the ReLU mask is one HSETP2-mask + one AND per element pair
    // (rounding to bf16 and zeroing commute), rows move as 256-bit sectors both ways
    if (row_ok) {
      const __nv_bfloat16* a = p.aux + g * p.aux_gs + static_cast<long long>(row) * p.aux_ld + col0;
      uint32_t raw[16], pk[16];
      ld_global_nc_256(a, raw);
      ld_global_nc_256(a + 16, raw + 8);
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1], false) & __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&raw[j]), zero2);
      __nv_bfloat16* c = static_cast<__nv_bfloat16*>(p.C) + g * p.c_gs + static_cast<long long>(row) * p.c_ld + col0;
      st_global_256(c, pk);
      st_global_256(c + 16, pk + 8);
    }
    return;
  }
  if (p.epi == kEpiBias || p.epi == kEpiBiasRelu || p.epi == kEpiLossGrad) {
    const float* b = p.bias + g * p.bias_gs + col0;
    if (ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {       // packed fp32x2 adds (sm_100): half the issue slots of 32 scalar FADDs
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b + j));
        const float2 s0 = __fadd2_rn(make_float2(f[j], f[j + 1]), make_float2(bb.x, bb.y));
        const float2 s1 = __fadd2_rn(make_float2(f[j + 2], f[j + 3]), make_float2(bb.z, bb.w));
        f[j] = s0.x; f[j + 1] = s0.y; f[j + 2] = s1.x; f[j + 3] = s1.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) f[j] += b[j];
    }
  } else if (p.epi == kEpiReluMask && row_ok) {
    const __nv_bfloat16* a = p.aux + g * p.aux_gs + static_cast<long long>(row) * p.aux_ld + col0;
    if (ncols == 32 && p.aux_v8) {
      uint32_t raw[16];
      ld_global_nc_256(a, raw);
      ld_global_nc_256(a + 16, raw + 8);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float2 m = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[q]));
        if (!(m.x > 0.f)) f[2 * q] = 0.f;
        if (!(m.y > 0.f)) f[2 * q + 1] = 0.f;
      }
    } else if (ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(a + j));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 m = __bfloat1622float2(h[q]);
          if (!(m.x > 0.f)) f[j + 2 * q] = 0.f;
          if (!(m.y > 0.f)) f[j + 2 * q + 1] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols && !(__bfloat162float(a[j]) > 0.f)) f[j] = 0.f;
    }
  }
  if (p.epi == kEpiLossGrad && row_ok) {
    // reconstruction loss fused into the output layer (reference model.py:25-28): the reconstruction never goes to HBM,
    // the epilogue reads the target, writes d loss / d recon and keeps the loss value
    const float* t = p.tgt + static_cast<long long>(row) * p.tgt_ld + col0;
    float tv[32];
    if (ncols == 32 && (reinterpret_cast<uintptr_t>(t) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 t4 = ldg_stream4(t + j);
        tv[j] = t4.x; tv[j + 1] = t4.y; tv[j + 2] = t4.z; tv[j + 3] = t4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) tv[j] = (j < ncols) ? __ldg(t + j) : 0.f;
    }
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float d = f[j] - tv[j];
      float val, gr;
      if (p.huber) {
        const float ad = fabsf(d);
        val = ad < 1.f ? 0.5f * d * d : ad - 0.5f;
        gr = fminf(fmaxf(d, -1.f), 1.f) * p.grad_scale;
      } else {
        val = d * d;
        gr = 2.f * d * p.grad_scale;
      }
      if (j < ncols) acc += val;
      f[j] = gr;
    }
    loss_acc += acc;
  }
  const bool relu = (p.epi == kEpiBiasRelu);
  if (p.c_dtype == kBF16) {
    __nv_bfloat16* cbase = static_cast<__nv_bfloat16*>(p.C) + g * p.c_gs + col0;
    if (ncols == 32 && p.c_v8) {
      // 32-byte-aligned rows: each lane writes its own 64 bytes as two 256-bit stores = two whole sectors, no exchange
      if (row_ok) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1], relu);
        __nv_bfloat16* c = cbase + static_cast<long long>(row) * p.c_ld;
        st_global_256(c, pk);
        st_global_256(c + 16, pk + 8);
      }
    } else if (ncols == 32) {
      // A lane holds 64 contiguous bytes of its row = two 32-byte sectors, but one store instruction moves 16 bytes per
      // lane.  Lanes (2i, 2i+1) trade halves so that every instruction writes whole sectors: {A0|A1}, {A2|A3} of the
      // even row, then {B0|B1}, {B2|B3} of the odd row -- half as many L2 write sectors as 32 row-strided 16-byte pieces.
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1], relu);
      const bool odd = lane & 1;
      uint32_t rc[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rc[i] = __shfl_xor_sync(0xffffffffu, odd ? pk[i] : pk[4 + i], 1);
        rc[4 + i] = __shfl_xor_sync(0xffffffffu, odd ? pk[8 + i] : pk[12 + i], 1);
      }
      const int row_e = row & ~1;
      __nv_bfloat16* ce = cbase + static_cast<long long>(row_e) * p.c_ld + (odd ? 8 : 0);
      __nv_bfloat16* co = ce + p.c_ld;
      if (row_e < p.M) {
        *reinterpret_cast<uint4*>(ce) = odd ? make_uint4(rc[0], rc[1], rc[2], rc[3]) : make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(ce + 16) = odd ? make_uint4(rc[4], rc[5], rc[6], rc[7]) : make_uint4(pk[8], pk[9], pk[10], pk[11]);
      }
      if (row_e + 1 < p.M) {
        *reinterpret_cast<uint4*>(co) = odd ? make_uint4(pk[4], pk[5], pk[6], pk[7]) : make_uint4(rc[0], rc[1], rc[2], rc[3]);
        *reinterpret_cast<uint4*>(co + 16) = odd ? make_uint4(pk[12], pk[13], pk[14], pk[15]) : make_uint4(rc[4], rc[5], rc[6], rc[7]);
      }
    } else if (row_ok) {
      __nv_bfloat16* c = cbase + static_cast<long long>(row) * p.c_ld;
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) c[j] = __float2bfloat16_rn(relu ? fmaxf(f[j], 0.f) : f[j]);
    }
  } else if (row_ok) {
    if (relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    float* c = static_cast<float*>(p.C) + g * p.c_gs + static_cast<long long>(row) * p.c_ld + col0;
    if (p.accumulate_atomic) {
      if (ncols == 32) {      // 128-bit vector reductions: 4x fewer L2 atomic operations than scalar red.f32
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c + j), "f"(f[j]), "f"(f[j + 1]), "f"(f[j + 2]), "f"(f[j + 3]) : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < ncols) atomicAdd(c + j, f[j]);
      }
    } else if (ncols == 32 && p.c_v8) {
      uint32_t u[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(f[j]);
#pragma unroll
      for (int j = 0; j < 32; j += 8) st_global_256(c + j, u + j);
    } else if (ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(c + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) c[j] = f[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, int CPS>
__global__ void __launch_bounds__(TcShape<CPS>::kThreads, CPS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_b2, const TcParams p) {
  using Cfg = TcCfg<BN, CPS>;
  constexpr int kEpiWarps = TcShape<CPS>::kEpiWarps;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;                                   // [kStages][16 KB]
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;          // [kStages][BN*128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                                    // TMA -> MMA
  uint64_t* empty = bars + kStages;                         // MMA -> TMA
  uint64_t* acc_full = bars + 2 * kStages;                  // MMA -> epilogue   [2]
  uint64_t* acc_empty = bars + 2 * kStages + 2;             // epilogue -> MMA   [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // programmatic dependent launch: let the next kernel of the stream begin its own prologue as CTAs of this one retire
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a); prefetch_tmap(&map_b);
    if (p.kb_switch > 0) prefetch_tmap(&map_b2);
    for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of the previous kernel;
  // from here on global memory written by it is read
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // work index -> (group, split, m-tile, n-tile): 32-bit arithmetic (64-bit div / mod was 15 % of this kernel's instructions)
  const uint32_t total_work = static_cast<uint32_t>(p.total_work);
  const uint32_t n_tiles = static_cast<uint32_t>(p.n_tiles), m_tiles = static_cast<uint32_t>(p.m_tiles), n_splits = static_cast<uint32_t>(p.splits);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (uint32_t w = blockIdx.x; w < total_work; w += gridDim.x) {
        const uint32_t q0 = w / n_tiles, q1 = q0 / m_tiles;
        const int nt = static_cast<int>(w - q0 * n_tiles);
        const int mt = static_cast<int>(q0 - q1 * m_tiles);
        const uint32_t gq = q1 / n_splits;
        const int ks = static_cast<int>(q1 - gq * n_splits);
        const int g = static_cast<int>(gq);
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          mbar_expect_tx(full + stage, Cfg::kStageBytes);
          uint8_t* sa = smem_a + stage * Cfg::kABytes;
          uint8_t* sb = smem_b + stage * Cfg::kBBytes;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_3d(sa + j * 8192, &map_a, full + stage, mt * BM + j * 64, kb * BK, g);
          } else {
            tma_load_3d(sa, &map_a, full + stage, kb * BK, mt * BM, g);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_3d(sb + j * 8192, &map_b, full + stage, nt * BN + j * 64, kb * BK, g);
          } else if (p.kb_switch > 0 && kb >= p.kb_switch) {
            tma_load_3d(sb, &map_b2, full + stage, (kb - p.kb_switch) * BK, nt * BN, g);
          } else {
            tma_load_3d(sb, &map_b, full + stage, kb * BK, nt * BN, g);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, A_MN, B_MN);
      // K-major SW128: 8-row groups 1024 B apart (SBO), K advance 32 B inside the swizzled row.
      // MN-major SW128: 64-element MN atoms (BK rows x 128 B = 8192 B apart, LBO), 8-k-row groups 1024 B apart (SBO),
      //                 K advance 16 rows = 2048 B.
      constexpr uint32_t a_lbo = A_MN ? 8192u : 16u, b_lbo = B_MN ? 8192u : 16u;
      constexpr uint32_t a_kstep = A_MN ? 2048u : 32u, b_kstep = B_MN ? 2048u : 32u;
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (uint32_t w = blockIdx.x; w < total_work; w += gridDim.x) {
        const uint32_t r = (w / n_tiles) / m_tiles;
        const int ks = static_cast<int>(r % n_splits);
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        mbar_wait(acc_empty + as, aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t sb = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc(sa + k * a_kstep, a_lbo, 1024u);
            const uint64_t db = make_smem_desc(sb + k * b_kstep, b_lbo, 1024u);
            umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty + stage);              // frees the smem slot once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full + as);                // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Warp (2 + i) may only touch TMEM lanes 32 * ((2 + i) & 3) ...; the two warps of a quarter take alternate
    // 32-column chunks.  The TMEM load of the next chunk is in flight while the current one is converted and stored.
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;              // which of the quarter's warps (0 .. kEpiWarps/4 - 1)
    int as = 0; uint32_t aphase = 0;
    float loss_acc = 0.f;
    for (uint32_t w = blockIdx.x; w < total_work; w += gridDim.x) {
      const uint32_t q0 = w / n_tiles, q1 = q0 / m_tiles;
      const int nt = static_cast<int>(w - q0 * n_tiles);
      const int mt = static_cast<int>(q0 - q1 * m_tiles);
      const int g = static_cast<int>(q1 / n_splits);
      mbar_wait(acc_full + as, aphase);
      tc_fence_after();
      const int row = mt * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
      constexpr int kChunks = BN / 32;
      const int ncols_tile = min(BN, p.N - nt * BN);
      const int nchunks = (ncols_tile + 31) / 32;
      if constexpr (kEpiWarps == 8) {
        uint32_t va[32], vb[32];
        int c = half;
        if (c < nchunks) tmem_ld32(taddr + c * 32, va);
#pragma unroll 1
        for (; c < nchunks; c += 4) {
          tmem_ld_wait();
          const bool more = (c + 2) < nchunks;
          if (more) tmem_ld32(taddr + (c + 2) * 32, vb);
          epilogue_chunk(p, g, row, nt * BN + c * 32, va, loss_acc);
          if (more) {
            tmem_ld_wait();
            if (c + 4 < nchunks) tmem_ld32(taddr + (c + 4) * 32, va);
            epilogue_chunk(p, g, row, nt * BN + (c + 2) * 32, vb, loss_acc);
          }
        }
      } else {
        uint32_t va[32];
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          tmem_ld32(taddr + c * 32, va);
          tmem_ld_wait();
          epilogue_chunk(p, g, row, nt * BN + c * 32, va, loss_acc);
        }
      }
      (void)kChunks;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + as);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (p.epi == kEpiLossGrad) {            // one partial per epilogue warp; summed in fixed order by loss_total_kernel
      loss_acc = warp_sum(loss_acc);
      if (lane == 0) p.loss_partials[blockIdx.x * kEpiWarps + (warp - 2)] = loss_acc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::kTmemCols); }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 [G][outer][inner] tensor (row pitch ld, group pitch gs, in elements) with a (box_inner x box_outer x 1) box,
// 128-byte swizzle: the one tensor-map shape every tensor-core kernel of this library uses (loads and stores).
int encode_tmap_bf16_3d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t G, int64_t ld, int64_t gs,
                        int box_inner, int box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  MFVAE_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  MFVAE_CHECK(reinterpret_cast<uintptr_t>(base) % 16 == 0, "tensor map: base must be 16-byte aligned");
  MFVAE_CHECK(ld % 8 == 0 && (G == 1 || (gs % 8 == 0 && gs > 0)), "tensor map: pitches must be multiples of 8 elements");
  MFVAE_CHECK(box_inner * 2 <= 128 && box_outer <= 256, "tensor map: box too large");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer), static_cast<cuuint64_t>(G)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(G == 1 ? outer * ld : gs) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer), 1}, estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MFVAE_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
  return 0;
}

struct TcPlan {
  GemmOp op;
  TcParams prm;
  CUtensorMap map_a, map_b, map_b2;
  int BN = 0; bool a_mn = false, b_mn = false;
  int cps = 1;                // CTAs per SM of the chosen launch shape
  int grid = 0;
};

// operand(r, k) with element strides (rs, cs): K-major when cs == 1, MN-major when rs == 1
static int encode_operand(CUtensorMap* map, const void* base, int rows, int K, int G, int64_t gs, int64_t rs, int64_t cs,
                          int box_rows_kmajor, bool* is_mn) {
  EncodeTiledFn enc = get_encode_fn();
  MFVAE_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  MFVAE_CHECK(reinterpret_cast<uintptr_t>(base) % 16 == 0, "tcgen05 GEMM: operand base must be 16-byte aligned");
  const bool mn = (rs == 1 && cs != 1);
  *is_mn = mn;
  MFVAE_CHECK(mn || cs == 1, "tcgen05 GEMM: one operand stride must be 1");
  const int64_t ld = mn ? cs : rs;
  MFVAE_CHECK(ld % 8 == 0, "tcgen05 GEMM: leading dimension must be a multiple of 8 elements (16 bytes)");
  MFVAE_CHECK(G == 1 || (gs % 8 == 0 && gs > 0), "tcgen05 GEMM: group stride must be a positive multiple of 8 elements");
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  const int inner = mn ? rows : K, outer = mn ? K : rows;
  dims[0] = static_cast<cuuint64_t>(inner); dims[1] = static_cast<cuuint64_t>(outer); dims[2] = static_cast<cuuint64_t>(G);
  strides[0] = static_cast<cuuint64_t>(ld) * 2;
  strides[1] = static_cast<cuuint64_t>(G == 1 ? static_cast<int64_t>(outer) * ld : gs) * 2;
  box[0] = 64; box[1] = static_cast<cuuint32_t>(mn ? BK : box_rows_kmajor); box[2] = 1;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MFVAE_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
  return 0;
}

template <int BN, bool A_MN, bool B_MN, int CPS>
static int launch_tc(const TcPlan* pl, cudaStream_t s) {
  using Cfg = TcCfg<BN, CPS>;
  constexpr int kTcThreads = TcShape<CPS>::kThreads;
  static bool attr_set = false;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, CPS>;
  if (!attr_set) {
    MFVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pl->grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  MFVAE_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->map_a, pl->map_b, pl->map_b2, pl->prm));
  MFVAE_LAUNCH_CHECK();
  return 0;
}

template <int BN, int CPS>
static int launch_tc_major(const TcPlan* pl, cudaStream_t s) {
  if (!pl->a_mn && !pl->b_mn) return launch_tc<BN, false, false, CPS>(pl, s);
  if (!pl->a_mn && pl->b_mn) return launch_tc<BN, false, true, CPS>(pl, s);
  if (pl->a_mn && pl->b_mn) return launch_tc<BN, true, true, CPS>(pl, s);
  return launch_tc<BN, true, false, CPS>(pl, s);
}

int gemm_tc_plan(const GemmOp& op, TcPlan** out) {
  MFVAE_CHECK(op.dtype == kBF16, "tcgen05 GEMM: operands must be bf16");
  MFVAE_CHECK(op.M > 0 && op.N > 0 && op.K > 0 && op.G > 0, "tcgen05 GEMM: empty problem");
  MFVAE_CHECK(op.epi != kEpiAccum || op.c_dtype == kF32, "tcgen05 GEMM: accumulate epilogue needs fp32 C");
  MFVAE_CHECK(op.split_k == 1 || op.epi == kEpiAccum, "tcgen05 GEMM: split-K needs the accumulate epilogue");
  MFVAE_CHECK(op.c_ld % (op.c_dtype == kBF16 ? 8 : 4) == 0 && op.c_gs % 8 == 0, "tcgen05 GEMM: C leading dimension / group stride alignment");
  MFVAE_CHECK(reinterpret_cast<uintptr_t>(op.C) % 16 == 0, "tcgen05 GEMM: C must be 16-byte aligned");
  MFVAE_CHECK(op.epi != kEpiReluMask || (op.aux && op.aux_ld % 8 == 0 && op.aux_gs % 8 == 0 && reinterpret_cast<uintptr_t>(op.aux) % 16 == 0),
              "tcgen05 GEMM: relu-mask aux alignment");
  MFVAE_CHECK(op.epi != kEpiLossGrad || (op.c_dtype == kBF16 && op.split_k == 1), "tcgen05 GEMM: the loss epilogue writes bf16 gradients, no split-K");
  MFVAE_CHECK((op.epi != kEpiBias && op.epi != kEpiBiasRelu && op.epi != kEpiLossGrad) || (op.bias && reinterpret_cast<uintptr_t>(op.bias) % 16 == 0 && op.bias_gs % 4 == 0),
              "tcgen05 GEMM: bias alignment");
  TcPlan* pl = new TcPlan();
  pl->op = op;
  const int m_tiles = (op.M + BM - 1) / BM;
  // tile width: the widest BN that still gives every SM a tile; otherwise the narrowest that covers N
  const int k_blocks = (op.K + BK - 1) / BK;
  int BN = 64, forced_splits = 0;
  {
    const int cands[3] = {256, 128, 64};
    for (int c : cands) {
      if (c > 64 && c / 2 >= op.N) continue;        // do not pad N by 2x or more
      const long long t = static_cast<long long>(op.G) * m_tiles * ((op.N + c - 1) / c);
      if (t >= kNumSMs) { BN = c; break; }
    }
    // K-long GEMMs are bound by operand traffic from L2 (a 128 x BN x 64 step moves (128 + BN) * 128 bytes for
    // 128 * BN * 128 flops): a narrower tile that fills more SMs loses more to bandwidth than the idle SMs cost.
    // Model (microseconds): rounds over the 148 SMs x (k-blocks per work item x time of one 128 x BN x 64 step + epilogue),
    // step times measured on this kernel: 0.57 / 0.44 / 0.41 us for BN = 256 / 128 / 64.  Single-matrix weight gradients also
    // choose their split-K here: a work item is (tile, K range), items beyond one round cost a whole extra round, and a
    // split epilogue pays fp32 atomics instead of plain stores.  (Grouped wgrads of the small layers keep the narrow shape.)
    forced_splits = 0;
    if (k_blocks > 16 && (op.epi != kEpiAccum || op.G == 1)) {
      const double step_us[3] = {0.57, 0.44, 0.41};
      double best = 1e30;
      for (int i = 0; i < 3; ++i) {
        const int c = cands[i];
        if (c > 64 && c / 2 >= op.N) continue;
        const long long t = static_cast<long long>(op.G) * m_tiles * ((op.N + c - 1) / c);
        const int max_s = (op.epi == kEpiAccum) ? std::max(1, std::min(16, k_blocks / 4)) : 1;
        for (int sp = 1; sp <= max_s; ++sp) {
          const int kbs = (k_blocks + sp - 1) / sp;
          if (sp > 1 && (k_blocks + kbs - 1) / kbs != sp) continue;          // this split count leaves an empty split
          const double epi = (c / 64) * (sp > 1 ? 2.0 : 1.0) + 2.0;
          const double cost = static_cast<double>((t * sp + kNumSMs - 1) / kNumSMs) * (kbs * step_us[i] + epi) * (sp > 1 ? 1.1 : 1.0);   // near-ties keep plain stores
          if (cost < best) { best = cost; BN = c; forced_splits = sp; }
        }
      }
    }
  }
  // Narrow tiles, three co-resident CTAs per SM (75 KB smem, 128 TMEM columns each) when
  //  * K <= 256: the tile is a handful of MMAs plus an epilogue, latency hiding comes from the neighbours; or
  //  * 128-wide tiles cannot give every SM one tile anyway (the small layers): a CTA of this shape leaves room for the
  //    kernels of the other backward chain (side stream) on the same SM, and 3 x 72 KB of operands in flight per SM
  //    covers the L2 latency-bandwidth product as well as one deep ring does.
  // (measured again with 256-wide tiles kept for K <= 256: enc layer 3 forward 37 -> 52 us, its dgrad 56 -> 69 us)
  if (k_blocks <= 4 || (BN == 64 && forced_splits == 0)) { BN = 64; pl->cps = 3; }     // BN == 64 here: not even 128-wide tiles reach 148
  const int n_tiles = (op.N + BN - 1) / BN;
  int splits = 1;
  if (op.epi == kEpiAccum && forced_splits > 0 && op.split_k <= 1) {
    splits = forced_splits;
  } else if (op.epi == kEpiAccum) {
    const long long tiles = static_cast<long long>(op.G) * m_tiles * n_tiles;
    long long want = op.split_k > 1 ? op.split_k : 1;
    // whole multiples only: splitting 80 tiles in two gives 160 work items = two rounds of half tiles on 148 SMs, i.e. the
    // time of one round of whole tiles plus the atomics
    want = std::max<long long>(want, (static_cast<long long>(kNumSMs) * pl->cps) / tiles);
    splits = static_cast<int>(std::max<long long>(1, std::min<long long>(want, std::max(1, k_blocks / 4))));
  }
  const int kb_per_split = (k_blocks + splits - 1) / splits;
  splits = (k_blocks + kb_per_split - 1) / kb_per_split;       // no empty splits
  TcParams& p = pl->prm;
  p.G = op.G; p.M = op.M; p.N = op.N; p.K = op.K;
  p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.k_blocks = k_blocks; p.splits = splits; p.kb_per_split = kb_per_split;
  p.total_work = static_cast<long long>(op.G) * splits * m_tiles * n_tiles;
  if (p.total_work >= (1LL << 31)) { delete pl; MFVAE_FAIL("tcgen05 GEMM: too many tiles"); }
  p.C = op.C; p.c_gs = op.c_gs; p.c_ld = op.c_ld; p.c_dtype = op.c_dtype;
  p.bias = op.bias; p.bias_gs = op.bias_gs; p.epi = op.epi;
  p.aux = static_cast<const __nv_bfloat16*>(op.aux); p.aux_gs = op.aux_gs; p.aux_ld = op.aux_ld;
  {
    const int per32 = (op.c_dtype == kBF16) ? 16 : 8;          // elements per 32-byte sector
    p.c_v8 = (op.c_ld % per32 == 0 && (op.G == 1 || op.c_gs % per32 == 0) && reinterpret_cast<uintptr_t>(op.C) % 32 == 0) ? 1 : 0;
  }
  p.aux_v8 = (op.aux && op.aux_ld % 16 == 0 && (op.G == 1 || op.aux_gs % 16 == 0) && reinterpret_cast<uintptr_t>(op.aux) % 32 == 0) ? 1 : 0;
  p.tgt = nullptr; p.tgt_ld = 0; p.grad_scale = 0.f; p.huber = 1; p.loss_partials = nullptr;
  p.accumulate_atomic = (op.epi == kEpiAccum && splits > 1) ? 1 : 0;   // single split: plain stores into the zeroed C
  pl->BN = BN;
  pl->grid = static_cast<int>(std::min<long long>(p.total_work, static_cast<long long>(kNumSMs - g_tc_sm_reserve) * pl->cps));
  int rc = encode_operand(&pl->map_a, op.A, op.M, op.K, op.G, op.a_gs, op.a_rs, op.a_cs, BM, &pl->a_mn);
  p.kb_switch = 0;
  pl->map_b2 = pl->map_a;                  // placeholder when there is no second segment (never dereferenced)
  if (op.B2) {
    if (!(op.b_cs == 1 && op.k_split % BK == 0 && op.k_split > 0 && op.k1 > 0 && op.k1 <= op.k_split && op.K == op.k_split + op.k2 && op.split_k == 1)) {
      delete pl; MFVAE_FAIL("tcgen05 GEMM: second B segment needs a K-major B, k_split % 64 == 0, K = k_split + k2, no split-K");
    }
    bool mn2 = false;
    if (rc == 0) rc = encode_operand(&pl->map_b, op.B, op.N, op.k1, op.G, op.b_gs, op.b_rs, 1, BN, &pl->b_mn);
    if (rc == 0) rc = encode_operand(&pl->map_b2, op.B2, op.N, op.k2, op.G, op.b2_gs, op.b2_rs, 1, BN, &mn2);
    p.kb_switch = op.k_split / BK;
  } else if (rc == 0) {
    rc = encode_operand(&pl->map_b, op.B, op.N, op.K, op.G, op.b_gs, op.b_rs, op.b_cs, BN, &pl->b_mn);
  }
  if (rc != 0) { delete pl; return rc; }
  *out = pl;
  return 0;
}

int gemm_tc_run(const TcPlan* pl, cudaStream_t s) {
  MFVAE_CHECK(pl != nullptr, "tcgen05 GEMM: null plan");
  if (pl->cps == 3) return launch_tc_major<64, 3>(pl, s);
  switch (pl->BN) {
    case 64: return launch_tc_major<64, 1>(pl, s);
    case 128: return launch_tc_major<128, 1>(pl, s);
    case 256: return launch_tc_major<256, 1>(pl, s);
  }
  MFVAE_FAIL("tcgen05 GEMM: unsupported tile width");
}

void gemm_tc_free(TcPlan* p) { delete p; }

int gemm_tc_set_loss(TcPlan* pl, const float* tgt, int64_t tgt_ld, float grad_scale, int huber, float* partials) {
  MFVAE_CHECK(pl && pl->prm.epi == kEpiLossGrad, "tcgen05 GEMM: not a loss-epilogue plan");
  pl->prm.tgt = tgt; pl->prm.tgt_ld = tgt_ld; pl->prm.grad_scale = grad_scale; pl->prm.huber = huber; pl->prm.loss_partials = partials;
  return 0;
}
int gemm_tc_loss_partials(const TcPlan* pl) {
  if (!pl) return 0;
  return pl->grid * (pl->cps == 1 ? TcShape<1>::kEpiWarps : TcShape<3>::kEpiWarps);
}
bool gemm_tc_overwrites(const TcPlan* p) { return p && p->prm.accumulate_atomic == 0; }

}  // namespace mfvae
