// enc_fused.cu — the 40 per-agent encoders as ONE persistent tcgen05 kernel (sm_100a), two CTAs per SM.
//
// Reference being reproduced (torch_ver/model.py): per agent, Encoder MLP in -> 64 -> 64 -> 256 -> 2L with ReLU (:43-57,144),
// mu / logvar split (:149-150), reparameterize (:77-81,151) and the agent's KL term (:35-37).  Layer by layer these are 160
// tiny GEMMs whose activations make a round trip through HBM each; here one CTA owns a (agent, 128-sample tile) unit and
// chains the layers on chip.  The id-embedding is folded into layer 0's bias (fold.cu), so the unit's input is the staged
// bf16 tile X0F = [obs | 0] the backward pass needs anyway:
//
//   control (1 thread)  TMA-loads the unit's X0F tile and, layer by layer, that layer's weight matrix into ONE 32 KB buffer
//                       (streamed from L2 per unit: nothing agent-sized stays resident, which is what lets two CTAs share an
//                       SM and hide each other's latencies), issues each layer's tcgen05.mma chain into the SAME TMEM columns
//                       (layer l+1 starts only after layer l's epilogue has drained them), and TMA-stores every hidden
//                       activation tile to HBM for the backward pass while the MMAs that read the same tile run
//   workers (8 warps)   every layer's epilogue: TMEM -> +bias -> ReLU -> bf16 -> shared memory (the next layer's A operand) and,
//                       for the last layer, mu / logvar -> LAT, reparameterize -> z, KL partial
//
// Synchronisation is mbarriers only: x0_full / w_full (TMA -> control), act_ready[l] (workers -> control: layer l's input is in
// shared memory), mma_done[l] (tcgen05.commit -> workers and control), tmem_free (workers -> control: the last layer's
// accumulator has been read, the next unit's layer 0 may overwrite it).
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "kernels.h"
#include "tc_ptx.cuh"

namespace mfvae {

constexpr int kEncRows = 128;                  // samples per unit = UMMA M
constexpr int kBoxBytes = kEncRows * 128;      // one [128 rows x 64 bf16] swizzled box
constexpr int kActBoxes = 4;                   // activations up to 256 columns wide
constexpr int kEncWorkers = 256;
constexpr int kEncThreads = kEncWorkers + 32;  // + the control warp
constexpr int kEncTmemCols = 256;              // one accumulator region, reused by every layer
constexpr size_t kEncSmemLimit = 113 * 1024 - 1024;   // two CTAs per SM

struct EncFwdParams {
  int A, B, tiles, total_units, nl;
  int N[kEncMaxL], kboxes[kEncMaxL], ksteps[kEncMaxL];
  int in_box[kEncMaxL], wait_store[kEncMaxL];   // first activation box of layer l's input; 1: pending TMA stores must drain before the epilogue
  uint32_t w_bytes[kEncMaxL], w_buf;           // bytes of layer l's weight boxes / size of the streaming buffer
  int bias_off[kEncMaxL], bias_n;              // bias l lives at [bias_off[l], +N_l) of the per-agent bias block
  const float* bias[kEncMaxL];                 // [A][N_l] fp32 (layer 0: the folded bias b0 + W0[:, :I] . emb[a])
  float* lat; long long lat_gs, lat_ld;
  __nv_bfloat16* zin; long long zin_ld;
  const float* eps; long long eps_ld;
  unsigned long long seed, step; long long sample0;
  int L;
  float kl_scale; float* kl_out; float* scratch;
};
struct alignas(64) EncMaps { CUtensorMap w[kEncMaxL]; CUtensorMap x[kEncMaxL]; };

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(kEncThreads, 2)
enc_fwd_kernel(const __grid_constant__ EncMaps maps, const EncFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float red[32];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* act = smem;                                    // [4][16 KB]: X0F, then every hidden activation
  uint8_t* wsm = smem + kActBoxes * kBoxBytes;            // the current layer's weight matrix, K-major boxes of [N_l x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + p.w_buf);
  uint64_t* act_ready = bars;                             // [kEncMaxL]  ([0] unused: layer 0's input arrives by TMA)
  uint64_t* mma_done = bars + kEncMaxL;                   // [kEncMaxL]
  uint64_t* w_full = bars + 2 * kEncMaxL;
  uint64_t* x0_full = bars + 2 * kEncMaxL + 1;
  uint64_t* tmem_free = bars + 2 * kEncMaxL + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kEncMaxL + 3);
  float* bias_s = reinterpret_cast<float*>(bars + 16);    // 128 bytes in: 16-byte aligned; [2][bias_n], double-buffered per agent

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int l = 0; l < kEncMaxL; ++l) { mbar_init(act_ready + l, kEncWorkers); mbar_init(mma_done + l, 1); }
    mbar_init(w_full, 1);
    mbar_init(x0_full, 1);
    mbar_init(tmem_free, kEncWorkers);
    fence_barrier_init();
  }
  if (warp == 8) {
    if (lane == 0) for (int l = 0; l < p.nl; ++l) { prefetch_tmap(&maps.w[l]); prefetch_tmap(&maps.x[l]); }
    __syncwarp();
    tmem_alloc(tmem_slot, kEncTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int u0 = static_cast<int>(static_cast<long long>(blockIdx.x) * p.total_units / gridDim.x);
  const int u1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.total_units / gridDim.x);
  const int nl = p.nl;
  float kl_acc = 0.f;

  if (warp == 8) {
    // ================================ control thread ================================
    if (lane == 0 && u0 < u1) {
      uint32_t par = 0, wpar = 0;
      const uint32_t act_s = smem_u32(act), w_s = smem_u32(wsm);
      auto load_w = [&](int l, int a) {
        mbar_expect_tx(w_full, p.w_bytes[l]);
        for (int kb = 0; kb < p.kboxes[l]; ++kb) tma_load_3d(wsm + kb * p.N[l] * 128, &maps.w[l], w_full, kb * 64, 0, a);
      };
      auto load_x0 = [&](int u) {
        const int a = u / p.tiles, t = u - a * p.tiles;
        mbar_expect_tx(x0_full, static_cast<uint32_t>(p.kboxes[0]) * kBoxBytes);
        for (int kb = 0; kb < p.kboxes[0]; ++kb) tma_load_3d(act + kb * kBoxBytes, &maps.x[0], x0_full, kb * 64, t * kEncRows, a);
      };
      load_x0(u0);
      load_w(0, u0 / p.tiles);
      for (int u = u0; u < u1; ++u, par ^= 1) {
        const int a = u / p.tiles, t = u - a * p.tiles;
        for (int l = 0; l < nl; ++l) {
          mbar_wait(w_full, wpar); wpar ^= 1;
          if (l == 0) {
            mbar_wait(x0_full, par);
            if (u > u0) mbar_wait(tmem_free, par ^ 1);          // the previous unit's last accumulator has been read
          } else {
            mbar_wait(act_ready + l, par);
          }
          tc_fence_after();
          const uint32_t in_s = act_s + p.in_box[l] * kBoxBytes;
          if (l > 0) {   // layer l's input tile: out to HBM for the backward pass while the MMAs below read the same bytes
            for (int kb = 0; kb < p.kboxes[l]; ++kb)
              tma_store_3d(&maps.x[l], act + (p.in_box[l] + kb) * kBoxBytes, kb * 64, t * kEncRows, a);
            tma_store_commit();
          }
          const uint32_t idesc = make_idesc(kEncRows, p.N[l], false, false);
          const uint32_t nbox = static_cast<uint32_t>(p.N[l]) * 128u;
          for (int ks = 0; ks < p.ksteps[l]; ++ks) {
            const uint64_t da = make_smem_desc(in_s + (ks >> 2) * kBoxBytes + (ks & 3) * 32, 16u, 1024u);
            const uint64_t db = make_smem_desc(w_s + (ks >> 2) * nbox + (ks & 3) * 32, 16u, 1024u);
            umma_bf16(tmem_base, da, db, idesc, ks > 0 ? 1u : 0u);
          }
          if (p.wait_store[l]) tma_store_wait_read();   // this layer's epilogue writes over boxes a store may still be reading
          umma_commit(mma_done + l);
          // the weight buffer (and, after the last layer, the activation boxes) are free once these MMAs have retired
          mbar_wait(mma_done + l, par);
          if (l + 1 < nl) {
            load_w(l + 1, a);
          } else if (u + 1 < u1) {
            tma_store_wait_read();                      // the last hidden activation's store has left shared memory
            load_x0(u + 1);
            load_w(0, (u + 1) / p.tiles);
          }
        }
      }
      tma_store_wait_all();
    }
  } else {
    // ================================ workers ================================
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;                                   // tile row == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int cur_a = -1, bsel = 0;
    uint32_t par = 0;
    for (int u = u0; u < u1; ++u, par ^= 1) {
      const int a = u / p.tiles, b0 = (u - a * p.tiles) * kEncRows;
      if (a != cur_a) {
        // the agent's biases -> shared memory (the other buffer may still be in use by slower warps of the previous unit)
        cur_a = a; bsel ^= 1;
        float* bs = bias_s + bsel * p.bias_n;
        for (int l = 0; l < nl; ++l)
          for (int i = threadIdx.x; i < p.N[l]; i += kEncWorkers) bs[p.bias_off[l] + i] = __ldg(p.bias[l] + static_cast<long long>(a) * p.N[l] + i);
        asm volatile("bar.sync 1, %0;" ::"n"(kEncWorkers) : "memory");
      }
      const float* const bias_u = bias_s + bsel * p.bias_n;
      // ---- hidden layers: accumulator -> +bias -> relu -> bf16 -> the next layer's A operand ----
      for (int l = 0; l + 1 < nl; ++l) {
        mbar_wait(mma_done + l, par);
        tc_fence_after();
        const int nch = p.N[l] >> 5;
        const float* bias = bias_u + p.bias_off[l];
        uint8_t* const outb = act + p.in_box[l + 1] * kBoxBytes;
        for (int c = half; c < nch; c += 2) {
          uint32_t v[32];
          tmem_ld32(lane_addr + c * 32, v);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 bb = reinterpret_cast<const float2*>(bias + c * 32)[j];
            o[j] = pack_bf16x2(__uint_as_float(v[2 * j]) + bb.x, __uint_as_float(v[2 * j + 1]) + bb.y, true);
          }
          uint8_t* box = outb + (c >> 1) * kBoxBytes;
          const int ch0 = (c & 1) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(box + sw128_chunk_off(row, ch0 + i)) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(act_ready + l + 1);
      }
      // ---- last layer: (mu | logvar) -> LAT, z = mu + eps * exp(logvar / 2) -> ZIN, KL partial ----
      {
        const int l = nl - 1;
        mbar_wait(mma_done + l, par);
        tc_fence_after();
        const int L = p.L, hw = L >> 1;
        const int jh = half * hw;                                   // this warp's share of the latent columns
        const float* bias = bias_u + p.bias_off[l];
        const long long b = b0 + row;
        const bool ok = b < p.B;
        float* latrow = p.lat + a * p.lat_gs + b * p.lat_ld;
        __nv_bfloat16* zrow = p.zin + b * p.zin_ld + a * L;
        for (int jj = 0; jj < hw; jj += 8) {
          const int j = jh + jj;
          uint32_t vm[8], vl[8];
          tmem_ld8(lane_addr + j, vm);
          tmem_ld8(lane_addr + L + j, vl);
          tmem_ld_wait();
          if (ok) {
            float mu[8], lv[8];
#pragma unroll
            for (int k = 0; k < 8; k += 4) {
              const float4 bm = *reinterpret_cast<const float4*>(bias + j + k);
              const float4 bl = *reinterpret_cast<const float4*>(bias + L + j + k);
              mu[k] = __uint_as_float(vm[k]) + bm.x; mu[k + 1] = __uint_as_float(vm[k + 1]) + bm.y;
              mu[k + 2] = __uint_as_float(vm[k + 2]) + bm.z; mu[k + 3] = __uint_as_float(vm[k + 3]) + bm.w;
              lv[k] = __uint_as_float(vl[k]) + bl.x; lv[k + 1] = __uint_as_float(vl[k + 1]) + bl.y;
              lv[k + 2] = __uint_as_float(vl[k + 2]) + bl.z; lv[k + 3] = __uint_as_float(vl[k + 3]) + bl.w;
              *reinterpret_cast<float4*>(latrow + j + k) = make_float4(mu[k], mu[k + 1], mu[k + 2], mu[k + 3]);
              *reinterpret_cast<float4*>(latrow + L + j + k) = make_float4(lv[k], lv[k + 1], lv[k + 2], lv[k + 3]);
            }
            float4 e4[2];
            if (p.eps) {
#pragma unroll
              for (int k = 0; k < 2; ++k) e4[k] = *reinterpret_cast<const float4*>(p.eps + b * p.eps_ld + a * L + j + 4 * k);
            } else {
#pragma unroll
              for (int k = 0; k < 2; ++k)
                e4[k] = philox_normal4(p.seed, p.step, static_cast<uint64_t>(p.sample0 + b), static_cast<uint32_t>((a * L + j + 4 * k) >> 2));
            }
            float z[8];
#pragma unroll
            for (int k = 0; k < 8; k += 4) {
              const float ev[4] = {e4[k >> 2].x, e4[k >> 2].y, e4[k >> 2].z, e4[k >> 2].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float sd = expf(0.5f * lv[k + i]);
                z[k + i] = mu[k + i] + ev[i] * sd;
                kl_acc += 1.f + lv[k + i] - mu[k + i] * mu[k + i] - sd * sd;
              }
            }
            *reinterpret_cast<uint4*>(zrow + j) = make_uint4(pack_bf16x2(z[0], z[1], false), pack_bf16x2(z[2], z[3], false),
                                                             pack_bf16x2(z[4], z[5], false), pack_bf16x2(z[6], z[7], false));
          }
        }
        tc_fence_before();
        mbar_arrive(tmem_free);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, kEncTmemCols); }
  const float tot = block_sum(kl_acc, red);
  finish_scalar(tot, p.scratch, -0.5f * p.kl_scale, p.kl_out, red);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct EncFusedPlan {
  EncMaps fmaps;
  EncFwdParams fp;
  int grid = 0;
  size_t fwd_smem = 0;
};

static size_t enc_smem_bytes(const EncFusedDesc& d, uint32_t* w_buf_out) {
  uint32_t wbuf = 0; int cols = 0;
  for (int l = 0; l < d.nl; ++l) {
    const int K = (l == 0) ? d.K0 : d.N[l - 1];
    wbuf = std::max<uint32_t>(wbuf, static_cast<uint32_t>((K + 63) / 64) * d.N[l] * 128);
    cols += d.N[l];
  }
  if (w_buf_out) *w_buf_out = wbuf;
  return static_cast<size_t>(kActBoxes) * kBoxBytes + wbuf + 256 + 2 * cols * sizeof(float) + 1024;
}

bool enc_fused_applicable(const EncFusedDesc& d) {
  if (d.nl < 2 || d.nl > kEncMaxL) return false;
  if (d.K0 > 64 * kActBoxes || d.K0 % 8 != 0 || d.L % 16 != 0 || 2 * d.L > 256) return false;
  for (int l = 0; l < d.nl; ++l) {
    const int N = d.N[l];
    if (N % 16 != 0 || N > kEncTmemCols) return false;
    if (l + 1 < d.nl && N % 64 != 0) return false;
  }
  return enc_smem_bytes(d, nullptr) <= 227 * 1024 - 2048;
}

int enc_fused_plan(const EncFusedDesc& d, EncFusedPlan** out) {
  MFVAE_CHECK(enc_fused_applicable(d), "fused encoder: shape not supported");
  EncFusedPlan* pl = new EncFusedPlan();
  EncFwdParams& p = pl->fp;
  memset(&p, 0, sizeof(p));
  p.A = d.A; p.B = d.B; p.tiles = (d.B + kEncRows - 1) / kEncRows; p.total_units = p.A * p.tiles; p.nl = d.nl;
  int col = 0;
  int rc = 0;
  for (int l = 0; l < d.nl && rc == 0; ++l) {
    const int N = d.N[l], K = (l == 0) ? d.K0 : d.N[l - 1];
    p.N[l] = N; p.kboxes[l] = (K + 63) / 64; p.ksteps[l] = (K + 15) / 16;
    p.w_bytes[l] = static_cast<uint32_t>(p.kboxes[l]) * N * 128;
    p.bias_off[l] = col; col += N;
    p.bias[l] = d.bias[l];
    // weights of layer l: K-major [A][N][w_ld], the K valid columns start at W[l]
    rc = encode_tmap_bf16_3d(&pl->fmaps.w[l], d.W[l], K, N, d.A, d.w_ld[l], d.w_gs[l], 64, N);
    // input of layer l in HBM: loaded (l = 0: the staged X0F tile) or stored for the backward pass (l >= 1)
    if (rc == 0) rc = encode_tmap_bf16_3d(&pl->fmaps.x[l], d.X[l], K, d.B, d.A, d.x_ld[l], d.x_gs[l], 64, kEncRows);
    // activation boxes: X0F starts at box 0; a hidden activation goes next to the tile whose TMA store may still be in
    // flight when it is written, or back to box 0 (then the control thread drains the pending stores first)
    if (l == 0) p.in_box[0] = 0;
    if (l + 1 < d.nl) {
      const int in_end = p.in_box[l] + p.kboxes[l];
      const int nb_out = N / 64;
      if (in_end + nb_out <= kActBoxes) { p.in_box[l + 1] = in_end; p.wait_store[l] = 0; }
      else { p.in_box[l + 1] = 0; p.wait_store[l] = 1; }
    } else {
      p.wait_store[l] = 0;
    }
  }
  if (rc != 0) { delete pl; return rc; }
  p.bias_n = col;
  p.lat = d.lat; p.lat_gs = d.lat_gs; p.lat_ld = d.lat_ld;
  p.zin = static_cast<__nv_bfloat16*>(d.zin); p.zin_ld = d.zin_ld; p.L = d.L;
  pl->fwd_smem = enc_smem_bytes(d, &p.w_buf);
  const int per_sm = pl->fwd_smem <= kEncSmemLimit ? 2 : 1;
  pl->grid = std::min(p.total_units, kNumSMs * per_sm);
  *out = pl;
  return 0;
}

void enc_fused_free(EncFusedPlan* p) { delete p; }

int enc_fused_forward(EncFusedPlan* pl, const EncFwdBatch& b, cudaStream_t s) {
  MFVAE_CHECK(pl != nullptr, "fused encoder: null plan");
  static size_t attr_bytes = 0;
  if (pl->fwd_smem > attr_bytes) {
    MFVAE_CUDA(cudaFuncSetAttribute(enc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl->fwd_smem)));
    attr_bytes = pl->fwd_smem;
  }
  EncFwdParams p = pl->fp;
  p.eps = b.eps; p.eps_ld = b.eps_ld; p.seed = b.seed; p.step = b.step; p.sample0 = b.sample0;
  p.kl_scale = b.kl_scale; p.kl_out = b.kl_out; p.scratch = b.scratch;
  enc_fwd_kernel<<<pl->grid, kEncThreads, pl->fwd_smem, s>>>(pl->fmaps, p);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mfvae
