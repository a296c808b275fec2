// enc_fused.cu — the 40 per-agent encoders as ONE persistent tcgen05 kernel per direction (sm_100a).
//
// Reference being reproduced (torch_ver/model.py): per agent, idx_emb lookup + cat with the observation (:142-143),
// Encoder MLP in -> 64 -> 64 -> 256 -> 2L with ReLU (:43-57,144), mu / logvar split (:149-150), reparameterize
// (:77-81,151) and the agent's KL term (:35-37).  Layer by layer these are 160 tiny GEMMs whose activations make a
// round trip through HBM each; here one CTA owns a (agent, 128-sample tile) unit and chains the layers on chip:
//
//   workers (8 warps)   build X0 = [idx_emb | obs_a | 0] as bf16 straight into 128B-swizzled shared memory, later run every
//                       layer's epilogue: TMEM -> +bias -> ReLU -> bf16 -> shared memory (the next layer's A operand, written
//                       in place over the previous activation) and, for the last layer, mu / logvar -> reparameterize -> z, KL
//   control (1 thread)  TMA-loads the agent's four weight matrices once per agent (they stay resident in shared memory),
//                       issues each layer's tcgen05.mma chain into its own TMEM columns, and TMA-stores every activation
//                       tile to HBM (the backward pass needs them) while the MMAs that read the same tile run
//
// Synchronisation is mbarriers only: act_ready[l] (workers -> control: layer l's input is in shared memory),
// mma_done[l] (tcgen05.commit -> workers: layer l's accumulator is complete AND its input tile may be overwritten).
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
constexpr int kEncTmemCols = 512;
constexpr size_t kEncSmemLimit = 227 * 1024 - 1024;   // dynamic + static shared memory must fit 227 KB per CTA

struct EncFwdParams {
  int A, B, tiles, total_units, nl;
  int N[kEncMaxL], kboxes[kEncMaxL], ksteps[kEncMaxL], tmem_col[kEncMaxL];
  uint32_t w_off[kEncMaxL], w_bytes;           // weight offsets inside the weight region / its size
  const float* bias[kEncMaxL];                 // [A][N_l] fp32
  const float* obs; long long obs_ld;
  const float* idx; int idx_ld;                // optional explicit agent-index column [B][A]
  const float* idx_emb; int I;
  const int32_t* obs_off; const int32_t* obs_dim;
  float* lat; long long lat_gs, lat_ld;
  __nv_bfloat16* zin; long long zin_ld;
  const float* eps; long long eps_ld;
  unsigned long long seed, step; long long sample0;
  int L;
  float kl_scale; float* kl_out; float* scratch;
};
struct alignas(64) EncMaps { CUtensorMap w[kEncMaxL]; CUtensorMap x[kEncMaxL]; };

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1], false), pack_bf16x2(f[2], f[3], false), pack_bf16x2(f[4], f[5], false),
                    pack_bf16x2(f[6], f[7], false));
}

__global__ void __launch_bounds__(kEncThreads, 1)
enc_fwd_kernel(const __grid_constant__ EncMaps maps, const EncFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float red[32];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* act = smem;                                    // [4][16 KB]: X0, then every hidden activation, in place
  uint8_t* wsm = smem + kActBoxes * kBoxBytes;            // the agent's weight matrices, K-major boxes of [N_l x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + p.w_bytes);
  uint64_t* act_ready = bars;
  uint64_t* mma_done = bars + kEncMaxL;
  uint64_t* w_full = bars + 2 * kEncMaxL;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kEncMaxL + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int l = 0; l < kEncMaxL; ++l) { mbar_init(act_ready + l, kEncWorkers); mbar_init(mma_done + l, 1); }
    mbar_init(w_full, 1);
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
    if (lane == 0) {
      uint32_t par = 0, wpar = 0;
      int cur_a = -1;
      const uint32_t act_s = smem_u32(act), w_s = smem_u32(wsm);
      for (int u = u0; u < u1; ++u, par ^= 1) {
        const int a = u / p.tiles, t = u - a * p.tiles;
        if (a != cur_a) {                      // every MMA of the previous unit has retired (wait at the loop's end)
          cur_a = a;
          mbar_expect_tx(w_full, p.w_bytes);
          for (int l = 0; l < nl; ++l)
            for (int kb = 0; kb < p.kboxes[l]; ++kb)
              tma_load_3d(wsm + p.w_off[l] + kb * p.N[l] * 128, &maps.w[l], w_full, kb * 64, 0, a);
          mbar_wait(w_full, wpar); wpar ^= 1;
        }
        for (int l = 0; l < nl; ++l) {
          mbar_wait(act_ready + l, par);
          tc_fence_after();
          // layer l's input tile: out to HBM for the backward pass, while the MMAs below read the same bytes
          for (int kb = 0; kb < p.kboxes[l]; ++kb) tma_store_3d(&maps.x[l], act + kb * kBoxBytes, kb * 64, t * kEncRows, a);
          tma_store_commit();
          const uint32_t idesc = make_idesc(kEncRows, p.N[l], false, false);
          const uint32_t wl = w_s + p.w_off[l];
          const uint32_t nbox = static_cast<uint32_t>(p.N[l]) * 128u;
          for (int ks = 0; ks < p.ksteps[l]; ++ks) {
            const uint64_t da = make_smem_desc(act_s + (ks >> 2) * kBoxBytes + (ks & 3) * 32, 16u, 1024u);
            const uint64_t db = make_smem_desc(wl + (ks >> 2) * nbox + (ks & 3) * 32, 16u, 1024u);
            umma_bf16(tmem_base + p.tmem_col[l], da, db, idesc, ks > 0 ? 1u : 0u);
          }
          tma_store_wait_read();               // the store engine is done with the tile: the epilogue may overwrite it
          umma_commit(mma_done + l);
        }
        mbar_wait(mma_done + (nl - 1), par);
      }
      tma_store_wait_all();
    }
  } else {
    // ================================ workers ================================
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;                                   // tile row == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nchunk0 = p.kboxes[0] * 8;                             // 16-byte chunks per X0 row
    int cur_a = -1, od = 0, off = 0;
    uint4 emb_pk = make_uint4(0, 0, 0, 0);
    uint32_t par = 0;
    for (int u = u0; u < u1; ++u, par ^= 1) {
      const int a = u / p.tiles, t = u - a * p.tiles;
      const int b0 = t * kEncRows;
      if (a != cur_a) {
        cur_a = a; od = p.obs_dim[a]; off = p.obs_off[a];
        if (!p.idx && lane * 8 < p.I) {
          float e[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) e[k] = __ldg(p.idx_emb + static_cast<long long>(a) * p.I + lane * 8 + k);
          emb_pk = pack8(e);
        }
      }
      // ---- X0 = [idx_emb | obs_a | 0]: warp w builds rows w, w + 8, ...; lane = 16-byte chunk (8 columns) of the row ----
      if (lane < nchunk0) {
        const int col0 = lane * 8;
        const int j0 = col0 - p.I;
        uint8_t* box = act + (lane >> 3) * kBoxBytes;
#pragma unroll 1
        for (int rr = 0; rr < kEncRows / 8; rr += 4) {
          uint4 val[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = warp + 8 * (rr + i);
            const long long b = b0 + r;
            val[i] = make_uint4(0, 0, 0, 0);
            if (b < p.B) {
              if (col0 < p.I) {
                if (p.idx) {
                  const int id = static_cast<int>(p.idx[b * p.idx_ld + a]);
                  float e[8];
#pragma unroll
                  for (int k = 0; k < 8; ++k) e[k] = __ldg(p.idx_emb + static_cast<long long>(id) * p.I + col0 + k);
                  val[i] = pack8(e);
                } else {
                  val[i] = emb_pk;
                }
              } else if (j0 < od) {
                const float* src = p.obs + b * p.obs_ld + off + j0;
                float e[8];
                if (j0 + 8 <= od && (reinterpret_cast<uintptr_t>(src) & 7) == 0) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 v2 = __ldg(reinterpret_cast<const float2*>(src) + k);
                    e[2 * k] = v2.x; e[2 * k + 1] = v2.y;
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < 8; ++k) e[k] = (j0 + k < od) ? __ldg(src + k) : 0.f;
                }
                val[i] = pack8(e);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = warp + 8 * (rr + i);
            *reinterpret_cast<uint4*>(box + sw128_chunk_off(r, lane & 7)) = val[i];
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(act_ready + 0);

      // ---- hidden layers: accumulator -> +bias -> relu -> bf16 -> the next layer's A operand (in place) ----
      for (int l = 0; l + 1 < nl; ++l) {
        mbar_wait(mma_done + l, par);
        tc_fence_after();
        const int nch = p.N[l] >> 5;
        const float* bias = p.bias[l] + static_cast<long long>(a) * p.N[l];
        for (int c = half; c < nch; c += 2) {
          uint32_t v[32];
          tmem_ld32(lane_addr + p.tmem_col[l] + c * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 bb = __ldg(reinterpret_cast<const float2*>(bias + c * 32) + j);
            pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]) + bb.x, __uint_as_float(v[2 * j + 1]) + bb.y, true);
          }
          uint8_t* box = act + (c >> 1) * kBoxBytes;
          const int ch0 = (c & 1) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(box + sw128_chunk_off(row, ch0 + i)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
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
        const float* bias = p.bias[l] + static_cast<long long>(a) * 2 * L;
        const long long b = b0 + row;
        const bool ok = b < p.B;
        float* latrow = p.lat + a * p.lat_gs + b * p.lat_ld;
        __nv_bfloat16* zrow = p.zin + b * p.zin_ld + a * L;
        const uint32_t tcol = lane_addr + p.tmem_col[l];
        for (int jj = 0; jj < hw; jj += 16) {
          const int j = jh + jj;
          uint32_t vm[16], vl[16];
          tmem_ld16(tcol + j, vm);
          tmem_ld16(tcol + L + j, vl);
          tmem_ld_wait();
          if (ok) {
            float mu[16], lv[16];
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
              const float4 bm = __ldg(reinterpret_cast<const float4*>(bias + j + k));
              const float4 bl = __ldg(reinterpret_cast<const float4*>(bias + L + j + k));
              mu[k] = __uint_as_float(vm[k]) + bm.x; mu[k + 1] = __uint_as_float(vm[k + 1]) + bm.y;
              mu[k + 2] = __uint_as_float(vm[k + 2]) + bm.z; mu[k + 3] = __uint_as_float(vm[k + 3]) + bm.w;
              lv[k] = __uint_as_float(vl[k]) + bl.x; lv[k + 1] = __uint_as_float(vl[k + 1]) + bl.y;
              lv[k + 2] = __uint_as_float(vl[k + 2]) + bl.z; lv[k + 3] = __uint_as_float(vl[k + 3]) + bl.w;
              *reinterpret_cast<float4*>(latrow + j + k) = make_float4(mu[k], mu[k + 1], mu[k + 2], mu[k + 3]);
              *reinterpret_cast<float4*>(latrow + L + j + k) = make_float4(lv[k], lv[k + 1], lv[k + 2], lv[k + 3]);
            }
            float z[16];
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
              const int col = a * L + j + k;
              float4 e;
              if (p.eps) e = *reinterpret_cast<const float4*>(p.eps + b * p.eps_ld + col);
              else       e = philox_normal4(p.seed, p.step, static_cast<uint64_t>(p.sample0 + b), static_cast<uint32_t>(col >> 2));
              const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float s = expf(0.5f * lv[k + i]);
                z[k + i] = mu[k + i] + ev[i] * s;
                kl_acc += 1.f + lv[k + i] - mu[k + i] * mu[k + i] - s * s;
              }
            }
            uint32_t pz[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) pz[k] = pack_bf16x2(z[2 * k], z[2 * k + 1], false);
            *reinterpret_cast<uint4*>(zrow + j) = make_uint4(pz[0], pz[1], pz[2], pz[3]);
            *reinterpret_cast<uint4*>(zrow + j + 8) = make_uint4(pz[4], pz[5], pz[6], pz[7]);
          }
        }
        tc_fence_before();
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

bool enc_fused_applicable(const EncFusedDesc& d) {
  if (d.nl < 2 || d.nl > kEncMaxL) return false;
  if (d.K0p > 64 * kActBoxes || d.I % 8 != 0 || d.L % 32 != 0 || 2 * d.L > 256) return false;
  int cols = 0;
  size_t wbytes = 0;
  for (int l = 0; l < d.nl; ++l) {
    const int N = d.N[l], K = (l == 0) ? d.K0p : d.N[l - 1];
    if (N % 16 != 0 || N > 256) return false;
    if (l + 1 < d.nl && N % 64 != 0) return false;
    cols += N;
    wbytes += static_cast<size_t>((K + 63) / 64) * N * 128;
  }
  if (cols > kEncTmemCols) return false;
  if (kActBoxes * kBoxBytes + wbytes + 256 + 1024 > kEncSmemLimit) return false;
  return true;
}

int enc_fused_plan(const EncFusedDesc& d, EncFusedPlan** out) {
  MFVAE_CHECK(enc_fused_applicable(d), "fused encoder: shape not supported");
  EncFusedPlan* pl = new EncFusedPlan();
  EncFwdParams& p = pl->fp;
  memset(&p, 0, sizeof(p));
  p.A = d.A; p.B = d.B; p.tiles = (d.B + kEncRows - 1) / kEncRows; p.total_units = p.A * p.tiles; p.nl = d.nl;
  int col = 0; uint32_t woff = 0;
  int rc = 0;
  for (int l = 0; l < d.nl && rc == 0; ++l) {
    const int N = d.N[l], K = (l == 0) ? d.K0p : d.N[l - 1];
    p.N[l] = N; p.kboxes[l] = (K + 63) / 64; p.ksteps[l] = (K + 15) / 16; p.tmem_col[l] = col; col += N;
    p.w_off[l] = woff; woff += static_cast<uint32_t>(p.kboxes[l]) * N * 128;
    p.bias[l] = d.bias[l];
    rc = encode_tmap_bf16_3d(&pl->fmaps.w[l], d.W[l], K, N, d.A, K, static_cast<int64_t>(N) * K, 64, N);
    if (rc == 0) rc = encode_tmap_bf16_3d(&pl->fmaps.x[l], d.X[l], K, d.B, d.A, d.x_ld[l], d.x_gs[l], 64, kEncRows);
  }
  if (rc != 0) { delete pl; return rc; }
  p.w_bytes = woff;
  p.obs_off = d.obs_off; p.obs_dim = d.obs_dim; p.idx_emb = d.idx_emb; p.I = d.I;
  p.lat = d.lat; p.lat_gs = d.lat_gs; p.lat_ld = d.lat_ld;
  p.zin = static_cast<__nv_bfloat16*>(d.zin); p.zin_ld = d.zin_ld; p.L = d.L;
  pl->grid = std::min(p.total_units, kNumSMs);
  pl->fwd_smem = static_cast<size_t>(kActBoxes) * kBoxBytes + woff + 256 + 1024;
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
  p.obs = b.obs; p.obs_ld = b.obs_ld; p.idx = b.idx; p.idx_ld = b.idx_ld;
  p.eps = b.eps; p.eps_ld = b.eps_ld; p.seed = b.seed; p.step = b.step; p.sample0 = b.sample0;
  p.kl_scale = b.kl_scale; p.kl_out = b.kl_out; p.scratch = b.scratch;
  enc_fwd_kernel<<<pl->grid, kEncThreads, pl->fwd_smem, s>>>(pl->fmaps, p);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mfvae
