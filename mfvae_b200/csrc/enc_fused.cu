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
  int in_box[kEncMaxL], wait_store[kEncMaxL];   // first activation box of layer l's input; 1: its TMA store must drain before the epilogue
  uint32_t w_off[kEncMaxL], w_bytes;           // weight offsets inside the weight region / its size
  int bias_n;                                  // sum of N_l (= TMEM columns in use): bias l lives at [tmem_col[l], +N_l)
  const float* bias[kEncMaxL];                 // [A][N_l] fp32
  const float* obs; long long obs_ld;
  const float* idx; int idx_ld;                // optional explicit agent-index column [B][A]
  const float* idx_emb; int I;
  const int32_t* obs_off; const int32_t* obs_dim;
  __nv_bfloat16* xout[kEncMaxL]; long long xout_gs[kEncMaxL], xout_ld[kEncMaxL]; int xout_w[kEncMaxL];   // input of layer l in HBM
  float* lat; long long lat_gs, lat_ld;
  __nv_bfloat16* zin; long long zin_ld;
  const float* eps; long long eps_ld;
  unsigned long long seed, step; long long sample0;
  int L;
  float kl_scale; float* kl_out; float* scratch;
};
struct alignas(64) EncMaps { CUtensorMap w[kEncMaxL]; CUtensorMap x[kEncMaxL]; };

enum { kStageNone = 0, kStageEmb, kStageEmbIdx, kStageVec, kStageScalar, kStageZero };
struct StageLane {                             // one lane's role in building X0 rows of the unit being staged
  int kind, a, b0, nvalid, off;
  const float* src0;
  uint4 emb;
};

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
  uint64_t* store_done = bars + 2 * kEncMaxL + 1;           // control -> workers: the last activation store has left shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kEncMaxL + 2);
  float* bias_s = reinterpret_cast<float*>(bars + 16);               // 128 bytes in: 16-byte aligned   // [2][bias_n]: the agent's biases, double-buffered

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int l = 0; l < kEncMaxL; ++l) { mbar_init(act_ready + l, kEncWorkers); mbar_init(mma_done + l, 1); }
    mbar_init(w_full, 1);
    mbar_init(store_done, 1);
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
          const uint32_t in_s = act_s + p.in_box[l] * kBoxBytes;
          {              // layer l's input tile: out to HBM for the backward pass while the MMAs below read the same bytes
            for (int kb = 0; kb < p.kboxes[l]; ++kb)
              tma_store_3d(&maps.x[l], act + (p.in_box[l] + kb) * kBoxBytes, kb * 64, t * kEncRows, a);
            tma_store_commit();
          }
          const uint32_t idesc = make_idesc(kEncRows, p.N[l], false, false);
          const uint32_t wl = w_s + p.w_off[l];
          const uint32_t nbox = static_cast<uint32_t>(p.N[l]) * 128u;
          for (int ks = 0; ks < p.ksteps[l]; ++ks) {
            const uint64_t da = make_smem_desc(in_s + (ks >> 2) * kBoxBytes + (ks & 3) * 32, 16u, 1024u);
            const uint64_t db = make_smem_desc(wl + (ks >> 2) * nbox + (ks & 3) * 32, 16u, 1024u);
            umma_bf16(tmem_base + p.tmem_col[l], da, db, idesc, ks > 0 ? 1u : 0u);
          }
          if (p.wait_store[l]) tma_store_wait_read();   // this layer's epilogue writes over boxes a store may still be reading
          umma_commit(mma_done + l);
        }
        tma_store_wait_read();
        mbar_arrive(store_done);                        // the next unit's X0 may overwrite the activation boxes
        mbar_wait(mma_done + (nl - 1), par);
      }
      tma_store_wait_all();
    }
  } else {
    // ================================ workers ================================
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;                                   // tile row == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nchunk0 = p.kboxes[0] * 8;                             // 16-byte chunks per X0 row held in shared memory
    const int col0 = lane * 8;                                       // this lane's X0 columns [col0, col0 + 8)
    uint8_t* const x0box = act + (lane >> 3) * kBoxBytes;
    // staging state of the unit being prefetched: lane role + packed bf16 chunks of its 16 rows (warp + 8 i)
    StageLane sl;
    float raw[8][8];
    int st_a = -1, st_bias = 0;
    auto stage_setup = [&](int u) {
      const int a = u / p.tiles, t = u - a * p.tiles;
      if (a != st_a) {
        st_a = a;
        // the agent's biases -> shared memory (the other buffer may still be in use by the unit in flight)
        st_bias ^= 1;
        float* bs = bias_s + st_bias * p.bias_n;
        for (int l = 0; l < nl; ++l)
          for (int i = threadIdx.x; i < p.N[l]; i += kEncWorkers) bs[p.tmem_col[l] + i] = __ldg(p.bias[l] + static_cast<long long>(a) * p.N[l] + i);
        asm volatile("bar.sync 1, %0;" ::"n"(kEncWorkers) : "memory");
        const int od = p.obs_dim[a];
        sl.nvalid = min(8, od - (col0 - p.I));
        sl.off = p.obs_off[a] + (col0 - p.I);
        sl.emb = make_uint4(0, 0, 0, 0);
        if (!p.idx && col0 < p.I) {
          float e[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) e[k] = __ldg(p.idx_emb + static_cast<long long>(a) * p.I + col0 + k);
          sl.emb = pack8(e);
        }
      }
      sl.a = a; sl.b0 = t * kEncRows;
      sl.src0 = p.obs + static_cast<long long>(sl.b0) * p.obs_ld + sl.off;
      sl.kind = kStageNone;
      if (lane < nchunk0) {
        if (col0 < p.I) sl.kind = p.idx ? kStageEmbIdx : kStageEmb;
        else if (sl.nvalid > 0 && ((reinterpret_cast<uintptr_t>(sl.src0) | (static_cast<uintptr_t>(p.obs_ld) << 2)) & 7) == 0) sl.kind = kStageVec;
        else if (sl.nvalid > 0) sl.kind = kStageScalar;
        else sl.kind = kStageZero;
      }
    };
    // global loads of half h (rows warp + 8 (8 h + i)) into `raw`: straight-line, all in flight together
    auto stage_load = [&](int h) {
      if (sl.kind == kStageVec) {
        // 8-byte loads; the lane at the end of the agent's slice predicates the pairs past it (never reads beyond the row)
        const int npair = sl.nvalid >> 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = min(warp + 8 * (8 * h + i), p.B - 1 - sl.b0);     // rows past the batch re-read its last row
          const float* src = sl.src0 + static_cast<long long>(r) * p.obs_ld;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float2 v2 = make_float2(0.f, 0.f);
            if (k < npair) v2 = __ldg(reinterpret_cast<const float2*>(src) + k);
            else if (2 * k < sl.nvalid) v2.x = __ldg(src + 2 * k);
            raw[i][2 * k] = v2.x; raw[i][2 * k + 1] = v2.y;
          }
        }
      } else if (sl.kind == kStageScalar) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = min(warp + 8 * (8 * h + i), p.B - 1 - sl.b0);
          const float* src = sl.src0 + static_cast<long long>(r) * p.obs_ld;
#pragma unroll
          for (int k = 0; k < 8; ++k) raw[i][k] = (k < sl.nvalid) ? __ldg(src + k) : 0.f;
        }
      } else if (sl.kind == kStageEmbIdx) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long b = min(static_cast<long long>(sl.b0 + warp + 8 * (8 * h + i)), static_cast<long long>(p.B - 1));
          const int id = max(0, min(static_cast<int>(__ldg(p.idx + b * p.idx_ld + sl.a)), p.A - 1));
#pragma unroll
          for (int k = 0; k < 8; ++k) raw[i][k] = __ldg(p.idx_emb + static_cast<long long>(id) * p.I + col0 + k);
        }
      }
    };
    // pack half h and write it to shared memory (A operand of layer 0; the control thread TMA-stores the tile to HBM)
    auto stage_store = [&](int h) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp + 8 * (8 * h + i);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (sl.kind == kStageEmb) v = sl.emb;
        else if (sl.kind == kStageVec || sl.kind == kStageScalar || sl.kind == kStageEmbIdx) v = pack8(raw[i]);
        if (sl.b0 + r >= p.B) v = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(x0box + sw128_chunk_off(r, lane & 7)) = v;
      }
    };

    uint32_t par = 0;
    for (int u = u0; u < u1; ++u, par ^= 1) {
      stage_setup(u);
      const int a = sl.a, b0 = sl.b0;
      const float* const bias_u = bias_s + st_bias * p.bias_n;
      // ---- X0 = [idx_emb | obs_a | 0]: two batches of 8 rows per warp, each one DRAM round trip ----
      if (u > u0) mbar_wait(store_done, par ^ 1);                    // previous unit's last activation store has drained
      if (sl.kind != kStageNone) {
        stage_load(0); stage_store(0);
        stage_load(1); stage_store(1);
      }
      fence_proxy_async();
      mbar_arrive(act_ready + 0);

      // ---- hidden layers: accumulator -> +bias -> relu -> bf16 -> the next layer's A operand (in place) + HBM ----
      for (int l = 0; l + 1 < nl; ++l) {
        mbar_wait(mma_done + l, par);
        tc_fence_after();
        const int nch = p.N[l] >> 5;
        const float* bias = bias_u + p.tmem_col[l];
        uint8_t* const outb = act + p.in_box[l + 1] * kBoxBytes;
        for (int c = half; c < nch; c += 2) {
          uint32_t v[32];
          tmem_ld32(lane_addr + p.tmem_col[l] + c * 32, v);
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
        const float* bias = bias_u + p.tmem_col[l];
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
              const float4 bm = *reinterpret_cast<const float4*>(bias + j + k);
              const float4 bl = *reinterpret_cast<const float4*>(bias + L + j + k);
              mu[k] = __uint_as_float(vm[k]) + bm.x; mu[k + 1] = __uint_as_float(vm[k + 1]) + bm.y;
              mu[k + 2] = __uint_as_float(vm[k + 2]) + bm.z; mu[k + 3] = __uint_as_float(vm[k + 3]) + bm.w;
              lv[k] = __uint_as_float(vl[k]) + bl.x; lv[k + 1] = __uint_as_float(vl[k + 1]) + bl.y;
              lv[k + 2] = __uint_as_float(vl[k + 2]) + bl.z; lv[k + 3] = __uint_as_float(vl[k + 3]) + bl.w;
              *reinterpret_cast<float4*>(latrow + j + k) = make_float4(mu[k], mu[k + 1], mu[k + 2], mu[k + 3]);
              *reinterpret_cast<float4*>(latrow + L + j + k) = make_float4(lv[k], lv[k + 1], lv[k + 2], lv[k + 3]);
            }
            float z[16];
            float4 e4[4];                                            // branch hoisted: the four Philox calls interleave
            if (p.eps) {
#pragma unroll
              for (int k = 0; k < 4; ++k) e4[k] = *reinterpret_cast<const float4*>(p.eps + b * p.eps_ld + a * L + j + 4 * k);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                e4[k] = philox_normal4(p.seed, p.step, static_cast<uint64_t>(p.sample0 + b), static_cast<uint32_t>((a * L + j + 4 * k) >> 2));
            }
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
              const float ev[4] = {e4[k >> 2].x, e4[k >> 2].y, e4[k >> 2].z, e4[k >> 2].w};
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
  if (kActBoxes * kBoxBytes + wbytes + 256 + 2 * cols * sizeof(float) + 1024 > kEncSmemLimit) return false;
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
    p.xout[l] = static_cast<__nv_bfloat16*>(d.X[l]); p.xout_gs[l] = d.x_gs[l]; p.xout_ld[l] = d.x_ld[l]; p.xout_w[l] = K;
    if (rc == 0) rc = encode_tmap_bf16_3d(&pl->fmaps.x[l], d.X[l], K, d.B, d.A, d.x_ld[l], d.x_gs[l], 64, kEncRows);
    // activation boxes: X0 starts at box 0; a hidden activation goes next to the tile whose TMA store may still be in
    // flight when it is written, or back to box 0 (then the control thread drains that store first)
    if (l == 0) p.in_box[0] = 0;
    if (l + 1 < d.nl) {
      const int in_end = p.in_box[l] + p.kboxes[l];                         // boxes a pending store may be reading
      const int nb_out = N / 64;
      if (in_end + nb_out <= kActBoxes) { p.in_box[l + 1] = in_end; p.wait_store[l] = 0; }
      else { p.in_box[l + 1] = 0; p.wait_store[l] = 1; }
    } else {
      p.wait_store[l] = 0;
    }
  }
  if (rc != 0) { delete pl; return rc; }
  p.w_bytes = woff; p.bias_n = col;
  p.obs_off = d.obs_off; p.obs_dim = d.obs_dim; p.idx_emb = d.idx_emb; p.I = d.I;
  p.lat = d.lat; p.lat_gs = d.lat_gs; p.lat_ld = d.lat_ld;
  p.zin = static_cast<__nv_bfloat16*>(d.zin); p.zin_ld = d.zin_ld; p.L = d.L;
  pl->grid = std::min(p.total_units, kNumSMs);
  pl->fwd_smem = static_cast<size_t>(kActBoxes) * kBoxBytes + woff + 256 + 2 * col * sizeof(float) + 1024;
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
