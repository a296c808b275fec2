// elementwise.cu — the HBM-bound kernels of the MAVAE train step (sm_100a).
//
//   stage            trainer.py:7-45 (create_dataset) + model.py:142-146 (id / action embedding gathers)
//   reparam_kl_fwd   model.py:77-81 (reparameterize) + model.py:35-37 (analytic Gaussian KL)
//   recon_loss       model.py:25-34 (Huber / MSE mean) forward value and d/d recon in one pass
//   reparam_kl_bwd   closed-form backward of the two above w.r.t. (mu, logvar)
//   colsum           bias gradients
//   adam             torch.optim.Adam single-pass fused update (+ bf16 shadow write)
//
// All of them move each byte once: 128-bit global accesses, grid sized to a multiple of the 148 SMs,
// grid-stride loops, warp-shuffle -> shared -> one partial per CTA, and a last-CTA-done ticket so the
// final scalar is produced in the same launch in a fixed (deterministic) order.
#include <algorithm>

#include "kernels.h"

namespace mfvae {

constexpr int kThreads = 256;
static inline int grid_for(int64_t work_items, int per_sm = 8) {
  int64_t blocks = (work_items + kThreads - 1) / kThreads;
  int64_t cap = static_cast<int64_t>(kNumSMs) * per_sm;     // multiple of the SM count
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ---------------------------------------------------------------------------------------------
// stage: packed fp32 rows -> encoder input X0[a][b][:] = [idx_emb[id] | obs_a | 0] and the action-embedding
// half of the decoder input.
// ---------------------------------------------------------------------------------------------
// grid = (row chunks, A); block = 256 threads = (256 / VL) rows x VL column strips of 4.  A thread keeps its strip and
// walks the rows four at a time: the strip's role (embedding / interior of the observation / edge) and the load width its
// alignment allows are fixed per thread, so the four rows' loads are straight-line and in flight together.
// 4 consecutive input columns starting at `src`, by the widest access its alignment class `mode` allows
// (fp32: 3 = 16 B, 2 = 8 B, 1 = 4 B;  bf16: 3 = 8 B, 2 = 4 B, 1 = 2 B)
template <typename TI> __device__ __forceinline__ void load_in4(const TI* src, int mode, float (&v)[4]);
template <> __device__ __forceinline__ void load_in4<float>(const float* src, int mode, float (&v)[4]) {
  if (mode == 3) {
    const float4 t = ldg_stream4(src);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if (mode == 2) {
    const float2 t0 = __ldg(reinterpret_cast<const float2*>(src)), t1 = __ldg(reinterpret_cast<const float2*>(src) + 1);
    v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
  } else {
    v[0] = __ldg(src); v[1] = __ldg(src + 1); v[2] = __ldg(src + 2); v[3] = __ldg(src + 3);
  }
}
template <> __device__ __forceinline__ void load_in4<__nv_bfloat16>(const __nv_bfloat16* src, int mode, float (&v)[4]) {
  if (mode == 3) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(src));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else if (mode == 2) {
    const uint32_t r0 = __ldg(reinterpret_cast<const uint32_t*>(src)), r1 = __ldg(reinterpret_cast<const uint32_t*>(src) + 1);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r0)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r1));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __bfloat162float(src[k]);
  }
}
template <typename TI> __device__ __forceinline__ float load_in1(const TI* p);
template <> __device__ __forceinline__ float load_in1<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_in1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T, typename TI>
__global__ void __launch_bounds__(kThreads) stage_kernel(StageArgs p) {
  const int a = blockIdx.y;
  const int nvec = p.x0_ld / 4;
  const int VL = nvec <= 32 ? 32 : (nvec <= 64 ? 64 : 128);      // strips handled per row pass
  const int rows_per_pass = kThreads / VL;
  const int strip0 = threadIdx.x % VL, rphase = threadIdx.x / VL;
  const int od = p.obs_dim[a], off = p.obs_off[a];
  T* x0 = static_cast<T*>(p.x0) + a * p.x0_gs;
  const int rstep = gridDim.x * rows_per_pass;
  constexpr int U = 4;
  for (int strip = strip0; strip < nvec; strip += VL) {
    const int c0 = strip * 4;
    const bool interior = c0 >= p.I && c0 + 3 < p.I + od;
    const TI* src0 = static_cast<const TI*>(p.obs) + off + (c0 - p.I);
    // row pitch and base decide the widest aligned load for every row of this strip
    const uintptr_t al = reinterpret_cast<uintptr_t>(src0) | (static_cast<uintptr_t>(p.obs_ld) * sizeof(TI));
    constexpr uintptr_t W = 4 * sizeof(TI);                      // bytes of 4 input columns
    const int mode = !interior ? 0 : ((al & (W - 1)) == 0 ? 3 : ((al & (W / 2 - 1)) == 0 ? 2 : 1));
    for (int b = blockIdx.x * rows_per_pass + rphase; b < p.B; b += U * rstep) {
      float v[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int bb = min(b + u * rstep, p.B - 1);                 // clamped rows are loaded twice, stored once
        const TI* src = src0 + static_cast<int64_t>(bb) * p.obs_ld;
        if (mode != 0) {
          load_in4<TI>(src, mode, v[u]);
        } else {
          int id = a;
          if (p.idx) id = max(0, min(static_cast<int>(p.idx[static_cast<int64_t>(bb) * p.A + a]), p.A - 1));   // same clamp as the scatter of backward
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = c0 + k;
            float x = 0.f;
            if (c < p.I) x = p.idx_emb[static_cast<int64_t>(id) * p.I + c];
            else if (c < p.I + od) x = load_in1<TI>(src + k);
            v[u][k] = x;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int bb = b + u * rstep;
        if (bb < p.B) store4<T>(x0 + static_cast<int64_t>(bb) * p.x0_ld + c0, make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) act_embed_kernel(StageArgs p) {
  const int cvec = p.C / 4;
  const int64_t total = static_cast<int64_t>(p.B) * p.A * cvec;
  T* zin = static_cast<T*>(p.zin);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cq = static_cast<int>(i % cvec);
    const int64_t r = i / cvec;
    const int a = static_cast<int>(r % p.A);
    const int64_t b = r / p.A;
    int act = static_cast<int>(p.act[b * p.act_ld + a]);
    act = max(0, min(act, p.n_act[a] - 1));
    const float4 e = *reinterpret_cast<const float4*>(p.act_table + a * p.act_table_gs +
                                                     static_cast<int64_t>(act) * p.C + cq * 4);
    store4<T>(zin + b * p.zin_ld + p.A * p.L + a * p.C + cq * 4, e);
  }
}

// continuous-action staging (reference model.py:148: actions[agent].to(device) feeds the ActionEncoder)
template <typename T>
__global__ void __launch_bounds__(kThreads) stage_actions_kernel(const float* __restrict__ act, int64_t act_ld,
                                                                 const int32_t* __restrict__ act_off, const int32_t* __restrict__ act_dim,
                                                                 T* __restrict__ act0, int A, int64_t B, int Kap) {
  const int64_t total = static_cast<int64_t>(A) * B * Kap;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % Kap);
    const int64_t r = i / Kap;
    const int64_t b = r % B;
    const int a = static_cast<int>(r / B);
    act0[i] = from_f<T>(k < act_dim[a] ? __ldg(act + b * act_ld + act_off[a] + k) : 0.f);
  }
}

int launch_stage_actions(const float* act, int64_t act_ld, const int32_t* act_off, const int32_t* act_dim, void* act0, int dtype,
                         int A, int64_t B, int Kap, cudaStream_t s) {
  const int grid = grid_for(static_cast<int64_t>(A) * B * Kap);
  if (dtype == kBF16) stage_actions_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(act, act_ld, act_off, act_dim, static_cast<__nv_bfloat16*>(act0), A, B, Kap);
  else                stage_actions_kernel<float><<<grid, kThreads, 0, s>>>(act, act_ld, act_off, act_dim, static_cast<float*>(act0), A, B, Kap);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

int launch_stage(const StageArgs& a, cudaStream_t s, bool do_x0, bool do_act) {
  MFVAE_CHECK(a.x0_ld % 4 == 0 && a.C % 4 == 0 && a.zin_ld % 4 == 0, "stage: widths must be multiples of 4");
  const int64_t t0 = static_cast<int64_t>(a.A) * a.B * (a.x0_ld / 4);
  const int64_t t1 = static_cast<int64_t>(a.A) * a.B * (a.C / 4);
  (void)t0;
  const int nvec = a.x0_ld / 4;
  const int rpp = kThreads / (nvec <= 32 ? 32 : (nvec <= 64 ? 64 : 128));
  const int chunks = std::max(1, std::min((a.B + rpp * 4 - 1) / (rpp * 4), std::max(1, kNumSMs * 16 / a.A)));
  dim3 sgrid(chunks, a.A);
  if (do_x0) {
    if (a.obs_dtype == kBF16) {
      if (a.dtype == kBF16) stage_kernel<__nv_bfloat16, __nv_bfloat16><<<sgrid, kThreads, 0, s>>>(a);
      else                  stage_kernel<float, __nv_bfloat16><<<sgrid, kThreads, 0, s>>>(a);
    } else {
      if (a.dtype == kBF16) stage_kernel<__nv_bfloat16, float><<<sgrid, kThreads, 0, s>>>(a);
      else                  stage_kernel<float, float><<<sgrid, kThreads, 0, s>>>(a);
    }
    MFVAE_LAUNCH_CHECK();
  }
  if (do_act) {
    if (a.dtype == kBF16) act_embed_kernel<__nv_bfloat16><<<grid_for(t1), kThreads, 0, s>>>(a);
    else                  act_embed_kernel<float><<<grid_for(t1), kThreads, 0, s>>>(a);
    MFVAE_LAUNCH_CHECK();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// reparameterize + KL.  One thread = 4 consecutive latent columns of one (sample, agent):
// 2 x 16-B loads (mu, logvar), one Philox4x32-10 call = 4 normals, one 16-B (fp32) / 8-B (bf16) store.
// Algorithmic bytes per sample: 3 * A*L * 4 (fp32 z) = SURVEY 8(d).
// ---------------------------------------------------------------------------------------------
// grid = (chunks, A): blockIdx.y is the agent, so the only index arithmetic per 4-element item is one 32-bit
// shift (L / 4 a power of two) or division -- the integer bookkeeping used to outweigh the Philox rounds.
__device__ __forceinline__ void split_item(uint32_t e, uint32_t lq, int lq_shift, uint32_t& b, uint32_t& jq) {
  if (lq_shift >= 0) { b = e >> lq_shift; jq = e & (lq - 1); }
  else               { b = e / lq; jq = e - b * lq; }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) reparam_kl_fwd_kernel(ReparamArgs p, uint32_t lq, int lq_shift, uint32_t items) {
  __shared__ float red[32];
  const int a = blockIdx.y;
  T* z = static_cast<T*>(p.z) + a * p.L;
  const float* mu_a = p.mu + a * p.lat_as;
  const float* lv_a = p.lv + a * p.lat_as;
  const float* eps_a = p.eps ? p.eps + a * p.L : nullptr;
  const uint32_t q0 = static_cast<uint32_t>(a * p.L) >> 2;           // Philox column-quad index of this agent's first column
  float acc = 0.f;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < items; e += gridDim.x * blockDim.x) {
    uint32_t b, jq;
    split_item(e, lq, lq_shift, b, jq);
    const int64_t off = static_cast<int64_t>(b) * p.lat_bs + jq * 4;
    const float4 mu = ldg_stream4(mu_a + off);
    const float4 lv = ldg_stream4(lv_a + off);
    float4 ep;
    if (eps_a) ep = ldg_stream4(eps_a + static_cast<int64_t>(b) * p.eps_ld + jq * 4);
    else       ep = philox_normal4(p.seed, p.step, static_cast<uint64_t>(p.sample0 + b), q0 + jq);
    float4 o;
    // sigma = exp(lv / 2); exp(lv) = sigma^2 (one transcendental per element instead of two; 1 ulp apart)
    const float sx = expf(0.5f * lv.x), sy = expf(0.5f * lv.y), sz = expf(0.5f * lv.z), sw = expf(0.5f * lv.w);
    const float ex = sx * sx, ey = sy * sy, ez = sz * sz, ew = sw * sw;
    o.x = mu.x + ep.x * sx;
    o.y = mu.y + ep.y * sy;
    o.z = mu.z + ep.z * sz;
    o.w = mu.w + ep.w * sw;
    store4<T>(z + static_cast<int64_t>(b) * p.z_ld + jq * 4, o);
    acc += (1.f + lv.x - mu.x * mu.x - ex) + (1.f + lv.y - mu.y * mu.y - ey) +
           (1.f + lv.z - mu.z * mu.z - ez) + (1.f + lv.w - mu.w * mu.w - ew);
  }
  const float tot = block_sum(acc, red);
  finish_scalar(tot, p.scratch, -0.5f * p.kl_scale, p.kl_out, red);
}

static int reparam_grid(int64_t B, int A, int L, dim3* grid, uint32_t* lq, int* lq_shift, uint32_t* items) {
  MFVAE_CHECK(B * (L / 4) < (1LL << 32), "reparam: batch * latent / 4 must fit 32 bits");
  *lq = static_cast<uint32_t>(L / 4);
  *lq_shift = -1;
  for (int sft = 0; sft < 31; ++sft) if ((1u << sft) == *lq) *lq_shift = sft;
  *items = static_cast<uint32_t>(B * (L / 4));
  const int per_agent = std::max(1, std::min<int>(grid_for(*items), std::max(1, kNumSMs * 8 / A)));
  MFVAE_CHECK(static_cast<int64_t>(per_agent) * A <= kMaxPartials, "reparam: too many blocks for the partial-sum scratch");
  *grid = dim3(per_agent, A);
  return 0;
}

int launch_reparam_kl_fwd(const ReparamArgs& a, cudaStream_t s) {
  MFVAE_CHECK(a.L % 4 == 0, "reparam: latent must be a multiple of 4");
  MFVAE_CHECK(a.lat_as % 4 == 0 && a.lat_bs % 4 == 0 && a.z_ld % 4 == 0, "reparam: strides must be multiples of 4");
  MFVAE_CHECK(a.A <= 65535, "reparam: too many agents for one launch");
  dim3 grid; uint32_t lq, items; int sh;
  MFVAE_TRY(reparam_grid(a.B, a.A, a.L, &grid, &lq, &sh, &items));
  if (a.z_dtype == kBF16) reparam_kl_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(a, lq, sh, items);
  else                    reparam_kl_fwd_kernel<float><<<grid, kThreads, 0, s>>>(a, lq, sh, items);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

template <typename TG, typename TD>
__global__ void __launch_bounds__(kThreads) reparam_kl_bwd_kernel(ReparamBwdArgs p, uint32_t lq, int lq_shift, uint32_t items) {
  const int a = blockIdx.y;
  const TG* gz = static_cast<const TG*>(p.gz) + a * p.L;
  TD* dl = static_cast<TD*>(p.dlat) + a * p.dlat_as;
  const float* mu_a = p.mu + a * p.lat_as;
  const float* lv_a = p.lv + a * p.lat_as;
  const float* eps_a = p.eps ? p.eps + a * p.L : nullptr;
  const float* glat_a = p.glat ? p.glat + a * p.lat_as : nullptr;
  const uint32_t q0 = static_cast<uint32_t>(a * p.L) >> 2;
  const float k = p.kl_scale;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < items; e += gridDim.x * blockDim.x) {
    uint32_t b, jq;
    split_item(e, lq, lq_shift, b, jq);
    const int64_t off = static_cast<int64_t>(b) * p.lat_bs + jq * 4;
    const float4 mu = *reinterpret_cast<const float4*>(mu_a + off);
    const float4 lv = *reinterpret_cast<const float4*>(lv_a + off);
    const float4 g = load4<TG>(gz + static_cast<int64_t>(b) * p.gz_ld + jq * 4);
    float4 ep;
    if (eps_a) ep = *reinterpret_cast<const float4*>(eps_a + static_cast<int64_t>(b) * p.eps_ld + jq * 4);
    else       ep = philox_normal4(p.seed, p.step, static_cast<uint64_t>(p.sample0 + b), q0 + jq);
    float4 dmu, dlv;
    dmu.x = g.x + k * mu.x; dmu.y = g.y + k * mu.y; dmu.z = g.z + k * mu.z; dmu.w = g.w + k * mu.w;
    const float sx = expf(0.5f * lv.x), sy = expf(0.5f * lv.y), sz = expf(0.5f * lv.z), sw = expf(0.5f * lv.w);
    dlv.x = g.x * ep.x * 0.5f * sx + k * 0.5f * (sx * sx - 1.f);
    dlv.y = g.y * ep.y * 0.5f * sy + k * 0.5f * (sy * sy - 1.f);
    dlv.z = g.z * ep.z * 0.5f * sz + k * 0.5f * (sz * sz - 1.f);
    dlv.w = g.w * ep.w * 0.5f * sw + k * 0.5f * (sw * sw - 1.f);
    if (glat_a) {
      const float4 um = *reinterpret_cast<const float4*>(glat_a + off);
      const float4 ul = *reinterpret_cast<const float4*>(glat_a + off + p.L);
      dmu.x += um.x; dmu.y += um.y; dmu.z += um.z; dmu.w += um.w;
      dlv.x += ul.x; dlv.y += ul.y; dlv.z += ul.z; dlv.w += ul.w;
    }
    TD* o = dl + static_cast<int64_t>(b) * p.dlat_bs + jq * 4;
    store4<TD>(o, dmu);
    store4<TD>(o + p.L, dlv);
  }
}

int launch_reparam_kl_bwd(const ReparamBwdArgs& a, cudaStream_t s) {
  MFVAE_CHECK(a.L % 4 == 0 && a.gz_ld % 4 == 0 && a.dlat_bs % 4 == 0, "reparam bwd: widths must be multiples of 4");
  MFVAE_CHECK(a.g_dtype == a.d_dtype, "reparam bwd: mixed dtypes unsupported");
  MFVAE_CHECK(a.A <= 65535, "reparam bwd: too many agents for one launch");
  dim3 grid; uint32_t lq, items; int sh;
  MFVAE_TRY(reparam_grid(a.B, a.A, a.L, &grid, &lq, &sh, &items));
  if (a.g_dtype == kBF16) reparam_kl_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, kThreads, 0, s>>>(a, lq, sh, items);
  else                    reparam_kl_bwd_kernel<float, float><<<grid, kThreads, 0, s>>>(a, lq, sh, items);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// reconstruction loss, forward value + gradient in one pass.
// Reference argument order F.huber_loss(target_data, recon) is symmetric in its two arguments.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void recon_elem(float r, float t, int huber, float gs, float& val, float& g) {
  const float d = r - t;
  if (huber) {
    const float ad = fabsf(d);
    val = ad < 1.f ? 0.5f * d * d : ad - 0.5f;
    g = fminf(fmaxf(d, -1.f), 1.f) * gs;
  } else {
    val = d * d;
    g = 2.f * d * gs;
  }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) recon_loss_kernel(ReconLossArgs p) {
  __shared__ float red[32];
  T* grad = static_cast<T*>(p.grad);
  float acc = 0.f;
  if (VEC) {
    const int wq = p.width / 4;
    const int64_t total = p.B * wq;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t b = i / wq;
      const int c = static_cast<int>(i - b * wq) * 4;
      const float4 r = p.recon16 ? load4<__nv_bfloat16>(p.recon16 + b * p.grad_ld + c) : ldg_stream4(p.recon + b * p.recon_ld + c);
      const float4 t = p.target16 ? load4<__nv_bfloat16>(p.target16 + b * p.target_ld + c) : ldg_stream4(p.target + b * p.target_ld + c);
      float4 g; float v0, v1, v2, v3;
      recon_elem(r.x, t.x, p.huber, p.grad_scale, v0, g.x);
      recon_elem(r.y, t.y, p.huber, p.grad_scale, v1, g.y);
      recon_elem(r.z, t.z, p.huber, p.grad_scale, v2, g.z);
      recon_elem(r.w, t.w, p.huber, p.grad_scale, v3, g.w);
      acc += (v0 + v1) + (v2 + v3);
      if (grad) store4<T>(grad + b * p.grad_ld + c, g);
    }
  } else {
    const int64_t total = p.B * p.width;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t b = i / p.width;
      const int c = static_cast<int>(i - b * p.width);
      float v, g;
      const float r = p.recon16 ? __bfloat162float(p.recon16[b * p.grad_ld + c]) : p.recon[b * p.recon_ld + c];
      const float t = p.target16 ? __bfloat162float(p.target16[b * p.target_ld + c]) : p.target[b * p.target_ld + c];
      recon_elem(r, t, p.huber, p.grad_scale, v, g);
      acc += v;
      if (grad) grad[b * p.grad_ld + c] = from_f<T>(g);
    }
  }
  const float tot = block_sum(acc, red);
  finish_scalar(tot, p.scratch, p.loss_scale, p.loss_out, red);
}

// Same pass, with the column sums of the gradient (= the bias gradient of the layer that produced `recon`) accumulated on the
// way: a thread owns one 4-column strip and walks rows (blockIdx.y, += gridDim.y), so its column sums stay in registers and
// leave with one fp32 atomic per column per CTA row-chunk.  The value summed is the STORED gradient (bf16-rounded when the
// gradient buffer is bf16), exactly what the separate column-sum kernel read back from HBM.
template <typename T>
__global__ void __launch_bounds__(kThreads) recon_loss_colsum_kernel(ReconLossArgs p) {
  __shared__ float red[32];
  T* grad = static_cast<T*>(p.grad);
  const int wq = p.width / 4;
  const int cq = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cq < wq) {
    const int c = cq * 4;
    constexpr int U = 4;
    for (int64_t b0 = blockIdx.y; b0 < p.B; b0 += static_cast<int64_t>(gridDim.y) * U) {
      float4 r[U], t[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t b = min(b0 + static_cast<int64_t>(u) * gridDim.y, p.B - 1);      // clamped rows are loaded twice, used once
        r[u] = p.recon16 ? load4<__nv_bfloat16>(p.recon16 + b * p.grad_ld + c) : ldg_stream4(p.recon + b * p.recon_ld + c);
        t[u] = p.target16 ? load4<__nv_bfloat16>(p.target16 + b * p.target_ld + c) : ldg_stream4(p.target + b * p.target_ld + c);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t b = b0 + static_cast<int64_t>(u) * gridDim.y;
        if (b >= p.B) break;
        float4 g; float v0, v1, v2, v3;
        recon_elem(r[u].x, t[u].x, p.huber, p.grad_scale, v0, g.x);
        recon_elem(r[u].y, t[u].y, p.huber, p.grad_scale, v1, g.y);
        recon_elem(r[u].z, t[u].z, p.huber, p.grad_scale, v2, g.z);
        recon_elem(r[u].w, t[u].w, p.huber, p.grad_scale, v3, g.w);
        acc += (v0 + v1) + (v2 + v3);
        store4<T>(grad + b * p.grad_ld + c, g);
        cs.x += to_f<T>(from_f<T>(g.x)); cs.y += to_f<T>(from_f<T>(g.y)); cs.z += to_f<T>(from_f<T>(g.z)); cs.w += to_f<T>(from_f<T>(g.w));
      }
    }
    float* o = p.colsum_out + c;
    atomicAdd(o, cs.x); atomicAdd(o + 1, cs.y); atomicAdd(o + 2, cs.z); atomicAdd(o + 3, cs.w);
  }
  const float tot = block_sum(acc, red);
  finish_scalar(tot, p.scratch, p.loss_scale, p.loss_out, red);
}

int launch_recon_loss(const ReconLossArgs& a, cudaStream_t s) {
  const bool vec = a.width % 4 == 0 && (a.recon16 || a.recon_ld % 4 == 0) && a.target_ld % 4 == 0 && a.grad_ld % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(a.recon16 ? static_cast<const void*>(a.recon16) : static_cast<const void*>(a.recon)) % 16 == 0) && (reinterpret_cast<uintptr_t>(a.target16 ? static_cast<const void*>(a.target16) : static_cast<const void*>(a.target)) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(a.grad) % 16 == 0);
  if (a.colsum_out && vec && a.grad) {
    const int wq = a.width / 4;
    dim3 grid((wq + kThreads - 1) / kThreads, 1);
    grid.y = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((a.B + 3) / 4, (kNumSMs * 8 + grid.x - 1) / grid.x)));
    MFVAE_CHECK(static_cast<int64_t>(grid.x) * grid.y <= kMaxPartials, "recon loss: too many CTAs for the partial buffer");
    if (a.grad_dtype == kBF16) recon_loss_colsum_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(a);
    else                       recon_loss_colsum_kernel<float><<<grid, kThreads, 0, s>>>(a);
    MFVAE_LAUNCH_CHECK();
    return 0;
  }
  const int64_t items = vec ? a.B * (a.width / 4) : a.B * a.width;
  const int grid = grid_for(items);
  if (a.grad_dtype == kBF16) {
    if (vec) recon_loss_kernel<__nv_bfloat16, true><<<grid, kThreads, 0, s>>>(a);
    else     recon_loss_kernel<__nv_bfloat16, false><<<grid, kThreads, 0, s>>>(a);
  } else {
    if (vec) recon_loss_kernel<float, true><<<grid, kThreads, 0, s>>>(a);
    else     recon_loss_kernel<float, false><<<grid, kThreads, 0, s>>>(a);
  }
  MFVAE_LAUNCH_CHECK();
  // ragged widths / unaligned pointers: the column sum runs as its own pass
  if (a.colsum_out && a.grad) return launch_colsum(a.grad, a.grad_dtype, 1, a.B, a.width, a.grad_ld, 0, a.colsum_out, 0, s);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// column sums (bias gradients): out[g][n] += sum_b X[g][b][n]
// One thread owns a 4-column strip (8-byte bf16 / 16-byte fp32 loads); a warp covers LPR strips of one row and
// 32 / LPR consecutive rows per load instruction, 4 independent loads in flight per thread.  Partial sums are folded
// across the CTA in shared memory and leave with one fp32 atomic per column per CTA.
// grid = (column tiles of 4 * LPR, row splits, groups); block = 256.
// ---------------------------------------------------------------------------------------------
template <typename T, int LPR>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int64_t B, int N, int64_t ld,
                                                     int64_t gs, float* __restrict__ out, int64_t out_gs) {
  constexpr int RPB = 256 / LPR;                  // rows covered by the CTA per pass
  __shared__ float4 sm[256];
  const int g = blockIdx.z;
  const int strip = threadIdx.x % LPR, rphase = threadIdx.x / LPR;
  const int c = (blockIdx.x * LPR + strip) * 4;
  const T* xg = x + g * gs;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < N) {
    const int64_t stride = static_cast<int64_t>(gridDim.y) * RPB;
    int64_t b = static_cast<int64_t>(blockIdx.y) * RPB + rphase;
    for (; b + 3 * stride < B; b += 4 * stride) {
      const float4 v0 = load4<T>(xg + b * ld + c);
      const float4 v1 = load4<T>(xg + (b + stride) * ld + c);
      const float4 v2 = load4<T>(xg + (b + 2 * stride) * ld + c);
      const float4 v3 = load4<T>(xg + (b + 3 * stride) * ld + c);
      acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; b < B; b += stride) {
      const float4 v = load4<T>(xg + b * ld + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  if (rphase == 0 && c < N) {
#pragma unroll 4
    for (int r = 1; r < RPB; ++r) {
      const float4 v = sm[r * LPR + strip];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* o = out + g * out_gs + c;
    atomicAdd(o, acc.x);
    if (c + 1 < N) atomicAdd(o + 1, acc.y);
    if (c + 2 < N) atomicAdd(o + 2, acc.z);
    if (c + 3 < N) atomicAdd(o + 3, acc.w);
  }
}

static int kColsumPerSm = 4;
template <typename T>
static int colsum_dispatch(const T* x, int G, int64_t B, int N, int64_t ld, int64_t gs, float* out, int64_t out_gs, cudaStream_t s) {
  // columns are read 4 at a time: the padded row (ld, a multiple of 8) always holds whole strips
  const int strips = (N + 3) / 4;
  const int lpr = strips >= 32 ? 32 : (strips > 8 ? 16 : 8);
  const int ctiles = (strips + lpr - 1) / lpr;
  const int rpb = 256 / lpr;
  // (two CTAs per SM instead of four, to leave thread slots to the GEMMs they run beside, was measured: 0.823 -> 0.839 ms)
  const int64_t want = (static_cast<int64_t>(kNumSMs) * kColsumPerSm) / std::max<int64_t>(1, static_cast<int64_t>(ctiles) * G);
  const int splits = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, (B + 4 * rpb - 1) / (4 * rpb))));
  dim3 grid(ctiles, splits, G);
  if (lpr == 32) colsum_kernel<T, 32><<<grid, 256, 0, s>>>(x, B, N, ld, gs, out, out_gs);
  else if (lpr == 16) colsum_kernel<T, 16><<<grid, 256, 0, s>>>(x, B, N, ld, gs, out, out_gs);
  else colsum_kernel<T, 8><<<grid, 256, 0, s>>>(x, B, N, ld, gs, out, out_gs);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

int launch_colsum(const void* x, int dtype, int G, int64_t B, int N, int64_t ld, int64_t gs,
                  float* out, int64_t out_gs, cudaStream_t s) {
  MFVAE_CHECK(ld % 4 == 0 && gs % 4 == 0 && ld >= (N + 3) / 4 * 4, "colsum: rows must be padded to whole 4-column strips");
  if (dtype == kBF16) return colsum_dispatch(static_cast<const __nv_bfloat16*>(x), G, B, N, ld, gs, out, out_gs, s);
  return colsum_dispatch(static_cast<const float*>(x), G, B, N, ld, gs, out, out_gs, s);
}

// ---------------------------------------------------------------------------------------------
// embedding gradients (scatter-add; tables are tiny, rows are many -> shared-memory pre-reduction)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) idx_emb_scatter_kernel(const T* __restrict__ gx0, int64_t gs, int64_t ld,
                                                                   const float* __restrict__ idx, int idx_ld, int A, int I,
                                                                   int64_t B, float* __restrict__ d_emb) {
  const int64_t total = static_cast<int64_t>(A) * B * I;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % I);
    const int64_t r = i / I;
    const int64_t b = r % B;
    const int a = static_cast<int>(r / B);
    int id = static_cast<int>(idx[b * idx_ld + a]);
    id = max(0, min(id, A - 1));
    atomicAdd(d_emb + static_cast<int64_t>(id) * I + c, to_f<T>(gx0[a * gs + b * ld + c]));
  }
}

int launch_idx_emb_scatter(const void* gx0, int dtype, int64_t gs, int64_t ld, const float* idx, int idx_ld,
                           int A, int I, int64_t B, float* d_emb, cudaStream_t s) {
  const int grid = grid_for(static_cast<int64_t>(A) * B * I);
  if (dtype == kBF16)
    idx_emb_scatter_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(gx0), gs, ld, idx, idx_ld, A, I, B, d_emb);
  else
    idx_emb_scatter_kernel<float><<<grid, kThreads, 0, s>>>(static_cast<const float*>(gx0), gs, ld, idx, idx_ld, A, I, B, d_emb);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

// grid = (row chunks, A); each CTA reduces its rows into smem[n_act][C] and flushes with atomics.
template <typename T>
__global__ void __launch_bounds__(kThreads) act_table_grad_kernel(const T* __restrict__ gzin, int64_t ld, int col0,
                                                                  const float* __restrict__ act, int act_ld,
                                                                  const int32_t* __restrict__ n_act, int C, int64_t B,
                                                                  float* __restrict__ d_table, int64_t table_gs) {
  extern __shared__ float acc[];
  const int a = blockIdx.y;
  const int na = n_act[a];
  for (int i = threadIdx.x; i < na * C; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int rows_per_iter = blockDim.x / C > 0 ? blockDim.x / C : 1;
  const int c = threadIdx.x % C;
  const int rphase = threadIdx.x / C;
  if (rphase < rows_per_iter) {
    for (int64_t b = static_cast<int64_t>(blockIdx.x) * rows_per_iter + rphase; b < B;
         b += static_cast<int64_t>(gridDim.x) * rows_per_iter) {
      int k = static_cast<int>(act[b * act_ld + a]);
      k = max(0, min(k, na - 1));
      atomicAdd(acc + k * C + c, to_f<T>(gzin[b * ld + col0 + a * C + c]));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < na * C; i += blockDim.x) atomicAdd(d_table + a * table_gs + i, acc[i]);
}

// Few actions (Discrete(5) in simple_tag): every thread owns 4 columns of one agent and keeps one running sum per
// action in registers (predicated adds, no shared-memory atomics); rows are strided over (blockIdx.x, row phase).
// grid = (row chunks, A); block = 256 = (C / 4 strips) x (256 / (C / 4) row phases)
template <typename T, int NA>
__global__ void __launch_bounds__(kThreads) act_table_grad_small_kernel(const T* __restrict__ gzin, int64_t ld, int col0,
                                                                        const float* __restrict__ act, int act_ld,
                                                                        const int32_t* __restrict__ n_act, int C, int64_t B,
                                                                        float* __restrict__ d_table, int64_t table_gs) {
  const int a = blockIdx.y;
  const int strips = C / 4;
  const int strip = threadIdx.x % strips, rphase = threadIdx.x / strips;
  const int rpb = blockDim.x / strips;
  const int na = n_act[a];
  float4 acc[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rphase < rpb) {
    for (int64_t b = static_cast<int64_t>(blockIdx.x) * rpb + rphase; b < B; b += static_cast<int64_t>(gridDim.x) * rpb) {
      int kk = static_cast<int>(__ldg(act + b * act_ld + a));
      kk = max(0, min(kk, na - 1));
      const float4 v = load4<T>(gzin + b * ld + col0 + a * C + strip * 4);
#pragma unroll
      for (int k = 0; k < NA; ++k) {
        const float m = (kk == k) ? 1.f : 0.f;
        acc[k].x = fmaf(m, v.x, acc[k].x); acc[k].y = fmaf(m, v.y, acc[k].y);
        acc[k].z = fmaf(m, v.z, acc[k].z); acc[k].w = fmaf(m, v.w, acc[k].w);
      }
    }
  }
  // fold every row phase of the CTA that shares a strip through shared memory, then one atomic per (action, column)
  __shared__ float4 fold[kThreads];
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    if (k >= na) break;                     // uniform across the CTA (one agent per CTA)
    __syncthreads();
    fold[threadIdx.x] = acc[k];
    __syncthreads();
    if (rphase == 0) {
      float4 t = fold[strip];
      for (int r = 1; r < rpb; ++r) {
        const float4 u = fold[r * strips + strip];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      float* o = d_table + a * table_gs + static_cast<int64_t>(k) * C + strip * 4;
      atomicAdd(o, t.x); atomicAdd(o + 1, t.y); atomicAdd(o + 2, t.z); atomicAdd(o + 3, t.w);
    }
  }
}

int launch_act_table_grad(const void* gzin, int dtype, int64_t ld, int col0, const float* act, int act_ld,
                          const int32_t* n_act, int A, int C, int64_t B, float* d_table, int64_t table_gs,
                          cudaStream_t s) {
  MFVAE_CHECK(C <= kThreads, "act_table_grad: act_features must be <= 256");
  const int na_max = static_cast<int>(table_gs / C);
  const int strips = C / 4;
  const bool small = na_max <= 8 && C % 4 == 0 && ld % 4 == 0 && col0 % 4 == 0 && (strips >= 32 ? strips % 32 == 0 : 32 % strips == 0) &&
                     strips <= kThreads;
  if (small) {
    const int rpb = kThreads / strips;
    int chunks = static_cast<int>(std::min<int64_t>((B + rpb * 8 - 1) / (rpb * 8), std::max(1, kNumSMs * 2 / A)));
    chunks = std::max(chunks, 1);
    dim3 grid(chunks, A);
    if (dtype == kBF16)
      act_table_grad_small_kernel<__nv_bfloat16, 8><<<grid, kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(gzin), ld, col0, act, act_ld, n_act, C, B, d_table, table_gs);
    else
      act_table_grad_small_kernel<float, 8><<<grid, kThreads, 0, s>>>(static_cast<const float*>(gzin), ld, col0, act, act_ld, n_act, C, B, d_table, table_gs);
    MFVAE_LAUNCH_CHECK();
    return 0;
  }
  MFVAE_CHECK(table_gs * sizeof(float) <= 48 * 1024, "act_table_grad: action table too large for the shared-memory path");
  const int rows_per_iter = kThreads / C;
  int chunks = static_cast<int>(std::min<int64_t>((B + rows_per_iter * 8 - 1) / (rows_per_iter * 8),
                                                   std::max(1, kNumSMs * 4 / A)));
  chunks = std::max(chunks, 1);
  dim3 grid(chunks, A);
  const size_t smem = static_cast<size_t>(table_gs) * sizeof(float);
  if (dtype == kBF16)
    act_table_grad_kernel<__nv_bfloat16><<<grid, kThreads, smem, s>>>(static_cast<const __nv_bfloat16*>(gzin), ld, col0, act, act_ld, n_act, C, B, d_table, table_gs);
  else
    act_table_grad_kernel<float><<<grid, kThreads, smem, s>>>(static_cast<const float*>(gzin), ld, col0, act, act_ld, n_act, C, B, d_table, table_gs);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// fused Adam: one pass, 28 B / parameter (+2 B when the bf16 shadow is refreshed in the same pass).
// Same operation order as torch.optim.Adam (single-tensor path): lerp, addcmul, sqrt / bc2_sqrt + eps, addcdiv.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        __nv_bfloat16* __restrict__ shadow, int64_t n4,
                                                        float one_minus_b1, float b2, float one_minus_b2,
                                                        float step_size, float bc2_sqrt, float eps) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = ldg_stream4(g + i * 4);
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
    if (shadow) store4<__nv_bfloat16>(shadow + i * 4, pp);
  }
}

int launch_adam(float* p, const float* g, float* m, float* v, __nv_bfloat16* shadow, int64_t n,
                float lr, float b1, float b2, float eps, int64_t t, cudaStream_t s, int blocks_per_sm) {
  MFVAE_CHECK(n % 4 == 0, "adam: element count must be a multiple of 4");
  MFVAE_CHECK(t >= 1, "adam: step count starts at 1");
  const double bc1 = 1.0 - pow(static_cast<double>(b1), static_cast<double>(t));
  const double bc2 = 1.0 - pow(static_cast<double>(b2), static_cast<double>(t));
  const float step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  const float bc2_sqrt = static_cast<float>(sqrt(bc2));
  adam_kernel<<<grid_for(n / 4, blocks_per_sm), kThreads, 0, s>>>(p, g, m, v, shadow, n / 4, 1.f - b1, b2, 1.f - b2,
                                                  step_size, bc2_sqrt, eps);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(kThreads) cast_bf16_kernel(const float* __restrict__ src,
                                                             __nv_bfloat16* __restrict__ dst, int64_t n4) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    store4<__nv_bfloat16>(dst + i * 4, ldg_stream4(src + i * 4));
}

int launch_cast_bf16(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t s) {
  MFVAE_CHECK(n % 4 == 0, "cast: element count must be a multiple of 4");
  cast_bf16_kernel<<<grid_for(n / 4), kThreads, 0, s>>>(src, dst, n / 4);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) cast2d_kernel(const float* __restrict__ src, int64_t src_ld, T* __restrict__ dst,
                                                          int64_t dst_ld, int64_t B, int width) {
  const int64_t total = B * width;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / width;
    const int c = static_cast<int>(i - b * width);
    dst[b * dst_ld + c] = from_f<T>(src ? src[b * src_ld + c] : 0.f);
  }
}

int launch_cast2d(const float* src, int64_t src_ld, void* dst, int64_t dst_ld, int dtype, int64_t B, int width, cudaStream_t s) {
  const int grid = grid_for(B * width);
  if (dtype == kBF16) cast2d_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(src, src_ld, static_cast<__nv_bfloat16*>(dst), dst_ld, B, width);
  else                cast2d_kernel<float><<<grid, kThreads, 0, s>>>(src, src_ld, static_cast<float*>(dst), dst_ld, B, width);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(kThreads) philox_normal_kernel(float* out, int64_t B, int wq, uint64_t seed,
                                                                 uint64_t step, int64_t sample0) {
  const int64_t total = B * wq;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / wq;
    const int q = static_cast<int>(i - b * wq);
    reinterpret_cast<float4*>(out)[i] = philox_normal4(seed, step, static_cast<uint64_t>(sample0 + b), q);
  }
}

int launch_philox_normal(float* out, int64_t B, int width, uint64_t seed, uint64_t step, int64_t sample0,
                         cudaStream_t s) {
  MFVAE_CHECK(width % 4 == 0, "philox_normal: width must be a multiple of 4");
  philox_normal_kernel<<<grid_for(B * (width / 4)), kThreads, 0, s>>>(out, B, width / 4, seed, step, sample0);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(kThreads) loss_total_kernel(float* losses, float s_weight, float r_weight, float kl_weight,
                                                              const float* partials, int n_partials, float partial_scale) {
  __shared__ float red[32];
  if (partials) {                         // fixed summation order: the value is reproducible run to run
    float v = 0.f;
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) v += __ldcg(partials + i);
    v = block_sum(v, red);
    if (threadIdx.x == 0) losses[1] = v * partial_scale;
  }
  if (threadIdx.x == 0) losses[0] = s_weight * losses[1] + r_weight * losses[2] + kl_weight * losses[3];
}
int launch_loss_total(float* losses, float s_weight, float r_weight, float kl_weight, cudaStream_t s, const float* partials, int n_partials,
                      float partial_scale) {
  loss_total_kernel<<<1, kThreads, 0, s>>>(losses, s_weight, r_weight, kl_weight, partials, n_partials, partial_scale);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mfvae
