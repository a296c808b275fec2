// ring.cu — HBM-resident replay ring: the device-side replacement for the sampling half of
// cpprb.ReplayBuffer (torch_ver/src/replay_buffer.py:83,102,108) and flashbax's item buffer
// (jax_ver/jax_buffer.py:86-91,132-140).  One row = one joint transition
//   [ obs(S) | act(W) | next_obs(S) | rew(A) | terminals(A) | truncations(A) | mask(1) | pad ]   fp32, row stride % 4 == 0
// -- every key of the reference's cpprb env_dict (replay_buffer.py:62-81): W = A for float-coded discrete actions, sum of the
// agents' action widths for continuous ones.  The first four blocks are in the column order create_dataset
// (torch_ver/trainer.py:7-45) produces, so a sampled batch is gathered straight into the packed matrices the step
// consumes (no host staging on the path); the flags are gathered only on request.
#include <algorithm>

#include "kernels.h"

struct MfvaeRing_ {
  int S = 0, A = 0, W = 0;
  int64_t capacity = 0, row = 0;
  float* storage = nullptr;
  int64_t head = 0, size = 0;      // host-side cursor: next write slot, number of valid rows
};

namespace mfvae {

__global__ void __launch_bounds__(256) ring_gather_kernel(const float* __restrict__ storage, int64_t row, int64_t size,
                                                          int S, int A, int W, int64_t batch, uint64_t seed, uint64_t step,
                                                          float* __restrict__ obs, float* __restrict__ act,
                                                          float* __restrict__ next, float* __restrict__ rew,
                                                          float* __restrict__ flags, int32_t* __restrict__ indices) {
  // one CTA per sampled row (grid-stride); uniform-with-replacement index from Philox(seed; counter = (i, step))
  for (int64_t i = blockIdx.x; i < batch; i += gridDim.x) {
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), 0x52494E47u,
                                             static_cast<uint32_t>(step)),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    const int64_t src = static_cast<int64_t>((static_cast<unsigned long long>(r.x) * static_cast<unsigned long long>(size)) >> 32);
    if (indices && threadIdx.x == 0) indices[i] = static_cast<int32_t>(src);
    const float* p = storage + src * row;
    const bool vec = (S % 4 == 0) && (A % 4 == 0) && (W % 4 == 0);
    if (vec) {
      const int sq = S / 4, aq = A / 4, wq = W / 4;
      for (int c = threadIdx.x; c < sq; c += blockDim.x) {
        reinterpret_cast<float4*>(obs + i * S)[c] = ldg_stream4(p + c * 4);
        reinterpret_cast<float4*>(next + i * S)[c] = ldg_stream4(p + S + W + c * 4);
      }
      for (int c = threadIdx.x; c < wq; c += blockDim.x) reinterpret_cast<float4*>(act + i * W)[c] = ldg_stream4(p + S + c * 4);
      for (int c = threadIdx.x; c < aq; c += blockDim.x) reinterpret_cast<float4*>(rew + i * A)[c] = ldg_stream4(p + 2 * S + W + c * 4);
    } else {
      for (int c = threadIdx.x; c < S; c += blockDim.x) { obs[i * S + c] = p[c]; next[i * S + c] = p[S + W + c]; }
      for (int c = threadIdx.x; c < W; c += blockDim.x) act[i * W + c] = p[S + c];
      for (int c = threadIdx.x; c < A; c += blockDim.x) rew[i * A + c] = p[2 * S + W + c];
    }
    if (flags) for (int c = threadIdx.x; c < 2 * A + 1; c += blockDim.x) flags[i * (2 * A + 1) + c] = p[2 * S + W + A + c];
  }
}

}  // namespace mfvae

using namespace mfvae;

extern "C" {

int64_t mfvae_ring_row_floats(int32_t state_dim, int32_t n_agents, int32_t act_cols) {
  return round_up(2LL * state_dim + act_cols + 3LL * n_agents + 1, 4);
}

int mfvae_ring_create(int32_t state_dim, int32_t n_agents, int32_t act_cols, int64_t capacity, float* d_storage, MfvaeRing* out) {
  MFVAE_CHECK(out && d_storage, "null argument");
  MFVAE_CHECK(state_dim >= 1 && n_agents >= 1 && act_cols >= n_agents && capacity >= 1, "ring dimensions must be positive (act_cols >= n_agents)");
  MFVAE_CHECK(reinterpret_cast<uintptr_t>(d_storage) % 16 == 0, "ring storage must be 16-byte aligned");
  MfvaeRing_* r = new MfvaeRing_();
  r->S = state_dim; r->A = n_agents; r->W = act_cols; r->capacity = capacity; r->row = mfvae_ring_row_floats(state_dim, n_agents, act_cols);
  r->storage = d_storage;
  *out = r;
  return 0;
}

int mfvae_ring_destroy(MfvaeRing r) { delete r; return 0; }
int64_t mfvae_ring_size(MfvaeRing r) { return r ? r->size : -1; }

int mfvae_ring_add(MfvaeRing r, const float* rows, int64_t n, int32_t rows_on_device, void* stream) {
  MFVAE_CHECK(r && rows, "null argument");
  MFVAE_CHECK(n >= 0, "negative row count");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const cudaMemcpyKind kind = rows_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (n > r->capacity) { rows += (n - r->capacity) * r->row; n = r->capacity; }   // only the newest survive
  int64_t done = 0;
  while (done < n) {
    const int64_t chunk = std::min(n - done, r->capacity - r->head);
    MFVAE_CUDA(cudaMemcpyAsync(r->storage + r->head * r->row, rows + done * r->row,
                               static_cast<size_t>(chunk * r->row) * sizeof(float), kind, s));
    r->head = (r->head + chunk) % r->capacity;
    done += chunk;
  }
  r->size = std::min(r->capacity, r->size + n);
  return 0;
}

int mfvae_ring_sample(MfvaeRing r, int64_t batch, uint64_t seed, uint64_t step, float* d_obs, float* d_act,
                      float* d_next, float* d_rew, float* d_flags_or_null, int32_t* d_indices_or_null, void* stream) {
  MFVAE_CHECK(r && d_obs && d_act && d_next && d_rew, "null argument");
  MFVAE_CHECK(r->size > 0, "cannot sample from an empty ring");
  MFVAE_CHECK(batch >= 1, "batch must be positive");
  const int grid = static_cast<int>(std::min<int64_t>(batch, static_cast<int64_t>(kNumSMs) * 16));
  ring_gather_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(r->storage, r->row, r->size, r->S, r->A, r->W, batch,
                                                                         seed, step, d_obs, d_act, d_next, d_rew, d_flags_or_null,
                                                                         d_indices_or_null);
  MFVAE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
