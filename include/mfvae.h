/* mfvae.h — C ABI of the B200-native MF-VAE (class MAVAE) training hot path.
 *
 * Drop-in boundary for the reference path /root/reference/torch_ver:
 *   model.py:134-173   MAVAE.forward          -> mfvae_forward
 *   model.py:19-40     loss_s_r_vae_fn        -> mfvae_loss       (forward value + d loss / d recon)
 *   model.py:8-16      loss_vae_fn            -> mfvae_loss with joint_mse = 1
 *   main.py:92-93 / trainer.py:98-100  zero_grad + loss.backward() -> mfvae_backward
 *   main.py:97   / trainer.py:102-103  torch.optim.Adam.step()     -> mfvae_adam_step
 *   trainer.py:7-45    create_dataset (host numpy) -> device staging kernel inside the forward entry (MfvaeBatch)
 *   src/replay_buffer.py:53-115 (cpprb, C++) / jax_ver/jax_buffer.py:80-140 (flashbax)
 *                                             -> mfvae_ring_* (HBM-resident ring + gather)
 *
 * Conventions: every entry point returns 0 on success, non-zero on failure;
 * mfvae_last_error() returns a thread-local message.  No exceptions or aborts cross the ABI.
 * All pointers named d_* are device pointers owned by the caller (the Python host allocates them as
 * torch tensors); the library never frees caller memory.  `stream` is a cudaStream_t passed as void*.
 * A handle is not thread-safe (one per rank / thread), matching the reference's single-threaded use.
 */
#ifndef MFVAE_H_
#define MFVAE_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFVAE_MAX_HIDDEN 8

enum { MFVAE_PREC_FP32 = 0, MFVAE_PREC_BF16 = 1 };
/* GEMM engines.  AUTO = tcgen05 for bf16, SIMT FFMA for fp32.  SIMT with bf16 is a debugging aid that
 * runs the same data flow without tensor cores; it is NOT a fallback: nothing selects it implicitly. */
enum { MFVAE_ENGINE_AUTO = 0, MFVAE_ENGINE_SIMT = 1, MFVAE_ENGINE_TCGEN05 = 2 };
/* Cross-layer fusions of the tcgen05 engine (bit mask).  All variants compute the same values with the same rounding
 * points; tests/test_gpu_fused.py pins them against each other.
 *   NONE     one kernel per layer
 *   ENCODER  the per-agent encoder chain + reparameterisation + KL as ONE persistent tcgen05 kernel on the staged [obs | 0]
 *            tile (enc_fused.cu; per-layer kernels are used when its shape constraints do not hold, or when the batch carries
 *            an explicit per-row agent-index column, which cannot be folded into layer 0's bias)
 *   LOSS     mfvae_fwd_bwd only: the state head's reconstruction loss and its gradient are the epilogue of the output-layer
 *            GEMM (recon_s is never written to HBM; MfvaeOutputs.d_recon_s is NULL)
 *   AUTO     the combination that measured fastest on B200 (today: ENCODER)
 *   NOFOLD_IDX / NOFOLD_ACT   keep a constant-input column block dense (csrc/fold.cu): the id-embedding columns of encoder
 *            layer 0 (instead of a per-agent bias) / the action-embedding half of decoder layer 0 (K = A*C instead of
 *            A*n_act one-hot columns).  Same mathematics, different bf16 rounding points; used by the tests that pin the
 *            folded path against the dense one.  ENCODER implies NOFOLD_IDX (that kernel stages the embedding columns). */
enum { MFVAE_FUSE_AUTO = 0, MFVAE_FUSE_NONE = 1, MFVAE_FUSE_ENCODER = 2, MFVAE_FUSE_LOSS = 4, MFVAE_FUSE_NOFOLD_IDX = 8,
       MFVAE_FUSE_NOFOLD_ACT = 16 };
enum { MFVAE_LOSS_DEFAULT = 0, MFVAE_LOSS_HUBER = 1, MFVAE_LOSS_MSE = 2, MFVAE_LOSS_JOINT_MSE = 3 };

typedef struct MfvaeConfig {
  int32_t n_agents;                      /* A                                   model.py:116 */
  int32_t idx_features;                  /* I  (IDX_FEATURES, main.py:30)                    */
  int32_t latent;                        /* L  (OBS_FEATURES, main.py:31)                    */
  int32_t act_features;                  /* C  (ACT_FEATURES, main.py:32)                    */
  int32_t n_enc_hidden;                  /* Encoder.HIDDEN  model.py:46 -> {64,64,256}       */
  int32_t enc_hidden[MFVAE_MAX_HIDDEN];
  int32_t n_dec_hidden;                  /* Decoder.HIDDEN  model.py:87 -> {1024,256,64,256,1024} */
  int32_t dec_hidden[MFVAE_MAX_HIDDEN];
  const int32_t* obs_dim;                /* [A] host, observation width per agent            */
  const int32_t* n_act;                  /* [A] host: the reference's action_dim dict -- #discrete actions per agent, or the
                                            action-vector width per agent when continuous_act = 1                       */
  float kl_weight;                       /* model.py:5                                       */
  float r_weight;                        /* model.py:6                                       */
  int32_t huber;                         /* 1: Huber(delta=1) (model.py:25-31), 0: MSE       */
  int32_t precision;                     /* MFVAE_PREC_*                                     */
  int32_t engine;                        /* MFVAE_ENGINE_*                                   */
  int32_t optimize_encoders;             /* 0 = reference semantics: encoders / action tables get
                                            gradients but no Adam update (model.py:112,114)  */
  int32_t fusion;                        /* MFVAE_FUSE_*: cross-layer kernel fusion (tcgen05 engine only)  */
  int32_t continuous_act;                /* 0: nn.Embedding per agent (DESCRETE_ACT = True, model.py:121); 1: ActionEncoder MLP
                                            action_dim -> act_hidden -> C per agent (model.py:60-74,123,148)                */
  int32_t act_hidden;                    /* ActionEncoder.HIDDEN[0] (model.py:63: 64); 0 = 64                              */
} MfvaeConfig;

/* One tensor of the parameter arena.  Offsets are in ELEMENTS and identical for the fp32 master,
 * gradient, Adam m / v and bf16 shadow arenas.  A tensor is rows x cols with row stride ld
 * (ld >= cols; padding columns are kept at zero). */
enum { MFVAE_T_IDX_EMB = 0, MFVAE_T_ENC_W, MFVAE_T_ENC_B, MFVAE_T_ACT_TABLE,
       MFVAE_T_SDEC_W, MFVAE_T_SDEC_B, MFVAE_T_RDEC_W, MFVAE_T_RDEC_B,
       MFVAE_T_RLIN_W, MFVAE_T_RLIN_B, MFVAE_T_ACTENC_W, MFVAE_T_ACTENC_B /* layer 0 / 1 of the ActionEncoder */ };
typedef struct MfvaeTensorInfo {
  int32_t kind;      /* MFVAE_T_*                         */
  int32_t agent;     /* agent index or -1                 */
  int32_t layer;     /* Linear index inside its MLP or -1 */
  int32_t rows, cols, ld;
  int64_t offset;
} MfvaeTensorInfo;

typedef struct MfvaeArenas {
  float* d_param;          /* fp32 master weights                       */
  float* d_grad;           /* fp32 gradients                            */
  float* d_m;              /* Adam exp_avg                              */
  float* d_v;              /* Adam exp_avg_sq                           */
  void*  d_shadow_bf16;    /* bf16 copy the tensor-core GEMMs read (may be NULL for fp32) */
} MfvaeArenas;

/* Packed device batch (what create_dataset builds on the host in the reference, trainer.py:7-45):
 * obs/next are the agents' observation vectors concatenated in codebook order. */
typedef struct MfvaeBatch {
  const float* d_obs;      /* [B, S]   fp32                                            */
  const float* d_act;      /* [B, A]   fp32-coded action index (replay_buffer.py:76); continuous_act: [B, sum(action_dim)]
                              action vectors of all agents concatenated in codebook order                   */
  const float* d_next;     /* [B, S]   target next state   (may be NULL for forward only) */
  const float* d_rew;      /* [B, A]   target rewards      (may be NULL for forward only) */
  const float* d_idx;      /* [B, A]   fp32-coded agent index column, or NULL = codebook order */
  const float* d_eps;      /* [B, A*L] explicit eps, or NULL = Philox(seed, step, sample)      */
  int32_t batch;           /* B on this rank                                            */
  int64_t sample0;         /* global index of row 0 (Philox counter; rank * B for DP)   */
  int64_t batch_global;    /* B summed over ranks: the loss means divide by this        */
  uint64_t seed;           /* Philox key                                                */
  uint64_t step;           /* Philox counter word 3                                     */
  int32_t obs_bf16;        /* 1: d_obs points at bf16 [B, S] (what a host ring that stores observations in bf16 ships: half
                              the PCIe bytes).  The bf16 engine rounds observations to bf16 on arrival anyway, so results
                              are bit-identical to feeding the fp32 values those bf16 numbers came from.               */
  int32_t next_bf16;       /* 1: d_next points at bf16 [B, S]: the reconstruction TARGET is rounded -- changes the loss
                              at the 1e-3 level, off by default                                                        */
} MfvaeBatch;

typedef struct MfvaeOutputs {
  const float* d_recon_s;  int32_t recon_s_ld;   /* [B, S]  fp32 (model.py:169)          */
  const float* d_recon_r;  int32_t recon_r_ld;   /* [B, A]  fp32 (model.py:170)          */
  const float* d_latent;                          /* [A][B][2L] fp32: cols [0,L)=mu, [L,2L)=logvar (model.py:149-150) */
  const float* d_losses;                          /* [4] loss, s_loss, r_loss, kl_loss (model.py:40); partial
                                                     (this rank's share of the global means) until all-reduced */
} MfvaeOutputs;

typedef struct MfvaeHandle_* MfvaeHandle;

const char* mfvae_last_error(void);
int mfvae_version(void);

/* lifecycle.  device = -1 makes a layout-only handle (arena / tensor table queries work, every compute
 * call fails): used by host-side tests on machines without a GPU. */
int mfvae_create(const MfvaeConfig* cfg, int device, MfvaeHandle* out);
int mfvae_destroy(MfvaeHandle h);

/* parameter arena description */
int64_t mfvae_arena_elems(MfvaeHandle h);                 /* total elements (multiple of 8)          */
int64_t mfvae_optimized_elems(MfvaeHandle h);             /* prefix [0, n) that Adam updates          */
int32_t mfvae_tensor_count(MfvaeHandle h);
int mfvae_tensor_table(MfvaeHandle h, MfvaeTensorInfo* out, int32_t capacity);
int mfvae_bind_arenas(MfvaeHandle h, const MfvaeArenas* arenas);
int mfvae_refresh_shadow(MfvaeHandle h, void* stream);    /* fp32 master -> bf16 shadow (whole arena) */
/* same over arena elements [begin, end) (4-element aligned): the drop-in forward re-casts reward_linear on every call,
 * because the reference's POP-ART idiom edits it through `.data` (torch_ver/trainer.py:73-74), invisibly to autograd */
int mfvae_refresh_shadow_range(MfvaeHandle h, int64_t begin, int64_t end, void* stream);

/* activation workspace: caller allocates mfvae_workspace_bytes(h, B) bytes (256-B aligned) */
int64_t mfvae_workspace_bytes(MfvaeHandle h, int32_t batch);
int mfvae_bind_workspace(MfvaeHandle h, void* d_ws, int64_t bytes, int32_t batch);

/* the train step, piecewise (reference call sites in the file header) */
int mfvae_forward(MfvaeHandle h, const MfvaeBatch* b, MfvaeOutputs* out, void* stream);
/* loss_kind: MFVAE_LOSS_DEFAULT (cfg.huber), _HUBER / _MSE (loss_s_r_vae_fn, model.py:19-40) or
 * _JOINT_MSE (loss_vae_fn, model.py:8-16) */
int mfvae_loss(MfvaeHandle h, const MfvaeBatch* b, int32_t loss_kind, void* stream);
/* the reference reads kl_weight / r_weight module globals at call time (model.py:5-6,34,39) */
int mfvae_set_loss_weights(MfvaeHandle h, float kl_weight, float r_weight);
/* jax_ver weighting (jax_ver/trainer.py:42-43,64): loss = s_weight * s + r_weight * r + kl_weight * kl with
 * s_weight = 1 - r_weight, kl_weight = 0.1, r_weight = 0.5.  mfvae_set_loss_weights resets s_weight to 1. */
int mfvae_set_loss_weights3(MfvaeHandle h, float kl_weight, float r_weight, float s_weight);
int mfvae_backward(MfvaeHandle h, const MfvaeBatch* b, void* stream);
/* backward seeded by caller-provided upstream gradients (autograd bridge for losses other than the fused
 * ELBO): d loss / d recon_s [B, ld_s], d loss / d recon_r [B, ld_r], d loss / d latent [A][B][2L]; fp32, any
 * of them may be NULL (= zero).  The analytic KL term is NOT added (it is part of the caller's loss). */
int mfvae_backward_ext(MfvaeHandle h, const MfvaeBatch* b, const float* d_g_recon_s, int64_t ld_s,
                       const float* d_g_recon_r, int64_t ld_r, const float* d_g_latent, void* stream);
int mfvae_adam_step(MfvaeHandle h, float lr, float beta1, float beta2, float eps, int64_t t, void* stream);
/* same update; the decoder block runs on an internal stream as soon as its gradient buckets are final (overlapping the
 * encoder half of backward).  Only valid when no collective has to run between backward and the update (1 GPU). */
int mfvae_adam_step_overlapped(MfvaeHandle h, float lr, float beta1, float beta2, float eps, int64_t t, void* stream);
/* data parallel: Adam over one arena range (a gradient bucket, 8-element aligned, inside the optimised prefix) on `stream`
 * -- issued on the communication stream right behind that bucket's all-reduce -- and the guard that makes `stream` wait
 * until the backward pass in flight no longer reads the decoder weights (needed before updating buckets 0..2). */
int mfvae_adam_range(MfvaeHandle h, int64_t begin, int64_t end, float lr, float beta1, float beta2, float eps, int64_t t,
                     void* stream);
int mfvae_wait_decoder_reads(MfvaeHandle h, void* stream);
/* finer guard: `stream` waits until backward has finished reading the weights of gradient bucket i (so that bucket's Adam
 * can start while later layers are still in backward) */
int mfvae_bucket_read_wait(MfvaeHandle h, int32_t i, void* stream);
/* persistent GEMM grids use (148 - n_sms) SMs, leaving room for the collective's kernels that run beside backward: a
 * persistent kernel whose CTAs cannot all be resident at once runs a second wave (process-wide setting) */
int mfvae_set_sm_reserve(MfvaeHandle h, int32_t n_sms);
/* forward + loss + backward in one call (no optimizer; the host all-reduces gradients in between).  With the bf16 /
 * tcgen05 engine the state reconstruction is consumed by the loss without an fp32 copy in HBM: out->d_recon_s is NULL
 * (d_recon_r, d_latent and d_losses are valid); call mfvae_forward when recon_s itself is wanted. */
int mfvae_fwd_bwd(MfvaeHandle h, const MfvaeBatch* b, MfvaeOutputs* out, void* stream);

/* ONE call = one train step: mfvae_fwd_bwd, then -- when mfvae_comm_bind has been called -- the data-parallel exchange of
 * every gradient bucket on the library's own communication stream with the bucket's Adam behind it, else the overlapped
 * single-GPU Adam.  t = 1-based optimizer step; pipeline = 1: do not order `stream` behind the decoder block's optimizer sweep
 * (mfvae_opt_join).  Equivalent to the call-by-call sequence of INTEGRATION.md section 4. */
int mfvae_train_step(MfvaeHandle h, const MfvaeBatch* b, float lr, float beta1, float beta2, float eps, int64_t t, int32_t pipeline,
                     MfvaeOutputs* out, void* stream);

/* instrumentation: number of kernels this library has launched so far (process-wide), and optional CUDA-event
 * timing of every GEMM launch of the step on its launching stream.  With profiling on, each step records
 * start/stop events around its GEMM launches; mfvae_profile_read synchronises and returns, per GEMM of the last
 * step, {M, N, K, groups, kind (0 fwd, 1 dgrad, 2 wgrad, 3 = the fused encoder chain: N = 1, K = MACs per (agent, sample)),
 * milliseconds}. */
uint64_t mfvae_launch_count(void);
typedef struct MfvaeGemmTiming { int32_t M, N, K, groups, kind; float ms; } MfvaeGemmTiming;
int mfvae_profile_enable(MfvaeHandle h, int32_t on);
int32_t mfvae_profile_read(MfvaeHandle h, MfvaeGemmTiming* out, int32_t capacity);

/* gradient buckets in backward-completion order, for the overlapped all-reduce: bucket i covers
 * arena elements [begin, end) and is final once event i (cudaEvent_t, returned as void*) fires. */
int32_t mfvae_bucket_count(MfvaeHandle h);
int mfvae_bucket(MfvaeHandle h, int32_t i, int64_t* begin, int64_t* end, void** event);
/* make `stream` wait until the four loss scalars of the step in flight are final (they are, before backward starts: their
 * all-reduce can run beside backward instead of after it) */
int mfvae_loss_wait(MfvaeHandle h, void* stream);
/* make `stream` (e.g. the NCCL stream) wait for bucket i's event */
int mfvae_bucket_wait(MfvaeHandle h, int32_t i, void* stream);

/* Data-parallel exchange step as the library's own kernels over NVLink / NVSwitch peer memory (csrc/comm.cu; SURVEY.md 8b asked
 * for mfvae_allreduce_grads next to the NCCL route).  The host maps ONE symmetric buffer of mfvae_comm_window_bytes() bytes per
 * rank into every rank (CUDA VMM / fabric handles; torch.distributed._symmetric_memory does it for the Python host) and hands in
 *   d_peer_windows   device array [world] of the windows' base addresses as mapped in THIS process (own window included)
 *   multicast_window multicast (NVLS) mapping of the same windows, or NULL -> plain peer loads / stores
 *   d_signal_pads    device array [world] of zero-initialised 32-bit flag pads of signal_pad_bytes each (slots [0, 64) are not used)
 * payload_bf16 = 1 ships gradients as bf16 (the switch / the reducing rank accumulates in fp32); 0 ships fp32.
 * Every rank must issue the same sequence of mfvae_allreduce_* calls.  max_blocks (0 = 64 at 2 ranks, 32 from 3 ranks: measured) caps the CTAs of a reduce; one 32-bit
 * flag per (CTA, peer) is used behind the first 64 slots of a pad, so signal_pad_bytes bounds it too. */
int mfvae_comm_bind(MfvaeHandle h, int32_t rank, int32_t world, void* const* d_peer_windows, void* multicast_window,
                    void* const* d_signal_pads, int64_t signal_pad_bytes, void* local_window, int64_t window_bytes, int32_t payload_bf16,
                    int32_t max_blocks);
int64_t mfvae_comm_window_bytes(MfvaeHandle h, int32_t payload_bf16);
/* sum over ranks of gradient-arena elements [begin, end) (a bucket of mfvae_bucket), on `stream`; do_adam = 1 runs the fused
 * Adam of that range right behind it (reduced gradient read from the window, fp32 copy left in the gradient arena) */
int mfvae_allreduce_grads(MfvaeHandle h, int64_t begin, int64_t end, int32_t do_adam, float lr, float beta1, float beta2, float eps,
                          int64_t t, void* stream);
/* do_adam = 1 sweeps run on an internal optimizer stream behind their reduce; this makes `stream` wait for all of them.
 * A caller may skip it ("pipelined" step): the next mfvae_forward / mfvae_fwd_bwd on this handle waits for the sweeps itself,
 * right before its decoder half -- the encoder half of the next step then overlaps the optimizer tail of this one.  Anything
 * else that reads parameters (the host copying them out) must call mfvae_opt_join first. */
int mfvae_opt_join(MfvaeHandle h, void* stream);
/* sum over ranks of the 4 loss scalars of the step in flight, in place */
int mfvae_allreduce_losses(MfvaeHandle h, void* stream);

/* standalone bandwidth-bound kernels (BASELINE config 5 microbenchmarks; also used by the step).
 * dtype: 0 = fp32 activations, 1 = bf16 activations (mu / logvar / recon / target stay fp32). */
int mfvae_reparam_kl(const float* d_mu, const float* d_logvar, const float* d_eps_or_null,
                     void* d_z, int32_t z_dtype, int64_t batch, int32_t width /* A*L */,
                     uint64_t seed, uint64_t step, int64_t sample0, int64_t batch_global,
                     float* d_kl_out /* [1] */, float* d_scratch /* >= 4096 floats */, void* stream);
int mfvae_recon_loss(const float* d_recon, int32_t recon_ld, const float* d_target, int32_t target_ld,
                     void* d_grad, int32_t grad_ld, int32_t grad_dtype, int64_t batch, int32_t width,
                     int32_t huber, float weight, int64_t count_global,
                     float* d_loss_out /* [1] */, float* d_scratch /* >= 4096 floats */, void* stream);
int mfvae_adam_flat(float* d_p, const float* d_g, float* d_m, float* d_v, void* d_shadow_bf16_or_null,
                    int64_t n, float lr, float beta1, float beta2, float eps, int64_t t, void* stream);
int mfvae_philox_normal(float* d_out, int64_t batch, int32_t width, uint64_t seed, uint64_t step,
                        int64_t sample0, void* stream);

/* generic GEMM entry (tests + microbench): C[M,N] (+)= A[M,K] * B[N,K]^T with element strides.
 * dtype 0 fp32 / 1 bf16 operands; C fp32 when c_dtype = 0 else bf16.  engine: MFVAE_ENGINE_*.
 * epilogue: 0 none, 1 +bias, 2 +bias,relu, 3 multiply by (aux > 0), 4 accumulate into fp32 C. */
int mfvae_gemm(int32_t engine, int32_t dtype, int32_t groups, int32_t M, int32_t N, int32_t K,
               const void* d_A, int64_t a_gs, int64_t a_rs, int64_t a_cs,
               const void* d_B, int64_t b_gs, int64_t b_rs, int64_t b_cs,
               void* d_C, int64_t c_gs, int64_t c_ld, int32_t c_dtype,
               const float* d_bias, int64_t bias_gs, int32_t epilogue,
               const void* d_aux, int64_t aux_gs, int64_t aux_ld, int32_t split_k, void* stream);

/* HBM-resident replay ring (replaces cpprb.ReplayBuffer / flashbax item buffer on this path).
 * Row = one joint transition with every key of the reference's cpprb env_dict (torch_ver/src/replay_buffer.py:62-81):
 *   [obs(S) | act(W) | next(S) | rew(A) | terminals(A) | truncations(A) | mask(1)] fp32,
 * W = act_cols = A for float-coded discrete actions, sum of the agents' action widths for continuous actions. */
typedef struct MfvaeRing_* MfvaeRing;
int mfvae_ring_create(int32_t state_dim, int32_t n_agents, int32_t act_cols, int64_t capacity, float* d_storage, MfvaeRing* out);
int mfvae_ring_destroy(MfvaeRing r);
int64_t mfvae_ring_row_floats(int32_t state_dim, int32_t n_agents, int32_t act_cols);
int64_t mfvae_ring_size(MfvaeRing r);
/* append n rows from a host or device staging buffer (cudaMemcpyAsync, wraps around) */
int mfvae_ring_add(MfvaeRing r, const float* rows, int64_t n, int32_t rows_on_device, void* stream);
/* uniform-with-replacement sample (Philox(seed, step)) and gather into the packed batch matrices; d_flags_or_null
 * [batch, 2 A + 1] receives terminals | truncations | mask of the sampled rows */
int mfvae_ring_sample(MfvaeRing r, int64_t batch, uint64_t seed, uint64_t step,
                      float* d_obs, float* d_act, float* d_next, float* d_rew, float* d_flags_or_null,
                      int32_t* d_indices_or_null, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFVAE_H_ */
