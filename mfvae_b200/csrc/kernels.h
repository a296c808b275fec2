// kernels.h — internal launch API shared by plan.cu / api.cu (not part of the C ABI).
#pragma once
#include "common.cuh"

struct CUtensorMap_st;            // CUtensorMap of <cuda.h>

namespace mfvae {

enum DType { kF32 = 0, kBF16 = 1 };
static inline size_t dtype_size(int dt) { return dt == kBF16 ? 2 : 4; }

// ---- bandwidth-bound kernels (elementwise.cu) -------------------------------------------------
struct StageArgs {
  const void* obs; int obs_ld; int obs_dtype;   // [B, S] fp32 (kF32) or bf16 (kBF16)
  const float* act; int act_ld;          // [B, A] fp32-coded
  const float* idx;                      // [B, A] fp32-coded or nullptr
  const float* idx_emb;                  // [A, I] fp32 master
  const float* act_table; int64_t act_table_gs; int n_act_max;  // [A][n_act_max][C] fp32 master
  const int32_t* obs_off; const int32_t* obs_dim; const int32_t* n_act;   // device [A]
  void* x0; int x0_ld; int64_t x0_gs;    // [A][B][K0p]
  void* zin; int zin_ld;                 // [B, Din]: act-emb written at column A*L + a*C
  int A, I, L, C, B;
  int dtype;
};
// do_x0: encoder inputs X0; do_act: action-embedding half of the decoder input
int launch_stage(const StageArgs& a, cudaStream_t s, bool do_x0 = true, bool do_act = true);

struct ReparamArgs {
  const float* mu; const float* lv;      // element (b,a,j) at p + a*lat_as + b*lat_bs + j
  int64_t lat_as, lat_bs;
  const float* eps; int64_t eps_ld;      // [B, A*L] or nullptr -> Philox
  void* z; int64_t z_ld; int z_dtype;    // z[b][a*L + j]
  int64_t B; int A, L;
  uint64_t seed, step; int64_t sample0;
  float kl_scale;                        // 1 / batch_global
  float* kl_out;                         // [1]
  float* scratch;                        // >= 4096 floats (partials + ticket)
};
int launch_reparam_kl_fwd(const ReparamArgs& a, cudaStream_t s);

struct ReparamBwdArgs {
  const void* gz; int64_t gz_ld; int g_dtype;   // dL/dz [B, >= A*L]
  const float* mu; const float* lv; int64_t lat_as, lat_bs;
  const float* eps; int64_t eps_ld;
  void* dlat; int64_t dlat_as, dlat_bs; int d_dtype;   // [A][B][2L]
  int64_t B; int A, L;
  uint64_t seed, step; int64_t sample0;
  float kl_scale;                        // kl_weight / batch_global
  const float* glat;                     // optional upstream d/d(mu, logvar), same layout as mu/lv base (fp32)
};
int launch_reparam_kl_bwd(const ReparamBwdArgs& a, cudaStream_t s);

struct ReconLossArgs {
  const float* recon; int64_t recon_ld;
  const __nv_bfloat16* recon16;                  // when set: the reconstruction in bf16 (same ld as grad; may alias grad: in-place)
  const float* target; int64_t target_ld;
  const __nv_bfloat16* target16;                 // when set: the target in bf16 (same ld), `target` is ignored
  void* grad; int64_t grad_ld; int grad_dtype;   // may be nullptr (forward value only)
  int64_t B; int width;
  int huber;
  float grad_scale;      // weight / count_global
  float loss_scale;      // 1 / count_global
  float* loss_out;       // [1]
  float* scratch;        // >= 4096 floats
  float* colsum_out = nullptr;   // optional [width] fp32: += column sums of the stored gradient (the producing layer's bias gradient)
};
int launch_recon_loss(const ReconLossArgs& a, cudaStream_t s);

// out[g][n] += sum_b X[g][b][n]   (fp32 atomics; `out` must be zeroed by the caller)
int launch_colsum(const void* x, int dtype, int G, int64_t B, int N, int64_t ld, int64_t gs,
                  float* out, int64_t out_gs, cudaStream_t s);

// d_emb[idx[b][a]][i] += gx0[a][b][i]  (general index path)
int launch_idx_emb_scatter(const void* gx0, int dtype, int64_t gs, int64_t ld, const float* idx, int idx_ld,
                           int A, int I, int64_t B, float* d_emb, cudaStream_t s);
// d_table[a][act[b][a]][c] += gzin[b][col0 + a*C + c]
int launch_act_table_grad(const void* gzin, int dtype, int64_t ld, int col0, const float* act, int act_ld,
                          const int32_t* n_act, int A, int C, int64_t B, float* d_table, int64_t table_gs,
                          cudaStream_t s);

// continuous actions: act [B, sum D_a] fp32 -> ACT0[a][b][0..Kap) (zero padded), the ActionEncoder's first operand
int launch_stage_actions(const float* act, int64_t act_ld, const int32_t* act_off, const int32_t* act_dim, void* act0, int dtype,
                         int A, int64_t B, int Kap, cudaStream_t s);

int launch_adam(float* p, const float* g, float* m, float* v, __nv_bfloat16* shadow, int64_t n,
                float lr, float b1, float b2, float eps, int64_t t, cudaStream_t s, int blocks_per_sm = 8);
// dst[b][c] = src[b][c] for c < width (fp32 -> activation dtype); src == nullptr writes zeros
int launch_cast2d(const float* src, int64_t src_ld, void* dst, int64_t dst_ld, int dtype, int64_t B, int width, cudaStream_t s);
int launch_cast_bf16(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t s);
int launch_philox_normal(float* out, int64_t B, int width, uint64_t seed, uint64_t step, int64_t sample0,
                         cudaStream_t s);
// losses[0] = s_w * s + r_w * r + kl_w * kl  given losses[1..3]
// also sums `n_partials` per-warp partials of the fused output-layer loss into losses[1] (scaled) when partials != nullptr
int launch_loss_total(float* losses, float s_weight, float r_weight, float kl_weight, cudaStream_t s, const float* partials = nullptr,
                      int n_partials = 0, float partial_scale = 0.f);

// ---- data-parallel exchange over peer memory (comm.cu) ------------------------------------------
struct CommCtx {
  int rank = 0, world = 1;
  void* const* d_peers = nullptr;      // device array [world]: base pointer of every rank's window of the symmetric buffer
  void* mc = nullptr;                  // multicast base of the windows (NVSwitch), or nullptr -> plain peer loads / stores
  uint32_t* const* d_pads = nullptr;   // device array [world]: every rank's signal pad
  void* local = nullptr;               // this rank's window
  int dtype = kF32;                    // payload type of the gradient region (element i of the arena at element i of the window)
  int64_t elems = 0;                   // gradient region: elements
  int64_t scalar_off = 0;              // byte offset of a 256-byte fp32 scalar region behind it
  int64_t small_off = 0, small_bytes = 0;   // fp32 scratch for tiny ranges reduced by one single-CTA kernel
  int max_blocks = 16;
  int split_sync = 0;                  // 1: pack / rendezvous / reduce / rendezvous as four launches (no wide kernel ever waits for a peer)
};
int comm_pack(const float* g, void* window, int dtype, int64_t begin, int64_t end, cudaStream_t s);
int comm_unpack(const void* window, int dtype, float* g, int64_t begin, int64_t end, cudaStream_t s);
int comm_allreduce(const CommCtx& c, int64_t begin, int64_t end, cudaStream_t s, const float* pack_from = nullptr);   // pack_from: fp32 arena whose [begin, end) is packed first
int comm_allreduce_scalars(const CommCtx& c, const float* src, int n, float* out, cudaStream_t s);
int comm_small_allreduce_adam(const CommCtx& c, int64_t n, float* p, float* g, float* m, float* v, __nv_bfloat16* shadow, int do_adam,
                              float lr, float b1, float b2, float eps, int64_t t, cudaStream_t s);
int launch_adam_payload(float* p, const void* gw, int dtype, float* g, float* m, float* v, __nv_bfloat16* shadow, int64_t n,
                        float lr, float b1, float b2, float eps, int64_t t, cudaStream_t s);

// ---- folded constant-input column blocks (fold.cu) ---------------------------------------------
int launch_onehot(const float* act, int act_ld, const int32_t* n_act, void* zin, int dtype, int64_t zin_ld, int col0, int A, int nmax,
                  int width, int64_t B, cudaStream_t s);
int launch_act_fold_fwd(const float* W0, int64_t w_ld, int wcol0, const float* table, int64_t table_gs, void* Tt, int dtype, int64_t t_ld,
                        int rows, int A, int C, int nmax, cudaStream_t s);
int launch_act_fold_bwd(const float* dT, int64_t t_ld, const float* W0, float* gW0, int64_t w_ld, int wcol0, const float* table, float* gtable,
                        int64_t table_gs, int rows, int A, int C, int nmax, cudaStream_t s);
int launch_enc_bias_fold(const float* W0, int64_t w_ld, const float* b0, const float* emb, float* eb, int A, int N, int I, cudaStream_t s);
int launch_emb_grad_fold(const float* W0, float* gW0, int64_t w_ld, const float* db0, const float* emb, float* g_emb, int A, int N, int I,
                         cudaStream_t s);

// ---- GEMM (gemm_simt.cu / gemm_tc.cu) --------------------------------------------------------
enum Epilogue { kEpiNone = 0, kEpiBias = 1, kEpiBiasRelu = 2, kEpiReluMask = 3, kEpiAccum = 4, kEpiLossGrad = 5 };

struct GemmOp {
  int G = 1, M = 0, N = 0, K = 0;
  int dtype = kF32;                           // operand type
  const void* A = nullptr; int64_t a_gs = 0, a_rs = 0, a_cs = 0;   // A(m,k)
  const void* B = nullptr; int64_t b_gs = 0, b_rs = 0, b_cs = 0;   // B(n,k)
  // optional second K segment of the B operand (K-major only): for k >= k_split, B(n,k) = B2(n, k - k_split).  k_split must
  // be a multiple of 64.  Lets a Linear layer read part of its weight columns from a derived table (decoder layer 0:
  // [W0 z-columns | W0_act . action tables]) without materialising the concatenation.
  const void* B2 = nullptr; int64_t b2_gs = 0, b2_rs = 0; int k_split = 0, k1 = 0, k2 = 0;   // k1 / k2 = valid columns of B / B2 (K = k_split + k2)
  void* C = nullptr; int64_t c_gs = 0, c_ld = 0; int c_dtype = kF32;
  const float* bias = nullptr; int64_t bias_gs = 0;
  int epi = kEpiNone;
  const void* aux = nullptr; int64_t aux_gs = 0, aux_ld = 0;      // relu-mask source, operand dtype
  int split_k = 1;                            // > 1 requires kEpiAccum into a zeroed fp32 C
};
int gemm_simt(const GemmOp& op, cudaStream_t s);

// tcgen05 path: tensor maps are built once per op (plan time) and reused every step.
extern int g_tc_sm_reserve;                             // SMs the persistent GEMM grids leave to concurrent collectives
struct TcPlan;
int gemm_tc_plan(const GemmOp& op, TcPlan** out);       // validates alignment, encodes CUtensorMaps
int gemm_tc_run(const TcPlan* p, cudaStream_t s);
void gemm_tc_free(TcPlan* p);
bool gemm_tc_overwrites(const TcPlan* p);
// kEpiLossGrad plans: per-call target / scale / loss-partial buffer (one float per epilogue warp of the grid)
int gemm_tc_set_loss(TcPlan* p, const float* tgt, int64_t tgt_ld, float grad_scale, int huber, float* partials);
int gemm_tc_loss_partials(const TcPlan* p);               // true: C is fully written by plain stores (no pre-zeroing needed)

// ---- fused per-agent encoder (enc_fused.cu) ---------------------------------------------------
constexpr int kEncMaxL = 4;
struct EncFusedDesc {                          // everything that is fixed once arenas + workspace are bound
  int A = 0, B = 0, nl = 0, L = 0, K0 = 0;     // K0 = width of the staged input [obs | 0] (the id-embedding is folded into bias[0])
  int N[kEncMaxL] = {0, 0, 0, 0};              // output width of layer l (last = 2L)
  const void* W[kEncMaxL] = {};                // bf16, first valid column of layer l's weights: element (a, n, k) at W + a*w_gs + n*w_ld + k
  int64_t w_ld[kEncMaxL] = {}, w_gs[kEncMaxL] = {};
  const float* bias[kEncMaxL] = {};            // fp32 [A][N_l]
  void* X[kEncMaxL] = {};                      // bf16 input of layer l, [A][B][x_ld]: X0F (read), XE_0, ... (written)
  int64_t x_ld[kEncMaxL] = {}, x_gs[kEncMaxL] = {};
  float* lat = nullptr; int64_t lat_gs = 0, lat_ld = 0;
  void* zin = nullptr; int64_t zin_ld = 0;
};
struct EncFwdBatch {
  const float* eps; int64_t eps_ld; uint64_t seed, step; int64_t sample0;
  float kl_scale; float* kl_out; float* scratch;
};
struct EncFusedPlan;
bool enc_fused_applicable(const EncFusedDesc& d);
int enc_fused_plan(const EncFusedDesc& d, EncFusedPlan** out);
void enc_fused_free(EncFusedPlan* p);
int enc_fused_forward(EncFusedPlan* p, const EncFwdBatch& b, cudaStream_t s);

// shared tensor-map encoder (gemm_tc.cu)
int encode_tmap_bf16_3d(::CUtensorMap_st* map, const void* base, int64_t inner, int64_t outer, int64_t G, int64_t ld, int64_t gs,
                        int box_inner, int box_outer);

}  // namespace mfvae
