// plan.cu — handle, arena / workspace layout and the op sequence of one MAVAE train step.
//
// Reference data flow being reproduced (all citations /root/reference/torch_ver):
//   forward   model.py:134-173      loss  model.py:19-40      backward  main.py:92-93      Adam  main.py:97
//
// HBM layout:
//   parameter arena (fp32 master | grad | m | v | bf16 shadow share one element layout)
//     [ idx_emb | dec hidden l: Ws_l, Wr_l, bs_l, br_l | Ws_out, bs_out, Wr_out, br_out, rlW, rlb ] <- Adam prefix
//     [ enc l: W[A][N_l][K_l], b[A][N_l] | action tables [A][n_act_max][C] ]
//   activation workspace for a bound batch B
//     X0[A][B][K0p]  XE_l[A][B][H_l]  LAT[A][B][2L](fp32)  ZIN[B][A(L+C)]  HD_l[B][2 H_l] (state | reward halves)
//     RS[B][Sp](fp32) RR0[B][Ap] RR[B][Ap](fp32)  and the matching gradient buffers.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels.h"

namespace mfvae {

unsigned long long g_launch_count = 0;
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(const char* file, int line, const std::string& msg) {
  const char* base = strrchr(file, '/');
  g_last_error = std::string(base ? base + 1 : file) + ":" + std::to_string(line) + ": " + msg;
  return 1;
}

struct Span { int64_t off = 0; int rows = 0, cols = 0, ld = 0; };

}  // namespace mfvae

using namespace mfvae;

struct MfvaeHandle_ {
  MfvaeConfig cfg{};
  int device = 0;
  std::vector<int32_t> obs_dim, n_act, obs_off;
  int A = 0, I = 0, L = 0, C = 0, S = 0, Sp = 0, Ap = 0, Ip = 0, Din = 0, K0p = 0, nact_max = 0;
  float s_weight = 1.0f;                     // weight of the state reconstruction term (1 in torch_ver; 1 - r_weight in jax_ver)
  bool cont_act = false; int Hact = 64, Kap = 8, act_total = 0;   // continuous actions: ActionEncoder D_a -> Hact -> C
  std::vector<int32_t> act_off;
  // folded constant-input column blocks (fold.cu): discrete actions -> one-hot columns against T = W0_act . tables;
  // codebook agent index -> per-agent bias of encoder layer 0
  bool fold_act = false; int Kz = 0, Kzp = 0, Koh = 0, Kohp = 0, K0f = 0;
  int ne = 0, nd = 0;                       // number of Linear layers in encoder / decoder
  std::vector<int> encN, encK;              // per encoder layer (K padded)
  std::vector<int> decH;                    // decoder hidden widths
  int dtype = kF32;
  bool use_tc = false;

  // arena layout
  Span idx_emb, rlW, rlb, sOutW, sOutB, rOutW, rOutB, actT;
  Span actW1, actB1, actW2, actB2;          // ActionEncoder (continuous actions): [A][Hact][Kap], [A][Hact], [A][C][Hact], [A][C]
  std::vector<Span> decW, decB;             // per hidden layer: rows = 2*H (state rows then reward rows)
  std::vector<Span> encW, encB;             // rows = A*N_l
  int64_t arena_elems = 0, optimized_elems = 0, reg3_begin = 0, reg2_begin = 0, enc_begin = 0;
  std::vector<MfvaeTensorInfo> table;
  MfvaeArenas ar{};

  // device constants
  int32_t* d_meta = nullptr;                // obs_off | obs_dim | n_act | act_off

  // workspace
  char* ws = nullptr; int64_t ws_bytes = 0; int B = 0;
  struct Buf { int64_t off = 0; int64_t ld = 0; int64_t gs = 0; };
  Buf X0, LAT, ZIN, RS, RR0, RR, DRS, DRR, DRR0, GZIN, DLAT, GX0, ACT0, HA, DHA;
  Buf X0F, TACT, DTACT, EB0;                  // folded paths: X0F[A][B][K0f] = [obs | 0]; TACT / DTACT [2 H0][Kohp]; EB0[A][N0] fp32
  std::vector<Buf> XE, DXE, HD, DHD;
  int64_t off_losses = 0, off_scratch = 0;

  // GEMM ops
  std::vector<GemmOp> gemms;
  std::vector<TcPlan*> tc;
  EncFusedPlan* enc_fused = nullptr;         // fused per-agent encoder chain (enc_fused.cu), when its shape constraints hold
  std::vector<int> g_enc_fwd, g_enc_wg, g_enc_dg, g_dec_fwd, g_dec_wg, g_dec_dg;
  int g_sout_fwd = -1, g_rout_fwd = -1, g_rl_fwd = -1;
  int g_sout_loss = -1;
  int g_sout_fwd16 = -1;                     // train step: recon_s written once, in bf16, into the D(recon_s) buffer (loss runs in place)                      // state output layer with the reconstruction loss + its gradient as the epilogue
  int g_act_fwd1 = -1, g_act_fwd2 = -1, g_act_wg2 = -1, g_act_dg2 = -1, g_act_wg1 = -1;
  int g_sout_wg = -1, g_sout_dg = -1, g_rout_wg = -1, g_rout_dg = -1, g_rl_wg = -1, g_rl_dg = -1;
  int g_dec_wgT = -1;                        // d T = d H0^T . onehot(act)            (fold_act)
  int g_enc_fwd0_f = -1, g_enc_wg0_f = -1;   // encoder layer 0 on [obs | 0] with the folded bias (codebook agent index)
  cudaEvent_t eb_ev = nullptr;               // folded encoder bias is ready (aux stream)

  // side stream for the wgrad / bias-gradient chain of backward, with its fork / join events
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> fork_ev;
  cudaEvent_t join_ev = nullptr;
  // third stream: the reward head (two tiny GEMM chains) and the action-embedding kernels run beside the big layers
  cudaStream_t aux = nullptr;
  bool grads_zeroed = false;                 // mfvae_fwd_bwd: the gradient arena was zeroed on csum beside the forward pass
  bool opt_pending = false;                  // optimizer sweeps on opt_stream that no stream of the next forward has waited for yet
  bool sout_bias_done = false;               // the state head's bias gradient was accumulated by the loss kernel (mfvae_fwd_bwd)
  cudaEvent_t zero_ev = nullptr, zero_fork_ev = nullptr;
  cudaEvent_t loss_ev = nullptr;             // the four loss scalars are final (recorded at the end of the loss phase)
  cudaStream_t csum = nullptr;               // fourth stream: the bias column sums (HBM-bound) run beside the wgrad GEMMs
  std::vector<cudaEvent_t> csum_ev;
  cudaEvent_t aux_fork_ev = nullptr, aux_join_ev = nullptr, aux_fork2_ev = nullptr, aux_join2_ev = nullptr;
  cudaStream_t opt_stream = nullptr;         // overlapped Adam
  cudaEvent_t opt_ev = nullptr, dec_read_ev = nullptr;
  cudaEvent_t read_ev[3] = {nullptr, nullptr, nullptr};   // backward has finished READING the weights of gradient bucket 0 / 1 / 2

  // optional per-GEMM event timing (bench.py roofline)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_ev;          // 2 per GEMM op
  std::vector<char> prof_hit;
  cudaEvent_t enc_prof_ev[2] = {nullptr, nullptr}; bool enc_prof_hit = false;   // the fused encoder chain, timed as one item

  std::vector<cudaEvent_t> ar_ev = std::vector<cudaEvent_t>(8, nullptr); size_t ar_ev_i = 0;   // reduce -> optimizer-stream hand-off
  cudaStream_t comm_stream = nullptr; cudaEvent_t comm_done_ev = nullptr;   // mfvae_train_step's own communication stream
  CommCtx comm;                              // data-parallel exchange over peer memory (mfvae_comm_bind); world == 1: unbound
  // gradient buckets (arena ranges) and their completion events
  struct Bucket { int64_t begin, end; cudaEvent_t ev; };
  std::vector<Bucket> buckets;
};

namespace mfvae {

static int64_t take(int64_t& cursor, int64_t n) { int64_t o = cursor; cursor += round_up(n, 8); return o; }

static void add_info(MfvaeHandle_* h, int kind, int agent, int layer, int rows, int cols, int ld, int64_t off) {
  MfvaeTensorInfo t{kind, agent, layer, rows, cols, ld, off};
  h->table.push_back(t);
}

static int build_layout(MfvaeHandle_* h) {
  const MfvaeConfig& c = h->cfg;
  h->A = c.n_agents; h->I = c.idx_features; h->L = c.latent; h->C = c.act_features;
  MFVAE_CHECK(h->A >= 1 && h->A <= 4096, "n_agents out of range");
  MFVAE_CHECK(h->I % 8 == 0 && h->L % 8 == 0 && h->C % 8 == 0, "idx_features, latent and act_features must be multiples of 8");
  MFVAE_CHECK(c.n_enc_hidden >= 1 && c.n_enc_hidden <= MFVAE_MAX_HIDDEN, "n_enc_hidden out of range");
  MFVAE_CHECK(c.n_dec_hidden >= 1 && c.n_dec_hidden <= MFVAE_MAX_HIDDEN, "n_dec_hidden out of range");
  int maxo = 0; h->S = 0; h->nact_max = 0;
  h->obs_off.resize(h->A);
  for (int a = 0; a < h->A; ++a) {
    MFVAE_CHECK(h->obs_dim[a] >= 1 && h->n_act[a] >= 1, "obs_dim / n_act must be positive");
    h->obs_off[a] = h->S; h->S += h->obs_dim[a];
    maxo = std::max(maxo, h->obs_dim[a]); h->nact_max = std::max(h->nact_max, h->n_act[a]);
  }
  h->Sp = static_cast<int>(round_up(h->S, 8)); h->Ap = static_cast<int>(round_up(h->A, 8));
  h->Ip = h->I;
  h->K0p = static_cast<int>(round_up(h->I + maxo, 8));
  h->Din = h->A * (h->L + h->C);
  h->fold_act = (c.continuous_act == 0) && !(c.fusion & MFVAE_FUSE_NOFOLD_ACT);
  h->Kz = h->A * h->L; h->Kzp = static_cast<int>(round_up(h->Kz, 64));
  h->Koh = h->A * h->nact_max; h->Kohp = static_cast<int>(round_up(h->Koh, 8));
  h->K0f = h->K0p - h->I;
  h->ne = c.n_enc_hidden + 1; h->nd = c.n_dec_hidden + 1;
  h->encN.clear(); h->encK.clear(); h->decH.clear();
  for (int l = 0; l < h->ne; ++l) {
    const int n = (l < c.n_enc_hidden) ? c.enc_hidden[l] : 2 * h->L;
    const int k = (l == 0) ? h->K0p : c.enc_hidden[l - 1];
    MFVAE_CHECK(n % 8 == 0, "encoder hidden widths must be multiples of 8");
    h->encN.push_back(n); h->encK.push_back(k);
  }
  for (int l = 0; l < c.n_dec_hidden; ++l) {
    MFVAE_CHECK(c.dec_hidden[l] % 8 == 0, "decoder hidden widths must be multiples of 8");
    h->decH.push_back(c.dec_hidden[l]);
  }
  h->dtype = (c.precision == MFVAE_PREC_BF16) ? kBF16 : kF32;
  int engine = c.engine;
  if (engine == MFVAE_ENGINE_AUTO) engine = (h->dtype == kBF16) ? MFVAE_ENGINE_TCGEN05 : MFVAE_ENGINE_SIMT;
  MFVAE_CHECK(!(engine == MFVAE_ENGINE_TCGEN05 && h->dtype != kBF16), "the tcgen05 engine computes in bf16");
  h->use_tc = (engine == MFVAE_ENGINE_TCGEN05);

  // ---- arena ----
  int64_t cur = 0;
  h->table.clear();
  h->idx_emb = {take(cur, static_cast<int64_t>(h->A) * h->I), h->A, h->I, h->I};
  add_info(h, MFVAE_T_IDX_EMB, -1, -1, h->A, h->I, h->I, h->idx_emb.off);
  h->reg2_begin = cur;
  h->decW.resize(c.n_dec_hidden); h->decB.resize(c.n_dec_hidden);
  for (int l = 0; l < c.n_dec_hidden; ++l) {
    const int H = h->decH[l], K = (l == 0) ? h->Din : h->decH[l - 1];
    h->decW[l] = {take(cur, 2LL * H * K), 2 * H, K, K};
    h->decB[l] = {take(cur, 2LL * H), 1, 2 * H, 2 * H};
    add_info(h, MFVAE_T_SDEC_W, -1, l, H, K, K, h->decW[l].off);
    add_info(h, MFVAE_T_RDEC_W, -1, l, H, K, K, h->decW[l].off + static_cast<int64_t>(H) * K);
    add_info(h, MFVAE_T_SDEC_B, -1, l, 1, H, H, h->decB[l].off);
    add_info(h, MFVAE_T_RDEC_B, -1, l, 1, H, H, h->decB[l].off + H);
  }
  h->reg3_begin = cur;
  const int HL = h->decH.back();
  h->sOutW = {take(cur, static_cast<int64_t>(h->S) * HL), h->S, HL, HL};
  h->sOutB = {take(cur, h->S), 1, h->S, h->S};
  h->rOutW = {take(cur, static_cast<int64_t>(h->A) * HL), h->A, HL, HL};
  h->rOutB = {take(cur, h->A), 1, h->A, h->A};
  h->rlW = {take(cur, static_cast<int64_t>(h->A) * h->Ap), h->A, h->A, h->Ap};
  h->rlb = {take(cur, h->A), 1, h->A, h->A};
  add_info(h, MFVAE_T_SDEC_W, -1, c.n_dec_hidden, h->S, HL, HL, h->sOutW.off);
  add_info(h, MFVAE_T_SDEC_B, -1, c.n_dec_hidden, 1, h->S, h->S, h->sOutB.off);
  add_info(h, MFVAE_T_RDEC_W, -1, c.n_dec_hidden, h->A, HL, HL, h->rOutW.off);
  add_info(h, MFVAE_T_RDEC_B, -1, c.n_dec_hidden, 1, h->A, h->A, h->rOutB.off);
  add_info(h, MFVAE_T_RLIN_W, -1, -1, h->A, h->A, h->Ap, h->rlW.off);
  add_info(h, MFVAE_T_RLIN_B, -1, -1, 1, h->A, h->A, h->rlb.off);
  h->enc_begin = cur;
  h->encW.resize(h->ne); h->encB.resize(h->ne);
  for (int l = 0; l < h->ne; ++l) {
    const int N = h->encN[l], K = h->encK[l];
    h->encW[l] = {take(cur, static_cast<int64_t>(h->A) * N * K), h->A * N, K, K};
    h->encB[l] = {take(cur, static_cast<int64_t>(h->A) * N), h->A, N, N};
    for (int a = 0; a < h->A; ++a) {
      const int kc = (l == 0) ? h->I + h->obs_dim[a] : K;
      add_info(h, MFVAE_T_ENC_W, a, l, N, kc, K, h->encW[l].off + static_cast<int64_t>(a) * N * K);
      add_info(h, MFVAE_T_ENC_B, a, l, 1, N, N, h->encB[l].off + static_cast<int64_t>(a) * N);
    }
  }
  h->cont_act = c.continuous_act != 0;
  h->Hact = c.act_hidden > 0 ? c.act_hidden : 64;
  MFVAE_CHECK(h->Hact % 8 == 0, "act_hidden must be a multiple of 8");
  h->act_off.resize(h->A); h->act_total = 0;
  for (int a = 0; a < h->A; ++a) { h->act_off[a] = h->act_total; h->act_total += h->cont_act ? h->n_act[a] : 1; }
  if (!h->cont_act) {
    h->actT = {take(cur, static_cast<int64_t>(h->A) * h->nact_max * h->C), h->A * h->nact_max, h->C, h->C};
    for (int a = 0; a < h->A; ++a)
      add_info(h, MFVAE_T_ACT_TABLE, a, -1, h->n_act[a], h->C, h->C, h->actT.off + static_cast<int64_t>(a) * h->nact_max * h->C);
  } else {
    const int A = h->A, Hh = h->Hact, C = h->C;
    h->Kap = static_cast<int>(round_up(h->nact_max, 8));          // K of the first layer, zero padded (16-byte TMA rows)
    h->actW1 = {take(cur, static_cast<int64_t>(A) * Hh * h->Kap), A * Hh, h->Kap, h->Kap};
    h->actB1 = {take(cur, static_cast<int64_t>(A) * Hh), A, Hh, Hh};
    h->actW2 = {take(cur, static_cast<int64_t>(A) * C * Hh), A * C, Hh, Hh};
    h->actB2 = {take(cur, static_cast<int64_t>(A) * C), A, C, C};
    for (int a = 0; a < A; ++a) {
      add_info(h, MFVAE_T_ACTENC_W, a, 0, Hh, h->n_act[a], h->Kap, h->actW1.off + static_cast<int64_t>(a) * Hh * h->Kap);
      add_info(h, MFVAE_T_ACTENC_B, a, 0, 1, Hh, Hh, h->actB1.off + static_cast<int64_t>(a) * Hh);
      add_info(h, MFVAE_T_ACTENC_W, a, 1, C, Hh, Hh, h->actW2.off + static_cast<int64_t>(a) * C * Hh);
      add_info(h, MFVAE_T_ACTENC_B, a, 1, 1, C, C, h->actB2.off + static_cast<int64_t>(a) * C);
    }
  }
  h->arena_elems = cur;
  h->optimized_elems = c.optimize_encoders ? cur : h->enc_begin;
  return 0;
}

// ---- workspace ---------------------------------------------------------------------------------
static int64_t layout_workspace(MfvaeHandle_* h, int B) {
  const int64_t es = dtype_size(h->dtype);
  int64_t cur = 0;
  auto alloc = [&](int64_t bytes) { int64_t o = cur; cur += round_up(bytes, 256); return o; };
  auto mk = [&](int64_t groups, int64_t ld, int64_t esz, bool grouped_rows) {
    MfvaeHandle_::Buf b;
    b.ld = ld; b.gs = grouped_rows ? static_cast<int64_t>(B) * ld : 0;
    b.off = alloc(groups * B * ld * esz);
    return b;
  };
  const int A = h->A;
  h->X0 = mk(A, h->K0p, es, true);
  h->XE.clear(); h->DXE.clear(); h->HD.clear(); h->DHD.clear();
  for (int l = 0; l + 1 < h->ne; ++l) h->XE.push_back(mk(A, h->encN[l], es, true));
  h->LAT = mk(A, 2 * h->L, 4, true);
  h->ZIN = mk(1, h->fold_act ? h->Kzp + h->Kohp : h->Din, es, false);
  for (int l = 0; l < h->cfg.n_dec_hidden; ++l) h->HD.push_back(mk(1, 2 * h->decH[l], es, false));
  h->RS = mk(1, h->Sp, 4, false);
  h->RR0 = mk(1, h->Ap, es, false);
  h->RR = mk(1, h->Ap, 4, false);
  h->DRS = mk(1, h->Sp, es, false);
  h->DRR = mk(1, h->Ap, es, false);
  h->DRR0 = mk(1, h->Ap, es, false);
  for (int l = 0; l < h->cfg.n_dec_hidden; ++l) h->DHD.push_back(mk(1, 2 * h->decH[l], es, false));
  h->GZIN = mk(1, h->fold_act ? h->Kzp : h->Din, es, false);
  h->DLAT = mk(A, 2 * h->L, es, true);
  for (int l = 0; l + 1 < h->ne; ++l) h->DXE.push_back(mk(A, h->encN[l], es, true));
  h->GX0 = mk(A, h->Ip, es, true);
  if (h->cont_act) { h->ACT0 = mk(A, h->Kap, es, true); h->HA = mk(A, h->Hact, es, true); h->DHA = mk(A, h->Hact, es, true); }
  h->X0F = mk(A, h->K0f, es, true);
  if (h->fold_act) {
    const int64_t rows = 2LL * h->decH[0];
    h->TACT.ld = h->Kohp; h->TACT.gs = 0; h->TACT.off = alloc(rows * h->Kohp * es);
    h->DTACT.ld = h->Kohp; h->DTACT.gs = 0; h->DTACT.off = alloc(rows * h->Kohp * 4);
  }
  h->EB0.ld = h->encN[0]; h->EB0.gs = h->encN[0]; h->EB0.off = alloc(static_cast<int64_t>(A) * h->encN[0] * 4);
  h->off_losses = alloc(64 * sizeof(float));
  h->off_scratch = alloc(3 * 4096 * sizeof(float));
  return cur;
}

static void free_plans(MfvaeHandle_* h) {
  for (TcPlan* p : h->tc) if (p) gemm_tc_free(p);
  if (h->enc_fused) { enc_fused_free(h->enc_fused); h->enc_fused = nullptr; }
  h->tc.clear(); h->gemms.clear();
  h->g_enc_fwd.clear(); h->g_enc_wg.clear(); h->g_enc_dg.clear();
  h->g_dec_fwd.clear(); h->g_dec_wg.clear(); h->g_dec_dg.clear();
}

static int pick_split_k(int G, int M, int N, int K) {
  const int64_t tiles = static_cast<int64_t>(G) * ((M + 127) / 128) * ((N + 127) / 128);
  int64_t want = (2LL * kNumSMs + tiles - 1) / tiles;
  int64_t maxs = std::max(1, K / 256);
  return static_cast<int>(std::max<int64_t>(1, std::min(want, maxs)));
}

static int build_ops(MfvaeHandle_* h) {
  free_plans(h);
  const int B = h->B, A = h->A, dt = h->dtype;
  const int64_t es = dtype_size(dt);
  char* ws = h->ws;
  const char* W = (dt == kBF16) ? static_cast<const char*>(h->ar.d_shadow_bf16) : reinterpret_cast<const char*>(h->ar.d_param);
  MFVAE_CHECK(W != nullptr, "arenas are not bound (bf16 precision needs d_shadow_bf16)");
  float* P = h->ar.d_param; float* Gd = h->ar.d_grad;
  MFVAE_CHECK(P && Gd, "arenas are not bound");
  auto wptr = [&](int64_t off) { return static_cast<const void*>(W + off * es); };
  auto buf = [&](const MfvaeHandle_::Buf& b) { return static_cast<void*>(ws + b.off); };
  auto push = [&](const GemmOp& op) { h->gemms.push_back(op); return static_cast<int>(h->gemms.size()) - 1; };
  auto fwd = [&](int G, int M, int N, int K, const void* Ain, int64_t a_gs, int64_t a_ld, const void* Wt, int64_t w_gs, int64_t w_ld,
                 void* Cout, int64_t c_gs, int64_t c_ld, int c_dtype, const float* bias, int64_t bias_gs, bool relu) {
    GemmOp o; o.G = G; o.M = M; o.N = N; o.K = K; o.dtype = dt;
    o.A = Ain; o.a_gs = a_gs; o.a_rs = a_ld; o.a_cs = 1;
    o.B = Wt; o.b_gs = w_gs; o.b_rs = w_ld; o.b_cs = 1;
    o.C = Cout; o.c_gs = c_gs; o.c_ld = c_ld; o.c_dtype = c_dtype;
    o.bias = bias; o.bias_gs = bias_gs; o.epi = relu ? kEpiBiasRelu : kEpiBias;
    return push(o);
  };
  // dW[g] (N_out x K_in) += D[g]^T (N_out x B) * X[g] (B x K_in)
  auto wgrad_to = [&](int G, int Nout, int Kin, const void* D, int64_t d_gs, int64_t d_ld, const void* X, int64_t x_gs, int64_t x_ld,
                      float* out, int64_t gw_gs, int64_t gw_ld) {
    GemmOp o; o.G = G; o.M = Nout; o.N = Kin; o.K = B; o.dtype = dt;
    o.A = D; o.a_gs = d_gs; o.a_rs = 1; o.a_cs = d_ld;
    o.B = X; o.b_gs = x_gs; o.b_rs = 1; o.b_cs = x_ld;
    o.C = out; o.c_gs = gw_gs; o.c_ld = gw_ld; o.c_dtype = kF32;
    o.epi = kEpiAccum; o.split_k = h->use_tc ? 1 : pick_split_k(G, Nout, Kin, B);     // the tcgen05 planner sizes its own split-K
    return push(o);
  };
  auto wgrad = [&](int G, int Nout, int Kin, const void* D, int64_t d_gs, int64_t d_ld, const void* X, int64_t x_gs, int64_t x_ld,
                   int64_t gw_off, int64_t gw_gs, int64_t gw_ld) {
    return wgrad_to(G, Nout, Kin, D, d_gs, d_ld, X, x_gs, x_ld, Gd + gw_off, gw_gs, gw_ld);
  };
  // dX[g] (B x K_in) = D[g] (B x N_out) * W[g] (N_out x K_in)  [* (mask > 0)]
  auto dgrad = [&](int G, int Kin, int Nout, const void* D, int64_t d_gs, int64_t d_ld, const void* Wt, int64_t w_gs, int64_t w_ld,
                   void* dX, int64_t dx_gs, int64_t dx_ld, const void* mask, int64_t m_gs, int64_t m_ld) {
    GemmOp o; o.G = G; o.M = B; o.N = Kin; o.K = Nout; o.dtype = dt;
    o.A = D; o.a_gs = d_gs; o.a_rs = d_ld; o.a_cs = 1;
    o.B = Wt; o.b_gs = w_gs; o.b_rs = 1; o.b_cs = w_ld;
    o.C = dX; o.c_gs = dx_gs; o.c_ld = dx_ld; o.c_dtype = dt;
    o.epi = mask ? kEpiReluMask : kEpiNone; o.aux = mask; o.aux_gs = m_gs; o.aux_ld = m_ld;
    return push(o);
  };

  // ---- encoders (G = A) ----
  for (int l = 0; l < h->ne; ++l) {
    const int N = h->encN[l], K = h->encK[l];
    const MfvaeHandle_::Buf& in = (l == 0) ? h->X0 : h->XE[l - 1];
    const bool last = (l + 1 == h->ne);
    const MfvaeHandle_::Buf& out = last ? h->LAT : h->XE[l];
    h->g_enc_fwd.push_back(fwd(A, B, N, K, buf(in), in.gs, in.ld, wptr(h->encW[l].off), static_cast<int64_t>(N) * K, K,
                               buf(out), out.gs, out.ld, last ? kF32 : dt, P + h->encB[l].off, N, !last));
  }
  for (int l = 0; l < h->ne; ++l) {
    const int N = h->encN[l], K = h->encK[l];
    const MfvaeHandle_::Buf& in = (l == 0) ? h->X0 : h->XE[l - 1];
    const MfvaeHandle_::Buf& D = (l + 1 == h->ne) ? h->DLAT : h->DXE[l];
    h->g_enc_wg.push_back(wgrad(A, N, K, buf(D), D.gs, D.ld, buf(in), in.gs, in.ld, h->encW[l].off, static_cast<int64_t>(N) * K, K));
    if (l == 0) {   // only the id-embedding columns of dX0 are needed
      h->g_enc_dg.push_back(dgrad(A, h->I, N, buf(D), D.gs, D.ld, wptr(h->encW[0].off), static_cast<int64_t>(N) * K, K,
                                  buf(h->GX0), h->GX0.gs, h->GX0.ld, nullptr, 0, 0));
    } else {
      const MfvaeHandle_::Buf& dx = h->DXE[l - 1];
      h->g_enc_dg.push_back(dgrad(A, K, N, buf(D), D.gs, D.ld, wptr(h->encW[l].off), static_cast<int64_t>(N) * K, K,
                                  buf(dx), dx.gs, dx.ld, buf(in), in.gs, in.ld));
    }
  }
  // encoder layer 0 with the id-embedding folded into a per-agent bias (fold.cu): input [obs | 0], weight columns [I, K0p)
  {
    const int N = h->encN[0], K = h->encK[0], I = h->I;
    const MfvaeHandle_::Buf& out = (h->ne == 1) ? h->LAT : h->XE[0];
    const MfvaeHandle_::Buf& D = (h->ne == 1) ? h->DLAT : h->DXE[0];
    const float* eb = reinterpret_cast<const float*>(ws + h->EB0.off);
    h->g_enc_fwd0_f = fwd(A, B, N, h->K0f, buf(h->X0F), h->X0F.gs, h->X0F.ld, wptr(h->encW[0].off + I), static_cast<int64_t>(N) * K, K,
                          buf(out), out.gs, out.ld, (h->ne == 1) ? kF32 : dt, eb, N, h->ne > 1);
    h->g_enc_wg0_f = wgrad(A, N, h->K0f, buf(D), D.gs, D.ld, buf(h->X0F), h->X0F.gs, h->X0F.ld, h->encW[0].off + I, static_cast<int64_t>(N) * K, K);
  }
  // ---- ActionEncoder (continuous actions, G = A): D_a -> Hact (ReLU) -> C, written straight into its ZIN columns ----
  if (h->cont_act) {
    const int Hh = h->Hact, C = h->C, Kap = h->Kap;
    char* zact = ws + h->ZIN.off + static_cast<int64_t>(A) * h->L * es;       // column A*L of ZIN: agent a at + a*C
    char* gzact = ws + h->GZIN.off + static_cast<int64_t>(A) * h->L * es;
    h->g_act_fwd1 = fwd(A, B, Hh, Kap, buf(h->ACT0), h->ACT0.gs, h->ACT0.ld, wptr(h->actW1.off), static_cast<int64_t>(Hh) * Kap, Kap,
                        buf(h->HA), h->HA.gs, h->HA.ld, dt, P + h->actB1.off, Hh, true);
    h->g_act_fwd2 = fwd(A, B, C, Hh, buf(h->HA), h->HA.gs, h->HA.ld, wptr(h->actW2.off), static_cast<int64_t>(C) * Hh, Hh,
                        zact, C, h->ZIN.ld, dt, P + h->actB2.off, C, false);
    h->g_act_wg2 = wgrad(A, C, Hh, gzact, C, h->GZIN.ld, buf(h->HA), h->HA.gs, h->HA.ld, h->actW2.off, static_cast<int64_t>(C) * Hh, Hh);
    h->g_act_dg2 = dgrad(A, Hh, C, gzact, C, h->GZIN.ld, wptr(h->actW2.off), static_cast<int64_t>(C) * Hh, Hh,
                         buf(h->DHA), h->DHA.gs, h->DHA.ld, buf(h->HA), h->HA.gs, h->HA.ld);
    h->g_act_wg1 = wgrad(A, Hh, Kap, buf(h->DHA), h->DHA.gs, h->DHA.ld, buf(h->ACT0), h->ACT0.gs, h->ACT0.ld,
                         h->actW1.off, static_cast<int64_t>(Hh) * Kap, Kap);
  }
  // ---- decoders: hidden layers (layer 0 fused over both decoders, others G = 2 over column halves) ----
  const int nh = h->cfg.n_dec_hidden;
  for (int l = 0; l < nh; ++l) {
    const int H = h->decH[l];
    if (l == 0 && h->fold_act) {
      // [z | 0 | onehot(act)] . [W0 z-columns | T]^T: the action-embedding columns are K = A*n_act one-hot columns (fold.cu)
      const int i0 = fwd(1, B, 2 * H, h->Kzp + h->Koh, buf(h->ZIN), 0, h->ZIN.ld, wptr(h->decW[0].off), 0, h->Din,
                         buf(h->HD[0]), 0, h->HD[0].ld, dt, P + h->decB[0].off, 0, true);
      GemmOp& o = h->gemms[i0];
      o.B2 = ws + h->TACT.off; o.b2_gs = 0; o.b2_rs = h->TACT.ld; o.k_split = h->Kzp; o.k1 = h->Kz; o.k2 = h->Koh;
      h->g_dec_fwd.push_back(i0);
      h->g_dec_wg.push_back(wgrad(1, 2 * H, h->Kz, buf(h->DHD[0]), 0, h->DHD[0].ld, buf(h->ZIN), 0, h->ZIN.ld, h->decW[0].off, 0, h->Din));
      h->g_dec_wgT = wgrad_to(1, 2 * H, h->Koh, buf(h->DHD[0]), 0, h->DHD[0].ld, ws + h->ZIN.off + static_cast<int64_t>(h->Kzp) * es, 0, h->ZIN.ld,
                              reinterpret_cast<float*>(ws + h->DTACT.off), 0, h->DTACT.ld);
      h->g_dec_dg.push_back(dgrad(1, h->Kz, 2 * H, buf(h->DHD[0]), 0, h->DHD[0].ld, wptr(h->decW[0].off), 0, h->Din,
                                  buf(h->GZIN), 0, h->GZIN.ld, nullptr, 0, 0));
    } else if (l == 0) {
      h->g_dec_fwd.push_back(fwd(1, B, 2 * H, h->Din, buf(h->ZIN), 0, h->ZIN.ld, wptr(h->decW[0].off), 0, h->Din,
                                 buf(h->HD[0]), 0, h->HD[0].ld, dt, P + h->decB[0].off, 0, true));
      h->g_dec_wg.push_back(wgrad(1, 2 * H, h->Din, buf(h->DHD[0]), 0, h->DHD[0].ld, buf(h->ZIN), 0, h->ZIN.ld,
                                  h->decW[0].off, 0, h->Din));
      h->g_dec_dg.push_back(dgrad(1, h->Din, 2 * H, buf(h->DHD[0]), 0, h->DHD[0].ld, wptr(h->decW[0].off), 0, h->Din,
                                  buf(h->GZIN), 0, h->GZIN.ld, nullptr, 0, 0));
    } else {
      const int K = h->decH[l - 1];
      char* in = ws + h->HD[l - 1].off; char* out = ws + h->HD[l].off;
      char* D = ws + h->DHD[l].off; char* dx = ws + h->DHD[l - 1].off;
      h->g_dec_fwd.push_back(fwd(2, B, H, K, in, K, h->HD[l - 1].ld, wptr(h->decW[l].off), static_cast<int64_t>(H) * K, K,
                                 out, H, h->HD[l].ld, dt, P + h->decB[l].off, H, true));
      h->g_dec_wg.push_back(wgrad(2, H, K, D, H, h->DHD[l].ld, in, K, h->HD[l - 1].ld, h->decW[l].off, static_cast<int64_t>(H) * K, K));
      h->g_dec_dg.push_back(dgrad(2, K, H, D, H, h->DHD[l].ld, wptr(h->decW[l].off), static_cast<int64_t>(H) * K, K,
                                  dx, K, h->DHD[l - 1].ld, in, K, h->HD[l - 1].ld));
    }
  }
  // ---- output layers ----
  {
    const int HL = h->decH.back();
    char* hs = ws + h->HD[nh - 1].off;                    // state half
    char* hr = hs + static_cast<int64_t>(HL) * es;        // reward half
    char* dhs = ws + h->DHD[nh - 1].off; char* dhr = dhs + static_cast<int64_t>(HL) * es;
    const int64_t hld = h->HD[nh - 1].ld;
    h->g_sout_fwd = fwd(1, B, h->S, HL, hs, 0, hld, wptr(h->sOutW.off), 0, HL, buf(h->RS), 0, h->RS.ld, kF32, P + h->sOutB.off, 0, false);
    if (h->use_tc) {
      h->g_sout_loss = fwd(1, B, h->S, HL, hs, 0, hld, wptr(h->sOutW.off), 0, HL, buf(h->DRS), 0, h->DRS.ld, dt, P + h->sOutB.off, 0, false);
      h->gemms[h->g_sout_loss].epi = kEpiLossGrad;
      h->g_sout_fwd16 = fwd(1, B, h->S, HL, hs, 0, hld, wptr(h->sOutW.off), 0, HL, buf(h->DRS), 0, h->DRS.ld, dt, P + h->sOutB.off, 0, false);
    }
    h->g_rout_fwd = fwd(1, B, A, HL, hr, 0, hld, wptr(h->rOutW.off), 0, HL, buf(h->RR0), 0, h->RR0.ld, dt, P + h->rOutB.off, 0, false);
    h->g_rl_fwd = fwd(1, B, A, A, buf(h->RR0), 0, h->RR0.ld, wptr(h->rlW.off), 0, h->Ap, buf(h->RR), 0, h->RR.ld, kF32, P + h->rlb.off, 0, false);
    h->g_rl_wg = wgrad(1, A, A, buf(h->DRR), 0, h->DRR.ld, buf(h->RR0), 0, h->RR0.ld, h->rlW.off, 0, h->Ap);
    h->g_rl_dg = dgrad(1, A, A, buf(h->DRR), 0, h->DRR.ld, wptr(h->rlW.off), 0, h->Ap, buf(h->DRR0), 0, h->DRR0.ld, nullptr, 0, 0);
    h->g_rout_wg = wgrad(1, A, HL, buf(h->DRR0), 0, h->DRR0.ld, hr, 0, hld, h->rOutW.off, 0, HL);
    h->g_rout_dg = dgrad(1, HL, A, buf(h->DRR0), 0, h->DRR0.ld, wptr(h->rOutW.off), 0, HL, dhr, 0, hld, hr, 0, hld);
    h->g_sout_wg = wgrad(1, h->S, HL, buf(h->DRS), 0, h->DRS.ld, hs, 0, hld, h->sOutW.off, 0, HL);
    h->g_sout_dg = dgrad(1, HL, h->S, buf(h->DRS), 0, h->DRS.ld, wptr(h->sOutW.off), 0, HL, dhs, 0, hld, hs, 0, hld);
  }
  h->tc.assign(h->gemms.size(), nullptr);
  if (h->use_tc) {
    for (size_t i = 0; i < h->gemms.size(); ++i) MFVAE_TRY(gemm_tc_plan(h->gemms[i], &h->tc[i]));
  }
  // ---- fused encoder chain (reads the staged X0F tile and the folded layer-0 bias) ----
  if (h->use_tc && (h->cfg.fusion & MFVAE_FUSE_ENCODER) && !(h->cfg.fusion & MFVAE_FUSE_NOFOLD_IDX) && h->ne <= kEncMaxL) {
    EncFusedDesc d;
    d.A = A; d.B = B; d.nl = h->ne; d.L = h->L; d.K0 = h->K0f;
    for (int l = 0; l < h->ne; ++l) {
      const MfvaeHandle_::Buf& in = (l == 0) ? h->X0F : h->XE[l - 1];
      d.N[l] = h->encN[l];
      d.W[l] = wptr(h->encW[l].off + (l == 0 ? h->I : 0));
      d.w_ld[l] = h->encK[l]; d.w_gs[l] = static_cast<int64_t>(h->encN[l]) * h->encK[l];
      d.bias[l] = (l == 0) ? reinterpret_cast<const float*>(ws + h->EB0.off) : P + h->encB[l].off;
      d.X[l] = buf(in); d.x_ld[l] = in.ld; d.x_gs[l] = in.gs;
    }
    d.lat = reinterpret_cast<float*>(ws + h->LAT.off); d.lat_gs = h->LAT.gs; d.lat_ld = h->LAT.ld;
    d.zin = ws + h->ZIN.off; d.zin_ld = h->ZIN.ld;
    if (enc_fused_applicable(d)) MFVAE_TRY(enc_fused_plan(d, &h->enc_fused));
  }
  return 0;
}

static int run_gemm(MfvaeHandle_* h, int i, cudaStream_t s) {
  const bool prof = h->profiling && static_cast<size_t>(2 * i + 1) < h->prof_ev.size();
  if (prof) MFVAE_CUDA(cudaEventRecord(h->prof_ev[2 * i], s));
  int rc = h->use_tc ? gemm_tc_run(h->tc[i], s) : gemm_simt(h->gemms[i], s);
  if (prof && rc == 0) { MFVAE_CUDA(cudaEventRecord(h->prof_ev[2 * i + 1], s)); h->prof_hit[i] = 1; }
  return rc;
}

static int check_ready(MfvaeHandle_* h, const MfvaeBatch* b) {
  MFVAE_CHECK(h != nullptr && b != nullptr, "null handle or batch");
  MFVAE_CHECK(h->device >= 0, "layout-only handle: there is no CPU path, create the handle on a CUDA device");
  MFVAE_CHECK(h->ws != nullptr, "workspace is not bound");
  MFVAE_CHECK(b->batch == h->B, "batch size differs from the bound workspace");
  MFVAE_CHECK(b->d_obs && b->d_act, "batch needs d_obs and d_act");
  MFVAE_CHECK(b->batch_global >= b->batch, "batch_global must be >= batch");
  return 0;
}

static float* losses_ptr(MfvaeHandle_* h) { return reinterpret_cast<float*>(h->ws + h->off_losses); }
static float* scratch_ptr(MfvaeHandle_* h, int i) { return reinterpret_cast<float*>(h->ws + h->off_scratch) + 4096 * i; }

static bool use_aux(const MfvaeHandle_* h) { return h->aux != nullptr && !h->profiling; }

// action-embedding half of the decoder input: independent of the encoders, so it runs beside them on the aux stream
static int act_embed_on(MfvaeHandle_* h, const StageArgs& st, cudaStream_t q) {
  if (h->fold_act) {
    // one-hot action columns + T = W0_act . tables (from the fp32 masters, every step: both factors train)
    MFVAE_TRY(launch_onehot(st.act, st.act_ld, h->d_meta + 2 * h->A, h->ws + h->ZIN.off, h->dtype, h->ZIN.ld, h->Kzp, h->A, h->nact_max,
                            h->Kohp, h->B, q));
    return launch_act_fold_fwd(h->ar.d_param + h->decW[0].off, h->Din, h->Kz, h->ar.d_param + h->actT.off,
                               static_cast<int64_t>(h->nact_max) * h->C, h->ws + h->TACT.off, h->dtype, h->TACT.ld, 2 * h->decH[0], h->A, h->C,
                               h->nact_max, q);
  }
  if (!h->cont_act) return launch_stage(st, q, false, true);
  // ActionEncoder MLP (reference model.py:148)
  MFVAE_TRY(launch_stage_actions(st.act, h->act_total, h->d_meta + 3 * h->A, h->d_meta + 2 * h->A, h->ws + h->ACT0.off, h->dtype,
                                 h->A, h->B, h->Kap, q));
  MFVAE_TRY(run_gemm(h, h->g_act_fwd1, q));
  return run_gemm(h, h->g_act_fwd2, q);
}

// encoder layer 0 reads [obs | 0] and a per-agent bias b0 + W0[:, :I] . emb[a] when the batch carries the codebook agent index
static bool fold_idx(const MfvaeHandle_* h, const MfvaeBatch* b) {
  return b->d_idx == nullptr && !(h->cfg.fusion & MFVAE_FUSE_NOFOLD_IDX);
}
static int enc_bias_on(MfvaeHandle_* h, cudaStream_t q) {
  return launch_enc_bias_fold(h->ar.d_param + h->encW[0].off, h->encK[0], h->ar.d_param + h->encB[0].off, h->ar.d_param + h->idx_emb.off,
                              reinterpret_cast<float*>(h->ws + h->EB0.off), h->A, h->encN[0], h->I, q);
}

static int do_forward_act_embed(MfvaeHandle_* h, const StageArgs& st, cudaStream_t s, bool fold_i) {
  if (!use_aux(h)) {
    if (fold_i) MFVAE_TRY(enc_bias_on(h, s));
    if (h->opt_pending) { MFVAE_CUDA(cudaStreamWaitEvent(s, h->opt_ev, 0)); h->opt_pending = false; }
    return act_embed_on(h, st, s);
  }
  MFVAE_CUDA(cudaEventRecord(h->aux_fork_ev, s));          // orders it after whatever last read ZIN on the caller's stream
  MFVAE_CUDA(cudaStreamWaitEvent(h->aux, h->aux_fork_ev, 0));
  if (fold_i) {
    MFVAE_TRY(enc_bias_on(h, h->aux));
    MFVAE_CUDA(cudaEventRecord(h->eb_ev, h->aux));
  }
  // a pipelined optimizer sweep of the decoder block (mfvae_allreduce_grads without mfvae_opt_join) may still be running:
  // the encoder half of this forward does not read what it writes, the action fold (decoder layer 0's master weights) does
  if (h->opt_pending) MFVAE_CUDA(cudaStreamWaitEvent(h->aux, h->opt_ev, 0));
  MFVAE_TRY(act_embed_on(h, st, h->aux));
  MFVAE_CUDA(cudaEventRecord(h->aux_join_ev, h->aux));
  return 0;
}

// loss_batch != nullptr (train step): the state output layer runs with the loss epilogue -- recon_s is not materialised,
// D(recon_s) and the per-warp loss partials come straight out of the GEMM (the target must be bound in the batch).
static int do_forward_decoders(MfvaeHandle_* h, cudaStream_t s, const MfvaeBatch* loss_batch = nullptr, int huber = 1, bool recon16 = false,
                               bool defer_reward_join = false) {
  const bool aux = use_aux(h);
  if (h->opt_pending) { MFVAE_CUDA(cudaStreamWaitEvent(s, h->opt_ev, 0)); h->opt_pending = false; }
  if (aux) MFVAE_CUDA(cudaStreamWaitEvent(s, h->aux_join_ev, 0));     // action embeddings are in ZIN
  for (int l = 0; l < h->cfg.n_dec_hidden; ++l) MFVAE_TRY(run_gemm(h, h->g_dec_fwd[l], s));
  // reward head (two tiny GEMMs) beside the state output layer.  It is enqueued FIRST: the output layer is a persistent
  // kernel that claims every SM for ~50 us, and whatever is launched behind it starves until it retires.
  cudaStream_t r = s;
  if (aux) {
    MFVAE_CUDA(cudaEventRecord(h->aux_fork2_ev, s));
    MFVAE_CUDA(cudaStreamWaitEvent(h->aux, h->aux_fork2_ev, 0));
    r = h->aux;
  }
  MFVAE_TRY(run_gemm(h, h->g_rout_fwd, r));
  MFVAE_TRY(run_gemm(h, h->g_rl_fwd, r));
  if (aux) MFVAE_CUDA(cudaEventRecord(h->aux_join2_ev, h->aux));
  if (loss_batch) {
    const double cs = static_cast<double>(loss_batch->batch_global) * h->S;
    MFVAE_TRY(gemm_tc_set_loss(h->tc[h->g_sout_loss], loss_batch->d_next, h->S, static_cast<float>(static_cast<double>(h->s_weight) / cs), huber,
                               scratch_ptr(h, 1)));
    MFVAE_TRY(run_gemm(h, h->g_sout_loss, s));
  } else if (recon16) {
    MFVAE_TRY(run_gemm(h, h->g_sout_fwd16, s));
  } else {
    MFVAE_TRY(run_gemm(h, h->g_sout_fwd, s));
  }
  if (aux && !defer_reward_join) MFVAE_CUDA(cudaStreamWaitEvent(s, h->aux_join2_ev, 0));
  return 0;
}

static int do_forward(MfvaeHandle_* h, const MfvaeBatch* b, MfvaeOutputs* out, cudaStream_t s, bool fuse_loss = false, bool recon16 = false,
                      bool defer_reward_join = false) {
  MFVAE_TRY(check_ready(h, b));
  const MfvaeBatch* lb = fuse_loss ? b : nullptr;
  StageArgs st{};
  st.obs = b->d_obs; st.obs_ld = h->S; st.obs_dtype = b->obs_bf16 ? kBF16 : kF32; st.act = b->d_act; st.act_ld = h->cont_act ? h->act_total : h->A; st.idx = b->d_idx;
  st.idx_emb = h->ar.d_param + h->idx_emb.off;
  st.act_table = h->cont_act ? nullptr : h->ar.d_param + h->actT.off; st.act_table_gs = static_cast<int64_t>(h->nact_max) * h->C; st.n_act_max = h->nact_max;
  st.obs_off = h->d_meta; st.obs_dim = h->d_meta + h->A; st.n_act = h->d_meta + 2 * h->A;
  st.x0 = h->ws + h->X0.off; st.x0_ld = static_cast<int>(h->X0.ld); st.x0_gs = h->X0.gs;
  st.zin = h->ws + h->ZIN.off; st.zin_ld = static_cast<int>(h->ZIN.ld);
  st.A = h->A; st.I = h->I; st.L = h->L; st.C = h->C; st.B = h->B; st.dtype = h->dtype;
  const float* lat = reinterpret_cast<const float*>(h->ws + h->LAT.off);
  if (out) {
    out->d_recon_s = reinterpret_cast<const float*>(h->ws + h->RS.off); out->recon_s_ld = static_cast<int32_t>(h->RS.ld);
    out->d_recon_r = reinterpret_cast<const float*>(h->ws + h->RR.off); out->recon_r_ld = static_cast<int32_t>(h->RR.ld);
    out->d_latent = lat; out->d_losses = losses_ptr(h);
    if (fuse_loss || recon16) { out->d_recon_s = nullptr; out->recon_s_ld = 0; }      // not materialised in fp32 on this path
  }
  const bool fold_i = fold_idx(h, b);
  MFVAE_TRY(do_forward_act_embed(h, st, s, fold_i));
  if (fold_i) {       // X0F[a][b][:] = [obs_a | 0]: the staging kernel with a zero-width embedding block
    st.I = 0; st.idx = nullptr;
    st.x0 = h->ws + h->X0F.off; st.x0_ld = static_cast<int>(h->X0F.ld); st.x0_gs = h->X0F.gs;
  }
  MFVAE_TRY(launch_stage(st, s, true, false));
  if (fold_i && use_aux(h)) MFVAE_CUDA(cudaStreamWaitEvent(s, h->eb_ev, 0));
  if (fold_i && h->enc_fused) {
    // the four encoder layers, reparameterisation and KL as ONE kernel (enc_fused.cu) on the staged tile
    EncFwdBatch eb{};
    eb.eps = b->d_eps; eb.eps_ld = static_cast<int64_t>(h->A) * h->L; eb.seed = b->seed; eb.step = b->step; eb.sample0 = b->sample0;
    eb.kl_scale = 1.0f / static_cast<float>(b->batch_global); eb.kl_out = losses_ptr(h) + 3; eb.scratch = scratch_ptr(h, 0);
    if (h->profiling && h->enc_prof_ev[0]) MFVAE_CUDA(cudaEventRecord(h->enc_prof_ev[0], s));
    MFVAE_TRY(enc_fused_forward(h->enc_fused, eb, s));
    if (h->profiling && h->enc_prof_ev[0]) { MFVAE_CUDA(cudaEventRecord(h->enc_prof_ev[1], s)); h->enc_prof_hit = true; }
    return do_forward_decoders(h, s, lb, h->cfg.huber, recon16, defer_reward_join);
  }
  MFVAE_TRY(run_gemm(h, fold_i ? h->g_enc_fwd0_f : h->g_enc_fwd[0], s));
  for (int l = 1; l < h->ne; ++l) MFVAE_TRY(run_gemm(h, h->g_enc_fwd[l], s));
  ReparamArgs rp{};
  rp.mu = lat; rp.lv = lat + h->L; rp.lat_as = h->LAT.gs; rp.lat_bs = h->LAT.ld;
  rp.eps = b->d_eps; rp.eps_ld = static_cast<int64_t>(h->A) * h->L;
  rp.z = h->ws + h->ZIN.off; rp.z_ld = h->ZIN.ld; rp.z_dtype = h->dtype;
  rp.B = h->B; rp.A = h->A; rp.L = h->L; rp.seed = b->seed; rp.step = b->step; rp.sample0 = b->sample0;
  rp.kl_scale = 1.0f / static_cast<float>(b->batch_global);
  rp.kl_out = losses_ptr(h) + 3; rp.scratch = scratch_ptr(h, 0);
  MFVAE_TRY(launch_reparam_kl_fwd(rp, s));
  return do_forward_decoders(h, s, lb, h->cfg.huber, recon16, defer_reward_join);
}

static int do_loss(MfvaeHandle_* h, const MfvaeBatch* b, int loss_kind, cudaStream_t s, bool state_fused = false, bool recon16 = false,
                   bool join_reward = false, bool fuse_bias = false) {
  MFVAE_TRY(check_ready(h, b));
  MFVAE_CHECK(loss_kind >= MFVAE_LOSS_DEFAULT && loss_kind <= MFVAE_LOSS_JOINT_MSE, "unknown loss kind");
  const int joint_mse = (loss_kind == MFVAE_LOSS_JOINT_MSE);
  const int use_huber = (loss_kind == MFVAE_LOSS_DEFAULT) ? h->cfg.huber : (loss_kind == MFVAE_LOSS_HUBER);
  MFVAE_CHECK(b->d_next && b->d_rew, "loss needs d_next and d_rew");
  const double Bg = static_cast<double>(b->batch_global);
  ReconLossArgs a{};
  a.recon = reinterpret_cast<const float*>(h->ws + h->RS.off); a.recon_ld = h->RS.ld;
  a.recon16 = recon16 ? reinterpret_cast<const __nv_bfloat16*>(h->ws + h->DRS.off) : nullptr;   // in place over D(recon_s)
  a.target = b->d_next; a.target_ld = h->S;
  a.target16 = b->next_bf16 ? reinterpret_cast<const __nv_bfloat16*>(b->d_next) : nullptr;
  a.grad = h->ws + h->DRS.off; a.grad_ld = h->DRS.ld; a.grad_dtype = h->dtype;
  a.B = h->B; a.width = h->S;
  const double cs = joint_mse ? Bg * (h->S + h->A) : Bg * h->S;
  const double cr = joint_mse ? Bg * (h->S + h->A) : Bg * h->A;
  const float rw = joint_mse ? 1.0f : h->cfg.r_weight;
  a.huber = joint_mse ? 0 : use_huber;
  const float sw = joint_mse ? 1.0f : h->s_weight;
  a.grad_scale = static_cast<float>(static_cast<double>(sw) / cs); a.loss_scale = static_cast<float>(1.0 / cs);
  a.loss_out = losses_ptr(h) + 1; a.scratch = scratch_ptr(h, 1);
  h->sout_bias_done = false;
  if (!state_fused && fuse_bias && h->grads_zeroed) {
    // train step: the gradient arena is already being zeroed beside the forward pass, so the loss kernel can leave the state
    // head's bias gradient (column sums of d recon_s) behind and backward skips that 46 MB re-read
    MFVAE_CUDA(cudaStreamWaitEvent(s, h->zero_ev, 0));
    a.colsum_out = h->ar.d_grad + h->sOutB.off;
    h->sout_bias_done = true;
  }
  if (!state_fused) MFVAE_TRY(launch_recon_loss(a, s));
  a.colsum_out = nullptr;
  if (join_reward && use_aux(h)) MFVAE_CUDA(cudaStreamWaitEvent(s, h->aux_join2_ev, 0));   // reward head (aux stream) is needed from here
  a.recon = reinterpret_cast<const float*>(h->ws + h->RR.off); a.recon_ld = h->RR.ld; a.recon16 = nullptr;
  a.target = b->d_rew; a.target_ld = h->A; a.target16 = nullptr;
  a.grad = h->ws + h->DRR.off; a.grad_ld = h->DRR.ld; a.width = h->A;
  a.grad_scale = static_cast<float>(static_cast<double>(rw) / cr); a.loss_scale = static_cast<float>(1.0 / cr);
  a.loss_out = losses_ptr(h) + 2; a.scratch = scratch_ptr(h, 2);
  MFVAE_TRY(launch_recon_loss(a, s));
  if (state_fused)
    MFVAE_TRY(launch_loss_total(losses_ptr(h), sw, rw, h->cfg.kl_weight, s, scratch_ptr(h, 1), gemm_tc_loss_partials(h->tc[h->g_sout_loss]),
                                static_cast<float>(1.0 / cs)));
  else
    MFVAE_TRY(launch_loss_total(losses_ptr(h), sw, rw, h->cfg.kl_weight, s));
  if (h->loss_ev) MFVAE_CUDA(cudaEventRecord(h->loss_ev, s));
  return 0;
}

// optimizer.zero_grad(): split-K wgrads and bias column sums accumulate with fp32 atomics.  The two largest weight
// gradients (decoder layer 0, state output layer) are written by single-split wgrads with plain stores and are
// skipped (they are 2/3 of the arena).
static int zero_grads(MfvaeHandle_* h, cudaStream_t q) {
  float* G = h->ar.d_grad;
  int64_t skip[2][2]; int ns = 0;
  if (h->use_tc && gemm_tc_overwrites(h->tc[h->g_dec_wg[0]])) { skip[ns][0] = h->decW[0].off; skip[ns][1] = h->decB[0].off; ++ns; }
  if (h->use_tc && gemm_tc_overwrites(h->tc[h->g_sout_wg])) { skip[ns][0] = h->sOutW.off; skip[ns][1] = h->sOutB.off; ++ns; }
  int64_t cur = 0;
  for (int i = 0; i <= ns; ++i) {
    const int64_t end = (i < ns) ? skip[i][0] : h->arena_elems;
    if (end > cur) MFVAE_CUDA(cudaMemsetAsync(G + cur, 0, static_cast<size_t>(end - cur) * sizeof(float), q));
    if (i < ns) cur = skip[i][1];
  }
  if (h->fold_act && !(h->use_tc && gemm_tc_overwrites(h->tc[h->g_dec_wgT])))
    MFVAE_CUDA(cudaMemsetAsync(h->ws + h->DTACT.off, 0, static_cast<size_t>(2) * h->decH[0] * h->DTACT.ld * sizeof(float), q));
  return 0;
}

// Backward.  The dgrad chain (which carries the dependency from layer to layer) runs on the caller's stream; every
// wgrad and bias column-sum only needs D_l and X_l, so they are forked onto a side stream as soon as D_l exists.  The
// small layers of this network launch far fewer CTAs than the GPU has SMs; running the two chains concurrently fills
// the machine.  Gradient-bucket events (data-parallel all-reduce) are recorded on the stream that finishes them.
static int do_backward(MfvaeHandle_* h, const MfvaeBatch* b, cudaStream_t s, const float* glat = nullptr, bool with_kl = true) {
  MFVAE_TRY(check_ready(h, b));
  const int dt = h->dtype, A = h->A, nh = h->cfg.n_dec_hidden;
  float* G = h->ar.d_grad;
  char* ws = h->ws;
  const bool overlap = h->side != nullptr && !h->profiling;     // profiling wants serialised, undisturbed durations
  cudaStream_t w = overlap ? h->side : s;                       // stream of the wgrad chain
  cudaStream_t cs = (overlap && h->csum) ? h->csum : w;         // stream of the bias column sums
  size_t ev_i = 0;
  auto fork = [&]() -> int {
    if (!overlap) return 0;
    MFVAE_CHECK(ev_i < h->fork_ev.size(), "fork event pool exhausted");
    MFVAE_CUDA(cudaEventRecord(h->fork_ev[ev_i], s));
    MFVAE_CUDA(cudaStreamWaitEvent(w, h->fork_ev[ev_i], 0));
    if (cs != w) MFVAE_CUDA(cudaStreamWaitEvent(cs, h->fork_ev[ev_i], 0));
    ++ev_i;
    return 0;
  };
  // the column sums of a gradient bucket finish on `cs`: fold them into `w` before the bucket's event is recorded there
  size_t cs_i = 0;
  auto csum_into_w = [&]() -> int {
    if (cs == w) return 0;
    MFVAE_CHECK(cs_i < h->csum_ev.size(), "column-sum event pool exhausted");
    MFVAE_CUDA(cudaEventRecord(h->csum_ev[cs_i], cs));
    MFVAE_CUDA(cudaStreamWaitEvent(w, h->csum_ev[cs_i], 0));
    ++cs_i;
    return 0;
  };
  if (!h->grads_zeroed) MFVAE_TRY(zero_grads(h, s));            // (mfvae_fwd_bwd zeroes them beside the forward pass)
  else MFVAE_CUDA(cudaStreamWaitEvent(s, h->zero_ev, 0));
  h->grads_zeroed = false;
  MFVAE_TRY(fork());                                            // D(recon_s), D(recon_r) and the zeroed arena are ready
  // reward head: reward_linear and the reward decoder's output layer are tiny; their whole backward chain runs on the
  // aux stream beside the state output layer (which owns the SMs for ~70 us on each of the other two streams)
  const bool auxo = overlap && h->aux != nullptr;
  cudaStream_t r = auxo ? h->aux : s, rw = auxo ? h->aux : w;
  if (auxo) {
    MFVAE_CUDA(cudaEventRecord(h->aux_fork_ev, s));
    MFVAE_CUDA(cudaStreamWaitEvent(h->aux, h->aux_fork_ev, 0));
  }
  // reward-head dgrads first (aux): the two output-layer GEMMs below are persistent kernels that claim every SM, and the
  // decoder dgrad chain on the caller's stream needs the reward half of D before it can go on
  MFVAE_TRY(run_gemm(h, h->g_rl_dg, r));
  MFVAE_TRY(run_gemm(h, h->g_rout_dg, r));
  // output layers
  MFVAE_TRY(run_gemm(h, h->g_sout_dg, s));
  MFVAE_TRY(run_gemm(h, h->g_sout_wg, w));
  if (!h->sout_bias_done) MFVAE_TRY(launch_colsum(ws + h->DRS.off, dt, 1, h->B, h->S, h->DRS.ld, 0, G + h->sOutB.off, 0, cs));
  h->sout_bias_done = false;
  if (auxo) {                                                   // D of the last hidden layer: state half (s) + reward half (aux)
    MFVAE_CUDA(cudaEventRecord(h->aux_join_ev, h->aux));
    MFVAE_CUDA(cudaStreamWaitEvent(s, h->aux_join_ev, 0));
  }
  MFVAE_CUDA(cudaEventRecord(h->read_ev[0], s));                // the three output-layer dgrads (readers of bucket 0's weights) are queued
  MFVAE_TRY(run_gemm(h, h->g_rl_wg, rw));
  MFVAE_TRY(launch_colsum(ws + h->DRR.off, dt, 1, h->B, A, h->DRR.ld, 0, G + h->rlb.off, 0, rw));
  if (!auxo) MFVAE_TRY(fork());                                 // D(reward decoder output) from the dgrad chain
  MFVAE_TRY(run_gemm(h, h->g_rout_wg, rw));
  MFVAE_TRY(launch_colsum(ws + h->DRR0.off, dt, 1, h->B, A, h->DRR0.ld, 0, G + h->rOutB.off, 0, rw));
  if (auxo) {
    MFVAE_CUDA(cudaEventRecord(h->aux_join2_ev, h->aux));
    MFVAE_CUDA(cudaStreamWaitEvent(w, h->aux_join2_ev, 0));
  }
  MFVAE_TRY(csum_into_w());
  MFVAE_CUDA(cudaEventRecord(h->buckets[0].ev, w));
  // decoder hidden layers, last to first
  // (a high-priority stream of its own for layer 0's wgrad -- the largest bucket -- was measured: no change at 1 or 2 GPUs;
  //  the persistent GEMM CTAs already resident decide the order, not the stream priority)
  for (int l = nh - 1; l >= 0; --l) {
    MFVAE_TRY(fork());                                          // D_l (both decoder halves) ready
    MFVAE_TRY(run_gemm(h, h->g_dec_wg[l], w));
    if (l == 0 && h->fold_act) {
      // d T = d H0^T . onehot, then exactly: d W0[:, action columns] = d T . tables, d tables = d T^T . W0[:, action columns]
      MFVAE_TRY(run_gemm(h, h->g_dec_wgT, w));
      MFVAE_TRY(launch_act_fold_bwd(reinterpret_cast<const float*>(ws + h->DTACT.off), h->DTACT.ld, h->ar.d_param + h->decW[0].off,
                                    G + h->decW[0].off, h->Din, h->Kz, h->ar.d_param + h->actT.off, G + h->actT.off,
                                    static_cast<int64_t>(h->nact_max) * h->C, 2 * h->decH[0], A, h->C, h->nact_max, w));
    }
    MFVAE_TRY(launch_colsum(ws + h->DHD[l].off, dt, 1, h->B, 2 * h->decH[l], h->DHD[l].ld, 0, G + h->decB[l].off, 0, cs));
    if (l == 1) { MFVAE_TRY(csum_into_w()); MFVAE_CUDA(cudaEventRecord(h->buckets[1].ev, w)); }
    MFVAE_TRY(run_gemm(h, h->g_dec_dg[l], s));
    if (l == 1) MFVAE_CUDA(cudaEventRecord(h->read_ev[1], s));    // decoder layers >= 1 have been read
  }
  if (nh == 1) { MFVAE_CUDA(cudaEventRecord(h->buckets[1].ev, w)); MFVAE_CUDA(cudaEventRecord(h->read_ev[1], s)); }
  MFVAE_TRY(csum_into_w());
  MFVAE_CUDA(cudaEventRecord(h->buckets[2].ev, w));
  if (h->dec_read_ev) MFVAE_CUDA(cudaEventRecord(h->dec_read_ev, s));   // last reader of the decoder weights (dgrad layer 0) is queued
  MFVAE_CUDA(cudaEventRecord(h->read_ev[2], s));
  // action tables (model.py:121: unregistered; gradients still flow)
  cudaStream_t at = s;
  if (auxo) {
    MFVAE_CUDA(cudaEventRecord(h->aux_fork2_ev, s));            // GZIN is complete
    MFVAE_CUDA(cudaStreamWaitEvent(h->aux, h->aux_fork2_ev, 0));
    at = h->aux;
  }
  if (h->fold_act) {
    // the action tables' gradient came out of act_fold_bwd above
  } else if (!h->cont_act) {
    MFVAE_TRY(launch_act_table_grad(ws + h->GZIN.off, dt, h->GZIN.ld, A * h->L, b->d_act, A, h->d_meta + 2 * A, A, h->C, h->B,
                                    G + h->actT.off, static_cast<int64_t>(h->nact_max) * h->C, at));
  } else {   // ActionEncoder backward: D = the action columns of d ZIN
    const int64_t es = dtype_size(dt);
    const char* gzact = ws + h->GZIN.off + static_cast<int64_t>(A) * h->L * es;
    MFVAE_TRY(run_gemm(h, h->g_act_wg2, at));
    MFVAE_TRY(launch_colsum(gzact, dt, A, h->B, h->C, h->GZIN.ld, h->C, G + h->actB2.off, h->C, at));
    MFVAE_TRY(run_gemm(h, h->g_act_dg2, at));
    MFVAE_TRY(run_gemm(h, h->g_act_wg1, at));
    MFVAE_TRY(launch_colsum(ws + h->DHA.off, dt, A, h->B, h->Hact, h->DHA.ld, h->DHA.gs, G + h->actB1.off, h->Hact, at));
  }
  if (auxo) MFVAE_CUDA(cudaEventRecord(h->aux_join_ev, h->aux));
  // reparameterization + KL backward
  ReparamBwdArgs rb{};
  const float* lat = reinterpret_cast<const float*>(ws + h->LAT.off);
  rb.gz = ws + h->GZIN.off; rb.gz_ld = h->GZIN.ld; rb.g_dtype = dt;
  rb.mu = lat; rb.lv = lat + h->L; rb.lat_as = h->LAT.gs; rb.lat_bs = h->LAT.ld;
  rb.eps = b->d_eps; rb.eps_ld = static_cast<int64_t>(A) * h->L;
  rb.dlat = ws + h->DLAT.off; rb.dlat_as = h->DLAT.gs; rb.dlat_bs = h->DLAT.ld; rb.d_dtype = dt;
  rb.B = h->B; rb.A = A; rb.L = h->L; rb.seed = b->seed; rb.step = b->step; rb.sample0 = b->sample0;
  rb.kl_scale = with_kl ? h->cfg.kl_weight / static_cast<float>(b->batch_global) : 0.f;
  rb.glat = glat;
  MFVAE_TRY(launch_reparam_kl_bwd(rb, s));
  // encoders, last layer to first
  const bool fold_i = fold_idx(h, b);
  for (int l = h->ne - 1; l >= 0; --l) {
    const MfvaeHandle_::Buf& D = (l + 1 == h->ne) ? h->DLAT : h->DXE[l];
    MFVAE_TRY(fork());
    MFVAE_TRY(run_gemm(h, (l == 0 && fold_i) ? h->g_enc_wg0_f : h->g_enc_wg[l], w));
    MFVAE_TRY(launch_colsum(ws + D.off, dt, A, h->B, h->encN[l], D.ld, D.gs, G + h->encB[l].off, h->encN[l], cs));
    if (l == 0 && fold_i) {
      // d emb[a] = W0_a[:, :I]^T . d b0_a and d W0_a[:, :I] = d b0_a (x) emb[a], behind the column sum that produces d b0
      MFVAE_TRY(launch_emb_grad_fold(h->ar.d_param + h->encW[0].off, G + h->encW[0].off, h->encK[0], G + h->encB[0].off,
                                     h->ar.d_param + h->idx_emb.off, G + h->idx_emb.off, A, h->encN[0], h->I, cs));
    } else {
      MFVAE_TRY(run_gemm(h, h->g_enc_dg[l], s));
    }
  }
  // id embedding (model.py:113,142)
  if (fold_i) {
    // done by emb_grad_fold above
  } else if (b->d_idx)
    MFVAE_TRY(launch_idx_emb_scatter(ws + h->GX0.off, dt, h->GX0.gs, h->GX0.ld, b->d_idx, A, A, h->I, h->B, G + h->idx_emb.off, s));
  else
    MFVAE_TRY(launch_colsum(ws + h->GX0.off, dt, A, h->B, h->I, h->GX0.ld, h->GX0.gs, G + h->idx_emb.off, h->I, s));
  if (overlap) {                                               // join
    MFVAE_TRY(csum_into_w());
    MFVAE_CUDA(cudaEventRecord(h->join_ev, w));
    MFVAE_CUDA(cudaStreamWaitEvent(s, h->join_ev, 0));
    if (auxo) MFVAE_CUDA(cudaStreamWaitEvent(s, h->aux_join_ev, 0));
  }
  for (size_t i = 3; i < h->buckets.size(); ++i) MFVAE_CUDA(cudaEventRecord(h->buckets[i].ev, s));
  return 0;
}

}  // namespace mfvae

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* mfvae_last_error(void) { return g_last_error.c_str(); }
int mfvae_version(void) { return 100; }

int mfvae_create(const MfvaeConfig* cfg, int device, MfvaeHandle* out) {
  MFVAE_CHECK(cfg && out, "null argument");
  MFVAE_CHECK(cfg->obs_dim && cfg->n_act, "config needs obs_dim and n_act");
  if (device >= 0) {
    int ndev = 0;
    MFVAE_CUDA(cudaGetDeviceCount(&ndev));
    MFVAE_CHECK(device < ndev, "no such CUDA device (this library has no CPU path)");
    MFVAE_CUDA(cudaSetDevice(device));
  }
  MfvaeHandle_* h = new MfvaeHandle_();
  h->cfg = *cfg; h->device = device;
  // AUTO = what measured fastest on B200: the fused encoder chain, the loss as its own kernel
  if (h->cfg.fusion == MFVAE_FUSE_AUTO) h->cfg.fusion = MFVAE_FUSE_ENCODER;
  h->obs_dim.assign(cfg->obs_dim, cfg->obs_dim + cfg->n_agents);
  h->n_act.assign(cfg->n_act, cfg->n_act + cfg->n_agents);
  h->cfg.obs_dim = h->obs_dim.data(); h->cfg.n_act = h->n_act.data();
  if (build_layout(h) != 0) { delete h; return 1; }
  // buckets in backward-completion order
  auto add_bucket = [&](int64_t b, int64_t e) {
    MfvaeHandle_::Bucket k{b, e, nullptr};
    if (device >= 0) cudaEventCreateWithFlags(&k.ev, cudaEventDisableTiming);
    h->buckets.push_back(k);
  };
  const int64_t l1 = (h->cfg.n_dec_hidden > 1) ? h->decW[1].off : h->reg3_begin;
  add_bucket(h->reg3_begin, h->enc_begin);           // output layers + reward_linear
  add_bucket(l1, h->reg3_begin);                      // decoder hidden layers >= 1 (may be empty)
  add_bucket(h->reg2_begin, l1);                      // decoder layer 0 (largest)
  add_bucket(0, h->reg2_begin);                       // idx_emb
  if (h->cfg.optimize_encoders) add_bucket(h->enc_begin, h->arena_elems);
  if (device < 0) { *out = h; return 0; }        // layout-only handle
  {
    // the wgrad chain gets the greater stream priority: every kernel of either chain fills the SMs' CTA slots, so the chains
    // interleave only at kernel boundaries; without the priority the wgrad chain lags ~200 us behind its inputs and leaves a
    // tail in which it runs alone (kernel timeline, tools/dp_trace.py)
    int lo = 0, hi = 0;
    const char* env = getenv("MFVAE_SIDE_PRIORITY");
    cudaDeviceGetStreamPriorityRange(&lo, &hi);            // hi = numerically smallest = greatest priority
    const int prio = (env && env[0] == '0') ? lo : hi;
    if (cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio) != cudaSuccess) h->side = nullptr;
  }
  for (int i = 0; i < 4 + 2 * MFVAE_MAX_HIDDEN + 4; ++i) {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    h->fork_ev.push_back(e);
  }
  cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming);
  if (cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking) != cudaSuccess) h->aux = nullptr;
  if (cudaStreamCreateWithFlags(&h->csum, cudaStreamNonBlocking) != cudaSuccess) h->csum = nullptr;
  for (int i = 0; i < 6; ++i) { cudaEvent_t e = nullptr; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); h->csum_ev.push_back(e); }
  cudaEventCreateWithFlags(&h->zero_ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->loss_ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->zero_fork_ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->eb_ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->aux_fork_ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->aux_join_ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->aux_fork2_ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->aux_join2_ev, cudaEventDisableTiming);
  if (cudaStreamCreateWithFlags(&h->opt_stream, cudaStreamNonBlocking) != cudaSuccess) h->opt_stream = nullptr;
  cudaEventCreateWithFlags(&h->opt_ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->dec_read_ev, cudaEventDisableTiming);
  for (int i = 0; i < 3; ++i) cudaEventCreateWithFlags(&h->read_ev[i], cudaEventDisableTiming);
  std::vector<int32_t> meta;
  meta.insert(meta.end(), h->obs_off.begin(), h->obs_off.end());
  meta.insert(meta.end(), h->obs_dim.begin(), h->obs_dim.end());
  meta.insert(meta.end(), h->n_act.begin(), h->n_act.end());
  meta.insert(meta.end(), h->act_off.begin(), h->act_off.end());
  if (cudaMalloc(&h->d_meta, meta.size() * sizeof(int32_t)) != cudaSuccess ||
      cudaMemcpy(h->d_meta, meta.data(), meta.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete h; MFVAE_FAIL("cudaMalloc / cudaMemcpy of the agent table failed");
  }
  *out = h;
  return 0;
}

int mfvae_destroy(MfvaeHandle h) {
  if (!h) return 0;
  free_plans(h);
  for (auto& b : h->buckets) if (b.ev) cudaEventDestroy(b.ev);
  for (auto e : h->prof_ev) cudaEventDestroy(e);
  for (auto e : h->enc_prof_ev) if (e) cudaEventDestroy(e);
  for (auto e : h->fork_ev) if (e) cudaEventDestroy(e);
  if (h->join_ev) cudaEventDestroy(h->join_ev);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->aux) cudaStreamDestroy(h->aux);
  if (h->csum) cudaStreamDestroy(h->csum);
  if (h->zero_ev) cudaEventDestroy(h->zero_ev);
  if (h->loss_ev) cudaEventDestroy(h->loss_ev);
  if (h->zero_fork_ev) cudaEventDestroy(h->zero_fork_ev);
  for (auto e : h->csum_ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : {h->aux_fork_ev, h->aux_join_ev, h->aux_fork2_ev, h->aux_join2_ev}) if (e) cudaEventDestroy(e);
  if (h->eb_ev) cudaEventDestroy(h->eb_ev);
  for (auto e : h->ar_ev) if (e) cudaEventDestroy(e);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->comm_done_ev) cudaEventDestroy(h->comm_done_ev);
  if (h->opt_ev) cudaEventDestroy(h->opt_ev);
  if (h->dec_read_ev) cudaEventDestroy(h->dec_read_ev);
  for (int i = 0; i < 3; ++i) if (h->read_ev[i]) cudaEventDestroy(h->read_ev[i]);
  if (h->opt_stream) cudaStreamDestroy(h->opt_stream);
  if (h->d_meta) cudaFree(h->d_meta);
  delete h;
  return 0;
}

int64_t mfvae_arena_elems(MfvaeHandle h) { return h ? h->arena_elems : -1; }
int64_t mfvae_optimized_elems(MfvaeHandle h) { return h ? h->optimized_elems : -1; }
int32_t mfvae_tensor_count(MfvaeHandle h) { return h ? static_cast<int32_t>(h->table.size()) : -1; }

int mfvae_tensor_table(MfvaeHandle h, MfvaeTensorInfo* out, int32_t capacity) {
  MFVAE_CHECK(h && out, "null argument");
  MFVAE_CHECK(capacity >= static_cast<int32_t>(h->table.size()), "tensor table capacity too small");
  std::copy(h->table.begin(), h->table.end(), out);
  return 0;
}

int mfvae_bind_arenas(MfvaeHandle h, const MfvaeArenas* a) {
  MFVAE_CHECK(h && a, "null argument");
  MFVAE_CHECK(h->device >= 0, "layout-only handle: there is no CPU path");
  MFVAE_CHECK(a->d_param && a->d_grad && a->d_m && a->d_v, "param / grad / m / v arenas are required");
  MFVAE_CHECK(h->dtype != kBF16 || a->d_shadow_bf16, "bf16 precision needs the bf16 shadow arena");
  MFVAE_CHECK(reinterpret_cast<uintptr_t>(a->d_param) % 256 == 0 && reinterpret_cast<uintptr_t>(a->d_grad) % 256 == 0 &&
              reinterpret_cast<uintptr_t>(a->d_shadow_bf16) % 256 == 0, "arenas must be 256-byte aligned");
  h->ar = *a;
  if (h->ws) return build_ops(h);
  return 0;
}

int mfvae_refresh_shadow(MfvaeHandle h, void* stream) {
  MFVAE_CHECK(h, "null handle");
  if (!h->ar.d_shadow_bf16) return 0;
  return launch_cast_bf16(h->ar.d_param, static_cast<__nv_bfloat16*>(h->ar.d_shadow_bf16), h->arena_elems,
                          static_cast<cudaStream_t>(stream));
}

int mfvae_refresh_shadow_range(MfvaeHandle h, int64_t begin, int64_t end, void* stream) {
  MFVAE_CHECK(h, "null handle");
  if (!h->ar.d_shadow_bf16) return 0;
  MFVAE_CHECK(begin >= 0 && end <= h->arena_elems && begin <= end && begin % 4 == 0 && (end - begin) % 4 == 0, "shadow range must be 4-element aligned and inside the arena");
  if (end == begin) return 0;
  return launch_cast_bf16(h->ar.d_param + begin, static_cast<__nv_bfloat16*>(h->ar.d_shadow_bf16) + begin, end - begin,
                          static_cast<cudaStream_t>(stream));
}

int64_t mfvae_workspace_bytes(MfvaeHandle h, int32_t batch) {
  if (!h || batch < 1) return -1;
  MfvaeHandle_ tmp = *h;            // layout only; does not touch plans
  tmp.tc.clear(); tmp.gemms.clear();
  return layout_workspace(&tmp, batch);
}

int mfvae_bind_workspace(MfvaeHandle h, void* d_ws, int64_t bytes, int32_t batch) {
  MFVAE_CHECK(h && d_ws, "null argument");
  MFVAE_CHECK(h->device >= 0, "layout-only handle: there is no CPU path");
  MFVAE_CHECK(batch >= 1, "batch must be positive");
  MFVAE_CHECK(reinterpret_cast<uintptr_t>(d_ws) % 256 == 0, "workspace must be 256-byte aligned");
  const int64_t need = layout_workspace(h, batch);
  MFVAE_CHECK(bytes >= need, "workspace too small");
  h->ws = static_cast<char*>(d_ws); h->ws_bytes = bytes; h->B = batch;
  // scalar slots and tickets start at zero
  MFVAE_CUDA(cudaMemset(h->ws + h->off_losses, 0, 64 * sizeof(float)));
  MFVAE_CUDA(cudaMemset(h->ws + h->off_scratch, 0, 3 * 4096 * sizeof(float)));
  if (h->ar.d_param) return build_ops(h);
  return 0;
}

int mfvae_forward(MfvaeHandle h, const MfvaeBatch* b, MfvaeOutputs* out, void* stream) {
  MFVAE_CHECK(h, "null handle");
  return do_forward(h, b, out, static_cast<cudaStream_t>(stream));
}
int mfvae_loss(MfvaeHandle h, const MfvaeBatch* b, int32_t loss_kind, void* stream) {
  MFVAE_CHECK(h, "null handle");
  return do_loss(h, b, loss_kind, static_cast<cudaStream_t>(stream));
}
int mfvae_set_loss_weights(MfvaeHandle h, float kl_weight, float r_weight) {
  MFVAE_CHECK(h, "null handle");
  h->cfg.kl_weight = kl_weight; h->cfg.r_weight = r_weight; h->s_weight = 1.0f;
  return 0;
}
int mfvae_set_loss_weights3(MfvaeHandle h, float kl_weight, float r_weight, float s_weight) {
  MFVAE_CHECK(h, "null handle");
  h->cfg.kl_weight = kl_weight; h->cfg.r_weight = r_weight; h->s_weight = s_weight;
  return 0;
}
int mfvae_backward(MfvaeHandle h, const MfvaeBatch* b, void* stream) {
  MFVAE_CHECK(h, "null handle");
  return do_backward(h, b, static_cast<cudaStream_t>(stream));
}
int mfvae_backward_ext(MfvaeHandle h, const MfvaeBatch* b, const float* d_g_recon_s, int64_t ld_s,
                       const float* d_g_recon_r, int64_t ld_r, const float* d_g_latent, void* stream) {
  MFVAE_CHECK(h, "null handle");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MFVAE_TRY(check_ready(h, b));
  MFVAE_TRY(launch_cast2d(d_g_recon_s, ld_s, h->ws + h->DRS.off, h->DRS.ld, h->dtype, h->B, h->S, s));
  MFVAE_TRY(launch_cast2d(d_g_recon_r, ld_r, h->ws + h->DRR.off, h->DRR.ld, h->dtype, h->B, h->A, s));
  return do_backward(h, b, s, d_g_latent, false);
}
int mfvae_fwd_bwd(MfvaeHandle h, const MfvaeBatch* b, MfvaeOutputs* out, void* stream) {
  MFVAE_CHECK(h, "null handle");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // tensor-core engine: reconstruction loss of the state head fused into its output-layer GEMM (recon_s is not written)
  const bool fuse = h->use_tc && (h->cfg.fusion & MFVAE_FUSE_LOSS) && h->g_sout_loss >= 0 && b && b->d_next && b->d_rew && !b->next_bf16;
  // bf16 engine: the train step needs recon_s only inside the loss, so the output layer writes it once in bf16 straight into
  // the D(recon_s) buffer and the loss kernel turns it into the gradient in place (no fp32 round trip: -140 MB per step)
  const bool r16 = !fuse && h->use_tc && h->g_sout_fwd16 >= 0 && b && b->d_next && b->d_rew;
  if (h->csum && !h->profiling) {
    // zero_grad beside the forward pass: ordered after everything already queued on the caller's stream (the previous
    // optimizer step read these gradients), awaited by backward
    MFVAE_CUDA(cudaEventRecord(h->zero_fork_ev, s));
    MFVAE_CUDA(cudaStreamWaitEvent(h->csum, h->zero_fork_ev, 0));
    MFVAE_TRY(zero_grads(h, h->csum));
    MFVAE_CUDA(cudaEventRecord(h->zero_ev, h->csum));
    h->grads_zeroed = true;
  }
  MFVAE_TRY(do_forward(h, b, out, s, fuse, r16, true));
  MFVAE_TRY(do_loss(h, b, MFVAE_LOSS_DEFAULT, s, fuse, r16, true, true));
  return do_backward(h, b, s);
}

int mfvae_adam_step(MfvaeHandle h, float lr, float beta1, float beta2, float eps, int64_t t, void* stream) {
  MFVAE_CHECK(h && h->ar.d_param, "arenas are not bound");
  return launch_adam(h->ar.d_param, h->ar.d_grad, h->ar.d_m, h->ar.d_v,
                     static_cast<__nv_bfloat16*>(h->ar.d_shadow_bf16), h->optimized_elems, lr, beta1, beta2, eps, t,
                     static_cast<cudaStream_t>(stream));
}

// Adam for the decoder block of the arena on a third stream as soon as its gradient buckets are final, so that the
// 28 B/parameter sweep overlaps the encoder half of backward still running on the caller's stream; the remaining
// ranges (idx_emb, and the encoders when they are optimised) follow on the caller's stream, which then joins.
int mfvae_adam_step_overlapped(MfvaeHandle h, float lr, float beta1, float beta2, float eps, int64_t t, void* stream) {
  MFVAE_CHECK(h && h->ar.d_param, "arenas are not bound");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!h->opt_stream || h->profiling) return mfvae_adam_step(h, lr, beta1, beta2, eps, t, stream);
  __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(h->ar.d_shadow_bf16);
  auto range = [&](int64_t b, int64_t e, cudaStream_t st) -> int {
    if (e <= b) return 0;
    // 4 blocks per SM when it runs beside the wgrad tail: 8 x 256 threads would take every thread slot of the SM (2 was measured too: slower)
    return launch_adam(h->ar.d_param + b, h->ar.d_grad + b, h->ar.d_m + b, h->ar.d_v + b, sh ? sh + b : nullptr, e - b,
                       lr, beta1, beta2, eps, t, st, st == s ? 8 : 4);
  };
  // one launch per gradient bucket, each as soon as that bucket's gradients are final AND backward no longer reads its weights:
  // the output layers' 174 MB sweep runs beside the decoder's latency-bound middle, the decoder-layer-0 sweep beside the
  // encoder half of backward, instead of one 523 MB sweep fighting the encoder wgrads for HBM at the end of the step
  for (int i = 0; i < 3; ++i) {
    MFVAE_CUDA(cudaStreamWaitEvent(h->opt_stream, h->buckets[i].ev, 0));
    MFVAE_CUDA(cudaStreamWaitEvent(h->opt_stream, h->read_ev[i], 0));
    MFVAE_TRY(range(h->buckets[i].begin, h->buckets[i].end, h->opt_stream));
  }
  MFVAE_CUDA(cudaEventRecord(h->opt_ev, h->opt_stream));
  MFVAE_TRY(range(0, h->reg2_begin, s));
  if (h->optimized_elems > h->enc_begin) MFVAE_TRY(range(h->enc_begin, h->optimized_elems, s));
  MFVAE_CUDA(cudaStreamWaitEvent(s, h->opt_ev, 0));
  return 0;
}

// Adam over one arena range (a gradient bucket) on `stream`: data parallel runs it per bucket on the communication stream
// right behind that bucket's all-reduce, so the optimizer sweep overlaps the rest of backward.
int mfvae_adam_range(MfvaeHandle h, int64_t begin, int64_t end, float lr, float beta1, float beta2, float eps, int64_t t, void* stream) {
  MFVAE_CHECK(h && h->ar.d_param, "arenas are not bound");
  MFVAE_CHECK(begin >= 0 && end <= h->optimized_elems && begin % 8 == 0, "adam range must lie inside the optimised prefix, 8-element aligned");
  if (end <= begin) return 0;
  __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(h->ar.d_shadow_bf16);
  return launch_adam(h->ar.d_param + begin, h->ar.d_grad + begin, h->ar.d_m + begin, h->ar.d_v + begin, sh ? sh + begin : nullptr,
                     end - begin, lr, beta1, beta2, eps, t, static_cast<cudaStream_t>(stream));
}
// make `stream` wait until the backward pass in flight has finished READING the weights gradient bucket i covers (buckets
// 0..2: output layers, decoder layers >= 1, decoder layer 0; later buckets have no reader left once their event has fired)
int mfvae_bucket_read_wait(MfvaeHandle h, int32_t i, void* stream) {
  MFVAE_CHECK(h && i >= 0 && i < static_cast<int32_t>(h->buckets.size()), "bucket index out of range");
  if (i < 3 && h->read_ev[i]) MFVAE_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->read_ev[i], 0));
  return 0;
}
// make `stream` wait until the backward pass in flight has finished READING the decoder / output-layer weights
int mfvae_wait_decoder_reads(MfvaeHandle h, void* stream) {
  MFVAE_CHECK(h && h->dec_read_ev, "no backward pass has been recorded");
  MFVAE_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->dec_read_ev, 0));
  return 0;
}

// Persistent GEMM grids leave `n_sms` SMs unclaimed (process-wide; takes effect for plans built afterwards: the handle's
// plans are rebuilt here).  Data parallel sets it to the number of SMs the NCCL kernels occupy.
int mfvae_set_sm_reserve(MfvaeHandle h, int32_t n_sms) {
  MFVAE_CHECK(n_sms >= 0 && n_sms < kNumSMs / 2, "sm reserve out of range");
  g_tc_sm_reserve = n_sms;
  if (h && h->ws && h->ar.d_param) return build_ops(h);
  return 0;
}

// ---- data-parallel exchange over peer memory (comm.cu) ----
constexpr int64_t kCommSmallBytes = 64 * 1024;      // ranges up to 16 K elements take the single-kernel path
int mfvae_comm_bind(MfvaeHandle h, int32_t rank, int32_t world, void* const* d_peer_windows, void* multicast_window,
                    void* const* d_signal_pads, int64_t signal_pad_bytes, void* local_window, int64_t window_bytes, int32_t payload_bf16,
                    int32_t max_blocks) {
  MFVAE_CHECK(h && h->device >= 0, "comm: needs a device handle");
  MFVAE_CHECK(world >= 2 && world <= 16 && rank >= 0 && rank < world, "comm: bad rank / world");
  MFVAE_CHECK(d_peer_windows && d_signal_pads && local_window, "comm: null window / pad pointers");
  MFVAE_CHECK(reinterpret_cast<uintptr_t>(local_window) % 256 == 0, "comm: the window must be 256-byte aligned");
  const int64_t es = payload_bf16 ? 2 : 4;
  const int64_t elems = round_up(h->optimized_elems, 8);
  const int64_t scalar_off = round_up(elems * es, 256);
  MFVAE_CHECK(window_bytes >= scalar_off + 256 + kCommSmallBytes, "comm: window too small (mfvae_comm_window_bytes)");
  CommCtx& c = h->comm;
  c.rank = rank; c.world = world; c.d_peers = d_peer_windows; c.mc = multicast_window;
  c.d_pads = reinterpret_cast<uint32_t* const*>(d_signal_pads); c.local = local_window;
  c.dtype = payload_bf16 ? kBF16 : kF32; c.elems = elems; c.scalar_off = scalar_off;
  c.small_off = scalar_off + 256; c.small_bytes = kCommSmallBytes;
  // one 32-bit flag per (block, peer) behind the 64 slots left to the host framework
  const int64_t slot_blocks = (signal_pad_bytes / 4 - 64) / world;
  MFVAE_CHECK(slot_blocks >= 1, "comm: signal pads too small (need >= 256 + 4 * world bytes)");
  // CTAs of a reduce: measured (cfg2, B = 4096 per GPU).  At 2 GPUs a rank moves half of every bucket and needs the bytes in
  // flight: 128 CTAs (0.985 ms vs 1.04 with 32).  From 4 GPUs the slices are small and the reduce mostly waits in its barriers,
  // where resident CTAs only take slots from backward's kernels: 32 CTAs (4 GPUs: 0.956 vs 0.997 with 112, 1.012 with 16;
  // 8 GPUs: 0.973 vs 0.999 with 56, 1.000 with 16).
  // (with the split synchronisation 64 CTAs do at 2 ranks what took 128 with the barriers inside: 0.955 vs 0.969 ms)
  const int64_t dflt = (world <= 2) ? 64 : 32;
  c.max_blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(max_blocks > 0 ? max_blocks : dflt, slot_blocks), kNumSMs)));
  // default: pack / rendezvous / reduce / rendezvous as four launches -- no wide kernel ever waits for a peer (4 GPUs: 0.953 ->
  // 0.919 ms against the single kernel with barriers inside); MFVAE_DP_SPLIT_SYNC=0 selects the single kernel
  { const char* e = getenv("MFVAE_DP_SPLIT_SYNC"); c.split_sync = (e && e[0] == '0') ? 0 : 1; }
  for (auto& e : h->ar_ev) if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  return 0;
}
int64_t mfvae_comm_window_bytes(MfvaeHandle h, int32_t payload_bf16) {
  if (!h) return -1;
  return round_up(round_up(h->optimized_elems, 8) * (payload_bf16 ? 2 : 4), 256) + 256 + kCommSmallBytes;
}
// all-reduce (sum) of the gradient arena's elements [begin, end) across ranks, on `stream`: pack -> two-shot reduce in the
// windows.  With do_adam the fused Adam of that range follows (reading the reduced gradient from the window and leaving its
// fp32 copy in the gradient arena); without, the reduced gradient is unpacked into the gradient arena.
int mfvae_allreduce_grads(MfvaeHandle h, int64_t begin, int64_t end, int32_t do_adam, float lr, float beta1, float beta2, float eps,
                          int64_t t, void* stream) {
  MFVAE_CHECK(h && h->comm.world >= 2, "comm: mfvae_comm_bind has not been called");
  MFVAE_CHECK(begin >= 0 && end <= h->comm.elems && begin % 8 == 0, "comm: range outside the optimised prefix or not 8-element aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t e8 = round_up(end, 8);
  if (e8 <= begin) return 0;
  const CommCtx& c = h->comm;
  // measurement switches (never set in production): MFVAE_DP_DEBUG=noar packs without reducing, =nosmall skips the tiny buckets'
  // cross-rank kernel -- they locate the data-parallel overhead; results are then per-rank, not reduced
  static const char* dbg = getenv("MFVAE_DP_DEBUG");
  const bool noar = dbg && strstr(dbg, "noar"), nosmall = dbg && strstr(dbg, "nosmall");
  if ((e8 - begin) * 4 <= c.small_bytes && nosmall) {
    __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(h->ar.d_shadow_bf16);
    return do_adam ? launch_adam(h->ar.d_param + begin, h->ar.d_grad + begin, h->ar.d_m + begin, h->ar.d_v + begin, sh ? sh + begin : nullptr,
                                 e8 - begin, lr, beta1, beta2, eps, t, s) : 0;
  }
  if ((e8 - begin) * 4 <= c.small_bytes) {
    __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(h->ar.d_shadow_bf16);
    return comm_small_allreduce_adam(c, e8 - begin, h->ar.d_param + begin, h->ar.d_grad + begin, h->ar.d_m + begin, h->ar.d_v + begin,
                                     sh ? sh + begin : nullptr, do_adam, lr, beta1, beta2, eps, t, s);
  }
  if (noar) MFVAE_TRY(comm_pack(h->ar.d_grad, c.local, c.dtype, begin, e8, s));
  else MFVAE_TRY(comm_allreduce(c, begin, e8, s, h->ar.d_grad));       // pack (fp32 -> payload) + two-shot reduce, one kernel
  if (do_adam) {
    // the optimizer sweep of this range runs on the handle's optimizer stream behind the reduce, so that the next bucket's
    // reduce (on `stream`) does not queue behind a 100-200 MB sweep; mfvae_opt_join makes the caller's stream wait for it
    const int64_t es = (c.dtype == kBF16) ? 2 : 4;
    __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(h->ar.d_shadow_bf16);
    cudaStream_t o = h->opt_stream ? h->opt_stream : s;
    if (o != s) {
      cudaEvent_t ev = h->ar_ev[h->ar_ev_i++ % h->ar_ev.size()];
      MFVAE_CUDA(cudaEventRecord(ev, s));
      MFVAE_CUDA(cudaStreamWaitEvent(o, ev, 0));
    }
    MFVAE_TRY(launch_adam_payload(h->ar.d_param + begin, static_cast<const char*>(c.local) + begin * es, c.dtype, h->ar.d_grad + begin,
                                  h->ar.d_m + begin, h->ar.d_v + begin, sh ? sh + begin : nullptr, e8 - begin, lr, beta1, beta2, eps, t, o));
    if (o != s) { MFVAE_CUDA(cudaEventRecord(h->opt_ev, o)); h->opt_pending = true; }
    return 0;
  }
  return comm_unpack(c.local, c.dtype, h->ar.d_grad, begin, e8, s);
}
// make `stream` wait for every optimizer sweep issued by mfvae_allreduce_grads(do_adam = 1) so far
int mfvae_opt_join(MfvaeHandle h, void* stream) {
  MFVAE_CHECK(h, "null handle");
  if (h->opt_stream && h->opt_ev) MFVAE_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->opt_ev, 0));
  h->opt_pending = false;
  return 0;
}
// sum of the 4 loss scalars of the step in flight across ranks (in place in the handle's loss slots), on `stream`
int mfvae_allreduce_losses(MfvaeHandle h, void* stream) {
  MFVAE_CHECK(h && h->comm.world >= 2 && h->ws, "comm: not bound");
  { static const char* dbg = getenv("MFVAE_DP_DEBUG"); if (dbg && strstr(dbg, "noloss")) return 0; }
  return comm_allreduce_scalars(h->comm, losses_ptr(h), 4, losses_ptr(h), static_cast<cudaStream_t>(stream));
}

// One call = one train step (SURVEY 8b `mfvae_train_step`): forward + ELBO + backward, the data-parallel exchange when
// mfvae_comm_bind has been called (per bucket, on the library's own high-priority communication stream, behind that bucket's
// events), and Adam (per bucket, overlapped with the rest of backward).  `t` is the 1-based optimizer step.  pipeline = 1
// returns without ordering `stream` behind the decoder block's optimizer sweep (see mfvae_opt_join).  This is the sequence the
// Python host issues call by call (mfvae_b200/model.py::train_step); a C host gets it in one.
int mfvae_train_step(MfvaeHandle h, const MfvaeBatch* b, float lr, float beta1, float beta2, float eps, int64_t t, int32_t pipeline,
                     MfvaeOutputs* out, void* stream) {
  MFVAE_CHECK(h && b, "null handle or batch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MFVAE_TRY(mfvae_fwd_bwd(h, b, out, stream));
  if (h->comm.world < 2) return mfvae_adam_step_overlapped(h, lr, beta1, beta2, eps, t, stream);
  if (!h->comm_stream) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    MFVAE_CUDA(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, hi));
    MFVAE_CUDA(cudaEventCreateWithFlags(&h->comm_done_ev, cudaEventDisableTiming));
  }
  cudaStream_t cs = h->comm_stream;
  MFVAE_TRY(mfvae_loss_wait(h, cs));
  MFVAE_TRY(mfvae_allreduce_losses(h, cs));
  for (int32_t i = 0; i < static_cast<int32_t>(h->buckets.size()); ++i) {
    const int64_t bb = h->buckets[i].begin, be = std::min(h->buckets[i].end, h->optimized_elems);
    if (be <= bb) continue;
    MFVAE_TRY(mfvae_bucket_wait(h, i, cs));
    MFVAE_TRY(mfvae_bucket_read_wait(h, i, cs));
    MFVAE_TRY(mfvae_allreduce_grads(h, bb, be, 1, lr, beta1, beta2, eps, t, cs));
  }
  MFVAE_CUDA(cudaEventRecord(h->comm_done_ev, cs));
  MFVAE_CUDA(cudaStreamWaitEvent(s, h->comm_done_ev, 0));
  if (!pipeline) MFVAE_TRY(mfvae_opt_join(h, stream));
  return 0;
}

uint64_t mfvae_launch_count(void) { return g_launch_count; }

int mfvae_profile_enable(MfvaeHandle h, int32_t on) {
  MFVAE_CHECK(h, "null handle");
  if (on) {
    while (h->prof_ev.size() < 2 * h->gemms.size()) {
      cudaEvent_t e; MFVAE_CUDA(cudaEventCreate(&e)); h->prof_ev.push_back(e);
    }
    h->prof_hit.assign(h->gemms.size(), 0);
    for (auto& e : h->enc_prof_ev) if (!e) MFVAE_CUDA(cudaEventCreate(&e));
    h->enc_prof_hit = false;
  }
  h->profiling = on != 0;
  return 0;
}

int32_t mfvae_profile_read(MfvaeHandle h, MfvaeGemmTiming* out, int32_t capacity) {
  if (!h || !out) return -1;
  int32_t n = 0;
  for (size_t i = 0; i < h->gemms.size() && i < h->prof_hit.size() && n < capacity; ++i) {
    if (!h->prof_hit[i]) continue;
    if (cudaEventSynchronize(h->prof_ev[2 * i + 1]) != cudaSuccess) return -1;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]) != cudaSuccess) return -1;
    const GemmOp& o = h->gemms[i];
    const int kind = (o.epi == kEpiAccum) ? 2 : ((o.b_rs == 1 && o.b_cs != 1) ? 1 : 0);
    out[n++] = MfvaeGemmTiming{o.M, o.N, o.K, o.G, kind, ms};
  }
  if (h->enc_prof_hit && n < capacity) {
    // the fused encoder chain as one item: kind 3, groups = agents, M = batch, N = 1, K = MACs per (agent, sample) of the chain,
    // so that 2 * groups * M * N * K is its flop count like every other row
    if (cudaEventSynchronize(h->enc_prof_ev[1]) != cudaSuccess) return -1;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->enc_prof_ev[0], h->enc_prof_ev[1]) != cudaSuccess) return -1;
    int macs = 0;
    for (int l = 0; l < h->ne; ++l) macs += ((l == 0) ? h->K0f : h->encK[l]) * h->encN[l];
    out[n++] = MfvaeGemmTiming{h->B, 1, macs, h->A, 3, ms};
    h->enc_prof_hit = false;
  }
  return n;
}

int32_t mfvae_bucket_count(MfvaeHandle h) { return h ? static_cast<int32_t>(h->buckets.size()) : -1; }
int mfvae_bucket(MfvaeHandle h, int32_t i, int64_t* begin, int64_t* end, void** event) {
  MFVAE_CHECK(h && i >= 0 && i < static_cast<int32_t>(h->buckets.size()), "bucket index out of range");
  if (begin) *begin = h->buckets[i].begin;
  if (end) *end = h->buckets[i].end;
  if (event) *event = h->buckets[i].ev;
  return 0;
}
int mfvae_loss_wait(MfvaeHandle h, void* stream) {
  MFVAE_CHECK(h && h->loss_ev, "no loss has been recorded");
  MFVAE_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->loss_ev, 0));
  return 0;
}
int mfvae_bucket_wait(MfvaeHandle h, int32_t i, void* stream) {
  MFVAE_CHECK(h && i >= 0 && i < static_cast<int32_t>(h->buckets.size()), "bucket index out of range");
  MFVAE_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->buckets[i].ev, 0));
  return 0;
}

// ---- standalone kernels ---------------------------------------------------------------------------
int mfvae_reparam_kl(const float* d_mu, const float* d_logvar, const float* d_eps_or_null, void* d_z, int32_t z_dtype,
                     int64_t batch, int32_t width, uint64_t seed, uint64_t step, int64_t sample0, int64_t batch_global,
                     float* d_kl_out, float* d_scratch, void* stream) {
  MFVAE_CHECK(d_mu && d_logvar && d_z && d_kl_out && d_scratch, "null argument");
  ReparamArgs a{};
  a.mu = d_mu; a.lv = d_logvar; a.lat_as = width; a.lat_bs = width;   // one "agent" of width A*L
  a.eps = d_eps_or_null; a.eps_ld = width; a.z = d_z; a.z_ld = width; a.z_dtype = z_dtype;
  a.B = batch; a.A = 1; a.L = width; a.seed = seed; a.step = step; a.sample0 = sample0;
  a.kl_scale = 1.0f / static_cast<float>(batch_global); a.kl_out = d_kl_out; a.scratch = d_scratch;
  return launch_reparam_kl_fwd(a, static_cast<cudaStream_t>(stream));
}

int mfvae_recon_loss(const float* d_recon, int32_t recon_ld, const float* d_target, int32_t target_ld, void* d_grad,
                     int32_t grad_ld, int32_t grad_dtype, int64_t batch, int32_t width, int32_t huber, float weight,
                     int64_t count_global, float* d_loss_out, float* d_scratch, void* stream) {
  MFVAE_CHECK(d_recon && d_target && d_loss_out && d_scratch, "null argument");
  ReconLossArgs a{};
  a.recon = d_recon; a.recon_ld = recon_ld; a.target = d_target; a.target_ld = target_ld;
  a.grad = d_grad; a.grad_ld = grad_ld; a.grad_dtype = grad_dtype; a.B = batch; a.width = width; a.huber = huber;
  a.grad_scale = static_cast<float>(static_cast<double>(weight) / static_cast<double>(count_global));
  a.loss_scale = static_cast<float>(1.0 / static_cast<double>(count_global));
  a.loss_out = d_loss_out; a.scratch = d_scratch;
  return launch_recon_loss(a, static_cast<cudaStream_t>(stream));
}

int mfvae_adam_flat(float* d_p, const float* d_g, float* d_m, float* d_v, void* d_shadow, int64_t n, float lr,
                    float beta1, float beta2, float eps, int64_t t, void* stream) {
  MFVAE_CHECK(d_p && d_g && d_m && d_v, "null argument");
  return launch_adam(d_p, d_g, d_m, d_v, static_cast<__nv_bfloat16*>(d_shadow), n, lr, beta1, beta2, eps, t,
                     static_cast<cudaStream_t>(stream));
}

int mfvae_philox_normal(float* d_out, int64_t batch, int32_t width, uint64_t seed, uint64_t step, int64_t sample0, void* stream) {
  MFVAE_CHECK(d_out, "null argument");
  return launch_philox_normal(d_out, batch, width, seed, step, sample0, static_cast<cudaStream_t>(stream));
}

int mfvae_gemm(int32_t engine, int32_t dtype, int32_t groups, int32_t M, int32_t N, int32_t K,
               const void* d_A, int64_t a_gs, int64_t a_rs, int64_t a_cs,
               const void* d_B, int64_t b_gs, int64_t b_rs, int64_t b_cs,
               void* d_C, int64_t c_gs, int64_t c_ld, int32_t c_dtype,
               const float* d_bias, int64_t bias_gs, int32_t epilogue,
               const void* d_aux, int64_t aux_gs, int64_t aux_ld, int32_t split_k, void* stream) {
  GemmOp o; o.G = groups; o.M = M; o.N = N; o.K = K; o.dtype = dtype;
  o.A = d_A; o.a_gs = a_gs; o.a_rs = a_rs; o.a_cs = a_cs;
  o.B = d_B; o.b_gs = b_gs; o.b_rs = b_rs; o.b_cs = b_cs;
  o.C = d_C; o.c_gs = c_gs; o.c_ld = c_ld; o.c_dtype = c_dtype;
  o.bias = d_bias; o.bias_gs = bias_gs; o.epi = epilogue; o.aux = d_aux; o.aux_gs = aux_gs; o.aux_ld = aux_ld;
  o.split_k = split_k < 1 ? 1 : split_k;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (engine == MFVAE_ENGINE_AUTO) engine = (dtype == kBF16) ? MFVAE_ENGINE_TCGEN05 : MFVAE_ENGINE_SIMT;
  if (engine == MFVAE_ENGINE_SIMT) return gemm_simt(o, s);
  TcPlan* p = nullptr;
  MFVAE_TRY(gemm_tc_plan(o, &p));
  int r = gemm_tc_run(p, s);
  if (r == 0 && cudaStreamSynchronize(s) != cudaSuccess) r = mfvae::fail(__FILE__, __LINE__, "tcgen05 GEMM failed at synchronize");
  gemm_tc_free(p);
  return r;
}

}  // extern "C"
