"""Drop-in for the reference ``torch_ver/model.py`` on a B200.

Same module-level names and call signatures as the reference (``/root/reference/torch_ver/model.py``):
``kl_weight``, ``r_weight`` (:5-6), ``loss_vae_fn`` (:8), ``loss_s_r_vae_fn`` (:19), ``Encoder`` (:43),
``ActionEncoder`` (:60), ``reparameterize`` (:77), ``Decoder`` (:84), ``MAVAE`` (:101) — but the arithmetic
of ``MAVAE.forward``, the ELBO loss, the backward pass and the Adam update runs in hand-written sm_100a
CUDA behind the C ABI of ``include/mfvae.h`` (``libmfvae_b200.so``).  There is no CPU / eager fallback:
calling the model without a CUDA device raises.

What stays host-side Python: packing the reference's dict-of-tensors inputs into the packed batch the
kernels read, parameter bookkeeping (every ``nn.Parameter`` is a *view* into one flat fp32 arena, its
``.grad`` a view into the gradient arena, so ``state_dict()`` / ``load_state_dict()`` / ``parameters()``
behave as in the reference without copies), and the autograd bridge that lets the reference's own call
sequence ``model(...) -> loss_s_r_vae_fn(...) -> loss.backward() -> optimizer.step()``
(``torch_ver/main.py:87-97``) drive the engine.

Reference quirks kept on purpose: the per-agent encoders and action tables live in
plain dicts and are therefore not registered / not optimised (model.py:112,114) unless
``optimize_encoders=True``; the never-called ``decoder`` (model.py:127) exists and appears in
``state_dict()``; outputs are ``(recon_state, recon_reward, mu_all: list, log_var_all: list)``.
"""
import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib as L

kl_weight = 0.0025
r_weight = 0.005

import os as _os
_DP_DEBUG = _os.environ.get("MFVAE_DP_DEBUG", "")     # "nocomm" / "noloss": scaling-loss attribution experiments only

_ENC_HIDDEN = (64, 64, 256)
_DEC_HIDDEN = (1024, 256, 64, 256, 1024)


# ----------------------------------------------------------------------------------------------
# small torch modules with the reference's attribute layout (``.net`` = Sequential of Linear / ReLU)
# ----------------------------------------------------------------------------------------------
def _sequential(widths: Sequence[int], device=None) -> nn.Sequential:
    mods: List[nn.Module] = []
    for i in range(len(widths) - 1):
        mods.append(nn.Linear(widths[i], widths[i + 1], device=device))
        if i + 2 < len(widths):
            mods.append(nn.ReLU())
    return nn.Sequential(*mods)


class _MLP(nn.Module):
    HIDDEN: Sequence[int] = ()

    def __init__(self, in_dim, out_dim, hidden=None, device=None):
        super().__init__()
        hidden = tuple(self.HIDDEN if hidden is None else hidden)
        self.net = _sequential((in_dim, *hidden, out_dim), device=device)

    def forward(self, x):
        return self.net(x)


class Encoder(_MLP):
    """in -> 64 -> 64 -> 256 -> out (reference model.py:43-57)."""
    HIDDEN = _ENC_HIDDEN


class ActionEncoder(_MLP):
    """in -> 64 -> out, continuous-action variant (reference model.py:60-74)."""
    HIDDEN = (64,)


class Decoder(_MLP):
    """in -> 1024 -> 256 -> 64 -> 256 -> 1024 -> out (reference model.py:84-98)."""
    HIDDEN = _DEC_HIDDEN


def reparameterize(mu, log_var):
    """mu + eps * exp(log_var / 2) with eps ~ N(0, 1) (reference model.py:77-81); plain torch, kept for API
    parity — ``MAVAE.forward`` draws its eps inside the fused CUDA kernel (counter-based Philox)."""
    return mu + torch.randn_like(mu) * torch.exp(0.5 * log_var)


# ----------------------------------------------------------------------------------------------
# packed batch
# ----------------------------------------------------------------------------------------------
class PackedBatch:
    """Device-resident batch in the layout the kernels read: what ``create_dataset``
    (reference trainer.py:7-45) produces, without the per-agent dict indirection.

    obs  [B, S] fp32   observations of all agents concatenated in codebook order; may be bfloat16 (a host ring that keeps
                       observations in bf16 ships half the PCIe bytes; the bf16 engine rounds them on arrival anyway, so the
                       step is bit-identical to feeding the fp32 values those bf16 numbers came from)
    act  [B, A] fp32   float-coded discrete action per agent (replay_buffer.py:76); continuous actions:
                       [B, sum(action_dim)] action vectors concatenated in codebook order
    next [B, S] fp32   next observations (= ``next_states`` target)      (optional; bfloat16 allowed: ROUNDS THE TARGET, the
                       loss moves at the 1e-3 level -- an explicit opt-in of the caller)
    rew  [B, A] fp32   rewards (= ``rewards`` target)                     (optional)
    idx  [B, A] fp32   agent-index column of ``idx_state`` or None = codebook order
    eps  [B, A*L] fp32 explicit normal draw or None = Philox(seed, step, sample0 + row)
    """

    def __init__(self, obs, act, next=None, rew=None, idx=None, eps=None, sample0=0, batch_global=None):
        self.obs, self.act, self.next, self.rew, self.idx, self.eps = obs, act, next, rew, idx, eps
        self.sample0 = int(sample0)
        self.batch_global = int(batch_global if batch_global is not None else obs.shape[0])

    @property
    def batch(self):
        return self.obs.shape[0]


def _f32c(t, device):
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


# ----------------------------------------------------------------------------------------------
# autograd bridge
# ----------------------------------------------------------------------------------------------
class _ForwardFn(torch.autograd.Function):
    """Outputs of the CUDA forward as autograd nodes: any torch loss built on them back-propagates into
    ``mfvae_backward_ext``."""

    @staticmethod
    def forward(ctx, anchor, model, recon_s, recon_r, latent):
        ctx.model = model
        ctx.serial = model._serial
        return recon_s.view_as(recon_s), recon_r.view_as(recon_r), latent.view_as(latent)

    @staticmethod
    def backward(ctx, g_rs, g_rr, g_lat):
        m = ctx.model
        if ctx.serial != m._serial:
            raise RuntimeError("mfvae_b200: backward through a stale forward (the engine keeps one batch of activations)")
        m._backward_ext(g_rs, g_rr, g_lat)
        return None, None, None, None, None


class _FusedLossFn(torch.autograd.Function):
    """The fused ELBO: values come from ``mfvae_loss``; ``backward`` runs ``mfvae_backward``."""

    @staticmethod
    def forward(ctx, anchor, model, losses):
        ctx.model = model
        ctx.serial = model._serial
        return losses[0].clone(), losses[1].clone(), losses[2].clone(), losses[3].clone()

    @staticmethod
    def backward(ctx, g0, g1, g2, g3):
        m = ctx.model
        if ctx.serial != m._serial:
            raise RuntimeError("mfvae_b200: backward through a stale forward (the engine keeps one batch of activations)")
        m._backward_fused()
        return None, None, None


def _owner(t):
    return getattr(t, "_mfvae_owner", None)


def loss_s_r_vae_fn(recon_s, recon_r, s_hat, r_hat, mean_all, logvar_all, device, using_huber_loss=True):
    """Reference model.py:19-40.  When the reconstructions come straight from a ``MAVAE`` of this package the
    fused CUDA loss (+ gradient seeds) is used; otherwise (e.g. ``Trainer.training_model``'s denormalised
    rewards) the same formula is evaluated with torch ops and autograd flows into ``mfvae_backward_ext``."""
    m = _owner(recon_s)
    if (m is not None and m is _owner(recon_r) and getattr(recon_s, "_mfvae_serial", -1) == m._serial
            and getattr(recon_r, "_mfvae_serial", -2) == m._serial and mean_all is m._last_mu
            and logvar_all is m._last_lv and torch.is_grad_enabled()):
        return m._fused_loss(s_hat, r_hat, L.LOSS_HUBER if using_huber_loss else L.LOSS_MSE)
    F = torch.nn.functional
    s_hat = s_hat.to(device); r_hat = r_hat.to(device)
    recon_s = recon_s.to(device); recon_r = recon_r.to(device)
    fn = F.huber_loss if using_huber_loss else F.mse_loss
    s_loss = fn(s_hat, recon_s)
    rr_loss = fn(r_hat, recon_r)
    kl = _kl_sum_of_means(mean_all, logvar_all)
    loss = s_loss + rr_loss * r_weight + kl * kl_weight
    return loss, s_loss, rr_loss, kl


def _joint_of(t):
    """The MAVAE whose current (recon_state, recon_reward) outputs `t` is the column concatenation of, or None.
    ``loss_vae_fn`` takes ONE reconstruction matrix [B, S + A] (the reference's legacy joint decoder output, model.py:168);
    with the two-headed model the caller builds it as ``torch.cat([recon_s, recon_r], 1)``, which autograd records as a
    CatBackward0 node fed by outputs 0 and 1 of this package's forward node."""
    fn = getattr(t, "grad_fn", None)
    if fn is None or fn.name() != "CatBackward0" or len(fn.next_functions) != 2:
        return None
    (n0, i0), (n1, i1) = fn.next_functions
    m = getattr(n0, "_mfvae_model", None) if n0 is not None else None
    if m is None or n0 is not n1 or (i0, i1) != (0, 1) or n0 is not m._fwd_node or t.dim() != 2:
        return None
    return m


def loss_vae_fn(y, y_hat, mean_all, logvar_all, device):
    """Reference model.py:8-16: joint MSE over [next_state | reward] + KL; returns the scalar loss.  When one of the two
    matrices is ``torch.cat([recon_s, recon_r], 1)`` of this package's current forward pass (the loss is symmetric in
    ``y`` / ``y_hat``), value and gradient seeds come from the fused CUDA loss (``MFVAE_LOSS_JOINT_MSE``); otherwise the
    same formula is evaluated with torch ops and autograd flows into ``mfvae_backward_ext``."""
    for recon, target in ((y_hat, y), (y, y_hat)):
        m = _joint_of(recon)
        if (m is not None and mean_all is m._last_mu and logvar_all is m._last_lv and torch.is_grad_enabled()
                and not target.requires_grad):
            S = m._cur.obs.shape[1]
            return m._fused_loss(target[:, :S], target[:, S:], L.LOSS_JOINT_MSE)[0]
    y = y.to(device); y_hat = y_hat.to(device)
    return torch.nn.functional.mse_loss(y_hat, y) + _kl_sum_of_means(mean_all, logvar_all) * kl_weight


def _kl_sum_of_means(mean_all, logvar_all):
    kl = 0.0
    for mu, lv in zip(mean_all, logvar_all):
        kl = kl + torch.mean(-0.5 * torch.sum(1 + lv - mu ** 2 - torch.exp(lv), 1), 0)
    return kl


# ----------------------------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------------------------
class MAVAE(nn.Module):
    """Reference ctor signature (model.py:102) plus keyword-only engine options."""

    def __init__(self, idx_features: int, obs_features: int, action_features: int, descrete_act: bool,
                 agents: list, obs_dim: dict, action_dim: dict, device: str, *,
                 precision: str = "bf16", engine: str = "auto", enc_hidden: Optional[Sequence[int]] = None,
                 dec_hidden: Optional[Sequence[int]] = None, optimize_encoders: bool = False,
                 include_dead_decoder: bool = True, seed: int = 0x5EED, huber: bool = True,
                 fusion: str = "auto"):
        super().__init__()
        self.obs_dim = obs_dim
        self.act_dim = action_dim
        self.feature = obs_features
        self.device = device
        self.descrete_act = descrete_act
        self.agents = list(agents)
        self.idx_features, self.action_features = idx_features, action_features
        self.enc_hidden = tuple(_ENC_HIDDEN if enc_hidden is None else enc_hidden)
        self.dec_hidden = tuple(_DEC_HIDDEN if dec_hidden is None else dec_hidden)
        self.precision = precision
        self.optimize_encoders = bool(optimize_encoders)
        self.philox_seed = int(seed)
        self.philox_step = 0
        self.data_parallel = False          # set by Trainer / enable_data_parallel()
        self._pg = None

        tdev = torch.device(device)
        self._tdev = tdev
        A = len(self.agents)
        cfg = L.MfvaeConfig()
        cfg.n_agents, cfg.idx_features, cfg.latent, cfg.act_features = A, idx_features, obs_features, action_features
        cfg.n_enc_hidden = len(self.enc_hidden)
        cfg.n_dec_hidden = len(self.dec_hidden)
        for i, w in enumerate(self.enc_hidden):
            cfg.enc_hidden[i] = w
        for i, w in enumerate(self.dec_hidden):
            cfg.dec_hidden[i] = w
        self._obs_arr = (C.c_int32 * A)(*[int(obs_dim[a]) for a in self.agents])
        self._act_arr = (C.c_int32 * A)(*[int(action_dim[a]) for a in self.agents])
        cfg.obs_dim = C.cast(self._obs_arr, C.POINTER(C.c_int32))
        cfg.n_act = C.cast(self._act_arr, C.POINTER(C.c_int32))
        cfg.kl_weight, cfg.r_weight, cfg.huber = kl_weight, r_weight, int(huber)
        cfg.precision = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16}[precision]
        cfg.engine = {"auto": L.ENGINE_AUTO, "simt": L.ENGINE_SIMT, "tcgen05": L.ENGINE_TCGEN05}[engine]
        cfg.optimize_encoders = int(self.optimize_encoders)
        cfg.continuous_act = 0 if descrete_act else 1
        cfg.act_hidden = ActionEncoder.HIDDEN[0]
        if fusion == "auto":          # measurement knob: MFVAE_FUSION overrides what "auto" resolves to
            fusion = _os.environ.get("MFVAE_FUSION", "auto")
        cfg.fusion = {"auto": L.FUSE_AUTO, "none": L.FUSE_NONE, "encoder": L.FUSE_ENCODER, "loss": L.FUSE_LOSS,
                      "encoder+loss": L.FUSE_ENCODER | L.FUSE_LOSS, "nofold": L.FUSE_NONE | L.FUSE_NOFOLD_IDX | L.FUSE_NOFOLD_ACT, "nofold_idx": L.FUSE_NONE | L.FUSE_NOFOLD_IDX,
                      "nofold_act": L.FUSE_NONE | L.FUSE_NOFOLD_ACT}[fusion]
        self._cfg = cfg
        lib = L.lib()
        self._h = C.c_void_p()
        dev_index = -1 if tdev.type != "cuda" else (tdev.index if tdev.index is not None else torch.cuda.current_device())
        L.check(lib.mfvae_create(C.byref(cfg), dev_index, C.byref(self._h)))
        self._on_gpu = dev_index >= 0

        # ---- arenas: one flat tensor each; parameters are views ----
        n = lib.mfvae_arena_elems(self._h)
        self._n_opt = lib.mfvae_optimized_elems(self._h)
        self._arena = torch.zeros(n, dtype=torch.float32, device=tdev)
        self._grad = torch.zeros(n, dtype=torch.float32, device=tdev)
        self._m = torch.zeros(n, dtype=torch.float32, device=tdev)
        self._v = torch.zeros(n, dtype=torch.float32, device=tdev)
        self._shadow = torch.zeros(n, dtype=torch.bfloat16, device=tdev) if precision == "bf16" else None
        self._adam_t = 0
        nt = lib.mfvae_tensor_count(self._h)
        table = (L.MfvaeTensorInfo * nt)()
        L.check(lib.mfvae_tensor_table(self._h, table, nt))
        self._table = list(table)
        if self._on_gpu:
            ar = L.MfvaeArenas(L.ptr(self._arena), L.ptr(self._grad), L.ptr(self._m), L.ptr(self._v), L.ptr(self._shadow))
            L.check(lib.mfvae_bind_arenas(self._h, C.byref(ar)))

        self._views: List[tuple] = []          # (parameter, grad view)
        self._build_modules(include_dead_decoder)
        self.reset_parameters()

        self._ws = None
        self._ws_batch = -1
        self._serial = 0
        self._arena_version = -1
        self._dirty = False
        self._grads_pending = False           # a backward pass has run and neither step() nor zero_grad() has consumed it
        self._losses_reduced = False          # data parallel: the 4 loss scalars of the step in flight are already all-reduced
        self._fwd_node = None
        rl = [t for t in self._table if t.kind in (L.T_RLIN_W, L.T_RLIN_B)]
        self._rl_range = (min(t.offset for t in rl) // 8 * 8, (max(t.offset + t.rows * t.ld for t in rl) + 7) // 8 * 8)
        self._anchor = torch.zeros(1, device=tdev, requires_grad=True)
        self._cur: Optional[PackedBatch] = None
        self._cb = None
        self._last_mu = self._last_lv = None
        self._comm_stream = None
        self._native_comm = False
        self._opt_pending = False
        self.comm_info = {"route": "none"}

    # ------------------------------------------------------------------ construction helpers
    def _view(self, info, arena):
        v = arena[info.offset: info.offset + info.rows * info.ld].view(info.rows, info.ld)
        if info.cols != info.ld:
            v = v[:, :info.cols]
        return v

    def _param(self, info, squeeze=False):
        v, g = self._view(info, self._arena), self._view(info, self._grad)
        if squeeze:
            v, g = v[0], g[0]
        p = nn.Parameter(v)
        p.grad = g
        self._views.append((p, g))
        return p

    def _arena_linear(self, w_info, b_info):
        lin = nn.Linear(w_info.cols, w_info.rows, device="meta")
        lin.weight = self._param(w_info)
        lin.bias = self._param(b_info, squeeze=True)
        return lin

    def _arena_mlp(self, cls, infos):
        """cls instance whose ``.net`` Linear layers are arena views; infos = [(w, b), ...]."""
        obj = cls.__new__(cls)
        nn.Module.__init__(obj)
        mods = []
        for i, (w, b) in enumerate(infos):
            mods.append(self._arena_linear(w, b))
            if i + 1 < len(infos):
                mods.append(nn.ReLU())
        obj.net = nn.Sequential(*mods)
        return obj

    def _build_modules(self, include_dead_decoder):
        by = {}
        for t in self._table:
            by[(t.kind, t.agent, t.layer)] = t
        A = len(self.agents)
        ne, nd = len(self.enc_hidden) + 1, len(self.dec_hidden) + 1
        self.encoders = {}
        self.idx_emb = nn.Embedding(A, self.idx_features, device="meta")
        self.idx_emb.weight = self._param(by[(L.T_IDX_EMB, -1, -1)])
        self.action_encoder = {}
        for ai, a in enumerate(self.agents):
            self.encoders[a] = self._arena_mlp(Encoder, [(by[(L.T_ENC_W, ai, l)], by[(L.T_ENC_B, ai, l)]) for l in range(ne)])
            if self.descrete_act:
                t = by[(L.T_ACT_TABLE, ai, -1)]
                emb = nn.Embedding(t.rows, t.cols, device="meta")
                emb.weight = self._param(t)
                self.action_encoder[a] = emb
            else:       # model.py:123: ActionEncoder(act_dim, action_features), arena-backed like the encoders
                self.action_encoder[a] = self._arena_mlp(ActionEncoder, [(by[(L.T_ACTENC_W, ai, l)], by[(L.T_ACTENC_B, ai, l)])
                                                                         for l in range(2)])
        S = sum(int(self.obs_dim[a]) for a in self.agents)
        din = (self.feature + self.action_features) * A
        if include_dead_decoder:   # model.py:127 — constructed, registered, never called
            self.decoder = Decoder(din, S + A, hidden=self.dec_hidden, device=self._tdev)
        self.state_decoder = self._arena_mlp(Decoder, [(by[(L.T_SDEC_W, -1, l)], by[(L.T_SDEC_B, -1, l)]) for l in range(nd)])
        self.reward_decoder = self._arena_mlp(Decoder, [(by[(L.T_RDEC_W, -1, l)], by[(L.T_RDEC_B, -1, l)]) for l in range(nd)])
        self.reward_linear = self._arena_linear(by[(L.T_RLIN_W, -1, -1)], by[(L.T_RLIN_B, -1, -1)])
        if self.optimize_encoders:      # extension: make the per-agent nets visible to parameters()
            self._extra = nn.ModuleList(list(self.encoders.values()) + list(self.action_encoder.values()))

    @torch.no_grad()
    def reset_parameters(self):
        """torch default initialisers (what the reference gets from nn.Linear / nn.Embedding) +
        reward_linear = ones / zeros (model.py:131-132)."""
        for p, _ in self._views:
            p.zero_()
        def init_linear(lin):
            nn.init.kaiming_uniform_(lin.weight, a=math.sqrt(5))
            bound = 1.0 / math.sqrt(lin.weight.shape[1])
            nn.init.uniform_(lin.bias, -bound, bound)
        nn.init.normal_(self.idx_emb.weight)
        for a in self.agents:
            for mod in self.encoders[a].net:
                if isinstance(mod, nn.Linear):
                    init_linear(mod)
            if self.descrete_act:
                nn.init.normal_(self.action_encoder[a].weight)
            else:
                for mod in self.action_encoder[a].net:
                    if isinstance(mod, nn.Linear):
                        init_linear(mod)
        for dec in (self.state_decoder, self.reward_decoder):
            for mod in dec.net:
                if isinstance(mod, nn.Linear):
                    init_linear(mod)
        nn.init.ones_(self.reward_linear.weight)
        nn.init.zeros_(self.reward_linear.bias)

    def named_arena_tensors(self) -> Dict[str, torch.Tensor]:
        """Every arena-backed tensor under the oracle's naming (encoders.<agent>.net.<i>.weight, ...)."""
        out = {"idx_emb.weight": self.idx_emb.weight}
        for a in self.agents:
            for n, p in self.encoders[a].named_parameters():
                out[f"encoders.{a}.{n}"] = p
            for n, p in self.action_encoder[a].named_parameters():     # Embedding: "weight"; ActionEncoder: "net.0.weight", ...
                out[f"action_encoder.{a}.{n}"] = p
        for pre in ("state_decoder", "reward_decoder", "reward_linear"):
            for n, p in getattr(self, pre).named_parameters():
                out[f"{pre}.{n}"] = p
        return out

    @torch.no_grad()
    def load_named(self, tensors: Dict[str, torch.Tensor]):
        mine = self.named_arena_tensors()
        for k, p in mine.items():
            p.copy_(tensors[k].to(p.device, torch.float32))
        if hasattr(self, "decoder"):
            for n, p in self.decoder.named_parameters():
                k = f"decoder.{n}"
                if k in tensors:
                    p.copy_(tensors[k].to(p.device, torch.float32))

    def save(self, path):
        """Reference model.py:175-176: ``torch.save(self.state_dict(), path)`` -- the 39 registered tensors only (the
        per-agent encoders and action tables live in plain dicts there and are NOT saved).  The file loads with
        ``strict=True`` into the reference ``MAVAE`` and back."""
        self.synchronize_optimizer()
        torch.save(self.state_dict(), path)

    # ---- the checkpoint the reference forgets: everything needed to resume bit-identically ----
    CHECKPOINT_FORMAT = 1

    def checkpoint(self) -> dict:
        """``state_dict`` (the reference's 39 keys, loadable by the reference itself) + the unregistered per-agent encoders /
        action tables under the oracle's names + Adam exp_avg / exp_avg_sq / step of the optimised prefix + the Philox
        (seed, step) of the reparameterisation stream.  CPU tensors."""
        self.synchronize_optimizer()
        n = self._n_opt
        sd = {k: v.detach().cpu().clone() for k, v in self.state_dict().items()}
        extra = {k: p.detach().cpu().clone() for k, p in self.named_arena_tensors().items() if k not in sd}
        return {"format": self.CHECKPOINT_FORMAT, "state_dict": sd, "unregistered": extra,
                "adam": {"exp_avg": self._m[:n].detach().cpu().clone(), "exp_avg_sq": self._v[:n].detach().cpu().clone(),
                         "step": int(self._adam_t), "optimized_elems": int(n)},
                "philox": {"seed": int(self.philox_seed), "step": int(self.philox_step)},
                "config": {"agents": list(self.agents), "latent": int(self.feature), "enc_hidden": list(self.enc_hidden),
                           "dec_hidden": list(self.dec_hidden), "optimize_encoders": bool(self.optimize_encoders)}}

    def save_checkpoint(self, path):
        torch.save(self.checkpoint(), path)

    @torch.no_grad()
    def load_checkpoint(self, path_or_dict):
        ck = torch.load(path_or_dict, map_location="cpu") if not isinstance(path_or_dict, dict) else path_or_dict
        if ck.get("format") != self.CHECKPOINT_FORMAT:
            raise RuntimeError("mfvae_b200: unknown checkpoint format")
        if int(ck["adam"]["optimized_elems"]) != int(self._n_opt):
            raise RuntimeError("mfvae_b200: checkpoint was written by a model with a different optimised parameter set")
        self.load_state_dict(ck["state_dict"], strict=True)
        mine = self.named_arena_tensors()
        for k, t in ck["unregistered"].items():
            mine[k].copy_(t.to(mine[k].device, torch.float32))
        n = self._n_opt
        self._m[:n].copy_(ck["adam"]["exp_avg"]); self._v[:n].copy_(ck["adam"]["exp_avg_sq"])
        self._adam_t = int(ck["adam"]["step"])
        self.philox_seed, self.philox_step = int(ck["philox"]["seed"]), int(ck["philox"]["step"])
        self._dirty = True
        self._grads_pending = False

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                L.lib().mfvae_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ engine plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self._tdev).cuda_stream)

    def _require_gpu(self):
        if not self._on_gpu:
            raise RuntimeError("mfvae_b200: MAVAE was created on a non-CUDA device; the hot path has no CPU fallback")

    def _bind(self, B):
        if B == self._ws_batch:
            return
        lib = L.lib()
        need = lib.mfvae_workspace_bytes(self._h, B)
        if need <= 0:
            raise RuntimeError("mfvae_b200: bad batch size")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.zeros(need, dtype=torch.uint8, device=self._tdev)
        torch.cuda.current_stream(self._tdev).synchronize()
        L.check(lib.mfvae_bind_workspace(self._h, L.ptr(self._ws), self._ws.numel(), B))
        self._ws_batch = B

    def mark_dirty(self):
        """Tell the engine that master weights were edited behind autograd's back (``p.data.mul_()``, ``p.data.copy_()``:
        a ``.data`` alias carries its own version counter, so the edit is invisible to ``_sync_shadow``): the bf16 weight
        shadow the tensor cores read is rebuilt before the next forward."""
        self._dirty = True

    def _sync_shadow(self, dropin=False):
        """Parameter views share the arena's version counter: any in-place edit made through them (optimizer,
        load_state_dict, ``Trainer.pop``) bumps it and the bf16 shadow is refreshed before the next forward.  Edits through
        ``.data`` do not (see ``mark_dirty``); the reference's own POP-ART idiom does exactly that to ``reward_linear``
        (``torch_ver/trainer.py:73-74``), so the drop-in ``forward`` re-casts that 1.6 K-element block on every call."""
        if self._shadow is None:
            return
        v = self._arena._version
        if v != self._arena_version or self._dirty:
            self.synchronize_optimizer()
            L.check(L.lib().mfvae_refresh_shadow(self._h, self._stream()))
            self._arena_version = v
            self._dirty = False
        elif dropin:
            b, e = self._rl_range
            L.check(L.lib().mfvae_refresh_shadow_range(self._h, b, e, self._stream()))

    def _cbatch(self, pb: PackedBatch):
        cb = L.MfvaeBatch()
        cb.d_obs, cb.d_act = pb.obs.data_ptr(), pb.act.data_ptr()
        cb.d_next = pb.next.data_ptr() if pb.next is not None else None
        cb.d_rew = pb.rew.data_ptr() if pb.rew is not None else None
        cb.d_idx = pb.idx.data_ptr() if pb.idx is not None else None
        cb.d_eps = pb.eps.data_ptr() if pb.eps is not None else None
        cb.obs_bf16 = int(pb.obs.dtype == torch.bfloat16)
        cb.next_bf16 = int(pb.next is not None and pb.next.dtype == torch.bfloat16)
        for name, t in (("obs", pb.obs), ("next", pb.next)):
            if t is not None and (t.dtype not in (torch.float32, torch.bfloat16) or not t.is_contiguous()):
                raise TypeError(f"mfvae_b200: PackedBatch.{name} must be a contiguous float32 or bfloat16 matrix")
        cb.batch, cb.sample0, cb.batch_global = pb.batch, pb.sample0, pb.batch_global
        cb.seed, cb.step = self.philox_seed, self.philox_step
        return cb

    def _ws_view(self, ptr, rows, ld, cols):
        off = ptr - self._ws.data_ptr()
        t = self._ws[off: off + rows * ld * 4].view(torch.float32).view(rows, ld)
        return t[:, :cols] if cols != ld else t

    def pack(self, idx_state: dict, actions: dict, eps=None) -> PackedBatch:
        """dict-of-tensors inputs of the reference forward -> PackedBatch on the device."""
        keys = list(idx_state.keys())
        if keys != self.agents:
            raise NotImplementedError("idx_state must list the agents in the model's agent order")
        dev = self._tdev
        cols = [_f32c(idx_state[a], dev) for a in keys]
        obs = torch.cat([c[:, 1:] for c in cols], dim=1)
        idx = torch.stack([c[:, 0] for c in cols], dim=1).contiguous()
        # nn.Embedding raises IndexError on an out-of-range index (model.py:142,146); so does the drop-in path (one
        # device->host flag per call; PackedBatch users skip this and the kernels clamp, forward and backward alike)
        A = len(self.agents)
        bad = ((idx < 0) | (idx >= A)).any()
        if self.descrete_act:
            act = torch.cat([_f32c(actions[a], dev).reshape(-1, 1) for a in keys], dim=1)
            nmax = torch.tensor([float(self.act_dim[a]) for a in keys], device=dev)
            bad = bad | ((act < 0) | (act >= nmax)).any()
        # create_dataset (trainer.py:21) writes the codebook index into column 0: row b of agent a carries a.  Then the
        # id-embedding is a per-agent constant and the engine folds it into encoder layer 0's bias (idx = None).
        codebook = (idx == torch.arange(A, device=dev, dtype=idx.dtype)).all()
        bad, codebook = (bool(x) for x in torch.stack([bad, codebook]).cpu())
        if bad:
            raise IndexError("index out of range in self")
        if codebook:
            idx = None
        if not self.descrete_act:       # continuous: [B, act_dim_a] vectors, concatenated in agent order
            act = torch.cat([_f32c(actions[a], dev).reshape(cols[0].shape[0], -1) for a in keys], dim=1).contiguous()
        if eps is not None:
            eps = _f32c(eps, dev)
        return PackedBatch(obs, act, idx=idx, eps=eps)

    def forward(self, idx_state, actions=None, eps=None):
        """Reference signature ``forward(idx_state: dict, actions: dict)``; a ``PackedBatch`` may be passed as
        the first argument instead.  ``eps`` ([B, A*L]) overrides the Philox draw (parity tests)."""
        self._require_gpu()
        pb = idx_state if isinstance(idx_state, PackedBatch) else self.pack(idx_state, actions, eps)
        if pb.batch < 2:
            # the reference's .squeeze() at model.py:159 makes B == 1 fail with a dimension error
            raise RuntimeError("Tensors must have same number of dimensions (B == 1 is unsupported, as in the reference)")
        lib = L.lib()
        self._bind(pb.batch)
        self._sync_shadow(dropin=True)
        self._serial += 1
        self._cur = pb
        self._cb = self._cbatch(pb)
        out = L.MfvaeOutputs()
        L.check(lib.mfvae_forward(self._h, C.byref(self._cb), C.byref(out), self._stream()))
        B, A, Lt = pb.batch, len(self.agents), self.feature
        S = pb.obs.shape[1]
        recon_s = self._ws_view(out.d_recon_s, B, out.recon_s_ld, S)
        recon_r = self._ws_view(out.d_recon_r, B, out.recon_r_ld, A)
        latent = self._ws_view(out.d_latent, A * B, 2 * Lt, 2 * Lt).view(A, B, 2 * Lt)
        self._losses = self._ws_view(out.d_losses, 1, 4, 4)[0]
        self._fwd_node = None
        if torch.is_grad_enabled():
            recon_s, recon_r, latent = _ForwardFn.apply(self._anchor, self, recon_s, recon_r, latent)
            self._fwd_node = recon_s.grad_fn
            if self._fwd_node is not None:
                self._fwd_node._mfvae_model = self
        for t in (recon_s, recon_r):
            t._mfvae_owner, t._mfvae_serial = self, self._serial
        mu_all = [latent[a, :, :Lt] for a in range(A)]
        lv_all = [latent[a, :, Lt:] for a in range(A)]
        self._last_mu, self._last_lv = mu_all, lv_all
        if self.training:
            self.philox_step += 1
        return recon_s, recon_r, mu_all, lv_all

    # ------------------------------------------------------------------ loss / backward / optimizer
    def _set_weights(self, weights=None):
        """(kl_weight, r_weight[, s_weight]); None = this module's globals, read at call time like the reference
        (model.py:5-6,34,39), with the state term weighted 1."""
        lib = L.lib()
        if weights is None:
            L.check(lib.mfvae_set_loss_weights(self._h, kl_weight, r_weight))
        else:
            kw, rw = float(weights[0]), float(weights[1])
            sw = float(weights[2]) if len(weights) > 2 else 1.0
            L.check(lib.mfvae_set_loss_weights3(self._h, kw, rw, sw))

    def _fused_loss(self, s_hat, r_hat, kind, weights=None):
        lib = L.lib()
        pb = self._cur
        pb.next = _f32c(s_hat, self._tdev)
        pb.rew = _f32c(r_hat, self._tdev)
        self._cb.d_next, self._cb.d_rew = pb.next.data_ptr(), pb.rew.data_ptr()
        self._set_weights(weights)
        L.check(lib.mfvae_loss(self._h, C.byref(self._cb), kind, self._stream()))
        if self.data_parallel:
            # each rank holds its share of the global means: reduce BEFORE the values are handed to the caller
            import torch.distributed as dist
            dist.all_reduce(self._losses, group=self._pg)
            self._losses_reduced = True
        return _FusedLossFn.apply(self._anchor, self, self._losses)

    def _attach_grads(self):
        for p, g in self._views:
            if p.grad is not g:
                p.grad = g

    def _check_accumulation(self):
        # the backward pass re-zeroes the gradient arena and its two largest wgrads use plain stores: a second backward
        # before step() / zero_grad() would silently keep only the last gradient where the reference accumulates
        if self._grads_pending:
            raise RuntimeError("mfvae_b200: a second backward() arrived before optimizer.step() / zero_grad(): gradient "
                               "accumulation across backward passes is not supported by the engine")

    def grads_consumed(self):
        """optimizer.zero_grad() / step(): the gradients of the last backward pass have been used or dropped."""
        self._grads_pending = False

    def _backward_fused(self):
        self._check_accumulation()
        L.check(L.lib().mfvae_backward(self._h, C.byref(self._cb), self._stream()))
        self._after_backward()

    def _backward_ext(self, g_rs, g_rr, g_lat):
        self._check_accumulation()
        # data parallel: a torch loss is a mean over the LOCAL batch, the all-reduce is a plain sum -> seeds / world
        scale = 1.0 / self._world() if self.data_parallel else 1.0

        def prep(g):
            if g is None:
                return None
            g = g.to(torch.float32)
            return (g * scale if scale != 1.0 else g).contiguous()
        g_rs, g_rr, g_lat = prep(g_rs), prep(g_rr), prep(g_lat)
        L.check(L.lib().mfvae_backward_ext(self._h, C.byref(self._cb), L.ptr(g_rs), g_rs.shape[1] if g_rs is not None else 0,
                                           L.ptr(g_rr), g_rr.shape[1] if g_rr is not None else 0, L.ptr(g_lat), self._stream()))
        self._after_backward()

    def _after_backward(self):
        self._attach_grads()
        self._grads_pending = True
        if self.data_parallel:
            self._allreduce_grads()

    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size(self._pg)

    # ---- data parallel: bucketed all-reduce on a side stream, launched as each bucket's event fires ----
    def _bind_native_comm(self, process_group):
        """The exchange step as the library's own kernels (csrc/comm.cu): one symmetric buffer per rank mapped into every
        rank (torch.distributed._symmetric_memory = CUDA VMM + fabric handles: plumbing), its peer / multicast pointers and
        signal pads handed to mfvae_comm_bind.  Returns False when symmetric memory cannot be set up on this machine."""
        import torch.distributed as dist
        mode = _os.environ.get("MFVAE_DP_COMM", "native")            # "nccl": the round-1 route (torch.distributed all_reduce)
        if mode == "nccl":
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            lib = L.lib()
            bf16 = self.precision == "bf16" and _os.environ.get("MFVAE_DP_PAYLOAD", "bf16") != "fp32"
            nbytes = lib.mfvae_comm_window_bytes(self._h, int(bf16))
            group = process_group if process_group is not None else dist.group.WORLD
            win = symm.empty(nbytes, dtype=torch.uint8, device=self._tdev)
            win.zero_()
            hdl = symm.rendezvous(win, group=group.group_name)
            mc = int(hdl.multicast_ptr) if _os.environ.get("MFVAE_DP_MULTICAST", "1") != "0" else 0
            L.check(lib.mfvae_comm_bind(self._h, hdl.rank, hdl.world_size, C.c_void_p(int(hdl.buffer_ptrs_dev)), C.c_void_p(mc) if mc else None,
                                        C.c_void_p(int(hdl.signal_pad_ptrs_dev)), int(hdl.signal_pad_size), C.c_void_p(win.data_ptr()), nbytes,
                                        int(bf16), int(_os.environ.get("MFVAE_DP_BLOCKS", "0"))))
            self._comm_win, self._comm_hdl = win, hdl
            self.comm_info = {"route": "native", "payload": "bf16" if bf16 else "fp32", "multicast": bool(mc), "window_bytes": int(nbytes)}
            torch.cuda.synchronize(self._tdev)
            dist.barrier(group=process_group)
            return True
        except Exception as e:                                        # no fabric / multicast support: keep the NCCL route
            if mode == "native_strict":
                raise
            self.comm_info = {"route": "nccl", "why": f"{type(e).__name__}: {e}"[:200]}
            return False

    def enable_data_parallel(self, process_group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._pg = process_group
        self.data_parallel = dist.get_world_size(process_group) > 1
        if self.data_parallel:
            # replicas must start identical: parameters, Adam moments / step count and the Philox stream of rank 0
            for t in (self._arena, self._m, self._v):
                dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
            meta = torch.tensor([self._adam_t, self.philox_step, self.philox_seed], dtype=torch.int64, device=self._tdev)
            dist.broadcast(meta, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
            self._adam_t, self.philox_step, self.philox_seed = (int(x) for x in meta.cpu())
            self._dirty = True
            self._native_comm = bool(self._on_gpu and self._bind_native_comm(process_group))
        if self._on_gpu and self._comm_stream is None:
            # greatest priority: the exchange kernels are small and latency-critical, and every GEMM of backward is a persistent
            # grid that owns all CTA slots -- a default-priority stream gets its CTAs in only at kernel boundaries, late
            self._comm_stream = torch.cuda.Stream(self._tdev, priority=-1)
        if self._on_gpu and self.data_parallel:
            # SMs the persistent GEMM grids leave to the NCCL kernels running beside backward.  Measured at 2 GPUs: 0 / 16 / 32
            # reserved -> 1.091 / 1.084 / 1.113 ms per step (NCCL's CTAs co-reside with ours), so the default is 0.
            reserve = int(_os.environ.get("MFVAE_SM_RESERVE", "0"))
            L.check(L.lib().mfvae_set_sm_reserve(self._h, reserve))
            self._ws_batch = -1 if self._ws is None else self._ws_batch

    def grad_buckets(self):
        lib = L.lib()
        out = []
        for i in range(lib.mfvae_bucket_count(self._h)):
            b, e = C.c_int64(), C.c_int64()
            L.check(lib.mfvae_bucket(self._h, i, C.byref(b), C.byref(e), None))
            if e.value > b.value:
                out.append((i, b.value, e.value))
        return out

    def synchronize_optimizer(self):
        """Order the current stream behind every optimizer sweep a pipelined ``train_step(..., pipeline=True)`` left in flight.
        ``forward`` / ``train_step`` / ``test_step`` / ``adam_step`` / checkpointing do it themselves; call it before reading
        parameter tensors directly after a pipelined step."""
        if self._opt_pending and self._on_gpu:
            L.check(L.lib().mfvae_opt_join(self._h, self._stream()))
        self._opt_pending = False

    def _allreduce_grads(self, adam=None, pipeline=False):
        """SUM all-reduce of the gradient buckets (the loss-gradient kernels already divide by the GLOBAL batch, so
        no averaging pass is needed) and of the 4 partial loss scalars.  On the GPU each bucket is reduced on the
        communication stream as soon as its completion event fires, overlapping the rest of backward; with
        ``adam=(lr, betas, eps)`` the fused Adam of that bucket follows on the same stream right behind its all-reduce,
        so the optimizer sweep is hidden behind the encoder half of backward as well."""
        import torch.distributed as dist
        buckets = self.grad_buckets()
        if not self._on_gpu:           # host-logic path exercised by the gloo tests; no compute happens on CPU
            for _, b, e in buckets:
                dist.all_reduce(self._grad[b:e], group=self._pg)
            if getattr(self, "_losses", None) is not None and not self._losses_reduced:
                dist.all_reduce(self._losses, group=self._pg)
            self._losses_reduced = False
            return
        lib = L.lib()
        main = torch.cuda.current_stream(self._tdev)
        cs = self._comm_stream
        csp = C.c_void_p(cs.cuda_stream)
        if adam is not None:
            self._adam_t += 1
        if self._native_comm and not any(k in _DP_DEBUG for k in ("nocomm", "lateloss")):
            # the library's own exchange kernels: per bucket pack -> two-shot reduce over peer memory -> Adam from the window
            lr, betas, eps = adam if adam is not None else (0.0, (0.9, 0.999), 1e-8)
            if not self._losses_reduced:
                L.check(lib.mfvae_loss_wait(self._h, csp))
                L.check(lib.mfvae_allreduce_losses(self._h, csp))
            self._losses_reduced = False
            for i, b, e in buckets:
                L.check(lib.mfvae_bucket_wait(self._h, i, csp))
                if adam is not None:
                    L.check(lib.mfvae_bucket_read_wait(self._h, i, csp))
                e_opt = min(e, self._n_opt)
                L.check(lib.mfvae_allreduce_grads(self._h, b, e_opt, int(adam is not None), float(lr), float(betas[0]), float(betas[1]),
                                                  float(eps), max(self._adam_t, 1), csp))
            main.wait_stream(cs)
            if adam is not None:
                if pipeline:
                    self._opt_pending = True        # joined by the next forward inside the library, or by synchronize_optimizer()
                else:
                    L.check(lib.mfvae_opt_join(self._h, C.c_void_p(main.cuda_stream)))
            return
        dbg = _DP_DEBUG            # measurement switches (MFVAE_DP_DEBUG): never set in production
        with torch.cuda.stream(cs):
            # the loss scalars are final before backward starts: reduce them first, beside backward, not in the step's tail
            early_loss = "nocomm" not in dbg and "noloss" not in dbg and "lateloss" not in dbg and not self._losses_reduced
            self._losses_reduced = False
            if early_loss:
                L.check(lib.mfvae_loss_wait(self._h, csp))
                dist.all_reduce(self._losses, group=self._pg, async_op=True).wait()
            for i, b, e in buckets:
                L.check(lib.mfvae_bucket_wait(self._h, i, csp))
                if "nocomm" not in dbg:
                    dist.all_reduce(self._grad[b:e], group=self._pg, async_op=True).wait()     # cs waits for NCCL
                if adam is not None:
                    L.check(lib.mfvae_bucket_read_wait(self._h, i, csp))     # backward may still be reading this bucket's weights
                    lr, betas, eps = adam
                    L.check(lib.mfvae_adam_range(self._h, b, min(e, self._n_opt), float(lr), float(betas[0]), float(betas[1]),
                                                 float(eps), self._adam_t, csp))
            if "lateloss" in dbg:
                cs.wait_stream(main)
                dist.all_reduce(self._losses, group=self._pg, async_op=True).wait()
        main.wait_stream(cs)

    def adam_step(self, lr, betas=(0.9, 0.999), eps=1e-8, overlapped=False):
        """Fused Adam over the registered (optimised) prefix of the arena; also refreshes the bf16 shadow.
        ``overlapped`` (single GPU, called right after backward): the decoder block is updated on an internal stream
        as soon as its gradients are final, concurrently with the encoder half of backward."""
        self._require_gpu()
        self.synchronize_optimizer()
        self._adam_t += 1
        self._grads_pending = False
        fn = L.lib().mfvae_adam_step_overlapped if overlapped else L.lib().mfvae_adam_step
        L.check(fn(self._h, float(lr), float(betas[0]), float(betas[1]), float(eps), self._adam_t, self._stream()))

    @torch.no_grad()
    def test_step(self, pb: PackedBatch, loss_weights=None):
        """Forward + ELBO without backward / update (jax_ver/trainer.py:86-90 ``test_step``; the torch driver has no
        evaluation loop).  Returns the device tensor [loss, s_loss, r_loss, kl_loss]; the Philox step is not advanced."""
        self._require_gpu()
        lib = L.lib()
        self._bind(pb.batch)
        self._sync_shadow()
        self._serial += 1
        self._cur, self._cb = pb, self._cbatch(pb)
        out = L.MfvaeOutputs()
        L.check(lib.mfvae_forward(self._h, C.byref(self._cb), C.byref(out), self._stream()))
        self._set_weights(loss_weights)
        L.check(lib.mfvae_loss(self._h, C.byref(self._cb), L.LOSS_DEFAULT, self._stream()))
        self._losses = self._ws_view(out.d_losses, 1, 4, 4)[0]
        if self.data_parallel:
            import torch.distributed as dist
            dist.all_reduce(self._losses, group=self._pg)
        return self._losses

    def train_step_c(self, pb: PackedBatch, lr: float, betas=(0.9, 0.999), eps=1e-8, loss_weights=None, pipeline=False):
        """``train_step`` through the single C entry ``mfvae_train_step`` (what a C host calls: the exchange runs on the
        library's own communication stream).  Same results as ``train_step``; data parallel needs the native exchange."""
        self._require_gpu()
        if self.data_parallel and not self._native_comm:
            raise RuntimeError("mfvae_b200: mfvae_train_step needs the native exchange (MFVAE_DP_COMM=native)")
        lib = L.lib()
        self._bind(pb.batch)
        self._sync_shadow()
        self._set_weights(loss_weights)
        self._serial += 1
        self._grads_pending = self._losses_reduced = False
        self._cur, self._cb = pb, self._cbatch(pb)
        out = L.MfvaeOutputs()
        self._adam_t += 1
        L.check(lib.mfvae_train_step(self._h, C.byref(self._cb), float(lr), float(betas[0]), float(betas[1]), float(eps), self._adam_t,
                                     int(bool(pipeline)), C.byref(out), self._stream()))
        self._losses = self._ws_view(out.d_losses, 1, 4, 4)[0]
        self.philox_step += 1
        self._opt_pending = bool(pipeline and self.data_parallel)
        return self._losses

    def train_step(self, pb: PackedBatch, lr: float, betas=(0.9, 0.999), eps=1e-8, loss_weights=None, pipeline=False):
        """Fast path: forward + fused ELBO + backward (+ all-reduce) + Adam, no autograd graph.
        Returns the device tensor [loss, s_loss, r_loss, kl_loss].  ``loss_weights`` = (kl_weight, r_weight[, s_weight])
        overrides the module globals (e.g. the jax_ver weighting (0.1, 0.5, 0.5)).
        ``pipeline`` (data parallel): return without ordering the caller's stream behind the decoder block's optimizer sweep;
        the next step's encoder half then overlaps it (the library orders the decoder half itself).  See
        ``synchronize_optimizer`` before touching parameter tensors by hand."""
        self._require_gpu()
        lib = L.lib()
        self._bind(pb.batch)
        self._sync_shadow()
        self._set_weights(loss_weights)
        self._serial += 1
        self._grads_pending = self._losses_reduced = False
        self._cur, self._cb = pb, self._cbatch(pb)
        out = L.MfvaeOutputs()
        L.check(lib.mfvae_fwd_bwd(self._h, C.byref(self._cb), C.byref(out), self._stream()))
        self._losses = self._ws_view(out.d_losses, 1, 4, 4)[0]
        self.philox_step += 1
        if self.data_parallel:
            self._allreduce_grads(adam=(lr, betas, eps), pipeline=pipeline)      # per-bucket all-reduce + Adam on the communication stream
        else:
            self.adam_step(lr, betas, eps, overlapped=True)
        return self._losses
