"""CPU oracle for the MF-VAE (class ``MAVAE``) training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mfvae_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and there only as the checker or as the timed CPU arm.

It is a from-scratch, functional restatement (plain dict of tensors + pure functions, torch
CPU arithmetic) of the reference algorithm in ``/root/reference/torch_ver``:

* forward            -> ``torch_ver/model.py:134-173`` (per-agent loop, id-embedding ``:142``,
                        encoder MLP ``:43-57``, latent split ``:149-150``, reparameterize ``:77-81``,
                        action embedding ``:145-146``, concat order ``:158-164``, two decoders
                        ``:84-98,169-170`` and ``reward_linear`` ``:130-132,170``)
* loss               -> ``torch_ver/model.py:19-40`` (``loss_s_r_vae_fn``) and ``:8-16`` (``loss_vae_fn``)
* staging            -> ``torch_ver/trainer.py:7-45`` (``create_dataset``)
* optimizer          -> ``torch.optim.Adam`` defaults as used at ``torch_ver/main.py:52`` /
                        ``torch_ver/trainer.py:62`` and ``CosineAnnealingLR(T_max=50, eta_min=1e-4)``
                        at ``torch_ver/main.py:53``

Third-party arithmetic (torch, version unpinned by the reference; 2.11.0 here) is restated from its
published formulas: ``nn.Linear`` = x W^T + b, Huber(delta=1, mean), Adam without amsgrad / weight
decay, closed-form cosine annealing.

PARITY PIN: the reference ships no tests or golden vectors for this path, so this
oracle is pinned against *outputs of the reference itself*: ``tests/golden/make_golden.py`` imports
``/root/reference/torch_ver/{model,trainer}.py`` unchanged, loads the weights produced by
``init_params`` below, injects the same eps, and records losses / gradients / post-Adam
parameters; ``tests/test_oracle_golden.py`` replays them through this file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

KL_WEIGHT = 0.0025   # torch_ver/model.py:5
R_WEIGHT = 0.005     # torch_ver/model.py:6


# ----------------------------------------------------------------------------------------------
# problem description
# ----------------------------------------------------------------------------------------------
@dataclass
class Spec:
    """Shape description of one MAVAE instance (reference ctor args, ``torch_ver/model.py:102``)."""
    agents: List[str]
    obs_dim: Dict[str, int]
    n_act: Dict[str, int]
    idx_features: int = 64        # IDX_FEATURES  torch_ver/main.py:30
    latent: int = 64              # OBS_FEATURES  torch_ver/main.py:31
    act_features: int = 64        # ACT_FEATURES  torch_ver/main.py:32
    enc_hidden: Sequence[int] = (64, 64, 256)              # torch_ver/model.py:46
    dec_hidden: Sequence[int] = (1024, 256, 64, 256, 1024)  # torch_ver/model.py:87
    include_dead_decoder: bool = False   # the never-called ``decoder`` of model.py:127
    discrete_act: bool = True            # DESCRETE_ACT torch_ver/main.py:33; False -> ActionEncoder MLP (model.py:60-74,123,148)
    act_dim: Optional[Dict[str, int]] = None     # continuous action width per agent (used when discrete_act is False)
    act_hidden: Sequence[int] = (64,)            # ActionEncoder.HIDDEN torch_ver/model.py:63

    @property
    def n_agents(self) -> int:
        return len(self.agents)

    @property
    def state_dim(self) -> int:
        return sum(self.obs_dim[a] for a in self.agents)

    @property
    def dec_in(self) -> int:
        return (self.latent + self.act_features) * self.n_agents


def simple_tag_spec(n_adv: int = 30, n_good: int = 10, n_obst: int = 20, **kw) -> Spec:
    """Dims of PettingZoo ``simple_tag_v3`` as configured at ``torch_ver/src/env.py:27``
    (env itself is not needed: only its integer dims enter the VAE)."""
    agents = [f"adversary_{i}" for i in range(n_adv)] + [f"agent_{i}" for i in range(n_good)]
    n = n_adv + n_good
    obs = {}
    for a in agents:
        base = 2 + 2 + 2 * n_obst + 2 * (n - 1)
        obs[a] = base + 2 * (n_good if a.startswith("adversary") else n_good - 1)
    return Spec(agents=agents, obs_dim=obs, n_act={a: 5 for a in agents}, **kw)


def tiny_spec(n_agents: int = 3, **kw) -> Spec:
    agents = [f"adversary_{i}" for i in range(n_agents - 1)] + ["agent_0"]
    obs = {a: (14 if a.startswith("adversary") else 12) for a in agents}
    return Spec(agents=agents, obs_dim=obs, n_act={a: 5 for a in agents}, **kw)


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
def _linear_names(prefix: str, n_layers: int) -> List[Tuple[str, str]]:
    # nn.Sequential(Linear, ReLU, Linear, ...) numbers the Linear modules 0,2,4,...
    return [(f"{prefix}.net.{2 * i}.weight", f"{prefix}.net.{2 * i}.bias") for i in range(n_layers)]


def layer_dims(spec: Spec) -> Dict[str, List[Tuple[int, int]]]:
    """(out,in) of every Linear, keyed by module prefix."""
    out = {}
    for a in spec.agents:
        widths = [spec.idx_features + spec.obs_dim[a], *spec.enc_hidden, 2 * spec.latent]
        out[f"encoders.{a}"] = [(widths[i + 1], widths[i]) for i in range(len(widths) - 1)]
    if not spec.discrete_act:
        for a in spec.agents:
            widths = [spec.act_dim[a], *spec.act_hidden, spec.act_features]
            out[f"action_encoder.{a}"] = [(widths[i + 1], widths[i]) for i in range(len(widths) - 1)]
    for name, od in (("state_decoder", spec.state_dim), ("reward_decoder", spec.n_agents),
                     ("decoder", spec.state_dim + spec.n_agents)):
        if name == "decoder" and not spec.include_dead_decoder:
            continue
        widths = [spec.dec_in, *spec.dec_hidden, od]
        out[name] = [(widths[i + 1], widths[i]) for i in range(len(widths) - 1)]
    return out


def init_params(spec: Spec, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic (numpy PCG64) weights from the same distribution families torch's default init
    uses: Linear W,b ~ U(+-1/sqrt(fan_in)); Embedding ~ N(0,1); reward_linear = ones / zeros
    (``torch_ver/model.py:131-132``).  The *values* are what both the reference and this oracle are
    loaded with when golden vectors are made; bit-equality with torch's own generator is irrelevant."""
    rng = np.random.default_rng(seed)
    P: Dict[str, torch.Tensor] = {}

    def put(name, arr):
        P[name] = torch.from_numpy(np.ascontiguousarray(arr.astype(np.float32))).to(dtype)

    put("idx_emb.weight", rng.standard_normal((spec.n_agents, spec.idx_features)))
    dims = layer_dims(spec)
    for prefix, lst in dims.items():
        for (wn, bn), (o, i) in zip(_linear_names(prefix, len(lst)), lst):
            bound = 1.0 / math.sqrt(i)
            put(wn, rng.uniform(-bound, bound, size=(o, i)))
            put(bn, rng.uniform(-bound, bound, size=(o,)))
    if spec.discrete_act:
        for a in spec.agents:
            put(f"action_encoder.{a}.weight", rng.standard_normal((spec.n_act[a], spec.act_features)))
    A = spec.n_agents
    put("reward_linear.weight", np.ones((A, A)))
    put("reward_linear.bias", np.zeros((A,)))
    return P


def registered_names(spec: Spec) -> List[str]:
    """Names the reference optimizer sees (``model.parameters()``): the per-agent encoders and action
    tables live in plain dicts (``torch_ver/model.py:112,114``) and are therefore NOT registered."""
    names = ["idx_emb.weight"]
    dims = layer_dims(spec)
    for prefix in ("decoder", "state_decoder", "reward_decoder"):
        if prefix in dims:
            for wn, bn in _linear_names(prefix, len(dims[prefix])):
                names += [wn, bn]
    names += ["reward_linear.weight", "reward_linear.bias"]
    return names


# ----------------------------------------------------------------------------------------------
# synthetic data + staging
# ----------------------------------------------------------------------------------------------
def synth_transition(spec: Spec, batch: int, seed: int = 0, reward_scale: float = 1.0) -> Dict[str, np.ndarray]:
    """A ``cpprb.sample``-shaped dict (``torch_ver/src/replay_buffer.py:62-81,107-108``):
    float32 arrays ``{agent}_{observations,next_observations,actions,rewards}`` of shape (B, dim)."""
    rng = np.random.default_rng(seed)
    t = {}
    for a in spec.agents:
        o = spec.obs_dim[a]
        t[f"{a}_observations"] = rng.standard_normal((batch, o)).astype(np.float32)
        t[f"{a}_next_observations"] = rng.standard_normal((batch, o)).astype(np.float32)
        if spec.discrete_act:
            t[f"{a}_actions"] = rng.integers(0, spec.n_act[a], size=(batch, 1)).astype(np.float32)
        else:       # MPE continuous actions live in [0, 1]
            t[f"{a}_actions"] = rng.uniform(0.0, 1.0, size=(batch, spec.act_dim[a])).astype(np.float32)
        t[f"{a}_rewards"] = (reward_scale * rng.standard_normal((batch, 1))).astype(np.float32)
    return t


def stage_batch(transition: Dict[str, np.ndarray], codebook: Dict[str, int]):
    """Restates ``create_dataset`` (``torch_ver/trainer.py:7-45``): per agent the codebook index is
    prepended as column 0 of the observation; rewards and next observations are column-concatenated
    in codebook order; the joint ``[next_states | rewards]`` matrix is also returned."""
    idx_state, acts, nxt, rew = {}, {}, [], []
    for a, k in codebook.items():
        obs = transition[a + "_observations"]
        col = np.full((obs.shape[0], 1), k, dtype=obs.dtype)
        idx_state[a] = torch.from_numpy(np.hstack([col, obs]))
        acts[a] = torch.from_numpy(transition[a + "_actions"])
        nxt.append(transition[a + "_next_observations"])
        rew.append(transition[a + "_rewards"])
    next_states = np.hstack(nxt)
    rewards = np.hstack(rew)
    joint = np.hstack([next_states, rewards])
    return idx_state, acts, torch.from_numpy(joint), torch.from_numpy(next_states), torch.from_numpy(rewards)


# ----------------------------------------------------------------------------------------------
# forward / loss
# ----------------------------------------------------------------------------------------------
class _RoundBF16(torch.autograd.Function):
    """Round-to-nearest-even to bfloat16 in the forward and / or the backward direction.  Used only by the
    ``emulate_bf16`` mode, which evaluates the reference algorithm with the operand rounding points of the bf16 CUDA
    path (GEMM operands and stored activations / activation gradients in bf16, fp32 accumulation, fp32 biases,
    losses and master weights) so that the tensor-core kernels can be checked to ~1e-3 instead of against an fp32
    run whose ReLU masks differ in a fraction of units."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.to(torch.bfloat16).to(x.dtype) if fwd else x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return (g.to(torch.bfloat16).to(g.dtype) if ctx.bwd else g), None, None


def _q(x, on, fwd=True, bwd=True):
    return _RoundBF16.apply(x, fwd, bwd) if on else x


def _mlp(P, prefix: str, n_layers: int, x: torch.Tensor, q: bool = False, out_fwd: bool = False, first=None) -> torch.Tensor:
    """``first``: pre-activation of layer 0 computed by the caller (the folded layer-0 forms of ``emulate_bf16``)."""
    names = _linear_names(prefix, n_layers)
    for i, (wn, bn) in enumerate(names):
        if i == 0 and first is not None:
            x = first
        else:
            x = torch.nn.functional.linear(x, _q(P[wn], q, True, False), P[bn])      # nn.Linear: x W^T + b
        if i + 1 < n_layers:
            x = _q(torch.relu(x), q)
    return _q(x, q, out_fwd, True)


def forward(P: Dict[str, torch.Tensor], spec: Spec, idx_state: Dict[str, torch.Tensor],
            actions: Dict[str, torch.Tensor], eps: Dict[str, torch.Tensor], emulate_bf16: bool = False, fold: bool = True,
            fold_idx: Optional[bool] = None):
    """``MAVAE.forward`` with the normal draw made explicit.  Returns
    ``(recon_state[B,S], recon_reward[B,A], mu_all, log_var_all)``.

    ``emulate_bf16`` re-evaluates the same algorithm with the operand rounding points of the bf16 CUDA path.  With
    ``fold`` (what the CUDA path does, csrc/fold.cu) two algebraically identical regroupings move rounding points:
    encoder layer 0 is ``bf16(obs) . bf16(W0[:, I:])^T + (b0 + W0[:, :I] . emb)`` with the bracket in fp32, and layer 0 of
    each decoder is ``bf16(z) . bf16(W0[:, :A L])^T + sum_a T_a[:, act_a] + b`` with ``T_a = bf16(W0[:, cols_a] . table_a^T)``.
    In fp32 (``emulate_bf16=False``) the regrouping is invisible at 1e-5 and the reference's literal form is evaluated.
    ``fold_idx`` (default = ``fold``) switches the encoder regrouping alone: the fused encoder kernel keeps those columns dense."""
    q = emulate_bf16
    fold_idx = fold if fold_idx is None else fold_idx
    L = spec.latent
    A = spec.n_agents
    n_enc = len(spec.enc_hidden) + 1
    n_dec = len(spec.dec_hidden) + 1
    I = spec.idx_features
    zs, embs, mus, lvs, act_idx = [], [], [], [], []
    for k_a, a in enumerate(idx_state.keys()):
        x = idx_state[a].to(P["idx_emb.weight"].dtype)
        ids = x[:, 0].to(torch.int32).long()                     # model.py:142  .int()
        e_id = torch.nn.functional.embedding(ids, P["idx_emb.weight"])
        codebook = bool((ids == k_a).all())                       # create_dataset's index column (trainer.py:21)
        if q and fold_idx and codebook:
            W0, b0 = P[f"encoders.{a}.net.0.weight"], P[f"encoders.{a}.net.0.bias"]
            first = torch.nn.functional.linear(_q(x[:, 1:], q), _q(W0[:, I:], q, True, False)) + (e_id @ W0[:, :I].t() + b0)
            lat = _mlp(P, f"encoders.{a}", n_enc, None, q, first=first)
        else:
            h = _q(torch.cat([e_id, x[:, 1:]], dim=1), q)
            lat = _mlp(P, f"encoders.{a}", n_enc, h, q)
        mu, lv = lat[:, :L], lat[:, L:]                           # model.py:149-150
        z = mu + eps[a].to(mu.dtype) * torch.exp(0.5 * lv)        # model.py:77-81
        if spec.discrete_act:
            ai = actions[a].to(torch.int32).long().reshape(-1)    # model.py:146
            act_idx.append(ai)
            embs.append(torch.nn.functional.embedding(ai, P[f"action_encoder.{a}.weight"]))
        else:                                                     # model.py:148: ActionEncoder MLP on the raw action vector
            embs.append(_mlp(P, f"action_encoder.{a}", len(spec.act_hidden) + 1, _q(actions[a].to(mu.dtype), q), q))
        zs.append(z); mus.append(mu); lvs.append(lv)
    if q and fold and spec.discrete_act:
        z_all = _q(torch.cat(zs, dim=-1), q)
        Kz, C = A * L, spec.act_features

        def layer0(prefix):
            W0, b0 = P[f"{prefix}.net.0.weight"], P[f"{prefix}.net.0.bias"]
            out = torch.nn.functional.linear(z_all, _q(W0[:, :Kz], q, True, False), b0)
            for i, a in enumerate(idx_state.keys()):
                T = _q(W0[:, Kz + i * C:Kz + (i + 1) * C] @ P[f"action_encoder.{a}.weight"].t(), q, True, False)   # [H, n_act]
                out = out + T.t()[act_idx[i]]
            return out
        recon_s = _mlp(P, "state_decoder", n_dec, None, q, first=layer0("state_decoder"))
        r = _mlp(P, "reward_decoder", n_dec, None, q, out_fwd=True, first=layer0("reward_decoder"))
    else:
        dec_in = _q(torch.cat(zs + embs, dim=-1), q)               # model.py:158-164: all z, then all act-emb
        recon_s = _mlp(P, "state_decoder", n_dec, dec_in, q)
        r = _mlp(P, "reward_decoder", n_dec, dec_in, q, out_fwd=True)
    recon_r = _q(torch.nn.functional.linear(r, _q(P["reward_linear.weight"], q, True, False), P["reward_linear.bias"]),
                 q, False, True)
    return recon_s, recon_r, mus, lvs


def huber_mean(x: torch.Tensor, y: torch.Tensor, delta: float = 1.0) -> torch.Tensor:
    d = (x - y).abs()
    return torch.where(d < delta, 0.5 * d * d, delta * (d - 0.5 * delta)).mean()


def kl_sum_of_means(mus, lvs) -> torch.Tensor:
    """model.py:35-37: sum over agents of batch-mean of -1/2 sum_latent(1 + lv - mu^2 - e^lv)."""
    kl = 0.0
    for mu, lv in zip(mus, lvs):
        kl = kl + (-0.5 * (1.0 + lv - mu * mu - torch.exp(lv)).sum(dim=1)).mean(dim=0)
    return kl


def loss_s_r(recon_s, recon_r, s_hat, r_hat, mus, lvs, huber: bool = True,
             kl_weight: float = KL_WEIGHT, r_weight: float = R_WEIGHT):
    """``loss_s_r_vae_fn`` (model.py:19-40).  Returns (loss, s_loss, r_loss, kl_loss)."""
    # same library entry points the reference calls (model.py:26-33); `huber_mean` above is their explicit
    # formula and tests/test_oracle_golden.py checks the two agree
    fn = torch.nn.functional.huber_loss if huber else torch.nn.functional.mse_loss
    s_loss = fn(s_hat, recon_s)
    r_loss = fn(r_hat, recon_r)
    kl = kl_sum_of_means(mus, lvs)
    return s_loss + r_weight * r_loss + kl_weight * kl, s_loss, r_loss, kl


def loss_joint_mse(y, y_hat, mus, lvs, kl_weight: float = KL_WEIGHT):
    """``loss_vae_fn`` (model.py:8-16)."""
    return ((y_hat - y) ** 2).mean() + kl_weight * kl_sum_of_means(mus, lvs)


# ----------------------------------------------------------------------------------------------
# optimizer / schedule
# ----------------------------------------------------------------------------------------------
def adam_update(p, g, m, v, t: int, lr: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam single-tensor formula (no amsgrad, no weight decay), t counted from 1:
    m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    bc1 = 1.0 - b1 ** t
    bc2 = 1.0 - b2 ** t
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def cosine_lr(step: int, base_lr: float = 0.005, t_max: int = 50, eta_min: float = 1e-4) -> float:
    """Closed form of CosineAnnealingLR; lr used by the ``step``-th optimizer.step() (0-based).
    The reference keeps stepping past T_max, which the closed form continues as a period-2*T_max wave."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * step / t_max)) / 2.0


# ----------------------------------------------------------------------------------------------
# one full train step (autograd), reference semantics
# ----------------------------------------------------------------------------------------------
@dataclass
class OracleState:
    spec: Spec
    P: Dict[str, torch.Tensor]
    m: Dict[str, torch.Tensor] = field(default_factory=dict)
    v: Dict[str, torch.Tensor] = field(default_factory=dict)
    t: int = 0
    optimize_encoders: bool = False


def grads(P, spec, idx_state, actions, eps, s_hat, r_hat, huber=True,
          kl_weight=KL_WEIGHT, r_weight=R_WEIGHT, emulate_bf16=False, fold=True, fold_idx=None):
    """Forward + loss + autograd backward.  Returns (losses 4-tuple of floats, dict of grads for every
    tensor that took part, outputs)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    rs, rr, mus, lvs = forward(leaves, spec, idx_state, actions, eps, emulate_bf16, fold, fold_idx)
    loss, sl, rl, kl = loss_s_r(rs, rr, s_hat.to(rs.dtype), r_hat.to(rs.dtype), mus, lvs, huber, kl_weight, r_weight)
    loss.backward()
    G = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    outs = (rs.detach(), rr.detach(), [m.detach() for m in mus], [l.detach() for l in lvs])
    return (float(loss.detach()), float(sl.detach()), float(rl.detach()), float(kl.detach())), G, outs


def train_step(st: OracleState, idx_state, actions, eps, s_hat, r_hat, lr: float, huber=True):
    """forward -> loss -> backward -> Adam on the registered tensors (or all, when optimize_encoders)."""
    losses, G, outs = grads(st.P, st.spec, idx_state, actions, eps, s_hat, r_hat, huber)
    st.t += 1
    names = list(st.P.keys()) if st.optimize_encoders else registered_names(st.spec)
    for n in names:
        if n not in G:
            continue                       # dead ``decoder``: grad None -> Adam skips
        if n not in st.m:
            st.m[n] = torch.zeros_like(st.P[n]); st.v[n] = torch.zeros_like(st.P[n])
        st.P[n], st.m[n], st.v[n] = adam_update(st.P[n], G[n], st.m[n], st.v[n], st.t, lr)
    return losses, G, outs


# ----------------------------------------------------------------------------------------------
# closed-form pieces of the loss tail (numpy, float64) for the bandwidth-kernel parity tests
# ----------------------------------------------------------------------------------------------
def np_reparam_kl(mu: np.ndarray, lv: np.ndarray, eps: np.ndarray, latent: int):
    """mu, lv, eps: [B, A*L].  z and the reference KL scalar (sum over agents of batch means ==
    total sum / B because every agent shares B)."""
    mu = mu.astype(np.float64); lv = lv.astype(np.float64); eps = eps.astype(np.float64)
    z = mu + eps * np.exp(0.5 * lv)
    kl = (-0.5 * (1.0 + lv - mu * mu - np.exp(lv))).sum() / mu.shape[0]
    return z, kl


def np_reparam_kl_bwd(dz, mu, lv, eps, kl_weight: float, batch_global: int):
    dz = dz.astype(np.float64); mu = mu.astype(np.float64); lv = lv.astype(np.float64); eps = eps.astype(np.float64)
    dmu = dz + kl_weight * mu / batch_global
    dlv = dz * eps * 0.5 * np.exp(0.5 * lv) + kl_weight * 0.5 * (np.exp(lv) - 1.0) / batch_global
    return dmu, dlv


def np_recon_loss(recon: np.ndarray, target: np.ndarray, huber: bool, weight: float, count_global: int):
    """mean Huber/MSE over ``count_global`` elements of (target, recon) and d(weight*loss)/d recon."""
    r = recon.astype(np.float64); t = target.astype(np.float64)
    d = r - t
    if huber:
        a = np.abs(d)
        val = np.where(a < 1.0, 0.5 * d * d, a - 0.5).sum() / count_global
        g = np.clip(d, -1.0, 1.0) * (weight / count_global)
    else:
        val = (d * d).sum() / count_global
        g = 2.0 * d * (weight / count_global)
    return val, g


def np_adam(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    p = p.astype(np.float64); g = g.astype(np.float64); m = m.astype(np.float64); v = v.astype(np.float64)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    p = p - (lr / (1 - b1 ** t)) * m / (np.sqrt(v) / math.sqrt(1 - b2 ** t) + eps)
    return p, m, v


# ----------------------------------------------------------------------------------------------
# Philox4x32-10 + Box-Muller, the counter-based eps the CUDA path draws (numpy restatement of the
# published Random123 algorithm; constants from Salmon et al., SC'11).
# ----------------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter: [...,4] uint32, key: [...,2] uint32 -> [...,4] uint32."""
    c = [counter[..., i].astype(np.uint32) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32); k1 = key[..., 1].astype(np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c[0].astype(np.uint64)
            p1 = _PHILOX_M1 * c[2].astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32); lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32); lo1 = p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + _PHILOX_W0).astype(np.uint32); k1 = (k1 + _PHILOX_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_normal(seed: int, step: int, sample0: int, n_samples: int, width: int) -> np.ndarray:
    """eps[B, width] exactly as the CUDA kernels draw it (see ``csrc/philox.cuh``):
    element (global sample s, column j) uses counter (s_lo, s_hi, j//4, step), key (seed_lo, seed_hi);
    the 4 outputs give 4 normals for columns 4*(j//4)..+3 via two Box-Muller pairs
    (u = ((x >> 9) + 0.5) * 2^-23, exactly representable in fp32;  r = sqrt(-2 ln u0);
    n0 = r cos(2 pi u1), n1 = r sin(2 pi u1))."""
    assert width % 4 == 0
    s = (np.arange(n_samples, dtype=np.uint64) + np.uint64(sample0))[:, None]
    q = np.arange(width // 4, dtype=np.uint64)[None, :]
    ctr = np.zeros((n_samples, width // 4, 4), dtype=np.uint32)
    ctr[..., 0] = (s & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[..., 1] = (s >> np.uint64(32)).astype(np.uint32)
    ctr[..., 2] = q.astype(np.uint32)
    ctr[..., 3] = np.uint32(step & 0xFFFFFFFF)
    key = np.zeros((n_samples, width // 4, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = (philox4x32_10(ctr, key) >> np.uint32(9)).astype(np.float64)
    u = (r + 0.5) * (1.0 / 8388608.0)
    rad0 = np.sqrt(-2.0 * np.log(u[..., 0])); rad1 = np.sqrt(-2.0 * np.log(u[..., 2]))
    th0 = 2.0 * np.pi * u[..., 1]; th1 = 2.0 * np.pi * u[..., 3]
    out = np.stack([rad0 * np.cos(th0), rad0 * np.sin(th0), rad1 * np.cos(th1), rad1 * np.sin(th1)], axis=-1)
    return out.reshape(n_samples, width)
