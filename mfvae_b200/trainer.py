"""Drop-in for the reference ``torch_ver/trainer.py``: ``create_dataset`` (:7-45) and ``Trainer`` (:48-119).

``Trainer`` keeps the reference's constructor, fields (``sigma, mu, nu, sigma_new, mu_new, beta, lr, loss_func,
loss, opt, device``) and methods (``art, pop, update_stats, normalize, denormalize, forward, backward, step,
training_model``).  ``opt`` is a :class:`FusedAdam`, a ``torch.optim.Optimizer`` whose ``step`` is one fused
CUDA kernel over the parameter arena (same update rule as ``torch.optim.Adam`` defaults, trainer.py:62).

Deliberate deviations from reference *bugs* (SURVEY.md section 8a, a16), each covered by a test:
  * ``training_model`` in the reference raises ``TypeError`` (adds a tuple to a float, trainer.py:112-113);
    here it accumulates the total loss (element 0 of the tuple).
  * mode ``'POPART'`` in the reference raises for every batch size (``pop()`` multiplies the [A, A] weight
    in place by a [B, A] ratio, trainer.py:72-74); here POP-ART statistics are per agent ([A], batch means),
    which is what the rescaling of ``reward_linear`` needs.  ``'Adam'`` and ``'ART'`` follow the reference
    formulas literally.
"""
from typing import Dict, Optional

import numpy as np
import torch
from tqdm import tqdm

from . import model as _model
from .model import MAVAE, PackedBatch


# ----------------------------------------------------------------------------------------------
# staging
# ----------------------------------------------------------------------------------------------
def _layout(transition, codebook):
    agents = list(codebook.keys())
    B = transition[agents[0] + "_observations"].shape[0]
    dims = [transition[a + "_observations"].shape[1] for a in agents]
    return agents, B, dims, int(sum(dims)), len(agents)


def _act_dims(transition, agents):
    """Action columns per agent: 1 for a float-coded discrete action, act_dim for a continuous action vector."""
    return [int(np.prod(transition[a + "_actions"].shape[1:])) or 1 for a in agents]


def _split_flat(flat, B, S, A, W=None):
    """Four dense matrices [obs | act | next | rew] carved out of one flat buffer of B * (2S + W + A) floats; W = total
    action columns (= A for discrete actions)."""
    W = A if W is None else W
    o0, o1, o2 = B * S, B * (S + W), B * (2 * S + W)
    return (flat[:o0].reshape(B, S), flat[o0:o1].reshape(B, W), flat[o1:o2].reshape(B, S), flat[o2:o2 + B * A].reshape(B, A))


def _pack_numpy(transition: Dict[str, np.ndarray], codebook: Dict[str, int], flat: Optional[np.ndarray] = None):
    """Single pass over the sampled dict into one flat float32 buffer holding obs[B,S], act[B,A], next[B,S],
    rew[B,A] back to back, agents in codebook order (no quadratic re-concatenation as in trainer.py:23-30)."""
    agents, B, dims, S, A = _layout(transition, codebook)
    adims = _act_dims(transition, agents)
    W = int(sum(adims))
    if flat is None:
        flat = np.empty(B * (2 * S + W + A), dtype=np.float32)
    obs, act, nxt, rew = _split_flat(flat, B, S, A, W)
    o = ao = 0
    for i, a in enumerate(agents):
        d = dims[i]
        obs[:, o:o + d] = transition[a + "_observations"]
        nxt[:, o:o + d] = transition[a + "_next_observations"]
        act[:, ao:ao + adims[i]] = transition[a + "_actions"].reshape(B, adims[i])
        rew[:, i] = transition[a + "_rewards"].reshape(B)
        o += d; ao += adims[i]
    return flat, (obs, act, nxt, rew), dims


def create_dataset(transition, codebook):
    """Reference trainer.py:7-45.  Returns ``(idx_state_all, action_all, next_state_rew, next_states, rewards)``:
    two dicts of CPU tensors ([B, 1 + O_a] with the codebook index in column 0; [B, 1]) and three CPU tensors
    ([B, S + A], [B, S], [B, A])."""
    _, (obs, act, nxt, rew), dims = _pack_numpy(transition, codebook)
    adims = _act_dims(transition, list(codebook.keys()))
    B = obs.shape[0]
    idx_state_all, action_all = {}, {}
    o = ao = 0
    for i, (agent_id, num) in enumerate(codebook.items()):
        d = dims[i]
        block = np.empty((B, d + 1), dtype=obs.dtype)
        block[:, 0] = num
        block[:, 1:] = obs[:, o:o + d]
        idx_state_all[agent_id] = torch.from_numpy(block)
        action_all[agent_id] = torch.from_numpy(np.ascontiguousarray(act[:, ao:ao + adims[i]]))
        o += d; ao += adims[i]
    joint = np.concatenate((nxt, rew), axis=1)
    return idx_state_all, action_all, torch.from_numpy(joint), torch.from_numpy(nxt), torch.from_numpy(rew)


class HostStager:
    """Pinned-host staging ring: a sampled transition dict is packed once into page-locked memory and shipped
    with ONE asynchronous H2D copy (the reference issues 44 small ``.to(device)`` copies per step,
    model.py:140,146 and :20-23).  ``depth`` slots let packing of batch k+1 overlap the copy of batch k.

    ``obs_dtype=torch.bfloat16`` keeps the observation block in bf16 on the host (rows ``[obs bf16 | act | next | rew fp32]``,
    25 % fewer PCIe bytes): the bf16 engine rounds observations on arrival anyway, so the train step is bit-identical to
    shipping fp32 (``tests/test_gpu_frontends.py::test_bf16_observation_feed_is_bit_identical``); this is the row format
    ``bench.py``'s ``e2e`` figure is measured on.  Reconstruction targets always travel in fp32."""

    def __init__(self, device, depth: int = 2, obs_dtype=torch.float32):
        if obs_dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("HostStager: obs_dtype must be torch.float32 or torch.bfloat16")
        self.device = torch.device(device)
        self.depth = depth
        self.obs_dtype = obs_dtype
        self._host = [None] * depth
        self._dev = [None] * depth
        self._ev = [None] * depth
        self._i = 0
        self.h2d_bytes = 0

    def stage(self, transition, codebook, sample0=0, batch_global=None) -> PackedBatch:
        agents, B, _, S, A = _layout(transition, codebook)
        W = int(sum(_act_dims(transition, agents)))
        k = self._i
        self._i = (k + 1) % self.depth
        cuda = self.device.type == "cuda"
        if self.obs_dtype == torch.float32:
            n = B * (2 * S + W + A)
            if self._host[k] is None or self._host[k].numel() != n:
                self._host[k] = torch.empty(n, dtype=torch.float32, pin_memory=cuda)
                self._dev[k] = torch.empty(n, dtype=torch.float32, device=self.device)
            elif self._ev[k] is not None:
                self._ev[k].synchronize()              # the previous copy out of this slot must have finished
            _pack_numpy(transition, codebook, flat=self._host[k].numpy())
            self._dev[k].copy_(self._host[k], non_blocking=True)
            self.h2d_bytes = n * 4
            obs, act, nxt, rew = _split_flat(self._dev[k], B, S, A, W)
        else:
            # byte buffer: [obs bf16 (B*S*2, padded to 16) | act, next, rew fp32]; fp32 rows are packed first, then the
            # observation block is rounded to bf16 in place on the host (one pass over B*S values)
            ob = (B * S * 2 + 15) // 16 * 16
            n32 = B * (S + W + A)
            total = ob + n32 * 4
            if self._host[k] is None or self._host[k].numel() != total:
                self._host[k] = torch.empty(total, dtype=torch.uint8, pin_memory=cuda)
                self._dev[k] = torch.empty(total, dtype=torch.uint8, device=self.device)
                self._scratch = torch.empty(B * (2 * S + W + A), dtype=torch.float32)
            elif self._ev[k] is not None:
                self._ev[k].synchronize()
            _pack_numpy(transition, codebook, flat=self._scratch.numpy())
            so, sa, sn, sr = _split_flat(self._scratch, B, S, A, W)
            h = self._host[k]
            h[:B * S * 2].view(torch.bfloat16).view(B, S).copy_(so)
            tail = h[ob:].view(torch.float32)
            tail[:B * W].view(B, W).copy_(sa)
            tail[B * W:B * (W + S)].view(B, S).copy_(sn)
            tail[B * (W + S):].view(B, A).copy_(sr)
            self._dev[k].copy_(h, non_blocking=True)
            self.h2d_bytes = total
            d = self._dev[k]
            dt = d[ob:].view(torch.float32)
            obs = d[:B * S * 2].view(torch.bfloat16).view(B, S)
            act, nxt, rew = dt[:B * W].view(B, W), dt[B * W:B * (W + S)].view(B, S), dt[B * (W + S):].view(B, A)
        if cuda:
            self._ev[k] = torch.cuda.Event()
            self._ev[k].record()
        return PackedBatch(obs, act, nxt, rew, sample0=sample0, batch_global=batch_global)


# ----------------------------------------------------------------------------------------------
# optimizer
# ----------------------------------------------------------------------------------------------
class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam(model.parameters(), lr)`` semantics (betas 0.9/0.999, eps 1e-8, no weight decay /
    amsgrad) executed as one kernel over the model's parameter arena.  LR schedulers that edit
    ``param_groups[0]['lr']`` (``CosineAnnealingLR`` at reference main.py:53) work unchanged."""

    def __init__(self, model: MAVAE, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self._model = model
        super().__init__(list(model.parameters()), dict(lr=lr, betas=betas, eps=eps))

    def zero_grad(self, set_to_none: bool = True):
        """Gradients live in the arena and are re-zeroed by the backward pass itself (its two largest wgrads overwrite), so
        nothing is cleared here; the call marks the last backward's gradients as consumed.  Two backward passes without a
        ``step()`` / ``zero_grad()`` in between raise instead of silently dropping the first (the engine does not
        accumulate across backward passes, where ``torch.optim`` users might expect it to)."""
        self._model.grads_consumed()
        return None

    def state_dict(self):
        """torch.optim.Adam-shaped: exp_avg / exp_avg_sq are the flat Adam arenas of the optimised prefix, `step` the
        shared step count -- enough to resume (the reference never saves optimizer state, main.py:111-112)."""
        m = self._model
        n = m._n_opt
        return {"state": {"exp_avg": m._m[:n].detach().clone(), "exp_avg_sq": m._v[:n].detach().clone(), "step": int(m._adam_t)},
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    @torch.no_grad()
    def load_state_dict(self, sd):
        m = self._model
        n = m._n_opt
        m._m[:n].copy_(sd["state"]["exp_avg"]); m._v[:n].copy_(sd["state"]["exp_avg_sq"])
        m._adam_t = int(sd["state"]["step"])
        for g, src in zip(self.param_groups, sd["param_groups"]):
            g.update(src)

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        self._model.adam_step(g["lr"], g["betas"], g["eps"])
        return None


# ----------------------------------------------------------------------------------------------
# Trainer
# ----------------------------------------------------------------------------------------------
class Trainer:
    def __init__(self, mode, model, lr, loss_func, beta=None, device='cuda:0'):
        self.mode = mode
        assert self.mode in ['Adam', 'ART', 'POPART']
        self.model = model
        self.sigma = torch.tensor(1., dtype=torch.float).to(device)
        self.sigma_new = None
        self.mu = torch.tensor(0., dtype=torch.float).to(device)
        self.mu_new = None
        self.nu = self.sigma ** 2 + self.mu ** 2
        self.beta = beta
        self.lr = lr
        self.loss_func = loss_func
        self.loss = None
        self.opt = FusedAdam(model, lr) if isinstance(model, MAVAE) else torch.optim.Adam(model.parameters(), lr)
        self.device = device

    # --- reward normalisation (ART / POP-ART) ---
    def art(self, y):
        y = y.to(self.device)
        if self.mode == 'POPART':
            y_mean, y_sq = y.mean(dim=0), (y ** 2).mean(dim=0)       # per-agent statistics, shape [A]
        else:
            y_mean, y_sq = y, y ** 2                                  # reference literal (trainer.py:67-68)
        self.mu_new = (1. - self.beta) * self.mu + self.beta * y_mean
        self.nu = (1. - self.beta) * self.nu + self.beta * y_sq
        self.sigma_new = torch.sqrt(self.nu - self.mu_new ** 2)

    def pop(self):
        relative_sigma = (self.sigma / self.sigma_new)
        lin = self.model.reward_linear
        with torch.no_grad():
            lin.weight.mul_(relative_sigma.reshape(-1, 1) if relative_sigma.dim() == 1 else relative_sigma)
            lin.bias.mul_(relative_sigma).add_((self.mu - self.mu_new) / self.sigma_new)

    def update_stats(self):
        if self.sigma_new is not None:
            self.sigma = self.sigma_new
        if self.mu_new is not None:
            self.mu = self.mu_new

    def normalize(self, y):
        return (y.to(self.device) - self.mu) / self.sigma

    def denormalize(self, y):
        return self.sigma * y.to(self.device) + self.mu

    # --- the step ---
    def forward(self, idx_state, actions, s_hat, r_hat):
        if self.mode in ['POPART', 'ART']:
            self.art(r_hat)
        if self.mode in ['POPART']:
            self.pop()
        self.update_stats()
        recon_s, recon_r, mean_all, logvar_all = self.model(idx_state, actions)
        self.loss, _, _, _ = self.loss_func(recon_s, recon_r, s_hat, self.normalize(r_hat), mean_all, logvar_all, self.device)
        return recon_s, recon_r, mean_all, logvar_all

    def backward(self):
        self.opt.zero_grad()
        self.loss.backward()

    def step(self):
        self.opt.step()

    def training_model(self, batched_replay_buffer, train_num, agent_id_codebook):
        loss_train = 0.0
        pbar = tqdm(range(train_num), desc="Training vae step", leave=False)
        for train_step_i in pbar:
            transitions = batched_replay_buffer.sample()
            idx_state, actions, next_state_rew, next_state, rewards = create_dataset(transitions, agent_id_codebook)
            self.forward(idx_state, actions, next_state, rewards)
            loss_train = loss_train + self.loss.detach()
            self.backward()
            self.step()
        return loss_train / (train_num * 1.0)


def cosine_lr(step: int, base_lr: float = 0.005, t_max: int = 50, eta_min: float = 1e-4) -> float:
    """lr that ``CosineAnnealingLR(T_max=50, eta_min=1e-4)`` (reference main.py:53) applies at optimizer step
    ``step`` (0-based), in closed form — what callers of ``MAVAE.train_step(pb, lr)`` pass when no scheduler object runs."""
    import math
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * step / t_max)) / 2.0
