"""Shape bookkeeping of one MAVAE instance that the host side needs without touching the CUDA library: the agent list
and integer dims of the reference's environment (``torch_ver/src/env.py:27``: simple_tag_v3 with 30 adversaries, 10
good agents, 20 obstacles) and the multiply-accumulate count of one train step (``SURVEY.md`` Appendix A)."""
from dataclasses import dataclass
from typing import Dict, List, Sequence

ENC_HIDDEN = (64, 64, 256)                 # Encoder.HIDDEN  torch_ver/model.py:46
DEC_HIDDEN = (1024, 256, 64, 256, 1024)    # Decoder.HIDDEN  torch_ver/model.py:87


@dataclass
class ModelDims:
    agents: List[str]
    obs_dim: Dict[str, int]
    n_act: Dict[str, int]
    idx_features: int = 64        # IDX_FEATURES  torch_ver/main.py:30
    latent: int = 64              # OBS_FEATURES  torch_ver/main.py:31
    act_features: int = 64        # ACT_FEATURES  torch_ver/main.py:32
    enc_hidden: Sequence[int] = ENC_HIDDEN
    dec_hidden: Sequence[int] = DEC_HIDDEN

    @property
    def n_agents(self) -> int:
        return len(self.agents)

    @property
    def state_dim(self) -> int:
        return sum(self.obs_dim[a] for a in self.agents)

    @property
    def dec_in(self) -> int:
        return (self.latent + self.act_features) * self.n_agents


def simple_tag_dims(n_adv: int = 30, n_good: int = 10, n_obst: int = 20, **kw) -> ModelDims:
    """Integer dims of PettingZoo ``simple_tag_v3`` as configured at ``torch_ver/src/env.py:27``: obs = 2 vel + 2 pos +
    2 per obstacle + 2 per other agent + 2 per other good agent's velocity; 5 discrete actions."""
    agents = [f"adversary_{i}" for i in range(n_adv)] + [f"agent_{i}" for i in range(n_good)]
    n = n_adv + n_good
    obs = {}
    for a in agents:
        base = 2 + 2 + 2 * n_obst + 2 * (n - 1)
        obs[a] = base + 2 * (n_good if a.startswith("adversary") else n_good - 1)
    return ModelDims(agents=agents, obs_dim=obs, n_act={a: 5 for a in agents}, **kw)


def macs_per_sample(d: ModelDims) -> int:
    """Multiply-accumulates of one forward pass through the reference's dense layers (``nn.Linear`` at model.py:50,53,
    91,94,130): 40 encoders, both decoders, ``reward_linear``.  The never-called ``decoder`` (model.py:127) is not counted."""
    macs = 0
    for a in d.agents:
        w = [d.idx_features + d.obs_dim[a], *d.enc_hidden, 2 * d.latent]
        macs += sum(w[i] * w[i + 1] for i in range(len(w) - 1))
    for out in (d.state_dim, d.n_agents):
        w = [d.dec_in, *d.dec_hidden, out]
        macs += sum(w[i] * w[i + 1] for i in range(len(w) - 1))
    return macs + d.n_agents ** 2


def flops_per_sample(d: ModelDims) -> int:
    """Conventional dense count of one train step: forward + dgrad + wgrad = 6 flops per MAC of the reference's layers."""
    return 6 * macs_per_sample(d)


def executed_macs_per_sample(d: ModelDims, n_act_max: int = 5) -> int:
    """MACs the CUDA path actually runs per sample once the two constant-input column blocks are folded away
    (``SURVEY.md`` Appendix A): the id-embedding columns of encoder layer 0 become a per-agent bias, and the
    action-embedding half of decoder layer 0 becomes ``n_act`` one-hot columns per agent."""
    macs = macs_per_sample(d)
    macs -= d.n_agents * d.idx_features * d.enc_hidden[0]
    macs -= 2 * d.dec_hidden[0] * d.n_agents * (d.act_features - n_act_max)
    return macs
