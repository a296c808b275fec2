"""Shared helpers: replay a golden case (tests/golden/*.npz, minted from the reference) through
any implementation and compare digests."""
import os

import numpy as np
import torch

from oracle import mavae_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_SAMPLES = 24

SPECS = {
    "tiny": lambda: O.tiny_spec(3, idx_features=16, latent=8, act_features=8, include_dead_decoder=True),
    "tiny_mse": lambda: O.tiny_spec(3, idx_features=16, latent=8, act_features=8, include_dead_decoder=True),
    "latent32": lambda: O.tiny_spec(4, idx_features=64, latent=32, act_features=64, include_dead_decoder=True),
    "default": lambda: O.simple_tag_spec(include_dead_decoder=True),
    "continuous": lambda: O.tiny_spec(4, idx_features=64, latent=32, act_features=64, include_dead_decoder=True, discrete_act=False,
                                      act_dim={"adversary_0": 5, "adversary_1": 5, "adversary_2": 5, "agent_0": 3}),
}


# Cases without golden vectors (too large to mint digests for, or shapes the reference's hard-coded widths cannot
# express): replayed against the oracle only.  name -> (spec factory, batch, param_seed, data_seed, reward_scale, huber)
SYNTH = {
    # the exact headline shape of BASELINE.json configs[1] / bench.py default
    "cfg2_b4096": (lambda: O.simple_tag_spec(latent=32), 4096, 0, 0, 3.0, True),
    # BASELINE.json configs[2] "wide": hidden 1024 x 4, latent 128 (few agents, small batch: the layer shapes are what matter)
    "wide": (lambda: O.tiny_spec(4, idx_features=64, latent=128, act_features=64, enc_hidden=(1024,) * 4, dec_hidden=(1024,) * 4),
             256, 11, 12, 3.0, True),
}


def load_case(name):
    if name in SYNTH:
        f, B, ps, ds, rs, hub = SYNTH[name]
        return f(), {"batch": B, "param_seed": ps, "data_seed": ds, "reward_scale": rs, "huber": int(hub)}
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return SPECS[name](), {k: z[k] for k in z.files}


def digest(t, name):
    x = torch.as_tensor(t).detach().double().cpu().reshape(-1).numpy()
    h = np.frombuffer(name.encode(), dtype=np.uint8).astype(np.uint64).sum()
    rng = np.random.default_rng(int(h) + x.size)
    idx = rng.integers(0, x.size, size=N_SAMPLES)
    return np.concatenate([[x.sum(), np.sqrt((x * x).sum())], x[idx]])


def digest_close(got, want, rtol, what):
    """Compare [sum, l2, samples...]: l2 relative, sum and samples relative to the tensor scale
    (l2 / sqrt(n) is unknown here, so scale by max |sample| and l2)."""
    l2 = max(abs(want[1]), 1e-30)
    assert abs(got[1] - want[1]) <= rtol * l2, f"{what}: l2 {got[1]} vs {want[1]}"
    scale = max(np.abs(want[2:]).max(), 1e-30)
    err = np.abs(got[2:] - want[2:]).max()
    assert err <= rtol * scale * 4, f"{what}: samples max err {err} scale {scale}"
    assert abs(got[0] - want[0]) <= rtol * max(abs(want[0]), l2) * 8, f"{what}: sum {got[0]} vs {want[0]}"


def step_inputs(spec, rec, step):
    B = int(rec["batch"])
    trans = O.synth_transition(spec, B, seed=int(rec["data_seed"]) + 1000 * step,
                               reward_scale=float(rec["reward_scale"]))
    codebook = {a: i for i, a in enumerate(spec.agents)}
    eps_all = torch.from_numpy(O.philox_normal(0x5EED, step, 0, B, spec.n_agents * spec.latent).astype(np.float32))
    return trans, codebook, eps_all
