"""Mint golden vectors by running the UNMODIFIED reference (``/root/reference/torch_ver``).

Run once in the build container (the reference is not present on the GPU box):

    python tests/golden/make_golden.py

For every case it
  * builds the reference ``MAVAE`` on CPU and overwrites every weight (registered and the
    unregistered per-agent encoders / action tables) with ``oracle.init_params(spec, seed)``,
  * stages a synthetic ``cpprb.sample``-shaped batch through the reference ``create_dataset``,
  * injects eps by replacing the module-level ``model.reparameterize`` (looked up as a global at
    ``torch_ver/model.py:151``) with a closure that consumes a fixed eps tensor per agent,
  * runs 3 train steps exactly as ``torch_ver/main.py:84-98`` does (Adam + CosineAnnealingLR),
  * records losses, and compact digests (sum, L2 norm, 24 sampled entries) of the step-1 outputs,
    step-1 gradients of every tensor, and the step-3 parameters.

The digests are written to ``tests/golden/<case>.npz`` (a few hundred KB in total).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/torch_ver")

from oracle import mavae_oracle as O  # noqa: E402

import model as ref_model      # noqa: E402  (reference)
import trainer as ref_trainer  # noqa: E402  (reference)

N_SAMPLES = 24


def digest(t: torch.Tensor, name: str) -> np.ndarray:
    """[sum, l2, sample_0..sample_23] in float64; sample positions derive from the tensor name."""
    x = t.detach().double().reshape(-1).numpy()
    h = np.frombuffer(name.encode(), dtype=np.uint8).astype(np.uint64).sum()
    rng = np.random.default_rng(int(h) + x.size)
    idx = rng.integers(0, x.size, size=N_SAMPLES)
    return np.concatenate([[x.sum(), np.sqrt((x * x).sum())], x[idx]])


def ref_named_tensors(m):
    """Every tensor of the reference model under the oracle's naming."""
    out = dict(m.named_parameters())
    for a, enc in m.encoders.items():
        for n, p in enc.named_parameters():
            out[f"encoders.{a}.{n}"] = p
    for a, emb in m.action_encoder.items():
        for n, p in emb.named_parameters():            # Embedding: "weight"; ActionEncoder: "net.0.weight", ...
            out[f"action_encoder.{a}.{n}"] = p
    return out


CASES = {
    # name: (spec, batch, param_seed, data_seed, reward_scale, huber)
    "tiny": (O.tiny_spec(3, idx_features=16, latent=8, act_features=8, include_dead_decoder=True), 16, 1, 2, 1.0, True),
    "tiny_mse": (O.tiny_spec(3, idx_features=16, latent=8, act_features=8, include_dead_decoder=True), 16, 3, 4, 1.0, False),
    "latent32": (O.tiny_spec(4, idx_features=64, latent=32, act_features=64, include_dead_decoder=True), 64, 5, 6, 10.0, True),
    "default": (O.simple_tag_spec(include_dead_decoder=True), 128, 0, 0, 1.0, True),
    # continuous actions: the ActionEncoder MLP replaces the action embedding (model.py:60-74,123,148)
    "continuous": (O.tiny_spec(4, idx_features=64, latent=32, act_features=64, include_dead_decoder=True, discrete_act=False,
                               act_dim={"adversary_0": 5, "adversary_1": 5, "adversary_2": 5, "agent_0": 3}), 64, 7, 8, 1.0, True),
}


def run_case(name, spec, B, pseed, dseed, rscale, huber):
    torch.manual_seed(0)
    m = ref_model.MAVAE(spec.idx_features, spec.latent, spec.act_features, spec.discrete_act, spec.agents,
                        spec.obs_dim, spec.n_act if spec.discrete_act else spec.act_dim, "cpu")
    P = O.init_params(spec, pseed)
    tensors = ref_named_tensors(m)
    assert set(tensors) == set(P), (set(tensors) ^ set(P))
    with torch.no_grad():
        for k, p in tensors.items():
            assert p.shape == P[k].shape, (k, p.shape, P[k].shape)
            p.copy_(P[k])

    codebook = {a: i for i, a in enumerate(spec.agents)}
    opt = torch.optim.Adam(m.parameters(), 0.005)                                   # main.py:52
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=50, eta_min=1e-4)  # main.py:53

    rec = {"batch": B, "param_seed": pseed, "data_seed": dseed, "reward_scale": rscale, "huber": int(huber)}
    losses, lrs = [], []
    L = spec.latent
    for step in range(3):
        trans = O.synth_transition(spec, B, seed=dseed + 1000 * step, reward_scale=rscale)
        idx_state, acts, joint, nxt, rew = ref_trainer.create_dataset(trans, codebook)
        eps_all = torch.from_numpy(O.philox_normal(0x5EED, step, 0, B, spec.n_agents * L).astype(np.float32))
        it = iter(range(spec.n_agents))

        def fixed_reparam(mu, log_var, _it=it, _eps=eps_all):
            a = next(_it)
            return mu + _eps[:, a * L:(a + 1) * L] * torch.exp(0.5 * log_var)

        ref_model.reparameterize = fixed_reparam
        with contextlib.redirect_stdout(io.StringIO()):      # debug prints at model.py:160-163
            rs, rr, mus, lvs = m(idx_state, acts)
        loss, sl, rl, kl = ref_model.loss_s_r_vae_fn(rs, rr, nxt, rew, mus, lvs, "cpu", using_huber_loss=huber)
        if step == 0:
            jl = ref_model.loss_vae_fn(torch.cat([rs, rr], 1), joint, mus, lvs, "cpu")
            rec["joint_mse_loss"] = float(jl.detach())
            rec["out.recon_s"] = digest(rs, "out.recon_s")
            rec["out.recon_r"] = digest(rr, "out.recon_r")
            rec["out.mu"] = digest(torch.cat(mus, 1), "out.mu")
            rec["out.logvar"] = digest(torch.cat(lvs, 1), "out.logvar")
        # unregistered tensors are never zeroed by the reference; zero them here so
        # every step's digest is a single-step gradient.
        for k, p in tensors.items():
            p.grad = None
        opt.zero_grad()
        loss.backward()
        if step == 0:
            for k, p in tensors.items():
                if p.grad is not None:
                    rec["grad." + k] = digest(p.grad, "grad." + k)
                else:
                    rec["nograd." + k] = np.zeros(1)
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
        losses.append([float(loss), float(sl), float(rl), float(kl)])
    rec["losses"] = np.array(losses, dtype=np.float64)
    rec["lrs"] = np.array(lrs, dtype=np.float64)
    for k, p in tensors.items():
        rec["param3." + k] = digest(p, "param3." + k)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "losses", losses, "lrs", lrs)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for name, args in CASES.items():
        if only and name not in only:
            continue
        run_case(name, *args)
    if only:
        sys.exit(0)
    # scheduler probe: lr the reference would use at selected steps
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], 0.005)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=50, eta_min=1e-4)
    lr = []
    for s in range(160):
        lr.append(opt.param_groups[0]["lr"])
        opt.step(); sched.step()
    np.savez_compressed(os.path.join(HERE, "cosine_lr.npz"), lr=np.array(lr))
