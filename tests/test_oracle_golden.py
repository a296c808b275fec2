"""CPU: the oracle (oracle/mavae_oracle.py) must reproduce the golden vectors minted from the
unmodified reference (tests/golden/make_golden.py).  fp32 oracle vs fp32 reference -> 2e-5."""
import numpy as np
import pytest
import torch

from oracle import mavae_oracle as O
from tests.golden_util import load_case, digest, digest_close, step_inputs

RTOL = 2e-5


@pytest.mark.parametrize("name", ["tiny", "tiny_mse", "latent32", "default", "continuous"])
def test_oracle_matches_reference_golden(name):
    spec, rec = load_case(name)
    L = spec.latent
    st = O.OracleState(spec, O.init_params(spec, int(rec["param_seed"])))
    huber = bool(rec["huber"])
    for step in range(3):
        trans, codebook, eps_all = step_inputs(spec, rec, step)
        idx_state, acts, joint, nxt, rew = O.stage_batch(trans, codebook)
        eps = {a: eps_all[:, i * L:(i + 1) * L] for i, a in enumerate(spec.agents)}
        lr = O.cosine_lr(step)
        assert abs(lr - rec["lrs"][step]) < 1e-12
        losses, G, outs = O.train_step(st, idx_state, acts, eps, nxt, rew, lr, huber)
        np.testing.assert_allclose(losses, rec["losses"][step], rtol=5e-5)
        if step == 0:
            rs, rr, mus, lvs = outs
            digest_close(digest(rs, "out.recon_s"), rec["out.recon_s"], RTOL, "recon_s")
            digest_close(digest(rr, "out.recon_r"), rec["out.recon_r"], RTOL, "recon_r")
            digest_close(digest(torch.cat(mus, 1), "out.mu"), rec["out.mu"], RTOL, "mu")
            digest_close(digest(torch.cat(lvs, 1), "out.logvar"), rec["out.logvar"], RTOL, "logvar")
            jl = O.loss_joint_mse(joint, torch.cat([rs, rr], 1), mus, lvs)
            assert abs(float(jl) - float(rec["joint_mse_loss"])) <= 1e-5 * abs(float(rec["joint_mse_loss"]))
            n_checked = 0
            for k in rec:
                if k.startswith("grad."):
                    digest_close(digest(G[k[5:]], k), rec[k], RTOL, k)
                    n_checked += 1
                elif k.startswith("nograd."):
                    assert k[7:] not in G, k          # dead ``decoder`` gets no gradient
            assert n_checked > 20
    for k in rec:
        if k.startswith("param3."):
            digest_close(digest(st.P[k[7:]], k), rec[k], 5e-5, k)


def test_cosine_lr_matches_reference_scheduler():
    import os
    from tests.golden_util import GOLDEN_DIR
    lr = np.load(os.path.join(GOLDEN_DIR, "cosine_lr.npz"))["lr"]
    for s in range(len(lr)):
        assert abs(O.cosine_lr(s) - lr[s]) < 1e-9, s


def test_stage_batch_layout():
    spec = O.tiny_spec(3)
    t = O.synth_transition(spec, 5, seed=7)
    cb = {a: i for i, a in enumerate(spec.agents)}
    idx_state, acts, joint, nxt, rew = O.stage_batch(t, cb)
    for a, i in cb.items():
        assert idx_state[a].shape == (5, 1 + spec.obs_dim[a])
        assert torch.all(idx_state[a][:, 0] == i)
        assert torch.equal(idx_state[a][:, 1:], torch.from_numpy(t[a + "_observations"]))
    assert nxt.shape == (5, spec.state_dim) and rew.shape == (5, 3) and joint.shape == (5, spec.state_dim + 3)


def test_philox_known_answer():
    # Random123 published KAT for philox4x32-10: counter = key = 0 and the all-ones vector
    z = O.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))[0]
    assert [int(x) for x in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    o = O.philox4x32_10(np.full((1, 4), 0xFFFFFFFF, np.uint32), np.full((1, 2), 0xFFFFFFFF, np.uint32))[0]
    assert [int(x) for x in o] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_philox_normal_moments():
    e = O.philox_normal(0x5EED, 3, 0, 4096, 64)
    assert abs(e.mean()) < 0.01 and abs(e.std() - 1.0) < 0.01
    assert abs((e ** 3).mean()) < 0.03 and abs((e ** 4).mean() - 3.0) < 0.1
    # counter-based: a shard starting at sample 1024 equals the slice of the global draw
    s = O.philox_normal(0x5EED, 3, 1024, 128, 64)
    assert np.array_equal(s, e[1024:1152])


def test_explicit_huber_formula_equals_library():
    g = torch.Generator().manual_seed(0)
    x = 3 * torch.randn(64, 37, generator=g, dtype=torch.float64); y = torch.randn(64, 37, generator=g, dtype=torch.float64)
    assert abs(float(O.huber_mean(x, y)) - float(torch.nn.functional.huber_loss(x, y))) < 1e-12
    v, gr = O.np_recon_loss(x.numpy(), y.numpy(), True, 0.5, x.numel())
    xr = x.clone().requires_grad_(True)
    (0.5 * torch.nn.functional.huber_loss(xr, y)).backward()
    assert abs(v - float(torch.nn.functional.huber_loss(x, y))) < 1e-12
    assert float((xr.grad - torch.from_numpy(gr)).abs().max()) < 1e-15
