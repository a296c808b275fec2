"""-m gpu: the callers either side of the hot path: evaluation step, the jax_ver ELBO weighting and its
train_step / test_step pair, and the replay ring feeding train steps."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(precision="fp32", **kw):
    import mfvae_b200 as M
    from oracle import mavae_oracle as O
    spec = O.tiny_spec(4, idx_features=64, latent=32, act_features=64)
    m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, "cuda:0",
                precision=precision, include_dead_decoder=False, **kw)
    m.load_named(O.init_params(spec, 11))
    return M, O, spec, m


def _batch(M, spec, B, seed=3):
    g = torch.Generator(device="cuda:0").manual_seed(seed)
    S, A = spec.state_dim, spec.n_agents
    return M.PackedBatch(torch.randn(B, S, device="cuda:0", generator=g),
                         torch.randint(0, 5, (B, A), device="cuda:0", generator=g).float(),
                         torch.randn(B, S, device="cuda:0", generator=g), torch.randn(B, A, device="cuda:0", generator=g) * 4)


def test_test_step_equals_train_step_losses_and_leaves_state_alone():
    M, O, spec, m = _model()
    pb = _batch(M, spec, 96)
    before = m._arena.clone()
    ev = m.test_step(pb).clone()
    assert torch.equal(m._arena, before) and m.philox_step == 0           # no update, Philox step not advanced
    tr = m.train_step(pb, 0.0).clone()                                    # same Philox step -> same eps
    assert torch.allclose(ev, tr, rtol=1e-6, atol=0)
    # against the oracle on the kernel's own eps stream
    eps_all = torch.from_numpy(O.philox_normal(m.philox_seed, 0, 0, 96, spec.n_agents * spec.latent).astype(np.float32))
    L = spec.latent
    idx_state = {a: torch.cat([torch.full((96, 1), float(i)), pb.obs[:, sum(spec.obs_dim[x] for x in spec.agents[:i]):][:, :spec.obs_dim[a]].cpu()], 1)
                 for i, a in enumerate(spec.agents)}
    acts = {a: pb.act[:, i:i + 1].cpu() for i, a in enumerate(spec.agents)}
    eps = {a: eps_all[:, i * L:(i + 1) * L] for i, a in enumerate(spec.agents)}
    want, _, _ = O.grads(O.init_params(spec, 11), spec, idx_state, acts, eps, pb.next.cpu(), pb.rew.cpu())
    for g, w in zip(ev.cpu().tolist(), want):
        assert abs(g - w) <= 1e-5 * abs(w)


def test_jax_weighting_is_the_weighted_sum_of_the_three_terms():
    """jax_ver/trainer.py:42-43,64: loss = 0.5 s + 0.5 r + 0.1 kl.  Gradients are linear in the three weights, so the jax
    step's gradients must equal the same combination of the single-term gradients (fp32 engine, 1e-5)."""
    M, O, spec, m = _model()
    from mfvae_b200 import jax_trainer as J
    pb = _batch(M, spec, 64)

    def grads(weights):
        m.philox_step = 0
        losses = m.train_step(pb, 0.0, loss_weights=weights).clone()
        return losses, m._grad.clone()

    l_s, g_s = grads((0.0, 0.0, 1.0))
    l_r, g_r = grads((0.0, 1.0, 0.0))
    l_k, g_k = grads((1.0, 0.0, 0.0))
    m.philox_step = 0
    l_j = J.train_step(m, pb, lr=0.0).clone()
    g_j = m._grad.clone()
    kw, rw, sw = J.loss_weights()
    assert (kw, rw, sw) == (0.1, 0.5, 0.5)
    want = sw * g_s + rw * g_r + kw * g_k
    assert float((g_j - want).norm() / want.norm()) < 1e-5
    assert abs(float(l_j[0]) - (sw * float(l_j[1]) + rw * float(l_j[2]) + kw * float(l_j[3]))) < 1e-6
    assert torch.allclose(l_j[1:], l_s[1:], rtol=1e-6)                   # the three raw terms do not depend on the weights
    m.philox_step = 0
    assert torch.allclose(J.test_step(m, pb), l_j, rtol=1e-6)


def test_ring_fed_train_steps_match_host_fed():
    """The replay ring (reference MultiAgentCPPRB / JaxFbxBuffer replacement) hands the step the same rows the host
    path would: sample indices -> gather on the device == indexing the host copy of the stored rows."""
    M, O, spec, m = _model()
    ring = M.DeviceRing(spec.agents, spec.obs_dim, capacity=300, device="cuda:0")
    g = torch.Generator().manual_seed(9)
    rows = torch.randn(300, ring.row, generator=g)
    S, A = ring.S, ring.A
    rows[:, S:S + A] = torch.randint(0, 5, (300, A), generator=g).float()
    ring.add_rows_device(rows.cuda())
    pb, idx = ring.sample_packed(128, seed=5, with_indices=True)
    take = rows[idx.cpu().long()]
    assert torch.equal(pb.obs.cpu(), take[:, :S]) and torch.equal(pb.act.cpu(), take[:, S:S + A])
    assert torch.equal(pb.next.cpu(), take[:, S + A:2 * S + A]) and torch.equal(pb.rew.cpu(), take[:, 2 * S + A:2 * S + 2 * A])
    a = m.test_step(pb).clone()
    host = M.PackedBatch(take[:, :S].cuda(), take[:, S:S + A].cuda(), take[:, S + A:2 * S + A].cuda(), take[:, 2 * S + A:2 * S + 2 * A].cuda())
    b = m.test_step(host).clone()
    assert torch.equal(a, b)


def test_full_size_shard_additivity_and_determinism():
    """Size-independent properties at the benchmark's full size (cfg2: 40 agents, B = 4096, bf16 / tcgen05): the
    gradients of two half-batch shards (scaled by the global batch, Philox keyed by the global sample index) add up to
    the full-batch gradients, and the loss scalars are bit-reproducible run to run."""
    import mfvae_b200 as M
    from oracle import mavae_oracle as O
    spec = O.simple_tag_spec(latent=32)
    torch.manual_seed(3)
    m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, "cuda:0",
                precision="bf16", include_dead_decoder=False)
    B = 4096
    pb = _batch(M, spec, B, seed=21)

    def run(lo, hi):
        m.philox_step = 0
        part = M.PackedBatch(pb.obs[lo:hi].contiguous(), pb.act[lo:hi].contiguous(), pb.next[lo:hi].contiguous(),
                             pb.rew[lo:hi].contiguous(), sample0=lo, batch_global=B)
        losses = m.train_step(part, 0.0).clone()
        return losses, m._grad.clone()

    l_full, g_full = run(0, B)
    l_again, _ = run(0, B)
    assert torch.equal(l_full, l_again)                       # fixed-order reductions: bit-identical scalars
    l0, g0 = run(0, B // 2)
    l1, g1 = run(B // 2, B)
    assert torch.allclose(l0 + l1, l_full, rtol=2e-5)
    n = m._n_opt
    # same per-sample bf16 rounding points in both schedules: only fp32 accumulation order differs
    for lo, hi, what in ((0, n, "optimised prefix"), (n, g_full.numel(), "encoders / action tables")):
        d = float((g0[lo:hi] + g1[lo:hi] - g_full[lo:hi]).norm() / g_full[lo:hi].norm())
        assert d < 2e-4, (what, d)


def test_reference_driver_loop_end_to_end():
    """The reference's training loop (torch_ver/main.py:62-102) with every piece swapped for its drop-in: transitions are
    added to MultiAgentCPPRB (device ring), sampled, staged by create_dataset, and trained both through the inline call
    sequence (main.py:84-98) and through Trainer.training_model (trainer.py:105-119); the loss must fall."""
    import mfvae_b200 as M
    from oracle import mavae_oracle as O
    spec = O.tiny_spec(4, idx_features=64, latent=32, act_features=64)
    rng = np.random.default_rng(0)
    buf = M.MultiAgentCPPRB(max_size=512, batch_size=128, agents=spec.agents, obs_dim=spec.obs_dim, device="cuda:0", seed=3)
    W = {a: rng.standard_normal((spec.obs_dim[a], spec.obs_dim[a])).astype(np.float32) * 0.3 for a in spec.agents}
    for t in range(300):        # a learnable toy dynamics: next_obs = tanh(W obs), reward = mean(obs)
        obs = {a: rng.standard_normal(spec.obs_dim[a]).astype(np.float32) for a in spec.agents}
        nxt = {a: np.tanh(W[a] @ obs[a]) for a in spec.agents}
        act = {a: rng.integers(0, 5) for a in spec.agents}
        rew = {a: float(obs[a].mean()) for a in spec.agents}
        flags = {a: False for a in spec.agents}
        buf.add(obs, nxt, act, rew, flags, flags)
    buf.on_episode_end()
    codebook = {a: i for i, a in enumerate(spec.agents)}
    model = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, "cuda:0")
    opt = M.FusedAdam(model, 0.005)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=50, eta_min=1e-4)
    hist = []
    for step in range(40):                                            # main.py:84-98, verbatim call sequence
        transitions = buf.sample()
        idx_state, actions, next_state_rew, next_state, rewards = M.create_dataset(transitions, codebook)
        recon_state, recon_reward, mu_all, logvar_all = model(idx_state, actions)
        loss, s_loss, r_loss, kl_loss = M.loss_s_r_vae_fn(recon_state, recon_reward, next_state, rewards, mu_all, logvar_all, "cuda:0")
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        hist.append(float(loss))
    assert all(np.isfinite(hist)) and np.mean(hist[-5:]) < 0.95 * np.mean(hist[:5]), hist      # 0.226 -> 0.205 measured
    tr = M.Trainer("Adam", model, 0.002, M.loss_s_r_vae_fn, device="cuda:0")
    mean_loss = tr.training_model(buf, 5, codebook)                   # raises TypeError in the reference (trainer.py:112)
    assert np.isfinite(float(mean_loss)) and float(mean_loss) < np.mean(hist[:5])


@pytest.mark.parametrize("fusion", ["none", "encoder+loss"])
@pytest.mark.parametrize("B", [77, 300])
def test_workspace_tail_canary_ragged_batches(fusion, B):
    """Own bounds check (compute-sanitizer is not available on this pool): ragged batch sizes (not multiples of the
    128-row tiles, fewer rows than one tile) must not write past the activation workspace.  The workspace is re-bound with a
    64 KB sentinel tail that has to survive forward, loss, backward and Adam."""
    import ctypes as C
    import mfvae_b200 as M
    from mfvae_b200 import _lib as L
    from oracle import mavae_oracle as O
    spec = O.simple_tag_spec(latent=32)
    m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, "cuda:0",
                precision="bf16", include_dead_decoder=False, fusion=fusion)
    need = L.lib().mfvae_workspace_bytes(m._h, B)
    tail = 1 << 16
    ws = torch.zeros(need + tail, dtype=torch.uint8, device="cuda:0")
    ws[need:] = 0xAB
    torch.cuda.synchronize()
    L.check(L.lib().mfvae_bind_workspace(m._h, L.ptr(ws), need, B))
    m._ws, m._ws_batch = ws, B
    pb = _batch(M, spec, B, seed=77)
    for _ in range(3):
        losses = m.train_step(pb, 1e-3)
    rs, rr, mus, lvs = m(M.PackedBatch(pb.obs, pb.act))
    torch.cuda.synchronize()
    assert bool((ws[need:] == 0xAB).all()), "write past the end of the workspace"
    assert bool(torch.isfinite(losses).all()) and bool(torch.isfinite(rs).all())


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """save_checkpoint -> load_checkpoint into a fresh model -> the next train steps (parameters, Adam moments,
    losses, Philox stream) are bit-identical to the uninterrupted run.  Loss scalars use fixed-order reductions; the
    parameters are compared after steps whose gradients carry fp32 atomics, so the model runs with a batch of one tile per
    agent and split-K off the critical tensors is not assumed: equality is asserted on the losses and within 1e-6 on parameters."""
    M, O, spec, m = _model("bf16")
    pbs = [_batch(M, spec, 128, seed=40 + i) for i in range(4)]
    for i in range(2):
        m.train_step(pbs[i], M.cosine_lr(i))
    p = str(tmp_path / "ck.pt")
    m.save_checkpoint(p)
    cont = [m.train_step(pbs[2 + i], M.cosine_lr(2 + i)).clone() for i in range(2)]
    _, _, _, m2 = _model("bf16")
    with torch.no_grad():
        for t in m2.named_arena_tensors().values():            # make sure everything really comes from the file
            t.normal_()
    m2.load_checkpoint(p)
    assert m2._adam_t == 2 and m2.philox_step == 2
    res = [m2.train_step(pbs[2 + i], M.cosine_lr(2 + i)).clone() for i in range(2)]
    torch.cuda.synchronize()
    for a, b in zip(cont, res):
        assert torch.allclose(a, b, rtol=1e-6, atol=0), (a, b)
    n = m._n_opt
    assert float((m._arena[:n] - m2._arena[:n]).norm() / m._arena[:n].norm()) < 1e-6
    assert float((m._m[:n] - m2._m[:n]).norm() / m._m[:n].norm()) < 1e-5


def test_data_edits_of_reward_linear_reach_the_tensor_cores():
    """the reference's POP-ART idiom edits reward_linear through `.data` (torch_ver/trainer.py:73-74), which does
    not bump the arena's version counter.  The drop-in forward must still see the edit in bf16 mode (stale bf16 shadow =
    silently wrong forward); mark_dirty() covers edits of any other tensor."""
    M, O, spec, m = _model("bf16")
    pb = _batch(M, spec, 64)
    with torch.no_grad():
        m.philox_step = 0
        _, rr0, _, _ = m(M.PackedBatch(pb.obs, pb.act))
        rr0 = rr0.clone()
        m.reward_linear.weight.data.mul_(2.0)                  # invisible to autograd's version counter
        m.reward_linear.bias.data.add_(1.0)
        m.philox_step = 0
        _, rr1, _, _ = m(M.PackedBatch(pb.obs, pb.act))
        assert torch.allclose(rr1, 2.0 * rr0 + 1.0, rtol=2e-2, atol=2e-2), float((rr1 - (2 * rr0 + 1)).abs().max())
        # any other tensor: .data edit + mark_dirty()
        m.philox_step = 0
        rs0 = m(M.PackedBatch(pb.obs, pb.act))[0].clone()
        m.state_decoder.net[10].bias.data.add_(3.0)            # bias is read in fp32: visible at once
        m.state_decoder.net[10].weight.data.mul_(0.0)
        m.mark_dirty()
        m.philox_step = 0
        rs1 = m(M.PackedBatch(pb.obs, pb.act))[0]
        want = m.state_decoder.net[10].bias.detach().expand_as(rs1)
        assert torch.allclose(rs1, want, atol=1e-5) and not torch.allclose(rs0, rs1)


def test_second_backward_without_step_raises_on_gpu():
    M, O, spec, m = _model()
    pb = _batch(M, spec, 32)
    opt = M.FusedAdam(m, 1e-3)
    rs, rr, mus, lvs = m(M.PackedBatch(pb.obs, pb.act))
    loss = M.loss_s_r_vae_fn(rs, rr, pb.next, pb.rew, mus, lvs, "cuda:0")[0]
    loss.backward()
    rs, rr, mus, lvs = m(M.PackedBatch(pb.obs, pb.act))
    loss = M.loss_s_r_vae_fn(rs, rr, pb.next, pb.rew, mus, lvs, "cuda:0")[0]
    with pytest.raises(RuntimeError, match="second backward"):
        loss.backward()
    opt.zero_grad()
    rs, rr, mus, lvs = m(M.PackedBatch(pb.obs, pb.act))
    M.loss_s_r_vae_fn(rs, rr, pb.next, pb.rew, mus, lvs, "cuda:0")[0].backward()
    opt.step()


def test_bf16_observation_feed_is_bit_identical():
    """PackedBatch.obs in bfloat16 (host ring keeps observations in bf16: half the PCIe bytes) = the same step as feeding the
    fp32 values those bf16 numbers came from, bit for bit (bf16 engine), and within fp32 round-off on the fp32 engine.  bf16
    next-observation targets (opt-in) move the loss only at the bf16 rounding level."""
    for precision in ("bf16", "fp32"):
        M, O, spec, m = _model(precision)
        pb = _batch(M, spec, 200, seed=8)
        obs16 = pb.obs.to(torch.bfloat16)
        m.philox_step = 0
        a = m.train_step(M.PackedBatch(obs16.float(), pb.act, pb.next, pb.rew), 0.0).clone()
        ga = m._grad.clone()
        m.philox_step = 0
        b = m.train_step(M.PackedBatch(obs16, pb.act, pb.next, pb.rew), 0.0).clone()
        gb = m._grad.clone()
        assert torch.equal(a, b), (precision, a, b)
        assert float((ga - gb).norm() / ga.norm()) < 1e-6          # split-K atomics: not bit-reproducible run to run
        m.philox_step = 0
        c = m.train_step(M.PackedBatch(obs16, pb.act, pb.next.to(torch.bfloat16), pb.rew), 0.0)
        assert abs(float(c[1]) - float(a[1])) < 5e-3 * abs(float(a[1])) and float(c[2]) == float(a[2])


def test_single_call_train_step_matches_the_call_by_call_sequence():
    """SURVEY 8b `mfvae_train_step`: one C call = fwd + ELBO + bwd + Adam; identical to MAVAE.train_step's call-by-call sequence."""
    for precision in ("fp32", "bf16"):
        M, O, spec, a = _model(precision)
        _, _, _, b = _model(precision)
        pbs = [_batch(M, spec, 160, seed=60 + i) for i in range(3)]
        for i, pb in enumerate(pbs):
            la = a.train_step(pb, M.cosine_lr(i)).clone()
            lb = b.train_step_c(pb, M.cosine_lr(i)).clone()
            # step 1 is the same kernels on the same data: identical loss scalars (fixed-order reductions); later steps start from
            # parameters that differ in the last bit (split-K atomics are not bit-reproducible run to run)
            assert torch.equal(la, lb) if i == 0 else torch.allclose(la, lb, rtol=1e-5 if precision == "fp32" else 2e-3), (precision, i, la, lb)
        torch.cuda.synchronize()
        n = a._n_opt
        assert float((a._arena[:n] - b._arena[:n]).norm() / a._arena[:n].norm()) < (1e-6 if precision == "fp32" else 1e-3)
        assert a._adam_t == b._adam_t == 3 and a.philox_step == b.philox_step == 3
