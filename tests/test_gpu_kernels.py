"""-m gpu: every CUDA kernel against the oracle's closed forms (oracle/mavae_oracle.py), through the C ABI."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import mavae_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mfvae_b200 import _lib as L
    return L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def test_philox_normal_matches_oracle(lib):
    B, W = 257, 96
    out = torch.empty(B, W, device="cuda")
    lib.check(lib.lib().mfvae_philox_normal(lib.ptr(out), B, W, 0x5EED, 7, 1000, _stream()))
    want = O.philox_normal(0x5EED, 7, 1000, B, W)
    # fp32 logf / sincospif vs float64: tolerance 2e-6 absolute (values are O(1), |eps| < 6)
    assert float(np.abs(out.cpu().numpy() - want).max()) < 2e-6


@pytest.mark.parametrize("B", [1, 33, 1024, 16384])
@pytest.mark.parametrize("zdt", [0, 1])
def test_reparam_kl_matches_oracle(lib, B, zdt):
    A, Lt = 5, 32
    W = A * Lt
    g = torch.Generator(device="cuda").manual_seed(B)
    mu = torch.randn(B, W, device="cuda", generator=g)
    lv = 0.5 * torch.randn(B, W, device="cuda", generator=g)
    z = torch.empty(B, W, device="cuda", dtype=torch.bfloat16 if zdt else torch.float32)
    kl = torch.zeros(1, device="cuda"); scratch = torch.zeros(4096, device="cuda")
    for _ in range(2):   # twice: the in-kernel ticket must re-arm itself
        lib.check(lib.lib().mfvae_reparam_kl(lib.ptr(mu), lib.ptr(lv), None, lib.ptr(z), zdt, B, W, 0x5EED, 3, 64, 2 * B,
                                             lib.ptr(kl), lib.ptr(scratch), _stream()))
    eps = O.philox_normal(0x5EED, 3, 64, B, W)
    zw, klw = O.np_reparam_kl(mu.cpu().numpy(), lv.cpu().numpy(), eps, Lt)
    tol = 1e-2 if zdt else 2e-5
    assert float(np.abs(z.float().cpu().numpy() - zw).max()) <= tol * max(1.0, float(np.abs(zw).max()))
    assert abs(float(kl) - klw / 2) <= 1e-5 * abs(klw / 2) + 1e-7      # batch_global = 2B


@pytest.mark.parametrize("huber", [1, 0])
@pytest.mark.parametrize("B,W", [(7, 3), (64, 40), (513, 5660), (2, 8)])
def test_recon_loss_matches_oracle(lib, huber, B, W):
    g = torch.Generator(device="cuda").manual_seed(W)
    ld = (W + 7) // 8 * 8
    recon = (2.0 * torch.randn(B, ld, device="cuda", generator=g))
    target = torch.randn(B, W, device="cuda", generator=g).contiguous()
    for gdt in (0, 1):
        grad = torch.zeros(B, ld, device="cuda", dtype=torch.bfloat16 if gdt else torch.float32)
        loss = torch.zeros(1, device="cuda"); scratch = torch.zeros(4096, device="cuda")
        lib.check(lib.lib().mfvae_recon_loss(lib.ptr(recon), ld, lib.ptr(target), W, lib.ptr(grad), ld, gdt, B, W, huber, 0.005,
                                             3 * B * W, lib.ptr(loss), lib.ptr(scratch), _stream()))
        val, gw = O.np_recon_loss(recon[:, :W].cpu().numpy(), target.cpu().numpy(), bool(huber), 0.005, 3 * B * W)
        assert abs(float(loss) - val) <= 2e-6 * abs(val) + 1e-9
        got = grad[:, :W].float().cpu().numpy()
        tol = 8e-3 if gdt else 1e-6
        assert float(np.abs(got - gw).max()) <= tol * float(np.abs(gw).max())


def test_adam_flat_matches_oracle(lib):
    n = 4096 * 3 + 8
    g = torch.Generator(device="cuda").manual_seed(1)
    p = torch.randn(n, device="cuda", generator=g); p0 = p.clone()
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    shadow = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    pn, mn, vn = p0.cpu().numpy(), np.zeros(n), np.zeros(n)
    tref = torch.nn.Parameter(p0.clone()); topt = torch.optim.Adam([tref], 5e-3)
    for t in range(1, 4):
        gr = torch.randn(n, device="cuda", generator=g) * (10.0 ** (t - 2))
        lib.check(lib.lib().mfvae_adam_flat(lib.ptr(p), lib.ptr(gr), lib.ptr(m), lib.ptr(v), lib.ptr(shadow), n,
                                            5e-3, 0.9, 0.999, 1e-8, t, _stream()))
        pn, mn, vn = O.np_adam(pn, gr.cpu().numpy(), mn, vn, t, 5e-3)
        tref.grad = gr.clone(); topt.step()
    assert float(np.abs(p.cpu().numpy() - pn).max()) <= 1e-6 * float(np.abs(pn).max())
    assert float((p - tref.detach()).abs().max()) <= 1e-6 * float(tref.abs().max())     # and torch's own Adam
    assert torch.equal(shadow, p.to(torch.bfloat16))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_simt_gemm_all_majors_and_epilogues(lib, dtype):
    from tests.gpu_util import gemm, ref_gemm
    g = torch.Generator(device="cuda").manual_seed(0)
    for (G, M, N, K, akm, bkm, epi) in [(1, 70, 50, 33, True, True, 1), (3, 128, 64, 208, True, True, 2),
                                        (2, 96, 72, 40, True, False, 3), (1, 40, 136, 300, False, False, 4),
                                        (2, 64, 64, 1000, False, False, 4)]:
        def mk(rows, cols, km):
            shape = (G, rows, cols) if km else (G, cols, rows)
            return torch.randn(*shape, device="cuda", generator=g).to(dtype)
        A, B = mk(M, K, akm), mk(N, K, bkm)
        bias = torch.randn(G, N, device="cuda", generator=g) if epi in (1, 2) else None
        aux = torch.randn(G, M, N, device="cuda", generator=g).to(dtype) if epi == 3 else None
        C0 = torch.zeros(G, M, (N + 7) // 8 * 8, device="cuda") if epi == 4 else None
        got = gemm(1, A, B, a_kmajor=akm, b_kmajor=bkm, bias=bias, epi=epi, aux=aux, C_init=C0, split_k=3 if epi == 4 else 1)
        want = ref_gemm(A, B, akm, bkm, bias, epi, aux) if epi != 4 else ref_gemm(A, B, akm, bkm)
        err = float((got.double() - want).abs().max()); scale = float(want.abs().max())
        assert err <= 2e-5 * scale * max(1, K / 64) ** 0.5, (M, N, K, akm, bkm, epi, err, scale)


def test_ring_sample_gathers_rows(lib):
    S, A, cap = 24, 4, 50
    row = lib.lib().mfvae_ring_row_floats(S, A, A)
    storage = torch.zeros(cap * row, device="cuda")
    r = C.c_void_p()
    lib.check(lib.lib().mfvae_ring_create(S, A, A, cap, lib.ptr(storage), C.byref(r)))
    host = torch.arange(70 * row, dtype=torch.float32).reshape(70, row)      # 70 rows into a 50-slot ring: wraps
    lib.check(lib.lib().mfvae_ring_add(r, C.c_void_p(host.data_ptr()), 30, 0, _stream()))
    lib.check(lib.lib().mfvae_ring_add(r, C.c_void_p(host[30:].data_ptr()), 40, 0, _stream()))
    torch.cuda.synchronize()
    assert lib.lib().mfvae_ring_size(r) == 50
    B = 64
    obs = torch.empty(B, S, device="cuda"); nxt = torch.empty(B, S, device="cuda")
    act = torch.empty(B, A, device="cuda"); rew = torch.empty(B, A, device="cuda")
    idx = torch.empty(B, dtype=torch.int32, device="cuda")
    flags = torch.empty(B, 2 * A + 1, device="cuda")
    lib.check(lib.lib().mfvae_ring_sample(r, B, 9, 1, lib.ptr(obs), lib.ptr(act), lib.ptr(nxt), lib.ptr(rew), lib.ptr(flags), lib.ptr(idx), _stream()))
    torch.cuda.synchronize()
    st = storage.view(cap, row)
    ii = idx.long()
    assert int(ii.min()) >= 0 and int(ii.max()) < 50 and len(set(ii.tolist())) > 20
    assert torch.equal(obs, st[ii, :S]) and torch.equal(act, st[ii, S:S + A])
    assert torch.equal(nxt, st[ii, S + A:2 * S + A]) and torch.equal(rew, st[ii, 2 * S + A:2 * S + 2 * A])
    assert torch.equal(flags, st[ii, 2 * S + 2 * A:2 * S + 4 * A + 1])          # terminals | truncations | mask
    # slots 0..19 were overwritten by rows 50..69 (ring semantics of cpprb / flashbax)
    assert torch.equal(st[:20].cpu(), host[50:70]) and torch.equal(st[20:30].cpu(), host[20:30])
    lib.lib().mfvae_ring_destroy(r)


def test_tcgen05_gemm_against_fp64_reference():
    """All three operand-major combinations, ragged edges, groups, split-K and every epilogue (own process)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tc_gemm_check.py"), "full"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-6000:]); sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0 and "TC_GEMM_ALL_OK" in r.stdout


def test_replay_wrappers_match_reference_contracts(lib):
    """MultiAgentCPPRB / JaxFbxBuffer over the device ring: key sets and shapes of the reference wrappers
    (replay_buffer.py:62-81,107-108; jax_buffer.py:45-54,186-188), every sampled row is a stored transition, and the
    dict feeds create_dataset / the packed batch feeds the train step."""
    import mfvae_b200 as M
    spec = O.tiny_spec(3)
    rng = np.random.default_rng(0)

    class Box:                      # duck-typed gymnasium spaces
        def __init__(self, n): self.shape = (n,)

    class Env:
        agents = spec.agents
        def observation_space(self, a): return Box(spec.obs_dim[a])

    buf = M.MultiAgentCPPRB(Env(), max_size=40, batch_size=16)
    rows = []
    for t in range(55):            # wraps the 40-slot ring
        obs = {a: rng.standard_normal(spec.obs_dim[a]).astype(np.float32) for a in spec.agents}
        nxt = {a: rng.standard_normal(spec.obs_dim[a]).astype(np.float32) for a in spec.agents}
        act = {a: int(rng.integers(0, 5)) for a in spec.agents}
        rew = {a: float(rng.standard_normal()) for a in spec.agents}
        term = {a: (t % 7 == 3 and a == spec.agents[0]) for a in spec.agents}; trunc = {a: t % 25 == 24 for a in spec.agents}
        buf.add(obs, nxt, act, rew, term, trunc)
        rows.append(np.concatenate([obs[a] for a in spec.agents]))
        if t % 25 == 24:
            buf.on_episode_end()
    d = buf.sample()
    want_keys = {f"{a}_{k}" for a in spec.agents for k in ("observations", "next_observations", "actions", "rewards", "terminals", "truncations")} | {"mask"}
    assert set(d) == want_keys
    for a in spec.agents:
        assert d[f"{a}_observations"].shape == (16, spec.obs_dim[a]) and d[f"{a}_observations"].dtype == np.float32
        assert d[f"{a}_actions"].shape == (16, 1) and d[f"{a}_rewards"].shape == (16, 1)
    stored = np.stack(rows[15:])                                        # the 40 newest survive
    got = np.concatenate([d[f"{a}_observations"] for a in spec.agents], axis=1)
    for k, r in enumerate(got):
        hit = np.flatnonzero(np.all(np.isclose(stored, r[None, :]), axis=1))
        assert hit.size == 1
        t = 15 + int(hit[0])                                            # the step this row was added at: per-agent flags as added
        assert d[f"{spec.agents[0]}_terminals"][k, 0] == float(t % 7 == 3) and d[f"{spec.agents[1]}_terminals"][k, 0] == 0.0
        assert all(d[f"{a}_truncations"][k, 0] == float(t % 25 == 24) for a in spec.agents) and d["mask"][k, 0] == 0.0
    idx_state, acts, joint, nxt_t, rew_t = M.create_dataset(d, {a: i for i, a in enumerate(spec.agents)})
    assert nxt_t.shape == (16, spec.state_dim) and idx_state[spec.agents[0]].shape == (16, 1 + spec.obs_dim[spec.agents[0]])
    pb = buf.sample_packed()
    assert pb.obs.shape == (16, spec.state_dim) and pb.obs.is_cuda and pb.rew.shape == (16, 3)

    jb = M.JaxFbxBuffer(max_length=32, min_length=8, batch_size=4)
    assert jb.can_sample() is None and jb.sample(0) is None            # reference: prints and returns None before init
    o = {a: rng.standard_normal(spec.obs_dim[a]).astype(np.float32) for a in spec.agents}
    jb.init_buffer(o, {a: 0.0 for a in o}, {a: 0 for a in o}, o, {a: False for a in o})
    for t in range(9):
        jb.add_trans(o, {a: 1.0 for a in o}, {a: 2 for a in o}, o, {a: False for a in o})
        assert bool(jb.can_sample()) == (t + 1 >= 8)
    e = jb.sample(123).experience
    assert set(e) == {f"{a}_{k}" for a in spec.agents for k in ("obs", "act", "next_obs", "rew")} | {"done"}
    a0 = spec.agents[0]
    assert e[f"{a0}_obs"].shape == (4, spec.obs_dim[a0], 1) and e[f"{a0}_act"].shape == (4, 1, 1) and e["done"].shape == (4, 1, 1)
    assert np.allclose(e[f"{a0}_obs"][:, :, 0], o[a0][None, :]) and np.all(e[f"{a0}_act"] == 2) and np.all(e["done"] == 0.0)

    # continuous actions (Box action spaces): the ring keeps whole action vectors, sample() hands them back per agent
    act_dim = {a: (5 if i < 2 else 3) for i, a in enumerate(spec.agents)}
    cb = M.MultiAgentCPPRB(max_size=16, batch_size=8, agents=spec.agents, obs_dim=spec.obs_dim, act_dim=act_dim)
    acts_added = []
    for t in range(10):
        act = {a: rng.uniform(0, 1, act_dim[a]).astype(np.float32) for a in spec.agents}
        cb.add(o, o, act, {a: 0.5 for a in o}, {a: False for a in o}, {a: False for a in o})
        acts_added.append(np.concatenate([act[a] for a in spec.agents]))
    dc = cb.sample()
    assert all(dc[f"{a}_actions"].shape == (8, act_dim[a]) for a in spec.agents)
    got = np.concatenate([dc[f"{a}_actions"] for a in spec.agents], axis=1)
    for r in got:
        assert np.any(np.all(np.isclose(np.stack(acts_added), r[None, :]), axis=1))
    assert cb.sample_packed().act.shape == (8, sum(act_dim.values()))
