"""CPU: the C-ABI library loads and exports every symbol include/mfvae.h declares; host-side logic
(arena table, parameter views, state_dict surface, staging) behaves like the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import mavae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from mfvae_b200 import build
    build.build()
    import mfvae_b200
    return mfvae_b200


def test_every_declared_symbol_is_exported(built):
    hdr = open(os.path.join(ROOT, "include", "mfvae.h")).read()
    declared = set(re.findall(r"\b(mfvae_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = C.CDLL(built._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(built._lib.SIGNATURES), declared ^ set(built._lib.SIGNATURES)
    assert built._lib.lib().mfvae_version() >= 100


def test_no_cpu_path(built):
    spec = O.tiny_spec(3, idx_features=16, latent=8, act_features=8)
    m = built.MAVAE(16, 8, 8, True, spec.agents, spec.obs_dim, spec.n_act, "cpu", precision="fp32")
    t = O.synth_transition(spec, 4, 0)
    idx_state, acts, *_ = built.create_dataset(t, {a: i for i, a in enumerate(spec.agents)})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(idx_state, acts)
    with pytest.raises(RuntimeError):
        m.adam_step(1e-3)


def test_state_dict_surface_matches_reference(built):
    """39 keys with the reference's names and the reference's registered parameter count."""
    spec = O.simple_tag_spec()
    m = built.MAVAE(64, 64, 64, True, spec.agents, spec.obs_dim, spec.n_act, "cpu")
    keys = list(m.state_dict().keys())
    want = ["idx_emb.weight"]
    for pre in ("decoder", "state_decoder", "reward_decoder"):
        for i in (0, 2, 4, 6, 8, 10):
            want += [f"{pre}.net.{i}.weight", f"{pre}.net.{i}.bias"]
    want += ["reward_linear.weight", "reward_linear.bias"]
    assert keys == want
    assert sum(p.numel() for p in m.parameters()) == 29_096_880           # SURVEY appendix A [probe]
    live = sum(p.numel() for n, p in m.named_parameters() if not n.startswith("decoder."))
    assert live == 17_451_820
    assert isinstance(m.encoders, dict) and not isinstance(m.encoders, torch.nn.ModuleDict)
    assert sum(p.numel() for e in m.encoders.values() for p in e.parameters()) == 2_676_480
    assert m.reward_linear.weight.shape == (40, 40) and bool((m.reward_linear.weight == 1).all())
    assert m.state_decoder.net[10].weight.shape == (5660, 1024)
    assert m.encoders["adversary_0"].net[0].weight.shape == (64, 206)
    assert m.encoders["agent_0"].net[0].weight.shape == (64, 204)
    # parameters are views of one arena and share its version counter
    v0 = m._arena._version
    with torch.no_grad():
        m.state_decoder.net[0].bias.add_(1.0)
    assert m._arena._version > v0
    sd = m.state_dict()
    m2 = built.MAVAE(64, 64, 64, True, spec.agents, spec.obs_dim, spec.n_act, "cpu")
    m2.load_state_dict(sd)
    assert torch.equal(m2.state_decoder.net[0].bias, m.state_decoder.net[0].bias)


def test_load_named_roundtrip(built):
    spec = O.tiny_spec(3, idx_features=16, latent=8, act_features=8)
    P = O.init_params(spec, 3)
    m = built.MAVAE(16, 8, 8, True, spec.agents, spec.obs_dim, spec.n_act, "cpu", precision="fp32")
    m.load_named(P)
    mine = m.named_arena_tensors()
    assert set(mine) == set(P)
    for k in P:
        assert torch.equal(mine[k], P[k]), k
    # padding columns of the first encoder layer stay zero
    for t in m._table:
        if t.cols != t.ld:
            blk = m._arena[t.offset:t.offset + t.rows * t.ld].view(t.rows, t.ld)
            assert float(blk[:, t.cols:].abs().max()) == 0.0


def test_create_dataset_matches_reference_restatement(built):
    spec = O.tiny_spec(4)
    t = O.synth_transition(spec, 9, seed=11)
    cb = {a: i for i, a in enumerate(spec.agents)}
    got = built.create_dataset(t, cb)
    want = O.stage_batch(t, cb)
    for a in spec.agents:
        assert torch.equal(got[0][a], want[0][a]) and torch.equal(got[1][a], want[1][a])
    for g, w in zip(got[2:], want[2:]):
        assert torch.equal(g, w)


def test_host_stager_cpu(built):
    spec = O.tiny_spec(3)
    t = O.synth_transition(spec, 6, seed=5)
    cb = {a: i for i, a in enumerate(spec.agents)}
    pb = built.HostStager("cpu").stage(t, cb)
    _, _, _, nxt, rew = O.stage_batch(t, cb)
    assert torch.equal(pb.next, nxt) and torch.equal(pb.rew, rew)
    assert pb.obs.shape == (6, spec.state_dim) and pb.act.shape == (6, 3)


def test_host_stager_bf16_observation_rows(built):
    """HostStager(obs_dtype=bfloat16): rows [obs bf16 | act | next | rew fp32] -- the e2e row format of bench.py; observations
    are the bf16 rounding of the fp32 values, everything else is untouched, 25 % fewer bytes cross PCIe."""
    spec = O.tiny_spec(3)
    t = O.synth_transition(spec, 6, seed=5)
    cb = {a: i for i, a in enumerate(spec.agents)}
    st32, st16 = built.HostStager("cpu"), built.HostStager("cpu", obs_dtype=torch.bfloat16)
    a, b = st32.stage(t, cb), st16.stage(t, cb)
    assert b.obs.dtype == torch.bfloat16 and torch.equal(b.obs, a.obs.to(torch.bfloat16))
    assert torch.equal(b.act, a.act) and torch.equal(b.next, a.next) and torch.equal(b.rew, a.rew)
    assert st16.h2d_bytes < st32.h2d_bytes and b.obs.is_contiguous() and b.next.is_contiguous()
    b2 = st16.stage(t, cb); b3 = st16.stage(t, cb)          # ring slots are reused
    assert torch.equal(b3.obs, b.obs) and torch.equal(b2.next, a.next)
    with pytest.raises(TypeError):
        built.HostStager("cpu", obs_dtype=torch.float16)


def test_trainer_surface(built):
    spec = O.tiny_spec(3, idx_features=16, latent=8, act_features=8)
    m = built.MAVAE(16, 8, 8, True, spec.agents, spec.obs_dim, spec.n_act, "cpu", precision="fp32")
    with pytest.raises(AssertionError):
        built.Trainer("SGD", m, 1e-3, built.loss_s_r_vae_fn, device="cpu")
    tr = built.Trainer("Adam", m, 5e-3, built.loss_s_r_vae_fn, device="cpu")
    for f in ("sigma", "mu", "nu", "sigma_new", "mu_new", "beta", "lr", "loss_func", "loss", "opt", "device"):
        assert hasattr(tr, f)
    y = torch.randn(5, 3)
    assert torch.equal(tr.normalize(y), y) and torch.equal(tr.denormalize(y), y)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(tr.opt, T_max=50, eta_min=1e-4)   # main.py:53 works on FusedAdam
    assert abs(tr.opt.param_groups[0]["lr"] - 5e-3) < 1e-12
    for s in range(1, 4):
        assert abs(built.cosine_lr(s) - O.cosine_lr(s)) < 1e-15
    # POP-ART with per-agent statistics keeps the un-normalised output of reward_linear unchanged
    tr2 = built.Trainer("POPART", m, 5e-3, built.loss_s_r_vae_fn, beta=0.3, device="cpu")
    x = torch.randn(7, 3)
    before = tr2.denormalize(torch.nn.functional.linear(x, m.reward_linear.weight, m.reward_linear.bias))
    tr2.art(torch.randn(64, 3) * 3 + 1); tr2.pop(); tr2.update_stats()
    after = tr2.denormalize(torch.nn.functional.linear(x, m.reward_linear.weight, m.reward_linear.bias))
    assert torch.allclose(before, after, atol=1e-5)


def test_gradient_buckets_cover_optimized_prefix(built):
    spec = O.simple_tag_spec()
    m = built.MAVAE(64, 64, 64, True, spec.agents, spec.obs_dim, spec.n_act, "cpu")
    b = sorted((lo, hi) for _, lo, hi in m.grad_buckets())
    assert b[0][0] == 0 and b[-1][1] == m._n_opt
    for (l0, h0), (l1, h1) in zip(b, b[1:]):
        assert h0 == l1


def test_continuous_action_model_surface_and_staging(built):
    """descrete_act=False (reference model.py:123,148): per-agent ActionEncoder MLPs act_dim -> 64 -> C backed by the arena
    (layer 0 padded to 8 input columns in the arena, views hide the padding), still unregistered like the reference's
    plain dict; create_dataset keeps [B, act_dim] action matrices."""
    act_dim = {"adversary_0": 5, "adversary_1": 5, "adversary_2": 5, "agent_0": 3}
    spec = O.tiny_spec(4, idx_features=64, latent=32, act_features=64, discrete_act=False, act_dim=act_dim)
    m = built.MAVAE(64, 32, 64, False, spec.agents, spec.obs_dim, act_dim, "cpu", precision="fp32")
    for a in spec.agents:
        enc = m.action_encoder[a]
        assert isinstance(enc, built.ActionEncoder)
        assert enc.net[0].weight.shape == (64, act_dim[a]) and enc.net[2].weight.shape == (64, 64)
        assert enc.net[0].weight.stride(0) == 8                      # arena row pitch: K padded to 8
    names = set(m.named_arena_tensors())
    assert "action_encoder.agent_0.net.0.weight" in names and "action_encoder.agent_0.net.2.bias" in names
    assert not any(k.startswith("action_encoder") for k in m.state_dict())       # unregistered, as in the reference
    P = O.init_params(spec, 3)
    m.load_named(P)
    assert torch.equal(m.action_encoder["agent_0"].net[0].weight.detach(), P["action_encoder.agent_0.net.0.weight"])
    t = O.synth_transition(spec, 6, 1)
    cb = {a: i for i, a in enumerate(spec.agents)}
    mine, ref = built.create_dataset(t, cb), O.stage_batch(t, cb)
    for a in spec.agents:
        assert mine[1][a].shape == (6, act_dim[a]) and torch.equal(mine[1][a], ref[1][a])
        assert torch.equal(mine[0][a], ref[0][a])
    assert torch.equal(mine[3], ref[3]) and torch.equal(mine[4], ref[4])


def test_jax_front_end_weights(built):
    """jax_ver/trainer.py:42-43,64: kl 0.1, r 0.5, state term weighted 1 - r."""
    from mfvae_b200 import jax_trainer as J
    assert (J.kl_weight, J.r_weight) == (0.1, 0.5)
    assert J.loss_weights() == (0.1, 0.5, 0.5)
    J.r_weight = 0.25                       # read at call time, like the reference's module globals
    try:
        assert J.loss_weights() == (0.1, 0.25, 0.75)
    finally:
        J.r_weight = 0.5


def test_launch_list_parser_on_committed_profile():
    """tools/parse_launches.py reproduces the per-kernel summary committed under profiles/."""
    import subprocess
    import sys
    csv = os.path.join(ROOT, "profiles", "r1_launches_dram_cfg2_b4096.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parse_launches.py"), csv], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "gemm_tc_kernel (all instantiations): 36 launches" in out.stdout
    assert "launches in step 59" in out.stdout


def test_r2_launch_list_summary_matches_the_committed_profile(tmp_path):
    """tools/summarize_launches.py reproduces profiles/r2_traffic.json (bench.py's roofline.traffic) from the committed ncu list."""
    import json
    import shutil
    import subprocess
    import sys
    want = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    keep = {f: open(os.path.join(ROOT, "profiles", f)).read() for f in ("r2_traffic.json", "r2_ncu_step_summary.txt")}
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"),
                              os.path.join(ROOT, "profiles", "r2_launches_cfg2_b4096.csv"), "r2"], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        got = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        assert got["cfg2_b4096"] == want["cfg2_b4096"]
        assert got["cfg2_b4096"]["tensor_core_kernels_all_launches"]["launches"] == 33
        assert got["cfg2_b4096"]["whole_step"]["launches"] == 58
    finally:
        for f, txt in keep.items():
            open(os.path.join(ROOT, "profiles", f), "w").write(txt)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys; under
    torchrun every rank but 0 exits 0 without work."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "torch_ver", "model.py"))
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["steps"] == 1 and line["warmup"] == 1                # the arm runs the --steps / --warmup it is given
    assert set(line["config"]) == {"workload", "batch_per_gpu", "global_batch", "parallelism", "source", "l2", "flop_per_sample",
                                   "flop_per_sample_executed"}
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    env1 = dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                        capture_output=True, text=True, timeout=120, env=env1, cwd=ROOT)
    assert r1.returncode == 0 and not [l for l in r1.stdout.splitlines() if l.startswith("{")]


# ---------------------------------------------------------------------------------------------------------------
# interoperability with the UNMODIFIED reference (imported from /root/reference in the build container; these tests are
# skipped where it is absent, e.g. on the GPU box)
# ---------------------------------------------------------------------------------------------------------------
REF = "/root/reference/torch_ver"
needs_reference = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model.py")), reason="reference sources not present")


def _reference():
    import importlib.util
    import sys
    sys.dont_write_bytecode = True
    mods = []
    for name in ("model", "trainer"):
        sp = importlib.util.spec_from_file_location(f"_ref_{name}", os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(sp)
        sp.loader.exec_module(mod)
        mods.append(mod)
    return mods


@needs_reference
def test_saved_state_dict_loads_strict_into_the_reference_and_back(built, tmp_path):
    """MAVAE.save (model.py:175-176) writes the reference's 39 keys: the file loads with strict=True into the unmodified
    reference MAVAE, and a file written by the reference loads back here; the full checkpoint additionally restores the
    unregistered encoders / action tables, Adam moments + step and the Philox position."""
    ref_model, _ = _reference()
    spec = O.simple_tag_spec(latent=32)
    m = built.MAVAE(64, 32, 64, True, spec.agents, spec.obs_dim, spec.n_act, "cpu")
    p = str(tmp_path / "test.pt")
    m.save(p)
    r = ref_model.MAVAE(64, 32, 64, True, spec.agents, spec.obs_dim, spec.n_act, "cpu")
    missing = r.load_state_dict(torch.load(p), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(r.state_dict()[k], v), k
    # and back: a file written by the reference's own save()
    with torch.no_grad():
        for q in r.parameters():
            q.add_(0.25)
    p2 = str(tmp_path / "ref.pt")
    r.save(p2)
    m.load_state_dict(torch.load(p2), strict=True)
    for k, v in r.state_dict().items():
        assert torch.equal(m.state_dict()[k], v), k
    # full checkpoint round trip (CPU tensors; the layout-only handle is enough for the bookkeeping)
    m._m.uniform_(); m._v.uniform_(); m._adam_t = 17; m.philox_step = 123
    p3 = str(tmp_path / "full.pt")
    m.save_checkpoint(p3)
    m2 = built.MAVAE(64, 32, 64, True, spec.agents, spec.obs_dim, spec.n_act, "cpu")
    m2.load_checkpoint(p3)
    n = m._n_opt
    assert torch.equal(m2._arena, m._arena) and torch.equal(m2._m[:n], m._m[:n]) and torch.equal(m2._v[:n], m._v[:n])
    assert (m2._adam_t, m2.philox_step, m2.philox_seed) == (17, 123, m.philox_seed)
    ck = torch.load(p3)
    assert set(ck["state_dict"]) == set(r.state_dict()) and "encoders.adversary_0.net.0.weight" in ck["unregistered"]
    r.load_state_dict(ck["state_dict"], strict=True)          # the reference can read the weights out of the full checkpoint too


@needs_reference
def test_reference_trainer_failures_this_package_fixes(built):
    """the reference's POP-ART mode raises for every batch (pop() multiplies the [A, A] weight in place by a
    [B, A] ratio, trainer.py:72-74) and training_model raises TypeError (adds the loss 4-tuple to a float,
    trainer.py:112-113).  Both are reproduced on the unmodified reference; the drop-in Trainer runs the same calls."""
    import contextlib
    import io
    ref_model, ref_trainer = _reference()
    spec = O.tiny_spec(3, idx_features=16, latent=8, act_features=8)
    cb = {a: i for i, a in enumerate(spec.agents)}
    B = 6
    t = O.synth_transition(spec, B, 0)
    r = ref_model.MAVAE(16, 8, 8, True, spec.agents, spec.obs_dim, spec.n_act, "cpu")
    idx_state, acts, joint, nxt, rew = ref_trainer.create_dataset(t, cb)
    tr = ref_trainer.Trainer("POPART", r, 1e-3, ref_model.loss_s_r_vae_fn, beta=0.3, device="cpu")
    with pytest.raises(RuntimeError), contextlib.redirect_stdout(io.StringIO()):
        tr.forward(idx_state, acts, nxt, rew)                      # trainer.py:72-74

    class OneBatch:
        def sample(self):
            return t
    tr = ref_trainer.Trainer("Adam", r, 1e-3, ref_model.loss_s_r_vae_fn, device="cpu")
    with pytest.raises(TypeError), contextlib.redirect_stdout(io.StringIO()):
        tr.training_model(OneBatch(), 1, cb)                       # trainer.py:112-113
    # the drop-in: POP-ART statistics are per agent, so pop() is well-formed for any batch (CPU: bookkeeping only)
    m = built.MAVAE(16, 8, 8, True, spec.agents, spec.obs_dim, spec.n_act, "cpu", precision="fp32")
    mine = built.Trainer("POPART", m, 1e-3, built.loss_s_r_vae_fn, beta=0.3, device="cpu")
    mine.art(rew); mine.pop(); mine.update_stats()
    assert mine.sigma.shape == (3,) and bool(torch.isfinite(m.reward_linear.weight).all())


def test_second_backward_without_step_raises(built):
    spec = O.tiny_spec(3, idx_features=16, latent=8, act_features=8)
    m = built.MAVAE(16, 8, 8, True, spec.agents, spec.obs_dim, spec.n_act, "cpu", precision="fp32")
    opt = built.FusedAdam(m, 1e-3)
    m._grads_pending = True                     # state after one backward pass
    with pytest.raises(RuntimeError, match="second backward"):
        m._check_accumulation()
    opt.zero_grad()
    m._check_accumulation()                     # consumed
    sd = opt.state_dict()
    assert sd["state"]["step"] == 0 and sd["state"]["exp_avg"].numel() == m._n_opt and sd["param_groups"][0]["lr"] == 1e-3
    m._m.fill_(2.0); m._adam_t = 5
    sd = opt.state_dict()
    m._m.zero_(); m._adam_t = 0
    opt.load_state_dict(sd)
    assert m._adam_t == 5 and float(m._m[0]) == 2.0


def test_pack_rejects_out_of_range_indices_like_nn_embedding(built):
    spec = O.tiny_spec(3, idx_features=16, latent=8, act_features=8)
    m = built.MAVAE(16, 8, 8, True, spec.agents, spec.obs_dim, spec.n_act, "cpu", precision="fp32")
    cb1 = {a: i + 1 for i, a in enumerate(spec.agents)}            # 1-based codebook: index 3 does not exist
    idx_state, acts, *_ = built.create_dataset(O.synth_transition(spec, 4, 0), cb1)
    with pytest.raises(IndexError):
        m.pack(idx_state, acts)
    cb = {a: i for i, a in enumerate(spec.agents)}
    idx_state, acts, *_ = built.create_dataset(O.synth_transition(spec, 4, 0), cb)
    acts[spec.agents[0]][0, 0] = 5.0                               # Discrete(5): 5 is out of range
    with pytest.raises(IndexError):
        m.pack(idx_state, acts)
