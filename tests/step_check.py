"""Run a golden case through the CUDA path and report its deviation from (a) the golden vectors minted from the
reference and (b) the oracle, as one JSON line.  Run in its own process (a device trap must not poison pytest):

    python tests/step_check.py <case> <fp32|bf16> <auto|simt|tcgen05> [dropin|torchloss|jointmse|fast] [eps|philox] [fusion] [optenc]

Cases are the golden cases of tests/golden/*.npz (minted from the unmodified reference) or the synthetic cases of
tests/golden_util.py::SYNTH (oracle only).  `optenc` = MAVAE(optimize_encoders=True) against OracleState(optimize_encoders=True).
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mavae_oracle as O                     # noqa: E402
from tests.golden_util import load_case, digest, step_inputs   # noqa: E402
import mfvae_b200 as M                                     # noqa: E402
import ctypes as C                                         # noqa: E402
from mfvae_b200 import _lib as Lb                          # noqa: E402


def oracle_joint_mse(P, spec, idx_state, acts, eps, joint):
    """autograd of the oracle's loss_vae_fn restatement (reference model.py:8-16)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    rs, rr, mus, lvs = O.forward(leaves, spec, idx_state, acts, eps)
    loss = O.loss_joint_mse(joint, torch.cat([rs, rr], 1), mus, lvs)
    loss.backward()
    return float(loss.detach()), {k: v.grad for k, v in leaves.items() if v.grad is not None}, (rs.detach(), rr.detach(), [x.detach() for x in mus], [x.detach() for x in lvs])


def rel_l2(got, want):
    got = got.detach().double().cpu(); want = want.detach().double().cpu()
    return float((got - want).norm() / max(float(want.norm()), 1e-30))


def digest_err(got, want):
    l2 = max(abs(want[1]), 1e-30)
    scale = max(np.abs(want[2:]).max(), 1e-30)
    return max(abs(got[1] - want[1]) / l2, float(np.abs(got[2:] - want[2:]).max() / scale))


def fp64_metrics(out, P, spec, idx_state, acts, eps, nxt, rew, huber, G32, mine_grads):
    """fp32 is not exact: a ReLU unit whose pre-activation lies within round-off of zero takes either branch, and one such
    unit moves a whole gradient row.  The yardstick for an fp32 implementation is therefore the exact (fp64) evaluation,
    with the fp32 ORACLE's own distance from it as the scale of what fp32 can deliver at this shape."""
    P64 = {k: v.double() for k, v in P.items()}
    _, G64, _ = O.grads(P64, spec, {a: t.double() for a, t in idx_state.items()}, acts, {a: t.double() for a, t in eps.items()},
                        nxt.double(), rew.double(), huber)
    keys = sorted(mine_grads)

    def cat(G):
        return torch.cat([G[k].detach().double().cpu().reshape(-1) for k in keys])
    ref = cat(G64)
    out["whole_grad_rel_vs_fp64"] = float((cat(mine_grads) - ref).norm() / ref.norm())
    out["oracle32_whole_grad_rel_vs_fp64"] = float((cat(G32) - ref).norm() / ref.norm())
    out["grad_rel_max_vs_fp64"] = max(rel_l2(mine_grads[k], G64[k]) for k in keys)
    out["oracle32_grad_rel_max_vs_fp64"] = max(rel_l2(G32[k], G64[k]) for k in keys)
    out["grad_rel_median_vs_fp64"] = float(np.median([rel_l2(mine_grads[k], G64[k]) for k in keys]))


def whole_grad(mine, ref):
    """relative L2 and cosine of the CONCATENATED gradient (every tensor that takes part) against `ref`."""
    a = torch.cat([p.grad.detach().double().cpu().reshape(-1) for _, p in sorted(mine.items())])
    b = torch.cat([ref[k].detach().double().cpu().reshape(-1) for k, _ in sorted(mine.items())])
    return float((a - b).norm() / b.norm()), float(torch.dot(a, b) / (a.norm() * b.norm()))


def main(case, precision, engine, mode="dropin", rng="eps", fusion="auto", flags=""):
    spec, rec = load_case(case)
    golden = "losses" in rec
    optenc = "optenc" in flags
    dev = "cuda:0"
    huber = bool(rec["huber"])
    m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, spec.discrete_act, spec.agents, spec.obs_dim,
                spec.n_act if spec.discrete_act else spec.act_dim, dev, enc_hidden=spec.enc_hidden, dec_hidden=spec.dec_hidden,
                precision=precision, engine=engine, huber=huber, fusion=fusion, optimize_encoders=optenc,
                include_dead_decoder=golden)
    P = O.init_params(spec, int(rec["param_seed"]))
    m.load_named(P)
    st = O.OracleState(spec, {k: v.clone() for k, v in P.items() if not k.startswith("decoder.")}, optimize_encoders=optenc)
    opt = M.FusedAdam(m, 0.005)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=50, eta_min=1e-4)
    L = spec.latent
    out = {"case": case, "precision": precision, "engine": engine, "mode": mode, "rng": rng,
           "loss_rel_golden": [], "loss_rel_oracle": [], "losses": []}
    for step in range(3):
        trans, codebook, eps_all = step_inputs(spec, rec, step)
        idx_state, acts, joint, nxt, rew = M.create_dataset(trans, codebook)
        m.philox_step = step
        eps_dev = eps_all.to(dev)
        if rng == "philox":
            # draw inside the kernels; the oracle is fed the kernel's own stream (dumped through the C ABI)
            dump = torch.empty(int(rec["batch"]), spec.n_agents * L, device=dev)
            Lb.check(Lb.lib().mfvae_philox_normal(Lb.ptr(dump), dump.shape[0], dump.shape[1], m.philox_seed, step, 0,
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
            out.setdefault("philox_vs_numpy_maxabs", []).append(float((dump.cpu() - eps_all).abs().max()))
            eps_all = dump.cpu()
            eps_dev = None
        lr = opt.param_groups[0]["lr"]
        if mode == "dropin":
            recon_s, recon_r, mu_all, lv_all = m(idx_state, acts, eps=eps_dev)
            loss, sl, rl, kl = M.loss_s_r_vae_fn(recon_s, recon_r, nxt, rew, mu_all, lv_all, dev, using_huber_loss=huber)
            opt.zero_grad()
            loss.backward()
        elif mode == "torchloss":
            # the reference's own loss arithmetic in torch ops on our outputs -> autograd bridge (backward_ext)
            recon_s, recon_r, mu_all, lv_all = m(idx_state, acts, eps=eps_dev)
            F = torch.nn.functional
            fn = F.huber_loss if huber else F.mse_loss
            sl = fn(nxt.to(dev), recon_s); rl = fn(rew.to(dev), recon_r)
            kl = 0.0
            for mu_, lv_ in zip(mu_all, lv_all):
                kl = kl + torch.mean(-0.5 * torch.sum(1 + lv_ - mu_ ** 2 - torch.exp(lv_), 1), 0)
            loss = sl + 0.005 * rl + 0.0025 * kl
            loss.backward()
        elif mode == "jointmse":
            # loss_vae_fn (model.py:8-16) on the concatenated reconstruction: routed to the fused CUDA joint-MSE loss
            recon_s, recon_r, mu_all, lv_all = m(idx_state, acts, eps=eps_dev)
            n0 = Lb.lib().mfvae_launch_count()
            loss = M.loss_vae_fn(joint, torch.cat([recon_s, recon_r], 1), mu_all, lv_all, dev)
            out["jointmse_on_cuda_path"] = bool(Lb.lib().mfvae_launch_count() > n0 and "FusedLossFn" in type(loss.grad_fn).__name__)
            sl = rl = kl = loss
            opt.zero_grad()
            loss.backward()
        else:
            pb = m.pack(idx_state, acts, eps=eps_dev)
            pb.next, pb.rew = nxt.to(dev), rew.to(dev)
            pb.idx = None
        eps = {a: eps_all[:, i * L:(i + 1) * L] for i, a in enumerate(spec.agents)}
        if mode == "fast":
            # fast path = fwd+loss+bwd+adam in one call; compare losses and the post-step parameters only
            Gq = None
            if step == 0 and precision == "bf16":      # gradients of the fused step against the bf16-emulating oracle
                _, Gq, _ = O.grads(st.P, spec, idx_state, acts, eps, nxt, rew, huber, emulate_bf16=True)
            P0 = {k: v.clone() for k, v in st.P.items()}
            losses_dev = m.train_step(pb, lr)
            sched.step()
            got_losses = [float(x) for x in losses_dev.cpu()]
            if Gq is not None:
                qerr = {k: rel_l2(p.grad, Gq[k]) for k, p in m.named_arena_tensors().items()}
                qw = max(qerr, key=qerr.get)
                out["grad_rel_max_vs_bf16_oracle"] = qerr[qw]; out["grad_rel_worst_vs_bf16_oracle"] = qw
                out["grad_rel_median_vs_bf16_oracle"] = float(np.median(list(qerr.values())))
                out["whole_grad_rel_vs_bf16_oracle"], out["whole_grad_cos_vs_bf16_oracle"] = whole_grad(m.named_arena_tensors(), Gq)
            if step == 0:
                g_snap = {k: p.grad.detach().clone() for k, p in m.named_arena_tensors().items()}
            o_losses, G, outs = O.train_step(st, idx_state, acts, eps, nxt, rew, lr, huber)
            if step == 0:
                class _G:                      # whole_grad wants .grad
                    def __init__(self, g): self.grad = g
                out["whole_grad_rel"], out["whole_grad_cos"] = whole_grad({k: _G(v) for k, v in g_snap.items()}, G)
                gerr = {k: rel_l2(v, G[k]) for k, v in g_snap.items()}
                out["grad_rel_max"] = max(gerr.values()); out["grad_rel_worst"] = max(gerr, key=gerr.get)
                out["grad_rel_median"] = float(np.median(list(gerr.values())))
                if not golden and precision == "fp32":
                    fp64_metrics(out, P0, spec, idx_state, acts, eps, nxt, rew, huber, G, g_snap)
        else:
            got_losses = [float(loss), float(sl), float(rl), float(kl)]
            P0 = st.P
            # oracle on the same inputs, same current parameters
            if mode == "jointmse":
                jl, G, outs = oracle_joint_mse(st.P, spec, idx_state, acts, eps, joint)
                o_losses = [jl] * 4
            else:
                o_losses, G, outs = O.grads(st.P, spec, idx_state, acts, eps, nxt, rew, huber)
            if step == 0:
                rs, rr, mus, lvs = outs
                out["recon_s_rel"] = rel_l2(recon_s, rs); out["recon_r_rel"] = rel_l2(recon_r, rr)
                out["mu_rel"] = rel_l2(torch.cat(list(mu_all), 1), torch.cat(mus, 1))
                out["logvar_rel"] = rel_l2(torch.cat(list(lv_all), 1), torch.cat(lvs, 1))
                if golden:
                    out["golden_out_err"] = max(
                        digest_err(digest(recon_s, "out.recon_s"), rec["out.recon_s"]),
                        digest_err(digest(recon_r, "out.recon_r"), rec["out.recon_r"]),
                        digest_err(digest(torch.cat(list(mu_all), 1), "out.mu"), rec["out.mu"]),
                        digest_err(digest(torch.cat(list(lv_all), 1), "out.logvar"), rec["out.logvar"]))
                    if mode == "jointmse":
                        out["jointmse_loss_rel_golden"] = abs(got_losses[0] - float(rec["joint_mse_loss"])) / abs(float(rec["joint_mse_loss"]))
                mine = m.named_arena_tensors()
                gerr, gold = {}, {}
                for k, p in mine.items():
                    gerr[k] = rel_l2(p.grad, G[k])
                    if golden and mode != "jointmse":
                        gold[k] = digest_err(digest(p.grad, "grad." + k), rec["grad." + k])
                out["whole_grad_rel"], out["whole_grad_cos"] = whole_grad(mine, G)
                if not golden and precision == "fp32" and mode != "jointmse":
                    fp64_metrics(out, P0, spec, idx_state, acts, eps, nxt, rew, huber, G, {k: p.grad for k, p in mine.items()})
                out["grad_rel_top"] = sorted(((round(v, 6), k) for k, v in gerr.items()), reverse=True)[:10]
                worst = max(gerr, key=gerr.get)
                out["grad_rel_max"] = gerr[worst]; out["grad_rel_worst"] = worst
                out["grad_rel_median"] = float(np.median(list(gerr.values())))
                big = {k: v for k, v in gerr.items() if k.startswith(("state_decoder", "reward_decoder", "idx_emb", "reward_linear"))}
                out["grad_rel_max_registered"] = max(big.values())
                if precision == "bf16" and mode != "jointmse":
                    # the same algorithm with the CUDA path's bf16 rounding points (oracle emulate_bf16): isolates kernel
                    # errors from the ReLU-mask flips any bf16 evaluation shows against an fp32 run
                    _, Gq, outs_q = O.grads(st.P, spec, idx_state, acts, eps, nxt, rew, huber, emulate_bf16=True)
                    qerr = {k: rel_l2(p.grad, Gq[k]) for k, p in mine.items()}
                    qw = max(qerr, key=qerr.get)
                    out["grad_rel_max_vs_bf16_oracle"] = qerr[qw]; out["grad_rel_worst_vs_bf16_oracle"] = qw
                    out["grad_rel_median_vs_bf16_oracle"] = float(np.median(list(qerr.values())))
                    out["grad_rel_top_vs_bf16_oracle"] = sorted(((round(v, 6), k) for k, v in qerr.items()), reverse=True)[:6]
                    out["recon_s_rel_vs_bf16_oracle"] = rel_l2(recon_s, outs_q[0])
                    out["fp32_vs_bf16_oracle_grad_rel_max"] = max(rel_l2(Gq[k], G[k]) for k in mine)
                    out["whole_grad_rel_vs_bf16_oracle"], out["whole_grad_cos_vs_bf16_oracle"] = whole_grad(mine, Gq)

                    class _G:
                        def __init__(self, g): self.grad = g
                    out["oracle_bf16_whole_grad_rel_vs_fp32"], _ = whole_grad({k: _G(Gq[k]) for k in mine}, G)
                if gold:
                    gw = max(gold, key=gold.get)
                    out["golden_grad_err"] = gold[gw]; out["golden_grad_worst"] = gw
            opt.step()
            sched.step()
            if mode == "jointmse":        # oracle Adam on the joint-MSE gradients
                st.t += 1
                for n_ in (list(st.P.keys()) if st.optimize_encoders else O.registered_names(spec)):
                    if n_ in G:
                        if n_ not in st.m:
                            st.m[n_] = torch.zeros_like(st.P[n_]); st.v[n_] = torch.zeros_like(st.P[n_])
                        st.P[n_], st.m[n_], st.v[n_] = O.adam_update(st.P[n_], G[n_], st.m[n_], st.v[n_], st.t, lr)
            else:
                O.train_step(st, idx_state, acts, eps, nxt, rew, lr, huber)
        out["losses"].append(got_losses)
        if golden and mode != "jointmse":
            out["loss_rel_golden"].append(max(abs(g - w) / max(abs(w), 1e-30) for g, w in zip(got_losses, rec["losses"][step])))
        out["loss_rel_oracle"].append(max(abs(g - w) / max(abs(w), 1e-30) for g, w in zip(got_losses, o_losses)))
    torch.cuda.synchronize()
    mine = m.named_arena_tensors()
    perr = {k: rel_l2(p, st.P[k]) for k, p in mine.items()}
    pw = max(perr, key=perr.get)
    out["param3_rel_max"] = perr[pw]; out["param3_worst"] = pw
    pa = torch.cat([p.detach().double().cpu().reshape(-1) for _, p in sorted(mine.items())])
    pb_ = torch.cat([st.P[k].double().reshape(-1) for k, _ in sorted(mine.items())])
    out["param3_whole_rel"] = float((pa - pb_).norm() / pb_.norm())
    if golden and mode != "jointmse" and not optenc:
        out["golden_param3_err"] = max(digest_err(digest(p, "param3." + k), rec["param3." + k]) for k, p in mine.items())
    print("STEP_CHECK " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main(*sys.argv[1:])
