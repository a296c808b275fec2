"""Fused kernels (fusion="encoder") against the layer-by-layer kernels (fusion="none") of the same library on identical
weights, batch and Philox stream.  Both evaluate the same operations with the same bf16 rounding points, so they must
agree far more tightly than either agrees with the fp32 oracle.  One JSON line; run in its own process:

    python tests/fused_check.py <latent> <batch> [explicit_idx]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mavae_oracle as O      # noqa: E402
import mfvae_b200 as M                    # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def main(latent, B, explicit_idx=False):
    dev = "cuda:0"
    spec = O.simple_tag_spec(latent=latent)
    models = {}
    for fusion in ("none", "encoder", "loss"):
        torch.manual_seed(7)
        m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, dev,
                    precision="bf16", fusion=fusion, include_dead_decoder=False)
        models[fusion] = m
    for k in ("encoder", "loss"):
        models[k].load_named(models["none"].named_arena_tensors())
    g = torch.Generator(device=dev).manual_seed(11)
    S, A = spec.state_dim, spec.n_agents
    obs = torch.randn(B, S, device=dev, generator=g)
    nxt = torch.randn(B, S, device=dev, generator=g)
    act = torch.randint(0, 5, (B, A), device=dev, generator=g).float()
    rew = torch.randn(B, A, device=dev, generator=g) * 3
    idx = None
    if explicit_idx:       # a permuted agent-index column exercises the per-sample embedding gather
        idx = torch.stack([torch.randperm(A, device=dev, generator=g) for _ in range(B)]).float()
    out = {"latent": latent, "batch": B, "explicit_idx": bool(explicit_idx)}
    res = {}
    for fusion, m in models.items():
        pb = M.PackedBatch(obs, act, nxt, rew, idx=idx)
        with torch.no_grad():
            rs, rr, mus, lvs = m(pb)
        fw = dict(rs=rs.clone(), rr=rr.clone(), mu=torch.stack(mus).clone(), lv=torch.stack(lvs).clone())
        m.philox_step = 0
        losses = []
        for step in range(2):
            losses.append(m.train_step(M.PackedBatch(obs, act, nxt, rew, idx=idx), 1e-3).clone())
        torch.cuda.synchronize()
        grads = {k: p.grad.clone() for k, p in m.named_arena_tensors().items()}
        params = {k: p.detach().clone() for k, p in m.named_arena_tensors().items()}
        res[fusion] = (fw, losses, grads, params)
    # train_step against the drop-in call sequence (forward -> loss -> backward: recon_s is materialised in fp32 and the
    # loss is its own kernel) on the same model and batch:
    #   "loss" model: the loss epilogue sees the same fp32 values as the separate kernel -> tight agreement
    #   "none" model: train_step keeps recon_s only as bf16 inside the D(recon_s) buffer -> one more bf16 rounding point
    for tag, key in (("lossfuse", "loss"), ("r16", "none")):
        m = models[key]
        m.load_named(res["none"][3])
        m.philox_step = 5
        rs, rr, mus, lvs = m(M.PackedBatch(obs, act, idx=idx))
        loss, sl, rl, kl = M.loss_s_r_vae_fn(rs, rr, nxt, rew, mus, lvs, dev)
        loss.backward()
        torch.cuda.synchronize()
        g_drop = {k: p.grad.clone() for k, p in m.named_arena_tensors().items()}
        l_drop = [float(loss), float(sl), float(rl), float(kl)]
        m.philox_step = 5
        l_fast = [float(x) for x in m.train_step(M.PackedBatch(obs, act, nxt, rew, idx=idx), 0.0).cpu()]
        g_fast = {k: p.grad.clone() for k, p in m.named_arena_tensors().items()}
        out[tag + "_loss_rel"] = max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(l_fast, l_drop))
        lf = {k: rel(g_fast[k], g_drop[k]) for k in g_fast}
        out[tag + "_grad_rel_max"] = max(lf.values())
        out[tag + "_grad_rel_median"] = sorted(lf.values())[len(lf) // 2]
        out[tag + "_grad_worst"] = max(lf, key=lf.get)
    fa, la, ga, pa = res["encoder"]
    fn, ln, gn, pn = res["none"]
    for k in fa:
        out["fwd_" + k] = rel(fa[k], fn[k])
    out["loss_rel"] = [max(abs(float(x) - float(y)) / max(abs(float(y)), 1e-30) for x, y in zip(a, b)) for a, b in zip(la, ln)]
    gr = {k: rel(ga[k], gn[k]) for k in ga}
    worst = sorted(gr.items(), key=lambda kv: -kv[1])[:5]
    out["grad_rel_max"] = worst[0][1]
    out["grad_rel_worst"] = worst
    out["grad_rel_median"] = sorted(gr.values())[len(gr) // 2]
    out["param_rel_max"] = max(rel(pa[k], pn[k]) for k in pa)
    out["finite"] = all(bool(torch.isfinite(v).all()) for v in ga.values())
    print("FUSED_CHECK " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), len(sys.argv) > 3 and sys.argv[3] == "explicit_idx")
