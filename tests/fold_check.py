"""The folded constant-input column blocks (csrc/fold.cu: action embeddings as one-hot columns against T = W0_act . tables,
agent-id embedding as a per-agent bias of encoder layer 0) against the dense evaluation of the same library
(fusion="nofold") on identical weights, batch and Philox stream.  One JSON line; run in its own process:

    python tests/fold_check.py <fp32|bf16> <batch> [perm_idx]

perm_idx: a per-row permuted agent-index column -- the id-embedding cannot be folded, the action fold still applies."""
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


def main(precision, B, perm_idx=False):
    dev = "cuda:0"
    spec = O.simple_tag_spec(latent=32)
    P = O.init_params(spec, 9)
    g = torch.Generator(device=dev).manual_seed(13)
    S, A = spec.state_dim, spec.n_agents
    obs = torch.randn(B, S, device=dev, generator=g)
    nxt = torch.randn(B, S, device=dev, generator=g)
    act = torch.randint(0, 5, (B, A), device=dev, generator=g).float()
    rew = torch.randn(B, A, device=dev, generator=g) * 3
    idx = torch.stack([torch.randperm(A, device=dev, generator=g) for _ in range(B)]).float() if perm_idx else None
    res = {}
    for fusion in ("auto", "nofold"):
        m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, dev,
                    precision=precision, fusion=fusion, include_dead_decoder=False)
        m.load_named(P)
        with torch.no_grad():
            rs, rr, mus, lvs = m(M.PackedBatch(obs, act, idx=idx))
        fw = dict(rs=rs.clone(), rr=rr.clone(), mu=torch.stack(mus).clone(), lv=torch.stack(lvs).clone())
        m.philox_step = 0
        losses = [m.train_step(M.PackedBatch(obs, act, nxt, rew, idx=idx), 1e-3).clone() for _ in range(2)]
        torch.cuda.synchronize()
        grads = {k: p.grad.clone() for k, p in m.named_arena_tensors().items()}
        params = {k: p.detach().clone() for k, p in m.named_arena_tensors().items()}
        res[fusion] = (fw, losses, grads, params)
    (fa, la, ga, pa), (fn, ln, gn, pn) = res["auto"], res["nofold"]
    out = {"precision": precision, "batch": B, "perm_idx": bool(perm_idx)}
    for k in fa:
        out["fwd_" + k] = rel(fa[k], fn[k])
    out["loss_rel"] = [max(abs(float(x) - float(y)) / max(abs(float(y)), 1e-30) for x, y in zip(a, b)) for a, b in zip(la, ln)]
    gr = {k: rel(ga[k], gn[k]) for k in ga}
    out["grad_rel_max"] = max(gr.values()); out["grad_rel_worst"] = max(gr, key=gr.get)
    out["grad_rel_median"] = sorted(gr.values())[len(gr) // 2]
    a = torch.cat([ga[k].double().reshape(-1) for k in sorted(ga)]); b = torch.cat([gn[k].double().reshape(-1) for k in sorted(ga)])
    out["whole_grad_rel"] = float((a - b).norm() / b.norm()); out["whole_grad_cos"] = float(torch.dot(a, b) / (a.norm() * b.norm()))
    out["param_rel_max"] = max(rel(pa[k], pn[k]) for k in pa)
    out["finite"] = all(bool(torch.isfinite(v).all()) for v in ga.values())
    print("FOLD_CHECK " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), len(sys.argv) > 3 and sys.argv[3] == "perm_idx")
