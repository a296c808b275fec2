"""Data-parallel train step on real GPUs (NCCL): run under torchrun with WORLD_SIZE ranks, one GPU each.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/dp_gpu_check.py <fp32|bf16>

Every rank holds the same parameters and takes rows [rank * Bl, (rank + 1) * Bl) of one global batch; after two
train steps (bucketed all-reduce overlapped with backward, fused Adam) rank 0 compares losses, gradients and parameters
with a single-process run of the FULL batch on its own GPU.  Philox eps is keyed by the global sample index, so both runs
see identical noise.  One line: DP_CHECK {json}."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.dont_write_bytecode = True
from oracle import mavae_oracle as O      # noqa: E402
import mfvae_b200 as M                    # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def main(precision):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    spec = O.simple_tag_spec(latent=32)
    Bl = 256
    Bg = Bl * world
    P = O.init_params(spec, 3)
    g = torch.Generator().manual_seed(5)
    S, A = spec.state_dim, spec.n_agents
    obs, nxt = torch.randn(Bg, S, generator=g), torch.randn(Bg, S, generator=g)
    act, rew = torch.randint(0, 5, (Bg, A), generator=g).float(), torch.randn(Bg, A, generator=g) * 3

    def make():
        m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, dev,
                    precision=precision, include_dead_decoder=False)
        m.load_named(P)
        return m

    def batch(lo, hi, sample0):
        return M.PackedBatch(obs[lo:hi].to(dev), act[lo:hi].to(dev), nxt[lo:hi].to(dev), rew[lo:hi].to(dev),
                             sample0=sample0, batch_global=Bg)

    m = make()
    m.enable_data_parallel()
    losses, g1 = [], None
    for step in range(2):
        losses.append(m.train_step(batch(rank * Bl, (rank + 1) * Bl, rank * Bl), 1e-3).clone())
        if step == 0:
            torch.cuda.synchronize()
            g1 = {k: p.grad.clone() for k, p in m.named_arena_tensors().items()}
    torch.cuda.synchronize()
    out = {"precision": precision, "world": world, "batch_global": Bg, "comm": m.comm_info}
    if rank == 0:
        ref = make()                       # single process, full batch
        rl, r1 = [], None
        for step in range(2):
            rl.append(ref.train_step(batch(0, Bg, 0), 1e-3).clone())
            if step == 0:
                torch.cuda.synchronize()
                r1 = {k: p.grad.clone() for k, p in ref.named_arena_tensors().items()}
        torch.cuda.synchronize()
        out["loss_rel"] = max(abs(float(a) - float(b)) / max(abs(float(b)), 1e-30) for x, y in zip(losses, rl) for a, b in zip(x, y))
        mine, theirs = m.named_arena_tensors(), ref.named_arena_tensors()
        reg = [k for k in mine if k.startswith(("state_decoder", "reward_decoder", "reward_linear", "idx_emb"))]
        gr = {k: rel(mine[k].grad, theirs[k].grad) for k in reg}        # all-reduced (optimised) tensors
        out["grad_rel_max"] = max(gr.values()); out["grad_rel_worst"] = max(gr, key=gr.get)
        out["grad_rel_max_step1"] = max(rel(g1[k], r1[k]) for k in reg)  # before any parameter has moved
        out["param_rel_max"] = max(rel(mine[k], theirs[k]) for k in reg)
    # the reference's own call sequence (main.py:87-97) under data parallel, on both loss routes: the fused CUDA ELBO
    # (global means; loss values all-reduced before they are returned) and a torch loss over the autograd bridge (local
    # means -> seeds scaled by 1 / world).  Gradients after the all-reduce must equal the full-batch gradients.
    F = torch.nn.functional

    def dropin(model, pb, route):
        model.philox_step = 11
        rs, rr, mus, lvs = model(M.PackedBatch(pb.obs, pb.act, sample0=pb.sample0, batch_global=pb.batch_global))
        if route == "fused":
            loss, sl, rl_, kl = M.loss_s_r_vae_fn(rs, rr, pb.next, pb.rew, mus, lvs, dev)
        else:
            kl = sum(torch.mean(-0.5 * torch.sum(1 + lv - mu ** 2 - torch.exp(lv), 1), 0) for mu, lv in zip(mus, lvs))
            loss = F.huber_loss(pb.next, rs) + 0.005 * F.huber_loss(pb.rew, rr) + 0.0025 * kl
        loss.backward()
        model.grads_consumed()
        torch.cuda.synchronize()
        return float(loss.detach())

    for route in ("fused", "torch"):
        lv_dp = dropin(m, batch(rank * Bl, (rank + 1) * Bl, rank * Bl), route)
        if rank == 0:
            ref.load_named(m.named_arena_tensors())
            lv_ref = dropin(ref, batch(0, Bg, 0), route)
            mine, theirs = m.named_arena_tensors(), ref.named_arena_tensors()
            gr = {k: rel(mine[k].grad, theirs[k].grad) for k in reg}
            out[f"dropin_{route}_grad_rel_max"] = max(gr.values())
            if route == "fused":
                out["dropin_fused_loss_rel"] = abs(lv_dp - lv_ref) / max(abs(lv_ref), 1e-30)
    # the single C entry (mfvae_train_step: exchange on the library's own communication stream) against the call-by-call step
    m2 = make(); m2.enable_data_parallel()
    m3 = make(); m3.enable_data_parallel()
    for step in range(2):
        l2 = m2.train_step(batch(rank * Bl, (rank + 1) * Bl, rank * Bl), 1e-3).clone()
        l3 = m3.train_step_c(batch(rank * Bl, (rank + 1) * Bl, rank * Bl), 1e-3).clone()
    torch.cuda.synchronize()
    if rank == 0:
        out["single_call_loss_rel"] = max(abs(float(x) - float(y)) / max(abs(float(y)), 1e-30) for x, y in zip(l3, l2))
        out["single_call_param_rel"] = rel(m3._arena[:m3._n_opt], m2._arena[:m2._n_opt])
    if rank == 0:
        print("DP_CHECK " + json.dumps(out), flush=True)
    # every rank must hold identical parameters after the step
    flat = m._arena[:m._n_opt].clone()
    ref0 = flat.clone()
    dist.broadcast(ref0, 0)
    same = torch.tensor([float(torch.equal(flat, ref0))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_SYNC " + json.dumps({"identical_params_on_all_ranks": bool(same.item() == 1.0)}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "fp32")
