"""CPU, world_size = 2, gloo: the data-parallel host logic — bucket ranges, SUM all-reduce of the gradient arena,
shard bookkeeping (sample0 / batch_global) — and, with the oracle, the design claim that per-rank gradients scaled
by 1/B_global add up to the full-batch gradient and that every rank draws its slice of the global Philox eps."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mavae_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mfvae_b200 as M
        spec = O.tiny_spec(3, idx_features=16, latent=8, act_features=8)
        torch.manual_seed(100 + rank)                       # replicas start DIFFERENT: torch-default init per rank
        m = M.MAVAE(16, 8, 8, True, spec.agents, spec.obs_dim, spec.n_act, "cpu", precision="fp32")
        m._m.fill_(float(rank)); m._adam_t = 3 + rank; m.philox_step = 10 * (rank + 1)
        m.enable_data_parallel()
        assert m.data_parallel
        # (0) enable_data_parallel broadcasts rank 0's parameters, Adam moments / step and Philox position
        ref = m._arena.clone(); dist.broadcast(ref, 0)
        ok0 = torch.equal(ref, m._arena) and float(m._m.abs().max()) == 0.0 and (m._adam_t, m.philox_step) == (3, 10)
        # the autograd-bridge path scales its seeds by 1 / world (a torch loss is a LOCAL-batch mean, the exchange a sum)
        ok0 = ok0 and m._world() == world
        # (1) arena all-reduce over the buckets: every optimised element is summed exactly once
        m._grad.fill_(0.0)
        m._grad[:m._n_opt] = float(rank + 1)
        m._grad[m._n_opt:] = 100.0 * (rank + 1)            # encoders: not optimised -> not communicated
        m._losses = torch.tensor([1.0, 2.0, 3.0, 4.0]) * (rank + 1)
        m._allreduce_grads()
        ok1 = bool((m._grad[:m._n_opt] == 3.0).all()) and bool((m._grad[m._n_opt:] == 100.0 * (rank + 1)).all())
        ok1 = ok1 and torch.equal(m._losses, torch.tensor([3.0, 6.0, 9.0, 12.0]))

        # (2) sharded oracle gradients (local mean * B_local / B_global) summed over ranks == full-batch gradients
        Bg = 16; Bl = Bg // world
        P = O.init_params(spec, 1)
        trans = O.synth_transition(spec, Bg, seed=2)
        cb = {a: i for i, a in enumerate(spec.agents)}
        Lt = spec.latent
        eps_g = torch.from_numpy(O.philox_normal(0x5EED, 0, 0, Bg, spec.n_agents * Lt).astype(np.float32))
        eps_l = torch.from_numpy(O.philox_normal(0x5EED, 0, rank * Bl, Bl, spec.n_agents * Lt).astype(np.float32))
        ok2 = torch.equal(eps_l, eps_g[rank * Bl:(rank + 1) * Bl])       # rank draws its slice of the global stream
        shard = {k: v[rank * Bl:(rank + 1) * Bl] for k, v in trans.items()}
        idx_state, acts, _, nxt, rew = O.stage_batch(shard, cb)
        eps = {a: eps_l[:, i * Lt:(i + 1) * Lt] for i, a in enumerate(spec.agents)}
        losses, G, _ = O.grads(P, spec, idx_state, acts, eps, nxt, rew)
        flat = torch.cat([G[k].reshape(-1) for k in sorted(G)]) * (Bl / Bg)
        lvec = torch.tensor(losses, dtype=torch.float64) * (Bl / Bg)
        dist.all_reduce(flat); dist.all_reduce(lvec)
        idx_state, acts, _, nxt, rew = O.stage_batch(trans, cb)
        epsg = {a: eps_g[:, i * Lt:(i + 1) * Lt] for i, a in enumerate(spec.agents)}
        losses_full, Gf, _ = O.grads(P, spec, idx_state, acts, epsg, nxt, rew)
        flat_full = torch.cat([Gf[k].reshape(-1) for k in sorted(Gf)])
        ok3 = float((flat - flat_full).abs().max()) <= 2e-6 * float(flat_full.abs().max())
        ok3 = ok3 and max(abs(float(a) - b) / abs(b) for a, b in zip(lvec, losses_full)) < 1e-5
        q.put((rank, ok0 and ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, ok2, ok3 in res:
        assert ok1, f"rank {rank}: state broadcast / bucketed all-reduce"
        assert ok2, f"rank {rank}: Philox shard"
        assert ok3, f"rank {rank}: sharded gradients do not add up to the full-batch gradient"
