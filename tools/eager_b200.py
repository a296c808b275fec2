#!/usr/bin/env python
"""Comparator: the reference's op structure (per-agent Python loop of F.linear / relu / embedding, torch autograd, per-tensor
Adam: SURVEY.md 2.1 "the existing Blackwell path is PyTorch eager") executed by torch eager ON THE B200, same model and
batch as bench.py's default workload.  The Python reference itself is not on the GPU box, so this runs the oracle's
restatement of it with every tensor on cuda:0 (fp32, cuBLAS + ATen kernels).  Measurement tooling only: prints one JSON line.

    python tools/eager_b200.py [--batch 4096] [--latent 32] [--steps 10]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
from oracle import mavae_oracle as O    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--latent", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    dev = "cuda:0"
    spec = O.simple_tag_spec(latent=args.latent)
    st = O.OracleState(spec, {k: v.to(dev) for k, v in O.init_params(spec, 0).items()})
    cb = {a: i for i, a in enumerate(spec.agents)}
    trans = O.synth_transition(spec, args.batch, seed=0)
    idx_state, acts, _, nxt, rew = O.stage_batch(trans, cb)
    idx_state = {a: t.to(dev) for a, t in idx_state.items()}
    acts = {a: t.to(dev) for a, t in acts.items()}
    nxt, rew = nxt.to(dev), rew.to(dev)
    L = spec.latent

    def step(i):
        eps_all = torch.randn(args.batch, spec.n_agents * L, device=dev)
        eps = {a: eps_all[:, k * L:(k + 1) * L] for k, a in enumerate(spec.agents)}
        return O.train_step(st, idx_state, acts, eps, nxt, rew, O.cosine_lr(i))

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    print(json.dumps({"comparator": "torch eager on B200, reference op structure (oracle restatement), fp32, batches resident on the device",
                      "batch": args.batch, "latent": args.latent, "ms_per_step": dt * 1e3, "samples_per_s": args.batch / dt,
                      "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
