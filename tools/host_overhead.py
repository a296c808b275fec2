#!/usr/bin/env python
"""How much of a train step is host enqueue time?  Runs MAVAE.train_step at the benchmark batch and at a batch so small that
the kernels are negligible: the latter's step time IS the host cost of enqueuing one step (launches + event calls + ctypes).
Prints one JSON line.    python tools/host_overhead.py [--latent 32]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import mfvae_b200 as M                       # noqa: E402
from mfvae_b200.spec import simple_tag_dims  # noqa: E402


def run(B, steps, latent, graph=False):
    dev = "cuda:0"
    spec = simple_tag_dims(latent=latent)
    m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, dev,
                precision="bf16", include_dead_decoder=False)
    g = torch.Generator(device=dev).manual_seed(1)
    S, A = spec.state_dim, spec.n_agents
    pb = M.PackedBatch(torch.randn(B, S, device=dev, generator=g), torch.randint(0, 5, (B, A), device=dev, generator=g).float(),
                       torch.randn(B, S, device=dev, generator=g), torch.randn(B, A, device=dev, generator=g))
    for i in range(10):
        m.train_step(pb, 1e-3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        m.train_step(pb, 1e-3)
    e1.record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    return {"batch": B, "steps": steps, "host_enqueue_ms_per_step": t_enq / steps * 1e3, "wall_ms_per_step": t_all / steps * 1e3,
            "device_ms_per_step": e0.elapsed_time(e1) / steps}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--latent", type=int, default=32)
    a = ap.parse_args()
    print(json.dumps({"big": run(4096, 300, a.latent), "tiny": run(128, 300, a.latent)}))
