#!/usr/bin/env python
"""A handful of cfg2 train steps and nothing else: the target of `ncu` launch lists / full captures.
    python tools/one_step.py [--steps 8] [--batch 4096] [--latent 32]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import mfvae_b200 as M                       # noqa: E402
from mfvae_b200.spec import simple_tag_dims  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--latent", type=int, default=32)
    a = ap.parse_args()
    dev = "cuda:0"
    spec = simple_tag_dims(latent=a.latent)
    m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, dev,
                precision="bf16", include_dead_decoder=False)
    g = torch.Generator(device=dev).manual_seed(1)
    S, A, B = spec.state_dim, spec.n_agents, a.batch
    pbs = [M.PackedBatch(torch.randn(B, S, device=dev, generator=g), torch.randint(0, 5, (B, A), device=dev, generator=g).float(),
                         torch.randn(B, S, device=dev, generator=g), torch.randn(B, A, device=dev, generator=g)) for _ in range(2)]
    torch.cuda.synchronize()
    for i in range(a.steps):
        m.train_step(pbs[i % 2], M.cosine_lr(i))
    torch.cuda.synchronize()
    print("launches", M._lib.lib().mfvae_launch_count())
