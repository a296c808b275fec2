#!/usr/bin/env python
"""Kernel timeline of a few data-parallel train steps (torch.profiler / CUPTI; no nsys in this image).  Run under torchrun;
rank 0 writes gpurun_out/dp_trace.json: [{name, ts_us, dur_us, stream}] for the profiled steps."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import mfvae_b200 as M                      # noqa: E402
from oracle import mavae_oracle as O        # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    spec = O.simple_tag_spec(latent=32)
    B = 4096
    m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, dev,
                precision="bf16", include_dead_decoder=False)
    if world > 1:
        m.enable_data_parallel()
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    S, A = spec.state_dim, spec.n_agents
    pbs = [M.PackedBatch(torch.randn(B, S, device=dev, generator=g), torch.randint(0, 5, (B, A), device=dev, generator=g).float(),
                         torch.randn(B, S, device=dev, generator=g), torch.randn(B, A, device=dev, generator=g),
                         sample0=rank * B, batch_global=world * B) for _ in range(2)]
    for i in range(10):
        m.train_step(pbs[i % 2], 1e-3)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(4):
            m.train_step(pbs[i % 2], 1e-3)
        torch.cuda.synchronize()
    if rank == 0:
        ev = []
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                ev.append({"name": e.name[:60], "ts_us": e.time_range.start, "dur_us": e.time_range.end - e.time_range.start,
                           "stream": getattr(e, "device_resource_id", None)})
        ev.sort(key=lambda d: d["ts_us"])
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(ev, open(os.path.join(ROOT, "gpurun_out", f"dp_trace_n{world}.json"), "w"))
        print("events", len(ev))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
