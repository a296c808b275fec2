#!/usr/bin/env python
"""bench.py — VAE train samples/sec (ELBO fwd + bwd + Adam step) on N B200s, next to the host-CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|ref_dims|wide]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic transitions: device staging + encoder/decoder
forward + reparameterisation + ELBO + backward + (N > 1: gradient all-reduce) + Adam.  Weak scaling: the per-GPU
batch is fixed, `value` is the whole-job samples/s (N * B * K / max-over-ranks device time).

JSON line keys follow the driver contract (metric, value, unit, n_gpus, steps, warmup, ms_per_step, ..., e2e,
gpu_launches, clocks) plus `roofline` (all tcgen05 GEMM launches of the step, timed live with CUDA events on the
launching stream), `cpu_baseline` (the unmodified reference modules from baseline/_ref on this box's host cores; the oracle port
only as the declared fallback) and `eager_b200` (the same reference modules on the same GPU through torch eager).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

METRIC = "VAE train samples/sec (ELBO fwd+bwd+step)"
UNIT = "samples/s"

WORKLOADS = {
    # BASELINE.json configs[1]: "MF-VAE MLP encoder/decoder, batch 4096, latent 32, bf16 on 1 B200"
    "cfg2": dict(latent=32, batch=4096, enc_hidden=(64, 64, 256), dec_hidden=(1024, 256, 64, 256, 1024), precision="bf16",
                 name="cfg2: MF-VAE (simple_tag dims: 40 agents, obs 142/140, 5 actions) batch 4096/GPU, latent 32, bf16"),
    # reference dims (latent 64) at the same batch
    "ref_dims": dict(latent=64, batch=4096, enc_hidden=(64, 64, 256), dec_hidden=(1024, 256, 64, 256, 1024), precision="bf16",
                     name="reference dims (latent 64) batch 4096/GPU, bf16"),
    # BASELINE.json configs[0]: torch_ver/main.py default (B = 128, fp32)
    "cfg1": dict(latent=64, batch=128, enc_hidden=(64, 64, 256), dec_hidden=(1024, 256, 64, 256, 1024), precision="fp32",
                 name="cfg1: torch_ver/main.py default, batch 128, fp32"),
    # BASELINE.json configs[3]: batches sampled from the jax_buffer-style replay buffer, batch 65536 over 8 GPUs.  The ring
    # (HBM-resident, mfvae_ring_*) is sampled inside the timed region of every step: gather + train step.
    "cfg4": dict(latent=64, batch=8192, enc_hidden=(64, 64, 256), dec_hidden=(1024, 256, 64, 256, 1024), precision="bf16", ring=True,
                 name="cfg4: reference dims, batch 8192/GPU sampled every step from the HBM replay ring (65536 global at 8 GPUs), bf16"),
    # BASELINE.json configs[2]: wide MF-VAE
    "wide": dict(latent=128, batch=4096, enc_hidden=(1024, 1024, 1024, 1024), dec_hidden=(1024, 1024, 1024, 1024),
                 precision="bf16", name="cfg3: wide (hidden 1024 x4, latent 128) batch 4096/GPU, bf16"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_numa(local_rank):
    """Pin this rank's host threads to the NUMA node its GPU hangs off BEFORE any pinned buffer is allocated (first touch puts
    the pages there): with 8 ranks feeding 8 GPUs from host memory, remote-node pinned buffers halve the PCIe feed rate."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # nvml pads the PCI domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:
        return {"numa": f"unbound ({type(e).__name__})"}


def make_spec(w):
    from mfvae_b200.spec import simple_tag_dims
    return simple_tag_dims(latent=w["latent"], enc_hidden=tuple(w["enc_hidden"]), dec_hidden=tuple(w["dec_hidden"]))


def make_config(w, spec, B, world, source):
    """`config` of the JSON line: identical for both arms (the driver compares them)."""
    from mfvae_b200.spec import flops_per_sample, executed_macs_per_sample
    row = 2 * spec.state_dim + 2 * spec.n_agents
    return {"workload": w["name"], "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"dp{world}",
            "source": source, "l2": f"inputs {B * row * 4 / 1e6:.0f} MB/step, 4 rotating device batches (> 126 MB L2)",
            "flop_per_sample": flops_per_sample(spec), "flop_per_sample_executed": 6 * executed_macs_per_sample(spec)}


def synth_transition(spec, batch, seed=0):
    """cpprb.sample-shaped dict (torch_ver/src/replay_buffer.py:62-81,107-108): float32 arrays
    {agent}_{observations,next_observations,actions,rewards} of shape (B, dim)."""
    rng = np.random.default_rng(seed)
    t = {}
    for a in spec.agents:
        o = spec.obs_dim[a]
        t[f"{a}_observations"] = rng.standard_normal((batch, o)).astype(np.float32)
        t[f"{a}_next_observations"] = rng.standard_normal((batch, o)).astype(np.float32)
        t[f"{a}_actions"] = rng.integers(0, spec.n_act[a], size=(batch, 1)).astype(np.float32)
        t[f"{a}_rewards"] = rng.standard_normal((batch, 1)).astype(np.float32)
    return t


# ------------------------------------------------------------------------------------------------
# reference arm: the UNMODIFIED reference modules (baseline/_ref/torch_ver/{model,trainer}.py, installed byte-for-byte by
# __graft_entry__.build() from /root/reference) driven exactly as torch_ver/main.py:84-98 drives them.  The oracle port is
# the declared fallback when the install is absent or the workload is not expressible with the reference's hard-coded
# layer widths (model.py:46,87).
# ------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref", "torch_ver")


def reference_modules():
    if not (os.path.exists(os.path.join(REF_DIR, "model.py")) and os.path.exists(os.path.join(REF_DIR, "trainer.py"))):
        return None
    import importlib.util
    mods = []
    for name in ("model", "trainer"):
        sp = importlib.util.spec_from_file_location(f"mfvae_reference_{name}", os.path.join(REF_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(sp)
        sp.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


def reference_expressible(w):
    from mfvae_b200.spec import ENC_HIDDEN, DEC_HIDDEN
    return tuple(w["enc_hidden"]) == ENC_HIDDEN and tuple(w["dec_hidden"]) == DEC_HIDDEN


def reference_step_factory(w, batch, device="cpu"):
    """One train step as the reference driver runs it (torch_ver/main.py:84-98): sample -> create_dataset -> MAVAE.forward
    -> loss_s_r_vae_fn -> zero_grad -> backward -> Adam.step -> CosineAnnealingLR.step.  `sample` is a pre-drawn
    cpprb-shaped dict (cpprb is not installed; sampling is not part of the timed arithmetic)."""
    import contextlib
    mods = reference_modules()
    assert mods is not None
    ref_model, ref_trainer = mods
    spec = make_spec(w)
    torch.manual_seed(0)
    m = ref_model.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, device)
    m.to(device)                                                                      # main.py:49
    opt = torch.optim.Adam(m.parameters(), 0.005)                                     # main.py:52
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=50, eta_min=1e-4)   # main.py:53
    codebook = {a: i for i, a in enumerate(spec.agents)}                              # main.py:39-42
    trans = synth_transition(spec, batch, seed=0)
    devnull = open(os.devnull, "w")

    def step():
        with contextlib.redirect_stdout(devnull):                                     # debug prints at model.py:160-163
            idx_state, acts, joint, nxt, rew = ref_trainer.create_dataset(trans, codebook)
            rs, rr, mus, lvs = m(idx_state, acts)
            loss, sl, rl, kl = ref_model.loss_s_r_vae_fn(rs, rr, nxt, rew, mus, lvs, device)
            opt.zero_grad()
            loss.backward()
            opt.step()
            sched.step()
        return loss
    return step


def port_step_factory(w, batch):
    from oracle import mavae_oracle as O
    spec = O.simple_tag_spec(latent=w["latent"], enc_hidden=tuple(w["enc_hidden"]), dec_hidden=tuple(w["dec_hidden"]))
    st = O.OracleState(spec, O.init_params(spec, 0))
    codebook = {a: i for i, a in enumerate(spec.agents)}
    trans = O.synth_transition(spec, batch, seed=0)
    eps_all = torch.randn(batch, spec.n_agents * spec.latent)
    L = spec.latent
    counter = [0]

    def step():
        # same work the reference does per step (main.py:84-98): staging, forward, loss, backward, Adam
        idx_state, acts, joint, nxt, rew = O.stage_batch(trans, codebook)
        eps = {a: eps_all[:, i * L:(i + 1) * L] for i, a in enumerate(spec.agents)}
        losses, _, _ = O.train_step(st, idx_state, acts, eps, nxt, rew, O.cosine_lr(counter[0]))
        counter[0] += 1
        return losses
    return step


def time_cpu(w, batch, steps, warmup):
    """(samples/s, ms/step, kind) of the reference's CPU train step on all host cores."""
    torch.set_num_threads(os.cpu_count() or 1)
    if reference_modules() is not None and reference_expressible(w):
        step, kind = reference_step_factory(w, batch), "reference"
    else:
        step, kind = port_step_factory(w, batch), "port"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, kind


def time_eager_gpu(w, batch, dev, steps=5, warmup=2):
    """The reference modules, unmodified, on the same B200 through torch eager (cuBLAS + ATen + torch.optim.Adam): the
    "existing Blackwell path" of SURVEY.md section 2.1.  None when the reference install is absent."""
    if reference_modules() is None or not reference_expressible(w):
        return None
    step = reference_step_factory(w, batch, device=dev)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": steps,
            "what": "unmodified reference model.py/trainer.py (baseline/_ref) on torch eager, fp32, same GPU, host batch as in main.py:84-98"}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = make_spec(w)
    batch = args.batch or w["batch"]
    steps, warm = max(1, args.steps), max(0, args.warmup)
    sps, ms, kind = time_cpu(w, batch, steps, warm)
    cores = torch.get_num_threads()
    what = ("unmodified reference torch_ver/model.py + trainer.py (baseline/_ref), driven as main.py:84-98 incl. create_dataset"
            if kind == "reference" else "oracle port of the torch_ver step incl. create_dataset (reference install absent or widths not expressible)")
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": make_config(w, spec, batch, max(1, args.gpus), "device-resident batches"),
            "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{steps} full train steps (+{warm} warm-up) of batch {batch}: {what}, torch CPU fp32, {cores} threads"},
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def synth_device_batches(spec, B, n, device, seed):
    import mfvae_b200 as M
    g = torch.Generator(device=device).manual_seed(seed)
    S, A = spec.state_dim, spec.n_agents
    out = []
    for _ in range(n):
        obs = torch.randn(B, S, device=device, generator=g)
        nxt = torch.randn(B, S, device=device, generator=g)
        act = torch.randint(0, 5, (B, A), device=device, generator=g).float()
        rew = torch.randn(B, A, device=device, generator=g)
        out.append(M.PackedBatch(obs, act, nxt, rew))
    return out


def flops_per_sample(spec):
    from mfvae_b200.spec import flops_per_sample as f
    return f(spec)


def run_ours(args, w):
    import torch.distributed as dist
    import mfvae_b200 as M
    from mfvae_b200 import _lib as L
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device: there is no CPU path")
    numa = bind_numa(local)
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    spec = make_spec(w)
    B = args.batch or w["batch"]
    m = M.MAVAE(spec.idx_features, spec.latent, spec.act_features, True, spec.agents, spec.obs_dim, spec.n_act, dev,
                precision=w["precision"], enc_hidden=w["enc_hidden"], dec_hidden=w["dec_hidden"], include_dead_decoder=False,
                fusion=os.environ.get("MFVAE_FUSION", "auto"))
    torch.manual_seed(0)
    m.reset_parameters()
    if world > 1:
        m.enable_data_parallel()
    nb = 4
    batches = synth_device_batches(spec, B, nb, dev, seed=1234 + rank)
    for i, pb in enumerate(batches):
        pb.sample0 = rank * B
        pb.batch_global = world * B
    lr = M.cosine_lr
    ring = None
    if w.get("ring"):
        # this rank's shard of the replay ring: 4 batches of synthetic transitions in the ring's row layout
        # [obs | act | next | rew | done]; every step draws B rows uniformly with replacement (Philox) and gathers them
        from mfvae_b200.replay_buffer import DeviceRing
        ring = DeviceRing(spec.agents, spec.obs_dim, capacity=nb * B, device=dev)
        for pb in batches:
            rows = torch.zeros(B, ring.row, device=dev)
            rows[:, :2 * ring.S + 2 * ring.A] = torch.cat([pb.obs, pb.act, pb.next, pb.rew], dim=1)
            ring.add_rows_device(rows)
        torch.cuda.synchronize()
        batches = None

    def get_batch(i):
        if ring is None:
            return batches[i % nb]
        return ring.sample_packed(B, seed=1234 + rank, sample0=rank * B, batch_global=world * B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (`value`) ----
    for i in range(args.warmup):
        m.train_step(get_batch(i), lr(i), pipeline=True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.lib().mfvae_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        # pipeline: the optimizer sweep of the decoder block may overlap the next step's encoder half (data parallel only;
        # everything has completed when the closing barrier + synchronize returns)
        losses = m.train_step(get_batch(i), lr(args.warmup + i), pipeline=True)
    e1.record()
    barrier()
    launches = L.lib().mfvae_launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    # the timed region can be shorter than nvidia-smi's sampling period: keep the same load running (untimed) for ~1.2 s
    # so the clocks / throttle reasons describe this workload under load.  The step contains collectives when N > 1, so
    # EVERY rank runs the same number of extra steps (derived from the all-reduced step time).
    n_extra = int(min(4000, max(20, 1.2e3 / max(ms_total / args.steps, 1e-3))))
    for j in range(n_extra):
        m.train_step(get_batch(j), lr(j))
        if j % 50 == 49:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    value = world * B * args.steps / (ms_total * 1e-3)
    loss_host = [float(x) for x in losses.cpu()]

    # ---- end to end: packed pinned host batch -> H2D -> step -> D2H loss, every step ----
    # Host rows as a pinned replay ring keeps them: [obs | act | next | rew].  Observations are kept in bf16 on the host (the
    # bf16 engine rounds them on arrival anyway: the step is bit-identical to shipping fp32), everything else fp32.  The
    # `e2e_bf16_targets` variant also keeps next-observations (the reconstruction TARGET) in bf16 -- that rounds the loss'
    # target, so it is reported beside the headline, not as it.
    S, A = spec.state_dim, spec.n_agents
    row = 2 * S + 2 * A
    n_host = 3
    e2e_steps = max(3, min(args.steps, 20))
    loss_pinned = torch.empty(4).pin_memory()
    copy_stream = torch.cuda.Stream(dev)

    def e2e_run(obs_bf16, next_bf16):
        ob, nb_ = (2 if obs_bf16 else 4), (2 if next_bf16 else 4)
        seg = [B * S * ob, B * A * 4, B * S * nb_, B * A * 4]                    # bytes: obs | act | next | rew
        offs = [0, seg[0], seg[0] + seg[1], seg[0] + seg[1] + seg[2]]
        total = sum(seg)
        host = []
        for _ in range(n_host):
            hb = torch.empty(total, dtype=torch.uint8).pin_memory()
            hb[offs[0]:offs[1]].view(torch.bfloat16 if obs_bf16 else torch.float32).copy_(torch.randn(B * S))
            hb[offs[1]:offs[2]].view(torch.float32).copy_(torch.randint(0, 5, (B * A,)).float())
            hb[offs[2]:offs[3]].view(torch.bfloat16 if next_bf16 else torch.float32).copy_(torch.randn(B * S))
            hb[offs[3]:].view(torch.float32).copy_(torch.randn(B * A))
            host.append(hb)
        dbuf = [torch.empty(total, dtype=torch.uint8, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def views(d):
            return (d[offs[0]:offs[1]].view(torch.bfloat16 if obs_bf16 else torch.float32).view(B, S), d[offs[1]:offs[2]].view(torch.float32).view(B, A),
                    d[offs[2]:offs[3]].view(torch.bfloat16 if next_bf16 else torch.float32).view(B, S), d[offs[3]:].view(torch.float32).view(B, A))

        def loop(n):
            main = torch.cuda.current_stream()
            for k in range(2):
                freed[k].record(main)
            with torch.cuda.stream(copy_stream):      # prologue: copy of step 0
                dbuf[0].copy_(host[0], non_blocking=True); ready[0].record(copy_stream)
            for i in range(n):
                k = i % 2
                if i + 1 < n:     # overlap the next batch's H2D with this step's compute
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(freed[(i + 1) % 2])
                        dbuf[(i + 1) % 2].copy_(host[(i + 1) % n_host], non_blocking=True); ready[(i + 1) % 2].record(copy_stream)
                main.wait_event(ready[k])
                obs, act, nxt, rew = views(dbuf[k])
                pb = M.PackedBatch(obs, act, nxt, rew, sample0=rank * B, batch_global=world * B)
                out = m.train_step(pb, lr(i))
                freed[k].record(main)
                loss_pinned.copy_(out, non_blocking=True)
            main.synchronize()

        loop(3)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        t0.record()
        loop(e2e_steps)
        t1.record()
        barrier()
        wall = time.perf_counter() - wall0
        ms2 = torch.tensor([max(t0.elapsed_time(t1), 0.0)], device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        return world * B * e2e_steps / (float(ms2) * 1e-3), total, wall

    e2e_value, e2e_bytes, wall = e2e_run(w["precision"] == "bf16", False)
    e2e_fp32 = e2e_run(False, False) if w["precision"] == "bf16" else None
    e2e_t16 = e2e_run(True, True) if w["precision"] == "bf16" else None

    # ---- roofline of the tensor-core GEMMs, timed live with CUDA events on the launching stream ----
    roof = None
    pk = peaks()
    L.check(L.lib().mfvae_profile_enable(m._h, 1))        # all ranks: the step holds collectives when N > 1
    tim = (L.MfvaeGemmTiming * 256)()
    acc = {}
    reps = max(3, min(args.steps, 10))
    for i in range(reps):
        m.train_step(get_batch(i), lr(i))
        n = L.lib().mfvae_profile_read(m._h, tim, 256)
        for j in range(n):
            t = tim[j]
            key = (t.kind, t.groups, t.M, t.N, t.K, j)
            acc.setdefault(key, []).append(t.ms)
    L.check(L.lib().mfvae_profile_enable(m._h, 0))
    barrier()
    if rank == 0:
        tot_ms, tot_fl, rows = 0.0, 0.0, []
        for (kind, G, Mm, Nn, Kk, j), v in acc.items():
            msj = float(np.mean(v)); fl = 2.0 * G * Mm * Nn * Kk
            tot_ms += msj; tot_fl += fl
            rows.append({"kind": ["fwd", "dgrad", "wgrad", "fused encoder chain (fwd, 4 layers + reparam + KL)"][kind], "G": G, "M": Mm, "N": Nn, "K": Kk,
                         "ms": round(msj, 4), "tflops": round(fl / (msj * 1e-3) / 1e12, 1) if msj > 0 else None})
        rows.sort(key=lambda r: -r["ms"])
        if tot_ms > 0:
            ach = tot_fl / (tot_ms * 1e-3) / 1e12
            # the same launches against the REFERENCE's dense flop count (6 flops per MAC of its nn.Linear layers):
            # the folded column blocks (csrc/fold.cu) do the same mathematics with 30 % fewer executed flops
            alg_fl = float(flops_per_sample(spec)) * B
            alg = alg_fl / (tot_ms * 1e-3) / 1e12
            bound = "tensor" if w["precision"] == "bf16" else "fp32-simt"
            # DRAM traffic of the same launches: dram__bytes_read.sum + dram__bytes_write.sum from the committed ncu launch list
            # (profiles/r2_traffic.json); only valid for the workload / batch it was captured on
            traffic = None
            tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
            if args.workload == "cfg2" and B == 4096 and os.path.exists(tp):
                traffic = json.load(open(tp))["cfg2_b4096"]["tensor_core_kernels_all_launches"]["dram_bytes_per_step"]
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "achieved_vs_reference_flop_count": alg, "frac_vs_reference_flop_count": alg / pk["tf_sustained"],
                    "traffic": traffic, "traffic_note": "bytes per step over the same launches (ncu, profiles/r2_traffic.json)",
                    "kernel": "gemm_tc_kernel + enc_fwd_kernel (all %d tensor-core launches of one step)" % len(rows),
                    "gemm_ms_per_step": tot_ms, "gemm_flop_per_step": tot_fl, "reference_flop_per_step": alg_fl,
                    "peak_source": pk["src"] + " (sustained cuBLAS bf16)",
                    "engine": bound, "top": rows[:8], "all": rows}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on this box's host cores, bounded sample ----
    cpu = eager = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = min(B, 4096)
        sps, msc, kind = time_cpu(w, cb, 3, 1)
        cpu = {"value": sps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"3 train steps (+1 warm-up) of batch {cb}, same model, " +
                         ("unmodified reference modules (baseline/_ref) driven as main.py:84-98" if kind == "reference"
                          else "oracle port of the reference step") + " incl. create_dataset, torch CPU fp32",
               "ms_per_step": msc}
        eager = time_eager_gpu(w, cb, dev)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": w["precision"], "data": "synthetic",
                "config": make_config(w, spec, B, world, "replay ring (device gather every step)" if ring is not None else "device-resident batches"),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_bytes, "d2h_bytes_per_step": 16,
                        "steps": e2e_steps, "wall_s": wall,
                        "api": "MAVAE.train_step(PackedBatch) fed from packed pinned host rows [obs bf16 | act | next | rew fp32] (bf16 engine: "
                               "bit-identical to fp32 observations)" if w["precision"] == "bf16" else "MAVAE.train_step(PackedBatch) fed from packed pinned fp32 host rows"},
                "e2e_fp32_rows": None if e2e_fp32 is None else {"value": e2e_fp32[0], "unit": UNIT, "h2d_bytes_per_step": e2e_fp32[1]},
                "e2e_bf16_targets": None if e2e_t16 is None else {"value": e2e_t16[0], "unit": UNIT, "h2d_bytes_per_step": e2e_t16[1],
                                                                  "note": "next-observation targets also bf16 on the host: rounds the loss target (opt-in)"},
                "comm": getattr(m, "comm_info", None), "host_numa": numa,
                "gpu_launches": int(launches), "clocks": clocks, "losses_last_step": loss_host,
                "model_tflops": value * flops_per_sample(spec) / 1e12 / world,
                "roofline": roof, "cpu_baseline": cpu, "eager_b200": eager}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        run_ours(args, w)


if __name__ == "__main__":
    main()
