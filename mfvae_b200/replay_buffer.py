"""HBM-resident replay ring behind the reference's two buffer wrappers.

* ``MultiAgentCPPRB`` mirrors ``/root/reference/torch_ver/src/replay_buffer.py:53-115`` (a wrapper over the C++
  ``cpprb.ReplayBuffer``): ``add(observations, next_observations, actions, rewards, terminals, trunctions)``,
  ``on_episode_end()``, ``sample() -> dict of (B, dim) float32 numpy arrays`` keyed
  ``{agent}_{observations,next_observations,actions,rewards,terminals,truncations}`` + ``mask``, and the iterator
  protocol.
* ``JaxFbxBuffer`` mirrors ``/root/reference/jax_ver/jax_buffer.py:80-140`` (flashbax item buffer):
  ``init_buffer / add_trans(obs, reward, actions, next_obs, done)``, ``can_sample()``, ``sample(rng_key)`` whose
  ``.experience[{agent}_{obs,act,next_obs,rew} | done]`` arrays have shape ``(B, dim, 1)`` (jax_buffer.py:186-188).

Both store one joint transition per ring row ``[obs(S) | act(W) | next_obs(S) | rew(A) | terminals(A) | truncations(A) |
mask | pad]`` in device memory (``mfvae_ring_*`` in ``include/mfvae.h``): every key of the reference's cpprb ``env_dict``
(replay_buffer.py:62-81), per agent, with ``W`` = the agents' action widths summed (1 each for a float-coded discrete action,
the Box width for a continuous one, as ``get_space_shape`` gives them).  ``add`` goes through a pinned host staging block that is flushed with one
asynchronous H2D copy; ``sample_packed()`` gathers a uniform-with-replacement batch on the device straight into the
``PackedBatch`` the train step consumes, so no host staging is on the training path.  The dict-returning ``sample()``
of the reference contract is kept (device gather + one D2H) for drop-in use with ``create_dataset``.

``sample()`` hands back each agent's own ``terminals`` / ``truncations`` and the ``mask`` exactly as they were added
(cpprb semantics); the jax-style wrapper stores its single ``done = any(agent done)`` (jax_buffer.py:37,49-52) in the
``mask`` slot.
"""
import ctypes as C
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .model import PackedBatch


def _space_dim(space) -> int:
    """Width a gymnasium-like space occupies in the buffer (reference get_space_shape, replay_buffer.py:35-50):
    Discrete -> 1, MultiBinary -> n, Box -> prod(shape); anything else raises NotImplementedError like the reference."""
    kind = type(space).__name__
    if kind == "Discrete":
        return 1
    if kind == "MultiBinary":
        return int(np.prod(space.n))
    shape = getattr(space, "shape", None)
    if shape:
        return int(np.prod(shape))
    raise NotImplementedError(f"unsupported space {space!r}")


class DeviceRing:
    """The ring itself: storage tensor + pinned staging + the C handle."""

    def __init__(self, agents: Sequence[str], obs_dim: Dict[str, int], capacity: int, device="cuda:0", stage_rows: int = 256,
                 act_dim: Optional[Dict[str, int]] = None):
        self.agents = list(agents)
        self.obs_dim = {a: int(obs_dim[a]) for a in self.agents}
        self.act_dim = {a: int(act_dim[a]) if act_dim else 1 for a in self.agents}
        self.S = sum(self.obs_dim.values())
        self.A = len(self.agents)
        self.W = sum(self.act_dim.values())
        self.capacity = int(capacity)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mfvae_b200: the replay ring lives in GPU memory; there is no CPU fallback")
        lib = L.lib()
        self.row = int(lib.mfvae_ring_row_floats(self.S, self.A, self.W))
        self.storage = torch.zeros(self.capacity * self.row, dtype=torch.float32, device=self.device)
        self._h = C.c_void_p()
        L.check(lib.mfvae_ring_create(self.S, self.A, self.W, self.capacity, L.ptr(self.storage), C.byref(self._h)))
        self._stage = torch.zeros(stage_rows, self.row, dtype=torch.float32).pin_memory()
        self._n_staged = 0
        self._step = 0
        self._offsets = np.cumsum([0] + [self.obs_dim[a] for a in self.agents])
        self._act_off = np.cumsum([0] + [self.act_dim[a] for a in self.agents])
        # column offsets inside a row
        self.c_act, self.c_next = self.S, self.S + self.W
        self.c_rew = 2 * self.S + self.W
        self.c_term, self.c_trunc, self.c_mask = self.c_rew + self.A, self.c_rew + 2 * self.A, self.c_rew + 3 * self.A

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                L.lib().mfvae_ring_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def __len__(self):
        return int(L.lib().mfvae_ring_size(self._h)) + self._n_staged

    def add_row(self, obs: Dict, act: Dict, next_obs: Dict, rew: Dict, terminals=None, truncations=None, mask: float = 0.0):
        """One joint transition.  ``terminals`` / ``truncations``: per-agent dicts (or None = zeros)."""
        if self._n_staged == self._stage.shape[0]:
            self.flush()
        r = self._stage[self._n_staged].numpy()
        for i, a in enumerate(self.agents):
            o0, o1 = self._offsets[i], self._offsets[i + 1]
            r[o0:o1] = np.asarray(obs[a], dtype=np.float32).reshape(-1)
            r[self.c_next + o0:self.c_next + o1] = np.asarray(next_obs[a], dtype=np.float32).reshape(-1)
            r[self.c_act + self._act_off[i]:self.c_act + self._act_off[i + 1]] = np.asarray(act[a], dtype=np.float32).reshape(-1)
            r[self.c_rew + i] = float(np.asarray(rew[a]).reshape(-1)[0])
            r[self.c_term + i] = float(np.asarray(terminals[a]).reshape(-1)[0]) if terminals is not None else 0.0
            r[self.c_trunc + i] = float(np.asarray(truncations[a]).reshape(-1)[0]) if truncations is not None else 0.0
        r[self.c_mask] = float(mask)
        self._n_staged += 1

    def flush(self):
        if self._n_staged:
            L.check(L.lib().mfvae_ring_add(self._h, C.c_void_p(self._stage.data_ptr()), self._n_staged, 0, self._stream()))
            torch.cuda.current_stream(self.device).synchronize()     # the staging block is reused immediately
            self._n_staged = 0

    def add_rows_device(self, rows: torch.Tensor):
        """Append already-packed rows that live on the device ([n, row] fp32)."""
        assert rows.dtype == torch.float32 and rows.shape[1] == self.row and rows.is_contiguous()
        L.check(L.lib().mfvae_ring_add(self._h, L.ptr(rows), rows.shape[0], 1, self._stream()))

    def sample_packed(self, batch: int, seed: int = 0, sample0: int = 0, batch_global: Optional[int] = None,
                      with_indices: bool = False, with_flags: bool = False):
        """Uniform-with-replacement batch gathered on the device into a ``PackedBatch``; ``with_flags`` also returns the
        [batch, 2 A + 1] matrix terminals | truncations | mask of the sampled transitions."""
        self.flush()
        dev = self.device
        obs = torch.empty(batch, self.S, device=dev); nxt = torch.empty(batch, self.S, device=dev)
        act = torch.empty(batch, self.W, device=dev); rew = torch.empty(batch, self.A, device=dev)
        idx = torch.empty(batch, dtype=torch.int32, device=dev) if with_indices else None
        flags = torch.empty(batch, 2 * self.A + 1, device=dev) if with_flags else None
        L.check(L.lib().mfvae_ring_sample(self._h, batch, seed, self._step, L.ptr(obs), L.ptr(act), L.ptr(nxt), L.ptr(rew),
                                          L.ptr(flags), L.ptr(idx), self._stream()))
        self._step += 1
        pb = PackedBatch(obs, act, nxt, rew, sample0=sample0, batch_global=batch_global)
        out = (pb,) + ((idx,) if with_indices else ()) + ((flags,) if with_flags else ())
        return out if len(out) > 1 else pb


class MultiAgentCPPRB:
    """Reference ``MultiAgentCPPRB(environment, max_size=10000, batch_size=32)`` over the device ring.
    ``environment`` needs ``.agents``, ``.observation_space(agent)`` and ``.action_space(agent)`` (gymnasium-like: Discrete
    -> 1 column, Box -> prod(shape) columns, replay_buffer.py:35-50); alternatively pass ``agents=`` / ``obs_dim=`` (and
    ``act_dim=`` for continuous actions) explicitly when no simulator object exists."""

    def __init__(self, environment=None, max_size=10000, batch_size=32, *, agents=None, obs_dim=None, act_dim=None, device="cuda:0",
                 seed=0):
        if environment is not None:
            agents = list(environment.agents)
            obs_dim = {a: _space_dim(environment.observation_space(a)) for a in agents}
            if hasattr(environment, "action_space"):
                act_dim = {a: _space_dim(environment.action_space(a)) for a in agents}
        if agents is None or obs_dim is None:
            raise NotImplementedError("need an environment or explicit agents / obs_dim")
        self._environment = environment
        self._max_size, self._batch_size, self._seed = max_size, batch_size, seed
        self.ring = DeviceRing(agents, obs_dim, max_size, device, act_dim=act_dim)

    def add(self, observations, next_observations, actions, rewards, terminals, trunctions):
        self.ring.add_row(observations, actions, next_observations, rewards, terminals, trunctions)

    def on_episode_end(self):
        self.ring.flush()

    def sample_packed(self, **kw) -> PackedBatch:
        return self.ring.sample_packed(self._batch_size, self._seed, **kw)

    def sample(self):
        pb, flags = self.ring.sample_packed(self._batch_size, self._seed, with_flags=True)
        out = {}
        obs, act, nxt, rew, fl = (t.cpu().numpy() for t in (pb.obs, pb.act, pb.next, pb.rew, flags))
        r = self.ring
        for i, a in enumerate(r.agents):
            o0, o1 = r._offsets[i], r._offsets[i + 1]
            out[f"{a}_observations"] = np.ascontiguousarray(obs[:, o0:o1])
            out[f"{a}_next_observations"] = np.ascontiguousarray(nxt[:, o0:o1])
            out[f"{a}_actions"] = np.ascontiguousarray(act[:, r._act_off[i]:r._act_off[i + 1]])
            out[f"{a}_rewards"] = np.ascontiguousarray(rew[:, i:i + 1])
            out[f"{a}_terminals"] = np.ascontiguousarray(fl[:, i:i + 1])
            out[f"{a}_truncations"] = np.ascontiguousarray(fl[:, r.A + i:r.A + i + 1])
        out["mask"] = np.ascontiguousarray(fl[:, 2 * r.A:2 * r.A + 1])
        return out

    def __iter__(self):
        return self

    def __next__(self):
        return self.sample()


class JaxFbxBuffer:
    """Reference ``JaxFbxBuffer(max_length, min_length, batch_size, add_batch)`` (jax_ver/jax_buffer.py:80-140) over the
    device ring.  ``sample(rng_key)`` takes any integer-like key (the Philox seed)."""

    def __init__(self, max_length: int = 50_000, min_length: int = 64, batch_size: int = 64, add_batch: bool = False,
                 device="cuda:0"):
        self.max_length, self.min_length, self.batch_size, self.add_batch = max_length, min_length, batch_size, add_batch
        self.device = device
        self.ring: Optional[DeviceRing] = None
        self.buffer_state = None

    def init_buffer(self, obs, reward, actions, next_obs, done):
        agents = list(obs.keys())
        self.ring = DeviceRing(agents, {a: int(np.asarray(obs[a]).size) for a in agents}, self.max_length, self.device)
        self.buffer_state = self.ring

    def add_trans(self, obs, reward, actions, next_obs, done):
        if self.buffer_state is None:
            print("buffer not init; please call init_buffer() first")
            return
        for a in obs.keys():
            if a not in reward or a not in actions or a not in next_obs or a not in done:
                print(f"agent id {a} not exist in action/reward/next_obs/done dict")
                return
        self.ring.add_row(obs, actions, next_obs, reward, done, None, mask=float(any(bool(v) for v in done.values())))

    def can_sample(self):
        if self.buffer_state is None:
            print("buffer not init; please call init_buffer() first")
            return
        return len(self.ring) >= self.min_length

    def sample_packed(self, rng_key=0, **kw) -> PackedBatch:
        return self.ring.sample_packed(self.batch_size, int(rng_key), **kw)

    def sample(self, rng_key):
        if self.buffer_state is None:
            print("buffer not init; please call init_buffer() first")
            return
        if not self.can_sample():
            print("can not sample now")
            return
        pb, flags = self.ring.sample_packed(self.batch_size, int(rng_key), with_flags=True)
        r = self.ring
        exp = {}
        obs, act, nxt, rew = (t.cpu().numpy() for t in (pb.obs, pb.act, pb.next, pb.rew))
        for i, a in enumerate(r.agents):
            o0, o1 = r._offsets[i], r._offsets[i + 1]
            exp[f"{a}_obs"] = obs[:, o0:o1, None].copy()
            exp[f"{a}_next_obs"] = nxt[:, o0:o1, None].copy()
            exp[f"{a}_act"] = act[:, i:i + 1, None].copy()
            exp[f"{a}_rew"] = rew[:, i:i + 1, None].copy()
        exp["done"] = flags[:, 2 * r.A].cpu().numpy().reshape(-1, 1, 1)       # ma_done of jax_buffer.py:37,49-52
        return SimpleNamespace(experience=exp)
