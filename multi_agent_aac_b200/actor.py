"""Batched actor: the reference's `choose_action` for every drone of every env in one launch.

Mirrors, for the network the tdCPA_forV2 configuration trains (`ActorNetwork_allnei_wRadar`,
V2/Nnetworks:273-340; selected by use_allNeigh_wRadar=True, one shared model, V2/ma_main:81,95,100),

    maddpg.choose_action(state, ..., noisy=...)        V2/maddpg_agent:1241-1310

which runs N sequential batch-1 forwards per env step on `[obs, obs_full_nei, obs_grid]`.  Here the three
inputs are the env step's `norm_own`, `norm_nbr`, `radar` tensors as they lie in device memory, and the
actions come back in the `[E, N, 2]` layout `BatchedDroneEnv.step` takes, so a rollout never leaves HBM:

    actor = BatchedActor.for_env(env); actor.load_state_dict(reference_actor.state_dict())
    obs = env.reset()
    for _ in range(T):
        obs, reward, done, info = env.step(actor(obs, noise_scale=var), autoreset=True)   # fresh noise at every call

Exploration noise is `act + var * randn(2)` drawn anew for every drone at every call (V2/maddpg_agent:1290-1294).  The
draws are counter based, keyed by (noise_seed, row); with no `noise_seed` the actor advances its own call counter, so
successive calls never repeat a draw; pass `noise_seed` to make a call reproducible.

Compute: bf16 tensor-core GEMM chain with fp32 accumulation inside one fused sm_100a kernel
(csrc/aac_actor.cu); there is no fallback path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _actor_capi as K

# reference module names (V2/Nnetworks:292-298) -> fields of AacActorParams
_MODULES = [("own_fc.0", "own"), ("own_full_nei.0", "nbr"), ("own_grid.0", "grid"), ("merge_feature.0", "merge"), ("act_out.0", "hid"),
            ("act_out.2", "out")]


class BatchedActor:
    def __init__(self, d_own, d_nbr, d_grid, max_rows, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise K.AacActorError("BatchedActor runs on a CUDA device only (no CPU path)")
        self.d_own, self.d_nbr, self.d_grid, self.max_rows = int(d_own), int(d_nbr), int(d_grid), int(max_rows)
        self._h = C.c_void_p()
        cfg = K.AacActorConfig(K.ABI_VERSION, self.d_own, self.d_nbr, self.d_grid, self.max_rows)
        with torch.cuda.device(self.device):
            K.check(K.lib().aac_actor_create(C.byref(cfg), C.byref(self._h)), "aac_actor_create")

    @classmethod
    def for_env(cls, env):
        """Actor sized for a `BatchedDroneEnv` of the tdCPA_forV2 preset (own 7, nbr 5 (N - 1), grid R)."""
        return cls(env.D, 5 * (env.N - 1), env.R, env.E * env.N, device=env.device)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and K is not None and getattr(K, "_lib", None) is not None:   # module globals may be gone at interpreter exit
            K._lib.aac_actor_destroy(h)

    def load_state_dict(self, sd):
        """`sd`: the reference module's state_dict (tensors or arrays keyed 'own_fc.0.weight', ...)."""
        keep, params = [], K.AacActorParams()
        shapes = {"own": (K.H1, self.d_own), "nbr": (K.H1, self.d_nbr), "grid": (K.H1, self.d_grid), "merge": (K.H2, 3 * K.H1),
                  "hid": (K.H3, K.H2), "out": (K.NACT, K.H3)}
        for mod, field in _MODULES:
            for kind, pre in (("weight", "w_"), ("bias", "b_")):
                v = sd[mod + "." + kind]
                a = np.ascontiguousarray(v.detach().cpu().numpy() if torch.is_tensor(v) else v, dtype=np.float32)
                want = shapes[field] if kind == "weight" else (shapes[field][0],)
                if a.shape != want:
                    raise ValueError("%s.%s has shape %s, expected %s" % (mod, kind, a.shape, want))
                keep.append(a)
                setattr(params, pre + field, a.ctypes.data)
        with torch.cuda.device(self.device):
            K.check(K.lib().aac_actor_load(self._h, C.byref(params)), "aac_actor_load")

    def _rows(self, own, nbr, grid):
        for t, d, name in ((own, self.d_own, "own"), (nbr, self.d_nbr, "nbr"), (grid, self.d_grid, "grid")):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device or t.shape[-1] != d:
                raise ValueError("%s must be a contiguous float32 tensor [..., %d] on %s" % (name, d, self.device))
        n = own.numel() // self.d_own
        if nbr.numel() != n * self.d_nbr or grid.numel() != n * self.d_grid:
            raise ValueError("own / nbr / grid disagree on the number of rows")
        return n

    def _seed(self, noise_seed):
        """None -> this actor's own call counter (a new draw at every call), else the caller's seed."""
        if noise_seed is not None:
            return int(noise_seed)
        self._calls = getattr(self, "_calls", 0) + 1
        return 0x5EED0000 + self._calls

    def forward(self, own, nbr, grid, noise_scale=0.0, noise_seed=None, out=None):
        """actions [..., 2] = clamp(actor(own, nbr, grid) + noise_scale * N(0, 1), -1, 1); leading dims follow `own`.
        noise_seed=None: fresh noise at every call; an int makes the call reproducible."""
        noise_seed = self._seed(noise_seed)
        n = self._rows(own, nbr, grid)
        if out is None:
            out = torch.empty(tuple(own.shape[:-1]) + (K.NACT,), dtype=torch.float32, device=self.device)
        elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != n * K.NACT or out.device != self.device:
            raise ValueError("out must be a contiguous float32 tensor with %d elements" % (n * K.NACT))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        K.check(K.lib().aac_actor_forward(self._h, own.data_ptr(), nbr.data_ptr(), grid.data_ptr(), n, float(noise_scale), int(noise_seed),
                                          out.data_ptr(), stream), "aac_actor_forward")
        return out

    def __call__(self, obs, noise_scale=0.0, noise_seed=None, out=None):
        """`obs`: the dict `BatchedDroneEnv.reset/step` returns ([obs, obs_full_nei, obs_grid] of V2/maddpg_agent:1243-1245)."""
        return self.forward(obs["norm_own"], obs["norm_nbr"], obs["radar"], noise_scale, noise_seed, out)

    def hidden(self, layer, own, nbr, grid):
        """Post-activation output of hidden layer 1 / 2 / 3 (parity aid)."""
        n = self._rows(own, nbr, grid)
        width = {1: 3 * K.H1, 2: K.H2, 3: K.H3}[layer]
        out = torch.empty((n, width), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        K.check(K.lib().aac_actor_hidden(self._h, own.data_ptr(), nbr.data_ptr(), grid.data_ptr(), n, layer, out.data_ptr(), stream),
                "aac_actor_hidden")
        return out

    @property
    def launch_count(self):
        return int(K.lib().aac_actor_launch_count(self._h))


class BatchedAttActor:
    """The attention actor of the one_model_att variant (`ActorNetwork_ATT_TwoPortion`, ATT/Nnetworks:177-213) for every
    drone of every env in one launch: `choose_action` (ATT/maddpg_agent:455-503) on `[obs, obs_grid, obs_nei]` = the env's
    `norm_own`, `radar`, `norm_nbr6` tensors.  fp32 on the CUDA cores (csrc/aac_actor_att.cu); no fallback path."""

    _MODULES = [("own_fc.0", "own", True), ("own_grid.0", "grid", True), ("neigh_fc.0", "nei", True), ("q", "q", False), ("k", "k", False),
                ("v", "v", False), ("merge_feature.0", "merge", True), ("act_out.0", "out", True)]

    def __init__(self, d_own, d_grid, n_nei, d_nei=6, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise K.AacActorError("BatchedAttActor runs on a CUDA device only (no CPU path)")
        self.d_own, self.d_grid, self.n_nei, self.d_nei = int(d_own), int(d_grid), int(n_nei), int(d_nei)
        self._h = C.c_void_p()
        cfg = K.AacActorAttConfig(K.ABI_VERSION, self.d_own, self.d_grid, self.d_nei, self.n_nei)
        with torch.cuda.device(self.device):
            rc = K.lib().aac_actor_att_create(C.byref(cfg), C.byref(self._h))
        self._check(rc, "aac_actor_att_create")

    @staticmethod
    def _check(rc, what):
        if rc != 0:
            raise K.AacActorError("%s failed (%d): %s" % (what, rc, K.lib().aac_actor_att_last_error().decode()))

    @classmethod
    def for_env(cls, env):
        """Actor sized for a `BatchedDroneEnv` of the att preset (own 6 + 4 (N - 1), grid R, N - 1 neighbour rows of 6)."""
        return cls(env.D, env.R, env.N - 1, device=env.device)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and K is not None and getattr(K, "_lib", None) is not None:
            K._lib.aac_actor_att_destroy(h)

    def load_state_dict(self, sd):
        """`sd`: the reference module's state_dict ('own_fc.0.weight', ..., 'q.weight', 'k.weight', 'v.weight')."""
        keep, params = [], K.AacActorAttParams()
        shapes = {"own": (64, self.d_own), "grid": (64, self.d_grid), "nei": (64, self.d_nei), "q": (64, 64), "k": (64, 64), "v": (64, 64),
                  "merge": (256, 192), "out": (2, 256)}
        for mod, field, has_bias in self._MODULES:
            for kind, pre in (("weight", "w_"), ("bias", "b_")) if has_bias else (("weight", "w_"),):
                v = sd[mod + "." + kind]
                a = np.ascontiguousarray(v.detach().cpu().numpy() if torch.is_tensor(v) else v, dtype=np.float32)
                want = shapes[field] if kind == "weight" else (shapes[field][0],)
                if a.shape != want:
                    raise ValueError("%s.%s has shape %s, expected %s" % (mod, kind, a.shape, want))
                keep.append(a)
                setattr(params, pre + field, a.ctypes.data)
        with torch.cuda.device(self.device):
            self._check(K.lib().aac_actor_att_load(self._h, C.byref(params)), "aac_actor_att_load")

    def forward(self, own, grid, nei, noise_scale=0.0, noise_seed=None, out=None):
        """actions [..., 2]; own [..., d_own], grid [..., d_grid], nei [..., n_nei, d_nei] (leading dims follow `own`).
        noise_seed=None: fresh exploration noise at every call (the actor's own call counter keys the draws)."""
        if noise_seed is None:
            self._calls = getattr(self, "_calls", 0) + 1
            noise_seed = 0x5EED0000 + self._calls
        for t, tail, name in ((own, (self.d_own,), "own"), (grid, (self.d_grid,), "grid"), (nei, (self.n_nei, self.d_nei), "nei")):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device or tuple(t.shape[-len(tail):]) != tail:
                raise ValueError("%s must be a contiguous float32 tensor [..., %s] on %s" % (name, ", ".join(map(str, tail)), self.device))
        n = own.numel() // self.d_own
        if grid.numel() != n * self.d_grid or nei.numel() != n * self.n_nei * self.d_nei:
            raise ValueError("own / grid / nei disagree on the number of rows")
        if out is None:
            out = torch.empty(tuple(own.shape[:-1]) + (K.NACT,), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(K.lib().aac_actor_att_forward(self._h, own.data_ptr(), grid.data_ptr(), nei.data_ptr(), n, float(noise_scale),
                                                  int(noise_seed), out.data_ptr(), stream), "aac_actor_att_forward")
        return out

    def __call__(self, obs, noise_scale=0.0, noise_seed=None, out=None):
        """`obs`: the dict an att-preset `BatchedDroneEnv` returns ([obs, obs_grid, obs_nei] of ATT/maddpg_agent:457-459)."""
        return self.forward(obs["norm_own"], obs["radar"], obs["norm_nbr6"], noise_scale, noise_seed, out)

    @property
    def launch_count(self):
        return int(K.lib().aac_actor_att_launch_count(self._h))
