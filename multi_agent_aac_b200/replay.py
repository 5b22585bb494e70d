"""Device-resident replay ring for batched rollouts (second half).

The reference keeps a Python list of namedtuples, one joint transition of ONE env per push,
    Experience(states_obs, states_nei, states_grid, actions, next_states_*, rewards, dones, ...)   V2/memory:3-4
    ReplayMemory.push / sample (random.sample)                                                      V2/memory:11-21
filled from the training loop after every env step (V2/ma_main:526).  At 10^9 agent-steps/s that list is the
wall, and so would be a copy of 240 MB of observations per step.  Here a slot of the ring IS the env's output buffer
for that step (`BatchedDroneEnv.bind_outputs`): the step kernel writes observations, rewards and done flags of all E
envs directly into the ring, the actor writes its actions into the ring, and `sample` gathers joint transitions of
random (step, env) pairs with the next observation taken from the following slot.  Nothing leaves HBM and nothing is
copied on the way in.

Slot s holds obs_s (what the actor saw), act_s, and the outcome of act_s: reward_s, done_s, terminated_s.  The
observation produced by act_s lands in slot s + 1; for an env whose episode ended at s (terminated_s != 0) that row is
already the NEXT episode's first observation (auto-reset), so it must never be bootstrapped from.  `done` alone does
not say that: the reference ends an episode with done = False at the step cap and when every drone has arrived, and
pushes the true next state in those cases (ATT/ma_main:448-462, V2/ma_main:578-589).  `sample` therefore returns
`bootstrap` = (terminated == 0): multiply the target critic's value by it instead of by (1 - done).  Crash endings are
treated exactly as the reference treats them; the two done = False endings are treated as terminal (time-limit
truncation handled as termination) rather than bootstrapped from an unrelated episode.
"""
from __future__ import annotations

import torch

OBS_KEYS = ("norm_own", "norm_nbr", "radar")


class DeviceReplay:
    def __init__(self, env, capacity_steps):
        if capacity_steps < 2:
            raise ValueError("capacity_steps must be at least 2")
        if any(k not in env.out for k in OBS_KEYS):
            raise ValueError("DeviceReplay needs the norm_own / norm_nbr / radar outputs (tdCPA_forV2 preset)")
        self.env, self.T = env, int(capacity_steps)
        dev = env.device
        self.obs = {k: torch.zeros((self.T,) + tuple(env.out[k].shape), dtype=env.out[k].dtype, device=dev) for k in OBS_KEYS}
        self.act = torch.zeros((self.T, env.E, env.N, 2), dtype=torch.float32, device=dev)
        self.res = {k: torch.zeros((self.T,) + tuple(env.out[k].shape), dtype=env.out[k].dtype, device=dev) for k in ("reward", "done", "terminated")}
        self.head = 0        # slot whose observation is current and whose action comes next
        self.filled = 0      # completed transitions in the ring

    def bytes(self):
        return sum(t.numel() * t.element_size() for t in list(self.obs.values()) + list(self.res.values()) + [self.act])

    def begin(self):
        """Binds slot 0 as the env's observation buffer and observes the current state into it."""
        self.head, self.filled = 0, 0
        self.env.bind_outputs({k: self.obs[k][0] for k in OBS_KEYS})
        return self.env.observe()

    def current_obs(self):
        return {k: self.obs[k][self.head] for k in OBS_KEYS}

    def action_slot(self):
        """The [E, N, 2] tensor the policy writes its actions for the current observation into."""
        return self.act[self.head]

    def step(self):
        """Applies `action_slot()`: outcome into the current slot, next observation into the next one."""
        s, nxt = self.head, (self.head + 1) % self.T
        bind = {k: self.obs[k][nxt] for k in OBS_KEYS}
        bind.update({k: self.res[k][s] for k in self.res})
        self.env.bind_outputs(bind)
        out = self.env.step(self.act[s], autoreset=True)
        self.head = nxt
        self.filled = min(self.filled + 1, self.T - 1)
        return out

    def sample(self, batch, generator=None):
        """`batch` joint transitions (all N drones of one env at one step), uniformly over the stored ones:
        dict of obs / act / reward / done / terminated / bootstrap / next_obs tensors with leading dimension `batch`.
        `bootstrap` [batch] is 1.0 where next_obs continues the same episode and 0.0 where the episode ended at this step
        (next_obs is then the first observation of the following episode: mask the bootstrap term with it)."""
        if self.filled == 0:
            raise ValueError("the ring holds no completed transition yet")
        dev = self.env.device
        age = torch.randint(1, self.filled + 1, (batch,), device=dev, generator=generator)   # 1 = newest completed transition
        s = (self.head - age) % self.T
        e = torch.randint(0, self.env.E, (batch,), device=dev, generator=generator)
        out = {"step_slot": s, "env": e, "act": self.act[s, e]}
        for k in OBS_KEYS:
            out[k] = self.obs[k][s, e]
            out["next_" + k] = self.obs[k][(s + 1) % self.T, e]
        for k, t in self.res.items():
            out[k] = t[s, e]
        out["bootstrap"] = (out["terminated"] == 0).to(torch.float32)
        return out
