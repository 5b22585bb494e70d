"""CPU restatement (numpy float64) of the reference actor forward pass.  TEST INFRASTRUCTURE ONLY: imported
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the product never imports it.

Follows ActorNetwork_allnei_wRadar (V2/Nnetworks:273-340):
    own_fc / own_full_nei / own_grid = Linear(d, 128) + LeakyReLU(0.01)      V2/Nnetworks:292-294
    merge_feature = Linear(384, 512) + LeakyReLU(0.01)                        V2/Nnetworks:295
    act_out = Linear(512, 256) + LeakyReLU(0.01) + Linear(256, 2) + Tanh      V2/Nnetworks:297-298
    forward: cat((own, nei, radar), dim=1) -> merge_feature -> act_out        V2/Nnetworks:331-337
and the exploration step of choose_action (V2/maddpg_agent:1290-1294): act + noise, clamp to [-1, 1].

Parity pinned: tests/golden/actor_v2.npz holds inputs and the outputs of the UNMODIFIED reference class
(float64) for the parameters `reference_like_params(seed)` generates (tests/golden/gen_golden_actor.py).
"""
import numpy as np

KEYS = ["own_fc.0", "own_full_nei.0", "own_grid.0", "merge_feature.0", "act_out.0", "act_out.2"]


def reference_like_params(d_own, d_nbr, d_grid, seed=0):
    """Parameters with torch.nn.Linear's default scale (uniform +-1/sqrt(fan_in)), from a numpy Generator so that
    the fixture generator, the tests and the bench all rebuild the same float32 values on any machine."""
    rng = np.random.default_rng(seed)
    shapes = {"own_fc.0": (128, d_own), "own_full_nei.0": (128, d_nbr), "own_grid.0": (128, d_grid),
              "merge_feature.0": (512, 384), "act_out.0": (256, 512), "act_out.2": (2, 256)}
    sd = {}
    for k in KEYS:
        out_f, in_f = shapes[k]
        lim = 1.0 / np.sqrt(in_f)
        sd[k + ".weight"] = rng.uniform(-lim, lim, (out_f, in_f)).astype(np.float32)
        sd[k + ".bias"] = rng.uniform(-lim, lim, (out_f,)).astype(np.float32)
    return sd


def _leaky(x):
    return np.where(x > 0, x, 0.01 * x)


def forward(sd, own, nbr, grid, hidden=False):
    """float64 forward; `sd` maps '<module>.weight' / '<module>.bias' to arrays ([out, in] / [out])."""
    p = {k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}
    lin = lambda name, x: x @ p[name + ".weight"].T + p[name + ".bias"]
    h1 = np.concatenate([_leaky(lin("own_fc.0", np.asarray(own, np.float64))),
                         _leaky(lin("own_full_nei.0", np.asarray(nbr, np.float64))),
                         _leaky(lin("own_grid.0", np.asarray(grid, np.float64)))], axis=1)
    h2 = _leaky(lin("merge_feature.0", h1))
    h3 = _leaky(lin("act_out.0", h2))
    act = np.tanh(lin("act_out.2", h3))
    return (act, h1, h2, h3) if hidden else act


def explore(act, noise, scale):
    """choose_action's exploration (V2/maddpg_agent:1290-1294)."""
    return np.clip(act + scale * noise, -1.0, 1.0)
