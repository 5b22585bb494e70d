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

from multi_agent_aac_b200.actor_params import ATT_KEYS, KEYS, reference_like_params, reference_like_params_att  # noqa: F401  (re-exported)


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


# ---------------------------------------------------------------------------------------------------------------------
# ActorNetwork_ATT_TwoPortion (ATT/Nnetworks:177-213): the attention actor of the one_model_att variant
#   own_fc / own_grid / neigh_fc = Linear(d, 64) + ReLU                         ATT/Nnetworks:181-183
#   q, k, v = Linear(64, 64, bias=False)                                         ATT/Nnetworks:188-190
#   score = k(x_e) . q(own) / sqrt(64), masked softmax over the neighbours, v_att = sum alpha v   ATT/Nnetworks:197-208
#   merge_feature = Linear(192, 256) + ReLU; act_out = Linear(256, 2) + Tanh     ATT/Nnetworks:184-185, :210-212
# Parity pinned: tests/golden/actor_att*.npz (outputs of the unmodified reference class in float64).





def forward_att(sd, own, grid, nei):
    """float64 forward; own [B, d_own], grid [B, R], nei [B, M, 6] -> actions [B, 2]."""
    p = {k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}
    relu = lambda x: np.maximum(x, 0.0)
    own, grid, nei = (np.asarray(a, np.float64) for a in (own, grid, nei))
    own_obs = relu(own @ p["own_fc.0.weight"].T + p["own_fc.0.bias"])
    own_grid = relu(grid @ p["own_grid.0.weight"].T + p["own_grid.0.bias"])
    x_e = relu(nei @ p["neigh_fc.0.weight"].T + p["neigh_fc.0.bias"])              # [B, M, 64]
    q = own_obs @ p["q.weight"].T
    k = x_e @ p["k.weight"].T
    v = x_e @ p["v.weight"].T
    mask = nei.mean(axis=2) != 0.0                                                   # [B, M]
    score = np.einsum("bmc,bc->bm", k, q) / np.sqrt(64.0)
    score = np.where(mask, score, -np.inf)
    with np.errstate(invalid="ignore"):
        e = np.exp(score - score.max(axis=1, keepdims=True))
        alpha = e / e.sum(axis=1, keepdims=True)                                     # all-masked rows: nan, zeroed below
    alpha = np.where(mask, alpha, 0.0)
    v_att = np.einsum("bm,bmc->bc", alpha, v)
    h = relu(np.concatenate([own_obs, own_grid, v_att], axis=1) @ p["merge_feature.0.weight"].T + p["merge_feature.0.bias"])
    return np.tanh(h @ p["act_out.0.weight"].T + p["act_out.0.bias"])
