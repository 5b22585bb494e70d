"""Random parameters with the reference actors' shapes and torch.nn.Linear's default scale (uniform +-1/sqrt(fan_in)), drawn from
a numpy Generator so that fixtures, tests and the bench rebuild the same float32 values on any machine (there are no
checkpoints to load here).  Keys follow the reference modules' state_dict names (V2/Nnetworks:292-298, ATT/Nnetworks:181-190)."""
import numpy as np

KEYS = ["own_fc.0", "own_full_nei.0", "own_grid.0", "merge_feature.0", "act_out.0", "act_out.2"]
ATT_KEYS = ["own_fc.0", "own_grid.0", "neigh_fc.0", "merge_feature.0", "act_out.0"]


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


def reference_like_params_att(d_own, d_grid, seed=0, d_nei=6):
    rng = np.random.default_rng(seed)
    shapes = {"own_fc.0": (64, d_own), "own_grid.0": (64, d_grid), "neigh_fc.0": (64, d_nei), "merge_feature.0": (256, 192),
              "act_out.0": (2, 256)}
    sd = {}
    for k in ATT_KEYS:
        out_f, in_f = shapes[k]
        lim = 1.0 / np.sqrt(in_f)
        sd[k + ".weight"] = rng.uniform(-lim, lim, (out_f, in_f)).astype(np.float32)
        sd[k + ".bias"] = rng.uniform(-lim, lim, (out_f,)).astype(np.float32)
    for k in ("k", "q", "v"):
        sd[k + ".weight"] = rng.uniform(-0.125, 0.125, (64, 64)).astype(np.float32)
    return sd
