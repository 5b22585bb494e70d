"""Generates tests/golden/actor_v2.npz from the UNMODIFIED reference actor class (build container only).

    python tests/golden/gen_golden_actor.py

The reference's Nnetworks module is imported from /root/reference with the same inert stubs ref_harness.py
uses for the env (it star-imports Utilities_own, which imports matplotlib / openpyxl at module top).  The
parameters come from oracle.actor_oracle.reference_like_params (numpy Generator, so they can be rebuilt
anywhere without torch's RNG); inputs have the value ranges of the env's observations.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import actor_oracle  # noqa: E402
from tests.golden import ref_harness  # noqa: E402

CASES = {"actor_v2": dict(d_own=7, d_nbr=45, d_grid=36, rows=96, seed=0),
         "actor_v2_r18_n4": dict(d_own=7, d_nbr=15, d_grid=18, rows=40, seed=1)}


def reference_actor(d_own, d_nbr, d_grid):
    ref_harness._install_stubs()
    path = os.path.join(ref_harness.REFERENCE, ref_harness.VARIANTS["v2"][0])
    sys.path.insert(0, path)
    try:
        mod = importlib.import_module("Nnetworks_randomOD_radar_sur_drones_N_Model_use_tdCPA_forV2")
    finally:
        sys.path.remove(path)
    return mod.ActorNetwork_allnei_wRadar([d_own, d_nbr, d_grid, 6], 2).double()


ATT_CASES = {"actor_att": dict(d_own=14, d_grid=36, n_nei=2, rows=96, seed=2),
             "actor_att_n5_r18": dict(d_own=22, d_grid=18, n_nei=4, rows=48, seed=3)}


def reference_actor_att(d_own, d_grid):
    ref_harness._install_stubs()
    path = os.path.join(ref_harness.REFERENCE, ref_harness.VARIANTS["att"][0])
    for k in [k for k in sys.modules if k.startswith(("Utilities_own", "Nnetworks"))]:
        del sys.modules[k]
    sys.path.insert(0, path)
    try:
        mod = importlib.import_module("Nnetworks_randomOD_radar_sur_drones_oneModel_att")
    finally:
        sys.path.remove(path)
    return mod.ActorNetwork_ATT_TwoPortion([d_own, d_grid, 6], 2).double()


def main_att():
    for name, c in ATT_CASES.items():
        sd = actor_oracle.reference_like_params_att(c["d_own"], c["d_grid"], c["seed"])
        net = reference_actor_att(c["d_own"], c["d_grid"])
        net.load_state_dict({k: torch.from_numpy(v).double() for k, v in sd.items()})
        rng = np.random.default_rng(200 + c["seed"])
        own = rng.uniform(-1, 1, (c["rows"], c["d_own"]))
        grid = np.where(rng.uniform(size=(c["rows"], c["d_grid"])) < 0.6, 15.0, rng.uniform(0, 15, (c["rows"], c["d_grid"])))
        nei = rng.uniform(-1, 1, (c["rows"], c["n_nei"], 6))
        nei[rng.uniform(size=nei.shape[:2]) < 0.25] = 0.0          # absent neighbours are all-zero rows: masked (ATT/Nnetworks:200)
        nei[:4] = 0.0                                               # rows with every neighbour masked
        with torch.no_grad():
            act = net([torch.from_numpy(own), torch.from_numpy(grid), torch.from_numpy(nei)]).numpy()
        mine = actor_oracle.forward_att(sd, own, grid, nei)
        assert np.abs(mine - act).max() < 1e-12, np.abs(mine - act).max()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), own=own, grid=grid, nei=nei, act=act,
                            dims=np.array([c["d_own"], c["d_grid"], c["n_nei"], c["seed"]]))
        print(name, act.shape, "oracle-vs-reference max abs diff", np.abs(mine - act).max(), "nan", int(np.isnan(act).sum()))


def main():
    for name, c in CASES.items():
        sd = actor_oracle.reference_like_params(c["d_own"], c["d_nbr"], c["d_grid"], c["seed"])
        net = reference_actor(c["d_own"], c["d_nbr"], c["d_grid"])
        net.load_state_dict({k: torch.from_numpy(v).double() for k, v in sd.items()})
        rng = np.random.default_rng(100 + c["seed"])
        own = rng.uniform(-1, 1, (c["rows"], c["d_own"]))
        own[:, -1] = rng.uniform(-np.pi, np.pi, c["rows"])          # heading, unnormalised
        nbr = rng.uniform(-1, 1, (c["rows"], c["d_nbr"]))
        grid = rng.uniform(0, 15, (c["rows"], c["d_grid"]))          # radar ranges in metres
        grid[rng.uniform(size=grid.shape) < 0.5] = 15.0
        with torch.no_grad():
            act = net([torch.from_numpy(own), torch.from_numpy(nbr), torch.from_numpy(grid)]).numpy()
        mine = actor_oracle.forward(sd, own, nbr, grid)
        assert np.abs(mine - act).max() < 1e-12, np.abs(mine - act).max()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), own=own, nbr=nbr, grid=grid, act=act,
                            dims=np.array([c["d_own"], c["d_nbr"], c["d_grid"], c["seed"]]))
        print(name, act.shape, "oracle-vs-reference max abs diff", np.abs(mine - act).max())


if __name__ == "__main__":
    main()
    main_att()
