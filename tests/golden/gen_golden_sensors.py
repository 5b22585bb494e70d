"""Golden vectors for the later fork's sensor classes, taken from the UNMODIFIED reference class

    CS = MADDPG_ownENV_randomOD_radar_N_model_use_tdCPA_forV2_changeskin/env_simulator_randomOD_radar_sur_drones_N_Model_use_tdCPA_forV2_changeskin.py

run through tests/golden/ref_harness.py (build container only; geometry backend recorded in the fixture):
  * `cloud_traj`: the clouds' positions over 150 real `env.step` calls (CS:4667-4681, calculate_next_position);
  * probes: the fork's `cur_state_norm_state_v3(..., include_other_AC=True, use_nearestN_neigh_wRadar=True, N_neigh=2, ...)`
    (CS:1327-1900) evaluated at teleported drone states next to clouds, boundaries and each other - its radar
    (boundary segments, cloud outlines, other aircraft's outlines, true minimum: CS:1379-1506) and its nearest-N
    neighbour block (CS:1799-1802).

    python tests/golden/gen_golden_sensors.py        # writes tests/golden/cs_sensors.npz
"""
import os
import random
import sys
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
for name in ("matplotlib.path", "cairosvg", "PIL", "PIL.Image"):   # plotting-only imports of the fork's utilities
    m = mock.MagicMock(name=name)
    m.__path__, m.__spec__ = [], None
    sys.modules.setdefault(name, m)
import ref_harness as H  # noqa: E402
from multi_agent_aac_b200.maps import synthetic_map  # noqa: E402

H.VARIANTS["cs"] = ("MADDPG_ownENV_randomOD_radar_N_model_use_tdCPA_forV2_changeskin",
                    "env_simulator_randomOD_radar_sur_drones_N_Model_use_tdCPA_forV2_changeskin")
N, R, N_NEIGH, BOUND = 4, 18, 2, [0, 200, 0, 200]


def main():
    gmap = synthetic_map(bound=BOUND, seed=0)
    sys.modules.pop("cloud", None)
    mod = H.load_reference_module("cs")
    env = mod.env_simulator(gmap.occ.astype(float), [], gmap.grid_length, list(gmap.bound), H.grid_polys(gmap), None)
    env.create_world(N, 2, 0.95, 0.01, 1, 0.15, 0.05, 0.15, (1800, 1300), 5, [-8, 8])
    args = types.SimpleNamespace(mode="train")
    random.seed(5)
    np.random.seed(5)
    devnull, old = open(os.devnull, "w"), sys.stdout
    sys.stdout = devnull
    try:
        env.reset_world_change_skin(N, False, False, True, True, N_NEIGH, args, 0)
        clouds = env.cloud_config
        cloud_cfg = np.array([[c.ini_pos.x, c.ini_pos.y, c.goal.x, c.goal.y, c.radius, c.vel] for c in clouds])
        rng = np.random.default_rng(0)
        traj = [[(c.pos.x, c.pos.y) for c in clouds]]
        for t in range(1, 151):     # real steps: only the clouds' motion is recorded (drone actions are irrelevant to it)
            env.step(rng.uniform(-0.2, 0.2, (N, 2)), t, 8, args, True, False, False, True, True, N_NEIGH)
            traj.append([(c.pos.x, c.pos.y) for c in clouds])
        traj = np.array(traj)
        Point = type(clouds[0].pos)
        K = 320
        rec = {k: [] for k in ("pos", "vel", "heading", "cloud_k", "radar", "radar_noac", "raw_nbr", "norm_nbr", "norm_own")}
        for n in range(K):
            k = int(rng.integers(0, 151))
            for ci, c in enumerate(clouds):
                c.pos = Point(float(traj[k, ci, 0]), float(traj[k, ci, 1]))
                c.cloud_actual_cur_shape = c.pos.buffer(c.radius)
            kind = n % 4
            anchor = traj[k, n % len(clouds)] if kind in (0, 1) else np.array([rng.choice([2.0, 100.0, 198.0]), rng.choice([2.0, 100.0, 198.0])])
            spread = [26.0, 11.0, 14.0, 6.0][kind]      # near a cloud's outline, inside it, near a boundary / corner, tightly clustered
            pos = anchor + rng.uniform(-spread, spread, (N, 2))
            pos = np.clip(pos, 0.5, 199.5)
            vel = rng.uniform(-3.5, 3.5, (N, 2))
            head = rng.uniform(-np.pi, np.pi, N)
            for i in range(N):
                ag = env.all_agents[i]
                ag.pos, ag.vel, ag.heading = pos[i].copy(), vel[i].copy(), float(head[i])
                ag.pre_pos, ag.pre_vel = pos[i] - 0.5 * vel[i], vel[i].copy()
                ag.surroundingNeighbor, ag.pre_surroundingNeighbor = {}, {}
            out = env.cur_state_norm_state_v3({}, False, True, True, N_NEIGH, args, False)
            st, nst = out[0], out[1]
            radar = np.stack([np.asarray(env.all_agents[i].observableSpace, dtype=np.float64) for i in range(N)])
            out2 = env.cur_state_norm_state_v3({}, False, False, True, N_NEIGH, args, False)    # radar without the other aircraft
            radar_noac = np.stack([np.asarray(env.all_agents[i].observableSpace, dtype=np.float64) for i in range(N)])
            rec["pos"].append(pos); rec["vel"].append(vel); rec["heading"].append(head); rec["cloud_k"].append(k)
            rec["radar"].append(radar); rec["radar_noac"].append(radar_noac)
            rec["raw_nbr"].append(np.stack([np.asarray(st[1][i], dtype=np.float64) for i in range(N)]))
            rec["norm_nbr"].append(np.stack([np.asarray(nst[1][i], dtype=np.float64) for i in range(N)]))
            rec["norm_own"].append(np.stack([np.asarray(nst[0][i], dtype=np.float64) for i in range(N)]))
    finally:
        sys.stdout = old
    outp = os.path.join(HERE, "cs_sensors.npz")
    np.savez_compressed(outp, meta=np.array([N, R, N_NEIGH, int(env.all_agents[0].protectiveBound)]), bound=np.array(BOUND, dtype=np.float64), occ=gmap.occ, cloud_cfg=cloud_cfg,
                        cloud_traj=traj, geometry=np.array(H.GEOMETRY), **{k: np.array(v) for k, v in rec.items()})
    r = np.array(rec["radar"])
    print("wrote", outp, "probes", K, "rays hitting something: %.1f%%" % (100 * (r < 15 - 1e-9).mean()),
          "hit only with the other aircraft: %.1f%%" % (100 * (r < np.array(rec["radar_noac"]) - 1e-9).mean()), "geometry:", H.GEOMETRY)


if __name__ == "__main__":
    main()
