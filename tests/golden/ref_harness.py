"""Run the UNMODIFIED reference env_simulator classes from /root/reference (build container only).

Test infrastructure: used by gen_golden.py to produce the committed fixtures under tests/golden/.
Nothing here runs on the GPU box (no /root/reference there).  Geometry: the REAL shapely when it can be
imported (SURVEY.md section 8c "upgrade path": the rollouts are then pinned to GEOS itself), else
oracle/geos_lite.py, the restatement of the GEOS operations the path uses - `GEOMETRY` says which one
is active.  Neither this container nor the GPU boxes of rounds 1-2 have shapely (probed through gpurun,
DESIGN.md section 5), so the committed fixtures were produced on geos_lite; tests/test_geos_conformance.py
compares the two whenever shapely is present.  matplotlib / rtree / openpyxl / geopandas are inert stubs
(the hot path never calls them).
"""
import importlib
import os
import random
import sys
import types
from types import SimpleNamespace
from unittest import mock

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFERENCE = "/root/reference"
sys.path.insert(0, REPO)

from oracle import geos_lite  # noqa: E402

try:   # the real thing first (AAC_FORCE_GEOS_LITE=1 keeps the restatement, e.g. to regenerate the committed fixtures)
    if os.environ.get("AAC_FORCE_GEOS_LITE"):
        raise ImportError("geos_lite forced")
    import shapely as _shapely
    import shapely.geometry as _geom
    GEOMETRY = "shapely " + _shapely.__version__
except ImportError:
    _shapely, _geom = None, geos_lite
    GEOMETRY = "geos_lite (oracle/geos_lite.py)"

VARIANTS = {
    "att": ("MADDPG_ownENV_randomOD_radar_one_model_att",
            "env_simulator_randomOD_radar_sur_drones_oneModel_att"),
    "v2": ("MADDPG_ownENV_randomOD_radar_N_model_use_tdCPA_forV2",
           "env_simulator_randomOD_radar_sur_drones_N_Model_use_tdCPA_forV2"),
    "mm": ("MADDPG_ownENV_randomOD_radar_multipleMap",
           "env_simulator_randomOD_radar_multipleMap"),
}


def _install_stubs():
    if _shapely is None:
        geos_lite.install_as_shapely()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.markers", "matplotlib.transforms",
                 "matplotlib.patches", "matplotlib.colors", "matplotlib.animation", "rtree", "openpyxl",
                 "geopandas", "jps"):
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__path__ = []
            sys.modules[name] = m


def load_reference_module(variant):
    """Import the reference env module for `variant` with stubs in place; returns the module."""
    _install_stubs()
    d, modname = VARIANTS[variant]
    path = os.path.join(REFERENCE, d)
    # each variant directory has same-named helper modules (jps_straight, ...): isolate them
    for k in [k for k in sys.modules if k.startswith(("Utilities_own", "agent_", "jps_straight", "env_simulator"))]:
        del sys.modules[k]
    sys.path.insert(0, path)
    try:
        mod = importlib.import_module(modname)
    finally:
        sys.path.remove(path)
    return mod


def grid_polys(gmap):
    """[[occupied squares, free squares]] exactly as ATT/grid_env_generation:169-180 builds them."""
    ones, zeros = [], []
    h = gmap.grid_length / 2
    for ix in range(gmap.gx):
        for iy in range(gmap.gy):
            cx, cy = gmap.cell_centre(ix, iy)
            sq = _geom.Point(cx, cy).buffer(h, cap_style=3)
            (ones if gmap.occ[ix, iy] else zeros).append(sq)
    return [[ones, zeros]]


def make_reference_env_mm(maps, n_agents, max_spd=5, acc_range=(-4, 4)):
    """multipleMap variant: the constructor takes per-map collections (MM:42) and the reset plans on cropped
    grid indices looked up through `cropped_coord_match_actual_coord` (MM:330-340)."""
    mod = load_reference_module("mm")
    world, bounds, polys, cropped = {}, {}, {}, {}
    for k, gmap in enumerate(maps):
        world[k] = gmap.occ.astype(float)
        bounds[k] = list(gmap.bound)
        polys[k] = grid_polys(gmap)
        cropped[k] = {}
        for ix in range(gmap.gx):
            for iy in range(gmap.gy):
                cx, cy = gmap.cell_centre(ix, iy)
                c = _geom.Point(cx, cy).buffer(gmap.grid_length / 2, cap_style=3).centroid
                cropped[k][(ix, iy)] = [c.x, c.y]
    env = mod.env_simulator(world, [], maps[0].grid_length, bounds, polys, None, cropped)
    env.current_observable_space = lambda agent: []
    env.create_world(n_agents, 2, 0.95, 0.01, 1, 0.15, 0.05, 0.15, (1800, 1300), max_spd, list(acc_range))
    return env, mod


def make_reference_env(variant, gmap, n_agents, max_spd=5, acc_max=8):
    mod = load_reference_module(variant)
    if variant == "mm":
        raise NotImplementedError("use make_reference_env_mm")
    env = mod.env_simulator(gmap.occ.astype(float), [], gmap.grid_length, list(gmap.bound), grid_polys(gmap), None)
    # current_observable_space (ATT:1607) only fills agent.observableSpace at reset, which
    # cur_state_norm_state_v3 overwrites (ATT:1170) before anything reads it.
    env.current_observable_space = lambda agent: []
    env.create_world(n_agents, 2, 0.95, 0.01, 1, 0.15, 0.05, 0.15, (1800, 1300), max_spd, [-acc_max, acc_max])
    return env, mod


def flat(x):
    return np.asarray(x, dtype=np.float64)


def snapshot_agents(env, variant):
    n = len(env.all_agents)
    out = {
        "pos": np.array([env.all_agents[i].pos for i in range(n)], dtype=np.float64),
        "vel": np.array([env.all_agents[i].vel for i in range(n)], dtype=np.float64),
        "reach": np.array([bool(env.all_agents[i].reach_target) for i in range(n)]),
        "n_wp": np.array([len(getattr(env.all_agents[i], "waypoints", None) or env.all_agents[i].goal) for i in range(n)]),
        "wall": np.array([env.all_agents[i].collide_wall_count for i in range(n)]),
        "heading": np.array([env.all_agents[i].heading for i in range(n)], dtype=np.float64),
    }
    return out


def ref_lines(env):
    n = len(env.all_agents)
    return [np.array(list(env.all_agents[i].ref_line.coords), dtype=np.float64) for i in range(n)]


def pack_state_att(state, n):
    """ATT state list -> (own[N, 6+4(N-1)], radar[N,R], nbr6[N,N-1,6])."""
    own = np.stack([flat(a) for a in state[0]])
    radar = np.stack([flat(a) for a in state[1]])
    nbr6 = np.stack([np.concatenate([flat(b) for b in a], axis=0) for a in state[2]]) if n > 1 else np.zeros((n, 0, 6))
    return own, radar, nbr6


def pack_state_v2(state, n):
    """V2 state list -> (own[N,7], nbr[N,5(N-1)], radar[N,R], nbr6[N,N-1,6])."""
    own = np.stack([flat(a) for a in state[0]])
    nbr = np.stack([flat(a) for a in state[1]])
    radar = np.stack([flat(a) for a in state[2]])
    nbr6 = np.stack([np.concatenate([flat(b) for b in a], axis=0) for a in state[3]])
    return own, nbr, radar, nbr6


def rollout(variant, gmap, n_agents, seed, n_steps, episode_length, action_scale=1.0, max_spd=5, acc_max=8,
            quiet=True, cluster_radius=None, n_rays=18, cluster_min_sep=0.0, policy="random", eval_by_step=False):
    """Seeded rollout of the reference env with auto-reset on (any done | all goal | step cap).

    Returns a dict of stacked per-step arrays plus the per-episode reset data."""
    env, mod = make_reference_env(variant, gmap, n_agents, max_spd, acc_max)
    if n_rays != 18:
        # the reference hard-codes `range(0, 360, 20)` (ATT:1058-1062); shadow the builtin inside the
        # reference module only, so the same source casts 360/n_rays-degree fans
        import builtins
        step = 360 // n_rays
        mod.range = lambda *a: builtins.range(0, 360, step) if a == (0, 360, 20) else builtins.range(*a)
    rng = np.random.default_rng(seed)
    crng = np.random.default_rng(seed + 7919)
    random.seed(seed)
    # eval_by_step: the forV2 evaluation mode "by sorties" (args.mode == 'eval', evaluation_by_episode == False):
    # terminal drones stay where they are, crashes do not end the episode (V2:3729-3734, :3128-3156, :3551-3587)
    args = SimpleNamespace(mode="eval" if eval_by_step else "train")
    by_episode = not eval_by_step
    rec = {k: [] for k in ("actions", "reward", "done", "check_goal", "bbc", "pos", "vel", "reach", "n_wp",
                           "heading", "episode_id", "step_in_ep", "srr")}
    obs_keys = ("own", "radar", "nbr6") if variant == "att" else ("own", "nbr", "radar", "nbr6")
    for k in obs_keys:
        rec["raw_" + k] = []
        rec["norm_" + k] = []
    episodes = []
    devnull = open(os.devnull, "w")

    def do_reset():
        if variant == "att":
            st, nst = env.reset_world(n_agents, None, 0)
        else:
            st, nst = env.reset_world(n_agents, False, 0)
        if cluster_radius is not None:
            # move drones 1.. next to drone 0 so radar / near-drone / collision branches fire;
            # the state is then rebuilt by the reference's own cur_state_norm_state_v3
            c = env.all_agents[0].pos.astype(float)
            placed = [c]
            for i in range(1, n_agents):
                for _try in range(200):
                    off = crng.uniform(-cluster_radius, cluster_radius, size=2)
                    p = c + off
                    ix, iy = gmap.cell_of(p[0], p[1])
                    free = 0 <= ix < gmap.gx and 0 <= iy < gmap.gy and not gmap.occ[ix, iy]
                    if cluster_min_sep <= 0.0 or (free and all(np.hypot(*(p - q)) >= cluster_min_sep for q in placed)):
                        break
                placed.append(p)
                env.all_agents[i].pos = p.copy()
                env.all_agents[i].pre_pos = p.copy()
            for i in range(n_agents):
                env.all_agents[i].surroundingNeighbor = {}
                env.all_agents[i].pre_surroundingNeighbor = {}
            out = env.cur_state_norm_state_v3({}, None if variant == "att" else False)
            st, nst = out[0], out[1]
        snap = snapshot_agents(env, variant)
        lines = ref_lines(env)
        pk = pack_state_att if variant == "att" else pack_state_v2
        episodes.append({"start": snap["pos"].copy(), "heading": snap["heading"].copy(), "ref_lines": lines,
                         "raw": pk(st, n_agents), "norm": pk(nst, n_agents)})

    old_stdout = sys.stdout
    if quiet:
        sys.stdout = devnull
    try:
        do_reset()
        ep_step = 0
        for t in range(n_steps):
            act = rng.uniform(-1.0, 1.0, size=(n_agents, 2)) * action_scale
            if policy == "seek":  # steer at the next waypoint, plus noise: reaches goals, pops waypoints
                for i in range(n_agents):
                    ag = env.all_agents[i]
                    to = np.array(ag.waypoints[0], dtype=float) - ag.pos
                    want = to / max(np.linalg.norm(to), 1e-9) * max_spd * 0.9
                    act[i] = np.clip((want - ag.vel) / (acc_max * 0.5) + 0.25 * act[i], -1.0, 1.0)
            ep_step += 1
            srr = [None] * n_agents
            scr = [[] for _ in range(n_agents)]
            if variant == "att":
                out = env.step(act, ep_step, acc_max, None)
                esh = [None] * n_agents
                rw = env.ss_reward(ep_step, srr, esh, scr, (None, None), True, args)
            else:
                out = env.step(act, ep_step, acc_max, args, by_episode, False)
                rw = env.ss_reward_Mar(ep_step, srr, scr, (None, None), False, args, by_episode)
            st, nst = out[0], out[1]
            reward, done, check_goal, srr_out, _, _, bbc = rw
            pk = pack_state_att if variant == "att" else pack_state_v2
            for k, a in zip(obs_keys, pk(st, n_agents)):
                rec["raw_" + k].append(a)
            for k, a in zip(obs_keys, pk(nst, n_agents)):
                rec["norm_" + k].append(a)
            snap = snapshot_agents(env, variant)
            rec["actions"].append(act)
            rec["reward"].append(np.array([float(r) for r in reward]))
            rec["done"].append(np.array(done, dtype=bool))
            rec["check_goal"].append(np.array(check_goal, dtype=bool))
            rec["bbc"].append(np.array(bbc, dtype=bool))
            rec["srr"].append(np.array([[float(v) for v in s] for s in srr_out]))
            for k in ("pos", "vel", "reach", "n_wp", "heading"):
                rec[k].append(snap[k])
            rec["episode_id"].append(len(episodes) - 1)
            rec["step_in_ep"].append(ep_step)
            all_reach = all(env.all_agents[i].reach_target for i in range(n_agents))
            if ep_step > episode_length or (True in done) or all(check_goal) or all_reach:
                do_reset()
                ep_step = 0
    finally:
        sys.stdout = old_stdout
        devnull.close()
    out = {k: np.stack(v) for k, v in rec.items() if len(v)}
    out["episodes"] = episodes
    return out


def pack_state_mm(state, n):
    """MM state list -> (own[N,6], radar[N,R]).  The third part (legacy neighbour block) is ragged: MM only
    lists neighbours within 17.5 m and never clears the dict (MM:754-770); its actors do not read it
    (MM/maddpg_agent:361-362, :399), so it is not part of the path."""
    own = np.stack([flat(a) for a in state[0]])
    radar = np.stack([flat(a) for a in state[1]])
    return own, radar


def rollout_mm(maps, n_agents, seed, n_steps, episode_length, n_rays=18, policy="random", max_spd=5, quiet=True,
               cluster_radius=None):
    """Seeded rollout of the multipleMap reference: a map is drawn per episode as in MM/ma_main:464."""
    env, mod = make_reference_env_mm(maps, n_agents, max_spd)
    if n_rays != 18:
        import builtins
        step = 360 // n_rays
        mod.range = lambda *a: builtins.range(0, 360, step) if a == (0, 360, 20) else builtins.range(*a)
    rng = np.random.default_rng(seed)
    crng = np.random.default_rng(seed + 7919)
    random.seed(seed)
    keys = ("actions", "reward", "done", "check_goal", "bbc", "pos", "vel", "reach", "wp_mask", "episode_id", "step_in_ep", "map_id",
            "raw_own", "norm_own", "radar", "wall")
    rec = {k: [] for k in keys}
    episodes = []
    state = {"map": 0}

    def wp_mask(i):
        ag = env.all_agents[i]
        line = list(ag.ref_line.coords)
        m = 0
        for g in ag.goal:
            k = [q for q in range(len(line)) if abs(line[q][0] - g[0]) < 1e-9 and abs(line[q][1] - g[1]) < 1e-9]
            m |= 1 << k[0]
        return m

    def do_reset():
        state["map"] = random.randrange(len(maps))
        st, nst = env.reset_world(n_agents, state["map"], 0)
        if cluster_radius is not None:
            gm = maps[state["map"]]
            c = env.all_agents[0].pos.astype(float)
            for i in range(1, n_agents):
                p = c + crng.uniform(-cluster_radius, cluster_radius, size=2)
                env.all_agents[i].pos = p.copy()
                env.all_agents[i].pre_pos = p.copy()
            for i in range(n_agents):
                env.all_agents[i].surroundingNeighbor = {}
                env.all_agents[i].pre_surroundingNeighbor = {}
            out = env.cur_state_norm_state_v3({}, state["map"])
            st, nst = out[0], out[1]
        snap = snapshot_agents(env, "mm")
        episodes.append({"start": snap["pos"].copy(), "heading": snap["heading"].copy(), "ref_lines": ref_lines(env),
                         "raw": pack_state_mm(st, n_agents), "norm": pack_state_mm(nst, n_agents), "map": state["map"]})

    devnull = open(os.devnull, "w")
    old_stdout = sys.stdout
    if quiet:
        sys.stdout = devnull
    try:
        do_reset()
        ep_step = 0
        for t in range(n_steps):
            act = rng.uniform(-1.0, 1.0, size=(n_agents, 2))
            if policy == "seek":   # steer at the first remaining waypoint; coe_a = 20 is hard-coded in MM.step
                for i in range(n_agents):
                    ag = env.all_agents[i]
                    to = np.array(ag.goal[0], dtype=float) - ag.pos
                    want = to / max(np.linalg.norm(to), 1e-9) * max_spd * 0.9
                    act[i] = np.clip((want - ag.vel) / (20 * 0.5) + 0.1 * act[i], -1.0, 1.0)
            ep_step += 1
            srr = [None] * n_agents
            scr = [[] for _ in range(n_agents)]
            esh = [None] * n_agents
            out = env.step(act, ep_step, state["map"])
            rw = env.ss_reward(ep_step, srr, esh, scr, state["map"])
            st, nst = out[0], out[1]
            reward, done, check_goal, _, _, _, bbc = rw
            own, radar = pack_state_mm(st, n_agents)
            nown, _ = pack_state_mm(nst, n_agents)
            snap = snapshot_agents(env, "mm")
            rec["actions"].append(act)
            rec["reward"].append(np.array([float(r) for r in reward]))
            rec["done"].append(np.array(done, dtype=bool))
            rec["check_goal"].append(np.array(check_goal, dtype=bool))
            rec["bbc"].append(np.array(bbc, dtype=bool))
            rec["raw_own"].append(own); rec["norm_own"].append(nown); rec["radar"].append(radar)
            for k in ("pos", "vel", "reach", "wall"):
                rec[k].append(snap[k])
            rec["wp_mask"].append(np.array([wp_mask(i) for i in range(n_agents)], dtype=np.int64))
            rec["episode_id"].append(len(episodes) - 1)
            rec["step_in_ep"].append(ep_step)
            rec["map_id"].append(state["map"])
            all_reach = all(env.all_agents[i].reach_target for i in range(n_agents))
            if ep_step > episode_length or (True in done) or all_reach:
                do_reset()
                ep_step = 0
    finally:
        sys.stdout = old_stdout
        devnull.close()
    out = {k: np.stack(v) for k, v in rec.items() if len(v)}
    out["episodes"] = episodes
    return out
