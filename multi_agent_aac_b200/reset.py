"""Episode reset on the host: random origin/destination, grid path, reference line.

Mirrors `env_simulator.reset_world` (ATT:251-405; ATT = MADDPG_ownENV_randomOD_radar_one_model_att/
env_simulator_randomOD_radar_sur_drones_oneModel_att.py, identical in the forV2 variant V2:257-383):

* per drone draw a start quadrant, a different target quadrant, a start cell and -- after the first
  drone -- redraw until the start is more than 2*protectiveBound from every earlier start
  (ATT:254-270); the draw order is kept so `random.seed(k)` reproduces the reference's episodes;
* 4-connected grid search start->goal (`jps_find_path`, ATT/jps_straight.py:17-70), then drop
  collinear interior points (ATT:321-331); the polyline through the surviving cell centres is the
  drone's `ref_line`, its vertices after the first are `goal` / `waypoints` (ATT:334-347);
* heading points at the first waypoint, velocity starts at zero (ATT:359-372).

Two samplers share the planner: `sample_episode_reference_order` draws from a `random.Random` in the
reference's exact call order (used by RefCompatEnv and the parity tests); `ScenarioBank` draws whole
banks with a NumPy generator for the batched env.
"""
from __future__ import annotations

import math
import random as _random

import numpy as np

from .maps import GridMap

_NEIGHBOURS = ((0, -1), (0, 1), (-1, 0), (1, 0))  # expansion order of ATT/jps_straight.py:46


def plan_path(occ: np.ndarray, start, goal):
    """Best-first grid search with the reference's tie-breaking (ATT/jps_straight.py:17-70).

    The open set is kept in discovery order and the first entry with the smallest f = g + Manhattan
    is expanded; a cell already discovered is never re-queued or re-costed.  Returns the list of
    (ix, iy) cells from start to goal, or None when the goal is unreachable."""
    gx, gy = occ.shape
    n = gx * gy
    UNSEEN, OPEN, CLOSED = 0, 1, 2
    status = [UNSEEN] * n
    g = [0] * n
    f = [0] * n
    parent = [-1] * n
    s = start[0] * gy + start[1]
    t = goal[0] * gy + goal[1]
    frontier = [s]
    status[s] = OPEN
    while frontier:
        k_best = 0
        f_best = f[frontier[0]]
        for k in range(1, len(frontier)):
            if f[frontier[k]] < f_best:
                f_best = f[frontier[k]]
                k_best = k
        cur = frontier.pop(k_best)
        status[cur] = CLOSED
        if cur == t:
            out = []
            while cur != -1:
                out.append((cur // gy, cur % gy))
                cur = parent[cur]
            return out[::-1]
        cx, cy = cur // gy, cur % gy
        for dx, dy in _NEIGHBOURS:
            nx, ny = cx + dx, cy + dy
            if nx < 0 or ny < 0 or nx >= gx or ny >= gy or occ[nx, ny] != 0:
                continue
            c = nx * gy + ny
            if status[c] != UNSEEN:
                continue
            g[c] = g[cur] + 1
            f[c] = g[c] + abs(nx - goal[0]) + abs(ny - goal[1])
            parent[c] = cur
            status[c] = OPEN
            frontier.append(c)
    return None


def prune_collinear(path):
    """Keep the first cell, every turn cell and the last cell (ATT:321-331)."""
    out = [path[0]]
    cur = (path[1][0] - path[0][0], path[1][1] - path[0][1])
    for k in range(2, len(path)):
        nxt = (path[k][0] - path[k - 1][0], path[k][1] - path[k - 1][1])
        if nxt != cur:
            out.append(path[k - 1])
            cur = nxt
    out.append(path[-1])
    return out


def ref_line_cells(gmap: GridMap, start_xy, goal_xy):
    s = gmap.cell_of(*start_xy)
    t = gmap.cell_of(*goal_xy)
    path = plan_path(gmap.occ, s, t)
    if path is None:
        raise ValueError("goal %s unreachable from %s" % (goal_xy, start_xy))
    return prune_collinear(path)


def sample_od_reference_order(rng: _random.Random, pools, n_agents, prot=2.5):
    """Start/goal cell centres for one episode, drawing from `rng` exactly as ATT:254-276 does."""
    starts, goals = [], []
    for _ in range(n_agents):
        si = rng.randint(0, len(pools) - 1)
        left = list(range(0, si)) + list(range(si + 1, len(pools)))
        ti = rng.choice(left)
        start = rng.choice(pools[si])
        if starts:
            while len(starts) < n_agents:
                si = rng.randint(0, len(pools) - 1)
                left = list(range(0, si)) + list(range(si + 1, len(pools)))
                ti = rng.choice(left)
                start = rng.choice(pools[si])
                if all(math.hypot(start[0] - p[0], start[1] - p[1]) > prot * 2 for p in starts):
                    break
        goal = rng.choice(pools[ti])
        starts.append(start)
        goals.append(goal)
    return starts, goals


class Episode:
    """Reset data for one env: start positions, headings and reference lines (cell lists)."""

    def __init__(self, gmap: GridMap, starts, goals):
        self.gmap = gmap
        self.starts = [tuple(float(v) for v in s) for s in starts]
        self.cells = [ref_line_cells(gmap, s, t) for s, t in zip(starts, goals)]
        self.lines = [np.array([gmap.cell_centre(ix, iy) for ix, iy in c], dtype=np.float64) for c in self.cells]
        self.headings = [math.atan2(l[1][1] - l[0][1], l[1][0] - l[0][0]) for l in self.lines]

    @property
    def n_agents(self):
        return len(self.starts)


def sample_episode_reference_order(rng: _random.Random, gmap: GridMap, n_agents, prot=2.5) -> Episode:
    starts, goals = sample_od_reference_order(rng, gmap.target_pools(), n_agents, prot)
    return Episode(gmap, starts, goals)


class ScenarioBank:
    """S pre-planned episodes for one map, packed for upload (see include/aac_env.h AacBank).

    `cells[s, i, k]` = ix * 256 + iy of vertex k of drone i's reference line, `w[s, i]` = vertex
    count.  Envs pick scenario hash(global_env_id, episode_index) mod S on the device, so the
    episodes an env sees do not depend on how envs are sharded over GPUs."""

    def __init__(self, gmap: GridMap, n_agents, n_scenarios, w_max=32, seed=0, prot=2.5):
        self.gmap, self.n_agents, self.w_max = gmap, n_agents, w_max
        rng = np.random.default_rng(seed)
        pools = gmap.target_pools()
        pool_arr = [np.array(p, dtype=np.int64) for p in pools]
        cells = np.zeros((n_scenarios, n_agents, w_max), dtype=np.uint16)
        w = np.zeros((n_scenarios, n_agents), dtype=np.uint8)
        cache = {}
        for s in range(n_scenarios):
            starts = []
            for i in range(n_agents):
                while True:
                    si = int(rng.integers(0, 4))
                    ti = int(rng.choice([q for q in range(4) if q != si]))
                    st = tuple(int(v) for v in pool_arr[si][rng.integers(0, len(pool_arr[si]))])
                    if all(math.hypot(st[0] - p[0], st[1] - p[1]) > prot * 2 for p in starts):
                        break
                gl = tuple(int(v) for v in pool_arr[ti][rng.integers(0, len(pool_arr[ti]))])
                starts.append(st)
                key = (st, gl)
                if key not in cache:
                    cache[key] = ref_line_cells(gmap, st, gl)
                c = cache[key]
                if len(c) > w_max:
                    raise ValueError("reference line with %d vertices exceeds w_max=%d" % (len(c), w_max))
                w[s, i] = len(c)
                for k, (ix, iy) in enumerate(c):
                    cells[s, i, k] = ix * 256 + iy
        self.cells, self.w = cells, w

    @property
    def n_scenarios(self):
        return self.cells.shape[0]


class MultiMapBank:
    """Scenario bank over several maps (multipleMap variant): scenario s is planned on maps[map_id[s]], maps
    are drawn uniformly as MM/ma_main:464 does per episode.  Same packing as ScenarioBank plus `map_id`."""

    def __init__(self, maps, n_agents, n_scenarios, w_max=32, seed=0, prot=2.5):
        self.maps, self.n_agents, self.w_max = list(maps), n_agents, w_max
        rng = np.random.default_rng(seed)
        self.map_id = rng.integers(0, len(self.maps), size=n_scenarios).astype(np.int32)
        self.cells = np.zeros((n_scenarios, n_agents, w_max), dtype=np.uint16)
        self.w = np.zeros((n_scenarios, n_agents), dtype=np.uint8)
        for k, gmap in enumerate(self.maps):
            idx = np.nonzero(self.map_id == k)[0]
            if len(idx) == 0:
                continue
            sub = ScenarioBank(gmap, n_agents, len(idx), w_max=w_max, seed=seed * 1000 + k, prot=prot)
            self.cells[idx] = sub.cells
            self.w[idx] = sub.w

    @property
    def n_scenarios(self):
        return self.cells.shape[0]


def plan_paths_device(gmap: GridMap, pairs, w_max=32):
    """Reference lines for many (start cell, goal cell) pairs in one launch of the library's device planner
    (`aac_plan_paths_device`: the reference's search and pruning, ATT/jps_straight.py:17-70 + ATT:321-331, one warp per
    pair).  pairs: int array [n, 4] = (sx, sy, tx, ty).  Returns (cells uint16[n, w_max] codes ix<<8|iy, length int32[n])."""
    import ctypes as C
    from . import _capi as K
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 4)
    code = np.ascontiguousarray(np.stack([pairs[:, 0] * 256 + pairs[:, 1], pairs[:, 2] * 256 + pairs[:, 3]], axis=1).astype(np.uint16))
    occ = np.ascontiguousarray(gmap.occ, dtype=np.uint8)
    cells = np.zeros((len(pairs), w_max), dtype=np.uint16)
    length = np.zeros(len(pairs), dtype=np.int32)
    K.check(K.lib().aac_plan_paths_device(occ.ctypes.data, gmap.gx, gmap.gy, code.ctypes.data, len(pairs), cells.ctypes.data,
                                          length.ctypes.data, w_max, None), "aac_plan_paths_device")
    return cells, length


class OdTable:
    """Origin / destination table of one map (include/aac_env.h AacOdTable): the four quadrant pools of free cells
    (ATT:154-197) and the pruned grid path between every start / goal pair in different quadrants, planned by the
    library's planner (the reference's search and tie-breaking, ATT/jps_straight.py:17-70): `planner="host"` calls
    `aac_plan_path` pair by pair, `planner="device"` plans every pair in one launch (`aac_plan_paths_device`, identical
    results).  With a table installed the device draws origins and destinations itself at every reset (ATT:254-276).
    `paths=False` builds the pools only: the reference line of every episode is then searched on the device when the episode
    starts, by the warp that re-initialises the env (`reset_world`'s per-episode `jps_find_path`, ATT:317) - same planner,
    same episodes bit for bit, no P^2 table (an unreachable goal or a path of more than w_max vertices cannot raise there: the
    line falls back to start -> goal and `read_stats()[10]` counts it)."""

    def __init__(self, gmap: GridMap, w_max=32, planner="host", paths=True):
        from . import _capi as K
        lib = K.lib()
        assert planner in ("host", "device")
        self.gmap, self.w_max, self.planner, self.has_paths = gmap, w_max, planner, bool(paths)
        pools = gmap.target_pools()
        cells, pool_off, quad = [], [0], []
        for q in range(4):
            for cx, cy in pools[q]:
                cells.append(gmap.cell_of(cx, cy))
                quad.append(q)
            pool_off.append(len(cells))
        P = len(cells)
        self.n_cells, self.pool_off = P, np.array(pool_off, dtype=np.int32)
        self.cell_code = np.array([ix * 256 + iy for ix, iy in cells], dtype=np.uint16)
        if not paths:
            self.path_off = self.path_len = self.path_cells = None
            return
        self.path_off = np.zeros(P * P, dtype=np.uint32)
        self.path_len = np.zeros(P * P, dtype=np.uint8)
        quad = np.array(quad)
        ss, tt = np.nonzero(quad[:, None] != quad[None, :])          # row-major: s outer, t inner
        ca = np.array(cells, dtype=np.int64).reshape(-1, 2)
        if planner == "device":
            found, n_of = plan_paths_device(gmap, np.concatenate([ca[ss], ca[tt]], axis=1), w_max)
        else:
            occ = np.ascontiguousarray(gmap.occ, dtype=np.uint8)
            found, n_of = np.zeros((len(ss), w_max), dtype=np.uint16), np.zeros(len(ss), dtype=np.int32)
            for k, (s, t) in enumerate(zip(ss, tt)):
                n_of[k] = lib.aac_plan_path(occ.ctypes.data, gmap.gx, gmap.gy, cells[s][0], cells[s][1], cells[t][0], cells[t][1],
                                            found[k].ctypes.data, w_max)
        for k in np.nonzero(n_of <= 0)[0]:
            if n_of[k] == 0:
                raise ValueError("cell %s is unreachable from %s" % (cells[tt[k]], cells[ss[k]]))
            raise ValueError("a reference line needs more than w_max=%d vertices" % w_max)
        pad = (n_of.astype(np.int64) + 7) // 8 * 8                   # the device copies paths in 16-byte chunks
        off = np.concatenate([[0], np.cumsum(pad)])
        self.path_off[ss * P + tt] = off[:-1]
        self.path_len[ss * P + tt] = n_of
        self.path_cells = np.zeros(max(int(off[-1]), 1), dtype=np.uint16)
        col = np.arange(w_max)
        keep = col[None, :] < n_of[:, None]
        self.path_cells[(off[:-1, None] + col[None, :])[keep]] = found[keep]

    def path(self, s, t):
        k = s * self.n_cells + t
        return self.path_cells[self.path_off[k]:self.path_off[k] + self.path_len[k]]
