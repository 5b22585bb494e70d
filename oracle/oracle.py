"""ctypes front-end of the float64 CPU oracle (oracle/aac_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
The product package multi_agent_aac_b200 must never import this module.

Parity status: control flow / layouts pinned against the unmodified reference run through
oracle/geos_lite.py (tests/golden/); GEOS geometry primitives "parity unpinned" (see aac_oracle.c).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

VAR_ATT, VAR_V2, VAR_MM = 0, 1, 2
RADAR_MIN, RADAR_LAST_HIT = 0, 1
VARIANT_IDS = {"att": VAR_ATT, "v2": VAR_V2, "mm": VAR_MM}


class OracleCfg(C.Structure):
    _fields_ = [("variant", C.c_int32), ("n_agents", C.c_int32), ("n_rays", C.c_int32), ("w_max", C.c_int32),
                ("radar_mode", C.c_int32), ("sum_reward", C.c_int32), ("gx", C.c_int32), ("gy", C.c_int32),
                ("dt", C.c_double), ("vmax", C.c_double), ("acc_max", C.c_double), ("prot", C.c_double),
                ("ray_len", C.c_double), ("goal_r", C.c_double), ("bound", C.c_double * 4),
                ("x0c", C.c_double), ("y0c", C.c_double), ("cell", C.c_double), ("eval_by_step", C.c_int32), ("radar_targets", C.c_int32),
                ("n_nbr_obs", C.c_int32), ("n_clouds", C.c_int32), ("clouds", (C.c_double * 6) * 8)]


_STATE_FIELDS = ["pos", "vel", "heading", "reach", "wp_cur", "wall_cnt", "prev_nn", "vflags", "ref_line", "ref_w", "wp_mask", "ep_step"]
_OUT_FIELDS = ["raw_own", "norm_own", "raw_nbr", "norm_nbr", "radar", "radar_min", "radar_hit", "raw_nbr6",
               "norm_nbr6", "nbr_order", "tcpa", "conflict", "reward", "done", "check_goal", "bbc", "parts",
               "margin", "branch", "tcpa_min", "cloud_contact"]


class _State(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _STATE_FIELDS]


class _Out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _OUT_FIELDS]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "aac_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        assert _LIB.oracle_sizeof_cfg() == C.sizeof(OracleCfg)
    return _LIB


def own_dim(variant, n):
    return {"att": 6 + 4 * (n - 1), "v2": 7, "mm": 6}[variant]


class OracleEnv:
    """E independent envs, float64, stepping through oracle_step_maps().  `gmap` is one GridMap or, for the
    multipleMap variant, a list of them (env e lives on maps[env_map[e]])."""

    def __init__(self, variant, gmap, n_envs, n_agents, n_rays=18, w_max=32, radar_mode=None, sum_reward=None,
                 vmax=5.0, acc_max=None, eval_by_step=False, radar_targets=0, n_nbr_obs=0, clouds=(), prot=2.5):
        """radar_targets / n_nbr_obs / clouds: the later fork's sensor classes (bit0 cells, bit1 boundary segments, bit2 clouds,
        bit3 other aircraft; nearest-N neighbour block; clouds = rows of start x, y, goal x, y, radius, speed)."""
        maps = list(gmap) if isinstance(gmap, (list, tuple)) else [gmap]
        self.variant, self.maps, self.gmap = variant, maps, maps[0]
        self.E, self.N, self.R, self.w_max = n_envs, n_agents, n_rays, w_max
        if radar_mode is None:
            radar_mode = RADAR_LAST_HIT if variant == "v2" else RADAR_MIN
        if sum_reward is None:
            sum_reward = 1 if variant == "att" else 0   # ATT/ma_main:77 vs V2/ma_main:81
        if acc_max is None:
            acc_max = 20.0 if variant == "mm" else 8.0  # MM.step hard-codes coe_a = 20 (MM:2025)
        self.cfgs = (OracleCfg * len(maps))()
        self.occ_stride = max(m.gx * m.gy for m in maps)
        self.occ = np.zeros((len(maps), self.occ_stride), dtype=np.uint8)
        for k, m in enumerate(maps):
            cfg = self.cfgs[k]
            cfg.variant, cfg.n_agents, cfg.n_rays, cfg.w_max = VARIANT_IDS[variant], n_agents, n_rays, w_max
            cfg.radar_mode, cfg.sum_reward, cfg.gx, cfg.gy = radar_mode, sum_reward, m.gx, m.gy
            cfg.dt, cfg.vmax, cfg.acc_max, cfg.prot, cfg.ray_len, cfg.goal_r = 0.5, vmax, acc_max, prot, 15.0, 1.0
            for q in range(4):
                cfg.bound[q] = float(m.bound[q])
            cfg.x0c, cfg.y0c, cfg.cell = m.x0c, m.y0c, float(m.grid_length)
            cfg.eval_by_step = int(bool(eval_by_step))
            cfg.radar_targets, cfg.n_nbr_obs, cfg.n_clouds = int(radar_targets), int(n_nbr_obs), len(clouds)
            for ci, row in enumerate(clouds):
                for q in range(6):
                    cfg.clouds[ci][q] = float(row[q])
            self.occ[k, :m.gx * m.gy] = np.ascontiguousarray(m.occ, dtype=np.uint8).reshape(-1)
        self.cfg = self.cfgs[0]
        self.env_map = np.zeros(n_envs, dtype=np.int32)
        E, N, R, M = n_envs, n_agents, n_rays, n_agents - 1
        f, i = np.float64, np.int32
        self.state = {
            "pos": np.zeros((E, N, 2), f), "vel": np.zeros((E, N, 2), f), "heading": np.zeros((E, N), f),
            "reach": np.zeros((E, N), i), "wp_cur": np.zeros((E, N), i), "wall_cnt": np.zeros((E, N), i),
            "prev_nn": np.full((E, N, 2), -1, i), "vflags": np.zeros((E, N), i),
            "ref_line": np.zeros((E, N, w_max, 2), f), "ref_w": np.full((E, N), 2, i), "wp_mask": np.full((E, N), 2, i),
            "ep_step": np.zeros((E,), i),
        }
        d = own_dim(variant, N)
        Mo = min(n_nbr_obs, M) if n_nbr_obs > 0 else M
        self.out = {
            "raw_own": np.zeros((E, N, d), f), "norm_own": np.zeros((E, N, d), f),
            "raw_nbr": np.zeros((E, N, 5 * Mo), f), "norm_nbr": np.zeros((E, N, 5 * Mo), f),
            "radar": np.zeros((E, N, R), f), "radar_min": np.zeros((E, N, R), f), "radar_hit": np.zeros((E, N, R), i),
            "raw_nbr6": np.zeros((E, N, M, 6), f), "norm_nbr6": np.zeros((E, N, M, 6), f),
            "nbr_order": np.zeros((E, N, M), i), "tcpa": np.zeros((E, N, M, 4), f), "conflict": np.zeros((E, N, 2), i),
            "reward": np.zeros((E, N), f), "done": np.zeros((E, N), i), "check_goal": np.zeros((E, N), i),
            "bbc": np.zeros((E, 4), i), "parts": np.zeros((E, N, 8), f), "margin": np.zeros((E, N), f),
            "branch": np.zeros((E, N), i), "tcpa_min": np.zeros((E, N, 4), f), "cloud_contact": np.zeros((E, N), i),
        }
        self._s = _State(*[self.state[n].ctypes.data for n in _STATE_FIELDS])
        self._o = _Out(*[self.out[n].ctypes.data for n in _OUT_FIELDS])

    # ---- reset ---------------------------------------------------------------------------------
    def set_episode(self, e, starts, lines, headings, map_id=0):
        """Install reset data for env `e` (what reset_world leaves behind, ATT:301-372)."""
        s = self.state
        self.env_map[e] = map_id
        for i in range(self.N):
            w = len(lines[i])
            assert w <= self.w_max
            s["pos"][e, i] = starts[i]
            s["vel"][e, i] = 0.0
            s["heading"][e, i] = headings[i]
            s["ref_line"][e, i, :w] = lines[i]
            s["ref_w"][e, i] = w
            s["wp_mask"][e, i] = (1 << w) - 2      # every vertex after the start is a waypoint (MM:345)
        s["reach"][e] = 0
        s["wp_cur"][e] = 0
        s["wall_cnt"][e] = 0
        s["vflags"][e] = 0
        s["prev_nn"][e] = -1
        s["ep_step"][e] = 0

    def observe(self):
        lib().oracle_observe_maps(self.cfgs, self.occ.ctypes.data_as(C.c_void_p), C.c_int(self.occ_stride),
                                  self.env_map.ctypes.data_as(C.c_void_p), C.c_int(self.E), C.byref(self._s), C.byref(self._o))
        return self.out

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float64)
        assert a.shape == (self.E, self.N, 2)
        lib().oracle_step_maps(self.cfgs, self.occ.ctypes.data_as(C.c_void_p), C.c_int(self.occ_stride),
                               self.env_map.ctypes.data_as(C.c_void_p), C.c_int(self.E), C.byref(self._s),
                               a.ctypes.data_as(C.c_void_p), C.byref(self._o))
        return self.out

    def radar_candidate(self, xy, k, cand_id, map_id=0):
        """Distance at which ray `k` of a drone at `xy` meets radar candidate `cand_id` (a cell index or gx*gy + line), or
        None when the reference's `intersects` is false for it (nan: the ray lies strictly inside the cell)."""
        d = C.c_double(0.0)
        ok = lib().oracle_radar_candidate(C.byref(self.cfgs[map_id]), self.occ[map_id].ctypes.data_as(C.c_void_p), C.c_double(float(xy[0])),
                                          C.c_double(float(xy[1])), C.c_int(int(k)), C.c_int(int(cand_id)), C.byref(d))
        return d.value if ok else None

    def radar_probe(self, pos, i, map_id=0, k_cloud=0):
        """Radar of drone `i` for the position set pos[N,2] -> (stored value[R], true min[R], hit id[R]); k_cloud = steps the
        clouds have moved (sensor configurations)."""
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        out, omin, hit = np.zeros(self.R), np.zeros(self.R), np.zeros(self.R, dtype=np.int32)
        lib().oracle_radar(C.byref(self.cfgs[map_id]), self.occ[map_id].ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p), C.c_int(i),
                           C.c_int(int(k_cloud)), out.ctypes.data_as(C.c_void_p), omin.ctypes.data_as(C.c_void_p), hit.ctypes.data_as(C.c_void_p))
        return out, omin, hit

    def cloud_positions(self, k):
        """[n_clouds, 2] after k steps."""
        out = np.zeros((self.cfg.n_clouds, 2))
        lib().oracle_clouds(C.byref(self.cfg), C.c_int(int(k)), out.ctypes.data_as(C.c_void_p))
        return out
