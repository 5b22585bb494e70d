"""Drop-in boundary: the reference's `env_simulator` method surface on top of the CUDA env.

The reference has no plugin / FFI layer; its hot path is called as plain Python methods from
`ma_main_*`:

    env = env_simulator(world_map, building_polygons, grid_length, bound, allGridPoly, agentConfig)   ATT:41
    env.create_world(total_agentNum, n_actions, gamma, tau, target_update, largest_Nsigma,
                     smallest_Nsigma, ini_Nsigma, max_xy, max_spd, acc_range)                          ATT:84
    state, norm_state = env.reset_world(total_agentNum, actor_dim, show)                              ATT:199
    next_state, norm_next_state, *plot_objects = env.step(actions, ts, acc_max, actor_dim)            ATT:2627
    reward, done, check_goal, srr, esh, scr, bbc = env.ss_reward(ts, srr, esh, scr, xy, flag, args)   ATT:2105
  forV2:
    env.reset_world(total_agentNum, full_observable_critic_flag, show)                                V2:201
    env.step(actions, ts, acc_max, args, evaluation_by_episode, full_observable_critic_flag)          V2:3703
    env.ss_reward_Mar(ts, srr, scr, xy, full_observable_critic_flag, args, evaluation_by_episode)     V2:2995

`RefCompatEnv` keeps those names, argument meanings and the nested list-of-arrays return structure for a
single env (E = 1), so the reference's training loop can swap its import.  `step` launches the fused
CUDA step (which already contains the reward pass) and `ss_reward*` returns that launch's reward / done
/ flags.  Plot-only members of the return tuples (shapely objects) are returned as empty lists.
Reset follows the reference's draw order from Python's global `random` (reset.py), so `random.seed(k)`
gives the same origins, destinations and reference lines as the reference.

ATT = MADDPG_ownENV_randomOD_radar_one_model_att/env_simulator_randomOD_radar_sur_drones_oneModel_att.py
V2  = MADDPG_ownENV_randomOD_radar_N_model_use_tdCPA_forV2/env_simulator_randomOD_radar_sur_drones_N_Model_use_tdCPA_forV2.py
"""
from __future__ import annotations

import random
import re

import numpy as np
import torch

from . import _capi as K
from .env import BatchedDroneEnv, preset
from .maps import GridMap
from .reset import sample_episode_reference_order


class AgentView:
    """The attributes of the reference's `Agent` record that callers read (ATT/agent:14-53)."""

    def __init__(self, idx, max_spd):
        self.agent_name = "agent_%s" % idx
        self.maxSpeed = max_spd
        self.protectiveBound = 2.5
        self.detectionRange = 30
        self.pos = self.pre_pos = self.ini_pos = None
        self.vel = self.pre_vel = None
        self.heading = None
        self.goal = None
        self.waypoints = None
        self.ref_line = None
        self.reach_target = False
        self.collide_wall_count = 0
        self.bound_collision = self.building_collision = self.drone_collision = False
        self.observableSpace = []


class _Normalizer:
    """NormalizeData (ATT/Utilities_own:554-607) -- host-side helper kept for callers that read env.normalizer."""

    def __init__(self, x_min_max, y_min_max, spd_max, acc_range):
        self.dis_min_x, self.dis_max_x = x_min_max
        self.dis_min_y, self.dis_max_y = y_min_max
        self.spd_max, self.acc_min, self.acc_max = spd_max, acc_range[0], acc_range[1]
        self.x_scale = 2.0 / (self.dis_max_x - self.dis_min_x)
        self.y_scale = 2.0 / (self.dis_max_y - self.dis_min_y)

    def nmlz_pos(self, pos_c):
        return np.array([2 * ((pos_c[0] - self.dis_min_x) / (self.dis_max_x - self.dis_min_x)) - 1,
                         2 * ((pos_c[1] - self.dis_min_y) / (self.dis_max_y - self.dis_min_y)) - 1])

    def scale_pos(self, pos_c):
        return np.array([-1 + (pos_c[0] - self.dis_min_x) * self.x_scale, -1 + (pos_c[1] - self.dis_min_y) * self.y_scale])

    def nmlz_vel(self, cur_vel):
        return np.array([cur_vel[0] / self.spd_max, cur_vel[1] / self.spd_max])

    def reverse_nmlz_pos(self, norm):
        return np.array([(norm[0] + 1) / 2 * (self.dis_max_x - self.dis_min_x) + self.dis_min_x,
                         (norm[1] + 1) / 2 * (self.dis_max_y - self.dis_min_y) + self.dis_min_y])


class RefCompatEnv:
    variant = "att"

    def __init__(self, world_map, building_polygons, grid_length, bound, allGridPoly=None, agentConfig=None, n_rays=18,
                 device="cuda:0"):
        self.world_map_2D = np.asarray(world_map)
        self.buildingPolygons = building_polygons
        self.gridlength = grid_length
        self.bound = list(bound)
        self.world_map_2D_polyList = allGridPoly
        self.agentConfig = agentConfig
        self.time_step = 0.5
        self.global_time = 0.0
        self.all_agents = None
        self.normalizer = None
        self.gmap = GridMap(list(bound), int(grid_length), (self.world_map_2D != 0).astype(np.uint8))
        self._n_rays = n_rays
        self._device = device
        self._env = None
        self._last = None

    # ---- create_world (ATT:84-197) ---------------------------------------------------------------
    def create_world(self, total_agentNum, n_actions, gamma, tau, target_update, largest_Nsigma, smallest_Nsigma, ini_Nsigma,
                     max_xy, max_spd, acc_range):
        self.max_agent_num = total_agentNum
        self._max_spd, self._acc_range = float(max_spd), list(acc_range)
        self.normalizer = _Normalizer([self.bound[0], self.bound[1]], [self.bound[2], self.bound[3]], max_spd, acc_range)
        self.all_agents = {i: AgentView(i, max_spd) for i in range(total_agentNum)}
        cfg = preset("att" if self.variant == "att" else "tdcpa_v2", n_envs=1, n_agents=total_agentNum, n_rays=self._n_rays,
                     w_max=32, sum_reward=False, vmax=float(max_spd), acc_max=float(acc_range[1]),
                     out_flags=K.OUT_RAW | K.OUT_NBR6 | K.OUT_PARTS)
        self._env = BatchedDroneEnv(cfg, self.gmap, device=self._device)
        self._pools = self.gmap.target_pools()
        self.target_pool = self._pools

    # ---- reset_world (ATT:199-511) ---------------------------------------------------------------
    def reset_world(self, total_agentNum, actor_dim_or_flag=None, show=0):
        self.global_time = 0.0
        ep = sample_episode_reference_order(random, self.gmap, total_agentNum)
        self._episode = ep
        self._env.set_episode(0, ep.starts, ep.lines, ep.headings)
        for i, ag in self.all_agents.items():
            ag.pos = np.array(ep.starts[i], dtype=np.float64)
            ag.ini_pos = ag.pos.copy()
            ag.pre_pos = ag.pos.copy()
            ag.vel = np.zeros(2)
            ag.pre_vel = np.zeros(2)
            ag.heading = ep.headings[i]
            ag.goal = [list(p) for p in ep.lines[i][1:]]
            ag.waypoints = [list(p) for p in ep.lines[i][1:]]
            ag.ref_line = ep.lines[i].copy()
            ag.reach_target = False
            ag.collide_wall_count = 0
            ag.bound_collision = ag.building_collision = ag.drone_collision = False
        self._env.observe()
        return self._pack_states()

    # ---- step (ATT:2627-2815) --------------------------------------------------------------------
    def step(self, actions, current_ts, acc_max=None, *unused):
        a = np.asarray([np.asarray(x, dtype=np.float32).reshape(2) for x in actions], dtype=np.float32)
        if a.shape != (len(self.all_agents), 2):
            raise ValueError("actions must hold one (ax, ay) pair per agent")
        for ag in self.all_agents.values():
            ag.pre_pos, ag.pre_vel = ag.pos.copy(), ag.vel.copy()
        t = torch.from_numpy(a[None]).to(self._env.device).contiguous()
        _, reward, done, info = self._env.step(t)
        self.global_time += self.time_step
        o = {k: v[0].cpu().numpy() for k, v in self._env.out.items()}
        s = self._env.agent_state()
        for i, ag in self.all_agents.items():
            ag.pos, ag.vel, ag.heading = s["pos"][0, i].copy(), s["vel"][0, i].copy(), float(s["heading"][0, i])
            ag.reach_target = bool(s["reach"][0, i])
            ag.collide_wall_count = int(s["wall_cnt"][0, i])
            ag.waypoints = [list(p) for p in self._episode.lines[i][1 + int(s["wp_cur"][0, i]):]]
            ag.observableSpace = o["radar"][i].astype(np.float64)
            vf = int(s["vflags"][0, i])
            ag.bound_collision, ag.building_collision, ag.drone_collision = bool(vf & 1), bool(vf & 2), bool(vf & 4)
        self._last = o
        st, nst = self._pack_states()
        return st, nst, [], [], [], [], [], []

    # ---- ss_reward (ATT:2105-2618) ---------------------------------------------------------------
    def ss_reward(self, current_ts, step_reward_record, eps_status_holder, step_collision_record, xy, full_observable_critic_flag,
                  args=None):
        reward, done, check_goal, bbc = self._reward_core(full_observable_critic_flag)
        o = self._last
        for i in range(len(self.all_agents)):
            p = o["parts"][i].astype(np.float64)
            step_reward_record[i] = [0.0 * p[4], p[0]]  # [dist_to_ref_line (coef_ref_line = 0, ATT:2368), dist_to_goal]
            if eps_status_holder is not None and eps_status_holder[i] is not None and hasattr(eps_status_holder[i], "append"):
                spd = float(np.linalg.norm(self.all_agents[i].vel))
                eps_status_holder[i].append([p[5], p[0], p[4], 0.0, p[2], p[3], spd, 0.0, 0.0])
            if step_collision_record is not None and step_collision_record[i] is not None and hasattr(step_collision_record[i], "append"):
                step_collision_record[i].append([0.0, 0, 0, int(self.all_agents[i].collide_wall_count > 0), 0, 0])
        return reward, done, check_goal, step_reward_record, eps_status_holder, step_collision_record, bbc

    def _reward_core(self, sum_flag):
        if self._last is None:
            raise RuntimeError("ss_reward called before step")
        o = self._last
        n = len(self.all_agents)
        r = o["reward"].astype(np.float64)
        if sum_flag:  # reward = [np.sum(reward) for _ in reward] (ATT:2602-2603)
            r = np.full(n, r.sum())
        reward = [np.array(v) for v in r]
        done = [bool(v) for v in o["done"]]
        check_goal = [bool(v) for v in o["check_goal"]]
        bbc = [bool(v) for v in o["bbc"]]
        return reward, done, check_goal, bbc

    # ---- nested-list state (ATT:1445-1493) -------------------------------------------------------
    def _pack_states(self):
        o = {k: v[0].cpu().numpy().astype(np.float64) for k, v in self._env.out.items() if k.startswith(("raw_", "norm_", "radar"))}
        n = len(self.all_agents)
        radar = [o["radar"][i] for i in range(n)]
        raw = [[o["raw_own"][i] for i in range(n)], radar, [[o["raw_nbr6"][i, k][None] for k in range(n - 1)] for i in range(n)]]
        norm = [[o["norm_own"][i] for i in range(n)], radar, [[o["norm_nbr6"][i, k][None] for k in range(n - 1)] for i in range(n)]]
        return raw, norm

    @staticmethod
    def agent_index(name):
        """`agent_<k>` -> k; the reference raises ValueError for a name without digits (ATT:861-866)."""
        m = re.search(r"\d+(\.\d+)?", name)
        if not m:
            raise ValueError("No number found in string")
        return int(m.group())


class RefCompatEnvV2(RefCompatEnv):
    variant = "v2"

    def step(self, actions, current_ts, acc_max=None, args=None, evaluation_by_episode=True, full_observable_critic_flag=False):
        # evaluation "by sorties" (V2:3729-3734): the underlying env is rebuilt once in that mode, keeping its state
        by_step = args is not None and getattr(args, "mode", "train") == "eval" and not evaluation_by_episode
        if by_step != bool(self._env.cfg.eval_by_step):
            import dataclasses
            sd = self._env.state_dict()
            self._env.close()
            self._env = BatchedDroneEnv(dataclasses.replace(self._env.cfg, eval_by_step=by_step), self.gmap, device=self._device)
            self._env.load_state_dict(sd)
        return super().step(actions, current_ts, acc_max)

    def ss_reward_Mar(self, current_ts, step_reward_record, step_collision_record, xy, full_observable_critic_flag, args=None,
                      evaluation_by_episode=True):
        reward, done, check_goal, bbc = self._reward_core(full_observable_critic_flag)
        o = self._last
        for i in range(len(self.all_agents)):
            step_reward_record[i] = [0.0, float(o["reward"][i])]  # [dist_to_ref_line, rew] (V2:3649)
            if step_collision_record is not None and step_collision_record[i] is not None and hasattr(step_collision_record[i], "append"):
                step_collision_record[i].append([0.0, 0, 0, int(self.all_agents[i].building_collision), 0, 0])
        return reward, done, check_goal, step_reward_record, None, step_collision_record, bbc

    def _pack_states(self):
        o = {k: v[0].cpu().numpy().astype(np.float64) for k, v in self._env.out.items() if k.startswith(("raw_", "norm_", "radar"))}
        n = len(self.all_agents)
        radar = [o["radar"][i] for i in range(n)]
        raw = [[o["raw_own"][i] for i in range(n)], [o["raw_nbr"][i] for i in range(n)], radar,
               [[o["raw_nbr6"][i, k][None] for k in range(n - 1)] for i in range(n)]]
        norm = [[o["norm_own"][i] for i in range(n)], [o["norm_nbr"][i] for i in range(n)], radar,
                [[o["norm_nbr6"][i, k][None] for k in range(n - 1)] for i in range(n)]]
        return raw, norm


class RefCompatEnvMM:
    """radar_multipleMap drop-in (MM:42): the constructor takes per-map collections, `reset_world / step / ss_reward`
    take the episode's `random_map_idx` (drawn by the caller, MM/ma_main:464).  States are `[p1, p2, p3]` with
    p1 = own block (6), p2 = radar; p3 (the reference's ragged, never-cleared legacy neighbour block, which its
    actors do not read, MM/maddpg_agent:361-399) is returned as one zero row per drone."""

    def __init__(self, world_map_collection, building_polygons, grid_length, bound_collection, allGridPoly_collection=None,
                 agentConfig=None, cropped_coord_match_actual_coord=None, n_rays=18, device="cuda:0"):
        keys = sorted(world_map_collection.keys())
        self._keys = {k: i for i, k in enumerate(keys)}
        self.world_map_2D_collection = world_map_collection
        self.bound_collection = bound_collection
        self.world_map_2D_polyList_collection = allGridPoly_collection
        self.cropped_coord_match_actual_coord = cropped_coord_match_actual_coord
        self.buildingPolygons, self.gridlength, self.agentConfig = building_polygons, grid_length, agentConfig
        self.maps = [GridMap(list(bound_collection[k]), int(grid_length), (np.asarray(world_map_collection[k]) != 0).astype(np.uint8))
                     for k in keys]
        self.time_step, self.global_time = 0.5, 0.0
        self.all_agents, self.normalizer = None, None
        self._n_rays, self._device, self._env, self._last = n_rays, device, None, None

    def create_world(self, total_agentNum, n_actions, gamma, tau, target_update, largest_Nsigma, smallest_Nsigma, ini_Nsigma,
                     max_xy, max_spd, acc_range):
        self._max_spd, self._acc_range = float(max_spd), list(acc_range)
        self.all_agents = {i: AgentView(i, max_spd) for i in range(total_agentNum)}
        cfg = preset("multimap", n_envs=1, n_agents=total_agentNum, n_rays=self._n_rays, w_max=32, vmax=float(max_spd),
                     out_flags=K.OUT_RAW | K.OUT_PARTS)
        self._env = BatchedDroneEnv(cfg, self.maps, device=self._device)
        self.target_pool_collection = {k: self.maps[i].target_pools() for k, i in self._keys.items()}

    def reset_world(self, total_agentNum, random_map_idx, show=0):
        self.global_time = 0.0
        m = self._keys[random_map_idx]
        gmap = self.maps[m]
        b = gmap.bound
        self.normalizer = _Normalizer([b[0], b[1]], [b[2], b[3]], self._max_spd, self._acc_range)   # rebuilt per episode (MM:257-261)
        ep = sample_episode_reference_order(random, gmap, total_agentNum)
        self._episode = ep
        self._env.set_episode(0, ep.starts, ep.lines, ep.headings, map_id=m)
        for i, ag in self.all_agents.items():
            ag.pos = np.array(ep.starts[i], dtype=np.float64)
            ag.ini_pos, ag.pre_pos = ag.pos.copy(), ag.pos.copy()
            ag.vel, ag.pre_vel = np.zeros(2), np.zeros(2)
            ag.heading = ep.headings[i]
            ag.goal = [list(p) for p in ep.lines[i][1:]]
            ag.ref_line = ep.lines[i].copy()
            ag.reach_target, ag.collide_wall_count = False, 0
        self._env.observe()
        return self._pack_states()

    def step(self, actions, current_ts, random_map_idx):
        a = np.asarray([np.asarray(x, dtype=np.float32).reshape(2) for x in actions], dtype=np.float32)
        for ag in self.all_agents.values():
            ag.pre_pos, ag.pre_vel = ag.pos.copy(), ag.vel.copy()
        self._env.step(torch.from_numpy(a[None]).to(self._env.device).contiguous())
        self.global_time += self.time_step
        o = {k: v[0].cpu().numpy() for k, v in self._env.out.items()}
        s = self._env.agent_state()
        for i, ag in self.all_agents.items():
            ag.pos, ag.vel = s["pos"][0, i].copy(), s["vel"][0, i].copy()
            ag.reach_target = bool(s["reach"][0, i])
            ag.collide_wall_count = int(s["wall_cnt"][0, i])
            mask = int(s["wp_mask"][0, i])
            ag.goal = [list(p) for k, p in enumerate(self._episode.lines[i]) if (mask >> k) & 1]
            ag.observableSpace = o["radar"][i].astype(np.float64)
        self._last = o
        st, nst = self._pack_states()
        return st, nst, [], [], [], [], [], []

    def ss_reward(self, current_ts, step_reward_record, eps_status_holder, step_collision_record, random_map_idx):
        if self._last is None:
            raise RuntimeError("ss_reward called before step")
        o = self._last
        n = len(self.all_agents)
        for i in range(n):
            step_reward_record[i] = [float(o["parts"][i][1]), float(o["parts"][i][0])]   # [dist_to_ref_line, dist_to_goal] (MM:2003)
            if step_collision_record is not None and step_collision_record[i] is not None and hasattr(step_collision_record[i], "append"):
                step_collision_record[i].append([0.0, 0, 0, int(o["branch"][i] == 1), 0, 0])
        reward = [np.array(float(v)) for v in o["reward"]]
        done = [bool(v) for v in o["done"]]
        check_goal = [bool(v) for v in o["check_goal"]]
        bbc = [bool(v) for v in o["bbc"][:2]]
        return reward, done, check_goal, step_reward_record, eps_status_holder, step_collision_record, bbc

    def _pack_states(self):
        o = {k: v[0].cpu().numpy().astype(np.float64) for k, v in self._env.out.items() if k in ("raw_own", "norm_own", "radar")}
        n = len(self.all_agents)
        radar = [o["radar"][i] for i in range(n)]
        p3 = [[np.zeros((1, 6))] for _ in range(n)]
        return [[o["raw_own"][i] for i in range(n)], radar, p3], [[o["norm_own"][i] for i in range(n)], radar, p3]


def env_simulator(variant, *a, **kw):
    """Factory named like the reference class: env_simulator("att" | "v2" | "mm", world_map, ...)."""
    return {"att": RefCompatEnv, "v2": RefCompatEnvV2, "mm": RefCompatEnvMM}[variant](*a, **kw)
