"""Batched multi-drone environment on one B200: torch tensors for memory and streams, the C ABI
(include/aac_env.h) for everything that computes.

`BatchedDroneEnv` is E independent copies of the reference's `env_simulator` (ATT:40, V2:40): `reset()`
is `reset_world` (ATT:199-511), `step(actions)` is `env.step` + `ss_reward` / `ss_reward_Mar`
(ATT:2627 + :2105, V2:3703 + :2995) and returns the same normalised observation blocks the reference's
actors consume (ATT/maddpg_agent:457-459, V2/maddpg_agent:1243-1245), one row per drone.

ATT = MADDPG_ownENV_randomOD_radar_one_model_att/env_simulator_randomOD_radar_sur_drones_oneModel_att.py
V2  = MADDPG_ownENV_randomOD_radar_N_model_use_tdCPA_forV2/env_simulator_randomOD_radar_sur_drones_N_Model_use_tdCPA_forV2.py
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, replace

import numpy as np
import torch

from . import _capi as K
from .maps import GridMap
from .reset import ScenarioBank

VARIANTS = {"att": K.VARIANT_ATT, "v2": K.VARIANT_V2, "mm": K.VARIANT_MM}


@dataclass(frozen=True)
class EnvConfig:
    """Constants the reference hard-codes in source (SURVEY.md section 5 "config / flags")."""
    variant: str = "att"          # "att" | "v2" | "mm"
    n_envs: int = 1
    n_agents: int = 3             # ATT/ma_main:100
    n_rays: int = 18              # range(0, 360, 20) (ATT:1058)
    w_max: int = 16               # reference-line vertex capacity
    radar_mode: int = K.RADAR_MIN  # V2 only: RADAR_LAST_HIT reproduces V2:1288
    sum_reward: bool = True       # full_observable_critic_flag (ATT/ma_main:77 True, V2/ma_main:81 False)
    episode_length: int = 50      # ATT/ma_main:914 (V2/ma_main:1118: 100)
    dt: float = 0.5               # ATT:60
    vmax: float = 5.0             # ATT/ma_main:150
    acc_max: float = 8.0          # ATT/ma_main:136
    prot: float = 2.5             # ATT/agent:43
    ray_len: float = 15.0         # detectionRange / 2 (ATT:1066)
    goal_r: float = 1.0           # ATT:2266
    out_flags: int = 0            # K.OUT_*
    seed: int = 0
    env_id_base: int = 0
    tile_envs: int = 0
    block_threads: int = 0
    eval_by_step: bool = False    # V2 only: evaluation "by sorties" (args.mode == 'eval', evaluation_by_episode == False)
    autoreset_launches: int = 0   # step(autoreset=True): 1 = fused launch, 2 = step launch + reset launch, 3 = one phased launch, 0 = by batch size
    # the later fork's sensor classes (v2 only; CS = ...forV2_changeskin/env_simulator_...:1379-1506)
    radar_targets: int = 0        # K.TARGET_*: 0 = the variant's own radar, else the fork's true-minimum radar over these classes
    n_nbr_obs: int = 0            # > 0: only the nearest n neighbours enter norm_nbr (use_nearestN_neigh_wRadar / N_neigh, CS:1799-1802)
    clouds: tuple = ()            # rows (start x, y, goal x, y, radius, speed), global metres (cloud.py; CS:598-664, :4667-4681)


def preset(name, **kw) -> EnvConfig:
    """`att`: one_model_att defaults; `tdcpa_v2`: tdCPA_forV2 defaults (last-hit radar as in the source);
    `multimap`: radar_multipleMap defaults; `changeskin_sensors`: tdCPA_forV2 with the later fork's sensor classes."""
    if name == "att":
        base = EnvConfig(variant="att", sum_reward=True, episode_length=50, out_flags=K.OUT_NBR6)
    elif name in ("tdcpa_v2", "v2"):
        base = EnvConfig(variant="v2", sum_reward=False, episode_length=100, radar_mode=K.RADAR_LAST_HIT)
    elif name in ("changeskin_sensors", "cs"):
        # tdCPA_forV2 with the later fork's sensors: boundary segments + clouds + other aircraft, true minimum, the two nearest
        # neighbours, protectiveBound 5 (CS agent file), the fork's two training clouds (CS:600-602; cloud.py: radius 12, 2 m/s)
        base = EnvConfig(variant="v2", sum_reward=False, episode_length=100, radar_mode=K.RADAR_MIN, prot=5.0,
                         radar_targets=K.TARGET_BOUNDS | K.TARGET_CLOUDS | K.TARGET_AIRCRAFT, n_nbr_obs=2,
                         clouds=((30.0, 185.0, 180.0, 80.0, 12.0, 2.0), (30.0, 100.0, 180.0, 30.0, 12.0, 2.0)))
    elif name in ("multimap", "mm"):
        # radar_multipleMap: per-episode maps, true-min radar, coe_a = 20 hard-coded in step (MM:2025),
        # 150-step episodes (MM/ma_main:970)
        base = EnvConfig(variant="mm", sum_reward=False, episode_length=150, acc_max=20.0)
    else:
        raise ValueError("unknown preset %r" % (name,))
    return replace(base, **kw)


def own_dim(variant, n_agents):
    return {"att": 6 + 4 * (n_agents - 1), "v2": 7, "mm": 6}[variant]


class BatchedDroneEnv:
    def __init__(self, cfg: EnvConfig, gmap, device="cuda:0", stream=None):
        """`gmap`: one GridMap, or a list of them for the multipleMap variant (env e lives on maps[map_id[e]])."""
        if not torch.cuda.is_available():
            raise K.AacError("BatchedDroneEnv needs a CUDA device; there is no CPU path")
        self.maps = list(gmap) if isinstance(gmap, (list, tuple)) else [gmap]
        if cfg.variant != "mm" and len(self.maps) != 1:
            raise ValueError("only the multipleMap variant takes several maps")
        gmap = self.maps[0]
        self.cfg, self.gmap = cfg, gmap
        self.device = torch.device(device)
        self.L = K.lib()
        E, N, R, W, M = cfg.n_envs, cfg.n_agents, cfg.n_rays, cfg.w_max, cfg.n_agents - 1
        self.E, self.N, self.R, self.M = E, N, R, M
        D = own_dim(cfg.variant, N)
        self.D = D
        self.origin = gmap.origin
        self.origins = np.array([m.origin for m in self.maps], dtype=np.float64)   # local-frame origin per map
        with torch.cuda.device(self.device):
            c = K.AacConfig(K.ABI_VERSION, VARIANTS[cfg.variant], E, N, R, W, cfg.radar_mode, int(cfg.sum_reward),
                            cfg.episode_length, cfg.out_flags, cfg.tile_envs, cfg.block_threads, cfg.env_id_base, cfg.seed,
                            cfg.dt, cfg.vmax, cfg.acc_max, cfg.prot, cfg.ray_len, cfg.goal_r, int(cfg.eval_by_step), int(cfg.autoreset_launches),
                            int(cfg.radar_targets), int(cfg.n_nbr_obs), len(cfg.clouds))
            for ci, row in enumerate(cfg.clouds):
                for q in range(6):
                    c.clouds[ci][q] = float(row[q])
            h = C.c_void_p()
            K.check(self.L.aac_create(C.byref(c), C.byref(h)), "aac_create")
            self.h = h
            self._set_maps(self.maps)
            dev = self.device
            f32, i32, u8 = torch.float32, torch.int32, torch.uint8
            z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
            self.state = {
                "px": z((E, N), f32), "py": z((E, N), f32), "vx": z((E, N), f32), "vy": z((E, N), f32),
                "heading": z((E, N), f32), "meta": z((E, N), i32), "ref_cells": z((E, N, W), torch.int16),
                "ref_w": torch.full((E, N), 2, dtype=u8, device=dev), "wall_count": z((E, N), i32), "ep_step": z((E,), i32),
                "ep_index": z((E,), i32), "ep_return": z((E,), f32),
            }
            if cfg.variant == "mm":
                self.state["map_id"] = z((E,), i32)
                self.state["wp_mask"] = torch.full((E, N), 2, dtype=i32, device=dev)
            self.state["meta"].fill_(-65536)  # 0xFFFF0000: no previous neighbours
            st = K.AacState(*[self.state[n].data_ptr() if n in self.state else None for n in K.STATE_FIELDS])
            K.check(self.L.aac_bind_state(self.h, C.byref(st)), "aac_bind_state")
            self.out = self.alloc_out(dev)
            self._out_c = self._out_struct(self.out)
        self.stream = stream
        self._bank = None
        self._host_out = None

    # ------------------------------------------------------------------ setup helpers
    def _set_maps(self, maps):
        descs = (K.AacMapDesc * len(maps))()
        occ = np.zeros((len(maps), K.MAP_STRIDE), dtype=np.uint8)
        for k, gmap in enumerate(maps):
            d = descs[k]
            d.gx, d.gy = gmap.gx, gmap.gy
            for q in range(4):
                d.bound[q] = float(gmap.bound[q])
            d.x0c, d.y0c, d.cell = gmap.x0c, gmap.y0c, float(gmap.grid_length)
            d.origin_x, d.origin_y = gmap.origin
            occ[k, :gmap.gx * gmap.gy] = np.ascontiguousarray(gmap.occ, dtype=np.uint8).reshape(-1)
        K.check(self.L.aac_set_maps(self.h, descs, occ.ctypes.data_as(C.c_void_p), len(maps)), "aac_set_maps")

    def alloc_out(self, dev, pin=False):
        cfg, E, N, R, M, D = self.cfg, self.E, self.N, self.R, self.M, self.D
        f32, u8 = torch.float32, torch.uint8
        kw = dict(device=dev)
        if pin:
            kw = dict(device="cpu", pin_memory=True)
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, **kw)
        o = {"norm_own": z((E, N, D), f32), "radar": z((E, N, R), f32), "reward": z((E, N), f32), "done": z((E, N), u8),
             "check_goal": z((E, N), u8), "bbc": z((E, 4), u8), "terminated": z((E,), u8), "tcpa_min": z((E, N, 4), f32)}
        Mo = cfg.n_nbr_obs if cfg.n_nbr_obs > 0 else M      # neighbours per drone in the neighbour block (nearest-N selection)
        if cfg.variant == "v2":
            o["norm_nbr"] = z((E, N, 5 * Mo), f32)
        if cfg.clouds:
            o["cloud_contact"] = z((E, N), u8)
        fl = cfg.out_flags
        if fl & K.OUT_NBR6:
            o["norm_nbr6"] = z((E, N, M, 6), f32)
        if fl & K.OUT_RAW:
            o["raw_own"] = z((E, N, D), f32)
            if cfg.variant == "v2":
                o["raw_nbr"] = z((E, N, 5 * Mo), f32)
            if fl & K.OUT_NBR6:
                o["raw_nbr6"] = z((E, N, M, 6), f32)
        if fl & K.OUT_TCPA_PAIR:
            o["tcpa_pair"] = z((E, N, M, 4), f32)
            o["nbr_order"] = z((E, N, M), torch.int8)
        if fl & K.OUT_RADAR_AUX:
            o["radar_min"] = z((E, N, R), f32)
            o["radar_hit"] = z((E, N, R), torch.int16)
        if fl & K.OUT_PARTS:
            o["parts"] = z((E, N, 8), f32)
            o["branch"] = z((E, N), torch.int8)
        return o

    @staticmethod
    def _out_struct(o, only=None):
        return K.AacOut(*[o[n].data_ptr() if n in o and (only is None or n in only) else None for n in K.OUT_FIELDS])

    def bind_outputs(self, tensors):
        """Make the kernels write the named outputs straight into caller-owned tensors (same shape / dtype / device,
        contiguous) from the next call on, e.g. a slot of a `replay.DeviceReplay` ring: no copy of the step's results."""
        for k, t in tensors.items():
            ref = self.out[k]
            if t.shape != ref.shape or t.dtype != ref.dtype or t.device != ref.device or not t.is_contiguous():
                raise ValueError("output %r must be a contiguous %s tensor of shape %s on %s" % (k, ref.dtype, tuple(ref.shape), ref.device))
            self.out[k] = t
        self._out_c = self._out_struct(self.out)

    def _on_stream(self):
        """Context in which torch ops are ordered with this env's launches (its own stream when one was given)."""
        import contextlib
        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def _stream_ptr(self):
        s = self.stream if self.stream is not None else torch.cuda.current_stream(self.device)
        return C.c_void_p(s.cuda_stream)

    def set_bank(self, bank):
        """`bank`: a ScenarioBank (one map) or a MultiMapBank (scenarios tagged with their map)."""
        assert bank.n_agents == self.N and bank.w_max == self.cfg.w_max
        cells = np.ascontiguousarray(bank.cells, dtype=np.uint16)
        w = np.ascontiguousarray(bank.w, dtype=np.uint8)
        map_id = getattr(bank, "map_id", None)
        if map_id is not None:
            map_id = np.ascontiguousarray(map_id, dtype=np.int32)
        b = K.AacBank(bank.n_scenarios, cells.ctypes.data, w.ctypes.data, map_id.ctypes.data if map_id is not None else None)
        with torch.cuda.device(self.device):
            K.check(self.L.aac_set_bank(self.h, C.byref(b)), "aac_set_bank")
        self._bank = bank
        self.build_radar_table()

    def set_od_tables(self, tables):
        """Install one OdTable per map: resets then draw origins / destinations on the device (ATT:254-276)."""
        tables = list(tables) if isinstance(tables, (list, tuple)) else [tables]
        assert len(tables) == len(self.maps)
        arr = (K.AacOdTable * len(tables))()
        for k, t in enumerate(tables):
            assert t.w_max <= self.cfg.w_max
            arr[k].n_cells = t.n_cells
            for q in range(5):
                arr[k].pool_off[q] = int(t.pool_off[q])
            arr[k].cell_code = t.cell_code.ctypes.data
            if t.path_cells is not None:       # else pools only: paths are searched per episode on the device
                arr[k].path_off, arr[k].path_len, arr[k].path_cells = t.path_off.ctypes.data, t.path_len.ctypes.data, t.path_cells.ctypes.data
                arr[k].n_path_cells = int(t.path_cells.size)
        with torch.cuda.device(self.device):
            K.check(self.L.aac_set_od_tables(self.h, arr, len(tables)), "aac_set_od_tables")
        self._od = tables
        self.build_radar_table()

    def build_radar_table(self, enable=True):
        """reset_world puts every drone on a cell centre (ATT:301-372), so the radar a freshly reset drone sees depends on
        (map, cell) only: it is computed once here - by this library's own observe kernel on drones placed at the centre of
        every cell, so the values are exactly the ones a reset would compute - and installed with aac_set_radar_table; the
        observation of a reset env then looks its ranges up.  Not for the att variant (its radar senses the other drones)."""
        self._rtab = None
        with torch.cuda.device(self.device):
            K.check(self.L.aac_set_radar_table(self.h, None, None, None, None), "aac_set_radar_table")
        if not enable or self.cfg.variant == "att":
            return
        import dataclasses
        N, R, n_maps, S = self.N, self.R, len(self.maps), K.MAP_STRIDE
        per_map = -(-S // N)                                   # probe envs per map, N cell slots each
        cfg = dataclasses.replace(self.cfg, n_envs=n_maps * per_map, tile_envs=0, block_threads=0)
        probe = BatchedDroneEnv(cfg, self.maps if self.cfg.variant == "mm" else self.maps[0], device=self.device, stream=self.stream)
        px = np.zeros((n_maps, per_map * N), dtype=np.float32)
        py = np.zeros_like(px)
        code = np.zeros((n_maps, per_map * N), dtype=np.int64)
        for k, m in enumerate(self.maps):
            # the kernel's own arithmetic for a cell centre (float32; the products are exact): ex0 + (ix + 0.5) * cell
            cell, ox, oy = np.float32(m.grid_length), np.float32(m.origin[0]), np.float32(m.origin[1])
            ex0 = (np.float32(m.x0c) - np.float32(0.5) * cell) - ox
            ey0 = (np.float32(m.y0c) - np.float32(0.5) * cell) - oy
            idx = np.minimum(np.arange(per_map * N), m.gx * m.gy - 1)    # slots past the last cell repeat it
            ix, iy = idx // m.gy, idx % m.gy
            px[k] = ex0 + (ix.astype(np.float32) + np.float32(0.5)) * cell
            py[k] = ey0 + (iy.astype(np.float32) + np.float32(0.5)) * cell
            code[k] = (ix << 8) | iy
        st, dev = probe.state, self.device
        with self._on_stream():
            self._fill_probe(probe, st, dev, px, py, code, n_maps, per_map, N, R, S)
        ptr = lambda key: C.c_void_p(self._rtab[key].data_ptr()) if key in self._rtab else None
        with torch.cuda.device(self.device):
            K.check(self.L.aac_set_radar_table(self.h, ptr("radar"), ptr("radar_min"), ptr("radar_hit"), ptr("min_bits")), "aac_set_radar_table")

    def _fill_probe(self, probe, st, dev, px, py, code, n_maps, per_map, N, R, S):
        st["px"].copy_(torch.from_numpy(px.reshape(-1, N)).to(dev))
        st["py"].copy_(torch.from_numpy(py.reshape(-1, N)).to(dev))
        cells = torch.from_numpy(code.reshape(-1, N).astype(np.int16)).to(dev)
        st["ref_cells"].zero_()
        st["ref_cells"][:, :, 0] = cells
        st["ref_cells"][:, :, 1] = cells
        st["ref_w"].fill_(2)
        if "map_id" in st:
            st["map_id"].copy_(torch.arange(n_maps, device=dev, dtype=torch.int32).repeat_interleave(per_map))
        probe.observe()
        grab = lambda key: probe.out[key].reshape(n_maps, per_map * N, R)[:, :S].contiguous().clone()
        tab = grab("radar")
        self._rtab = {"radar": tab, "min_bits": tab.view(torch.int32).min(dim=2).values.contiguous()}   # ranges are >= 0 (nan sorts last): bit patterns order like the values
        if "radar_min" in probe.out:
            self._rtab["radar_min"], self._rtab["radar_hit"] = grab("radar_min"), grab("radar_hit")
        torch.cuda.synchronize(dev)
        probe.close()

    # ------------------------------------------------------------------ the env surface
    def reset(self, mask=None):
        """reset_world for the masked envs (all when mask is None) from the scenario bank."""
        m = None
        if mask is not None:
            with self._on_stream():
                m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            if self.stream is not None:
                m.record_stream(self.stream)   # the reset kernel reads it after this call returns
        with torch.cuda.device(self.device):
            K.check(self.L.aac_reset(self.h, C.c_void_p(m.data_ptr()) if m is not None else None, C.byref(self._out_c),
                                     self._stream_ptr()), "aac_reset")
        return self.obs()

    def observe(self):
        with torch.cuda.device(self.device):
            K.check(self.L.aac_observe(self.h, C.byref(self._out_c), self._stream_ptr()), "aac_observe")
        return self.obs()

    def step(self, actions: torch.Tensor, autoreset=False, fused=False):
        """actions [E, N, 2] float32 on the device, in [-1, 1] -> (obs, reward, done, info).  autoreset: the envs that
        terminate are re-initialised and their observation rows carry the reset observation (aac_step_autoreset: one fused
        launch, a step launch + a reset launch, or one phased launch - the library's rule or `autoreset_launches`;
        fused=True: the single-launch variant aac_step_fused; bit-identical results either way)."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous() \
                or tuple(actions.shape) != (self.E, self.N, 2):
            raise ValueError("actions must be a contiguous float32 [E, N, 2] tensor on %s" % self.device)
        with torch.cuda.device(self.device):
            fn = (self.L.aac_step_fused if fused else self.L.aac_step_autoreset) if autoreset else self.L.aac_step
            K.check(fn(self.h, C.c_void_p(actions.data_ptr()), C.byref(self._out_c), self._stream_ptr()), "aac_step")
        o = self.out
        info = {k: o[k] for k in ("check_goal", "bbc", "terminated", "tcpa_min", "tcpa_pair", "nbr_order", "radar_min",
                                  "radar_hit", "parts", "branch", "cloud_contact") if k in o}
        return self.obs(), o["reward"], o["done"], info

    def autoreset(self):
        with torch.cuda.device(self.device):
            K.check(self.L.aac_autoreset(self.h, C.byref(self._out_c), self._stream_ptr()), "aac_autoreset")
        return self.obs()

    def host_buffers(self, fields=("norm_own", "norm_nbr", "radar", "norm_nbr6", "reward", "done", "check_goal", "bbc",
                                   "terminated")):
        """Pinned host copies of the step outputs for `step_host`."""
        full = self.alloc_out(None, pin=True)
        self._host_out = {k: v for k, v in full.items() if k in fields}
        self._host_c = self._out_struct(self._host_out)
        return self._host_out

    def step_host(self, actions_host: torch.Tensor, autoreset=True):
        """One step through host memory: H2D of the actions, step (+ auto-reset), D2H of the outputs
        selected by `host_buffers()`; returns after the stream has drained."""
        if self._host_out is None:
            self.host_buffers()
        assert actions_host.dtype == torch.float32 and actions_host.is_contiguous() and not actions_host.is_cuda
        with torch.cuda.device(self.device):
            K.check(self.L.aac_step_host(self.h, C.c_void_p(actions_host.data_ptr()), C.byref(self._out_c), C.byref(self._host_c),
                                         int(autoreset), self._stream_ptr()), "aac_step_host")
        return self._host_out

    def obs(self):
        o = self.out
        return {k: o[k] for k in ("norm_own", "norm_nbr", "radar", "norm_nbr6", "raw_own", "raw_nbr", "raw_nbr6") if k in o}

    def read_stats(self, reset=False):
        buf = (C.c_double * K.N_STATS)()
        with torch.cuda.device(self.device):
            K.check(self.L.aac_read_stats(self.h, buf, int(reset), self._stream_ptr()), "aac_read_stats")
        return np.array(buf[:], dtype=np.float64)

    @property
    def launch_count(self):
        return int(self.L.aac_launch_count(self.h))

    def state_dict(self):
        """Everything needed to resume: the per-drone records and episode counters (host copies).  The reference
        checkpoints actor weights only (ATT/maddpg_agent:131-139); env state is plain tensors here, so it is cheap."""
        return {k: v.detach().cpu().clone() for k, v in self.state.items()}

    def load_state_dict(self, sd):
        with self._on_stream():
            for k, v in sd.items():
                self.state[k].copy_(v.to(self.device))
        return self.observe()

    def close(self):
        if getattr(self, "h", None):
            self.L.aac_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ explicit episode install / readback
    def set_episode(self, e, starts, lines, headings, map_id=0):
        """Install reset data for env `e` exactly as reset_world leaves it (ATT:301-372): start positions
        (global metres), reference lines (lists of cell-centre vertices) and headings."""
        with self._on_stream():
            self._set_episode(e, starts, lines, headings, map_id)

    def _set_episode(self, e, starts, lines, headings, map_id):
        st, g = self.state, self.maps[map_id]
        ox, oy = g.origin
        N, W = self.N, self.cfg.w_max
        cells = np.zeros((N, W), dtype=np.uint16)
        ws = np.zeros(N, dtype=np.uint8)
        for i in range(N):
            w = len(lines[i])
            if w > W:
                raise ValueError("reference line with %d vertices exceeds w_max=%d" % (w, W))
            for k in range(w):
                ix, iy = g.cell_of(lines[i][k][0], lines[i][k][1])
                cx, cy = g.cell_centre(ix, iy)
                if abs(cx - lines[i][k][0]) > 1e-9 or abs(cy - lines[i][k][1]) > 1e-9:
                    raise ValueError("reference-line vertices must be cell centres")
                cells[i, k] = ix * 256 + iy
            ws[i] = w
        dev = self.device
        pos = np.asarray(starts, dtype=np.float64)
        st["px"][e] = torch.tensor(pos[:, 0] - ox, dtype=torch.float32, device=dev)
        st["py"][e] = torch.tensor(pos[:, 1] - oy, dtype=torch.float32, device=dev)
        st["vx"][e] = 0.0
        st["vy"][e] = 0.0
        st["heading"][e] = torch.tensor(np.asarray(headings, dtype=np.float32), device=dev)
        st["meta"][e] = -65536
        st["ref_cells"][e] = torch.tensor(cells.view(np.int16), device=dev)
        st["ref_w"][e] = torch.tensor(ws, device=dev)
        st["wall_count"][e] = 0
        st["ep_step"][e] = 0
        st["ep_return"][e] = 0.0
        if self.cfg.variant == "mm":
            st["map_id"][e] = int(map_id)
            st["wp_mask"][e] = torch.tensor(((1 << ws.astype(np.int64)) - 2).astype(np.int32), device=dev)

    def _env_origins(self):
        if self.cfg.variant == "mm":
            return self.origins[self.state["map_id"].cpu().numpy()]
        return np.broadcast_to(self.origins[0], (self.E, 2))

    def agent_state(self):
        """Host copy of the per-drone records in the reference's terms (global metres)."""
        st = self.state
        org = self._env_origins()
        meta = st["meta"].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        pos = np.stack([st["px"].cpu().numpy().astype(np.float64) + org[:, None, 0],
                        st["py"].cpu().numpy().astype(np.float64) + org[:, None, 1]], -1)
        vel = np.stack([st["vx"].cpu().numpy(), st["vy"].cpu().numpy()], -1).astype(np.float64)
        extra = {}
        if self.cfg.variant == "mm":
            extra = {"wp_mask": st["wp_mask"].cpu().numpy().astype(np.int64) & 0xFFFFFFFF, "map_id": st["map_id"].cpu().numpy()}
        return {**extra, "pos": pos, "vel": vel, "heading": st["heading"].cpu().numpy().astype(np.float64),
                "reach": ((meta >> 8) & 1).astype(np.int32), "wp_cur": (meta & 0xFF).astype(np.int32),
                "vflags": ((meta >> 9) & 7).astype(np.int32), "wall_cnt": st["wall_count"].cpu().numpy(),
                "prev_nn": np.stack([(meta >> 16) & 0xFF, (meta >> 24) & 0xFF], -1).astype(np.int32),
                "ref_w": st["ref_w"].cpu().numpy().astype(np.int32)}

    def load_agent_state(self, pos, vel, heading=None, reach=None, wp_cur=None, prev_nn=None):
        """Overwrite the kinematic state (global metres) -- used by parity tests to keep the float32
        env on the float64 trajectory."""
        with self._on_stream():
            self._load_agent_state(pos, vel, heading, reach, wp_cur, prev_nn)

    def _load_agent_state(self, pos, vel, heading, reach, wp_cur, prev_nn):
        st, dev = self.state, self.device
        org = self._env_origins()
        pos = np.asarray(pos, dtype=np.float64)
        vel = np.asarray(vel, dtype=np.float64)
        st["px"].copy_(torch.tensor(pos[..., 0] - org[:, None, 0], dtype=torch.float32, device=dev))
        st["py"].copy_(torch.tensor(pos[..., 1] - org[:, None, 1], dtype=torch.float32, device=dev))
        st["vx"].copy_(torch.tensor(vel[..., 0], dtype=torch.float32, device=dev))
        st["vy"].copy_(torch.tensor(vel[..., 1], dtype=torch.float32, device=dev))
        if heading is not None:
            st["heading"].copy_(torch.tensor(np.asarray(heading), dtype=torch.float32, device=dev))
        if reach is not None or wp_cur is not None or prev_nn is not None:
            meta = st["meta"].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            if reach is not None:
                meta = (meta & ~(1 << 8)) | (np.asarray(reach, dtype=np.int64) << 8)
            if wp_cur is not None:
                meta = (meta & ~0xFF) | np.asarray(wp_cur, dtype=np.int64)
            if prev_nn is not None:
                pn = np.asarray(prev_nn, dtype=np.int64) & 0xFF
                meta = (meta & 0xFFFF) | (pn[..., 0] << 16) | (pn[..., 1] << 24)
            st["meta"].copy_(torch.tensor(meta.astype(np.uint32).view(np.int32), device=dev))
