"""Parity harness: the CUDA env (through the C ABI) against the float64 oracle.

Two drivers:
  * `GpuGoldenAdapter` lets tests/replay.py replay the committed reference rollouts (tests/golden/*.npz)
    through the CUDA env with teacher forcing (float32 state is re-seeded from the float64 rollout
    after every compared step);
  * `lockstep()` steps a batched CUDA env and the oracle side by side on identical inputs: after every
    step the oracle's state is overwritten with the CUDA env's (float32 values are exact in float64),
    so every comparison isolates one step of arithmetic.

Tolerances (north_star: 1e-4 relative for ranges, tdCPA, rewards and states; flags bit-exact except
counted boundary-epsilon ties):
  * real-valued outputs: |got - want| <= RTOL * |want| + atol, RTOL = 1e-4, atol a per-quantity floor that
    covers float32 resolution of the inputs (positions are float32 in a +-150 m local frame: 1.5e-5 m);
  * a step of an env is a TIE and its flag / reward comparisons are skipped (and counted) when the
    oracle's own `margin` (smallest |quantity - threshold| over all predicates of that drone) is below
    TIE_EPS, or two neighbour distances of a drone differ by less than TIE_EPS (V2 sort order);
  * a radar mismatch is a tie when the CUDA value lies inside the range the oracle itself spans for the
    drone displaced by +-RADAR_EPS (5e-5 m, float32 position resolution): a grazing ray or an edge nearly
    parallel to the ray; anything else is a failure.
"""
import numpy as np
import torch

from multi_agent_aac_b200 import _capi as K
from multi_agent_aac_b200.env import BatchedDroneEnv, preset
from multi_agent_aac_b200.maps import multimap_set, synthetic_map
from multi_agent_aac_b200.reset import MultiMapBank, ScenarioBank
from oracle.oracle import OracleEnv, RADAR_LAST_HIT, RADAR_MIN

RTOL = 1e-4
TIE_EPS = 2e-4
RADAR_EPS = 5e-5
ALL_OUT = K.OUT_RAW | K.OUT_NBR6 | K.OUT_TCPA_PAIR | K.OUT_RADAR_AUX | K.OUT_PARTS
ATOL = {"norm_own": 2e-6, "norm_nbr": 2e-6, "norm_nbr6": 2e-6, "raw_own": 1e-4, "raw_nbr": 1e-4, "raw_nbr6": 1e-4, "radar": 2e-4,
        "radar_min": 2e-4, "reward": 2e-4, "parts": 5e-4, "pos": 5e-5, "vel": 5e-6, "heading": 5e-6}


def np_out(env):
    return {k: v.detach().cpu().numpy() for k, v in env.out.items()}


class GpuGoldenAdapter:
    """OracleEnv surface (set_episode / observe / step / .state) over a 1-env BatchedDroneEnv."""

    def __init__(self, variant, gmap, n_agents, n_rays, device="cuda:0", eval_by_step=False):
        cfg = preset("att" if variant == "att" else "tdcpa_v2", n_envs=1, n_agents=n_agents, n_rays=n_rays, w_max=32, out_flags=ALL_OUT,
                     eval_by_step=eval_by_step)
        self.env = BatchedDroneEnv(cfg, gmap, device=device)
        self.variant = variant
        self.state = {}

    def _collect(self):
        o = np_out(self.env)
        s = self.env.agent_state()
        self.state = {"pos": s["pos"], "vel": s["vel"], "heading": s["heading"], "reach": s["reach"], "wp_cur": s["wp_cur"]}
        return o

    def set_episode(self, e, starts, lines, headings):
        self.env.set_episode(e, starts, lines, headings)

    def observe(self):
        self.env.observe()
        return self._collect()

    def step(self, actions):
        a = torch.tensor(np.asarray(actions, dtype=np.float32), device=self.env.device).contiguous()
        self.env.step(a)
        return self._collect()

    def resync(self, env, d, t):
        """Teacher forcing from the golden float64 rollout."""
        self.env.load_agent_state(d["pos"][t][None], d["vel"][t][None],
                                  heading=d["heading"][t][None] if self.variant == "v2" else None)


class GpuGoldenAdapterMM:
    """OracleEnv surface for tests/replay.replay_mm over a 1-env multipleMap BatchedDroneEnv."""

    def __init__(self, maps, n_agents, n_rays, device="cuda:0"):
        cfg = preset("multimap", n_envs=1, n_agents=n_agents, n_rays=n_rays, w_max=32, out_flags=K.OUT_RAW | K.OUT_RADAR_AUX | K.OUT_PARTS)
        self.env = BatchedDroneEnv(cfg, maps, device=device)
        self.state = {}

    def _collect(self):
        o = np_out(self.env)
        s = self.env.agent_state()
        self.state = {"pos": s["pos"], "vel": s["vel"], "reach": s["reach"], "wp_mask": s["wp_mask"], "wall_cnt": s["wall_cnt"]}
        return o

    def set_episode(self, e, starts, lines, headings, map_id=0):
        self.env.set_episode(e, starts, lines, headings, map_id=map_id)

    def observe(self):
        self.env.observe()
        return self._collect()

    def step(self, actions):
        a = torch.tensor(np.asarray(actions, dtype=np.float32), device=self.env.device).contiguous()
        self.env.step(a)
        return self._collect()

    def resync(self, env, d, t):
        self.env.load_agent_state(d["pos"][t][None], d["vel"][t][None])


def cells_to_lines(gmap, ref_cells, ref_w):
    """uint16 cell codes [.., W] -> float64 vertex coordinates [.., W, 2]."""
    c = ref_cells.astype(np.int64) & 0xFFFF
    x = gmap.x0c + (c >> 8) * gmap.grid_length
    y = gmap.y0c + (c & 255) * gmap.grid_length
    return np.stack([x, y], -1).astype(np.float64)


def sync_oracle(orc, env, envs=None, full=False):
    """Copy the CUDA env's state into the oracle (all envs, or the listed ones)."""
    s = env.agent_state()
    sel = slice(None) if envs is None else envs
    st = orc.state
    st["pos"][sel] = s["pos"][sel]
    st["vel"][sel] = s["vel"][sel]
    st["heading"][sel] = s["heading"][sel]
    st["reach"][sel] = s["reach"][sel]
    st["wp_cur"][sel] = s["wp_cur"][sel]
    st["wall_cnt"][sel] = s["wall_cnt"][sel]
    st["vflags"][sel] = s["vflags"][sel]
    pn = s["prev_nn"].copy()
    pn[pn == 255] = -1
    st["prev_nn"][sel] = pn[sel]
    st["ep_step"][sel] = env.state["ep_step"].cpu().numpy()[sel]     # sensor configurations: the clouds' clock
    if "wp_mask" in s:
        st["wp_mask"][sel] = s["wp_mask"][sel].astype(np.int32)
        orc.env_map[sel] = s["map_id"][sel]
    if full:
        cells = env.state["ref_cells"].cpu().numpy().view(np.uint16)
        if "map_id" in s:   # vertices are cell centres of the env's own map
            lines = np.stack([cells_to_lines(env.maps[int(m)], cells[e], None) for e, m in enumerate(s["map_id"])])
        else:
            lines = cells_to_lines(env.gmap, cells, None)
        st["ref_line"][sel] = lines[sel][..., :orc.w_max, :]
        st["ref_w"][sel] = s["ref_w"][sel]
    return s


class Tally:
    def __init__(self):
        self.n = {}
        self.ties = {}
        self.worst = {}
        self.fail = []

    def count(self, key, n=1):
        self.n[key] = self.n.get(key, 0) + int(n)

    def tie(self, key, n=1):
        self.ties[key] = self.ties.get(key, 0) + int(n)

    def close(self, key, got, want, where, atol, rtol=RTOL, mask=None):
        got = np.asarray(got, dtype=np.float64)
        want = np.asarray(want, dtype=np.float64).reshape(got.shape)
        both_nan = np.isnan(got) & np.isnan(want)
        err = np.abs(got - want) - rtol * np.abs(want) - atol
        err = np.where(both_nan | ((got == want)), -1.0, err)
        err = np.where(np.isnan(err), np.inf, err)
        if mask is not None:
            err = np.where(mask, err, -1.0)
        self.count(key, err.size if mask is None else int(np.sum(mask)))
        w = float(np.max(np.abs(got - want)[np.isfinite(got - want)])) if np.isfinite(got - want).any() else 0.0
        self.worst[key] = max(self.worst.get(key, 0.0), w)
        bad = err > 0
        if bad.any():
            idx = np.unravel_index(np.argmax(err), err.shape)
            if len(self.fail) < 50:
                self.fail.append("%s %s idx=%s got=%r want=%r (%d bad)" % (key, where, idx, got[idx], want[idx], int(bad.sum())))
        return bad

    def equal(self, key, got, want, where, mask=None):
        got = np.asarray(got).astype(np.int64)
        want = np.asarray(want).astype(np.int64).reshape(got.shape)
        bad = got != want
        if mask is not None:
            bad &= mask
        self.count(key, got.size if mask is None else int(np.sum(mask)))
        if bad.any() and len(self.fail) < 50:
            idx = np.unravel_index(np.argmax(bad), bad.shape)
            self.fail.append("%s %s idx=%s got=%r want=%r (%d bad)" % (key, where, idx, got[idx], want[idx], int(bad.sum())))
        return bad

    def summary(self):
        return {"compared": self.n, "ties": self.ties, "worst_abs_err": {k: float("%.3g" % v) for k, v in self.worst.items()},
                "failures": len(self.fail)}


def _bcast(mask_env, shape):
    m = np.asarray(mask_env)
    return np.broadcast_to(m.reshape(m.shape + (1,) * (len(shape) - m.ndim)), shape)


def compare_obs(T, variant, g, o, where, env_ok, s_gpu, orc, rows=None):
    """Observation blocks + radar.  env_ok[e] False => order-dependent blocks of env e are skipped."""
    E = g["norm_own"].shape[0]
    rows = np.ones(E, dtype=bool) if rows is None else rows
    ok = env_ok & rows
    keys = ["norm_own", "raw_own"] + (["norm_nbr6", "raw_nbr6"] if variant != "mm" else []) + (["norm_nbr", "raw_nbr"] if variant == "v2" else [])
    for k in keys:
        if k in g:
            T.close(k, g[k], o[k], where, ATOL[k], mask=_bcast(ok, g[k].shape))
    # radar: every mismatch must be explained by a grazing ray
    for k in ("radar", "radar_min"):
        bad = T.close(k + "(pre-tie)", g[k], o[k], where, ATOL[k], mask=_bcast(rows, g[k].shape))
        T.fail = [f for f in T.fail if not f.startswith(k + "(pre-tie)")]
        radar_tie_env = np.zeros(E, dtype=bool)
        for e, i in {(int(e), int(i)) for e, i, _ in zip(*np.nonzero(bad))}:
            # the CUDA value must lie inside the range the oracle itself spans when the drone is displaced
            # by +-RADAR_EPS (float32 position resolution): grazing rays and near-parallel edges
            gv = g[k][e, i].astype(np.float64)
            nominal = o[k][e, i]
            lo, hi, any_nan = nominal.copy(), nominal.copy(), np.isnan(nominal)
            for dx, dy in ((RADAR_EPS, 0), (-RADAR_EPS, 0), (0, RADAR_EPS), (0, -RADAR_EPS), (RADAR_EPS, RADAR_EPS),
                           (-RADAR_EPS, -RADAR_EPS), (RADAR_EPS, -RADAR_EPS), (-RADAR_EPS, RADAR_EPS)):
                pos = orc.state["pos"][e].copy()
                pos[i] += (dx, dy)
                out, omin, _ = orc.radar_probe(pos, i, map_id=int(orc.env_map[e]), k_cloud=int(orc.state["ep_step"][e]))
                ref = out if k == "radar" else omin
                any_nan |= np.isnan(ref)
                lo, hi = np.fmin(lo, ref), np.fmax(hi, ref)
            tol = RTOL * np.abs(hi) + ATOL[k]
            explained = ((gv >= lo - tol) & (gv <= hi + tol)) | (np.isnan(gv) & any_nan)
            still = bad[e, i] & ~explained
            T.tie(k, int((bad[e, i] & explained).sum()))
            radar_tie_env[e] = True
            if still.any() and len(T.fail) < 50:
                r = int(np.argmax(still))
                T.fail.append("%s %s env=%d drone=%d ray=%d got=%r want=%r pos=%r" % (k, where, e, i, r, g[k][e, i, r], o[k][e, i, r],
                                                                                      orc.state["pos"][e, i].tolist()))
        T.count(k, int(rows.sum()) * g[k].shape[1] * g[k].shape[2])
    # Hit ids (bit-exact except counted boundary-epsilon ties).  A mismatch whose range agrees is a tie only if the oracle
    # itself says so: either the candidate the CUDA env names is, for the oracle, at the same distance as its own hit (a ray
    # through a corner or an edge shared by two cells, a cell face lying on a boundary line), or the oracle names that very
    # candidate when the drone is displaced by +-RADAR_EPS (a grazing candidate: hit on one side, missed on the other).
    # Anything else fails.
    hit_ok = (np.abs(g["radar"].astype(np.float64) - o["radar"]) <= 1e-3) & _bcast(rows & ~radar_tie_env, g["radar"].shape)
    badh = (g["radar_hit"].astype(np.int64) != o["radar_hit"]) & hit_ok
    T.count("radar_hit", int(hit_ok.sum()))
    for e, i, r in zip(*np.nonzero(badh)):
        e, i, r = int(e), int(i), int(r)
        gh, oh = int(g["radar_hit"][e, i, r]), int(o["radar_hit"][e, i, r])
        mid = int(orc.env_map[e])
        pos = orc.state["pos"][e]
        want = float(o["radar"][e, i, r])
        tie = False
        if variant != "att" and not orc.cfg.radar_targets:
            d_g = float(orc.cfg.ray_len) if gh < 0 else orc.radar_candidate(pos[i], r, gh, mid)
            d_o = float(orc.cfg.ray_len) if oh < 0 else orc.radar_candidate(pos[i], r, oh, mid)
            if d_g is not None and d_o is not None and abs(d_g - d_o) <= RTOL * abs(d_o) + ATOL["radar"] and abs(d_g - want) <= RTOL * abs(want) + ATOL["radar"]:
                tie = True
        if not tie:
            for dx, dy in ((RADAR_EPS, 0), (-RADAR_EPS, 0), (0, RADAR_EPS), (0, -RADAR_EPS), (RADAR_EPS, RADAR_EPS),
                           (-RADAR_EPS, -RADAR_EPS), (RADAR_EPS, -RADAR_EPS), (-RADAR_EPS, RADAR_EPS)):
                pp = pos.copy()
                pp[i] += (dx, dy)
                if int(orc.radar_probe(pp, i, map_id=mid, k_cloud=int(orc.state["ep_step"][e]))[2][r]) == gh:
                    tie = True
                    break
        if tie:
            T.tie("radar_hit", 1)
        elif len(T.fail) < 50:
            T.fail.append("radar_hit %s env=%d drone=%d ray=%d got=%d want=%d range=%r pos=%r" % (where, e, i, r, gh, oh, want, pos[i].tolist()))
    return radar_tie_env


def lockstep(variant, n_envs, n_agents, n_rays, steps, seed=0, radar_mode=None, n_scen=128, cluster=None, map_seed=0,
             device="cuda:0", autoreset=True, tile_envs=0, block_threads=0, action_scale=1.0, eval_by_step=False, sensors=None):
    """Returns a Tally.  `cluster` = radius (m): after every reset drones 1.. are moved next to drone 0 so
    that drone-radar / near-drone / collision branches fire.  `sensors` = dict(radar_targets, n_nbr_obs, clouds, prot, bound):
    the later fork's sensor classes on the tdCPA_forV2 variant."""
    E, N, R, M = n_envs, n_agents, n_rays, n_agents - 1
    if radar_mode is None:
        radar_mode = RADAR_LAST_HIT if variant == "v2" else RADAR_MIN
    if variant == "mm":
        gmap = multimap_set(seed=map_seed)
        cfg = preset("multimap", n_envs=E, n_agents=N, n_rays=R, w_max=32, out_flags=K.OUT_RAW | K.OUT_RADAR_AUX | K.OUT_PARTS,
                     seed=seed, tile_envs=tile_envs, block_threads=block_threads)
        bank = MultiMapBank(gmap, N, n_scen, w_max=32, seed=seed)
        M = 0   # no neighbour terms on the multipleMap path
    else:
        sk = dict(sensors or {})
        gmap = synthetic_map(bound=sk.pop("bound", None), seed=map_seed)
        out_flags = ALL_OUT if not sk.get("n_nbr_obs") else (K.OUT_RAW | K.OUT_RADAR_AUX | K.OUT_PARTS)   # per-pair outputs cover every neighbour
        cfg = preset("att" if variant == "att" else "tdcpa_v2", n_envs=E, n_agents=N, n_rays=R, w_max=32, out_flags=out_flags,
                     radar_mode=radar_mode, seed=seed, tile_envs=tile_envs, block_threads=block_threads, eval_by_step=eval_by_step, **sk)
        bank = ScenarioBank(gmap, N, n_scen, w_max=32, seed=seed)
    env = BatchedDroneEnv(cfg, gmap, device=device)
    env.set_bank(bank)
    okw = {} if not sensors else dict(radar_targets=cfg.radar_targets, n_nbr_obs=cfg.n_nbr_obs, clouds=cfg.clouds, prot=cfg.prot)
    orc = OracleEnv(variant, gmap, E, N, R, w_max=32, radar_mode=radar_mode, eval_by_step=eval_by_step, **okw)
    pairs_out = "nbr_order" in env.out    # per-pair optional outputs are off with a nearest-N neighbour block
    rng = np.random.default_rng(seed + 1)
    T = Tally()

    def scatter_cluster(envs):
        if cluster is None or N < 2:
            return
        st = env.state
        px, py = st["px"].cpu().numpy(), st["py"].cpu().numpy()
        for e in envs:
            for i in range(1, N):
                px[e, i] = px[e, 0] + rng.uniform(-cluster, cluster)
                py[e, i] = py[e, 0] + rng.uniform(-cluster, cluster)
        st["px"].copy_(torch.tensor(px, device=env.device))
        st["py"].copy_(torch.tensor(py, device=env.device))

    def check_reset(envs, where):
        scatter_cluster(envs)
        env.observe()
        g = np_out(env)
        sync_oracle(orc, env, full=True)
        orc.state["prev_nn"][envs] = -1
        o = {k: v.copy() for k, v in orc.observe().items()}
        rows = np.zeros(E, dtype=bool)
        rows[envs] = True
        env_ok = sort_ok(orc, o)
        compare_obs(T, variant, g, o, where, env_ok, None, orc, rows=rows)
        if "cloud_contact" in g:
            T.equal("cloud_contact", g["cloud_contact"], o["cloud_contact"], where, mask=_bcast(rows, g["cloud_contact"].shape))
        if M > 0 and pairs_out:
            T.equal("nbr_order", g["nbr_order"], o["nbr_order"], where, mask=_bcast(env_ok & rows, g["nbr_order"].shape))

    def sort_ok(orc_, o):
        if variant != "v2" or M < 2:
            return np.ones(E, dtype=bool)
        pos = orc_.state["pos"]
        d = np.linalg.norm(pos[:, :, None, :] - pos[:, None, :, :], axis=-1)          # [E,N,N]
        order = o["nbr_order"].astype(np.int64)                                      # [E,N,M]
        ds = np.take_along_axis(d, order, axis=2)
        gap = np.min(np.diff(ds, axis=2), axis=(1, 2))
        bad = (gap < TIE_EPS) & (gap > 0.0)   # exact ties (lattice starts) resolve by index on both sides
        T.tie("sort_order", int(bad.sum()))
        return ~bad

    env.reset()
    check_reset(np.arange(E), "reset")
    for t in range(steps):
        act = (rng.uniform(-1.0, 1.0, size=(E, N, 2)) * action_scale).astype(np.float32)
        where = "t%d" % t
        env.step(torch.tensor(act, device=env.device))
        g = np_out(env)
        s_gpu = env.agent_state()
        o = {k: v.copy() for k, v in orc.step(act.astype(np.float64)).items()}
        so = {k: v.copy() for k, v in orc.state.items()}
        env_ok = sort_ok(orc, o)
        # kinematics
        T.close("pos", s_gpu["pos"], so["pos"], where, ATOL["pos"])
        T.close("vel", s_gpu["vel"], so["vel"], where, ATOL["vel"])
        if variant == "v2":
            # heading = atan2 of a float32 displacement: compare as an angle difference, skip stopped drones
            dh = np.abs(np.angle(np.exp(1j * (s_gpu["heading"] - so["heading"]))))
            moving = np.linalg.norm(so["vel"], axis=-1) > 1e-3
            T.close("heading", np.where(moving, dh, 0.0), np.zeros_like(dh), where, 2e-5)
        radar_tie_env = compare_obs(T, variant, g, o, where, env_ok, s_gpu, orc)
        if "cloud_contact" in g:   # overlap of two 64-gons: compared away from the touching distance
            cl = np.array([orc.cloud_positions(int(k)) for k in orc.state["ep_step"]])                      # [E, C, 2]
            gap = np.abs(np.linalg.norm(so["pos"][:, :, None, :] - cl[:, None, :, :], axis=-1) - (cfg.prot + np.array([c[4] for c in cfg.clouds])))
            T.equal("cloud_contact", g["cloud_contact"], o["cloud_contact"], where, mask=(gap > 0.02).all(axis=2))
        if M > 0 and pairs_out:
            T.equal("nbr_order", g["nbr_order"], o["nbr_order"], where, mask=_bcast(env_ok, g["nbr_order"].shape))
            # tdCPA: t = (r.w)/|w|^2 and d = |-r + w t| are ill-conditioned for nearly equal velocities.  The
            # float32 state carries dv = 2e-6 m/s of rounding per velocity and dr = 2e-5 m per position, so the
            # first-order floor is |dt| <= (|r| dv + |w| dr) / |w|^2 + 2 |t| dv / |w|, |dd| <= |w| |dt| + |t| dv + dr
        if M > 0 and pairs_out:
            tp_g, tp_o = g["tcpa_pair"].astype(np.float64), o["tcpa"]
            order = o["nbr_order"].astype(np.int64)
            vel, pos = so["vel"], so["pos"]
            take = lambda x: np.take_along_axis(x[:, None, :, :].repeat(N, 1), order[..., None].repeat(2, -1), axis=2)
            w = np.linalg.norm(take(vel) - vel[:, :, None, :], axis=-1)
            rr = np.linalg.norm(take(pos) - pos[:, :, None, :], axis=-1)
            dv, dr = 2e-6, 2e-5
            ws = np.maximum(w, 1e-9)
            okp = _bcast(env_ok, tp_g.shape[:3])
            for c, name in ((0, "tcpa"), (1, "d_tcpa")):
                want, got = tp_o[..., c], tp_g[..., c]
                t_abs = np.abs(tp_o[..., 0])
                floor_t = (rr * dv + ws * dr) / ws ** 2 + 2 * t_abs * dv / ws
                floor = floor_t if c == 0 else ws * floor_t + t_abs * dv + dr
                special = (tp_o[..., 0] == -10.0) | (tp_g[..., 0] == -10.0)   # exact zero relative velocity on one side only: tie
                err = np.abs(got - want) - RTOL * np.abs(want) - 1e-4 - 4 * floor
                T.count(name, int(okp.sum()))
                T.worst[name] = max(T.worst.get(name, 0.0), float(np.max(np.where(okp & ~special, np.abs(got - want) / (1 + 4 * floor / 1e-4), 0.0))))
                badp = (err > 0) & okp & ~(special & (tp_o[..., 0] != tp_g[..., 0]))
                if badp.any() and len(T.fail) < 50:
                    idx = np.unravel_index(np.argmax(np.where(badp, err, -1)), err.shape)
                    T.fail.append("%s %s idx=%s got=%r want=%r |w|=%r |r|=%r floor=%r" % (name, where, idx, got[idx], want[idx], w[idx], rr[idx], floor[idx]))
        # flags and rewards: skip envs with a predicate within TIE_EPS of its threshold
        margin_ok = (o["margin"] >= TIE_EPS).all(axis=1)
        T.tie("predicate_margin", int((~margin_ok).sum()))
        flag_ok = margin_ok & env_ok & ~radar_tie_env
        fm = _bcast(flag_ok, (E, N))
        if variant == "mm":   # no coupling between the drones of an env in this variant's reward: mask per drone
            fm = (o["margin"] >= TIE_EPS) & _bcast(env_ok & ~radar_tie_env, (E, N))
        bad_env = np.zeros(E, dtype=bool)
        wp_key = ("wp_mask", s_gpu["wp_mask"], so["wp_mask"]) if variant == "mm" else ("wp_cur", s_gpu["wp_cur"], so["wp_cur"])
        for key, gv, ov in (("done", g["done"], o["done"]), ("check_goal", g["check_goal"], o["check_goal"]),
                            ("branch", g["branch"], o["branch"]), ("reach", s_gpu["reach"], so["reach"]),
                            wp_key, ("wall_cnt", s_gpu["wall_cnt"], so["wall_cnt"])):
            bad_env |= T.equal(key, gv, ov, where, mask=fm).any(axis=1)
        if variant == "v2":
            T.equal("vflags", s_gpu["vflags"], so["vflags"], where, mask=fm)
        T.equal("bbc", g["bbc"], o["bbc"], where, mask=_bcast(flag_ok, (E, 4)))
        T.close("reward", g["reward"], o["reward"], where, ATOL["reward"], mask=fm)
        T.close("parts", g["parts"], o["parts"], where, ATOL["parts"], mask=np.broadcast_to(fm[..., None], g["parts"].shape))
        if M > 0:
            tm_g, tm_o = g["tcpa_min"].astype(np.float64), o["tcpa_min"]
            # conflict counters flip when d_tcpa crosses 2*prot or tcpa crosses 0 / 1: compare away from those
            tp = o["tcpa"]
            near = (np.abs(tp[..., 1] - 5.0) < 5e-3) | (np.abs(tp[..., 3] - 5.0) < 5e-3) | (np.abs(tp[..., 0]) < 1e-3) | \
                   (np.abs(tp[..., 0] - 1.0) < 1e-3) | (np.abs(tp[..., 2]) < 1e-3) | (np.abs(tp[..., 2] - 1.0) < 1e-3)
            cm = fm & ~near.any(axis=2)
            T.equal("conflict_counts", tm_g[..., 3], tm_o[..., 3], where, mask=cm)
        # terminated: step cap | any done | all reached
        term_o = ((env.state["ep_step"].cpu().numpy() > cfg.episode_length).astype(np.int64)
                  | (o["done"].any(axis=1).astype(np.int64) << 1) | (so["reach"].all(axis=1).astype(np.int64) << 2))
        T.equal("terminated", g["terminated"], term_o, where, mask=flag_ok)
        # lock step: the oracle continues from the CUDA env's float32 state
        sync_oracle(orc, env)
        if autoreset:
            term = np.nonzero(g["terminated"])[0]
            if len(term):
                env.autoreset()
                check_reset(term, where + ".reset")
    T.launches = env.launch_count
    env.close()
    return T
