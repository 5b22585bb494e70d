"""Replay a golden rollout (tests/golden/*.npz) through an env implementation and diff every output.

`env` is anything with the OracleEnv surface: set_episode(e, starts, lines, headings), observe(),
step(actions[E,N,2]) -> dict of arrays, and a `.state` dict with pos / vel / reach / wp_cur.
"""
import os

import numpy as np

from multi_agent_aac_b200.maps import GridMap

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    variant = str(d["meta_variant"])
    n, seed, steps, ep_len, rays, mseed = [int(v) for v in d["meta"][:6]]   # a 7th entry = eval_by_step (forV2 "by sorties")
    gmap = GridMap([int(v) for v in d["bound"]], 10, d["occ"].astype(np.uint8))
    return d, variant, n, rays, ep_len, gmap


def obs_parts(variant):
    return ("own", "radar", "nbr6") if variant == "att" else ("own", "nbr", "radar", "nbr6")


class Diff:
    def __init__(self, rtol, atol):
        self.rtol, self.atol = rtol, atol
        self.worst = {}
        self.fail = []

    def close(self, key, got, want, where):
        got = np.asarray(got, dtype=np.float64).reshape(np.shape(want))
        want = np.asarray(want, dtype=np.float64)
        err = np.abs(got - want) - self.rtol * np.abs(want)
        m = float(np.max(err)) if err.size else 0.0
        self.worst[key] = max(self.worst.get(key, -np.inf), m)
        if m > self.atol:
            idx = np.unravel_index(np.argmax(err), err.shape)
            self.fail.append("%s @%s idx=%s got=%r want=%r" % (key, where, idx, got[idx], want[idx]))

    def equal(self, key, got, want, where):
        got = np.asarray(got).reshape(np.shape(want)).astype(np.int64)
        want = np.asarray(want).astype(np.int64)
        if not np.array_equal(got, want):
            self.fail.append("%s @%s got=%s want=%s" % (key, where, got.tolist(), want.tolist()))


def replay(env, d, variant, rtol=1e-9, atol=1e-9, resync=None, max_steps=None):
    """Free-running replay of env slot 0.  `resync(env, d, t)` (optional) is called after every
    compared step to overwrite the env state with the golden state (teacher forcing for fp32)."""
    diff = Diff(rtol, atol)
    parts = obs_parts(variant)
    T = d["actions"].shape[0] if max_steps is None else min(max_steps, d["actions"].shape[0])
    N = d["actions"].shape[1]

    def install(ep):
        w = d["ep_ref_w"][ep]
        lines = [d["ep_ref_line"][ep, i, :w[i]] for i in range(N)]
        env.set_episode(0, d["ep_start"][ep], lines, d["ep_heading"][ep])
        out = env.observe()
        for pi, p in enumerate(parts):
            for kind in ("raw", "norm"):
                key = "%s_%s" % (kind, p) if p != "radar" else "radar"
                diff.close("reset." + key, out[key][0], d["ep_%s_%d" % (kind, pi)][ep], "ep%d" % ep)

    ep = 0
    install(0)
    for t in range(T):
        if int(d["episode_id"][t]) != ep:
            ep = int(d["episode_id"][t])
            install(ep)
        out = env.step(d["actions"][t][None])
        where = "t%d(ep%d,s%d)" % (t, ep, int(d["step_in_ep"][t]))
        for p in parts:
            for kind in ("raw", "norm"):
                key = "%s_%s" % (kind, p) if p != "radar" else "radar"
                diff.close(key, out[key][0], d["%s_%s" % (kind, p)][t], where)
        diff.close("reward", out["reward"][0], d["reward"][t], where)
        diff.equal("done", out["done"][0], d["done"][t], where)
        diff.equal("check_goal", out["check_goal"][0], d["check_goal"][t], where)
        diff.equal("bbc", out["bbc"][0], d["bbc"][t], where)
        diff.close("pos", env.state["pos"][0], d["pos"][t], where)
        diff.close("vel", env.state["vel"][0], d["vel"][t], where)
        diff.equal("reach", env.state["reach"][0], d["reach"][t], where)
        n_wp = d["ep_ref_w"][ep] - 1 - np.asarray(env.state["wp_cur"][0])
        diff.equal("n_wp", n_wp, d["n_wp"][t], where)
        if variant == "v2":
            diff.close("heading", env.state["heading"][0], d["heading"][t], where)
        if resync is not None:
            resync(env, d, t)
    return diff


def oracle_margins_mm(d, n, rays, maps):
    """Float64 oracle replay of a multipleMap rollout, returning margin[t, i] for replay_mm(..., margins=)."""
    from oracle.oracle import OracleEnv
    env = OracleEnv("mm", maps, 1, n, rays)
    T = d["actions"].shape[0]
    out = np.zeros((T, n))
    ep = -1
    for t in range(T):
        if int(d["episode_id"][t]) != ep:
            ep = int(d["episode_id"][t])
            w = d["ep_ref_w"][ep]
            env.set_episode(0, d["ep_start"][ep], [d["ep_ref_line"][ep, i, :w[i]] for i in range(n)], d["ep_heading"][ep],
                            map_id=int(d["ep_map"][ep]))
            env.observe()
        out[t] = env.step(d["actions"][t][None])["margin"][0]
    return out


def load_case_mm(name):
    from multi_agent_aac_b200.maps import multimap_set
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, seed, steps, ep_len, rays, mseed, n_maps = [int(v) for v in d["meta"]]
    return d, n, rays, ep_len, multimap_set(seed=mseed)[:n_maps]


def replay_mm(env, d, rtol=1e-9, atol=1e-9, resync=None, margins=None, tie_eps=2e-4):
    """multipleMap rollouts: a map per episode, observation = own block + radar, waypoint mask in the state.
    `margins[t, i]` (optional, from a float64 oracle replay of the same rollout) = smallest |quantity - threshold|
    over drone i's predicates at step t: steps of a drone closer than `tie_eps` to a threshold are ties and their
    reward / flag comparisons are skipped and counted in `diff.ties`."""
    diff = Diff(rtol, atol)
    diff.ties = 0
    T, N = d["actions"].shape[0], d["actions"].shape[1]
    ep = -1
    for t in range(T):
        if int(d["episode_id"][t]) != ep:
            ep = int(d["episode_id"][t])
            w = d["ep_ref_w"][ep]
            lines = [d["ep_ref_line"][ep, i, :w[i]] for i in range(N)]
            env.set_episode(0, d["ep_start"][ep], lines, d["ep_heading"][ep], map_id=int(d["ep_map"][ep]))
            out = env.observe()
            diff.close("reset.raw_own", out["raw_own"][0], d["ep_raw_own"][ep], "ep%d" % ep)
            diff.close("reset.norm_own", out["norm_own"][0], d["ep_norm_own"][ep], "ep%d" % ep)
            diff.close("reset.radar", out["radar"][0], d["ep_radar"][ep], "ep%d" % ep)
        out = env.step(d["actions"][t][None])
        where = "t%d(ep%d,s%d,map%d)" % (t, ep, int(d["step_in_ep"][t]), int(d["map_id"][t]))
        for k in ("raw_own", "norm_own", "radar"):
            diff.close(k, out[k][0], d[k][t], where)
        diff.close("pos", env.state["pos"][0], d["pos"][t], where)
        diff.close("vel", env.state["vel"][0], d["vel"][t], where)
        tie = margins is not None and bool((margins[t] < tie_eps).any())
        if tie:
            diff.ties += 1
        else:
            diff.close("reward", out["reward"][0], d["reward"][t], where)
            diff.equal("done", out["done"][0], d["done"][t], where)
            diff.equal("check_goal", out["check_goal"][0], d["check_goal"][t], where)
            diff.equal("bbc", np.asarray(out["bbc"][0])[:2], d["bbc"][t], where)
            diff.equal("reach", env.state["reach"][0], d["reach"][t], where)
            diff.equal("wp_mask", env.state["wp_mask"][0], d["wp_mask"][t], where)
            diff.equal("wall", env.state["wall_cnt"][0], d["wall"][t], where)
        if resync is not None:
            resync(env, d, t)
    return diff
