"""Generate the committed golden fixtures by running the UNMODIFIED reference env classes.

Build-container only:  python tests/golden/gen_golden.py
Reads /root/reference (never copied), writes tests/golden/<case>.npz.  Geometry underneath the
reference code is oracle/geos_lite.py (shapely is not installable here).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from multi_agent_aac_b200.maps import synthetic_map  # noqa: E402
from tests.golden import ref_harness as H  # noqa: E402

CASES = {
    # name: (variant, N, seed, steps, episode_length, cluster_radius, min_sep, n_rays, map_seed, policy)
    "att_n3_plain": ("att", 3, 0, 110, 50, None, 0.0, 18, 0, "random"),
    "att_n3_seek": ("att", 3, 10, 150, 50, None, 0.0, 18, 0, "seek"),
    "att_n3_cluster": ("att", 3, 1, 100, 50, 9.0, 0.0, 18, 0, "random"),
    "att_n4_near": ("att", 4, 11, 120, 50, 12.0, 5.5, 18, 0, "random"),
    "att_n5_near_r36": ("att", 5, 2, 80, 50, 14.0, 5.5, 36, 0, "seek"),
    "v2_n3_plain": ("v2", 3, 3, 120, 100, None, 0.0, 18, 0, "random"),
    "v2_n3_seek": ("v2", 3, 12, 200, 100, None, 0.0, 18, 0, "seek"),
    "v2_n4_cluster": ("v2", 4, 4, 80, 100, 10.0, 0.0, 18, 0, "random"),
    "v2_n4_near": ("v2", 4, 13, 120, 100, 12.0, 5.5, 18, 0, "seek"),
    "v2_n6_near_r36": ("v2", 6, 5, 70, 100, 16.0, 5.5, 36, 1, "random"),
    # forV2 evaluation "by sorties" (an 11th field = eval_by_step): crashed / arrived drones stay put, episodes run on
    "v2_n3_evalstep_seek": ("v2", 3, 35, 160, 80, None, 0.0, 18, 0, "seek", 1),
    "v2_n4_evalstep_cluster": ("v2", 4, 41, 100, 40, 7.0, 5.2, 18, 0, "random", 1),
    "v2_n5_evalstep_cluster": ("v2", 5, 31, 90, 30, 9.0, 0.0, 18, 0, "random", 1),
}


MM_CASES = {
    # name: (N, seed, steps, episode_length, n_rays, map_seed, n_maps, policy, cluster_radius)
    "mm_n3_seek": (3, 20, 260, 150, 18, 0, 14, "seek", None),
    "mm_n3_random": (3, 21, 160, 150, 18, 0, 14, "random", None),
    "mm_n4_seek_r36": (4, 22, 160, 150, 36, 1, 6, "seek", None),
}


def pack(r):
    out = {k: v for k, v in r.items() if k != "episodes"}
    eps = r["episodes"]
    n = eps[0]["start"].shape[0]
    wmax = max(len(l) for e in eps for l in e["ref_lines"])
    lines = np.zeros((len(eps), n, wmax, 2))
    w = np.zeros((len(eps), n), dtype=np.int32)
    for ei, e in enumerate(eps):
        for i, l in enumerate(e["ref_lines"]):
            lines[ei, i, :len(l)] = l
            w[ei, i] = len(l)
    out["ep_start"] = np.stack([e["start"] for e in eps])
    out["ep_heading"] = np.stack([e["heading"] for e in eps])
    out["ep_ref_line"] = lines
    out["ep_ref_w"] = w
    for part in range(len(eps[0]["raw"])):
        out["ep_raw_%d" % part] = np.stack([e["raw"][part] for e in eps])
        out["ep_norm_%d" % part] = np.stack([e["norm"][part] for e in eps])
    return out


def main_mm(names=None):
    from multi_agent_aac_b200.maps import multimap_set
    for name, (n, seed, steps, ep_len, rays, mseed, n_maps, policy, cl) in MM_CASES.items():
        if names and name not in names:
            continue
        t = time.time()
        maps = multimap_set(seed=mseed)[:n_maps]
        r = H.rollout_mm(maps, n, seed, steps, ep_len, n_rays=rays, policy=policy, cluster_radius=cl)
        eps = r.pop("episodes")
        d = dict(r)
        wmax = max(len(l) for e in eps for l in e["ref_lines"])
        lines = np.zeros((len(eps), n, wmax, 2))
        w = np.zeros((len(eps), n), dtype=np.int32)
        for ei, e in enumerate(eps):
            for i, l in enumerate(e["ref_lines"]):
                lines[ei, i, :len(l)] = l
                w[ei, i] = len(l)
        d["ep_start"] = np.stack([e["start"] for e in eps])
        d["ep_heading"] = np.stack([e["heading"] for e in eps])
        d["ep_ref_line"], d["ep_ref_w"] = lines, w
        d["ep_map"] = np.array([e["map"] for e in eps])
        d["ep_raw_own"] = np.stack([e["raw"][0] for e in eps])
        d["ep_norm_own"] = np.stack([e["norm"][0] for e in eps])
        d["ep_radar"] = np.stack([e["raw"][1] for e in eps])
        d["meta_variant"] = np.array("mm")
        d["meta"] = np.array([n, seed, steps, ep_len, rays, mseed, n_maps])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print("%s: %d steps, %d episodes, done=%d goal=%d maps=%s (%.1fs)" % (
            name, steps, len(eps), int(r["done"].any(1).sum()), int(r["check_goal"].sum()), sorted(set(d["ep_map"].tolist())), time.time() - t))


def main(names=None):
    for name, case in CASES.items():
        variant, n, seed, steps, ep_len, cl, sep, rays, mseed, policy = case[:10]
        evs = int(case[10]) if len(case) > 10 else 0
        if names and name not in names:
            continue
        t = time.time()
        gmap = synthetic_map(seed=mseed)
        r = H.rollout(variant, gmap, n, seed, steps, ep_len, cluster_radius=cl, n_rays=rays, cluster_min_sep=sep, policy=policy,
                      eval_by_step=bool(evs))
        d = pack(r)
        d["meta_variant"] = np.array(variant)
        d["meta"] = np.array([n, seed, steps, ep_len, rays, mseed] + ([evs] if evs else []))
        d["occ"] = gmap.occ
        d["bound"] = np.array(gmap.bound, dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print("%s: %d steps, %d episodes, done=%d goal=%d  (%.1fs)" % (
            name, steps, len(r["episodes"]), int(r["done"].any(1).sum()), int(r["check_goal"].sum()), time.time() - t))


if __name__ == "__main__":
    main(sys.argv[1:])
    main_mm(sys.argv[1:])
