"""CPU: the float64 oracle against the rollouts recorded from the UNMODIFIED reference classes
(tests/golden/*.npz, made by tests/golden/gen_golden.py), plus properties the domain offers."""
import glob
import os

import numpy as np
import pytest

from multi_agent_aac_b200.maps import synthetic_map
from oracle.oracle import OracleEnv, RADAR_MIN
from tests.replay import GOLDEN_DIR, load_case, load_case_mm, replay, replay_mm

ALL = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not os.path.basename(f).startswith(("actor", "cs_", "mapgen", "jps_")))
GOLDEN = [n for n in ALL if not n.startswith("mm_")]
GOLDEN_MM = [n for n in ALL if n.startswith("mm_")]


def test_fixture_inventory():
    assert len(GOLDEN) >= 10 and len(GOLDEN_MM) >= 3
    assert any(n.startswith("att") for n in GOLDEN) and any(n.startswith("v2") for n in GOLDEN)


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_replays_reference_rollout(name):
    d, variant, n, rays, ep_len, gmap = load_case(name)
    env = OracleEnv(variant, gmap, 1, n, rays, eval_by_step=len(d["meta"]) > 6 and bool(d["meta"][6]))
    diff = replay(env, d, variant, rtol=1e-9, atol=1e-9)
    assert not diff.fail, "\n".join(diff.fail[:10])


@pytest.mark.parametrize("name", GOLDEN_MM)
def test_oracle_replays_multimap_reference_rollout(name):
    d, n, rays, ep_len, maps = load_case_mm(name)
    env = OracleEnv("mm", maps, 1, n, rays)
    diff = replay_mm(env, d, rtol=1e-9, atol=1e-9)
    assert not diff.fail, "\n".join(diff.fail[:10])
    assert len(set(d["ep_map"].tolist())) >= 4          # several maps per rollout


def _two_drone_env(variant, p0, p1, goal0=(640.0, 320.0), rays=18):
    gmap = synthetic_map(seed=0)
    gmap.occ[:] = 0
    env = OracleEnv(variant, gmap, 1, 2, rays)
    lines = [np.array([[470.0, 270.0], goal0]), np.array([[470.0, 370.0], [640.0, 370.0]])]
    env.set_episode(0, [p0, p1], lines, [0.0, 0.0])
    return env


def test_kat_geometry_test_sample_reaches_goal():
    """ATT/geometry_test.py:12-14: cur=(534.12,355.86), goal=(536,356): centre distance 1.885 < 3.5."""
    env = _two_drone_env("att", (534.12, 355.86), (600.0, 300.0), goal0=(536.0, 356.0))
    env.state["ref_line"][0, 0, 1] = (536.0, 356.0)
    env.observe()
    out = env.step(np.zeros((1, 2, 2)))
    assert out["check_goal"][0, 0] == 1 and out["branch"][0, 0] == 3 and out["done"][0, 0] == 0
    assert env.state["reach"][0, 0] == 1


def test_goal_polygon_band():
    """64-gon(2.5) n 64-gon(1) is non-empty up to 3.5*cos(pi/64) in every direction, empty
    beyond 3.5, and reaches 3.5 exactly along the shared vertex directions (angle 0)."""
    for dist, ang, want in ((3.4957, 0.3, 1), (3.5001, 0.0, 0), (3.4999, 0.0, 1), (3.4990, np.pi / 64, 0)):
        g = (500.0, 300.0)
        p = (g[0] + dist * np.cos(ang), g[1] + dist * np.sin(ang))
        env = _two_drone_env("att", p, (600.0, 350.0), goal0=g)
        env.observe()
        out = env.step(np.zeros((1, 2, 2)))
        assert out["check_goal"][0, 0] == want, (dist, ang)


def test_tcpa_special_value_iff_zero_relative_velocity():
    env = _two_drone_env("att", (500.0, 300.0), (520.0, 300.0))
    env.observe()
    out = env.step(np.zeros((1, 2, 2)))          # both at rest
    assert (out["tcpa"][0, :, 0, 0] == -10.0).all()
    assert np.allclose(out["tcpa"][0, :, 0, 1], 20.0)
    act = np.zeros((1, 2, 2))
    act[0, 0, 0] = 1.0                           # drone 0 accelerates towards drone 1
    out = env.step(act)
    assert (out["tcpa"][0, :, 0, 0] != -10.0).all()
    v = env.state["vel"][0, 0, 0]
    assert np.isclose(out["tcpa"][0, 0, 0, 0], (20.0 - v * 0.5) / v)


def test_radar_properties_and_crash_flags():
    rng = np.random.default_rng(0)
    gmap = synthetic_map(seed=0)
    from multi_agent_aac_b200.reset import ScenarioBank
    for variant in ("att", "v2"):
        E, N, R = 32, 4, 36
        bank = ScenarioBank(gmap, N, E, w_max=32, seed=3)
        env = OracleEnv(variant, gmap, E, N, R, radar_mode=RADAR_MIN)
        g = gmap.grid_length
        for e in range(E):
            lines = []
            for i in range(N):
                w = int(bank.w[e, i])
                c = bank.cells[e, i, :w].astype(np.int64)
                lines.append(np.stack([gmap.x0c + (c >> 8) * g, gmap.y0c + (c & 255) * g], -1).astype(np.float64))
            env.set_episode(e, [l[0] for l in lines], lines, [0.0] * N)
        env.observe()
        for t in range(40):
            out = env.step(rng.uniform(-1, 1, size=(E, N, 2)))
            r = out["radar"]
            assert np.nanmax(r) <= 15.0 + 1e-9 and np.nanmin(r) >= 0.0
            done_env = out["done"].any(axis=1)
            assert (out["bbc"][done_env, :3].any(axis=1)).all(), "done implies a bound/building/drone flag"
            assert (~out["bbc"][~done_env, :3].any(axis=1)).all()
            assert np.all(np.linalg.norm(env.state["vel"], axis=-1) <= 5.0 + 1e-9)


def load_sensor_fixture():
    """tests/golden/cs_sensors.npz: vectors of the later fork's unmodified class (tests/golden/gen_golden_sensors.py)."""
    import numpy as np
    from multi_agent_aac_b200.maps import GridMap
    d = np.load(os.path.join(GOLDEN_DIR, "cs_sensors.npz"))
    n, r, n_neigh = (int(v) for v in d["meta"][:3])
    gmap = GridMap(bound=[float(v) for v in d["bound"]], grid_length=10, occ=d["occ"].astype(np.uint8))
    return d, n, r, n_neigh, gmap


def test_sensor_classes_match_the_later_forks_class():
    """Clouds' kinematics, the true-minimum radar over boundary segments / cloud outlines / other aircraft's outlines and the
    nearest-N neighbour block against vectors recorded from the fork's own `step` and `cur_state_norm_state_v3`."""
    import numpy as np
    from oracle.oracle import OracleEnv, RADAR_MIN
    d, n, r, n_neigh, gmap = load_sensor_fixture()
    clouds = [tuple(row) for row in d["cloud_cfg"]]
    prot = float(d["meta"][3]) if d["meta"][3] else 5.0       # the fork's protectiveBound (CS agent file: 5)
    orc = OracleEnv("v2", gmap, 1, n, r, w_max=32, radar_mode=RADAR_MIN, radar_targets=2 | 4 | 8, n_nbr_obs=n_neigh, clouds=clouds, prot=prot)
    noac = OracleEnv("v2", gmap, 1, n, r, w_max=32, radar_mode=RADAR_MIN, radar_targets=2 | 4, n_nbr_obs=n_neigh, clouds=clouds, prot=prot)
    for k in range(len(d["cloud_traj"])):                       # calculate_next_position, step after step
        assert np.allclose(orc.cloud_positions(k), d["cloud_traj"][k], rtol=0, atol=1e-9), k
    hits = 0
    for q in range(len(d["pos"])):
        for env, key in ((orc, "radar"), (noac, "radar_noac")):
            st = env.state
            st["pos"][0], st["vel"][0], st["heading"][0], st["ep_step"][0] = d["pos"][q], d["vel"][q], d["heading"][q], int(d["cloud_k"][q])
            st["ref_line"][0, :, 0], st["ref_line"][0, :, 1], st["ref_w"][0] = d["pos"][q], d["pos"][q] + 10.0, 2
            o = env.observe()
            assert np.allclose(o["radar"][0], d[key][q], rtol=1e-9, atol=1e-9), (q, key, np.abs(o["radar"][0] - d[key][q]).max())
        hits += int((d["radar"][q] < 15 - 1e-9).sum())
        assert np.allclose(o["raw_nbr"][0], d["raw_nbr"][q], rtol=1e-9, atol=1e-9), q
        assert np.allclose(o["norm_nbr"][0], d["norm_nbr"][q], rtol=1e-9, atol=1e-9), q
        assert o["norm_nbr"].shape[-1] == 5 * n_neigh
    assert hits > 0.5 * d["radar"].size


def _expand(cells):
    """Pruned line (vertices where the direction changes) -> every cell along it."""
    out = [cells[0]]
    for a, b in zip(cells[:-1], cells[1:]):
        dx, dy = np.sign(b[0] - a[0]), np.sign(b[1] - a[1])
        assert (dx == 0) != (dy == 0), (a, b)              # axis-parallel legs only
        x, y = a
        while (x, y) != tuple(b):
            x, y = x + dx, y + dy
            out.append((int(x), int(y)))
    return out


def test_planner_equals_the_reference_jps():
    """aac_plan_path (host; the search the origin / destination tables and the per-episode device search share) against paths
    of the UNMODIFIED reference jps_find_path (tests/golden/gen_golden_jps.py): the pruned line, expanded back to cells, is
    the reference's path cell for cell - same tie-breaking - and unreachable goals are reported as such."""
    import ctypes
    from multi_agent_aac_b200 import _capi
    _capi.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    z = np.load(os.path.join(GOLDEN_DIR, "jps_paths.npz"))
    n_none = 0
    for name in ("single", "multi5", "walled"):
        occ, pairs, flat, off = z[name + "_occ"], z[name + "_pairs"], z[name + "_cells"], z[name + "_off"]
        gx, gy = occ.shape
        buf = np.zeros(64, dtype=np.uint16)
        for k, (sx, sy, tx, ty) in enumerate(pairs):
            want = [divmod(int(c), 256) for c in flat[off[k]:off[k + 1]]]
            n = lib.aac_plan_path(occ.ctypes.data_as(ctypes.c_void_p), int(gx), int(gy), int(sx), int(sy), int(tx), int(ty),
                                  buf.ctypes.data_as(ctypes.c_void_p), 64)
            if not want:
                assert n == 0, (name, k)
                n_none += 1
                continue
            assert n >= 2, (name, k, n)
            got = [divmod(int(c), 256) for c in buf[:n]]
            assert got[0] == (sx, sy) and got[-1] == (tx, ty)
            assert _expand(got) == want, (name, k)
    assert n_none > 20
