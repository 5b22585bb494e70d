"""The device planner (aac_plan_paths_device, SURVEY 8f rank 1: the reference's grid search and collinear pruning,
ATT/jps_straight.py:17-70 + ATT:321-331, one warp per origin / destination pair) against the host planner aac_plan_path,
which tests/test_host_logic.py pins to the restatement that reproduces the reference's episodes.  Integer work: bit-exact."""
import ctypes as C
import types

import numpy as np
import pytest

from multi_agent_aac_b200 import _capi
from multi_agent_aac_b200.maps import multimap_set, synthetic_map
from multi_agent_aac_b200.reset import OdTable, plan_paths_device

pytestmark = pytest.mark.gpu


def host_plan(occ, pairs, w_max):
    lib = _capi.lib()
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    cells, length = np.zeros((len(pairs), w_max), dtype=np.uint16), np.zeros(len(pairs), dtype=np.int32)
    for k, (sx, sy, tx, ty) in enumerate(pairs):
        length[k] = lib.aac_plan_path(occ.ctypes.data, occ.shape[0], occ.shape[1], int(sx), int(sy), int(tx), int(ty), cells[k].ctypes.data, w_max)
        if length[k] <= 0:
            cells[k] = 0
    return cells, length


def grid(occ):
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    return types.SimpleNamespace(occ=occ, gx=occ.shape[0], gy=occ.shape[1])


def test_device_od_table_equals_host_od_table():
    for m in (synthetic_map(seed=0), multimap_set(seed=0)[2], multimap_set(seed=0)[6]):
        a, b = OdTable(m, w_max=32, planner="host"), OdTable(m, w_max=32, planner="device")
        for f in ("n_cells", "pool_off", "cell_code", "path_off", "path_len", "path_cells"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f
        assert int(a.path_len.astype(np.int64).sum()) > 0


def test_device_planner_equals_the_reference_jps():
    """The device search against paths of the UNMODIFIED reference jps_find_path (tests/golden/jps_paths.npz, made by
    tests/golden/gen_golden_jps.py): 1200 pairs on three grids, 112 of them unreachable; the pruned line expanded back to cells
    is the reference's path cell for cell."""
    import os
    from tests.test_oracle_golden import _expand
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jps_paths.npz"))
    for name in ("single", "multi5", "walled"):
        occ, pairs, flat, off = z[name + "_occ"], z[name + "_pairs"], z[name + "_cells"], z[name + "_off"]
        cells, length = plan_paths_device(grid(occ), pairs, 64)
        for k in range(len(pairs)):
            want = [divmod(int(c), 256) for c in flat[off[k]:off[k + 1]]]
            if not want:
                assert length[k] == 0, (name, k)
                continue
            got = [divmod(int(c), 256) for c in cells[k, :length[k]]]
            assert _expand(got) == want, (name, k)


def test_device_planner_edge_cases():
    # a walled pocket (unreachable goal), start == goal, adjacent cells, a start on an occupied cell, and a path
    # with more vertices than max_cells
    occ = np.zeros((9, 7), dtype=np.uint8)
    occ[3, 2:5] = occ[5, 2:5] = 1
    occ[4, 2] = occ[4, 4] = 1                     # (4, 3) is enclosed
    occ[7, 0:6] = 1                               # a wall with a gap at the top: detours with several turns
    pairs = np.array([[0, 0, 4, 3], [4, 3, 0, 0], [2, 2, 2, 2], [0, 0, 0, 1], [1, 0, 0, 0], [3, 3, 0, 0], [0, 0, 8, 0], [8, 0, 0, 6],
                      [0, 6, 8, 6], [6, 0, 6, 6]])
    for w_max in (32, 3, 2, 1):
        want_c, want_n = host_plan(occ, pairs, w_max)
        got_c, got_n = plan_paths_device(grid(occ), pairs, w_max)
        assert np.array_equal(want_n, got_n), (w_max, want_n, got_n)
        ok = want_n > 0
        assert np.array_equal(want_c[ok], got_c[ok])
    _, n32 = plan_paths_device(grid(occ), pairs, 32)
    assert n32[0] == 0 and n32[1] == 0 and n32[2] == 1 and n32[3] == 2 and n32[6] > 2
    _, n3 = plan_paths_device(grid(occ), pairs, 3)
    assert (n3 == -1).any()


@pytest.mark.parametrize("shape,density,seed", [((23, 13), 0.25, 1), ((31, 21), 0.3, 2), ((200, 150), 0.2, 3), ((255, 255), 0.1, 4), ((1, 40), 0.0, 5), ((40, 1), 0.0, 6)])
def test_device_planner_random_grids(shape, density, seed):
    # includes grids far larger than the reference's (the all-pairs table is not an option there): scratch lives in
    # global memory, 9 bytes per cell per warp
    rng = np.random.default_rng(seed)
    occ = (rng.random(shape) < density).astype(np.uint8)
    free = np.argwhere(occ == 0)
    n = 400 if shape[0] * shape[1] < 2000 else 120
    pairs = np.concatenate([free[rng.integers(0, len(free), n)], free[rng.integers(0, len(free), n)]], axis=1)
    want_c, want_n = host_plan(occ, pairs, 64)
    got_c, got_n = plan_paths_device(grid(occ), pairs, 64)
    assert np.array_equal(want_n, got_n)
    ok = want_n > 0
    assert np.array_equal(want_c[ok], got_c[ok])
    assert (got_c[~ok] == 0).all()
    assert ok.any()


def test_device_planner_rejects_bad_arguments():
    occ = np.zeros((4, 4), dtype=np.uint8)
    with pytest.raises(_capi.AacError):
        plan_paths_device(grid(occ), np.array([[0, 0, 4, 0]]), 8)      # goal outside the grid
    lib = _capi.lib()
    assert lib.aac_plan_paths_device(None, 4, 4, None, 1, None, None, 8, None) == -1
    c, n = plan_paths_device(grid(occ), np.zeros((0, 4), dtype=np.int64), 8)
    assert c.shape == (0, 8) and n.shape == (0,)


@pytest.mark.gpu
# (shapes that run the same run-time-shape kernels with and without the search: the instantiations specialised on the drone
#  count keep their neighbour distances with the 5 low mantissa bits replaced by the sort index)
@pytest.mark.parametrize("variant,n,r,E", [("tdcpa_v2", 7, 24, 3000), ("tdcpa_v2", 5, 24, 1500), ("att", 4, 18, 2000), ("multimap", 2, 18, 3000)])
def test_per_episode_planning_equals_the_table(variant, n, r, E):
    """Pools-only origin / destination tables: every episode's reference lines are searched on the device by the warp that
    re-initialises the env (reset_world's per-episode jps_find_path, ATT:317-331).  State, observations, rewards and counters
    must equal those of a handle that looks the same paths up in the all-pairs table, bit for bit, through resets and
    auto-resets; no line may have fallen back."""
    import numpy as np
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import multimap_set, synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    maps = multimap_set(seed=0)[:4] if variant == "multimap" else [synthetic_map(seed=0)]
    envs = []
    for paths in (True, False):
        env = BatchedDroneEnv(preset(variant, n_envs=E, n_agents=n, n_rays=r, w_max=32, seed=31), maps if variant == "multimap" else maps[0])
        env.set_od_tables([OdTable(m, w_max=32, planner="device", paths=paths) for m in maps])
        env.reset()
        envs.append(env)
    def same(tag):
        for k in envs[0].out:
            assert torch.equal(envs[0].out[k].view(torch.uint8), envs[1].out[k].view(torch.uint8)), (tag, k)
        for k in envs[0].state:
            assert torch.equal(envs[0].state[k].view(torch.uint8), envs[1].state[k].view(torch.uint8)), (tag, k)
    same("reset")
    gen = torch.Generator(device="cuda")
    gen.manual_seed(2)
    for t in range(15):
        act = (torch.rand((E, n, 2), device="cuda", generator=gen) * 2 - 1).contiguous()
        for env in envs:
            env.step(act, autoreset=True)
        same(t)
    mask = (torch.arange(E, device="cuda") % 3 == 0).to(torch.uint8)
    for env in envs:
        env.reset(mask)
    same("masked reset")
    s0, s1 = envs[0].read_stats(), envs[1].read_stats()
    assert s0[0] == s1[0] > 0 and s1[10] == 0 and np.array_equal(s0[[1, 3, 4, 5, 6, 7, 8, 9]], s1[[1, 3, 4, 5, 6, 7, 8, 9]])
