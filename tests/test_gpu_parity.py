"""GPU parity tests: every call goes through the C ABI (libaac_env.so); the oracle is the checker."""
import glob
import os

import pytest

from tests import parity
from tests.replay import GOLDEN_DIR, load_case, load_case_mm, oracle_margins_mm, replay, replay_mm

pytestmark = pytest.mark.gpu

ALL_GOLDEN = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not os.path.basename(f).startswith(("actor", "cs_", "mapgen", "jps_")))
GOLDEN = [n for n in ALL_GOLDEN if not n.startswith("mm_")]
GOLDEN_MM = [n for n in ALL_GOLDEN if n.startswith("mm_")]


@pytest.fixture(scope="module", autouse=True)
def _built():
    from multi_agent_aac_b200 import _capi
    _capi.lib()  # raises if the extension is missing: there is no fallback


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_replay(name):
    """The unmodified reference's recorded rollouts, replayed through the CUDA env (teacher forced)."""
    d, variant, n, rays, ep_len, gmap = load_case(name)
    ad = parity.GpuGoldenAdapter(variant, gmap, n, rays, eval_by_step=len(d["meta"]) > 6 and bool(d["meta"][6]))
    diff = replay(ad, d, variant, rtol=1e-4, atol=2e-4, resync=ad.resync)
    assert not diff.fail, "\n".join(diff.fail[:10])


@pytest.mark.parametrize("name", GOLDEN_MM)
def test_golden_replay_multimap(name):
    """multipleMap reference rollouts (a different map per episode) through the CUDA env, teacher forced."""
    d, n, rays, ep_len, maps = load_case_mm(name)
    ad = parity.GpuGoldenAdapterMM(maps, n, rays)
    # ties are frequent in this variant: a full-speed first step that leaves the reference line beyond its start
    # vertex ends EXACTLY protectiveBound (= vmax * dt = 2.5 m) from it, where the cross-track reward jumps by 3
    # (MM:1812-1817); the float64 oracle's own margins say which steps those are
    diff = replay_mm(ad, d, rtol=1e-4, atol=2e-4, resync=ad.resync, margins=oracle_margins_mm(d, n, rays, maps))
    assert not diff.fail, "\n".join(diff.fail[:10])
    assert diff.ties <= 0.25 * d["actions"].shape[0]


def _run(**kw):
    T = parity.lockstep(**kw)
    print(T.summary())
    assert not T.fail, "\n".join(T.fail[:12])
    n_flags = T.n.get("done", 0)
    assert n_flags > 0
    # hit ids: bit-exact except oracle-proven ties, and those stay rare
    assert T.ties.get("radar_hit", 0) <= 0.002 * T.n.get("radar_hit", 0) + 2, T.summary()
    n_ties = T.ties.get("predicate_margin", 0) + T.ties.get("sort_order", 0)
    # multipleMap: cross-track error == protectiveBound is hit exactly by full-speed first steps (see
    # test_golden_replay_multimap), so a larger share of env-steps sits on a threshold
    if kw["variant"] == "mm":   # flags are masked per drone there; at least 80 % of drone-steps must have been compared
        assert n_flags >= 0.8 * kw["n_envs"] * kw["n_agents"] * kw["steps"], T.summary()
    else:
        assert n_ties * kw["n_agents"] <= 0.02 * (n_flags + n_ties * kw["n_agents"]) + 5, T.summary()
    return T


def test_lockstep_att_default():
    _run(variant="att", n_envs=256, n_agents=3, n_rays=18, steps=60, seed=1)


def test_lockstep_att_cluster_r36():
    _run(variant="att", n_envs=128, n_agents=5, n_rays=36, steps=40, seed=2, cluster=10.0)


def test_lockstep_v2_default():
    _run(variant="v2", n_envs=256, n_agents=3, n_rays=18, steps=110, seed=3)


def test_lockstep_v2_n10_r36_last_hit():
    _run(variant="v2", n_envs=256, n_agents=10, n_rays=36, steps=40, seed=4)


def test_lockstep_v2_n10_r36_true_min():
    _run(variant="v2", n_envs=128, n_agents=10, n_rays=36, steps=30, seed=5, radar_mode=parity.RADAR_MIN)


def test_lockstep_v2_cluster():
    _run(variant="v2", n_envs=128, n_agents=6, n_rays=36, steps=40, seed=6, cluster=12.0)


def test_lockstep_multimap():
    """14 maps of different sizes in one batch, map drawn per episode on the device."""
    _run(variant="mm", n_envs=256, n_agents=3, n_rays=18, steps=80, seed=10)


def test_lockstep_multimap_n8_r36():
    _run(variant="mm", n_envs=96, n_agents=8, n_rays=36, steps=40, seed=11)


def test_lockstep_ragged_tile_and_single_drone():
    _run(variant="v2", n_envs=37, n_agents=4, n_rays=18, steps=30, seed=7, tile_envs=5, block_threads=96)
    T = parity.lockstep(variant="v2", n_envs=9, n_agents=1, n_rays=18, steps=20, seed=8)
    assert not T.fail, "\n".join(T.fail[:12])


def test_lockstep_r72_n20():
    _run(variant="v2", n_envs=32, n_agents=20, n_rays=72, steps=20, seed=9)


SENSORS = dict(radar_targets=2 | 4 | 8, n_nbr_obs=2, prot=5.0, bound=[0, 200, 0, 200],
               clouds=((30.0, 185.0, 180.0, 80.0, 12.0, 2.0), (30.0, 100.0, 180.0, 30.0, 12.0, 2.0)))   # the later fork's training set-up (CS:600-602)


def test_sensor_classes_reference_vectors():
    """The later fork's sensor classes against vectors of its UNMODIFIED class (tests/golden/cs_sensors.npz):
    true-minimum radar over boundary segments, moving clouds' outlines and other aircraft's outlines, and the nearest-N
    neighbour block, observed by the CUDA env at the fixture's 320 states (all of them in one batch)."""
    import numpy as np
    import torch
    from multi_agent_aac_b200 import _capi as K
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from oracle.oracle import OracleEnv, RADAR_MIN
    from tests.test_oracle_golden import load_sensor_fixture
    d, n, r, n_neigh, gmap = load_sensor_fixture()
    E = len(d["pos"])
    clouds = tuple(tuple(float(v) for v in row) for row in d["cloud_cfg"])
    for targets, key in ((2 | 4 | 8, "radar"), (2 | 4, "radar_noac")):
        cfg = preset("changeskin_sensors", n_envs=E, n_agents=n, n_rays=r, w_max=32, radar_targets=targets, n_nbr_obs=n_neigh, clouds=clouds,
                     out_flags=K.OUT_RAW | K.OUT_RADAR_AUX)
        env = BatchedDroneEnv(cfg, gmap)
        env.load_agent_state(d["pos"], d["vel"], heading=d["heading"])
        env.state["ep_step"].copy_(torch.tensor(d["cloud_k"].astype(np.int32), device=env.device))
        env.observe()
        got = {k: v.cpu().numpy().astype(np.float64) for k, v in env.out.items()}
        assert got["norm_nbr"].shape == (E, n, 5 * n_neigh)
        assert np.allclose(got["raw_nbr"], d["raw_nbr"], rtol=1e-4, atol=1e-4) and np.allclose(got["norm_nbr"], d["norm_nbr"], rtol=1e-4, atol=2e-6)
        bad = np.abs(got["radar"] - d[key]) > 1e-4 * np.abs(d[key]) + 2e-4
        # every mismatch must be bracketed by the checker's own answers at positions displaced by the float32 resolution (grazing rays)
        orc = OracleEnv("v2", gmap, 1, n, r, w_max=32, radar_mode=RADAR_MIN, radar_targets=targets, n_nbr_obs=n_neigh, clouds=clouds, prot=cfg.prot)
        for q, i, k in zip(*np.nonzero(bad)):
            lo = hi = d[key][q, i, k]
            for dx, dy in ((5e-5, 0), (-5e-5, 0), (0, 5e-5), (0, -5e-5), (5e-5, 5e-5), (-5e-5, -5e-5), (5e-5, -5e-5), (-5e-5, 5e-5)):
                pp = d["pos"][q].copy()
                pp[i] += (dx, dy)
                v = orc.radar_probe(pp, i, k_cloud=int(d["cloud_k"][q]))[0][k]
                lo, hi = min(lo, v), max(hi, v)
            assert lo - 1e-3 <= got["radar"][q, i, k] <= hi + 1e-3, (key, q, i, k, got["radar"][q, i, k], d[key][q, i, k])
        assert bad.sum() <= 0.002 * bad.size
        assert (d[key] < 15 - 1e-9).mean() > 0.2      # the fixture's rays do hit things
        env.close()


def test_lockstep_sensor_classes():
    """Sensor configuration in lock step with the checker (itself pinned to the fork's class by tests/test_oracle_golden.py):
    clouds move with the episode clock, resets restart them, clustered starts make the aircraft outlines fire."""
    T = _run(variant="v2", n_envs=192, n_agents=4, n_rays=18, steps=70, seed=12, radar_mode=parity.RADAR_MIN, cluster=14.0, sensors=SENSORS)
    assert T.n.get("cloud_contact", 0) > 0
    _run(variant="v2", n_envs=96, n_agents=6, n_rays=36, steps=40, seed=13, radar_mode=parity.RADAR_MIN,
         sensors=dict(SENSORS, radar_targets=1 | 2 | 4 | 8, n_nbr_obs=3))      # + the grid cells of the older variants


def test_fused_autoreset_equals_step_then_autoreset():
    """aac_step_fused (one launch) and aac_step_autoreset must leave exactly the state and outputs of aac_step followed
    by aac_autoreset, bit for bit."""
    import numpy as np
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import ScenarioBank
    gmap = synthetic_map(seed=0)
    from multi_agent_aac_b200.maps import multimap_set
    from multi_agent_aac_b200.reset import MultiMapBank
    from multi_agent_aac_b200 import _capi as K
    for variant, n, r, E in (("tdcpa_v2", 10, 36, 300), ("att", 3, 18, 257), ("multimap", 3, 18, 203)):
        envs = []
        for idx in range(3):   # env 2: aac_step_autoreset as two launches whatever the batch size
            if variant == "multimap":
                maps = multimap_set(seed=0)
                cfg = preset(variant, n_envs=E, n_agents=n, n_rays=r, w_max=32, seed=5, out_flags=K.OUT_RAW | K.OUT_RADAR_AUX | K.OUT_PARTS,
                             autoreset_launches=2 if idx == 2 else 0)
                env = BatchedDroneEnv(cfg, maps)
                env.set_bank(MultiMapBank(maps, n, 64, w_max=32, seed=5))
            else:
                cfg = preset(variant, n_envs=E, n_agents=n, n_rays=r, w_max=32, seed=5, out_flags=parity.ALL_OUT, autoreset_launches=2 if idx == 2 else 0)
                env = BatchedDroneEnv(cfg, gmap)
                env.set_bank(ScenarioBank(gmap, n, 64, w_max=32, seed=5))
            env.reset()
            envs.append(env)
        gen = torch.Generator(device="cuda")
        gen.manual_seed(3)
        n_term = 0
        for t in range(25):
            act = (torch.rand((E, n, 2), device="cuda", generator=gen) * 2 - 1).contiguous()
            envs[0].step(act, autoreset=False)
            term = envs[0].out["terminated"].clone()
            envs[0].autoreset()
            envs[1].step(act, autoreset=True, fused=True)
            envs[2].step(act, autoreset=True)
            n_term += int((term != 0).sum())
            for other in envs[1:]:
                for k in envs[0].out:
                    a, b = envs[0].out[k], other.out[k]
                    assert torch.equal(a.view(torch.uint8), b.view(torch.uint8)), (variant, t, k)
                for k in envs[0].state:
                    assert torch.equal(envs[0].state[k].view(torch.uint8), other.state[k].view(torch.uint8)), (variant, t, k)
        assert n_term > 0
        s0 = envs[0].read_stats()
        for other in envs[1:]:
            s1 = other.read_stats()
            assert np.array_equal(s0[[0, 1, 3, 4, 5, 6, 7, 8, 9]], s1[[0, 1, 3, 4, 5, 6, 7, 8, 9]]) and s0[0] == n_term
            assert abs(s0[2] - s1[2]) <= 1e-3 * max(1.0, abs(s0[2]))


def test_partial_reset_leaves_other_envs_untouched():
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import ScenarioBank
    gmap = synthetic_map(seed=0)
    E, n, r = 50, 4, 18
    env = BatchedDroneEnv(preset("tdcpa_v2", n_envs=E, n_agents=n, n_rays=r, w_max=32, seed=9), gmap)
    env.set_bank(ScenarioBank(gmap, n, 32, w_max=32, seed=9))
    env.reset()
    act = torch.zeros((E, n, 2), device="cuda")
    act[..., 0] = 0.3
    env.step(act)
    before_state = {k: v.clone() for k, v in env.state.items()}
    before_out = {k: v.clone() for k, v in env.out.items()}
    mask = torch.zeros(E, dtype=torch.uint8, device="cuda")
    mask[[3, 17, 18, 49]] = 1
    env.reset(mask)
    keep = mask == 0
    for k, v in env.state.items():
        assert torch.equal(v[keep], before_state[k][keep]), k
    for k in ("norm_own", "norm_nbr", "radar", "reward", "done"):
        assert torch.equal(env.out[k][keep], before_out[k][keep]), k
    assert (env.state["ep_step"][mask != 0] == 0).all() and (env.state["ep_index"][mask != 0] == 2).all()
    assert (env.state["vx"][mask != 0] == 0).all()


@pytest.mark.parametrize("name", ["att_n3_plain", "v2_n3_plain", "v2_n3_evalstep_seek"])
def test_ref_compat_env_reproduces_reference_episode(name):
    """The drop-in class (reference method names / nested-list tuples, E = 1): with `random.seed(k)` it draws
    the reference's first episode, and free-running on the recorded actions it tracks the reference's
    float64 rollout for that episode."""
    import random
    from types import SimpleNamespace
    import numpy as np
    from multi_agent_aac_b200.ref_compat import RefCompatEnv, RefCompatEnvV2
    d, variant, n, rays, ep_len, gmap = load_case(name)
    seed = int(d["meta"][1])
    cls = RefCompatEnv if variant == "att" else RefCompatEnvV2
    env = cls(gmap.occ.astype(float), [], gmap.grid_length, list(gmap.bound), None, None, n_rays=rays)
    env.create_world(n, 2, 0.95, 0.01, 1, 0.15, 0.05, 0.15, (1800, 1300), 5, [-8, 8])
    random.seed(seed)
    state, norm_state = env.reset_world(n, None if variant == "att" else False, 0)
    assert np.allclose(np.array([env.all_agents[i].pos for i in range(n)]), d["ep_start"][0])
    for i in range(n):
        w = int(d["ep_ref_w"][0, i])
        assert np.allclose(env.all_agents[i].ref_line, d["ep_ref_line"][0, i, :w])
    assert len(state) == (3 if variant == "att" else 4) and len(state[0]) == n
    assert np.allclose(np.stack(norm_state[0]), d["ep_norm_0"][0], rtol=1e-4, atol=1e-5)
    T = int(np.sum(d["episode_id"] == 0))
    for t in range(T):
        srr, scr = [None] * n, [[] for _ in range(n)]
        if variant == "att":
            out = env.step(d["actions"][t], t + 1, 8, None)
            rw = env.ss_reward(t + 1, srr, [None] * n, scr, (None, None), True, None)
        else:
            by_step = len(d["meta"]) > 6 and bool(d["meta"][6])   # forV2 evaluation "by sorties"
            args = SimpleNamespace(mode="eval" if by_step else "train")
            out = env.step(d["actions"][t], t + 1, 8, args, not by_step, False)
            rw = env.ss_reward_Mar(t + 1, srr, scr, (None, None), False, args, not by_step)
        assert len(out) == 8 and len(rw) == 7
        reward, done, check_goal, _, _, _, bbc = rw
        assert np.allclose(np.stack(out[0][0])[:, :4], d["raw_own"][t][:, :4], rtol=1e-4, atol=2e-3), t
        assert np.allclose([float(r) for r in reward], d["reward"][t], rtol=1e-3, atol=5e-3), t
        assert list(done) == [bool(v) for v in d["done"][t]] and list(bbc) == [bool(v) for v in d["bbc"][t]], t
        assert isinstance(reward[0], np.ndarray) and reward[0].ndim == 0 and isinstance(done[0], bool)


@pytest.mark.parametrize("n_envs,launches", [(100, 0), (7001, 0), (7001, 2), (7001, 3)])
def test_step_host_pipeline_equals_device_step(n_envs, launches):
    """aac_step_host (pinned host buffers, chunks pipelined over three streams for large batches) must return
    exactly what aac_step_autoreset leaves on the device."""
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import ScenarioBank
    gmap = synthetic_map(seed=0)
    n, r = 10, 36
    envs = []
    for _ in range(2):
        env = BatchedDroneEnv(preset("tdcpa_v2", n_envs=n_envs, n_agents=n, n_rays=r, w_max=32, seed=3, autoreset_launches=launches), gmap)
        env.set_bank(ScenarioBank(gmap, n, 64, w_max=32, seed=3))
        env.reset()
        envs.append(env)
    host = envs[1].host_buffers(("norm_own", "norm_nbr", "radar", "reward", "done", "check_goal", "bbc", "terminated", "tcpa_min"))
    gen = torch.Generator(device="cpu")
    gen.manual_seed(1)
    for t in range(6):
        act = (torch.rand((n_envs, n, 2), generator=gen) * 2 - 1).contiguous().pin_memory()
        envs[0].step(act.cuda(), autoreset=True)
        out = envs[1].step_host(act, autoreset=True)
        for k, v in out.items():
            assert torch.equal(v.view(torch.uint8), envs[0].out[k].cpu().view(torch.uint8)), (t, k)
        for k in envs[0].state:
            assert torch.equal(envs[0].state[k].view(torch.uint8), envs[1].state[k].view(torch.uint8)), (t, k)
    assert abs(envs[0].read_stats()[0] - envs[1].read_stats()[0]) == 0


@pytest.mark.parametrize("variant", ["tdcpa_v2", "multimap"])
def test_device_side_origin_destination_sampling(variant):
    """With an origin / destination table the device draws every episode itself (ATT:254-276): starts in the quadrant
    pools, more than 2 * protectiveBound apart, goal in a different quadrant, reference line = the host planner's
    path for that pair; every pool cell gets used; results do not depend on how envs are sharded."""
    import numpy as np
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import multimap_set, synthetic_map
    from multi_agent_aac_b200.reset import OdTable, ref_line_cells
    maps = multimap_set(seed=0)[:5] if variant == "multimap" else [synthetic_map(seed=0)]
    tabs = [OdTable(m, w_max=32) for m in maps]
    E, n, r = 4096, 4, 18

    def make(n_envs, base):
        env = BatchedDroneEnv(preset(variant, n_envs=n_envs, n_agents=n, n_rays=r, w_max=32, seed=77, env_id_base=base),
                              maps if variant == "multimap" else maps[0])
        env.set_od_tables(tabs)
        env.reset()
        return env
    env = make(E, 0)
    s = env.agent_state()
    cells = env.state["ref_cells"].cpu().numpy().view(np.uint16)
    map_id = s["map_id"] if variant == "multimap" else np.zeros(E, dtype=np.int64)
    used = [set() for _ in maps]
    for e in range(0, E, 7):
        m, tab = maps[map_id[e]], tabs[map_id[e]]
        pool_of = {int(c): int(np.searchsorted(tab.pool_off, k, side="right") - 1) for k, c in enumerate(tab.cell_code)}
        starts = []
        for i in range(n):
            w = int(s["ref_w"][e, i])
            line = [divmod(int(c), 256) for c in cells[e, i, :w]]
            c0, c1 = int(cells[e, i, 0]), int(cells[e, i, w - 1])
            assert c0 in pool_of and c1 in pool_of and pool_of[c0] != pool_of[c1]
            assert line == [tuple(c) for c in ref_line_cells(m, m.cell_centre(*line[0]), m.cell_centre(*line[-1]))]
            assert np.allclose(s["pos"][e, i], m.cell_centre(*line[0]), atol=1e-4)
            starts.append(np.array(m.cell_centre(*line[0])))
            used[map_id[e]].add(c0)
        for i in range(n):
            for j in range(i):
                assert np.hypot(*(starts[i] - starts[j])) > 5.0
    if variant != "multimap":
        assert len(used[0]) > 0.8 * tabs[0].n_cells
    else:
        assert len(set(map_id.tolist())) == len(maps)
    # sharding independence: envs [1000, 1500) of a second handle with env_id_base = 1000
    env2 = make(500, 1000)
    for k in ("px", "py", "ref_cells", "ref_w"):
        assert torch.equal(env.state[k][1000:1500], env2.state[k]), k
    # episodes keep being drawn on the device through the fused auto-reset
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    for t in range(10):
        env.step((torch.rand((E, n, 2), device="cuda", generator=gen) * 2 - 1).contiguous(), autoreset=True)
    assert env.read_stats()[0] > 0 and int(env.state["ep_index"].max()) >= 2


def test_ref_compat_env_multimap_reproduces_reference_episode():
    import random
    import numpy as np
    from multi_agent_aac_b200.ref_compat import RefCompatEnvMM
    d, n, rays, ep_len, maps = load_case_mm("mm_n3_seek")
    seed = int(d["meta"][1])
    env = RefCompatEnvMM({k: m.occ.astype(float) for k, m in enumerate(maps)}, [], 10, {k: list(m.bound) for k, m in enumerate(maps)},
                         None, None, None, n_rays=rays)
    env.create_world(n, 2, 0.95, 0.01, 1, 0.15, 0.05, 0.15, (1800, 1300), 5, [-4, 4])
    random.seed(seed)
    map_idx = random.randrange(len(maps))                 # MM/ma_main:464, drawn by the caller
    assert map_idx == int(d["ep_map"][0])
    state, norm_state = env.reset_world(n, map_idx, 0)
    assert np.allclose(np.array([env.all_agents[i].pos for i in range(n)]), d["ep_start"][0])
    assert np.allclose(np.stack(norm_state[0]), d["ep_norm_own"][0], rtol=1e-4, atol=1e-5)
    margins = oracle_margins_mm(d, n, rays, maps)
    T = int(np.sum(d["episode_id"] == 0))
    for t in range(T):
        srr, scr = [None] * n, [[] for _ in range(n)]
        out = env.step(d["actions"][t], t + 1, map_idx)
        reward, done, check_goal, _, _, _, bbc = env.ss_reward(t + 1, srr, [None] * n, scr, map_idx)
        assert np.allclose(np.stack(out[0][0])[:, :4], d["raw_own"][t][:, :4], rtol=1e-4, atol=5e-3), t
        if (margins[t] >= 1e-3).all():
            assert np.allclose([float(r) for r in reward], d["reward"][t], rtol=1e-3, atol=2e-2), t
            assert list(done) == [bool(v) for v in d["done"][t]] and list(bbc) == [bool(v) for v in d["bbc"][t]], t
            assert list(check_goal) == [bool(v) for v in d["check_goal"][t]], t


def test_state_dict_roundtrip():
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    gmap = synthetic_map(seed=0)
    tab = OdTable(gmap, w_max=32)
    envs = []
    for _ in range(2):
        env = BatchedDroneEnv(preset("tdcpa_v2", n_envs=64, n_agents=5, n_rays=36, w_max=32, seed=4), gmap)
        env.set_od_tables([tab])
        env.reset()
        envs.append(env)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(2)
    acts = [(torch.rand((64, 5, 2), device="cuda", generator=gen) * 2 - 1).contiguous() for _ in range(12)]
    for t in range(6):
        envs[0].step(acts[t], autoreset=True)
    envs[1].load_state_dict(envs[0].state_dict())
    for t in range(6, 12):
        envs[0].step(acts[t], autoreset=True)
        envs[1].step(acts[t], autoreset=True)
        for k in ("norm_own", "norm_nbr", "radar", "reward", "done"):
            assert torch.equal(envs[0].out[k], envs[1].out[k]), (t, k)


def test_full_size_c2_lockstep():
    """BASELINE config C2 at full size (one_model_att, 4096 envs x 3 drones x 36 rays): EVERY env of the batch against
    the oracle in lock step, auto-reset included."""
    _run(variant="att", n_envs=4096, n_agents=3, n_rays=36, steps=12, seed=31)


def test_full_size_c4_lockstep():
    """BASELINE config C4 at full size (radar_multipleMap, 65 536 envs x 3 drones x 18 rays, 14 heterogeneous maps drawn
    per episode): every env of the batch against the oracle in lock step."""
    T = parity.lockstep(variant="mm", n_envs=65536, n_agents=3, n_rays=18, steps=4, seed=32)
    print(T.summary())
    assert not T.fail, "\n".join(T.fail[:12])
    # every episode is young here: the exact cross-track == protectiveBound tie of a full-speed first step (see
    # test_golden_replay_multimap) masks a larger share of drone-steps than in the long small-batch runs
    assert T.n.get("done", 0) >= 0.7 * 65536 * 3 * 4, T.summary()


@pytest.mark.parametrize("E,N,R,n_check,n_steps", [(65536, 10, 36, 4096, 6), (131072, 20, 72, 2048, 4)])
def test_full_size_c3_properties_and_oracle_check(E, N, R, n_check, n_steps):
    """BASELINE config C3 at full size (65 536 envs x 10 drones x 36 rays) and one GPU's shard of C5 (1M envs x 20 drones x
    72 rays over 8 GPUs), on the path bench.py times: no optional outputs (lean kernels), the mode-specialised
    instantiations - C3: the phased launch (step loop + reset loop in one kernel), C5: step launch + reset launch.  Every step: size-independent properties on every env, determinism
    across two handles, and `n_check` randomly chosen envs (4096 for C3) against the oracle - the transition (reward,
    done, goal, bound_building_check, terminated and the stepped observation) from a snapshot taken before the step,
    and the reset observation of the envs that finished from the state the reset left."""
    import numpy as np
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    from oracle.oracle import OracleEnv, RADAR_LAST_HIT
    gmap = synthetic_map(seed=0)
    tab = OdTable(gmap, w_max=32)
    cfg = preset("tdcpa_v2", n_envs=E, n_agents=N, n_rays=R, w_max=32, seed=21)      # out_flags = 0: the benchmark's kernels
    envs = [BatchedDroneEnv(cfg, gmap) for _ in range(2)]
    for env in envs:
        env.set_od_tables([tab])
        env.reset()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    hx, hy = 0.5 * (gmap.bound[1] - gmap.bound[0]), 0.5 * (gmap.bound[3] - gmap.bound[2])
    idx = np.sort(torch.randperm(E, generator=torch.Generator().manual_seed(1))[:n_check].numpy())
    orc = OracleEnv("v2", gmap, n_check, N, R, w_max=32, radar_mode=RADAR_LAST_HIT)

    def load_oracle():
        s = envs[0].agent_state()
        cells = envs[0].state["ref_cells"].cpu().numpy().view(np.uint16)
        for k in ("pos", "vel", "heading", "reach", "wp_cur", "wall_cnt", "vflags"):
            orc.state[k][:] = s[k][idx]
        pn = s["prev_nn"][idx].copy()
        pn[pn == 255] = -1
        orc.state["prev_nn"][:] = pn
        orc.state["ref_line"][:] = parity.cells_to_lines(gmap, cells[idx], None)
        orc.state["ref_w"][:] = s["ref_w"][idx]

    def close(a, b, atol):
        a, b = a.astype(np.float64), b.astype(np.float64)
        return (np.abs(a - b) <= 1e-4 * np.abs(b) + atol) | (np.isnan(a) & np.isnan(b))

    def sort_ok(pos, order):
        d = np.linalg.norm(pos[:, :, None, :] - pos[:, None, :, :], axis=-1)
        ds = np.take_along_axis(d, order.astype(np.int64), axis=2)
        gap = np.min(np.diff(ds, axis=2), axis=(1, 2))
        return (gap >= parity.TIE_EPS) | (gap == 0)

    # rays at 45 degrees (they exist when 8 divides the ray count) from a cell centre - where every reset puts a drone - run
    # exactly through grid corners: whether the cells around the corner are touched is a boundary-epsilon tie by construction
    # (the lock-step tests bracket such rays one by one); here they are left out of the per-env radar verdict
    diag = (np.arange(R) * (360 // R)) % 90 == 45
    n_cmp = n_term = n_radar_tie = 0
    launches0 = envs[0].launch_count
    for t in range(n_steps):
        act = (torch.rand((E, N, 2), device="cuda", generator=gen) * 2 - 1).contiguous()
        load_oracle()
        ep_step0 = envs[0].state["ep_step"].cpu().numpy()[idx]
        for env in envs:
            env.step(act, autoreset=True)
        o = envs[0].out
        # determinism: two handles, same seed and inputs, bit-identical
        for k in ("norm_own", "norm_nbr", "radar", "reward", "done", "terminated"):
            assert torch.equal(o[k], envs[1].out[k]), (t, k)
        radar = o["radar"]
        assert float(radar[~radar.isnan()].max()) <= 15.0 and float(radar[~radar.isnan()].min()) >= 0.0
        spd = torch.sqrt(envs[0].state["vx"] ** 2 + envs[0].state["vy"] ** 2)
        assert float(spd.max()) <= 5.0 * (1 + 1e-5)
        done_env = o["done"].any(dim=1)
        assert bool((o["bbc"][:, :3].any(dim=1) == done_env).all())            # done <=> a bound / building / drone flag
        assert bool((((o["terminated"] >> 1) & 1).bool() == done_env).all())
        assert bool(torch.isfinite(o["reward"]).all())
        term_all = o["terminated"] != 0
        # observations describe the state the step (or the reset) left
        assert torch.allclose(o["norm_own"][..., 0], envs[0].state["px"] / hx, atol=1e-6)
        assert torch.allclose(o["norm_own"][..., 1], envs[0].state["py"] / hy, atol=1e-6)
        nb = o["norm_nbr"].view(E, N, N - 1, 5)
        d2 = (nb[..., 0] * hx) ** 2 + (nb[..., 1] * hy) ** 2
        assert bool((d2[..., 1:] >= d2[..., :-1] * (1 - 1e-5) - 1e-4).all())      # neighbour blocks sorted by distance
        assert bool((envs[0].state["ep_step"][term_all] == 0).all()) and bool((envs[0].state["vx"][term_all] == 0).all())
        # ---- the sampled envs against the oracle: the transition
        want = {k: v.copy() for k, v in orc.step(act.cpu().numpy().astype(np.float64)[idx]).items()}
        got = {k: v.cpu().numpy()[idx] for k, v in o.items()}
        stepped_pos = orc.state["pos"].copy()
        term_o = ((ep_step0 + 1 > cfg.episode_length).astype(np.int64) | (want["done"].any(axis=1).astype(np.int64) << 1)
                  | (orc.state["reach"].all(axis=1).astype(np.int64) << 2))
        ok = sort_ok(stepped_pos, want["nbr_order"]) & (want["margin"] >= parity.TIE_EPS).all(axis=1)     # ties masked exactly as in the lock-step tests
        for k in ("done", "check_goal", "bbc"):
            assert np.array_equal(got[k][ok].astype(np.int64), want[k][ok].astype(np.int64)), (t, k)
        assert np.array_equal(got["terminated"][ok].astype(np.int64), term_o[ok]), t
        assert close(got["reward"][ok], want["reward"][ok], 2e-4).all(), t
        n_cmp += int(ok.sum())
        alive = ok & (got["terminated"] == 0)
        rad_ok = (close(got["radar"], want["radar"], 2e-4) | diag).all(axis=(1, 2))   # grazing rays are classified by the lock-step tests
        n_radar_tie += int((alive & ~rad_ok).sum())
        for k, atol in (("norm_own", 2e-6), ("norm_nbr", 2e-6)):
            assert close(got[k][alive], want[k][alive], atol).all(), (t, k)
        # ---- ... and the reset observation of the sampled envs that finished
        fin = np.nonzero(got["terminated"] != 0)[0]
        n_term += len(fin)
        if len(fin):
            load_oracle()
            wr = {k: v.copy() for k, v in orc.observe().items()}
            okr = np.zeros(n_check, dtype=bool)
            okr[fin] = True
            okr &= sort_ok(orc.state["pos"], wr["nbr_order"])
            rad_ok = (close(got["radar"], wr["radar"], 2e-4) | diag).all(axis=(1, 2))
            n_radar_tie += int((okr & ~rad_ok).sum())
            for k, atol in (("norm_own", 2e-6), ("norm_nbr", 2e-6)):
                assert close(got[k][okr], wr[k][okr], atol).all(), (t, k, "reset")
    print({"oracle_checked_env_steps": n_cmp, "of": n_check * n_steps, "finished": n_term, "radar_tie_envs": n_radar_tie})
    assert n_cmp >= 0.9 * n_check * n_steps and n_term > 0
    assert n_radar_tie <= 0.02 * n_check * n_steps
    # the benchmark's path: C3 = one phased launch per step (step loop, then reset loop), C5's shard = step launch + reset launch
    assert envs[0].launch_count - launches0 == (1 if N == 10 else 2) * n_steps


@pytest.mark.parametrize("variant,n,r", [("tdcpa_v2", 10, 36), ("tdcpa_v2", 7, 24), ("att", 3, 18), ("multimap", 3, 18)])
def test_lean_kernel_equals_full_kernel(variant, n, r):
    """With no optional output requested the library runs an instantiation with that code compiled out; it must
    produce exactly the core outputs and state of the full kernel (which the oracle tests exercise)."""
    import torch
    from multi_agent_aac_b200 import _capi as K
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import multimap_set, synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    maps = multimap_set(seed=0)[:4] if variant == "multimap" else [synthetic_map(seed=0)]
    tabs = [OdTable(m, w_max=32) for m in maps]
    E = 500
    full_flags = K.OUT_RAW | K.OUT_RADAR_AUX | K.OUT_PARTS | (0 if variant == "multimap" else K.OUT_NBR6 | K.OUT_TCPA_PAIR)
    envs = []
    for flags in (0, full_flags):
        env = BatchedDroneEnv(preset(variant, n_envs=E, n_agents=n, n_rays=r, w_max=32, seed=6, out_flags=flags),
                              maps if variant == "multimap" else maps[0])
        env.set_od_tables(tabs)
        env.reset()
        envs.append(env)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(9)
    for t in range(20):
        act = (torch.rand((E, n, 2), device="cuda", generator=gen) * 2 - 1).contiguous()
        for env in envs:
            env.step(act, autoreset=True)
        for k in envs[0].out:
            assert torch.equal(envs[0].out[k].view(torch.uint8), envs[1].out[k].view(torch.uint8)), (t, k)
        for k in envs[0].state:
            assert torch.equal(envs[0].state[k].view(torch.uint8), envs[1].state[k].view(torch.uint8)), (t, k)
    assert envs[0].read_stats()[0] == envs[1].read_stats()[0] > 0


def test_lockstep_v2_eval_by_step():
    """forV2's evaluation "by sorties": terminal drones stay put, crash flags are live along the drone loop, crashes do not
    end the episode - clustered starts so that drone collisions, frozen neighbours and goal contacts all occur."""
    T = parity.lockstep(variant="v2", n_envs=96, n_agents=6, n_rays=36, steps=40, seed=21, cluster=9.0, eval_by_step=True)
    assert not T.fail, "\n".join(T.fail[:10])
    T = parity.lockstep(variant="v2", n_envs=64, n_agents=10, n_rays=18, steps=30, seed=22, eval_by_step=True)
    assert not T.fail, "\n".join(T.fail[:10])


@pytest.mark.parametrize("variant,n,r,flags", [("tdcpa_v2", 10, 36, 0), ("tdcpa_v2", 6, 24, "all"), ("multimap", 3, 18, "all")])
def test_radar_table_equals_direct_cast(variant, n, r, flags):
    """A freshly reset drone stands on a cell centre, so its radar is looked up in the per-(map, cell) table the library builds
    with its own observe kernel: the looked-up observation must equal, bit for bit, what the direct ray casting computes for
    the same state (observe), and a rollout with the table must equal one without it."""
    import torch
    from multi_agent_aac_b200 import _capi as K
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import multimap_set, synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    maps = multimap_set(seed=0)[:5] if variant == "multimap" else [synthetic_map(seed=0)]
    tabs = [OdTable(m, w_max=32) for m in maps]
    of = 0 if flags == 0 else (K.OUT_RAW | K.OUT_RADAR_AUX | K.OUT_PARTS)
    envs = []
    for use_table in (True, False):
        env = BatchedDroneEnv(preset(variant, n_envs=700, n_agents=n, n_rays=r, w_max=32, seed=8, out_flags=of), maps if variant == "multimap" else maps[0])
        env.set_od_tables(tabs)
        if not use_table:
            env.build_radar_table(enable=False)
        envs.append(env)
    with_tab, without = envs
    assert with_tab._rtab is not None and without._rtab is None
    with_tab.reset()
    looked_up = {k: with_tab.out[k].clone() for k in ("radar", "radar_min", "radar_hit") if k in with_tab.out}
    with_tab.observe()                                    # job 0 of the observe mode always casts the rays
    for k, v in looked_up.items():
        assert torch.equal(v.view(torch.uint8), with_tab.out[k].view(torch.uint8)), k
    without.reset()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    for t in range(25):
        act = (torch.rand((700, n, 2), device="cuda", generator=gen) * 2 - 1).contiguous()
        for env in envs:
            env.step(act, autoreset=True)
        for k in with_tab.out:
            assert torch.equal(with_tab.out[k].view(torch.uint8), without.out[k].view(torch.uint8)), (t, k)
    assert with_tab.read_stats()[0] == without.read_stats()[0] > 0


def test_every_instantiation_runs():
    """tests/tools/sanitize.py: a short rollout through every kernel instantiation and entry point (no launch error, no
    barrier-wait trap, finite outputs)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "tools", "sanitize.py")], cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-1500:]
    assert out.stdout.strip().endswith("done")


def test_phased_launches_on_concurrent_streams():
    """Two handles stepping at the same time on their own streams (phased launches: every warp of either kernel may wait on a
    per-group flag that a warp of the SAME kernel publishes; the kernels share the SMs) must each leave what a handle stepping
    alone leaves, bit for bit."""
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    gmap = synthetic_map(seed=0)
    tab = OdTable(gmap, w_max=32, planner="device")
    E, n, r, steps = 30000, 10, 36, 12
    streams = [None, torch.cuda.Stream(), torch.cuda.Stream()]
    envs = []
    for st in streams:
        env = BatchedDroneEnv(preset("tdcpa_v2", n_envs=E, n_agents=n, n_rays=r, w_max=32, seed=13, autoreset_launches=3), gmap, stream=st)
        env.set_od_tables([tab])
        env.reset()
        envs.append(env)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    acts = [(torch.rand((E, n, 2), device="cuda", generator=gen) * 2 - 1).contiguous() for _ in range(steps)]
    torch.cuda.synchronize()
    for t in range(steps):          # the reference run, alone on the device
        envs[0].step(acts[t], autoreset=True)
    torch.cuda.synchronize()
    l0 = [e.launch_count for e in envs]
    for t in range(steps):          # the two others interleaved, nothing between them but the launches
        envs[1].step(acts[t], autoreset=True)
        envs[2].step(acts[t], autoreset=True)
    torch.cuda.synchronize()
    assert envs[1].launch_count - l0[1] == steps       # one phased launch per step
    for other in envs[1:]:
        for k in envs[0].out:
            assert torch.equal(envs[0].out[k].view(torch.uint8), other.out[k].view(torch.uint8)), k
        for k in envs[0].state:
            assert torch.equal(envs[0].state[k].view(torch.uint8), other.state[k].view(torch.uint8)), k
    assert envs[0].read_stats()[0] == envs[1].read_stats()[0] == envs[2].read_stats()[0] > 0


@pytest.mark.parametrize("variant,n,r,E", [("tdcpa_v2", 10, 36, 3001), ("tdcpa_v2", 20, 72, 700), ("tdcpa_v2", 10, 36, 40000),
                                           ("att", 3, 36, 9001), ("att", 3, 18, 2000), ("multimap", 3, 18, 9001)])
def test_mode_specialised_launches_equal_fused_launch(variant, n, r, E):
    """The benchmark shapes without optional outputs run their two-launch auto-reset through kernels specialised on the
    mode (step without reset code, reset without reward code): state, outputs and counters must equal the fused launch's
    and the generic-mode path's (step, then autoreset through the runtime-mode kernel of a non-lean handle is covered by
    test_fused_autoreset_equals_step_then_autoreset), bit for bit."""
    import numpy as np
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import multimap_set, synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    maps = multimap_set(seed=0)[:4] if variant == "multimap" else [synthetic_map(seed=0)]
    tabs = [OdTable(m, w_max=32, planner="device") for m in maps]
    envs = []
    for launches in (1, 2, 3):   # 3: the phased launch (step loop, then reset loop behind per-group completion flags)
        env = BatchedDroneEnv(preset(variant, n_envs=E, n_agents=n, n_rays=r, w_max=32, seed=9, autoreset_launches=launches),
                              maps if variant == "multimap" else maps[0])
        env.set_od_tables(tabs)
        env.reset()          # MODE_RESET through the specialised kernel on both handles
        envs.append(env)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    n_term = 0
    for t in range(20):
        act = (torch.rand((E, n, 2), device="cuda", generator=gen) * 2 - 1).contiguous()
        for env in envs:
            env.step(act, autoreset=True)
        n_term += int((envs[0].out["terminated"] != 0).sum())
        for other in envs[1:]:
            for k in envs[0].out:
                assert torch.equal(envs[0].out[k].view(torch.uint8), other.out[k].view(torch.uint8)), (t, k)
            for k in envs[0].state:
                assert torch.equal(envs[0].state[k].view(torch.uint8), other.state[k].view(torch.uint8)), (t, k)
    assert n_term > 0
    assert envs[1].launch_count - envs[0].launch_count == 20 and envs[2].launch_count == envs[0].launch_count
    s0 = envs[0].read_stats()
    for other in envs[1:]:
        s1 = other.read_stats()
        assert np.array_equal(s0[[0, 1, 3, 4, 5, 6, 7, 8, 9]], s1[[0, 1, 3, 4, 5, 6, 7, 8, 9]]) and s0[0] == n_term
