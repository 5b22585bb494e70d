"""Exercise every kernel instantiation of both libraries on small batches (env step / reset / observe / auto-reset /
host-buffer step for the three variants with and without the optional outputs, specialised and generic shapes, ragged
tiles, evaluation by sorties; both actors; the device planner).  Written for a run under compute-sanitizer:

    compute-sanitizer --tool memcheck python tests/tools/sanitize.py [env|actor|plan ...]

(closed on the round-1 GPU pool, so there it only runs plain, as tests/test_gpu_parity.py::test_every_instantiation_runs:
launch errors, traps of the bounded barrier waits and non-finite outputs are what it catches.)  No oracle here."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from multi_agent_aac_b200 import _capi as K
from multi_agent_aac_b200 import actor_params
from multi_agent_aac_b200.actor import BatchedActor, BatchedAttActor
from multi_agent_aac_b200.env import BatchedDroneEnv, preset
from multi_agent_aac_b200.maps import multimap_set, synthetic_map
from multi_agent_aac_b200.reset import MultiMapBank, OdTable, ScenarioBank, plan_paths_device

ALL_OUT = K.OUT_RAW | K.OUT_NBR6 | K.OUT_TCPA_PAIR | K.OUT_RADAR_AUX | K.OUT_PARTS
what = set(sys.argv[1:]) or {"env", "actor", "plan"}
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(0)


def rollout(name, n, r, flags, source, steps=6, envs=96, **kw):
    maps = multimap_set(seed=0) if name == "multimap" else synthetic_map(bound=[0, 200, 0, 200] if name == "changeskin_sensors" else None, seed=0)
    if name == "multimap":
        flags &= ~(K.OUT_NBR6 | K.OUT_TCPA_PAIR)
    cfg = preset(name, n_envs=envs, n_agents=n, n_rays=r, w_max=32, seed=3, out_flags=flags, **kw)
    env = BatchedDroneEnv(cfg, maps, device=dev)
    if source == "bank":
        env.set_bank(MultiMapBank(maps, n, 64, w_max=32, seed=1) if name == "multimap" else ScenarioBank(maps, n, 64, w_max=32, seed=1))
    else:
        env.set_od_tables([OdTable(m, w_max=32, planner="device", paths=(source != "pools")) for m in (maps if isinstance(maps, list) else [maps])])
    env.reset()
    env.observe()
    for t in range(steps):
        act = (torch.rand((envs, n, 2), device=dev, generator=gen) * 2 - 1).contiguous()
        env.step(act, autoreset=(t % 3 != 2))
        if t % 3 == 2:
            env.autoreset()
    # the host-buffer entry point (chunks rotating over three streams)
    env.step_host((torch.rand((envs, n, 2)) * 2 - 1).contiguous().pin_memory(), autoreset=True)
    torch.cuda.synchronize(dev)
    stats = env.read_stats()
    env.close()
    print("ok env", name, n, r, hex(flags), source, kw, "episodes", stats[0])


if "env" in what:
    for flags in (0, ALL_OUT):
        rollout("att", 3, 18, flags, "od")
        rollout("tdcpa_v2", 10, 36, flags, "od")          # the specialised <V2, 10, 36> instantiation
        rollout("tdcpa_v2", 4, 24, flags, "bank", tile_envs=5, block_threads=96, envs=37)   # generic shapes, ragged tiles
        rollout("multimap", 3, 18, flags, "bank")
    rollout("tdcpa_v2", 20, 72, 0, "od", envs=40)
    rollout("tdcpa_v2", 5, 36, ALL_OUT, "od", eval_by_step=True)
    rollout("tdcpa_v2", 1, 18, 0, "bank", envs=9)
    rollout("tdcpa_v2", 10, 36, ALL_OUT, "pools", envs=130)      # pools-only tables: the per-episode search inside the reset
    rollout("multimap", 3, 18, 0, "pools", envs=200)
    rollout("att", 3, 18, 0, "pools", envs=70, tile_envs=4, block_threads=64)
    rollout("changeskin_sensors", 4, 18, K.OUT_RAW | K.OUT_RADAR_AUX | K.OUT_PARTS, "od", steps=12)    # the later fork's sensor classes
    rollout("changeskin_sensors", 6, 36, 0, "bank", radar_targets=15, n_nbr_obs=3)

if "actor" in what:
    for rows, dims in ((300, (7, 45, 36)), (128 * 150 + 17, (7, 45, 36)), (200, (7, 95, 72))):
        a = BatchedActor(*dims, rows, device=dev)
        a.load_state_dict(actor_params.reference_like_params(*dims, seed=0))
        own, nbr, grid = (torch.rand((rows, d), device=dev, generator=gen) for d in dims)
        out = a.forward(own, nbr, grid, noise_scale=0.1, noise_seed=1)
        for layer in (1, 2, 3):
            a.hidden(layer, own, nbr, grid)
        torch.cuda.synchronize(dev)
        assert bool(torch.isfinite(out).all())
        print("ok actor", rows, dims)
    att = BatchedAttActor(6, 36, 2, device=dev)
    att.load_state_dict(actor_params.reference_like_params_att(6, 36, seed=0))
    rows = 1000
    own, grid = torch.rand((rows, 6), device=dev, generator=gen), torch.rand((rows, 36), device=dev, generator=gen)
    nei = torch.rand((rows, 2, 6), device=dev, generator=gen)
    nei[::7] = 0.0        # masked neighbour rows
    out = att.forward(own, grid, nei, noise_scale=0.1, noise_seed=2)
    torch.cuda.synchronize(dev)
    assert bool(torch.isfinite(out).all())
    print("ok att actor", rows)

if "plan" in what:
    import types
    rng = np.random.default_rng(0)
    for shape in ((23, 13), (90, 70)):
        occ = (rng.random(shape) < 0.25).astype(np.uint8)
        free = np.argwhere(occ == 0)
        pairs = np.concatenate([free[rng.integers(0, len(free), 200)], free[rng.integers(0, len(free), 200)]], axis=1)
        cells, n = plan_paths_device(types.SimpleNamespace(occ=occ, gx=shape[0], gy=shape[1]), pairs, 16)
        print("ok plan", shape, np.bincount(np.clip(n, -1, 1) + 1))
print("done")
