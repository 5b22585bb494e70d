"""GPU parity of the batched actor (include/aac_actor.h) against the float64 oracle and the reference fixtures.

Tolerances (stated per the floating-point rule): the kernel rounds weights and every layer's input to bf16
(8-bit mantissa, relative 2^-9 per operand) and accumulates in fp32, so against float64
  * vs the float64 oracle / the reference's own outputs (tests/golden/actor_v2*.npz): |d action| <= 2e-2
    (tanh output in [-1, 1]; measured ~3e-3);
  * vs a float64 evaluation that applies the SAME bf16 roundings (weights, layer inputs): <= 1e-3 on actions (typically 1e-6; an activation that sits on a bf16 rounding boundary flips by 2^-9 relative) and
    5e-3 relative-to-scale on hidden layers (what is left is fp32 accumulation order and bf16 re-rounding flips).
"""
import os

import numpy as np
import pytest

from oracle import actor_oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def bf16(x):
    """Round-to-nearest-even to bfloat16, returned as float64."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def forward_bf16_points(sd, own, nbr, grid):
    """float64 arithmetic with the kernel's rounding points: bf16 weights, bf16 layer inputs, fp32-exact biases."""
    w = {k: (bf16(v) if k.endswith(".weight") and not k.startswith("act_out.2") else np.asarray(v, np.float64)) for k, v in sd.items()}
    leaky = lambda x: np.where(x > 0, x, 0.01 * x)
    lin = lambda name, x: x @ w[name + ".weight"].T + w[name + ".bias"]
    h1 = np.concatenate([leaky(lin("own_fc.0", bf16(own))), leaky(lin("own_full_nei.0", bf16(nbr))), leaky(lin("own_grid.0", bf16(grid)))], axis=1)
    h2 = leaky(lin("merge_feature.0", bf16(h1)))
    h3 = leaky(lin("act_out.0", bf16(h2)))
    return np.tanh(lin("act_out.2", h3)), h1, h2, h3


def make_actor(d_own, d_nbr, d_grid, rows, seed):
    from multi_agent_aac_b200.actor import BatchedActor
    sd = actor_oracle.reference_like_params(d_own, d_nbr, d_grid, seed)
    actor = BatchedActor(d_own, d_nbr, d_grid, rows)
    actor.load_state_dict(sd)
    return actor, sd


def to_dev(*arrs):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda() for a in arrs]


@pytest.mark.parametrize("name", ["actor_v2", "actor_v2_r18_n4"])
def test_actor_matches_reference_fixture(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    d_own, d_nbr, d_grid, seed = (int(v) for v in d["dims"])
    actor, sd = make_actor(d_own, d_nbr, d_grid, len(d["own"]), seed)
    own, nbr, grid = to_dev(d["own"], d["nbr"], d["grid"])
    act = actor.forward(own, nbr, grid).cpu().numpy().astype(np.float64)
    assert np.abs(act - d["act"]).max() <= 2e-2, np.abs(act - d["act"]).max()
    same_points = forward_bf16_points(sd, own.cpu().numpy(), nbr.cpu().numpy(), grid.cpu().numpy())[0]
    assert np.abs(act - same_points).max() <= 1e-3, np.abs(act - same_points).max()
    assert actor.launch_count == 1


@pytest.mark.parametrize("rows", [1, 127, 128, 129, 148 * 128 + 5, 40000])
def test_actor_hidden_layers_and_ragged_tiles(rows):
    """Every layer against the float64 oracle evaluated at the kernel's rounding points; row counts around the tile
    size, more tiles than SMs (persistent loop, ring phases wrapping), single row."""
    rng = np.random.default_rng(rows)
    actor, sd = make_actor(7, 45, 36, rows, 3)
    own = rng.uniform(-1, 1, (rows, 7)); own[:, -1] = rng.uniform(-np.pi, np.pi, rows)
    nbr = rng.uniform(-1, 1, (rows, 45))
    grid = np.where(rng.uniform(size=(rows, 36)) < 0.5, 15.0, rng.uniform(0, 15, (rows, 36)))
    t_own, t_nbr, t_grid = to_dev(own, nbr, grid)
    ref = forward_bf16_points(sd, t_own.cpu().numpy(), t_nbr.cpu().numpy(), t_grid.cpu().numpy())
    for layer in (1, 2, 3):
        h = actor.hidden(layer, t_own, t_nbr, t_grid).cpu().numpy().astype(np.float64)
        scale = np.abs(ref[layer]).max()
        assert np.abs(h - ref[layer]).max() <= 5e-3 * scale, (layer, np.abs(h - ref[layer]).max(), scale)
    act = actor.forward(t_own, t_nbr, t_grid).cpu().numpy().astype(np.float64)
    assert np.abs(act - ref[0]).max() <= 1e-3, np.abs(act - ref[0]).max()
    assert np.abs(act - actor_oracle.forward(sd, own, nbr, grid)).max() <= 2e-2


def test_actor_wide_inputs_and_noise():
    """N = 20 drones / 72 rays (C5 shapes: padded K pieces of 64 + 32 and 64 + 16); exploration noise is repeatable,
    clamped, and has the requested scale."""
    import torch
    rows = 5000
    rng = np.random.default_rng(5)
    actor, sd = make_actor(7, 95, 72, rows, 4)
    own, nbr, grid = rng.uniform(-1, 1, (rows, 7)), rng.uniform(-1, 1, (rows, 95)), rng.uniform(0, 15, (rows, 72))
    t = to_dev(own, nbr, grid)
    act = actor.forward(*t)
    ref = forward_bf16_points(sd, *(x.cpu().numpy() for x in t))[0]
    assert np.abs(act.cpu().numpy() - ref).max() <= 1e-3
    n1, n2, n3 = actor.forward(*t, noise_scale=0.3, noise_seed=7), actor.forward(*t, noise_scale=0.3, noise_seed=7), actor.forward(*t, noise_scale=0.3, noise_seed=8)
    assert torch.equal(n1, n2) and not torch.equal(n1, n3)
    assert float(n1.abs().max()) <= 1.0
    z = ((n1 - act) / 0.3)[(n1.abs() < 1.0)]          # unclamped entries: standard normal draws
    assert abs(float(z.mean())) < 0.05 and abs(float(z.std()) - 1.0) < 0.08


def test_policy_rollout_through_env():
    """The actor consumes the env's observation tensors in place and its actions drive the next step."""
    import torch
    from multi_agent_aac_b200.actor import BatchedActor
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    gmap = synthetic_map(seed=0)
    env = BatchedDroneEnv(preset("tdcpa_v2", n_envs=300, n_agents=10, n_rays=36, w_max=32, seed=2), gmap)
    env.set_od_tables([OdTable(gmap, w_max=32)])
    actor = BatchedActor.for_env(env)
    sd = actor_oracle.reference_like_params(7, 45, 36, 9)
    actor.load_state_dict(sd)
    obs = env.reset()
    for t in range(5):
        act = actor(obs)
        ref = actor_oracle.forward(sd, *(obs[k].reshape(3000, -1).cpu().numpy() for k in ("norm_own", "norm_nbr", "radar")))
        assert act.shape == (300, 10, 2) and np.abs(act.reshape(3000, 2).cpu().numpy() - ref).max() <= 2e-2
        obs, reward, done, info = env.step(act, autoreset=True)
    assert torch.isfinite(reward).all()


def test_device_replay_ring_is_written_in_place():
    """Slots of the replay ring are the env's output buffers: a rollout through the ring equals the same rollout of a
    twin env without it, slot by slot, and sampled transitions chain obs -> next_obs across slots (incl. wrap-around)."""
    import torch
    from multi_agent_aac_b200.actor import BatchedActor
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.replay import DeviceReplay, OBS_KEYS
    from multi_agent_aac_b200.reset import OdTable
    gmap = synthetic_map(seed=0)
    tab = [OdTable(gmap, w_max=32)]
    envs = []
    for _ in range(2):
        env = BatchedDroneEnv(preset("tdcpa_v2", n_envs=200, n_agents=6, n_rays=36, w_max=32, seed=4), gmap)
        env.set_od_tables(tab)
        env.reset()
        envs.append(env)
    ring_env, twin = envs
    actor = BatchedActor.for_env(ring_env)
    actor.load_state_dict(actor_oracle.reference_like_params(7, 25, 36, 2))
    T = 6
    ring = DeviceReplay(ring_env, T)
    obs = ring.begin()
    twin_obs = twin.observe()
    log = []
    for t in range(9):                                   # wraps the 6-slot ring
        assert all(torch.equal(obs[k], twin_obs[k]) for k in OBS_KEYS)
        actor(ring.current_obs(), noise_scale=0.2, noise_seed=t, out=ring.action_slot())
        act = ring.action_slot().clone()
        obs, reward, done, info = ring.step()
        twin_obs, r2, d2, i2 = twin.step(act, autoreset=True)
        assert torch.equal(reward, r2) and torch.equal(done, d2) and torch.equal(info["terminated"], i2["terminated"])
        log.append((act, reward.clone(), {k: obs[k].clone() for k in OBS_KEYS}))
    assert ring.filled == T - 1
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    b = ring.sample(4096, generator=gen)
    age = (ring.head - b["step_slot"]) % T               # 1 = newest
    assert int(age.min()) >= 1 and int(age.max()) <= T - 1
    for a in range(1, T):
        sel = age == a
        act, reward, nobs = log[-a]
        e = b["env"][sel]
        assert torch.equal(b["act"][sel], act[e]) and torch.equal(b["reward"][sel], reward[e])
        assert all(torch.equal(b["next_" + k][sel], nobs[k][e]) for k in OBS_KEYS)


def test_replay_bootstrap_mask_covers_step_cap_endings():
    """An episode that ends at the step cap carries done = 0 (ATT/ma_main:448-462): the ring's `bootstrap` mask - not
    (1 - done) - must switch the bootstrap term off, because next_obs already belongs to the next episode."""
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.replay import DeviceReplay
    from multi_agent_aac_b200.reset import OdTable
    gmap = synthetic_map(seed=0)
    env = BatchedDroneEnv(preset("tdcpa_v2", n_envs=64, n_agents=3, n_rays=18, w_max=32, seed=2, episode_length=2), gmap)
    env.set_od_tables([OdTable(gmap, w_max=32)])
    env.reset()
    ring = DeviceReplay(env, 8)
    ring.begin()
    saw_cap = False
    for t in range(5):
        ring.action_slot().zero_()          # hovering drones neither crash nor arrive: the only ending is the step cap
        _, _, done, info = ring.step()
        cap = (info["terminated"] & 1).bool() & ~done.any(dim=1).bool()
        saw_cap |= bool(cap.any())
    assert saw_cap
    b = ring.sample(2048, generator=torch.Generator(device="cuda").manual_seed(1))
    ended = b["terminated"] != 0
    assert bool(ended.any()) and bool((b["bootstrap"][ended] == 0).all()) and bool((b["bootstrap"][~ended] == 1).all())
    capped = ended & ~b["done"].any(dim=1).bool()
    assert bool(capped.any())               # (1 - done) would have bootstrapped these from an unrelated episode


def test_actor_noise_is_fresh_at_every_call_unless_seeded():
    import torch
    from multi_agent_aac_b200.actor import BatchedActor
    actor = BatchedActor(7, 45, 36, 256)
    actor.load_state_dict(actor_oracle.reference_like_params(7, 45, 36, 3))
    g = torch.Generator(device="cuda").manual_seed(0)
    own, nbr, grid = (torch.rand((256, d), device="cuda", generator=g) for d in (7, 45, 36))
    clean = actor.forward(own, nbr, grid)
    a, b = actor.forward(own, nbr, grid, noise_scale=0.1), actor.forward(own, nbr, grid, noise_scale=0.1)
    assert not torch.equal(a, b)                                  # choose_action draws np.random.randn(2) per call (V2/maddpg_agent:1290-1294)
    assert 0.03 < float((a - clean).std()) < 0.3 and abs(float(((a - clean) * (b - clean)).mean())) < 2e-3   # independent draws
    c, d = actor.forward(own, nbr, grid, noise_scale=0.1, noise_seed=5), actor.forward(own, nbr, grid, noise_scale=0.1, noise_seed=5)
    assert torch.equal(c, d)


@pytest.mark.parametrize("name", ["actor_att", "actor_att_n5_r18"])
def test_att_actor_matches_reference_fixture(name):
    """ActorNetwork_ATT_TwoPortion in fp32 against the float64 outputs of the unmodified reference class: 2e-5 absolute
    (fp32 accumulation over <= 192 terms, folded attention algebra); rows with masked and all-masked neighbours included."""
    from multi_agent_aac_b200.actor import BatchedAttActor
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    d_own, d_grid, n_nei, seed = (int(v) for v in d["dims"])
    actor = BatchedAttActor(d_own, d_grid, n_nei)
    actor.load_state_dict(actor_oracle.reference_like_params_att(d_own, d_grid, seed))
    own, grid, nei = to_dev(d["own"], d["grid"], d["nei"])
    act = actor.forward(own, grid, nei).cpu().numpy().astype(np.float64)
    assert np.isfinite(act).all()
    assert np.abs(act - d["act"]).max() <= 2e-5, np.abs(act - d["act"]).max()
    assert actor.launch_count == 1


def test_att_actor_rollout_through_env():
    """The att-preset env's observation tensors feed the attention actor in place; ragged last block; oracle agreement."""
    import torch
    from multi_agent_aac_b200.actor import BatchedAttActor
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    gmap = synthetic_map(seed=0)
    env = BatchedDroneEnv(preset("att", n_envs=1111, n_agents=3, n_rays=36, w_max=32, seed=3), gmap)
    env.set_od_tables([OdTable(gmap, w_max=32)])
    actor = BatchedAttActor.for_env(env)
    sd = actor_oracle.reference_like_params_att(14, 36, 5)
    actor.load_state_dict(sd)
    obs = env.reset()
    for t in range(4):
        act = actor(obs)
        ref = actor_oracle.forward_att(sd, obs["norm_own"].reshape(3333, 14).cpu().numpy(), obs["radar"].reshape(3333, 36).cpu().numpy(),
                                       obs["norm_nbr6"].reshape(3333, 2, 6).cpu().numpy())
        assert act.shape == (1111, 3, 2) and np.abs(act.reshape(3333, 2).cpu().numpy() - ref).max() <= 2e-5
        obs, reward, done, info = env.step(actor(obs, noise_scale=0.3, noise_seed=t), autoreset=True)
    assert torch.isfinite(reward).all()


@pytest.mark.gpu
def test_example_driver_runs():
    """examples/rollout_v2.py: the reference's loop skeleton (choose_action -> step -> replay -> sample) on the batched path."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "rollout_v2.py"), "--envs", "2048", "--steps", "40"],
                         cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-1500:]
    assert "agent-steps/s" in out.stdout and "episodes" in out.stdout

