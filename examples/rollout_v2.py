"""Thin example driver: the reference's training-loop skeleton on the batched path (`ma_main_*.py` is the
caller of the hot path and stays host Python; this is the few lines of it that touch the path).

The reference's loop, one env in one thread (V2/ma_main:373-382, :526, :578-637):

    for episode ...:                                   here, for E envs at once:
        cur_state, norm_cur_state = env.reset_world()      ring.begin()                  (aac_reset + observe)
        while True:
            action = model.choose_action(norm_cur_state)   actor(obs, out=ring.action_slot())   (one launch, all E * N drones)
            next_state, ... = env.step(action)             ring.step()                   (step + ss_reward_Mar + episode rule,
            reward, done, ... = env.ss_reward_Mar(...)                                    finished envs start their next episode)
            model.memory.push(...)                         - nothing: the ring slot IS the env's output buffer
            model.update_myown(...)                        batch = ring.sample(B)        (joint transitions, `bootstrap` mask)
        every 100 episodes: print collision / goal counters     env.read_stats()         (the same counters, summed on device)

The learner itself (critics, optimiser) is out of scope: `update` below only shows where it plugs in.  Parameters are random
(`actor_params.reference_like_params`: the reference architecture, no checkpoint in this tree); pass the reference module's
`state_dict()` to `BatchedActor.load_state_dict` to run a trained policy.

    python examples/rollout_v2.py --envs 4096 --steps 200
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def update(batch):
    """Where `MADDPG.update_myown` goes (V2/maddpg_agent:502-...): the batch is on the device, `bootstrap` masks the target
    critic's value where next_obs belongs to the following episode."""
    return float(batch["reward"].mean())


def main():
    import torch
    from multi_agent_aac_b200 import actor_params
    from multi_agent_aac_b200.actor import BatchedActor
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.replay import DeviceReplay
    from multi_agent_aac_b200.reset import OdTable
    from multi_agent_aac_b200.stats import reduce_episode_stats

    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--drones", type=int, default=10)
    ap.add_argument("--rays", type=int, default=36)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--ring", type=int, default=16, help="replay capacity in steps (each slot holds all envs)")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--update-every", type=int, default=10)
    ap.add_argument("--noise", type=float, default=0.1, help="exploration noise scale (`var` of V2/maddpg_agent:1290)")
    args = ap.parse_args()

    gmap = synthetic_map(seed=0)
    env = BatchedDroneEnv(preset("tdcpa_v2", n_envs=args.envs, n_agents=args.drones, n_rays=args.rays, w_max=32, seed=1), gmap)
    env.set_od_tables([OdTable(gmap, w_max=32, planner="device")])   # reset_world's origin / destination draw + path, on the device
    env.reset()
    actor = BatchedActor.for_env(env)
    actor.load_state_dict(actor_params.reference_like_params(env.D, 5 * (env.N - 1), env.R, seed=0))
    ring = DeviceReplay(env, args.ring)
    ring.begin()

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    last = None
    for k in range(args.steps):
        actor(ring.current_obs(), noise_scale=args.noise, out=ring.action_slot())
        ring.step()
        if (k + 1) % args.update_every == 0:
            last = update(ring.sample(args.batch))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rep = reduce_episode_stats(env.read_stats())
    print("%d steps x %d envs x %d drones in %.3f s = %.3g agent-steps/s (actor + step + auto-reset + replay, one GPU)"
          % (args.steps, env.E, env.N, dt, args.steps * env.E * env.N / dt))
    print("episodes %d, mean length %.2f, mean return %.2f, crash rate %.3f (bound %d / building %d / drone %d), all reached %d; "
          "replay ring %.1f MB, last batch mean reward %s"
          % (rep["episodes"], rep["mean_length"], rep["mean_return"], rep["crash_rate"], rep["bound_crash"], rep["building_crash"],
             rep["drone_crash"], rep["all_reached"], ring.bytes() / 1e6, "%.3f" % last if last is not None else "-"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
