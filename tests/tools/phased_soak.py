"""Soak of the phased launch (step loop + reset loop behind per-group release / acquire flags) against the two-launch path:
the same seeds and actions for `--steps` steps, state and outputs compared bit for bit every `--every` steps, over several
shapes / group sizes / CTA sizes (more groups than resident warps, fewer, ragged last group).

    python tests/tools/phased_soak.py --steps 3000
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--every", type=int, default=100)
    args = ap.parse_args()
    gmap = synthetic_map(seed=0)
    tab = OdTable(gmap, w_max=32, planner="device")
    cases = [dict(n_envs=65536, n_agents=10, n_rays=36), dict(n_envs=40001, n_agents=10, n_rays=36, tile_envs=1),
             dict(n_envs=20000, n_agents=10, n_rays=36, tile_envs=2, block_threads=128), dict(n_envs=9000, n_agents=20, n_rays=72)]
    for case in cases:
        envs = []
        for launches in (2, 3):
            env = BatchedDroneEnv(preset("tdcpa_v2", w_max=32, seed=17, autoreset_launches=launches, **case), gmap)
            env.set_od_tables([tab])
            env.reset()
            envs.append(env)
        gen = torch.Generator(device="cuda")
        gen.manual_seed(3)
        E, n = case["n_envs"], case["n_agents"]
        acts = [(torch.rand((E, n, 2), device="cuda", generator=gen) * 2 - 1).contiguous() for _ in range(8)]
        l0 = [e.launch_count for e in envs]
        for t in range(args.steps):
            for env in envs:
                env.step(acts[t % 8], autoreset=True)
            if (t + 1) % args.every == 0:
                for k in envs[0].out:
                    assert torch.equal(envs[0].out[k].view(torch.uint8), envs[1].out[k].view(torch.uint8)), (case, t, k)
                for k in envs[0].state:
                    assert torch.equal(envs[0].state[k].view(torch.uint8), envs[1].state[k].view(torch.uint8)), (case, t, k)
        assert envs[0].launch_count - l0[0] == 2 * args.steps and envs[1].launch_count - l0[1] == args.steps
        s = envs[1].read_stats()
        print("ok", case, "steps", args.steps, "episodes", int(s[0]), flush=True)
    print("done")
    return 0


if __name__ == "__main__":
    sys.exit(main())
