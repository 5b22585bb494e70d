"""What the per-episode path search costs: the C3 shape (65 536 envs x 10 drones x 36 rays, random actions, ~45 % of the envs
re-initialised per step) with the all-pairs table of paths and with a pools-only table (every reference line searched on the
device when its episode starts).  Prints ms per step for both."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    gmap = synthetic_map(seed=0)
    E, n, r, steps = 65536, 10, 36, 100
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    acts = [(torch.rand((E, n, 2), device="cuda", generator=gen) * 2 - 1).contiguous() for _ in range(8)]
    for paths, launches in ((True, 0), (True, 1), (False, 0)):
        env = BatchedDroneEnv(preset("tdcpa_v2", n_envs=E, n_agents=n, n_rays=r, w_max=32, seed=1, autoreset_launches=launches), gmap)
        env.set_od_tables([OdTable(gmap, w_max=32, planner="device", paths=paths)])
        env.reset()
        for t in range(10):
            env.step(acts[t % 8], autoreset=True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for t in range(steps):
            env.step(acts[t % 8], autoreset=True)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / steps
        s = env.read_stats()
        print("paths=%s launches=%d: %.4f ms per step, %.3g agent-steps/s, episodes %d, fallbacks %d" % (paths, launches, ms, E * n / ms * 1e3, s[0], s[10]), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
