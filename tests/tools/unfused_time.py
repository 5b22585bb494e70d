"""Tuning aid: aac_step_fused (one launch) against aac_step_autoreset (step launch + reset launch) on bench workloads."""
import sys, torch
sys.path.insert(0, ".")
import bench
from multi_agent_aac_b200.env import BatchedDroneEnv, preset
from multi_agent_aac_b200.reset import OdTable
dev = torch.device("cuda", 0)
for wl in sys.argv[1:] or ["c2", "c4", "c3", "c5"]:
    preset_name, envs, n, r, desc = bench.WORKLOADS[wl]
    gmap, _ = bench.build_world(wl, 64, seed=1000)
    env = BatchedDroneEnv(preset(preset_name, n_envs=envs, n_agents=n, n_rays=r, w_max=32, seed=1000), gmap, device=dev)
    env.set_od_tables([OdTable(m, w_max=32, planner="device") for m in (gmap if isinstance(gmap, list) else [gmap])])
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    acts = [(torch.rand((envs, n, 2), device=dev, generator=gen) * 2 - 1).contiguous() for _ in range(8)]
    def run(fused, steps=400):
        for k in range(30):
            env.step(acts[k % 8], autoreset=True, fused=fused)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize(); e[0].record()
        for k in range(steps):
            env.step(acts[k % 8], autoreset=True, fused=fused)
        e[1].record(); torch.cuda.synchronize()
        return e[0].elapsed_time(e[1]) / steps
    for rep in range(2):
        print("%s: fused %.4f ms   step + reset launches %.4f ms" % (wl, run(True), run(False)))
    env.close()
