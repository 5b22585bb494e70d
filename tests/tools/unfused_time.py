"""Tuning aid: fused step+auto-reset launch against step launch followed by auto-reset launch (C3 shapes)."""
import sys, torch
sys.path.insert(0, ".")
from multi_agent_aac_b200.env import BatchedDroneEnv, preset
from multi_agent_aac_b200.maps import synthetic_map
from multi_agent_aac_b200.reset import OdTable
E_, N, R = 65536, 10, 36
gmap = synthetic_map(seed=0)
dev = torch.device("cuda", 0)
env = BatchedDroneEnv(preset("tdcpa_v2", n_envs=E_, n_agents=N, n_rays=R, w_max=32, seed=1000), gmap, device=dev)
env.set_od_tables([OdTable(gmap, w_max=32, planner="device")])
env.reset()
gen = torch.Generator(device=dev); gen.manual_seed(1)
acts = [(torch.rand((E_, N, 2), device=dev, generator=gen) * 2 - 1).contiguous() for _ in range(8)]
def run(fused, steps=600):
    for k in range(30):
        env.step(acts[k % 8], autoreset=True)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); e[0].record()
    for k in range(steps):
        if fused:
            env.step(acts[k % 8], autoreset=True)
        else:
            env.step(acts[k % 8], autoreset=False); env.autoreset()
    e[1].record(); torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]) / steps
for rep in range(2):
    print("fused %.4f ms   step + autoreset launches %.4f ms" % (run(True), run(False)))
