import torch, sys
from multi_agent_aac_b200 import _capi as K
from multi_agent_aac_b200.env import BatchedDroneEnv, preset
from multi_agent_aac_b200.maps import multimap_set, synthetic_map
from multi_agent_aac_b200.reset import OdTable
variant, n, r = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
maps = multimap_set(seed=0)[:4] if variant == "multimap" else [synthetic_map(seed=0)]
tabs = [OdTable(m, w_max=32) for m in maps]
E = 500
full_flags = K.OUT_RAW | K.OUT_RADAR_AUX | K.OUT_PARTS | (0 if variant == "multimap" else K.OUT_NBR6 | K.OUT_TCPA_PAIR)
envs = []
for flags in (0, full_flags):
    env = BatchedDroneEnv(preset(variant, n_envs=E, n_agents=n, n_rays=r, w_max=32, seed=6, out_flags=flags), maps if variant == "multimap" else maps[0])
    env.set_od_tables(tabs)
    env.reset()
    envs.append(env)
def cmp(tag):
    for name, d0, d1 in (("out", envs[0].out, envs[1].out), ("state", envs[0].state, envs[1].state)):
        for k in d0:
            a, b = d0[k], d1[k]
            if a is None or b is None: continue
            a = a.reshape(E, -1); b = b.reshape(E, -1)
            bad = ((a != b) & ~((a != a) & (b != b))).any(1).nonzero().flatten().tolist() if a.dtype.is_floating_point else (a != b).any(1).nonzero().flatten().tolist()
            if bad:
                print(tag, name, k, "envs differing:", bad[:20], "n", len(bad))
                e = bad[0]
                print("   lean", a[e].tolist()[:40]); print("   full", b[e].tolist()[:40])
cmp("after reset")
gen = torch.Generator(device="cuda"); gen.manual_seed(9)
for t in range(3):
    act = (torch.rand((E, n, 2), device="cuda", generator=gen) * 2 - 1).contiguous()
    for env in envs: env.step(act, autoreset=True)
    term = envs[0].out["terminated"].flatten().nonzero().flatten().tolist()
    print("t", t, "terminated envs", term[:40])
    cmp("t%d" % t)
