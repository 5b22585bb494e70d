"""Tuning aid: times the attention actor kernel on the C2 shape (4096 envs x 3 drones, 36 rays)."""
import sys, torch
sys.path.insert(0, ".")
from multi_agent_aac_b200.actor import BatchedAttActor
from oracle import actor_oracle
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12288
actor = BatchedAttActor(14, 36, 2); actor.load_state_dict(actor_oracle.reference_like_params_att(14, 36, 0))
own = torch.rand((rows, 14), device="cuda") * 2 - 1; grid = torch.rand((rows, 36), device="cuda") * 15; nei = torch.rand((rows, 2, 6), device="cuda") * 2 - 1
out = torch.empty((rows, 2), device="cuda")
for _ in range(5): actor.forward(own, grid, nei, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(50): actor.forward(own, grid, nei, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print("att actor kernel: %.4f ms for %d drones = %.3e drones/s" % (ms, rows, rows / ms * 1e3))
