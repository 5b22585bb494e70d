"""Tuning aid: a few actor launches for `ncu -k regex:actor_kernel`."""
import sys, torch
sys.path.insert(0, ".")
from multi_agent_aac_b200.actor import BatchedActor
from oracle import actor_oracle
rows = 655360
actor = BatchedActor(7, 45, 36, rows); actor.load_state_dict(actor_oracle.reference_like_params(7, 45, 36, 0))
own = torch.rand((rows, 7), device="cuda") * 2 - 1
nbr = torch.rand((rows, 45), device="cuda") * 2 - 1
grid = torch.rand((rows, 36), device="cuda") * 15
out = torch.empty((rows, 2), device="cuda")
for _ in range(3): actor.forward(own, nbr, grid, out=out)
torch.cuda.synchronize()
