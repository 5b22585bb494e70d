"""Tuning aid: times the actor kernel (CUDA events) and, with AAC_ACTOR_PROF=1, prints its per-tile phase clocks."""
import sys, os, ctypes, torch, numpy as np
sys.path.insert(0, ".")
from multi_agent_aac_b200.actor import BatchedActor
from multi_agent_aac_b200 import _actor_capi as K
from oracle import actor_oracle
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 655360
sd = actor_oracle.reference_like_params(7, 45, 36, 0)
actor = BatchedActor(7, 45, 36, rows); actor.load_state_dict(sd)
g = torch.Generator(device="cuda"); g.manual_seed(0)
own = torch.rand((rows, 7), device="cuda", generator=g) * 2 - 1
nbr = torch.rand((rows, 45), device="cuda", generator=g) * 2 - 1
grid = torch.rand((rows, 36), device="cuda", generator=g) * 15
out = torch.empty((rows, 2), device="cuda")
for _ in range(5): actor.forward(own, nbr, grid, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
Kn = 50
e0.record()
for _ in range(Kn): actor.forward(own, nbr, grid, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / Kn
flop = 2.0 * rows * (7*128 + 45*128 + 36*128 + 384*512 + 512*256 + 256*2)
print("actor kernel: %.4f ms  %.3e rows/s  %.1f TFLOP/s (algorithmic)" % (ms, rows / ms * 1e3, flop / ms / 1e9))
ref = actor_oracle.forward(sd, own[:512].cpu().numpy(), nbr[:512].cpu().numpy(), grid[:512].cpu().numpy())
print("max |action - float64 oracle| on 512 rows: %.2e" % np.abs(out[:512].cpu().numpy() - ref).max())
if os.environ.get("AAC_ACTOR_PROF"):
    buf = (ctypes.c_longlong * (256 * 8 + 128))()
    L = K.lib(); L.aac_actor_prof.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    n = L.aac_actor_prof(actor._h, buf)
    a = np.array(buf[: n * 8], dtype=np.float64).reshape(n, 8)
    tiles = (rows + 127) // 128 / n
    print("per-tile cycles (mean over CTAs): stage %.0f | wait L1 %.0f | E1 %.0f | wait L2 %.0f | E2 %.0f | wait L3 %.0f | E3 %.0f | total %.0f" % (*(a[:, :7].mean(0) / tiles), a[:, :7].sum(1).mean() / tiles))
    tl = np.array(buf[n * 8: n * 8 + 128], dtype=np.int64)
    names = ["stage end", "L1 done", "E1 end", "L2 done", "E2 end", "L3 done", "E3 end"]
    print("CTA 0, last tile, cycles after the end of staging:", ", ".join("%s %d" % (names[i], tl[i] - tl[0]) for i in range(7)))
    print("MMA thread issues chunk c at:", " ".join("%d:%d" % (c, tl[64 + c] - tl[0]) for c in range(32) if tl[64 + c]))
