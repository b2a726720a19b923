"""Phase-2 iteration time (CRVAE on a pruned graph + VRAE4E error compensation, CRVAE_lorenz96.py:609-643) at p=100, B=256:
CUDA events around graph replays, the same way bench.py times the phase-1 iteration."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vae_connexe_b200 as V
from vae_connexe_b200.data import lorenz_96, lorenz_96_graph

p, B, H, steps = int(os.environ.get("P", 100)), 256, 64, 200
X = torch.tensor(lorenz_96(d=p, t=1000, t_eval=0, f=10.0, seed=0).T.copy())
wins = V.arrange_input(X, 20)[0]
Xb = wins[np.random.RandomState(0).randint(0, wins.shape[0], B)].cuda()
gc = lorenz_96_graph(p)                       # the true graph as the phase-2 structure: 4 inputs per head
torch.manual_seed(0)
c, v = V.CRVAE(p, gc, 64), V.VRAE4E(p, 64)
run = V.Phase2Runner(c, v, Xb, 5e-2, 0.0, 0.0)
gen = torch.Generator().manual_seed(1)
eps = [(torch.randn(B, H, generator=gen).cuda(), torch.randn(B, H, generator=gen).cuda()) for _ in range(8)]
run.forward(*eps[0]); run.update(); run.forward(*eps[1]); run.capture()
for i in range(10):
    run.iterate(*eps[i % 8])
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(steps):
    run.iterate(*eps[i % 8])
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / steps
print(f"phase-2 iteration p={p} B={B}: {ms:.3f} ms  ({B * 10 * p / (ms * 1e-3):.3e} timesteps*vars/s), loss {float(c.engine.loss):.5f}")
