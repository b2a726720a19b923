"""Iteration time of ONE rank's head shard on one GPU (p = 100, rank 0 of WORLD, default 8), the other ranks' dz contribution
injected: the critical path of a head-sharded iteration without any waiting for peers.  CUDA events around graph replays."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vae_connexe_b200 as V
import bench


class InjectComm:
    def __init__(self, extra): self.extra = extra
    def allreduce_dz(self, dz_part): dz_part.add_(self.extra)
    def close(self): pass


world = int(os.environ.get("WORLD", "8"))
p, B = 100, 256
Xb = bench.make_batch(p, 1000, B).cuda()
torch.manual_seed(0)
extra = (torch.randn(B, 64) * 1e-3).cuda()
m = V.CRVAE(p, np.ones((p, p)), 64, rank=0, world_size=world, comm=InjectComm(extra))
run = V.Phase1Runner(m, Xb, 0.05, 0.1, 0.0, 0.1, use_graphs=True)
eps = torch.randn(16, B, 64, generator=torch.Generator().manual_seed(1234)).cuda()
run.forward(eps[0]); run.update(); run.forward(eps[1]); run.capture()
for i in range(10):
    run.iterate(eps[i % 16])
torch.cuda.synchronize()
for rep in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(500):
        run.iterate(eps[i % 16])
    e.record(); torch.cuda.synchronize()
    print(f"world {world}: heads {m.engine.P}, rec {m.engine.rec_mode}, {s.elapsed_time(e) / 500 * 1e3:.1f} us per iteration", flush=True)
