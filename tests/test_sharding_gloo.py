"""Head sharding on CPU: world_size-2 gloo, checker backend.  The sharded run
must reproduce the single-process run: same losses, same weights per head, same GC on every rank."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import GOLDEN, ROOT


def test_head_range_partitions():
    from vae_connexe_b200.sharding import head_range
    for p in (1, 7, 10, 100, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [head_range(p, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == p
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        head_range(10, 2, 2)


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import vae_connexe_b200 as V
    import vae_connexe_b200.lib as L
    from tests.cpu_backend import OracleKernels
    L.set_test_backend(OracleKernels())
    traj = np.load(os.path.join(GOLDEN, "p10_traj.npz"))
    Xt = torch.from_numpy(traj["data"].T.copy())[None]
    torch.manual_seed(0); np.random.seed(0)
    m = V.CRVAE(10, np.ones((10, 10)), 64, rank=rank, world_size=world, group=dist.group.WORLD)
    log = []
    V.train_phase1(m, Xt, context=20, lam=0.1, lam_ridge=0.01, lr=5e-2, max_iter=12, check_every=5, verbose=0, log=log)
    gc = m.GC()
    sd = {k: v.numpy().copy() for k, v in m.state_dict().items()}   # numpy: no shared-memory handles in the queue
    q.put((rank, log, gc.numpy(), sd, (m.head_lo, m.head_hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_training_equals_single_process(world, cpu_backend):
    import vae_connexe_b200 as V
    traj = np.load(os.path.join(GOLDEN, "p10_traj.npz"))
    Xt = torch.from_numpy(traj["data"].T.copy())[None]
    torch.manual_seed(0); np.random.seed(0)
    m = V.CRVAE(10, np.ones((10, 10)), 64)
    log = []
    V.train_phase1(m, Xt, context=20, lam=0.1, lam_ridge=0.01, lr=5e-2, max_iter=12, check_every=5, verbose=0, log=log)
    ref_sd, ref_gc = m.state_dict(), m.GC().numpy()

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    results = [q.get(timeout=300) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    covered = set()
    for rank, rlog, gc, sd, (lo, hi) in results:
        assert np.array_equal(gc, ref_gc)                                  # all-gathered GC identical on every rank
        assert [r["it"] for r in rlog] == [r["it"] for r in log]
        for a, b in zip(rlog, log):
            assert abs(a["mean_loss"] - b["mean_loss"]) < 1e-5 and abs(a["kl"] - b["kl"]) < 1e-6
        for k, v in sd.items():
            assert np.allclose(v, ref_sd[k].numpy(), rtol=1e-4, atol=1e-6), (rank, k)    # shard heads + replicated encoder
        covered |= set(range(lo, hi))
    assert covered == set(range(10))
