"""Multi-rank parity script, run under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/gpu_dist_parity.py
Head-sharded train_phase1 / train_phase2 must reproduce the single-GPU run: identical check log, GC equal on every
rank, weights within 1e-4.  Two configurations:
  * one rank per GPU over NCCL (default; the production configuration), CUDA graphs on;
  * DIST_ONE_GPU=1: every rank on cuda:0 over gloo (NCCL refuses two ranks on one device), CUDA graphs off for the
    sharded model -- this is how tests/test_gpu_dist.py exercises the sharded path inside `pytest -m gpu` on a
    single-GPU box (the kernels, the shard arithmetic and the collectives' placement are the same; only the transport
    differs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vae_connexe_b200 as V  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    one_gpu = os.environ.get("DIST_ONE_GPU") == "1"
    if one_gpu:
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    quick = os.environ.get("DIST_QUICK") == "1"
    traj = np.load(os.path.join(ROOT, "tests", "golden", "p10_traj.npz"))
    ph2 = np.load(os.path.join(ROOT, "tests", "golden", "p10_phase2.npz"))
    Xt = torch.from_numpy(traj["data"].T.copy())[None].cuda()
    ok = True
    closers = []

    def run(sharded, phase):
        torch.manual_seed(0); np.random.seed(0)
        kw = dict(rank=rank, world_size=world, group=dist.group.WORLD) if sharded else {}
        log = []
        if phase == 1:
            m = V.CRVAE(10, np.ones((10, 10)), 64, **kw)
            closers.append(m)
            V.train_phase1(m, Xt, context=20, lam=0.1, lam_ridge=0.01, lr=5e-2, max_iter=61, check_every=20, verbose=0, log=log,
                           use_graphs=not (one_gpu and sharded))
            return m, None, log
        m = V.CRVAE(10, ph2["connection"], 64, **kw)
        closers.append(m)
        v = V.VRAE4E(10, 64)
        V.train_phase2(m, v, Xt, context=20, lam=0., lam_ridge=0, lr=5e-2, max_iter=21, check_every=10, verbose=0, log=log,
                       use_graphs=not (one_gpu and sharded))
        return m, v, log

    for phase in (1, 2):
        ms, vs, logs = run(True, phase)
        m1, v1, log1 = run(False, phase)          # every rank also runs the unsharded model on its own GPU
        for a, b in zip(logs, log1):
            for key in a:
                if key == "gc":
                    if a[key] is not None and not np.array_equal(a[key], b[key]):
                        ok = False; print(f"[rank {rank}] phase {phase} GC mismatch at check {a['it']}")
                elif a[key] is not None and abs(a[key] - b[key]) > 1e-4 * abs(b[key]) + 1e-6:
                    ok = False; print(f"[rank {rank}] phase {phase} log mismatch", key, a, b)
        if phase == 1 and not torch.equal(ms.GC(), m1.GC()):
            ok = False; print(f"[rank {rank}] GC mismatch")
        if phase == 1:      # test-mode generation on the head shard (per-step all-gather of the heads' outputs) == single GPU
            Xg = torch.zeros(32, 20, 10, device="cuda")
            torch.manual_seed(5); ga = ms(Xg, mode="test")
            torch.manual_seed(5); gb = m1(Xg, mode="test")
            relg = float((ga - gb).abs().max() / gb.abs().max())
            if ga.shape != (32, 21, 10) or relg > 1e-5:
                ok = False; print(f"[rank {rank}] sharded generation mismatch {relg:.2e}")
        sd1 = m1.state_dict()
        for k, t in ms.state_dict().items():
            rel = float((t - sd1[k]).abs().max() / sd1[k].abs().max().clamp_min(1e-30))
            if rel > 1e-4:
                ok = False; print(f"[rank {rank}] phase {phase} weight mismatch {k}: {rel:.2e}")
        if vs is not None:
            rel = float((vs.engine.theta.flat - v1.engine.theta.flat).abs().max() / v1.engine.theta.flat.abs().max())
            if rel > 1e-4:
                ok = False; print(f"[rank {rank}] VRAE weights mismatch {rel:.2e}")
    # phase 1 again at a size where every shard runs the tcgen05 recurrent / projection kernels (>= 8 heads per rank up to
    # world 8): p = 64 synthetic Lorenz-96, 41 iterations, sharded vs unsharded
    from vae_connexe_b200.data import lorenz_96
    p_tc = 64
    Xtc = torch.tensor(lorenz_96(d=p_tc, t=600, t_eval=0, f=10.0, seed=1).T.copy())[None].cuda()
    res = []
    for sharded in (True, False):
        torch.manual_seed(0); np.random.seed(0)
        kw = dict(rank=rank, world_size=world, group=dist.group.WORLD) if sharded else {}
        m = V.CRVAE(p_tc, np.ones((p_tc, p_tc)), 64, **kw)
        closers.append(m)
        log = []
        V.train_phase1(m, Xtc, context=20, lam=0.1, lam_ridge=0.0, lr=5e-2, max_iter=41, check_every=20, verbose=0, log=log,
                       use_graphs=not (one_gpu and sharded))
        res.append((m, log))
    (ms, logs), (m1, log1) = res
    if rank == 0:
        print(f"shard kernels at p={p_tc}, world={world}: recurrence {ms.engine.rec_mode}, projection {ms.engine.proj_mode}")
    for a, b in zip(logs, log1):
        for key in a:
            if key == "gc":
                if a[key] is not None and not np.array_equal(a[key], b[key]):
                    ok = False; print(f"[rank {rank}] tc-size GC mismatch at check {a['it']}")
            elif a[key] is not None and abs(a[key] - b[key]) > 1e-4 * abs(b[key]) + 1e-6:
                ok = False; print(f"[rank {rank}] tc-size log mismatch", key, a, b)
    if not torch.equal(ms.GC(), m1.GC()):
        ok = False; print(f"[rank {rank}] tc-size GC mismatch")
    sd1 = m1.state_dict()
    for k, t in ms.state_dict().items():
        diff = float((t - sd1[k]).abs().max())
        rel = diff / float(sd1[k].abs().max().clamp_min(1e-30))
        if rel > 1e-4 and diff > 5e-6:        # (1,)-shaped biases near zero: judge those by the absolute difference
            ok = False; print(f"[rank {rank}] tc-size weight mismatch {k}: rel {rel:.2e} abs {diff:.2e}")
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_PARITY", "PASS" if flag.item() == 1.0 else "FAIL", f"world={world}", flush=True)
    # orderly teardown: captured graphs that contain collectives (the sharded generator) and the peer-memory communicator go
    # first, then the process group
    import gc
    for mdl in closers:
        mdl.close()
    del closers[:]
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
