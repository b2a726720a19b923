"""The reference's golden 5,000-iteration phase-1 run (tests/golden/p10_traj.npz, BASELINE.md's GC hash) on the GPU:
how far does the free-running trajectory stay on the reference's check log, and how long does the run take."""
import hashlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vae_connexe_b200 as V

traj = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "p10_traj.npz"))
Xt = torch.from_numpy(traj["data"].T.copy())[None].cuda()
torch.manual_seed(0); np.random.seed(0)
m = V.CRVAE(10, np.ones((10, 10)), 64)
log = []
torch.cuda.synchronize(); t0 = time.time()
V.train_phase1(m, Xt, context=20, lam=0.1, lam_ridge=0, lr=5e-2, max_iter=int(traj["max_iter"]), check_every=50, verbose=0, log=log)
torch.cuda.synchronize(); dt = time.time() - t0
usage = np.array([r["usage"] for r in log]); gold = traj["log_usage"][: len(usage)]
loss = np.array([r["mean_loss"] for r in log]); gl = traj["log_loss"][: len(loss)]
same = usage == gold
first = int(np.argmin(same)) if not same.all() else -1
gc = m.GC().cpu().numpy()
sha = hashlib.sha256(np.ascontiguousarray(gc.astype(np.int32)).tobytes()).hexdigest()
print(f"wall {dt:.2f} s for {int(traj['max_iter'])} iterations + {len(log)} check blocks")
print(f"check-log usage equal at {int(same.sum())}/{len(same)} checks; first difference at check {first} (it {int(traj['log_it'][first]) if first >= 0 else -1})")
print(f"max rel loss deviation over the checks that precede the first usage difference: {float(np.max(np.abs(loss[:first if first > 0 else len(loss)] - gl[:first if first > 0 else len(gl)]) / gl[:first if first > 0 else len(gl)])):.2e}")
print(f"best_it {m.best_it} (golden {int(traj['best_it'])}); final GC == golden: {np.array_equal(gc.astype(np.int8), traj['final_GC'].astype(np.int8))}; differing entries {int((gc.astype(np.int8) != traj['final_GC'].astype(np.int8)).sum())}/100")
print(f"sha256(int32 GC) {sha[:16]}  golden {str(traj['final_GC_sha256'])[:16]}")
