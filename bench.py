#!/usr/bin/env python
"""bench.py -- CR-VAE phase-1 training throughput (timesteps*vars/s) on Lorenz-96 p=100, T=1000.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one steady-state training iteration of train_phase1 (CRVAE_lorenz96.py:497-515:
backward, GD on all parameters, group-lasso prox, forward, loss) on the fixed batch of B=256
windows; one unit = one (batch row, decoder timestep, variable) prediction, B*10*p units per step.
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "crvae_phase1_train_timesteps_x_vars_per_s"
UNIT = "timesteps*vars/s"
TD = 10
H = 64
G = 192
L2_BYTES = 126e6
LR, LAM = 5e-2, 0.1


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sus=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start(); return self

    def __exit__(self, *a):
        self.stop.set(); self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_batch(p, T, B):
    """Synthetic Lorenz-96 (reference generator semantics, seed 0) -> the fixed training batch."""
    from vae_connexe_b200.data import lorenz_96
    X = lorenz_96(d=p, t=T, t_eval=0, f=10.0, seed=0)                   # (p, T)
    series = torch.from_numpy(np.ascontiguousarray(X.T))                # (T, p)
    n = T - 20
    idx_w = torch.arange(n)[:, None] + torch.arange(20)[None, :]
    wins = series[idx_w]                                                # arrange_input (:332-350)
    np.random.seed(0)
    idx = np.random.randint(n, size=(B,))                               # :470
    return wins[idx].contiguous()


def cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip(); break
    except Exception:
        pass
    return model, os.cpu_count() or 1


def workload_config(args, world):
    """The `config` object: derived from the command line only, so both arms print the identical dict."""
    p_total = args.p * world if args.scaling == "weak" else args.p
    heads = -(-p_total // world)
    ws = args.batch * TD * heads * (768 + 512)                          # gate buffer + h / gh_n, per GPU
    if ws > 1.5 * L2_BYTES:
        l2 = "per-GPU activation working set %.0f MB exceeds the 126 MB L2: timed steps run back to back, no flush" % (ws / 1e6)
    else:
        l2 = ("per-GPU activation working set %.0f MB fits the 126 MB L2: a 256 MB buffer is overwritten between timed "
              "steps (outside the per-step CUDA events) so every step starts from a cold L2" % (ws / 1e6))
    return {"workload": f"CRVAE phase-1 iteration (backward+GD+prox+forward+loss), Lorenz-96 p={p_total} T={args.T}, "
                        f"B={args.batch}, hidden=64, context=20, lam={LAM}, lr={LR}",
            "parallelism": f"head-shard x{world}" if world > 1 else "single GPU", "heads_per_gpu": heads, "l2": l2}, ws


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the per-head torch-CPU port (oracle/ref_port.py)
# ------------------------------------------------------------------------------------------------
def time_cpu_port(Xb, p, steps, warmup, budget_s=120.0, lam=LAM, lr=LR):
    from oracle import ref_port as RP                                   # checker/baseline only
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    B = Xb.shape[0]
    torch.manual_seed(0)
    # bounded sample: the per-head loop is linear in heads -> time a head subset if the full model would take too long
    probe_heads = min(p, 10)
    conn = np.ones((p, p))

    def build(nh):
        m = RP.PortCRVAE(p, conn, 64)
        m.networks = torch.nn.ModuleList(list(m.networks)[:nh]); m.p = nh
        return m

    m = build(probe_heads)
    smooth, _, _ = RP.smooth_loss(m, Xb, 0.0, 0.1)
    smooth, _, _ = RP.iteration(m, Xb, smooth, lr, lam)                 # warm-up: the first iteration pays allocator / thread-pool start-up
    t0 = time.perf_counter()
    for _ in range(2):
        smooth, _, _ = RP.iteration(m, Xb, smooth, lr, lam)
    t_probe = (time.perf_counter() - t0) / 2
    # per-iteration cost model: encoder (fixed) + heads (linear); the probe over-estimates the full model slightly
    est_full = t_probe * p / probe_heads
    heads = p if est_full * (steps + warmup) <= budget_s else max(1, min(p, int(p * budget_s / (est_full * (steps + warmup)))))
    if heads != probe_heads:
        torch.manual_seed(0)
        m = build(heads)
        smooth, _, _ = RP.smooth_loss(m, Xb, 0.0, 0.1)
    for _ in range(warmup):
        smooth, _, _ = RP.iteration(m, Xb, smooth, lr, lam)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        smooth, _, _ = RP.iteration(m, Xb, smooth, lr, lam)
        times.append(time.perf_counter() - t0)
    total = float(sum(times))
    units = B * TD * heads
    return dict(value=units * steps / total, ms_per_step=1e3 * total / steps, heads=heads, threads=threads,
                sample=f"{steps} iterations (after {warmup} warm-up) of the per-head torch-CPU port on {heads} of {p} heads "
                       f"(K={p} inputs each, full encoder), B={B}; units = B*10*heads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cfg, _ = workload_config(args, max(world, 1))
    p_total = args.p * world if args.scaling == "weak" else args.p
    Xb = make_batch(p_total, args.T, args.batch)
    r = time_cpu_port(Xb, p_total, args.steps, args.warmup)
    model, cores = cpu_info()
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic Lorenz-96 (F=10, seed 0)",
        "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"], "cpu": model},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
STAGE_FNS = ["proj_fwd", "proj_fwd_tc", "split_tf32", "split_tf32_gate_rows", "proj_wgrad_tc", "gru_fwd_tc", "gru_bwd_deferred", "gru_bwd_tc",
             "gru_dwhh_tc", "gru_fwd", "gemm", "latent_fwd", "latent_head_fwd", "latent_head_bwd", "mse_fwd_bwd", "dot_small", "gru_bwd",
             "proj_wgrad", "latent_bwd", "gd_prox_gc", "gd_step", "axpy", "gru_fwd_ll", "gru_bwd_ll", "gru_fwd_mma", "gru_bwd_mma",
             "dz_allreduce", "enc_chain_fwd"]
P_ARG = {"proj_fwd": 4, "gru_fwd": 11, "gru_bwd": 16, "proj_wgrad": 4, "proj_fwd_tc": 6, "proj_wgrad_tc": 5, "gru_fwd_tc": 12,
         "gru_bwd_deferred": 15, "gru_bwd_tc": 14, "gru_dwhh_tc": 6, "gru_fwd_ll": 11, "gru_bwd_ll": 15,
         "gru_fwd_mma": 11, "gru_bwd_mma": 15}


def profile_stages(run, eps_list, reps):
    """Per-stage CUDA-event timing of the eager iteration (same kernels the graph replays), each kernel alone on its
    stream (side streams disabled), so a stage's time is that kernel's launch duration + its launch gap."""
    eng, k = run.eng, run.eng.k
    orig = {}
    records = []

    def wrap(name):
        fn = getattr(k, name)

        def w(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(*a, **kw); e.record()
            tag = name
            if name in P_ARG:
                tag = f"{name}[{'enc' if a[P_ARG[name]] == 1 and eng.P != 1 else 'dec'}]"
            records.append((tag, s, e))
        return w

    names = [n for n in STAGE_FNS if hasattr(k, n)]
    for n in names:
        orig[n] = getattr(k, n)
        setattr(k, n, wrap(n))
    side_was = eng.use_side_stream
    eng.use_side_stream = False
    coll = []
    eng._stage_hook = lambda s, e: coll.append((s, e))          # the dz all-reduce (a collective, not a library kernel call)
    try:
        for r in range(reps):
            run.eng.eps_next.copy_(eps_list[r % len(eps_list)])
            run.update(); run.forward_noeps()
        torch.cuda.synchronize()
    finally:
        for n in names:
            setattr(k, n, orig[n])
        eng.use_side_stream = side_was
        eng._stage_hook = None
    agg = {}
    for tag, s, e in records:
        agg.setdefault(tag, []).append(s.elapsed_time(e))
    if coll:
        agg["allreduce_dz"] = [s.elapsed_time(e) for s, e in coll]
    return {t: {"ms_per_step": float(np.sum(v)) / reps, "calls_per_step": len(v) / reps} for t, v in agg.items()}


def measure_tf32_peak(dev, seconds=1.5):
    """cuBLAS TF32 8192^3 matmul timed like MEASURED_PEAKS.json's bf16 figure: best of 10 (burst) and back to back (sustained)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a, b = torch.randn(n, n, device=dev), torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        flop = 2.0 * n ** 3
        best = 0.0
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); torch.matmul(a, b, out=c); e.record(); torch.cuda.synchronize()
            best = max(best, flop / (s.elapsed_time(e) * 1e-3) / 1e12)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(seconds / (flop / (best * 1e12))))
        s.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e.record(); torch.cuda.synchronize()
        sus = flop * reps / (s.elapsed_time(e) * 1e-3) / 1e12
        return {"tf32_tflops": best, "tf32_tflops_sustained": sus, "how": f"torch.matmul fp32 8192^3, allow_tf32 (cuBLAS TF32): best of 10, {reps} back to back"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def time_gpu_eager(Xb, p, dev, steps=5, warmup=2):
    """The reference's own execution structure (one nn.GRU per head, Python loop, autograd; oracle/ref_port.py) on the B200
    through PyTorch eager + cuDNN with TF32 off: the 'existing GPU kernels' comparator of SURVEY.md 8(d)."""
    from oracle import ref_port as RP                                   # baseline only
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        m = RP.PortCRVAE(p, np.ones((p, p)), 64).to(dev)
        X = Xb.to(dev)
        smooth, _, _ = RP.smooth_loss(m, X, 0.0, 0.1)
        for _ in range(warmup):
            smooth, _, _ = RP.iteration(m, X, smooth, LR, LAM)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            smooth, _, _ = RP.iteration(m, X, smooth, LR, LAM)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        return {"value": Xb.shape[0] * TD * p / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
                "what": f"oracle/ref_port.py (per-head nn.GRU loop + autograd) on cuda, cuDNN, TF32 off, {steps} iterations after {warmup} warm-up"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def time_phase2(V, Xb_dev, p, B, steps, packed=None):
    """Phase-2 iteration (CRVAE on the pruned Lorenz-96 graph + VRAE4E + Adam, CRVAE_lorenz96.py:609-643), graph replay."""
    from vae_connexe_b200.data import lorenz_96_graph
    torch.manual_seed(0)
    c, v = V.CRVAE(p, lorenz_96_graph(p), 64, packed=packed), V.VRAE4E(p, 64)
    run = V.Phase2Runner(c, v, Xb_dev, LR, 0.0, 0.0)
    gen = torch.Generator().manual_seed(1)
    eps = [(torch.randn(B, H, generator=gen).cuda(), torch.randn(B, H, generator=gen).cuda()) for _ in range(8)]
    run.forward(*eps[0]); run.update(); run.forward(*eps[1]); run.capture()
    for i in range(5):
        run.iterate(*eps[i % 8])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        run.iterate(*eps[i % 8])
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    out = {"ms_per_step": ms, "value": B * TD * p / (ms * 1e-3), "unit": UNIT, "steps": steps,
           "what": "train_phase2 iteration (:609-643): VRAE4E backward + Adam, CRVAE backward + GD, forward, residual, VRAE4E forward; "
                   "connection = Lorenz-96 stencil (4 inputs per head, %s)" % ("gather-packed" if c.engine.packed else "masked-dense"),
           "w_ih_bytes": int(c.engine.theta["w_ih"].numel() * 4), "loss": float(c.engine.loss), "loss_e": float(v.engine.loss)}
    run.g_full = run.g_update = run.g_fwd = None
    return out


def time_check_block(V, m, run, eps_dev, check_every=50, reps=5):
    """The check block's extras (:518-555): one more forward + loss read-back, GC readout, best-model snapshot and the
    21-step test-mode generation, amortised over check_every iterations."""
    eng = m.engine
    Xd = torch.zeros(eng.B, 20, eng.p, device=eng.device)
    m(Xd, mode="test")                                     # first call captures the generator's CUDA graph
    torch.cuda.synchronize()
    t_gen = []
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps):
        run.run_forward(eps_dev[i % len(eps_dev)])
        _ = float(eng.loss); _ = float(eng.kl)
        _ = float(100 * torch.mean(m.GC().float()))
        snap = eng.snapshot()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); m(Xd, mode="test"); g1.record()
        t_gen.append((g0, g1))
    e.record(); torch.cuda.synchronize()
    per = s.elapsed_time(e) / reps
    gen = float(np.mean([a.elapsed_time(b) for a, b in t_gen]))
    return {"check_block_ms": per, "generation_ms": gen, "check_every": check_every, "check_block_ms_amortised": per / check_every}


def parity_vs_n1(V, args, Xb_dev, p_total, rank, world, group, steps=25):
    """Sharded == single-GPU: every rank replays `steps` iterations of the sharded model AND of a full single-GPU model
    (same seeds, same noise) and compares its shard against the full model's rows; results are reduced over ranks."""
    import torch.distributed as dist
    B = args.batch
    gen = torch.Generator().manual_seed(99)
    eps = [torch.randn(B, H, generator=gen).cuda() for _ in range(steps + 1)]
    torch.manual_seed(0)
    ms = V.CRVAE(p_total, np.ones((p_total, p_total)), 64, rank=rank, world_size=world, group=group)
    rs = V.Phase1Runner(ms, Xb_dev, LR, LAM, 0.0, 0.1, use_graphs=False)
    rs.forward(eps[0])
    for i in range(steps):
        rs.update(); rs.forward(eps[i + 1])
    loss_s = ms.engine.loss.clone()
    dist.all_reduce(loss_s, group=group)
    gc_s = ms.GC()
    torch.manual_seed(0)
    mf = V.CRVAE(p_total, np.ones((p_total, p_total)), 64)
    rf = V.Phase1Runner(mf, Xb_dev, LR, LAM, 0.0, 0.1, use_graphs=False)
    rf.forward(eps[0])
    for i in range(steps):
        rf.update(); rf.forward(eps[i + 1])
    lo, hi = ms.head_lo, ms.head_hi
    worst = 0.0
    for k in ("w_ih", "w_hh", "b_ih", "b_hh", "w_lin", "b_lin"):
        a, b = ms.engine.theta[k].double(), mf.engine.theta[k][lo:hi].double()
        worst = max(worst, float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)))
    for k in ("enc_w_ih", "enc_w_hh", "enc_b_ih", "enc_b_hh", "lat_w", "lat_b"):
        a, b = ms.engine.theta[k].double(), mf.engine.theta[k].double()
        worst = max(worst, float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)))
    t = torch.tensor([worst, 0.0 if torch.equal(gc_s, mf.GC()) else 1.0], device=Xb_dev.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return {"steps": steps, "loss_sharded_allreduced": float(loss_s), "loss_single_gpu": float(mf.engine.loss),
            "loss_rel_diff": abs(float(loss_s) - float(mf.engine.loss)) / abs(float(mf.engine.loss)),
            "gc_equal_all_ranks": bool(float(t[1]) == 0.0), "max_rel_weight_diff_all_ranks": float(t[0]), "tolerance": 1e-4,
            "pass": bool(float(t[1]) == 0.0 and float(t[0]) < 1e-4)}


def run_ours(args):
    import torch.distributed as dist
    import vae_connexe_b200 as V
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    p_total = args.p * world if args.scaling == "weak" else args.p
    B = args.batch
    cfg, working_set = workload_config(args, world)
    flush = working_set <= 1.5 * L2_BYTES
    Xb = make_batch(p_total, args.T, B)
    torch.manual_seed(0)
    m = V.CRVAE(p_total, np.ones((p_total, p_total)), 64, rank=rank, world_size=world, group=group)
    eng = m.engine
    if os.environ.get("CRVAE_BATCH_TILE"):
        eng.k.set_batch_tile(int(os.environ["CRVAE_BATCH_TILE"]))
    Xb_dev = Xb.to(dev)
    run = V.Phase1Runner(m, Xb_dev, LR, LAM, 0.0, 0.1, use_graphs=not args.no_graphs)
    n_eps = 16
    gen = torch.Generator().manual_seed(1234)
    eps_host = torch.randn(n_eps, B, H, generator=gen).pin_memory()
    eps_dev = eps_host.to(dev)
    X_host = Xb.pin_memory()
    flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev) if flush else None     # 256 MB > L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn):
        """K steps; continuous (one event pair) or, when the working set fits L2, per-step events with an L2 flush between."""
        barrier()
        if not flush:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(args.steps):
                step_fn(i)
            e.record()
            barrier()
            return s.elapsed_time(e)
        evs = []
        for i in range(args.steps):
            flush_buf.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); step_fn(i); e.record()
            evs.append((s, e))
        barrier()
        return float(sum(s.elapsed_time(e) for s, e in evs))

    # warm-up: eager once (also counts launches per step), capture, then W replays
    run.forward(eps_dev[0])
    eng.k.reset_launch_count()
    run.update(); run.forward(eps_dev[1])
    launches_per_step = eng.k.launch_count()
    run.capture()
    W = max(args.warmup, 3)
    for i in range(W):
        run.iterate(eps_dev[i % n_eps])
    barrier()

    # ---- timed region 1: inputs resident in HBM (the `value`) ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler: sampler.__enter__()
    ms_dev = timed(lambda i: run.iterate(eps_dev[i % n_eps]))
    loss_t = run.loss.clone()                   # completes the pending forward of the last flow replay
    if world > 1:
        dist.all_reduce(loss_t)                 # loss = sum over ALL heads: comparable across N
    loss_end = float(loss_t)

    # ---- timed region 2: end to end from HOST buffers (the `e2e`) ----
    # every step: H2D of the step's window batch and noise from pinned host memory (copy stream, two staging slots),
    # update of the previous step, re-bind, forward, D2H of the step's loss into a pinned ring; one synchronisation at the end
    slots = []
    ms_e2e = timed(lambda i: slots.append(run.iterate_from_host(X_host, eps_host[i % n_eps])))
    last = float(run.losses_from_host()[slots[-1]])
    if sampler: sampler.__exit__()

    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    units = B * TD * p_total
    value = units * args.steps / (ms_dev * 1e-3)
    e2e = units * args.steps / (ms_e2e * 1e-3)

    stages = profile_stages(run, [eps_dev[i] for i in range(n_eps)], reps=max(3, min(args.steps, 20)))
    extra = {}
    if world == 1 and not args.lean:
        extra["check_block"] = time_check_block(V, m, run, eps_dev)
    # CUDA graphs must be gone before anything is torn down
    flow_info = {"enabled": run.g_flow is not None,
                 "head_groups": [hi - lo for lo, hi in eng._flow["groups"]] if getattr(eng, "_flow", None) else None}
    run.finish_forward()
    run.g_full = run.g_update = run.g_fwd = run.g_flow = run.g_recupd = run.g_pre = run.g_rec = None
    torch.cuda.synchronize()
    parity = None
    if world > 1:
        parity = parity_vs_n1(V, args, Xb_dev, p_total, rank, world, group)
        dist.barrier()
    if rank == 0:
        pk = _peaks()
        if not args.lean:
            pk.update(measure_tf32_peak(dev))
        P_loc, K = eng.P, p_total
        units_loc = B * TD * P_loc
        tf32_peak = pk.get("tf32_tflops", pk["bf16"] / 2.0)
        tf32_src = "measured here (cuBLAS TF32 8192^3, burst)" if "tf32_tflops" in pk else "bf16 burst / 2 (not measured: --lean)"
        # algorithmic HBM bytes / flops per launch (DESIGN.md "Kernels")
        alg = {}
        for nm in ("gru_bwd", "gru_bwd_deferred", "gru_bwd_tc", "gru_bwd_ll", "gru_bwd_mma"):
            alg[nm + "[dec]"] = ("hbm", 1796.0 * units_loc)
        for nm in ("gru_fwd", "gru_fwd_tc", "gru_fwd_ll", "gru_fwd_mma"):
            alg[nm + "[dec]"] = ("hbm", 1028.0 * units_loc)
        for nm in ("proj_fwd", "proj_wgrad", "proj_fwd_tc", "proj_wgrad_tc"):
            alg[nm + "[dec]"] = ("tensor", 2.0 * K * G * 0.9 * units_loc)
        alg["gd_prox_gc"] = ("hbm", 12.0 * P_loc * G * K)
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")     # ncu dram bytes per launch, re-measured this round (same config only)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("p") == p_total and tj.get("B") == B and world == 1:
                traffic = tj
        roof_all = {}
        for tag, (bound, amount) in alg.items():
            if tag not in stages: continue
            dur = stages[tag]["ms_per_step"] * 1e-3
            if bound == "hbm":
                ach, peak, unit = amount / dur / 1e9, pk["hbm"], "GB/s"
            else:
                ach, peak, unit = amount / dur / 1e12, tf32_peak, "TFLOP/s"
            roof_all[tag] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "ms": dur * 1e3,
                             "traffic": traffic.get(tag), "algorithmic": amount}
        step_ms = ms_dev / args.steps
        dom = max((t for t in roof_all), key=lambda t: roof_all[t]["ms"]) if roof_all else None
        roofline = dict(roof_all[dom], kernel=dom, share_of_step=roof_all[dom]["ms"] / step_ms,
                        peak_source=pk["source"] + "; tensor peak: " + tf32_src) if dom else None

        cpu = gpu_eager = phase2 = None
        if world == 1 and not args.lean:
            phase2 = time_phase2(V, Xb_dev, p_total, B, steps=min(args.steps, 200))
            gpu_eager = time_gpu_eager(Xb, p_total, dev)
        if world == 1 and not args.no_cpu_baseline:
            r = time_cpu_port(Xb, p_total, steps=8, warmup=2, budget_s=25.0)
            model, cores = cpu_info()
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"], "cpu": model}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic Lorenz-96 (F=10, seed 0; reference generator semantics), random-init weights (seed 0)",
            "config": cfg,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(X_host.numel() * 4 + B * H * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "cuda_graphs": not args.no_graphs, "l2_flush_between_steps": bool(flush),
            "flow": flow_info,
            "clocks": sampler.summary() if sampler else None,
            "roofline": roofline, "roofline_all": roof_all, "stages_ms": {k: round(v["ms_per_step"], 4) for k, v in stages.items()},
            "tensor_peak": {k: pk[k] for k in ("tf32_tflops", "tf32_tflops_sustained", "how") if k in pk},
            "cpu_baseline": cpu, "gpu_eager_baseline": gpu_eager, "phase2": phase2, "loss_after_timed": loss_end, "loss_after_e2e": last,
            "parity_vs_n1": parity,
        }
        line.update(extra)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.barrier()
        if hasattr(eng, "close"):
            eng.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--p", type=int, default=100)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true", help="skip the extra legs (phase 2, check block, TF32 peak, GPU-eager baseline)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)      # bounded: time_cpu_port samples a head subset so that steps + warmup fit its time budget
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
