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


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the per-head torch-CPU port (oracle/ref_port.py)
# ------------------------------------------------------------------------------------------------
def time_cpu_port(Xb, p, steps, warmup, budget_s=120.0, lam=0.1, lr=5e-2):
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
    t0 = time.perf_counter()
    smooth, _, _ = RP.iteration(m, Xb, smooth, lr, lam)
    t_probe = time.perf_counter() - t0
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
    Xb = make_batch(args.p, args.T, args.batch)
    r = time_cpu_port(Xb, args.p, args.steps, args.warmup)
    model, cores = cpu_info()
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic Lorenz-96 (F=10, seed 0)",
        "config": {"workload": f"CRVAE phase-1 iteration, Lorenz-96 p={args.p} T={args.T}, B={args.batch}, hidden=64, context=20, lam=0.1, lr=0.05"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"], "cpu": model},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def profile_stages(run, eps_list, reps):
    """Per-stage CUDA-event timing of the eager iteration (same kernels the graph replays)."""
    import vae_connexe_b200.lib as L
    eng, k = run.eng, run.eng.k
    names, evs = [], []
    orig = {}
    stage_fns = ["proj_fwd", "proj_fwd_tc", "split_tf32", "proj_wgrad_tc", "gru_fwd_tc", "gru_bwd_deferred", "gru_bwd_tc", "gru_dwhh_tc", "gru_fwd", "gemm", "latent_fwd", "latent_head_fwd", "latent_head_bwd", "mse_fwd_bwd", "dot_small", "gru_bwd", "proj_wgrad", "latent_bwd",
                 "gd_prox_gc", "gd_step", "axpy"]
    records = []

    def wrap(name):
        fn = getattr(k, name)

        def w(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(*a, **kw); e.record()
            tag = name
            if name in ("proj_fwd", "gru_fwd", "gru_bwd", "proj_wgrad", "proj_fwd_tc", "proj_wgrad_tc", "gru_fwd_tc", "gru_bwd_deferred", "gru_bwd_tc", "gru_dwhh_tc"):
                P = {"proj_fwd": 4, "gru_fwd": 11, "gru_bwd": 16, "proj_wgrad": 4, "proj_fwd_tc": 6, "proj_wgrad_tc": 5, "gru_fwd_tc": 12, "gru_bwd_deferred": 15, "gru_bwd_tc": 14, "gru_dwhh_tc": 6}[name]
                P = a[P]
                tag = f"{name}[{'dec' if P == eng.P and eng.P != 1 else 'enc' if P == 1 else 'dec'}]"
            records.append((tag, s, e))
        return w

    for n in stage_fns:
        orig[n] = getattr(k, n)
        setattr(k, n, wrap(n))
    try:
        for r in range(reps):
            run.eng.eps_next.copy_(eps_list[r % len(eps_list)])
            run.update(); run.forward_noeps()
        torch.cuda.synchronize()
    finally:
        for n in stage_fns:
            setattr(k, n, orig[n])
    agg = {}
    for tag, s, e in records:
        agg.setdefault(tag, []).append(s.elapsed_time(e))
    return {t: {"ms_per_step": float(np.sum(v)) / reps, "calls_per_step": len(v) / reps} for t, v in agg.items()}


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
    Xb = make_batch(p_total, args.T, B)
    torch.manual_seed(0)
    m = V.CRVAE(p_total, np.ones((p_total, p_total)), 64, rank=rank, world_size=world, group=group)
    eng = m.engine
    lr, lam = 5e-2, 0.1
    if os.environ.get("CRVAE_BATCH_TILE"):
        eng.k.set_batch_tile(int(os.environ["CRVAE_BATCH_TILE"]))
    run = V.Phase1Runner(m, Xb.to(dev), lr, lam, 0.0, 0.1, use_graphs=not args.no_graphs)
    n_eps = 16
    gen = torch.Generator().manual_seed(1234)
    eps_host = torch.randn(n_eps, B, H, generator=gen).pin_memory()
    eps_dev = eps_host.to(dev)
    X_host = Xb.pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s.record()
    for i in range(args.steps):
        run.iterate(eps_dev[i % n_eps])
    e.record()
    barrier()
    ms_dev = s.elapsed_time(e)
    loss_end = float(eng.loss)

    # ---- timed region 2: end to end from HOST buffers (the `e2e`) ----
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    last = 0.0
    # every step: H2D of the step's window batch and noise from pinned host memory (copy stream, two staging slots),
    # re-bind + backward+GD+prox+forward, D2H of the step's loss into a pinned ring; one synchronisation at the end
    slot = 0
    for i in range(args.steps):
        slot = run.iterate_from_host(X_host, eps_host[i % n_eps])
    e2.record()
    barrier()
    last = float(run.losses_from_host()[slot])
    ms_e2e = s2.elapsed_time(e2)
    if sampler: sampler.__exit__()

    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    units = B * TD * p_total
    value = units * args.steps / (ms_dev * 1e-3)
    e2e = units * args.steps / (ms_e2e * 1e-3)

    stages = profile_stages(run, [eps_dev[i] for i in range(n_eps)], reps=max(3, min(args.steps, 20)))
    # CUDA graphs that captured NCCL kernels must be gone before the communicator is torn down
    run.g_full = run.g_update = run.g_fwd = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank != 0:
        _hard_exit()
    pk = _peaks()
    P_loc, K = eng.P, p_total
    units_loc = B * TD * P_loc
    # algorithmic HBM bytes / flops per launch (DESIGN.md "Kernels")
    alg = {
        "gru_bwd[dec]": ("hbm", 1796.0 * units_loc), "gru_bwd_deferred[dec]": ("hbm", 1796.0 * units_loc), "gru_bwd_tc[dec]": ("hbm", 1796.0 * units_loc), "gru_fwd[dec]": ("hbm", 1028.0 * units_loc), "gru_fwd_tc[dec]": ("hbm", 1028.0 * units_loc),
        "proj_fwd[dec]": ("tensor", 2.0 * K * G * 0.9 * units_loc), "proj_wgrad[dec]": ("tensor", 2.0 * K * G * 0.9 * units_loc),
        "proj_fwd_tc[dec]": ("tensor", 2.0 * K * G * 0.9 * units_loc), "proj_wgrad_tc[dec]": ("tensor", 2.0 * K * G * 0.9 * units_loc),
        "gd_prox_gc": ("hbm", 12.0 * P_loc * G * K),
    }
    roof_all = {}
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")     # ncu dram bytes per launch (same config only)
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("p") == p_total and tj.get("B") == B and world == 1:
            traffic = tj
    for tag, (bound, amount) in alg.items():
        if tag not in stages: continue
        dur = stages[tag]["ms_per_step"] * 1e-3
        if bound == "hbm":
            ach, peak, unit = amount / dur / 1e9, pk["hbm"], "GB/s"
        else:   # no fp32 tensor mode exists; TF32 dense peak = half the measured bf16 burst
            ach, peak, unit = amount / dur / 1e12, pk["bf16"] / 2.0, "TFLOP/s"
        roof_all[tag] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "ms": dur * 1e3, "traffic": traffic.get(tag),
                         "algorithmic": amount}
    total_stage_ms = sum(v["ms_per_step"] for v in stages.values())
    dom = max((t for t in roof_all), key=lambda t: roof_all[t]["ms"]) if roof_all else None
    roofline = dict(roof_all[dom], kernel=dom, share_of_step=roof_all[dom]["ms"] / total_stage_ms, peak_source=pk["source"]) if dom else None

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = time_cpu_port(Xb, args.p, steps=8, warmup=2, budget_s=25.0)
        model, cores = cpu_info()
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"], "cpu": model}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic Lorenz-96 (F=10, seed 0; reference generator semantics), random-init weights (seed 0)",
        "config": {"workload": f"CRVAE phase-1 iteration (backward+GD+prox+forward+loss), Lorenz-96 p={p_total} T={args.T}, B={B}, hidden=64, context=20, lam=0.1, lr=0.05",
                   "heads_per_gpu": P_loc, "parallelism": f"head-shard x{world}" if world > 1 else "single GPU",
                   "cuda_graphs": not args.no_graphs,
                   "l2": "per-step working set (gate buffer %.0f MB + h/gh_n %.0f MB per GPU) exceeds the 126 MB L2; no explicit flush"
                         % (units_loc * 768 / 1e6, units_loc * 512 / 1e6)},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(X_host.numel() * 4 + B * H * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": int(launches_per_step),
        "clocks": sampler.summary() if sampler else None,
        "roofline": roofline, "roofline_all": roof_all, "stages_ms": {k: round(v["ms_per_step"], 4) for k, v in stages.items()},
        "cpu_baseline": cpu, "loss_after_timed": loss_end, "loss_after_e2e": last,
    }
    print(json.dumps(line))
    if world > 1:
        _hard_exit()


def _hard_exit():
    """Multi-rank runs leave through os._exit after flushing: communicator teardown with captured
    NCCL graphs alive has been seen to hang, and the line is already printed."""
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(0)


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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)      # bounded: time_cpu_port samples a head subset so that steps + warmup fit its time budget
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
