"""train_phase1 / train_phase2 with the reference's signatures and semantics
(CRVAE_lorenz96.py:457-560, :562-698), driven by the fused engine.

What is kept exactly: the single np.random.randint batch draw (:470), the order and size of every
torch.randn draw on the CPU default generator (one (1,B,H) draw per forward :214, in a check block one
more forward draw :522 and one generation draw :225), beta=0.1 / 1 (:475, :582), plain GD on every
CRVAE parameter (:498-499), the group-lasso prox (:502-504), the check-block bookkeeping incl.
best-checkpoint selection with a fresh noise draw (:518-547) and the final restore (:558).

What is different by design: no autograd graph (hand-written backward kernels), the steady-state
iteration is replayed from CUDA graphs, the noise is drawn in blocks (bit-identical stream, see
_NoiseFeed) and the best model is a device copy of the parameter arena instead of deepcopy().
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .engine import H as _H
from .functional import arrange_input


class _NoiseFeed:
    """Sequential (B,H) standard-normal draws from the CPU default generator, identical to calling
    torch.randn(size=(1,B,H)) once per use: a (n,B,H) draw consumes the generator exactly like n
    consecutive (1,B,H) draws (B*H is a multiple of 16, the CPU normal kernel's block).  `total`
    draws are consumed in all, so the generator ends where the reference's would."""

    def __init__(self, B: int, total: int, device, chunk: int = 64):
        self.B, self.left, self.device, self.chunk = B, int(total), device, chunk
        self.buf: Optional[torch.Tensor] = None
        self.pos = 0

    def next(self) -> torch.Tensor:
        if self.buf is None or self.pos >= self.buf.shape[0]:
            if self.left <= 0:
                raise RuntimeError("noise feed exhausted: draw count mis-planned")
            n = min(self.chunk, self.left)
            cuda = torch.device(self.device).type == "cuda"
            host = torch.randn(size=(n, self.B, _H), pin_memory=cuda)
            self.buf = host.to(self.device, non_blocking=True) if cuda else host
            self.left -= n
            self.pos = 0
        out = self.buf[self.pos]
        self.pos += 1
        return out


class Phase1Runner:
    """One CRVAE on one fixed batch: forward / (backward + GD + prox) with optional CUDA-graph
    replay.  Used by train_phase1, bench.py and the parity tests."""

    def __init__(self, crvae, Xb: torch.Tensor, lr: float, lam: float, lam_ridge: float, beta: float,
                 use_graphs: bool = True):
        self.m, self.eng = crvae, crvae.engine
        self.lr, self.lam, self.lam_ridge, self.beta = lr, lam, lam_ridge, beta
        self.eng.bind_batch(Xb)
        self.use_graphs = use_graphs
        self.g_full = self.g_update = self.g_fwd = None
        # software-pipelined iteration (engine.flow_body): graphs of [rec ; update ; pre], [rec ; update], [pre], [rec].
        # state "A": a forward is complete (activations + loss valid); state "P": only its pre half ran (gi and z are
        # ready, the recurrence is pending) -- the state the flow graph starts from and ends in.
        import os as _os
        self.use_flow = use_graphs and _os.environ.get("CRVAE_FLOW", "1") != "0"
        self.g_flow = self.g_recupd = self.g_pre = self.g_rec = None
        self.state = "A"

    # eager pieces ------------------------------------------------------------------------------
    def forward(self, eps: Optional[torch.Tensor]):
        self.eng.forward(eps)
        self.state = "A"

    def update(self):
        self.finish_forward()
        self.eng.backward(self.beta, self.lam_ridge)
        self.eng.step(self.lr, self.lam)

    def finish_forward(self):
        """State P -> A: run the pending recurrence + loss of a forward whose pre half ran inside a flow replay."""
        if self.state == "P":
            if self.g_rec is not None:
                self.g_rec.replay()
            else:
                self.eng.flow_rec()
            self.state = "A"

    # graph capture -------------------------------------------------------------------------------
    def capture(self):
        """Capture [backward+GD+prox], [forward] and their concatenation.  Must be called after at
        least one eager forward+update (kernel attributes set, activations valid)."""
        if not self.use_graphs or self.eng.device.type != "cuda":
            self.use_graphs = False
            return
        torch.cuda.synchronize()
        snap = self.eng.snapshot()
        pool = None
        graphs = []
        for body in ((self.update,), (self.forward_noeps,), (self.update, self.forward_noeps)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                for fn in body:
                    fn()
            pool = g.pool()
            graphs.append(g)
        self.g_update, self.g_fwd, self.g_full = graphs
        if self.use_flow and self.eng.flow_supported():
            eng = self.eng
            eng._flow_setup()                 # streams / workspaces exist before the capture
            bodies = (lambda: eng.flow_body(self.lr, self.lam, self.lam_ridge, self.beta, True),
                      lambda: eng.flow_body(self.lr, self.lam, self.lam_ridge, self.beta, False),
                      eng.flow_pre, eng.flow_rec)
            flow_graphs = []
            for body in bodies:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    body()
                pool = g.pool()
                flow_graphs.append(g)
            self.g_flow, self.g_recupd, self.g_pre, self.g_rec = flow_graphs
        torch.cuda.synchronize()
        self.eng.restore(snap)              # capture does not execute, but stay defensive

    def forward_noeps(self):
        self.eng.forward_staged()
        self.state = "A"

    def iterate(self, eps: torch.Tensor):
        """One steady-state iteration (:497-515): backward, GD, prox, then forward with `eps`.  With the flow graphs the
        forward is left half done (state P: its recurrence runs at the start of the next iteration, overlapped with that
        iteration's backward); finish_forward() -- called by everything that reads activations or the loss -- completes it."""
        self.eng.eps_next.copy_(eps, non_blocking=True)
        if self.g_flow is not None:
            if self.state == "P":
                self.g_flow.replay()
            else:
                self.g_update.replay()
                self.g_pre.replay()
            self.state = "P"
        elif self.g_full is not None:
            self.g_full.replay()
        else:
            self.update()
            self.forward_noeps()

    @property
    def loss(self) -> torch.Tensor:
        """Loss of the most recent forward (completes it first when it is pending)."""
        self.finish_forward()
        return self.eng.loss

    # ------------------------------------------------------------------ host-fed pipeline
    def iterate_from_host(self, X_host: torch.Tensor, eps_host: torch.Tensor) -> int:
        """One iteration whose window batch (B, 20, p) and noise (B, H) come from PINNED HOST memory and whose loss goes
        back to the host, without stalling the device: the two H2D copies run on a copy stream into one of two staging
        slots (so the copy of iteration i+1 overlaps the kernels of iteration i), the compute stream waits on the copy
        event, re-binds the batch and replays the iteration, and the loss is copied (async) into a pinned ring.  Returns
        the ring index of this iteration's loss; `losses_from_host()` synchronises and returns the ring."""
        dev = self.eng.device
        st = getattr(self, "_host", None)
        if st is None or st["X"][0].shape != X_host.shape:
            st = self._host = {
                "copy": torch.cuda.Stream(device=dev),
                "X": [torch.empty(X_host.shape, dtype=torch.float32, device=dev) for _ in range(2)],
                "eps": [torch.empty(eps_host.shape, dtype=torch.float32, device=dev) for _ in range(2)],
                "copied": [torch.cuda.Event() for _ in range(2)], "consumed": [torch.cuda.Event() for _ in range(2)],
                "loss": torch.zeros(4096, dtype=torch.float32).pin_memory(), "i": 0}
            for ev in st["consumed"]:
                ev.record(torch.cuda.current_stream(dev))
        i = st["i"]; j = i & 1
        with torch.cuda.stream(st["copy"]):
            st["copy"].wait_event(st["consumed"][j])                  # slot j was last read two iterations ago
            st["X"][j].copy_(X_host, non_blocking=True)
            st["eps"][j].copy_(eps_host, non_blocking=True)
            st["copied"][j].record(st["copy"])
        cur = torch.cuda.current_stream(dev)
        # Order matters: the backward of the PREVIOUS forward still reads the batch it was computed on (the weight
        # gradients multiply dgates by enc_in / dec_in), so the update runs first, THEN the new batch is bound, THEN
        # the forward -- exactly the reference's order when its loop resamples (CR-CS-RAE.py:557-558 precede the forward).
        slot = i % st["loss"].numel()
        if self.g_flow is not None:
            # flow form: [rec of the previous forward ; update] -> bind -> [pre]; this step's recurrence (and loss) runs at
            # the start of the next call, overlapped with its backward, so the loss reaches its ring slot one call later
            if self.state == "P":
                self.g_recupd.replay()
                prev = st.get("pending")
                if prev is not None:
                    st["loss"][prev:prev + 1].copy_(self.eng.loss, non_blocking=True)
            else:
                self.g_update.replay()
            cur.wait_event(st["copied"][j])
            self.eng.bind_batch(st["X"][j])
            self.eng.eps_next.copy_(st["eps"][j], non_blocking=True)
            self.g_pre.replay()
            self.state = "P"
            st["pending"] = slot
        else:
            self.run_update()
            cur.wait_event(st["copied"][j])
            self.eng.bind_batch(st["X"][j])
            self.run_forward(st["eps"][j])
            st["loss"][slot:slot + 1].copy_(self.eng.loss, non_blocking=True)
        st["consumed"][j].record(cur)
        st["i"] = i + 1
        return slot

    def losses_from_host(self) -> torch.Tensor:
        st = self._host
        if self.state == "P":
            self.finish_forward()
            if st.get("pending") is not None:
                st["loss"][st["pending"]:st["pending"] + 1].copy_(self.eng.loss, non_blocking=True)
                st["pending"] = None
        torch.cuda.synchronize(self.eng.device)
        return st["loss"]

    def run_update(self):
        self.finish_forward()
        if self.g_update is not None:
            self.g_update.replay()
        else:
            self.update()

    def run_forward(self, eps: torch.Tensor):
        self.eng.eps_next.copy_(eps, non_blocking=True)
        if self.g_fwd is not None:
            self.g_fwd.replay()
            self.state = "A"
        else:
            self.forward_noeps()


def _planned_draws(max_iter: int, check_every: int) -> int:
    checks = len(range(0, max_iter, check_every)) if max_iter > 0 else 0
    return 1 + max_iter + 2 * checks


def train_phase1(crvae, X, context, lr, max_iter, lam=0, lam_ridge=0,
                 lookback=5, check_every=50, verbose=1, sparsity=100, batch_size=256,
                 use_graphs=True, log: Optional[List[dict]] = None, generate_in_check: bool = False):
    """Phase-1 trainer, same signature as the reference (:457-458) (+ keyword-only extras: use_graphs, log,
    generate_in_check = also materialise the check block's test-mode sample (:550-555; the reference computes it,
    min-max scales it and throws it away -- off by default, its noise draw is consumed either way; when on, the last
    sample is left in crvae.last_sample).  X: (n_series, T, p) fp32 on the model's device.  Returns the (always
    empty, :463/:560) train_loss_list; the model is left at its best checkpoint (:558)."""
    p = X.shape[-1]
    eng = crvae.engine
    train_loss_list = []
    # Set up data (:466-473): windows of every series, ONE index draw, fixed batch.
    X_all = torch.cat([arrange_input(x, context)[0] for x in X], dim=0)
    idx = np.random.randint(len(X_all), size=(batch_size,))
    Xb = X_all[torch.from_numpy(idx).to(X_all.device)]
    beta = 0.1                                                    # :475
    best_it, best_loss, best_snap = None, np.inf, None
    B = Xb.shape[0]
    feed = _NoiseFeed(B, _planned_draws(max_iter, check_every), eng.device)
    run = Phase1Runner(crvae, Xb, lr, lam, lam_ridge, beta, use_graphs=use_graphs)
    rank0 = getattr(crvae, "rank", 0) == 0
    world = getattr(crvae, "world_size", 1)

    run.forward(feed.next())                                      # :482-489
    captured = False
    for it in range(max_iter):
        check = it % check_every == 0
        eps_train = feed.next()                                   # :508 draw
        if not check and captured:
            run.iterate(eps_train)                                # :497-515, one graph replay
            continue
        if check:
            eps_check = feed.next()                               # :522 draw
            h0_gen = feed.next()                                  # :550 -> :225 generation draw
        run.run_update()                                          # :497-506 (completes a pending flow forward first)
        if check:
            # The check-block forward (:522) and the training forward (:508) use the same weights,
            # so they are evaluated in the opposite order: the activations left in the engine are
            # then the ones `smooth` is built on, ready for the next backward.
            run.run_forward(eps_check)
            loss_t = eng.loss.clone()
            ridge_t = eng.ridge_value(lam_ridge)
            if world > 1:
                torch.distributed.all_reduce(loss_t, group=crvae.group)
                if lam_ridge != 0:
                    ridge_t = ridge_t.clone()
                    torch.distributed.all_reduce(ridge_t, group=crvae.group)
        run.run_forward(eps_train)
        if use_graphs and not captured:
            run.capture()                      # no-op (eager replay) on a non-CUDA test backend
            captured = True
        if not check:
            continue
        if generate_in_check:                                     # :550 (one CUDA-graph replay, generate.py)
            from .generate import crvae_generate
            crvae.last_sample = crvae_generate(crvae, Xb, phase=0, h0=h0_gen)
        # ---- rest of the check block (:524-547) ------------------------------------------------
        mean_loss = np.float32(np.float32(float(loss_t) + float(ridge_t)) / np.float32(p))   # :530-533
        kl_val = float(eng.kl)
        usage = gc_now = None
        if lam > 0:
            gc_now = crvae.GC()
            usage = float(100 * torch.mean(gc_now.float()))       # :541-542
        if verbose > 0 and rank0:
            print(('-' * 10 + 'Iter = %d' + '-' * 10) % (it))
            print('Loss = %f' % mean_loss)
            print('KL = %f' % kl_val)
            if lam > 0:
                print('Variable usage = %.2f%%' % usage)
        if log is not None:
            log.append(dict(it=it, mean_loss=float(mean_loss), kl=kl_val, usage=usage,
                            gc=None if gc_now is None else gc_now.cpu().numpy().astype(np.int8)))
        if mean_loss < best_loss:                                 # :544-547
            best_loss, best_it = mean_loss, it
            best_snap = eng.snapshot()
    run.finish_forward()
    if best_snap is not None:                                     # :558 restore best model
        eng.restore(best_snap)
    crvae.best_it = best_it
    return train_loss_list


class Phase2Runner:
    """CRVAE + VRAE4E on one fixed batch (phase 2, :609-643): [VRAE backward + Adam, CRVAE backward +
    GD (+prox)] and [CRVAE forward -> residual -> VRAE forward], with optional CUDA-graph replay."""

    def __init__(self, crvae, vrae, Xb, lr, lam, lam_ridge, beta=1.0, beta_e=1.0, use_graphs=True):
        self.c, self.v = crvae.engine, vrae.engine
        self.lr, self.lam, self.lam_ridge, self.beta, self.beta_e = lr, lam, lam_ridge, beta, beta_e
        self.c.bind_batch(Xb)
        self.use_graphs = use_graphs
        self.g_full = self.g_update = self.g_fwd = None

    def update(self):
        # The two backward passes are independent (:611-623 touch disjoint parameters): the VRAE4E chain (some thirty small,
        # latency-bound launches) runs on a twin stream next to the CRVAE's big kernels; fork / join are capturable.
        twin = None
        if self.c.device.type == "cuda" and self.c.use_side_stream:
            if getattr(self, "_twin", None) is None:
                self._twin = torch.cuda.Stream(device=self.c.device)
            twin = self._twin
            ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream()); twin.wait_event(ev)
        import contextlib
        with (torch.cuda.stream(twin) if twin is not None else contextlib.nullcontext()):
            self.v.backward(self.beta_e)                  # smooth_e.backward() (:611)
            if self.lam == 0:
                self.v.adam_step()                        # optimizer.step(); optimizer.zero_grad() (:612-614)
        self.c.backward(self.beta, self.lam_ridge)    # smooth.backward() (:616)
        self.c.step(self.lr, self.lam)                # GD (:617-618) + prox (:621-623)
        if twin is not None:
            ev = torch.cuda.Event(); ev.record(twin); torch.cuda.current_stream().wait_event(ev)

    def forward_staged(self):
        self.c.forward_staged(want_err=True)          # :630-637
        self.v.bind_error(self.c.residual())          # :639
        self.v.forward_staged()                       # :640-643

    def forward(self, eps_c, eps_e):
        self.c.eps_next.copy_(eps_c, non_blocking=True)
        self.v_stage(eps_e)
        if self.g_fwd is not None:
            self.g_fwd.replay()
        else:
            self.forward_staged()

    def v_stage(self, eps_e):
        if self.v.B is None:
            self.v._alloc(self.c.B)
        self.v.eps_next.copy_(eps_e, non_blocking=True)

    def capture(self):
        if not self.use_graphs or self.c.device.type != "cuda":
            self.use_graphs = False
            return
        torch.cuda.synchronize()
        pool, graphs = None, []
        for body in ((self.update,), (self.forward_staged,), (self.update, self.forward_staged)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                for fn in body:
                    fn()
            pool = g.pool()
            graphs.append(g)
        self.g_update, self.g_fwd, self.g_full = graphs
        torch.cuda.synchronize()

    def run_update(self):
        if self.g_update is not None:
            self.g_update.replay()
        else:
            self.update()

    def iterate(self, eps_c, eps_e):
        self.c.eps_next.copy_(eps_c, non_blocking=True)
        self.v.eps_next.copy_(eps_e, non_blocking=True)
        if self.g_full is not None:
            self.g_full.replay()
        else:
            self.update()
            self.forward_staged()


def train_phase2(crvae, vrae, X, context, lr, max_iter, lam=0, lam_ridge=0,
                 lookback=5, check_every=50, verbose=1, sparsity=100, batch_size=256,
                 use_graphs=True, log: Optional[List[dict]] = None):
    """Phase-2 trainer, same signature as the reference (:562-563).  Keeps: Adam(lr=1e-3) on the VRAE
    stepped only when lam == 0 (:565, :612-614); GD (+prox) on the CRVAE; beta = beta_e = 1
    (:582-583); the unused np.random.randint draw of every iteration (:628); noise-draw order
    (CRVAE :630, VRAE :640; check block: CRVAE :648, VRAE test :173, CRVAE test :266); best CRVAE
    checkpoint restore (:673-676, :696).  The check block's generated sample is not materialised
    (the reference only plots / saves it, :685-693) but its noise draws are consumed."""
    p = X.shape[-1]
    ce, ve = crvae.engine, vrae.engine
    train_loss_list = []
    X_all = torch.cat([arrange_input(x, context)[0] for x in X], dim=0)
    idx = np.random.randint(len(X_all), size=(batch_size,))                 # :576
    Xb = X_all[torch.from_numpy(idx).to(X_all.device)]
    B = Xb.shape[0]
    best_it, best_loss, best_snap = None, np.inf, None
    checks = len(range(0, max_iter, check_every)) if max_iter > 0 else 0
    feed = _NoiseFeed(B, 2 + 2 * max_iter + 3 * checks, ce.device)
    run = Phase2Runner(crvae, vrae, Xb, lr, lam, lam_ridge, use_graphs=use_graphs)
    rank0 = getattr(crvae, "rank", 0) == 0
    world = getattr(crvae, "world_size", 1)

    run.forward(feed.next(), feed.next())                                   # :590-603
    captured = False
    for it in range(max_iter):
        check = it % check_every == 0
        eps_c, eps_e = feed.next(), feed.next()                             # :630, :640 draws
        if check:
            eps_check = feed.next()                                         # :648
            feed.next(); feed.next()                                        # :679 -> :173 and :681 -> :266 (samples not used)
        if not check and captured:
            np.random.randint(len(X_all), size=(batch_size,))               # :628 (result unused, stream advanced)
            run.iterate(eps_c, eps_e)
            continue
        run.run_update()                                                    # :611-625
        np.random.randint(len(X_all), size=(batch_size,))                   # :628
        if check:
            ce.eps_next.copy_(eps_check, non_blocking=True)
            ce.forward_staged()                                             # :648 (evaluated before :630, same weights)
            loss_t = ce.loss.clone()
            ridge_t = ce.ridge_value(lam_ridge)
            if world > 1:
                torch.distributed.all_reduce(loss_t, group=crvae.group)
                if lam_ridge != 0:
                    ridge_t = ridge_t.clone()
                    torch.distributed.all_reduce(ridge_t, group=crvae.group)
        run.forward(eps_c, eps_e)                                           # :630-643
        if use_graphs and not captured:
            run.capture()
            captured = True
        if not check:
            continue
        mean_loss = np.float32(np.float32(float(loss_t) + float(ridge_t)) / np.float32(p))    # :654-659
        kl_val, kl_e = float(ce.kl), float(ve.kl)
        smooth_e = float(np.float32(float(ve.loss)) + np.float32(kl_e))
        usage = float(100 * torch.mean(crvae.GC().float())) if lam > 0 else None
        if verbose > 0 and rank0:
            print(('-' * 10 + 'Iter = %d' + '-' * 10) % (it))
            print('Loss = %f' % mean_loss)
            print('KL = %f' % kl_val)
            print('Loss_e = %f' % smooth_e)
            print('KL_e = %f' % kl_e)
            if lam > 0:
                print('Variable usage = %.2f%%' % usage)
        if log is not None:
            log.append(dict(it=it, mean_loss=float(mean_loss), kl=kl_val, loss_e=smooth_e, kl_e=kl_e, usage=usage))
        if mean_loss < best_loss:                                           # :673-676 (CRVAE only)
            best_loss, best_it = mean_loss, it
            best_snap = ce.snapshot()
    if best_snap is not None:                                               # :696
        ce.restore(best_snap)
    crvae.best_it = best_it
    return train_loss_list
