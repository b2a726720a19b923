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

    # eager pieces ------------------------------------------------------------------------------
    def forward(self, eps: Optional[torch.Tensor]):
        self.eng.forward(eps)

    def update(self):
        self.eng.backward(self.beta, self.lam_ridge)
        self.eng.step(self.lr, self.lam)

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
        torch.cuda.synchronize()
        self.eng.restore(snap)              # capture does not execute, but stay defensive

    def forward_noeps(self):
        self.eng.forward_staged()

    def iterate(self, eps: torch.Tensor):
        """One steady-state iteration (:497-515): backward, GD, prox, then forward with `eps`."""
        self.eng.eps_next.copy_(eps, non_blocking=True)
        if self.g_full is not None:
            self.g_full.replay()
        else:
            self.update()
            self.forward_noeps()

    def run_update(self):
        if self.g_update is not None:
            self.g_update.replay()
        else:
            self.update()

    def run_forward(self, eps: torch.Tensor):
        self.eng.eps_next.copy_(eps, non_blocking=True)
        if self.g_fwd is not None:
            self.g_fwd.replay()
        else:
            self.forward_noeps()


def _planned_draws(max_iter: int, check_every: int) -> int:
    checks = len(range(0, max_iter, check_every)) if max_iter > 0 else 0
    return 1 + max_iter + 2 * checks


def train_phase1(crvae, X, context, lr, max_iter, lam=0, lam_ridge=0,
                 lookback=5, check_every=50, verbose=1, sparsity=100, batch_size=256,
                 use_graphs=True, log: Optional[List[dict]] = None):
    """Phase-1 trainer, same signature as the reference (:457-458) (+ two keyword-only extras:
    use_graphs, log).  X: (n_series, T, p) fp32 on the model's device.  Returns the (always
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
            feed.next()                                           # :550 -> :225 generation draw (output unused)
        run.run_update()                                          # :497-506
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
        # ---- rest of the check block (:524-547) ------------------------------------------------
        mean_loss = np.float32(np.float32(float(loss_t) + float(ridge_t)) / np.float32(p))   # :530-533
        kl_val = float(eng.kl)
        usage = None
        if lam > 0:
            usage = float(100 * torch.mean(crvae.GC().float()))   # :541-542
        if verbose > 0 and rank0:
            print(('-' * 10 + 'Iter = %d' + '-' * 10) % (it))
            print('Loss = %f' % mean_loss)
            print('KL = %f' % kl_val)
            if lam > 0:
                print('Variable usage = %.2f%%' % usage)
        if log is not None:
            log.append(dict(it=it, mean_loss=float(mean_loss), kl=kl_val, usage=usage))
        if mean_loss < best_loss:                                 # :544-547
            best_loss, best_it = mean_loss, it
            best_snap = eng.snapshot()
    if best_snap is not None:                                     # :558 restore best model
        eng.restore(best_snap)
    crvae.best_it = best_it
    return train_loss_list
