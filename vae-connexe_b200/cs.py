"""CR-CS-RAE variant (reference CR-CS-RAE.py): the same multi-head recurrent engine with the KL term
replaced by a Cauchy-Schwarz divergence to a learnable equal-weight GMM prior.

Mirrors `CRVAE(num_series, connection, hidden, K, lambda_cs)` (:249-374, adds `.prior` :268) and the
rewritten `train_phase1` (:529-651): the mini-batch is RESAMPLED every iteration (:557-558), every
parameter incl. the prior gets the plain GD step (:591-594), the check block evaluates on ALL windows
without gradients (:606-622) and compares (recon + ridge + lambda_cs*cs)/p (:622).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import lib as L
from .engine import CRVAEEngine, H as _H
from .functional import arrange_input
from .modules import CRVAE as _BaseCRVAE


class GMMPrior(nn.Module):
    """Learnable isotropic Gaussian mixture prior with equal weights (:107-121), stored in a small arena."""

    def __init__(self, K: int, latent_dim: int, device, _init: bool = True):
        super().__init__()
        self.K, self.latent_dim = K, latent_dim
        self.flat = torch.zeros(2 * K * latent_dim, dtype=torch.float32, device=device)
        self.gflat = torch.zeros_like(self.flat)
        n = K * latent_dim
        self.mu = nn.Parameter(self.flat[:n].view(K, latent_dim))
        self.logvar = nn.Parameter(self.flat[n:].view(K, latent_dim))
        self.mu.grad = self.gflat[:n].view(K, latent_dim)
        self.logvar.grad = self.gflat[n:].view(K, latent_dim)
        if _init:
            with torch.no_grad():
                self.mu.copy_(torch.randn(K, latent_dim) * 0.05)      # :114, drawn on the CPU default generator

    @property
    def var(self):
        return self.logvar.exp()

    def forward(self):
        return self.mu, self.var


class CRVAE(_BaseCRVAE):
    """CRVAE(num_series, connection, hidden, K, lambda_cs) of CR-CS-RAE.py (:249-374)."""

    def __init__(self, num_series, connection, hidden, K, lambda_cs, **kw):
        self._K = int(K)
        self._prior_holder = []
        super().__init__(num_series, connection, hidden, **kw)
        self.lambda_cs = lambda_cs

    def _init_extra(self):
        # declaration order of CR-CS-RAE.py (:259-271): gru_left, fc_mu, fc_std, PRIOR, then the heads -- the prior's
        # randn draw happens here (after the encoder's init draws, before the heads'); the module itself is registered
        # by _register_extra so that parameters() keeps the reference's order
        self._prior_holder.append(GMMPrior(self._K, _H, self.device))

    def _register_extra(self):
        if not self._prior_holder:          # _init=False (deepcopy): no generator draw
            self._prior_holder.append(GMMPrior(self._K, _H, self.device, _init=False))
        self.prior = self._prior_holder[0]

    def _clone_args(self):
        return super()._clone_args() + (self._K, self.lambda_cs)

    def _copy_extra_to(self, new):
        new.prior.flat.copy_(self.prior.flat)
        new.prior.gflat.copy_(self.prior.gflat)


class CSPhase1Runner:
    """One iteration of CR-CS-RAE's train_phase1 (:553-601) on the fused engine + the CS head kernels."""

    def __init__(self, crvae: CRVAE, lr, lam, lam_ridge, lambda_cs):
        self.m, self.eng = crvae, crvae.engine
        self.lr, self.lam, self.lam_ridge, self.lambda_cs = lr, lam, lam_ridge, lambda_cs
        self.k = self.eng.k
        dev = self.eng.device
        self.cs_mean = torch.zeros(1, device=dev)
        self.ws = None

    def _head(self, eng, want_grad: bool):
        B, K = eng.B, self.m.prior.K
        need = self.k.cs_div_workspace(B, K) // 4 + 4
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.zeros(need, dtype=torch.float32, device=eng.device)
        if not hasattr(eng, "dlat_cs") or eng.dlat_cs.shape[0] != B:
            eng.dlat_cs = torch.zeros(B, 2 * _H, device=eng.device)
        pr = self.m.prior
        n = K * _H
        self.k.cs_div_fwd_bwd(eng.lat, pr.flat[:n], pr.flat[n:], B, K, self.lambda_cs if want_grad else 0.0, self.cs_mean,
                              eng.dlat_cs, pr.gflat[:n], pr.gflat[n:], self.ws)

    def iteration(self, Xb, eps):
        """forward (:563) -> recon + ridge + lambda_cs*cs (:566-582) -> backward (:585) -> GD on every parameter
        incl. the prior (:591-594) -> prox (:597-599)."""
        eng = self.eng
        eng.bind_batch(Xb)
        eng.forward(eps)
        self._head(eng, True)
        eng.backward(beta=0.0, lam_ridge=self.lam_ridge, dlat_extra=eng.dlat_cs)
        eng.step(self.lr, self.lam)
        pr = self.m.prior
        self.k.gd_step(pr.flat, pr.gflat, pr.flat.numel(), float(np.float32(self.lr)))


def train_phase1(crvae, X, context, lr, max_iter, lam=0, lam_ridge=0, lookback=5, check_every=50, verbose=1,
                 sparsity=100, batch_size=2048, lambda_cs=0.1, log: Optional[List[dict]] = None):
    """CR-CS-RAE.py train_phase1 (:529-651), same signature."""
    p = X.shape[-1]
    eng = crvae.engine
    X_all = torch.cat([arrange_input(x, context)[0] for x in X], dim=0)
    run = CSPhase1Runner(crvae, lr, lam, lam_ridge, lambda_cs)
    # evaluation engine on ALL windows (:606) sharing the parameter arena
    ev = CRVAEEngine(eng.p, eng.mask_np, head_off=eng.head_off, device=eng.device, group=eng.group, packed=eng.packed)
    ev.theta = eng.theta
    if hasattr(eng, "w_ih_hi"):
        pass
    best_it, best_loss, best_snap, best_prior = None, np.inf, None, None
    dev = eng.device
    cuda = dev.type == "cuda"
    for it in range(max_iter):
        idx = np.random.randint(len(X_all), size=(batch_size,))                 # :557
        Xb = X_all[torch.from_numpy(idx).to(X_all.device)]
        eps = torch.randn(size=(1, Xb.shape[0], _H))[0].to(dev)                  # forward draw (:285)
        run.iteration(Xb, eps)
        if it % check_every == 0:                                                # :604-640
            ev.bind_batch(X_all)
            ev.forward(torch.randn(size=(1, X_all.shape[0], _H))[0].to(dev))
            run._head(ev, False)
            loss_t = float(ev.loss)
            ridge_t = float(ev.ridge_value(lam_ridge))
            cs_t = float(run.cs_mean)
            mean_loss = np.float32(np.float32(np.float32(loss_t) + np.float32(ridge_t) + np.float32(lambda_cs) * np.float32(cs_t))
                                   / np.float32(p))
            usage = float(100 * torch.mean(crvae.GC().float())) if lam > 0 else None
            if verbose > 0:
                print(('-' * 10 + 'Iter = %d' + '-' * 10) % (it))
                print('Mean Loss = %f' % mean_loss)
                print('Recon Loss = %f' % (loss_t / p))
                print('CS_Div = %f' % cs_t)
                if lam > 0:
                    print('Variable usage = %.2f%%' % usage)
            if log is not None:
                log.append(dict(it=it, mean_loss=float(mean_loss), recon=loss_t / p, cs=cs_t, usage=usage))
            if mean_loss < best_loss:
                best_loss, best_it = mean_loss, it
                best_snap, best_prior = eng.snapshot(), crvae.prior.flat.clone()
                if verbose > 0:
                    print(f"*** New best model at iter {best_it} with loss {best_loss:.4f} ***")
    if best_snap is not None:
        eng.restore(best_snap)
        crvae.prior.flat.copy_(best_prior)
    crvae.best_it = best_it
    return []
