"""Mirrors of the reference's training helpers (CRVAE_lorenz96.py:308-350) for engine-backed heads.

Same names, argument meaning and results; the difference is ownership: the reference REBINDS
`W.data` (:312) / `params.data` (:330), which would silently detach a view from the fused arena,
so these write INTO the fused buffers instead (SURVEY.md 8(b) aliasing hazard)."""
from __future__ import annotations

import numpy as np
import torch

from . import lib as L
from .engine import G as _G, H as _H


def _f32(x: float) -> float:
    return float(np.float32(x))


def _head(network):
    owner = network._owner[0]
    return owner.engine, network._local_idx


def prox_update(network, lam, lr):
    """In-place group-lasso proximal update of one head's weight_ih_l0 (:308-314):
    W[:,j] <- W[:,j] / max(||W[:,j]||, lam*lr) * max(||W[:,j]|| - lr*lam, 0)."""
    eng, i = _head(network)
    mask = None if eng.mask_u8 is None else eng.mask_u8[i:i + 1]
    eng.k.gd_prox_gc(eng.theta["w_ih"][i:i + 1], None, mask, eng.col_norm[i:i + 1], 1, eng.Kw,
                     0.0, _f32(lam * lr), True)
    network.gru.flatten_parameters()


def regularize(network, lam):
    """lam * sum_j ||W[:,j]||_2 over the head's input columns (:316-319)."""
    eng, i = _head(network)
    mask = None if eng.mask_u8 is None else eng.mask_u8[i:i + 1]
    eng.k.gd_prox_gc(eng.theta["w_ih"][i:i + 1], None, mask, eng.col_norm[i:i + 1], 1, eng.Kw, 0.0, 0.0, False)
    return lam * torch.sum(eng.col_norm[i])


class _RidgeFn(torch.autograd.Function):
    """Value by the sumsq kernel; backward adds d/dW = 2*lam*W (times the incoming gradient) into the head's slices
    of the gradient arena (== linear.weight.grad / gru.weight_hh_l0.grad), so `(loss + ridge + ...).backward()` written
    like the reference trainer (:488-497) trains the ridge term too."""

    @staticmethod
    def forward(ctx, anchor, network, lam):
        eng, i = _head(network)
        out = torch.zeros(2, dtype=torch.float32, device=eng.device)
        eng.k.sumsq(eng.theta["w_lin"][i:i + 1], _H, out[0:1])
        eng.k.sumsq(eng.theta["w_hh"][i:i + 1], _G * _H, out[1:2])
        ctx.network, ctx.lam = network, float(lam)
        return lam * (out[0] + out[1])

    @staticmethod
    def backward(ctx, gout):
        eng, i = _head(ctx.network)
        owner = ctx.network._owner[0]
        alpha = 2.0 * ctx.lam * float(gout)
        if alpha != 0.0:
            # The fused BPTT WRITES the gradient arena (it does not accumulate) and autograd does not order the two
            # nodes: if the model's backward has not run yet in this pass, the addition is queued and applied at its end.
            if owner._bwd_ran:
                apply_ridge_grad(eng, i, alpha)
            else:
                owner._ridge_pending.append((i, alpha))
        return None, None, None


def apply_ridge_grad(eng, i, alpha):
    eng.k.axpy(eng.grad["w_hh"][i:i + 1], eng.theta["w_hh"][i:i + 1], _G * _H, alpha)
    eng.k.axpy(eng.grad["w_lin"][i:i + 1], eng.theta["w_lin"][i:i + 1], _H, alpha)


def ridge_regularize(network, lam):
    """lam * (||linear.weight||^2 + ||weight_hh_l0||^2) (:321-325), differentiable (see _RidgeFn)."""
    owner = network._owner[0]
    if torch.is_grad_enabled() and lam != 0:
        return _RidgeFn.apply(owner._anchor, network, lam)
    return _RidgeFn.forward(_NoCtx(), None, network, lam)


class _NoCtx:
    pass


def restore_parameters(model, best_model):
    """Move parameter values from best_model to model (:327-330) -- a device copy of the arena."""
    if hasattr(model, "engine") and hasattr(best_model, "engine"):
        model.engine.theta.flat.copy_(best_model.engine.theta.flat)
        return
    with torch.no_grad():
        for params, best_params in zip(model.parameters(), best_model.parameters()):
            params.copy_(best_params)


def arrange_input(data, context):
    """Arrange a single time series (T, dim) into overlapping windows (:332-350):
    input[n] = data[n:n+context], target[n] = data[n+1:n+context+1]."""
    assert context >= 1 and isinstance(context, int)
    n = len(data) - context
    idx = torch.arange(n, device=data.device)[:, None] + torch.arange(context, device=data.device)[None, :]
    inp = data[idx].to(torch.float32)
    tgt = data[idx + 1].to(torch.float32)
    return inp.detach(), tgt.detach()
