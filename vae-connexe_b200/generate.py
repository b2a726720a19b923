"""Test-mode autoregressive generation (CRVAE.forward(mode='test') :223-243 / :264-284 and
VRAE4E.forward(mode='test') :171-179): 21 one-step updates in which every head's next input is the
vector of all heads' previous outputs.  Built from the same kernels as training with T = 1:
projection -> one recurrent step (per-head hidden state carried in [P,B,H]) -> transpose of the p
scalar outputs into the next input row."""
from __future__ import annotations

import torch

from .engine import G, H

GEN_STEPS = 21          # int(20/1)+1  (:229, :174)


def crvae_generate(model, X, noise=None, phase=0):
    """Returns X_seq (B, 21, p) like the reference.  phase=1 adds 0.1*noise[:, i] to every generated
    step (:281-283).  Draws h_0 ~ N(0,1) of size (1,B,H) on the CPU generator (:225 / :266)."""
    eng, k = model.engine, model.engine.k
    if model.world_size > 1:
        raise NotImplementedError("test-mode generation needs every head's output per step; run it on one rank")
    B, P, p = X.shape[0], eng.P, eng.p
    dev = eng.device
    h0 = torch.randn(size=(1, B, H)).to(dev)[0]
    h = h0.unsqueeze(0).expand(P, B, H).contiguous()                 # every head starts from the same h_0 (:227-228)
    h_next = torch.empty_like(h)
    x = torch.zeros(1, B, p, device=dev)                               # X_seq starts as one zero step (:224)
    gates = torch.empty(P, 1, B, G, device=dev)
    ghn = torch.empty(P, 1, B, H, device=dev)
    pred = torch.empty(P, 1, B, device=dev)
    th = eng.theta
    out = torch.empty(B, GEN_STEPS, p, device=dev)
    xt = torch.empty(B, p, device=dev)
    for i in range(GEN_STEPS):
        k.proj_fwd(x, th["w_ih"], th["b_ih"], gates, P, 1, B, p, 0)
        k.gru_fwd(gates, th["b_ih"], th["w_hh"], th["b_hh"], h, B * H, th["w_lin"], th["b_lin"],
                  h_next.view(P, 1, B, H), ghn, pred, P, 1, B, 0)
        k.transpose(pred.view(P, B), xt, P, B)                         # X_t = cat(out_j) over heads (:233-236)
        step = xt
        if phase == 1:
            step = xt + 0.1 * noise[:, i, :].to(dev)                   # :281-283 (only the stored sequence is perturbed)
        out[:, i, :] = step
        x = step.reshape(1, B, p).contiguous()                         # next input = last element of X_seq (:232)
        h, h_next = h_next, h
    return out


def vrae_generate(model, X):
    """VRAE4E test mode (:171-179): X_seq (B, 22, p): the zero step followed by 21 generated steps."""
    eng, k = model.engine, model.engine.k
    B, p = X.shape[0], eng.p
    dev = eng.device
    h = torch.randn(size=(1, B, H)).to(dev)[0].contiguous()            # :173
    h_next = torch.empty_like(h)
    th = eng.theta
    x = torch.zeros(1, B, p, device=dev)
    gates = torch.empty(1, 1, B, G, device=dev)
    ghn = torch.empty(1, 1, B, H, device=dev)
    out = torch.zeros(B, GEN_STEPS + 1, p, device=dev)
    y = torch.empty(B, p, device=dev)
    from . import lib as L
    for i in range(GEN_STEPS):
        k.proj_fwd(x, th["dec_w_ih"].view(1, G, p), th["dec_b_ih"], gates, 1, 1, B, p, 0)
        k.gru_fwd(gates, th["dec_b_ih"], th["dec_w_hh"], th["dec_b_hh"], h, 0, None, None,
                  h_next.view(1, 1, B, H), ghn, None, 1, 1, B, 0)
        k.gemm(L.GEMM_NT, 1, B, p, H, h_next, H, 0, th["out_w"], H, 0, y, p, 0, th["out_b"], 0)
        out[:, i + 1, :] = y
        x = y.reshape(1, B, p).clone()
        h, h_next = h_next, h
    return out
