"""TEST / BASELINE INFRASTRUCTURE ONLY -- a CPU port of the reference's phase-1 iteration that keeps
the reference's EXECUTION STRUCTURE (one torch.nn.GRU + nn.Linear per variable, a Python loop over
heads, torch autograd for the backward, one elementwise update per parameter tensor), so that timing
it on the GPU box's host cores measures what a user of the reference experiences on CPU
(bench.py: cpu_baseline kind "port", and the `--impl reference` arm; the reference's own file cannot
travel to the GPU box).  Algorithm restated from CRVAE_lorenz96.py:97-121, :181-221, :308-314,
:484-515; pinned against tests/golden/p4_step.npz by tests/test_oracle_golden.py::test_ref_port.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


class PortHead(nn.Module):
    def __init__(self, k_in, hidden):
        super().__init__()
        self.gru = nn.GRU(k_in, hidden, batch_first=True)       # :104
        self.linear = nn.Linear(hidden, 1)                      # :106


class PortCRVAE(nn.Module):
    def __init__(self, p, connection, hidden=64):
        super().__init__()
        self.p, self.hidden, self.connection = p, hidden, np.asarray(connection)
        self.gru_left = nn.GRU(p, hidden, batch_first=True)     # :192
        self.fc_mu = nn.Linear(hidden, hidden)                  # :195
        self.fc_std = nn.Linear(hidden, hidden)                 # :196
        self.networks = nn.ModuleList([PortHead(int(self.connection[:, i].sum()), hidden) for i in range(p)])
        self.cols = [np.where(self.connection[:, i] != 0)[0] for i in range(p)]

    def forward(self, X):
        Xz = torch.cat((torch.zeros_like(X[:, 0:1, :]), X), 1)                          # :205
        _, h_t = self.gru_left(Xz[:, 1:11, :], torch.zeros(1, X.shape[0], self.hidden, device=X.device))   # :207-208
        mu, log_var = self.fc_mu(h_t), self.fc_std(h_t)                                 # :210-211
        # the reference draws on the CPU default generator and moves the noise to mu's device (:214-215)
        z = mu + torch.exp(0.5 * log_var) * torch.randn(size=mu.size()).type_as(mu)     # :213-216
        pred = []
        for i, net in enumerate(self.networks):                                         # :218-219
            Xi = Xz[:, :, self.cols[i]]                                                 # :115
            out, _ = net.gru(torch.cat((Xi[:, 0:1, :], Xi[:, 11:-1, :]), 1), z)         # :119
            pred.append(net.linear(out))                                                # :120
        return pred, log_var, mu                                                        # :221


def smooth_loss(model, X, lam_ridge, beta):
    pred, mu, log_var = model(X)                    # swapped unpacking, exactly as the trainer (:482/:508)
    mse = nn.MSELoss()
    loss = sum([mse(pred[i][:, :, 0], X[:, 10:, i]) for i in range(model.p)])                       # :484
    mmd = (-0.5 * (1 + log_var - mu ** 2 - torch.exp(log_var)).sum(dim=-1).sum(dim=0)).mean(dim=0)   # :486
    ridge = sum([lam_ridge * (torch.sum(n.linear.weight ** 2) + torch.sum(n.gru.weight_hh_l0 ** 2))
                 for n in model.networks])                                                          # :321-325
    return loss + ridge + beta * mmd, loss, mmd


def iteration(model, X, smooth, lr, lam, lam_ridge=0.0, beta=0.1):
    """One steady-state iteration (:497-515): backward, GD, prox, zero_grad, forward, loss."""
    smooth.backward()                                                   # :497
    for param in model.parameters():                                    # :498-499
        param.data -= lr * param.grad
    if lam > 0:                                                         # :502-504, prox_update :308-314
        for net in model.networks:
            W = net.gru.weight_ih_l0
            norm = torch.norm(W, dim=0, keepdim=True)
            W.data = ((W / torch.clamp(norm, min=(lam * lr))) * torch.clamp(norm - (lr * lam), min=0.0))
    model.zero_grad()                                                   # :506
    return smooth_loss(model, X, lam_ridge, beta)                       # :508-515
