"""Golden vectors for the Family-B CR-VAE (reference CRVAE.py:55-199) from the reference itself: family_b.npz =
init state_dict of CRVAE(D=6, H=64, Z=32, tau=10) under torch.manual_seed(0), a batch of 64 Henon-like windows, the
gradients of the first stage-1 step, and losses / parameters / Granger matrix after 4 stage-1 steps followed by 3 stage-2
steps of CRVAETrainer(lam_l1=0.03, lr=2e-3) -- the reference driver's hyper-parameters (:242-243)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_loader import load_reference  # noqa: E402


def main():
    ref = load_reference("CRVAE.py")
    D, H, Z, tau, B = 6, 64, 32, 10, 64
    # a bounded multivariate series in [0, 1] (the reference's henon_map_steps overflows for most seeds): each series is
    # a lagged nonlinear function of its left neighbour plus its own oscillation
    rs = np.random.RandomState(0)
    t = np.arange(400, dtype=np.float64)
    series = np.zeros((400, D))
    series[:, 0] = 0.5 + 0.4 * np.sin(0.31 * t)
    for d in range(1, D):
        series[1:, d] = 0.5 + 0.25 * np.sin(0.17 * (d + 1) * t[1:]) + 0.2 * np.tanh(3 * (series[:-1, d - 1] - 0.5))
    series = (series + 0.01 * rs.randn(400, D)).astype(np.float32)
    wins = np.stack([series[s:s + 2 * tau] for s in range(0, 400 - 2 * tau + 1)])
    idx = np.random.RandomState(1).randint(len(wins), size=(5, B))
    xb = torch.from_numpy(wins[idx].astype(np.float32))                 # five batches [5, B, 20, D]
    torch.manual_seed(0)
    m = ref.CRVAE(D, H, Z, tau)
    out = {"x": xb.numpy(), "D": D, "Z": Z, "tau": tau}
    out.update({"init." + k: v.numpy().copy() for k, v in m.state_dict().items()})
    # gradients of the first stage-1 step (recomputed here exactly like step_stage1 does, without the update)
    st = torch.get_rng_state()
    x_past, x_cur = torch.split(xb[0], tau, dim=1)
    recon, mu, ls, *_ = m(x_past, x_cur, phase=1)
    loss = F.mse_loss(recon, x_cur) + (-0.5 * torch.mean(1 + 2 * ls - mu.pow(2) - torch.exp(2 * ls)))
    m.zero_grad(); loss.backward()
    out["step0.loss"] = float(loss); out["step0.recon"] = recon.detach().numpy().copy()
    out.update({"grad0." + k: p.grad.numpy().copy() for k, p in m.named_parameters() if p.grad is not None})
    m.zero_grad()
    torch.set_rng_state(st)
    tr = ref.CRVAETrainer(m, λ_l1=0.03, lr=2e-3)
    losses1 = [tr.step_stage1(xb[i % 5]) for i in range(4)]
    out["stage1.losses"] = np.array(losses1)
    out.update({"stage1." + k: v.numpy().copy() for k, v in m.state_dict().items()})
    out["stage1.granger"] = m.granger_matrix().numpy()
    # make the graph non-trivial for stage 2: zero a few rows of W_in (as a long stage 1 would)
    with torch.no_grad():
        for p_ in range(D):
            m.W_in[p_].data[(p_ + 2) % D].zero_()
    out["stage2.granger_in"] = m.granger_matrix().numpy()
    losses2 = [tr.step_stage2(xb[(i + 1) % 5]) for i in range(3)]
    out["stage2.losses"] = np.array(losses2)
    out.update({"stage2." + k: v.numpy().copy() for k, v in m.state_dict().items()})
    out["rng_after"] = torch.get_rng_state().numpy()
    np.savez_compressed(os.path.join(HERE, "family_b.npz"), **out)
    print("wrote family_b.npz", losses1, losses2, out["stage1.granger"].sum())


if __name__ == "__main__":
    main()
