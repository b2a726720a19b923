"""Golden vectors for the CS-RAE variant (reference CR-CS-RAE.py) from the reference itself:
cs_p10.npz = CRVAE(p, ones, 64, K=10, lambda_cs=0.1) init, a standalone cs_divergence_gmm evaluation with its
autograd gradients, and the check log + final weights of an 11-iteration train_phase1 run."""
from __future__ import annotations

import contextlib
import io
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import crvae_oracle as O          # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402


def main():
    ref = load_reference("CR-CS-RAE.py")
    traj = np.load(os.path.join(HERE, "p10_traj.npz"))
    X = traj["data"]
    Xt = torch.tensor(X.T[None], dtype=torch.float32)
    p, H, K = 10, 64, 10
    conn = np.ones((p, p))
    out = {}
    # (1) the divergence itself on random inputs (heterogeneous prior variances, some clamped samples)
    g = torch.Generator().manual_seed(3)
    B = 48
    lat = (torch.randn(B, 2 * H, generator=g) * 0.3).requires_grad_(True)
    pm = (torch.randn(K, H, generator=g) * 0.3).requires_grad_(True)
    pl = (torch.randn(K, H, generator=g) * 0.2).requires_grad_(True)
    cs = ref.cs_divergence_gmm(lat[:, H:], torch.exp(lat[:, :H]), pm, pl.exp())
    (0.1 * cs.mean()).backward()
    out.update(cs_lat=lat.detach().numpy(), cs_pm=pm.detach().numpy(), cs_pl=pl.detach().numpy(), cs_vals=cs.detach().numpy(),
               cs_dlat=lat.grad.numpy(), cs_dpm=pm.grad.numpy(), cs_dpl=pl.grad.numpy())
    # (2) model init + short training run
    torch.manual_seed(0); np.random.seed(0)
    m = ref.CRVAE(p, conn, H, K, 0.1)
    out.update({"init." + k: v.numpy() for k, v in O.params_from_state_dict(
        {k: v for k, v in m.state_dict().items() if not k.startswith("prior")}, conn).items()})
    out["init.prior_mu"], out["init.prior_logvar"] = m.prior.mu.detach().numpy().copy(), m.prior.logvar.detach().numpy().copy()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref.train_phase1(m, Xt, context=20, lam=0.5, lam_ridge=0.01, lr=5e-2, max_iter=11, check_every=5, batch_size=128,
                         lambda_cs=0.1)
    txt = buf.getvalue()
    out["log_it"] = np.array([int(x) for x in re.findall(r"Iter = (\d+)", txt)])
    out["log_mean"] = np.array([float(x) for x in re.findall(r"Mean Loss = ([-\d.eE+]+)", txt)])
    out["log_recon"] = np.array([float(x) for x in re.findall(r"Recon Loss = ([-\d.eE+]+)", txt)])
    out["log_cs"] = np.array([float(x) for x in re.findall(r"CS_Div = ([-\d.eE+]+)", txt)])
    out["log_usage"] = np.array([float(x) for x in re.findall(r"usage = ([\d.]+)%", txt)])
    out.update({"final." + k: v.numpy() for k, v in O.params_from_state_dict(
        {k: v for k, v in m.state_dict().items() if not k.startswith("prior")}, conn).items()})
    out["final.prior_mu"], out["final.prior_logvar"] = m.prior.mu.detach().numpy().copy(), m.prior.logvar.detach().numpy().copy()
    out["rng_after"] = torch.get_rng_state().numpy()
    # (3) the "group-lasso prox sweep over lambda" of BASELINE config 5: the same 11-iteration run for every lambda
    sweep = [0.05, 0.1, 0.2, 0.5, 1.0]
    out["sweep_lams"] = np.array(sweep)
    for i, lam in enumerate(sweep):
        torch.manual_seed(0); np.random.seed(0)
        m = ref.CRVAE(p, conn, H, K, 0.1)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref.train_phase1(m, Xt, context=20, lam=lam, lam_ridge=0.01, lr=5e-2, max_iter=11, check_every=5, batch_size=128,
                             lambda_cs=0.1)
        txt = buf.getvalue()
        out[f"sweep{i}.log_mean"] = np.array([float(x) for x in re.findall(r"Mean Loss = ([-\d.eE+]+)", txt)])
        out[f"sweep{i}.log_cs"] = np.array([float(x) for x in re.findall(r"CS_Div = ([-\d.eE+]+)", txt)])
        out[f"sweep{i}.log_usage"] = np.array([float(x) for x in re.findall(r"usage = ([\d.]+)%", txt)])
        prm = O.params_from_state_dict({k: v for k, v in m.state_dict().items() if not k.startswith("prior")}, conn)
        out[f"sweep{i}.final_w_ih"] = prm["w_ih"].numpy()
        out[f"sweep{i}.final_enc_w_hh"] = prm["enc_w_hh"].numpy()
        out[f"sweep{i}.final_GC"] = m.GC().numpy()
    np.savez_compressed(os.path.join(HERE, "cs_p10.npz"), **out)
    print("wrote cs_p10.npz", out["log_it"], out["log_mean"], out["log_cs"], "clamped", int((cs == 0).sum()))


if __name__ == "__main__":
    main()
