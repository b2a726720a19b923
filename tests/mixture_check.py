"""Shared body of the MixtureCSRAE parity test (reference CSRAE_new.py:113-150): checker backend on the CPU and CUDA kernels
on the GPU against tests/golden/mixture_csrae.npz (produced by the reference, tests/golden/make_golden_mixture.py)."""
import os

import numpy as np
import torch

from tests.conftest import GOLDEN


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(device, tol=1e-4):
    from vae_connexe_b200.mixture_csrae import MixtureCSRAE
    g = np.load(os.path.join(GOLDEN, "mixture_csrae.npz"))
    torch.manual_seed(0)
    m = MixtureCSRAE(96, (64, 48), 20, 10, 0.7)
    with torch.no_grad():
        m.theta["prior_logvar"].copy_(0.3 * torch.randn(10, 20, generator=torch.Generator().manual_seed(1)))
    sd = m.state_dict()
    keys = [k[len("init."):] for k in g.files if k.startswith("init.")]
    assert list(sd.keys()) == keys
    for k in keys:
        assert np.array_equal(sd[k].cpu().numpy(), g["init." + k]), k
    x = torch.from_numpy(g["x"]).to(device)
    torch.manual_seed(5)
    total, recon, cs = m.loss_and_grad(x)
    assert abs(float(recon) - float(g["recon"])) < tol * float(g["recon"]) and abs(float(cs) - float(g["cs"])) < tol * float(g["cs"])
    assert abs(float(total) - float(g["total"])) < tol * float(g["total"])
    gd = m.grad_dict()
    for k in keys:
        assert _rel(gd[k], g["grad." + k]) < tol, k
    torch.manual_seed(5)
    t2, r2, c2 = m.loss(x)
    assert float(t2) == float(total)
