"""Shared body of the Family-B CR-VAE parity test (reference CRVAE.py:55-199 vs vae_connexe_b200.family_b): runs on the CPU
checker backend (tests/test_host_logic.py) and on the CUDA kernels (tests/test_gpu_train.py) against the same fixture,
tests/golden/family_b.npz, produced by the reference itself (tests/golden/make_golden_family_b.py)."""
import os

import numpy as np
import torch

from tests.conftest import GOLDEN


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(device, tol=1e-4):
    from vae_connexe_b200 import family_b as FB
    g = np.load(os.path.join(GOLDEN, "family_b.npz"))
    D, Z, tau = int(g["D"]), int(g["Z"]), int(g["tau"])
    xb = torch.from_numpy(g["x"]).to(device)
    torch.manual_seed(0)
    m = FB.CRVAE(D, 64, Z, tau)
    sd = m.state_dict()
    init_keys = [k[len("init."):] for k in g.files if k.startswith("init.")]
    assert list(sd.keys()) == init_keys                                   # same names, same order as the reference's state_dict
    for k in init_keys:
        assert np.array_equal(sd[k].cpu().numpy(), g["init." + k]), k      # seed parity (declaration order of the draws)
    # one stage-1 forward / backward: loss, reconstruction and every gradient
    st = torch.get_rng_state()
    x_past, x_cur = torch.split(xb[0], tau, dim=1)
    recon, mu, ls, _, _ = m(x_past, x_cur, phase=1)
    loss = m.loss_and_backward()
    assert _rel(recon, g["step0.recon"]) < tol and abs(float(loss) - float(g["step0.loss"])) < tol * float(g["step0.loss"])
    gd = m.grad_dict()
    for k in g.files:
        if k.startswith("grad0."):
            assert _rel(gd[k[len("grad0."):]], g[k]) < tol, k
    torch.set_rng_state(st)
    # 4 stage-1 steps (Adam + ISTA), then 3 stage-2 steps (ErrorVAE embedded in the H = 64 kernels, two Adam step counters)
    tr = FB.CRVAETrainer(m, λ_l1=0.03, lr=2e-3)
    l1 = [tr.step_stage1(xb[i % 5]) for i in range(4)]
    assert np.allclose(l1, g["stage1.losses"], rtol=tol)
    sd = m.state_dict()
    for k in init_keys:
        assert _rel(sd[k], g["stage1." + k]) < 5 * tol, k
    assert np.array_equal(m.granger_matrix().cpu().numpy(), g["stage1.granger"])
    with torch.no_grad():
        for p_ in range(D):
            m.theta["W_in"][p_][(p_ + 2) % D].zero_()
    assert np.array_equal(m.granger_matrix().cpu().numpy(), g["stage2.granger_in"])
    l2 = [tr.step_stage2(xb[(i + 1) % 5]) for i in range(3)]
    assert np.allclose(l2, g["stage2.losses"], rtol=tol)
    sd = m.state_dict()
    for k in init_keys:
        assert _rel(sd[k], g["stage2." + k]) < 5 * tol, k
    assert np.array_equal(torch.get_rng_state().numpy(), g["rng_after"])
    # the zero padding of the embedded ErrorVAE stayed exactly zero (weights, gradients and Adam moments)
    th = m.theta
    assert float(th["e_enc_w_hh"][32:64].abs().sum() + th["e_enc_w_hh"][:, 32:].abs().sum() + th["e_out_w"][:, 32:].abs().sum()
                 + th["e_z2h_w"][32:].abs().sum() + th["e_lat_w"][:, 32:].abs().sum()) == 0.0
    return m
