"""Shared body of the teacher-forcing < 1 parity test for the generic VRAE (reference VRAE.py:85-100): checker backend on
the CPU and CUDA kernels on the GPU against tests/golden/vrae_tf.npz (tests/golden/make_golden_vrae_tf.py)."""
import os

import numpy as np
import torch

from tests.conftest import GOLDEN


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(device, tol=1e-4):
    from vae_connexe_b200 import vrae as VR
    g = np.load(os.path.join(GOLDEN, "vrae_tf.npz"))
    B, T, D, Z = 24, 8, 10, 32
    torch.manual_seed(0)
    data = torch.randn(B, T, D)
    model = VR.VRAE(D, 64, Z, "gru", "tanh")
    e = model.engine
    torch.manual_seed(3)
    recon, mu, logvar = model(data.to(device), teacher_forcing_ratio=0.5)
    assert e.free is not None and not all(e.free["use_tf"]) and any(e.free["use_tf"])       # a mixed sequence
    assert _rel(recon, g["recon"]) < tol
    e.backward_free(0.5)
    rec, kld = float(e.sse) / B, float(e.kl)
    assert abs(rec - float(g["rec"])) < tol * float(g["rec"]) and abs(kld - float(g["kld"])) < tol * float(g["kld"])
    sd_names = {"encoder.rnn.weight_ih_l0": "enc_w_ih", "encoder.rnn.weight_hh_l0": "enc_w_hh", "encoder.rnn.bias_ih_l0": "enc_b_ih",
                "encoder.rnn.bias_hh_l0": "enc_b_hh", "decoder.fc_z2h.weight": "z2h_w", "decoder.fc_z2h.bias": "z2h_b",
                "decoder.cell.weight_ih": "dec_w_ih", "decoder.cell.weight_hh": "dec_w_hh", "decoder.cell.bias_ih": "dec_b_ih",
                "decoder.cell.bias_hh": "dec_b_hh", "decoder.fc_out.weight": "out_w", "decoder.fc_out.bias": "out_b"}
    for ref_name, ours in sd_names.items():
        assert _rel(e.grad[ours], g["grad." + ref_name]) < tol, ref_name
    assert _rel(e.grad["lat_w"][:Z], g["grad.encoder.fc_mu.weight"]) < tol and _rel(e.grad["lat_w"][Z:], g["grad.encoder.fc_logvar.weight"]) < tol
    assert "grad.decoder.start_token" not in g.files                                        # ratio > 0: the start token is not used
    # scheduled training: 6 epochs of exponential decay, then one epoch at ratio 0 (decoder starts from start_token)
    torch.manual_seed(4)
    losses = []
    for epoch in range(7):
        ratio = VR.exponential_teacher_forcing_schedule(epoch, decay_rate=0.25) if epoch < 6 else 0.0
        assert abs(ratio - float(g["ratios"][epoch])) < 1e-12
        model(data.to(device), teacher_forcing_ratio=ratio)
        (e.backward_free if e.free is not None else e.backward)(0.5)
        e.adam_step(1e-3)
        losses.append(float(e.sse) / B + 0.5 * float(e.kl))
    assert np.allclose(losses, g["losses"], rtol=tol)
    sd = model.state_dict()
    for k in sd:
        assert _rel(sd[k], g["final." + k]) < 5 * tol, k
    assert np.array_equal(torch.get_rng_state().numpy(), g["rng_after"])
