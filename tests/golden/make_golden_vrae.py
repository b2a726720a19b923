"""Golden vectors for config 4 (reference VRAE.py, generic VRAE, GRU cell, teacher forcing 1.0) from the
reference itself: vrae_generic.npz = init, one forward/backward (all gradients), and the losses + final weights
of a 6-epoch train() run.  Small shapes (B=48, T=12, D=10, H=64, Z=32) so every tensor can be committed."""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import crvae_oracle as O  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_vrae", "/root/reference/VRAE.py")
    ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
    B, T, D, H, Z = 48, 12, 10, 64, 32
    torch.manual_seed(0)
    data = torch.randn(B, T, D)
    model = ref.VRAE(D, H, Z, "gru", "tanh")
    out = {"data": data.numpy()}
    out.update({"init." + k: v.numpy() for k, v in O.gvrae_params_from_state_dict(model.state_dict()).items()})
    out["init.start_token"] = model.decoder.start_token.detach().numpy().copy()
    st = torch.get_rng_state()
    eps = torch.randn(B, Z)
    torch.set_rng_state(st)
    out["eps"] = eps.numpy()
    recon, mu, logvar = model(data, teacher_forcing_ratio=1.0)
    total, rec, kld = model.loss(recon, data, mu, logvar, 0.5)
    total.backward()
    out.update(recon=recon.detach().numpy(), mu=mu.detach().numpy(), logvar=logvar.detach().numpy(), total=float(total),
               rec=float(rec), kld=float(kld))
    out.update({"grad." + k: v.numpy() for k, v in O.gvrae_params_from_state_dict(
        {n: p.grad for n, p in model.named_parameters() if p.grad is not None}).items()})
    out["start_token_has_grad"] = model.decoder.start_token.grad is not None
    # training run (fresh model, same seeds)
    torch.manual_seed(0)
    data = torch.randn(B, T, D)
    model = ref.VRAE(D, H, Z, "gru", "tanh")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref.train(model, data, epochs=11, lr=1e-3, beta=0.5)
    txt = buf.getvalue()
    out["log_total"] = np.array([float(x) for x in re.findall(r"Total: ([-\d.]+)", txt)])
    out["log_rec"] = np.array([float(x) for x in re.findall(r"Rec: ([-\d.]+)", txt)])
    out["log_kld"] = np.array([float(x) for x in re.findall(r"KLD: ([-\d.]+)", txt)])
    out.update({"final." + k: v.numpy() for k, v in O.gvrae_params_from_state_dict(model.state_dict()).items()})
    out["rng_after"] = torch.get_rng_state().numpy()
    np.savez_compressed(os.path.join(HERE, "vrae_generic.npz"), **out)
    print("wrote vrae_generic.npz", out["log_total"], out["log_rec"], out["log_kld"], out["start_token_has_grad"])


if __name__ == "__main__":
    main()
