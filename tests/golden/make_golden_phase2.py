"""Golden vectors for phase 2 (train_phase2, CRVAE_lorenz96.py:562-698) from the reference itself:
p10_phase2.npz = pruned CRVAE (the golden phase-1 GC as `connection`) + VRAE4E, the state before and
after the first 3 iterations (all gradients of iteration 0, Adam/GD-updated weights) and the check log
of a 21-iteration run."""
from __future__ import annotations

import contextlib
import io
import os
import re
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import crvae_oracle as O          # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402


def _np(d, prefix=""):
    return {prefix + k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def make_phase2():
    ref = load_reference()
    traj = np.load(os.path.join(HERE, "p10_traj.npz"))
    GC = traj["final_GC"].astype(int)
    X = traj["data"]
    Xt = torch.tensor(X.T[None], dtype=torch.float32)
    p, B, H = 10, 256, 64
    out = {"connection": GC}

    # (1) one hand-executed iteration with the reference's own statements (:590-603, :611-625)
    torch.manual_seed(0); np.random.seed(0)
    cg = ref.CRVAE(p, GC, H)
    vr = ref.VRAE4E(p, H)
    out.update(_np(O.params_from_state_dict(cg.state_dict(), GC), "c_init."))
    out.update(_np(O.vrae_params_from_state_dict(vr.state_dict()), "v_init."))
    wins = ref.arrange_input(Xt[0], 20)[0]
    idx = np.random.randint(len(wins), size=(B,))
    Xb = wins[idx]
    out["idx"] = idx
    st = torch.get_rng_state()
    eps_c = torch.randn(size=(1, B, H))[0]; eps_e = torch.randn(size=(1, B, H))[0]
    torch.set_rng_state(st)
    out["eps_c"], out["eps_e"] = eps_c.numpy(), eps_e.numpy()
    opt = torch.optim.Adam(vr.parameters(), lr=1e-3)
    loss_fn = torch.nn.MSELoss()
    pred, mu, log_var = cg(Xb)
    loss = sum([loss_fn(pred[i][:, :, 0], Xb[:, 10:, i]) for i in range(p)])
    mmd = (-0.5 * (1 + log_var - mu ** 2 - torch.exp(log_var)).sum(dim=-1).sum(dim=0)).mean(dim=0)
    smooth = loss + 1 * mmd
    error = (-torch.stack(pred)[:, :, :, 0].permute(1, 2, 0) + Xb[:, 10:, :]).detach()
    pred_e, mu_e, log_var_e = vr(error)
    loss_e = loss_fn(pred_e, error)
    mmd_e = (-0.5 * (1 + log_var_e - mu_e ** 2 - torch.exp(log_var_e)).sum(dim=-1).sum(dim=0)).mean(dim=0)
    smooth_e = loss_e + 1 * mmd_e
    out.update(loss=float(loss), kl=float(mmd), loss_e=float(loss_e), kl_e=float(mmd_e), error=error.numpy(),
               pred_e=pred_e.detach().numpy())
    smooth_e.backward()
    out.update(_np(O.vrae_params_from_state_dict({k: v.grad for k, v in vr.named_parameters()}), "v_grad."))
    opt.step(); opt.zero_grad()
    smooth.backward()
    out.update(_np({k: v for k, v in O.params_from_state_dict({k: v.grad for k, v in cg.named_parameters()}, GC).items()
                    if k != "mask"}, "c_grad."))
    for param in cg.parameters():
        param.data -= 5e-2 * param.grad
    out.update(_np(O.params_from_state_dict(cg.state_dict(), GC), "c_post."))
    out.update(_np(O.vrae_params_from_state_dict(vr.state_dict()), "v_post."))

    # (2) the reference's train_phase2 for 21 iterations (check_every=10): its printed log + final weights
    torch.manual_seed(0); np.random.seed(0)
    cg = ref.CRVAE(p, GC, H)
    vr = ref.VRAE4E(p, H)
    buf = io.StringIO()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)          # the reference writes ori_henon.npy / syn_henon.npy into cwd (:692-693)
        try:
            with contextlib.redirect_stdout(buf):
                ref.train_phase2(cg, vr, Xt, context=20, lam=0., lam_ridge=0, lr=5e-2, max_iter=21, check_every=10)
        finally:
            os.chdir(cwd)
    txt = buf.getvalue()
    out["log_it"] = np.array([int(x) for x in re.findall(r"Iter = (\d+)", txt)])
    out["log_loss"] = np.array([float(x) for x in re.findall(r"\nLoss = ([-\d.eE+]+)", txt)])
    out["log_kl"] = np.array([float(x) for x in re.findall(r"\nKL = ([-\d.eE+]+)", txt)])
    out["log_loss_e"] = np.array([float(x) for x in re.findall(r"Loss_e = ([-\d.eE+]+)", txt)])
    out["log_kl_e"] = np.array([float(x) for x in re.findall(r"KL_e = ([-\d.eE+]+)", txt)])
    out.update(_np(O.params_from_state_dict(cg.state_dict(), GC), "c_final."))
    out.update(_np(O.vrae_params_from_state_dict(vr.state_dict()), "v_final."))
    out["rng_after"] = torch.get_rng_state().numpy()
    out["np_rng_after_draw"] = np.random.randint(1 << 30)
    np.savez_compressed(os.path.join(HERE, "p10_phase2.npz"), **out)
    print("wrote p10_phase2.npz", out["log_it"], out["log_loss"], out["log_loss_e"], out["log_kl_e"])


if __name__ == "__main__":
    make_phase2()
