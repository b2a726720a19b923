"""Generate the committed golden vectors by RUNNING THE REFERENCE ITSELF on this container's CPU.

    python tests/golden/make_golden.py step      # p4_step.npz        (seconds)
    python tests/golden/make_golden.py traj      # p10_traj.npz       (~20 min: 5000 iterations)
    python tests/golden/make_golden.py phase2    # p10_phase2_step.npz (seconds)

Needs /root/reference (build container only); the fixtures travel, this script's inputs do not.
Everything is loaded through oracle/ref_loader.py (definitions only: the reference's module-level
training driver, CRVAE_lorenz96.py:730-796, is never executed).  torch 2.11.0+cu128 CPU, seeds as
stated in SURVEY.md 8(d).
"""
from __future__ import annotations

import contextlib
import hashlib
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


def _np(d, prefix=""):
    return {prefix + k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
            for k, v in d.items()}


def lorenz_data(ref, d, t):
    X = ref.lorenz_96(d=d, t=t, t_eval=0, f=10.0, seed=0)     # (p, T) float32
    return X


def make_step():
    """One reference iteration at p=4, B=64 (small enough to commit every tensor)."""
    ref = load_reference()
    p, B, H = 4, 64, 64
    lr, lam, lam_ridge, beta = 5e-2, 0.1, 0.01, 0.1
    X = lorenz_data(ref, p, 200)
    Xt = torch.tensor(X.T, dtype=torch.float32)
    torch.manual_seed(0); np.random.seed(0)
    m = ref.CRVAE(p, np.ones((p, p)), H)
    conn = np.ones((p, p))
    wins = ref.arrange_input(Xt, 20)[0]
    idx = np.random.randint(len(wins), size=(B,))
    Xb = wins[idx]
    out = {"X": Xb.numpy(), "lr": lr, "lam": lam, "lam_ridge": lam_ridge, "beta": beta}
    out.update(_np(O.params_from_state_dict(m.state_dict(), conn), "init."))
    st = torch.get_rng_state()
    eps = torch.randn(size=(1, B, H))[0]
    torch.set_rng_state(st)
    # -- exactly the statements of train_phase1 (:482-489, :497-506), executed on the reference model
    pred, mu, log_var = m(Xb)                                   # NB swapped names, as in :482
    loss_fn = torch.nn.MSELoss()
    loss = sum([loss_fn(pred[i][:, :, 0], Xb[:, 10:, i]) for i in range(p)])
    mmd = (-0.5 * (1 + log_var - mu ** 2 - torch.exp(log_var)).sum(dim=-1).sum(dim=0)).mean(dim=0)
    ridge = sum([ref.ridge_regularize(net, lam_ridge) for net in m.networks])
    smooth = loss + ridge + beta * mmd
    smooth.backward()
    grads = O.params_from_state_dict({k: v.grad for k, v in m.named_parameters()}, conn)
    for param in m.parameters():
        param.data -= lr * param.grad
    pre_prox = O.params_from_state_dict(m.state_dict(), conn)["w_ih"]
    for net in m.networks:
        ref.prox_update(net, lam, lr)
    out.update(eps=eps.numpy(), fc_std_out=mu[0].detach().numpy(), fc_mu_out=log_var[0].detach().numpy(),
               pred=torch.stack(pred)[..., 0].permute(0, 2, 1).detach().numpy(),   # [P,Td,B]
               loss=float(loss), kl=float(mmd), ridge=float(ridge), smooth=float(smooth),
               pre_prox_w_ih=pre_prox.numpy(), GC=m.GC().numpy(), GC_norm=m.GC(False).detach().numpy())
    out.update(_np({k: v for k, v in grads.items() if k != "mask"}, "grad."))
    out.update(_np(O.params_from_state_dict(m.state_dict(), conn), "post."))
    np.savez_compressed(os.path.join(HERE, "p4_step.npz"), **out)
    print("wrote p4_step.npz", {k: np.asarray(v).shape for k, v in out.items() if "w_ih" in k})


def make_traj(max_iter=5000, p=10, T=1000, ckpt_its=(150, 200, 250), name="p10_traj.npz"):
    """The survey's golden run: train_phase1 on Lorenz-96 p=10 (SURVEY.md 8(d) cfg 1)."""
    ref = load_reference()
    X = lorenz_data(ref, p, T)
    Xt = torch.tensor(X.T[None], dtype=torch.float32)
    torch.manual_seed(0); np.random.seed(0)
    conn = np.ones((p, p))
    m = ref.CRVAE(p, conn, 64)
    out = {"data": X, "max_iter": max_iter}
    out.update(_np(O.params_from_state_dict(m.state_dict(), conn), "init."))
    st = np.random.get_state()
    out["idx"] = np.random.randint(T - 20, size=(256,))
    np.random.set_state(st)
    tst = torch.get_rng_state()
    out["eps0"] = torch.randn(size=(1, 256, 64))[0].numpy()
    torch.set_rng_state(tst)

    # hook: MinMaxScaler is called twice at the very end of every check block (:554-555)
    calls = {"n": 0}
    gcs, ckpts = [], {}
    orig = ref.MinMaxScaler
    # RNG state in front of each of the last three forwards: [:508 train, :522 check, :550 test]
    from collections import deque
    pre_states = deque(maxlen=3)
    m.register_forward_pre_hook(lambda mod, args: pre_states.append(torch.get_rng_state().clone()))

    def hook(data):
        if calls["n"] % 2 == 0:
            it = 50 * (calls["n"] // 2)
            gcs.append(m.GC().numpy().astype(np.int8))
            if it in ckpt_its:
                ckpts[it] = O.params_from_state_dict(m.state_dict(), conn)
                # generator state just before the :508 forward of iteration `it`: a resumed run
                # re-draws that eps, then the check-block draws, and continues in lock step
                ckpts[it]["_torch_rng"] = pre_states[0].clone()
        calls["n"] += 1
        return orig(data)

    ref.train_phase1.__globals__["MinMaxScaler"] = hook
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref.train_phase1(m, Xt, context=20, lam=0.1, lam_ridge=0, lr=5e-2, max_iter=max_iter,
                         check_every=50)
    txt = buf.getvalue()
    its = [int(x) for x in re.findall(r"Iter = (\d+)", txt)]
    losses = [float(x) for x in re.findall(r"Loss = ([-\d.eE+naninf]+)", txt)]
    kls = [float(x) for x in re.findall(r"KL = ([-\d.eE+naninf]+)", txt)]
    usage = [float(x) for x in re.findall(r"usage = ([\d.]+)%", txt)]
    out.update(log_it=np.array(its), log_loss=np.array(losses), log_kl=np.array(kls),
               log_usage=np.array(usage), log_gc=np.stack(gcs))
    best = np.minimum.accumulate(np.array(losses))
    out["best_it"] = its[int(np.argmin(losses))]     # printed to 6 decimals; ties -> first
    out.update(_np(O.params_from_state_dict(m.state_dict(), conn), "final."))
    gc = m.GC().numpy().astype(np.int32)
    out["final_GC"] = gc
    out["final_GC_sha256"] = hashlib.sha256(gc.tobytes()).hexdigest()
    for it, c in ckpts.items():
        out.update(_np(c, f"ckpt{it}."))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, "best_it", out["best_it"], "sha", out["final_GC_sha256"])
    print("usage tail", usage[-5:], "loss tail", losses[-3:], "best running", best[-1])


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "step"
    if what == "step":
        make_step()
    elif what == "traj":
        make_traj()
    elif what == "traj_short":
        make_traj(max_iter=301, name="p10_traj_short.npz")
    elif what == "phase2":
        from make_golden_phase2 import make_phase2   # noqa
        make_phase2()
    else:
        raise SystemExit(f"unknown target {what}")
