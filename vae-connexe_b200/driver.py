"""The reference script's driver (CRVAE_lorenz96.py:730-796) as a function: data file -> phase 1 ->
`GC_lorenz96.npy` -> phase 2, with the same files, shapes and dtypes on disk, so either phase can be
run here and the other by the reference.

    python -m vae_connexe_b200.driver            # what `python CRVAE_lorenz96.py` does, without the plots

Wire formats (the reference's, unchanged):
  * `2_x.npy`        float32 (p, T) -- or (1, p, T) -- z-scored Lorenz-96 series (:731-745, :747-750)
  * `GC_lorenz96.npy` int32 (p, p)   -- `cgru.GC(threshold=True).cpu().numpy()` (:775, :787): entry (i, j) != 0
                                        iff series j Granger-causes series i.  It is fed back as the `connection`
                                        argument of the phase-2 CRVAE (:788-789), whose head i reads COLUMN i of it
                                        (:201, the reference's transposed indexing, reproduced).
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .data import lorenz_96, lorenz_96_graph
from .modules import CRVAE, VRAE4E
from .train import train_phase1, train_phase2

DATA_FILE = "2_x.npy"
GC_FILE = "GC_lorenz96.npy"


def load_or_generate_series(path: str = DATA_FILE, p: int = 10, T: int = 2048, f: float = 10.0, seed: int = 0,
                            verbose: bool = True) -> np.ndarray:
    """:731-750 -- load the series file, else generate it with the reference's generator and save it; always
    returns (1, p, T)."""
    try:
        X_np = np.load(path)
        if verbose:
            print(f"Loaded `{path}` with shape {X_np.shape}")
    except FileNotFoundError:
        X_np = lorenz_96(d=p, t=T, t_eval=0, f=f, seed=seed)
        np.save(path, X_np)
        if verbose:
            print(f"Generated and saved `{path}` with shape {X_np.shape}")
    if X_np.ndim == 2:
        X_np = X_np[np.newaxis, :, :]
    return X_np


def save_gc(path: str, crvae) -> np.ndarray:
    """:775 + :787 -- the thresholded graph as the reference writes it: int32 (p, p)."""
    gc = crvae.GC(threshold=True).cpu().numpy()
    assert gc.dtype == np.int32 and gc.shape == (crvae.p, crvae.p)
    np.save(path, gc)
    return gc


def load_gc(path: str, p: Optional[int] = None) -> np.ndarray:
    """The phase-2 `connection` (:788): any integer / bool / float (p, p) array whose non-zeros mark the edges."""
    gc = np.load(path)
    if gc.ndim != 2 or gc.shape[0] != gc.shape[1] or (p is not None and gc.shape[0] != p):
        raise ValueError(f"{path}: expected a square (p, p) graph, got {gc.shape}")
    return gc


def run(workdir: str = ".", p: int = 10, T: int = 2048, hidden: int = 64, context: int = 20, lam: float = 0.1,
        lr: float = 5e-2, max_iter_phase1: int = 5000, max_iter_phase2: int = 10000, check_every: int = 50,
        phases=(1, 2), device: Optional[str] = None, verbose: int = 1, **crvae_kw):
    """:752-796.  `phases=(1,)` stops after writing GC_lorenz96.npy, `phases=(2,)` starts from an existing file
    (possibly written by the reference).  Returns dict(GC_true, GC_est, loss_phase1, loss_phase2)."""
    dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    X_np = load_or_generate_series(os.path.join(workdir, DATA_FILE), p=p, T=T, verbose=bool(verbose))
    if X_np.shape[1] != p:
        raise ValueError(f"series file holds {X_np.shape[1]} variables, expected p={p}")
    X = torch.tensor(X_np.transpose(0, 2, 1), dtype=torch.float32, device=dev)          # (batch, T, dim), :747
    out = {"GC_true": lorenz_96_graph(p), "GC_est": None, "loss_phase1": None, "loss_phase2": None}
    gc_path = os.path.join(workdir, GC_FILE)
    if 1 in phases:
        cgru = CRVAE(p, np.ones((p, p)), hidden=hidden, **crvae_kw).to(dev)              # :767-769 (the VRAE4E built there is unused)
        out["loss_phase1"] = train_phase1(cgru, X, context=context, lam=lam, lam_ridge=0, lr=lr, max_iter=max_iter_phase1,
                                          check_every=check_every, verbose=verbose)      # :772-774
        out["GC_est"] = save_gc(gc_path, cgru)                                           # :776, :787
        if verbose:
            print("Estimated GC:\n", out["GC_est"])
    if 2 in phases:
        full_connect = load_gc(gc_path, p)                                               # :788
        out["GC_est"] = full_connect
        cgru = CRVAE(p, full_connect, hidden=hidden, **crvae_kw).to(dev)                 # :789
        vrae = VRAE4E(p, hidden=hidden).to(dev)                                          # :790
        out["loss_phase2"] = train_phase2(cgru, vrae, X, context=context, lam=0., lam_ridge=0, lr=lr,
                                          max_iter=max_iter_phase2, check_every=check_every, verbose=verbose)   # :792-794
        if verbose:
            print("Phase 2 completed!")
    return out


if __name__ == "__main__":
    run()
