"""Synthetic Lorenz-96 data with the reference generator's semantics (CRVAE_lorenz96.py:700-728):
same ODE, same integrator (scipy.integrate.odeint), same noise / burn-in / per-variable z-score and
the same use of numpy's global RNG -- only the right-hand side is vectorised with np.roll, which is
element-wise the same float64 expression as the reference's Python loop (:705)."""
from __future__ import annotations

import numpy as np


def _lorenz_rhs(x, t, F):
    return (np.roll(x, -1) - np.roll(x, 2)) * np.roll(x, 1) - x + F


def lorenz_96(d, t, t_eval, f, seed, delta_t=0.1, sd=0.1, burn_in=1000):
    """Returns (d, t) float32, like the reference (:708-728)."""
    from scipy.integrate import odeint
    if seed is not None:
        np.random.seed(seed)
    x0 = np.random.normal(scale=0.01, size=d)
    tm = np.linspace(0, (t + t_eval + burn_in) * delta_t, t + t_eval + burn_in)
    X = odeint(_lorenz_rhs, x0, tm, args=(f,))
    X += np.random.normal(scale=sd, size=(t + t_eval + burn_in, d))
    X = X[burn_in:]
    X = (X - X.mean(axis=0, keepdims=True)) / (X.std(axis=0, keepdims=True) + 1e-8)
    return X.T.astype(np.float32)


def lorenz_96_graph(p):
    """Ground-truth Granger graph of Lorenz-96: i <- {i, i-1, i-2, i+1} (:757-764)."""
    gc = np.zeros((p, p), dtype=int)
    for i in range(p):
        for j in (i, (i - 1) % p, (i - 2) % p, (i + 1) % p):
            gc[i, j] = 1
    return gc
