"""vae-connexe_b200 -- B200-native (sm_100a) implementation of the CR-VAE training hot path of
anonyme-Zheng/VAE-connexe (CRVAE_lorenz96.py), behind the reference's own Python module API.

    from vae_connexe_b200 import CRVAE, VRAE4E, train_phase1, train_phase2, prox_update, ...

(the directory name carries a hyphen; the importable alias `vae_connexe_b200` is provided by the
shim module vae_connexe_b200.py at the repository root).
"""
from .functional import arrange_input, prox_update, regularize, restore_parameters, ridge_regularize
from .modules import CRVAE, GRU, VRAE4E
from .sharding import allgather_rows, head_range
from . import driver, family_b
from .train import Phase1Runner, Phase2Runner, train_phase1, train_phase2

__all__ = ["CRVAE", "GRU", "VRAE4E", "train_phase1", "train_phase2", "Phase1Runner", "Phase2Runner", "prox_update", "regularize", "ridge_regularize",
           "restore_parameters", "arrange_input", "head_range", "allgather_rows"]
