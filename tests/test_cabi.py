"""The C-ABI shared library loads and exports every symbol include/crvae_b200.h declares.
No compute calls here (no GPU needed)."""
import ctypes
import os
import re

import pytest

from tests.conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "crvae_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(crvae_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for must in ("crvae_proj_fwd", "crvae_proj_wgrad", "crvae_gru_fwd", "crvae_gru_bwd", "crvae_latent_fwd",
                 "crvae_latent_bwd", "crvae_mse_fwd_bwd", "crvae_gd_prox_gc", "crvae_gd_step", "crvae_adam_step"):
        assert must in syms


def test_library_builds_loads_and_exports_every_declared_symbol():
    import vae_connexe_b200.lib as L
    lib = L.load()
    assert lib.crvae_abi_version() == 1
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in L.SIGNATURES, f"{s} has no ctypes signature"
    for s in L.SIGNATURES:
        assert s in _declared_symbols(), f"{s} bound but not declared in the header"


def test_library_is_plain_c_abi():
    """No torch / C++ types at the boundary: every exported crvae_* symbol is unmangled."""
    import subprocess
    import vae_connexe_b200.build as B
    out = subprocess.run(["nm", "-D", "--defined-only", B.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(not s.startswith("_Z") or "crvae" in s for s in exported)
    for s in _declared_symbols():
        assert s in exported


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import vae_connexe_b200 as V
    import vae_connexe_b200.lib as L
    assert L._kernels is None or getattr(L._kernels, "device_type", "cuda") == "cuda"
    with pytest.raises(L.CrvaeLibraryError):
        V.CRVAE(4, np.ones((4, 4)), 64)
