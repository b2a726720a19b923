"""Golden vectors for test-mode generation from the reference itself (CRVAE_lorenz96.py:223-243, :264-284, VRAE4E
:171-179): gen_p8.npz = the weights of a seeded CRVAE / VRAE4E pair, the inputs, and the sequences the reference's own
`forward(mode='test')` generates (CRVAE phase 0, VRAE4E, CRVAE phase 1 fed with the VRAE4E sample), each call made right
after `torch.manual_seed(s)` so that the consumer can reproduce the h_0 draw on the CPU generator."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_loader import load_reference  # noqa: E402


def main():
    ref = load_reference()
    out = {}
    for tag, p, B in (("a", 8, 32), ("b", 100, 256)):
        conn = np.ones((p, p))
        torch.manual_seed(11)
        rm, rv = ref.CRVAE(p, conn, 64), ref.VRAE4E(p, 64)
        g = torch.Generator().manual_seed(5)
        X, err = torch.randn(B, 20, p, generator=g), torch.randn(B, 10, p, generator=g)
        with torch.no_grad():
            torch.manual_seed(21); s0 = rm(X, mode="test")
            torch.manual_seed(22); s1 = rv(err, mode="test")
            torch.manual_seed(23); s2 = rm(X, s1[:, 1:], mode="test", phase=1)
        # inputs are re-created by the consumer from Generator().manual_seed(5); of the big case only the first 16 batch
        # rows are kept (rows are independent of each other)
        keep = slice(None) if tag == "a" else slice(0, 16)
        out.update({f"{tag}.p": p, f"{tag}.B": B,
                    f"{tag}.gen_phase0": s0[keep].numpy(), f"{tag}.gen_vrae": s1[keep].numpy(), f"{tag}.gen_phase1": s2[keep].numpy()})
        if tag == "a":          # small case: ship the weights; the big case re-creates them from the seed (bit-identical init)
            out.update({f"{tag}.crvae." + k: v.numpy() for k, v in rm.state_dict().items()})
            out.update({f"{tag}.vrae." + k: v.numpy() for k, v in rv.state_dict().items()})
    np.savez_compressed(os.path.join(HERE, "gen_p8.npz"), **out)
    print("wrote gen_p8.npz", {k: v.shape for k, v in out.items() if "gen" in k})


if __name__ == "__main__":
    main()
