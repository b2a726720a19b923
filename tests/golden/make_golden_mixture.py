"""Golden vectors for MixtureCSRAE (reference CSRAE_new.py:113-150) from the reference itself: mixture_csrae.npz = init
state_dict under torch.manual_seed(0) of MixtureCSRAE(input_dim=96, hidden_dims=(64, 48), latent_dim=20, K=10,
lambda_cs=0.7), a batch of 40 binary-ish inputs, the three loss terms and every autograd gradient of loss(x)."""
import os, sys
import numpy as np
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import importlib.util


def main():
    spec = importlib.util.spec_from_file_location("ref_csrae_new", "/root/reference/CSRAE_new.py")
    ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
    torch.manual_seed(0)
    m = ref.MixtureCSRAE(96, (64, 48), 20, 10, 0.7)
    with torch.no_grad():                                   # a non-trivial prior (logvar is initialised to zero)
        m.prior.logvar.copy_(0.3 * torch.randn(10, 20, generator=torch.Generator().manual_seed(1)))
    out = {"init." + k: v.numpy().copy() for k, v in m.state_dict().items()}
    x = (torch.rand(40, 96, generator=torch.Generator().manual_seed(2)) < 0.4).float()
    out["x"] = x.numpy()
    torch.manual_seed(5)
    total, recon, cs = m.loss(x)
    total.backward()
    out.update(total=float(total), recon=float(recon), cs=float(cs))
    out.update({"grad." + k: p.grad.numpy().copy() for k, p in m.named_parameters()})
    np.savez_compressed(os.path.join(HERE, "mixture_csrae.npz"), **out)
    print("wrote mixture_csrae.npz", float(total), float(recon), float(cs))


if __name__ == "__main__":
    main()
