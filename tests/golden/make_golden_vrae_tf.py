"""Golden vectors for VRAE.py with teacher forcing < 1 (VRAE.py:85-100, schedules :173-182) from the reference itself:
vrae_tf.npz = one forward/backward at ratio 0.5 (every gradient) and the per-epoch losses + final weights of a 6-epoch Adam run
under exponential_teacher_forcing_schedule(decay 0.25) followed by one epoch at ratio 0.0 (decoder starts from start_token)."""
import importlib.util, os, sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_vrae", "/root/reference/VRAE.py")
    ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
    B, T, D, H, Z = 24, 8, 10, 64, 32
    torch.manual_seed(0)
    data = torch.randn(B, T, D)
    model = ref.VRAE(D, H, Z, "gru", "tanh")
    out = {}
    torch.manual_seed(3)
    recon, mu, logvar = model(data, teacher_forcing_ratio=0.5)
    total, rec, kld = model.loss(recon, data, mu, logvar, 0.5)
    total.backward()
    out.update(recon=recon.detach().numpy(), total=float(total), rec=float(rec), kld=float(kld))
    out.update({"grad." + k: p.grad.numpy().copy() for k, p in model.named_parameters() if p.grad is not None})
    model.zero_grad(set_to_none=True)
    torch.manual_seed(4)
    optim = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses, ratios = [], []
    for epoch in range(7):
        ratio = ref.exponential_teacher_forcing_schedule(epoch, decay_rate=0.25) if epoch < 6 else 0.0
        recon, mu, logvar = model(data, teacher_forcing_ratio=ratio)
        total, rec, kld = model.loss(recon, data, mu, logvar, 0.5)
        optim.zero_grad(); total.backward(); optim.step()
        losses.append(float(total)); ratios.append(ratio)
    out["losses"], out["ratios"] = np.array(losses), np.array(ratios)
    out.update({"final." + k: v.numpy().copy() for k, v in model.state_dict().items()})
    out["rng_after"] = torch.get_rng_state().numpy()
    np.savez_compressed(os.path.join(HERE, "vrae_tf.npz"), **out)
    print("wrote vrae_tf.npz", losses, ratios)


if __name__ == "__main__":
    main()
