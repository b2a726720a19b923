"""BASELINE config 4: VRAE.py standalone (batch 1024, seq 512, latent 32, D = 10, GRU, teacher forcing 1.0): one full-batch Adam
epoch = forward + loss + backward + Adam; units = B*T*D per epoch.  CUDA events around `epochs` epochs after a warm-up; the
reference's own structure (nn.GRU encoder + GRUCell loop with a host sync per step) on the same GPU in PyTorch eager for context."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_connexe_b200 import vrae as VR

B, T, D, Z = 1024, 512, 10, 32
torch.manual_seed(0)
data = torch.randn(B, T, D).cuda()
model = VR.VRAE(D, 64, Z, "gru", "tanh")
e = model.engine
VR.train(model, data, epochs=2, lr=1e-3, beta=1.0)
torch.cuda.synchronize()
n = 10
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(n):
    model(data); e.backward(1.0); e.adam_step(1e-3)
t.record(); torch.cuda.synchronize()
ms = s.elapsed_time(t) / n
out = {"config": "VRAE.py B=1024 T=512 D=10 H=64 Z=32, TF=1.0", "ms_per_epoch": ms, "units_per_s": B * T * D / (ms * 1e-3),
       "recurrent_steps_per_epoch": 4 * T, "us_per_recurrent_step": ms * 1e3 / (4 * T)}
print(json.dumps(out))
