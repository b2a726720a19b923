"""One launch of each MMA recurrent kernel for an ncu capture (default P = 1, T = 64, B = 1024: 64 CTAs, one per SM)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_connexe_b200 import lib
k = lib.kernels()
H, G = 64, 192
P, T, B = int(os.environ.get("P", 1)), int(os.environ.get("T", 64)), int(os.environ.get("B", 1024))
g = torch.Generator(device="cuda").manual_seed(0)
gi = torch.randn(P, T, B, G, device="cuda", generator=g)
w_hh = torch.randn(P, G, H, device="cuda", generator=g) * 0.125
b_ih, b_hh = torch.randn(P, G, device="cuda", generator=g) * 0.2, torch.randn(P, G, device="cuda", generator=g) * 0.2
h0 = torch.randn(B, H, device="cuda", generator=g)
w_lin, b_lin = torch.randn(P, H, device="cuda", generator=g) * 0.2, torch.randn(P, device="cuda", generator=g)
hs, ghn, pred = torch.zeros(P, T, B, H, device="cuda"), torch.zeros(P, T, B, H, device="cuda"), torch.zeros(P, T, B, device="cuda")
dpred = torch.randn(P, T, B, device="cuda", generator=g)
gr = lambda *sh: torch.zeros(*sh, device="cuda")
db_hh, db_ih, dw_lin, db_lin, dh0 = gr(P, G), gr(P, G), gr(P, H), gr(P), gr(P, B, H)
ws = torch.zeros(k.gru_bwd_workspace(P, B) // 4 + 4, device="cuda")
for _ in range(2):
    gates = gi.clone()
    k.gru_fwd_mma(gates, b_ih, w_hh, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0)
    k.gru_bwd_mma(gates, ghn, hs, h0, 0, w_hh, w_lin, dpred, None, None, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws)
torch.cuda.synchronize()
print("ok")
