"""Per-warp cycle breakdown of the ping-pong MMA recurrent kernels (library built with -DCRVAE_MMA_TIMING)."""
import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_connexe_b200 import lib
k = lib.kernels()
H, G = 64, 192
P, T, B = int(os.environ.get("P", 100)), int(os.environ.get("T", 10)), int(os.environ.get("B", 256))
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
names = ["M clk", "P clk", "barrier", "plumb", "#M", "#P", "slots", "total"]


def dump(tag):
    out = (ctypes.c_longlong * 128)()
    k.lib.crvae_debug_mma_timing(out)
    print(tag)
    for w in range(16):
        v = [out[w * 8 + i] for i in range(8)]
        n = max(v[4], 1)
        if w < 8:
            print(f"  M warp {w}: wait h_ready {v[0] / n:6.0f}  mma+handoff {v[1] / n:6.0f}  clk/op; ops {v[4]}  total {v[7]} = {v[7] / n:.0f} /op")
        else:
            print(f"  P warp {w}: wait slab {v[0] / n:5.0f}  wait acc {v[1] / n:6.0f}  gate math {v[2] / n:6.0f}  P-sync {v[3] / n:5.0f}  after {v[5] / n:5.0f}"
                  f"  clk/op; ops {v[4]}  total {v[7]} = {v[7] / n:.0f} /op")


for _ in range(2):
    gates = gi.clone()
    k.gru_fwd_mma(gates, b_ih, w_hh, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0)
torch.cuda.synchronize()
dump("forward")
