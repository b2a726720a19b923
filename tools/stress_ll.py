"""Stress the low-latency / MMA / exact recurrent kernels for run-to-run differences (fresh process, first launches included)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_connexe_b200 import lib
k = lib.kernels()
H, G = 64, 192
P, T, B = 1, 10, 256
g = torch.Generator().manual_seed(1)
gi = torch.randn(P, T, B, G, generator=g)
b_ih, w_hh, b_hh = torch.randn(P, G, generator=g) * 0.2, torch.randn(P, G, H, generator=g) * 0.125, torch.randn(P, G, generator=g) * 0.2
h0 = torch.randn(P, B, H, generator=g)
c = lambda t: t.cuda()
outs = {}
for name in ("gru_fwd_ll", "gru_fwd", "gru_fwd_mma"):
    fn = getattr(k, name)
    ref = None
    bad = 0
    for it in range(300):
        gates = c(gi.clone()); hs = torch.zeros(P, T, B, H, device="cuda"); ghn = torch.zeros(P, T, B, H, device="cuda")
        fn(gates, c(b_ih), c(w_hh), c(b_hh), c(h0), B * H, None, None, hs, ghn, None, P, T, B, 0)
        torch.cuda.synchronize()
        cur = (gates.cpu(), hs.cpu(), ghn.cpu())
        if ref is None:
            ref = cur
        else:
            d = max(float((a - b).abs().max()) for a, b in zip(cur, ref))
            if d != 0.0:
                bad += 1
                if bad <= 3:
                    print(name, "iteration", it, "differs from the first run by", d, flush=True)
    outs[name] = ref
    print(name, "runs differing from the first:", bad, flush=True)
for n in ("gru_fwd_ll", "gru_fwd_mma"):
    print(n, "vs exact: max abs diff", max(float((a - b).abs().max()) for a, b in zip(outs[n], outs["gru_fwd"])))

# ---- backward (encoder shape: gradient enters through h_T only) ----
gates_f = c(gi.clone()); hs = torch.zeros(P, T, B, H, device="cuda"); ghn_f = torch.zeros(P, T, B, H, device="cuda")
k.gru_fwd(gates_f, c(b_ih), c(w_hh), c(b_hh), c(h0), B * H, None, None, hs, ghn_f, None, P, T, B, 0)
dh_last = c(torch.randn(P, B, H, generator=g) * 0.1)
for name in ("gru_bwd_ll", "gru_bwd_mma", "gru_bwd_deferred"):
    fn = getattr(k, name)
    ref = None; bad = 0
    for it in range(300):
        gates, ghn = gates_f.clone(), ghn_f.clone()
        z = lambda *s: torch.zeros(*s, device="cuda")
        db_hh, db_ih, dh0, dw_hh = z(P, G), z(P, G), z(P, B, H), z(P, G, H)
        ws = torch.zeros(k.gru_bwd_workspace(P, B) // 4 + 4, device="cuda")
        ws2 = torch.zeros(k.gru_dwhh_tc_workspace(P, T, B) // 4 + 4, device="cuda")
        fn(gates, ghn, hs, c(h0), B * H, c(w_hh), None, None, dh_last, None, db_hh, db_ih, None, None, dh0, P, T, B, ws)
        k.gru_dwhh_tc(gates, ghn, hs, c(h0), B * H, dw_hh, P, T, B, ws2)
        torch.cuda.synchronize()
        cur = (gates.cpu(), ghn.cpu(), db_hh.cpu(), db_ih.cpu(), dh0.cpu(), dw_hh.cpu())
        if ref is None:
            ref = cur
        else:
            ds = [float((a - b).abs().max()) for a, b in zip(cur, ref)]
            if max(ds) != 0.0:
                bad += 1
                if bad <= 3:
                    print(name, "iteration", it, "differs:", ds, flush=True)
    print(name, "+ dwhh_tc runs differing from the first:", bad, flush=True)
