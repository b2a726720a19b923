"""Timing of the recurrent kernels at SMALL head counts (CUDA events, L2-cold between launches): exact FFMA (gru_fwd /
gru_bwd_deferred), tcgen05 (gru_fwd_tc / gru_bwd_tc) and low-latency (gru_fwd_ll / gru_bwd_ll)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_connexe_b200 import lib


def main():
    k = lib.kernels()
    H, G = 64, 192
    cases = [(1, 10, 256), (13, 10, 256), (25, 10, 256), (50, 10, 256), (100, 10, 256), (1, 512, 1024)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for P, T, B in cases:
        g = torch.Generator(device="cuda").manual_seed(0)
        gi = torch.randn(P, T, B, G, device="cuda", generator=g)
        w_hh = torch.randn(P, G, H, device="cuda", generator=g) * 0.125
        b_ih, b_hh = torch.randn(P, G, device="cuda", generator=g) * 0.2, torch.randn(P, G, device="cuda", generator=g) * 0.2
        h0 = torch.randn(B, H, device="cuda", generator=g)
        w_lin, b_lin = torch.randn(P, H, device="cuda", generator=g) * 0.2, torch.randn(P, device="cuda", generator=g)
        hs, ghn, pred = torch.zeros(P, T, B, H, device="cuda"), torch.zeros(P, T, B, H, device="cuda"), torch.zeros(P, T, B, device="cuda")
        n = 3 if T > 100 else 8

        def run(fn):
            ts = []
            for _ in range(n):
                gates = gi.clone()
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(gates); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            return ts[len(ts) // 2]
        res = {}
        res["fwd ffma"] = run(lambda gates: k.gru_fwd(gates, b_ih, w_hh, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0))
        res["fwd ll"] = run(lambda gates: k.gru_fwd_ll(gates, b_ih, w_hh, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0))
        res["fwd mma"] = run(lambda gates: k.gru_fwd_mma(gates, b_ih, w_hh, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0))
        if T <= 100:
            res["fwd tc"] = run(lambda gates: k.gru_fwd_tc(gates, b_ih, w_hh, None, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0))
        dpred = torch.randn(P, T, B, device="cuda", generator=g)
        gr = lambda *sh: torch.zeros(*sh, device="cuda")
        db_hh, db_ih, dw_lin, db_lin, dh0 = gr(P, G), gr(P, G), gr(P, H), gr(P), gr(P, B, H)
        ws = torch.zeros(k.gru_bwd_workspace(P, B) // 4 + 4, device="cuda")
        gates_f = gi.clone(); k.gru_fwd(gates_f, b_ih, w_hh, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0)

        def run_b(fn):
            ts = []
            for _ in range(n):
                gates, gh = gates_f.clone(), ghn.clone()
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(gates, gh); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            return ts[len(ts) // 2]
        res["bwd ffma"] = run_b(lambda gates, gh: k.gru_bwd_deferred(gates, gh, hs, h0, 0, w_hh, w_lin, dpred, None, None, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws))
        res["bwd ll"] = run_b(lambda gates, gh: k.gru_bwd_ll(gates, gh, hs, h0, 0, w_hh, w_lin, dpred, None, None, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws))
        res["bwd mma"] = run_b(lambda gates, gh: k.gru_bwd_mma(gates, gh, hs, h0, 0, w_hh, w_lin, dpred, None, None, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws))
        if T <= 100:
            res["bwd tc"] = run_b(lambda gates, gh: k.gru_bwd_tc(gates, gh, hs, h0, 0, w_hh, w_lin, dpred, None, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws))
        print(f"P={P:4d} T={T:4d} B={B:5d}  " + "  ".join(f"{k_} {v:8.1f} us ({v / T:5.2f}/step)" for k_, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
