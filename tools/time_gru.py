"""Scratch timing of the recurrent kernels for a list of head counts (CUDA events, L2-cold between launches)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_connexe_b200 import lib

def main():
    k = lib.kernels()
    H, G, T, B = 64, 192, 10, int(os.environ.get("B", 256))
    Ps = [int(x) for x in sys.argv[1:]] or [26, 74, 100, 148]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for P in Ps:
        g = torch.Generator(device="cuda").manual_seed(0)
        gi = torch.randn(P, T, B, G, device="cuda", generator=g)
        w_hh = torch.randn(P, G, H, device="cuda", generator=g) * 0.125
        b_ih, b_hh = torch.randn(P, G, device="cuda", generator=g) * 0.2, torch.randn(P, G, device="cuda", generator=g) * 0.2
        h0 = torch.randn(B, H, device="cuda", generator=g)
        w_lin, b_lin = torch.randn(P, H, device="cuda", generator=g) * 0.2, torch.randn(P, device="cuda", generator=g)
        hs, ghn, pred = torch.zeros(P, T, B, H, device="cuda"), torch.zeros(P, T, B, H, device="cuda"), torch.zeros(P, T, B, device="cuda")
        wh, wl = torch.empty_like(w_hh), torch.empty_like(w_hh)
        k.split_tf32(w_hh, wh, wl, w_hh.numel())
        def run(fn, n=8):
            ts = []
            for _ in range(n):
                gates = gi.clone()
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(gates); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            return ts[len(ts) // 2]
        t_ffma = run(lambda gates: k.gru_fwd(gates, b_ih, w_hh, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0))
        t_tc = run(lambda gates: k.gru_fwd_tc(gates, b_ih, wh, wl, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0))
        # backward (deferred dW_hh): FFMA vs tcgen05
        dpred = torch.randn(P, T, B, device="cuda", generator=g)
        gr = lambda *sh: torch.zeros(*sh, device="cuda")
        db_hh, db_ih, dw_lin, db_lin, dh0 = gr(P, G), gr(P, G), gr(P, H), gr(P), gr(P, B, H)
        ws = torch.zeros(k.gru_bwd_workspace(P, B) // 4 + 4, device="cuda")
        k.gru_fwd(gi.clone(), b_ih, w_hh, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0)
        gates_f = gi.clone(); k.gru_fwd(gates_f, b_ih, w_hh, b_hh, h0, 0, w_lin, b_lin, hs, ghn, pred, P, T, B, 0)
        def run_b(fn, n=8):
            ts = []
            for _ in range(n):
                gates, gh = gates_f.clone(), ghn.clone()
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(gates, gh); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            return ts[len(ts) // 2]
        tb_ffma = run_b(lambda gates, gh: k.gru_bwd_deferred(gates, gh, hs, h0, 0, w_hh, w_lin, dpred, None, None, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws))
        tb_tc = run_b(lambda gates, gh: k.gru_bwd_tc(gates, gh, hs, h0, 0, w_hh, w_lin, dpred, None, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws))
        print(f"P={P:4d}   bwd ffma {tb_ffma:7.1f} us  bwd tc {tb_tc:7.1f} us", flush=True)
        mb = P * T * B * 2052 / 1e6
        print(f"P={P:4d} tiles={P * ((B + 127) // 128):4d}  fwd ffma {t_ffma:7.1f} us  fwd tc {t_tc:7.1f} us  ({mb / t_tc / 1e3:.2f} TB/s incl. stash)", flush=True)

if __name__ == "__main__":
    main()
