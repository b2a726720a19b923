"""P1 -- kernel-level parity: every C-ABI entry point on the B200 against the CPU oracle on the
same seeded inputs (fp32, tolerance 1e-4 relative as BASELINE.json states; prox / GC decisions
bit-exact).  Calls go through the C ABI (ctypes) exactly as the package does."""
import numpy as np
import pytest
import torch

from oracle import crvae_oracle as O
from tests.cpu_backend import OracleKernels

pytestmark = pytest.mark.gpu
H, G = 64, 192
TOL = 1e-4


def _k():
    import vae_connexe_b200.lib as L
    return L.Kernels()


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _rand(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*s, generator=g) * scale


@pytest.mark.parametrize("form,M,N,K", [(0, 256, 128, 64), (0, 300, 192, 100), (0, 2304, 192, 10), (1, 128, 64, 256),
                                        (1, 192, 100, 2304), (1, 1, 128, 256), (2, 256, 64, 128), (2, 77, 33, 19)])
def test_gemm_forms(form, M, N, K):
    k, o = _k(), OracleKernels()
    if form == 0: A, B = _rand(M, K, seed=1), _rand(N, K, seed=2)
    elif form == 1: A, B = _rand(K, M, seed=1), _rand(K, N, seed=2)
    else: A, B = _rand(M, K, seed=1), _rand(K, N, seed=2)
    bias = _rand(N, seed=3)
    C_ref = torch.zeros(M, N)
    o.gemm(form, 1, M, N, K, A, A.shape[1], 0, B, B.shape[1], 0, C_ref, N, 0, bias, 0)
    C = torch.zeros(M, N, device="cuda")
    k.gemm(form, 1, M, N, K, A.cuda(), A.shape[1], 0, B.cuda(), B.shape[1], 0, C, N, 0, bias.cuda(), 0)
    assert _rel(C, C_ref) < 1e-5
    k.gemm(form, 1, M, N, K, A.cuda(), A.shape[1], 0, B.cuda(), B.shape[1], 0, C, N, 0, None, 0, accumulate=True)
    assert _rel(C, 2 * C_ref - bias) < 1e-5


@pytest.mark.parametrize("P,T,B,K,t_skip", [(3, 10, 64, 10, 1), (5, 10, 256, 100, 1), (1, 10, 256, 100, 0), (2, 4, 40, 7, 1)])
def test_projection_fwd_and_wgrad(P, T, B, K, t_skip):
    k, o = _k(), OracleKernels()
    x, w, b = _rand(T, B, K, seed=1), _rand(P, G, K, seed=2, scale=0.1), _rand(P, G, seed=3)
    mask = (torch.rand(P, K, generator=torch.Generator().manual_seed(4)) < 0.6).to(torch.uint8)
    g_ref = torch.full((P, T, B, G), 7.0)
    o.proj_fwd(x, w, b, g_ref, P, T, B, K, t_skip)
    g_gpu = torch.full((P, T, B, G), 7.0, device="cuda")
    k.proj_fwd(x.cuda(), w.cuda(), b.cuda(), g_gpu, P, T, B, K, t_skip)
    assert _rel(g_gpu, g_ref) < 1e-5
    assert torch.equal(g_gpu[:, :t_skip].cpu(), g_ref[:, :t_skip])          # skipped steps untouched
    dg = _rand(P, T, B, G, seed=5)
    for m in (None, mask):
        dw_ref = torch.zeros(P, G, K)
        o.proj_wgrad(dg, x, m, dw_ref, P, T, B, K, t_skip, None)
        ws = torch.zeros(k.proj_wgrad_workspace(P, T, B, K) // 4 + 4, device="cuda")
        dw = torch.zeros(P, G, K, device="cuda")
        k.proj_wgrad(dg.cuda(), x.cuda(), None if m is None else m.cuda(), dw, P, T, B, K, t_skip, ws)
        assert _rel(dw, dw_ref) < 1e-5
        if m is not None:
            assert torch.equal(dw.cpu()[(m == 0)[:, None, :].expand(P, G, K)], torch.zeros(int((m == 0).sum()) * G))


@pytest.mark.parametrize("P,T,B,K,t_skip", [(1, 10, 256, 100, 0), (3, 10, 64, 12, 1), (5, 10, 256, 100, 1), (2, 4, 40, 36, 1),
                                            (100, 10, 256, 100, 1), (2, 10, 256, 1000, 1),
                                            # resident-W variant (K = 32 nfull + tail <= 8, many row tiles per head, grid full)
                                            (30, 10, 256, 68, 1), (40, 10, 200, 64, 0), (30, 10, 256, 40, 1), (33, 10, 300, 100, 1)])
def test_projection_tensor_core_3xtf32(P, T, B, K, t_skip):
    """tcgen05 + TMA projection (3xTF32) against the fp64 result: fp32-grade accuracy required."""
    k = _k()
    x, w, b = _rand(T, B, K, seed=1), _rand(P, G, K, seed=2, scale=0.1), _rand(P, G, seed=3)
    ref = (torch.einsum("tbk,pgk->ptbg", x.double(), w.double()) + b.double()[:, None, None, :])
    xc, wc = x.cuda(), w.cuda()
    xh, xl, wh, wl = (torch.empty_like(t) for t in (xc, xc, wc, wc))
    k.split_tf32(xc, xh, xl, xc.numel()); k.split_tf32_gate_rows(wc, wh, wl, P * G, K)
    u = torch.arange(32)
    perm = (8 * ((u >> 1) & 3) + 2 * (u >> 3) + (u & 1))                    # row r of W lands at (r & ~31) | perm[r & 31]
    dst = ((torch.arange(P * G) & ~31) | perm[torch.arange(P * G) & 31]).cuda()
    assert torch.equal((wh + wl).view(P * G, K)[dst], wc.view(P * G, K))
    assert torch.equal(xh + xl, xc) and float((xh.view(torch.int32) & 0x1FFF).abs().sum()) == 0     # exact split, hi is tf32
    g = torch.full((P, T, B, G), 7.0, device="cuda")
    k.proj_fwd_tc(xh, xl, wh, wl, b.cuda(), g, P, T, B, K, t_skip)
    torch.cuda.synchronize()
    err_tc = _rel(g[:, t_skip:], ref[:, t_skip:])
    g32 = torch.full((P, T, B, G), 7.0, device="cuda")
    k.proj_fwd(xc, wc, b.cuda(), g32, P, T, B, K, t_skip)
    err_f32 = _rel(g32[:, t_skip:], ref[:, t_skip:])
    assert err_tc < (2e-6 if K <= 128 else 2e-5), (err_tc, err_f32)      # TMEM accumulation rounds differently from FFMA
    assert torch.equal(g[:, :t_skip].cpu(), torch.full((P, t_skip, B, G), 7.0))


@pytest.mark.parametrize("P,T,B,K,t_skip", [(1, 10, 256, 100, 0), (3, 10, 64, 12, 1), (5, 10, 256, 100, 1), (2, 4, 40, 36, 1),
                                            (100, 10, 256, 100, 1), (2, 10, 256, 1000, 1), (3, 10, 256, 200, 1)])
def test_projection_wgrad_tensor_core_3xtf32(P, T, B, K, t_skip):
    """tcgen05 weight gradient with MN-major operands + in-kernel tf32 split, against fp64."""
    k = _k()
    x, dg = _rand(T, B, K, seed=1), _rand(P, T, B, G, seed=5)
    mask = (torch.rand(P, K, generator=torch.Generator().manual_seed(4)) < 0.6).to(torch.uint8)
    ref = torch.einsum("ptbg,tbk->pgk", dg[:, t_skip:].double(), x[t_skip:].double())
    xc = x.cuda()
    xh, xl = torch.empty_like(xc), torch.empty_like(xc)
    k.split_tf32(xc, xh, xl, xc.numel())
    dgc = dg.cuda()
    for m in (None, mask):
        dw = torch.full((P, G, K), 3.0, device="cuda")
        ws = torch.zeros(k.proj_wgrad_tc_workspace(P, T, B, K, t_skip) // 4 + 4, device="cuda")
        k.proj_wgrad_tc(dgc, xh, xl, None if m is None else m.cuda(), dw, P, T, B, K, t_skip, ws)
        torch.cuda.synchronize()
        r = ref if m is None else ref * m[:, None, :].double()
        assert _rel(dw, r) < 2e-5, (_rel(dw, r), P, K)
    assert torch.equal(dgc.cpu(), dg)                      # the gate-gradient buffer itself is not modified


@pytest.mark.parametrize("P,T,B,tile,lin,t_skip,shared_h0", [
    (3, 10, 64, 16, True, 1, True), (3, 10, 64, 32, True, 1, True), (3, 10, 64, 64, True, 1, True),
    (2, 10, 100, 64, True, 1, False), (1, 10, 256, 0, False, 0, True), (7, 3, 37, 32, True, 0, False),
    (40, 10, 256, 0, True, 1, True)])
def test_gru_forward_backward(P, T, B, tile, lin, t_skip, shared_h0):
    k, o = _k(), OracleKernels()
    k.set_batch_tile(tile)
    try:
        gi = _rand(P, T, B, G, seed=1)
        b_ih, w_hh, b_hh = _rand(P, G, seed=2, scale=0.2), _rand(P, G, H, seed=3, scale=0.125), _rand(P, G, seed=4, scale=0.2)
        h0 = _rand(B, H, seed=5) if shared_h0 else _rand(P, B, H, seed=5)
        stride = 0 if shared_h0 else B * H
        w_lin, b_lin = (_rand(P, H, seed=6, scale=0.2), _rand(P, seed=7)) if lin else (None, None)
        c = lambda t: None if t is None else t.cuda()
        out_ref = dict(g=gi.clone(), hs=torch.zeros(P, T, B, H), ghn=torch.zeros(P, T, B, H), pred=torch.zeros(P, T, B) if lin else None)
        o.gru_fwd(out_ref["g"], b_ih, w_hh, b_hh, h0, stride, w_lin, b_lin, out_ref["hs"], out_ref["ghn"], out_ref["pred"], P, T, B, t_skip)
        out = dict(g=gi.clone().cuda(), hs=torch.zeros(P, T, B, H, device="cuda"), ghn=torch.zeros(P, T, B, H, device="cuda"),
                   pred=torch.zeros(P, T, B, device="cuda") if lin else None)
        k.gru_fwd(out["g"], c(b_ih), c(w_hh), c(b_hh), c(h0), stride, c(w_lin), c(b_lin), out["hs"], out["ghn"], out["pred"], P, T, B, t_skip)
        torch.cuda.synchronize()
        for name in ("g", "hs", "ghn") + (("pred",) if lin else ()):
            assert _rel(out[name], out_ref[name]) < 2e-5, name
        # backward, fed with the ORACLE's forward state so the comparison isolates the backward kernel
        dpred = _rand(P, T, B, seed=8) if lin else None
        dh_last = _rand(P, B, H, seed=9, scale=0.1)
        dhs = _rand(P, T, B, H, seed=10, scale=0.1)
        z = lambda *s: torch.zeros(*s)
        ref = dict(g=out_ref["g"].clone(), dw_hh=z(P, G, H), db_hh=z(P, G), db_ih=z(P, G), dw_lin=z(P, H) if lin else None,
                   db_lin=z(P) if lin else None, dh0=z(P, B, H))
        o.gru_bwd(ref["g"], out_ref["ghn"], out_ref["hs"], h0, stride, w_hh, w_lin, dpred, dh_last, dhs, ref["dw_hh"], ref["db_hh"],
                  ref["db_ih"], ref["dw_lin"], ref["db_lin"], ref["dh0"], P, T, B, None)
        gpu = {n: (None if v is None else torch.zeros_like(v).cuda()) for n, v in ref.items()}
        gpu["g"] = out_ref["g"].clone().cuda()
        ws = torch.zeros(k.gru_bwd_workspace(P, B) // 4 + 4, device="cuda")
        k.gru_bwd(gpu["g"], c(out_ref["ghn"]), c(out_ref["hs"]), c(h0), stride, c(w_hh), c(w_lin), c(dpred), c(dh_last), c(dhs),
                  gpu["dw_hh"], gpu["db_hh"], gpu["db_ih"], gpu["dw_lin"], gpu["db_lin"], gpu["dh0"], P, T, B, ws)
        torch.cuda.synchronize()
        for name, v in ref.items():
            if v is not None:
                assert _rel(gpu[name], v) < 5e-5, name
    finally:
        k.set_batch_tile(0)


@pytest.mark.parametrize("P,T,B,lin,t_skip,shared_h0", [(3, 10, 256, True, 1, True), (2, 10, 100, True, 1, False),
                                                         (2, 3, 300, False, 0, True), (1, 10, 128, True, 0, False),
                                                         (20, 10, 256, True, 1, True)])
def test_gru_forward_tensor_core(P, T, B, lin, t_skip, shared_h0):
    """tcgen05 recurrent forward (3xTF32 gate GEMM, fast sigmoid/tanh) against the CPU oracle."""
    k, o = _k(), OracleKernels()
    gi = _rand(P, T, B, G, seed=1)
    b_ih, w_hh, b_hh = _rand(P, G, seed=2, scale=0.2), _rand(P, G, H, seed=3, scale=0.125), _rand(P, G, seed=4, scale=0.2)
    h0 = _rand(B, H, seed=5) if shared_h0 else _rand(P, B, H, seed=5)
    stride = 0 if shared_h0 else B * H
    w_lin, b_lin = (_rand(P, H, seed=6, scale=0.2), _rand(P, seed=7)) if lin else (None, None)
    c = lambda t: None if t is None else t.cuda()
    ref = dict(g=gi.clone(), hs=torch.zeros(P, T, B, H), ghn=torch.zeros(P, T, B, H), pred=torch.zeros(P, T, B) if lin else None)
    o.gru_fwd(ref["g"], b_ih, w_hh, b_hh, h0, stride, w_lin, b_lin, ref["hs"], ref["ghn"], ref["pred"], P, T, B, t_skip)
    out = dict(g=gi.clone().cuda(), hs=torch.zeros(P, T, B, H, device="cuda"), ghn=torch.zeros(P, T, B, H, device="cuda"),
               pred=torch.zeros(P, T, B, device="cuda") if lin else None)
    wc = w_hh.cuda()
    wh, wl = torch.empty_like(wc), torch.empty_like(wc)
    k.split_tf32(wc, wh, wl, wc.numel())
    k.gru_fwd_tc(out["g"], c(b_ih), wh, wl, c(b_hh), c(h0), stride, c(w_lin), c(b_lin), out["hs"], out["ghn"], out["pred"], P, T, B, t_skip)
    torch.cuda.synchronize()
    for name in ("g", "hs", "ghn") + (("pred",) if lin else ()):
        assert _rel(out[name], ref[name]) < 2e-5, (name, _rel(out[name], ref[name]))


@pytest.mark.parametrize("P,T,B,lin,t_skip,shared_h0,last,with_dhs", [
    (1, 10, 256, False, 0, False, True, False),      # the encoder (gru_left, :208): gradient enters through h_T only
    (13, 10, 256, True, 1, True, False, False),      # a 13-head decoder shard (p = 100 over 8 GPUs): h0 = z shared
    (20, 10, 256, True, 1, True, False, False),      # 320 CTAs > 148 SMs: the two-CTAs-per-SM ring depths
    (2, 10, 100, True, 1, False, False, False),      # ragged last tile (4 of 16 rows)
    (1, 25, 33, True, 0, True, True, True),          # long sequence (ring wraps many times), ragged, per-step dhs (VRAE.py decoder)
    (3, 1, 16, True, 0, False, False, False),        # a single step
    (2, 2, 48, False, 2, False, True, False),        # every step is a zero-input step
    (30, 3, 144, True, 1, True, False, False),       # 270 tiles > 148 SMs: persistent CTAs cross tile and head boundaries
    (12, 4, 200, True, 0, False, False, False),      # 156 tiles: tile pairs, odd tile count per head (idle second stream), ragged last tile
    (22, 10, 256, True, 1, True, False, False),      # 176 tile pairs on 148 CTAs: head boundaries inside a CTA's range at a zero-input step
    (20, 4, 256, True, 0, False, True, True),        # tile pairs with per-head h0, dh_last and per-step dhs (2-deep rings of the MMA BPTT)
    (40, 1, 80, True, 0, True, True, False),         # 200 tiles of a single step: every step is a tile start and a tile end
])
@pytest.mark.parametrize("family", ["ll", "mma"])
def test_gru_low_latency_forward_backward(P, T, B, lin, t_skip, shared_h0, last, with_dhs, family):
    """crvae_gru_fwd_ll / crvae_gru_bwd_ll (16-row tiles, bulk-copy slab ring, packed fp32 FMAs) and crvae_gru_fwd_mma /
    crvae_gru_bwd_mma (16-row tiles, W_hh in registers, warp-level 3xTF32 MMAs) against the CPU oracle, and to fp32 rounding
    (ll) / 3xTF32 rounding (mma) against the exact FFMA kernels they stand in for."""
    k, o = _k(), OracleKernels()
    fwd_fn, bwd_fn, tol_exact = (k.gru_fwd_ll, k.gru_bwd_ll, 2e-6) if family == "ll" else (k.gru_fwd_mma, k.gru_bwd_mma, 1e-5)
    gi = _rand(P, T, B, G, seed=1)
    b_ih, w_hh, b_hh = _rand(P, G, seed=2, scale=0.2), _rand(P, G, H, seed=3, scale=0.125), _rand(P, G, seed=4, scale=0.2)
    h0 = _rand(B, H, seed=5) if shared_h0 else _rand(P, B, H, seed=5)
    stride = 0 if shared_h0 else B * H
    w_lin, b_lin = (_rand(P, H, seed=6, scale=0.2), _rand(P, seed=7)) if lin else (None, None)
    c = lambda t: None if t is None else t.cuda()
    ref = dict(g=gi.clone(), hs=torch.zeros(P, T, B, H), ghn=torch.zeros(P, T, B, H), pred=torch.zeros(P, T, B) if lin else None)
    o.gru_fwd(ref["g"], b_ih, w_hh, b_hh, h0, stride, w_lin, b_lin, ref["hs"], ref["ghn"], ref["pred"], P, T, B, t_skip)
    outs = []
    for fn in (fwd_fn, k.gru_fwd):
        out = dict(g=gi.clone().cuda(), hs=torch.zeros(P, T, B, H, device="cuda"), ghn=torch.zeros(P, T, B, H, device="cuda"),
                   pred=torch.zeros(P, T, B, device="cuda") if lin else None)
        fn(out["g"], c(b_ih), c(w_hh), c(b_hh), c(h0), stride, c(w_lin), c(b_lin), out["hs"], out["ghn"], out["pred"], P, T, B, t_skip)
        torch.cuda.synchronize()
        outs.append(out)
    ll, ex = outs
    for name in ("g", "hs", "ghn") + (("pred",) if lin else ()):
        assert _rel(ll[name], ref[name]) < 2e-5, (name, _rel(ll[name], ref[name]))
    # exact fp32 on both sides; the low-latency kernel sums k in two fixed halves, so agreement is to rounding, not bit-for-bit
    assert _rel(ll["hs"], ex["hs"]) < tol_exact and _rel(ll["g"], ex["g"]) < tol_exact and _rel(ll["ghn"], ex["ghn"]) < tol_exact
    # ---- backward on the forward's outputs ----
    dpred = _rand(P, T, B, seed=8) if lin else None
    dh_last = _rand(P, B, H, seed=9, scale=0.1) if last else None
    dhs = _rand(P, T, B, H, seed=10, scale=0.1) if with_dhs else None
    z = lambda *s: torch.zeros(*s)
    rb = dict(g=ref["g"].clone(), dw_hh=z(P, G, H), db_hh=z(P, G), db_ih=z(P, G), dw_lin=z(P, H) if lin else None,
              db_lin=z(P) if lin else None, dh0=z(P, B, H))
    o.gru_bwd(rb["g"], ref["ghn"], ref["hs"], h0, stride, w_hh, w_lin, dpred, dh_last, dhs, rb["dw_hh"], rb["db_hh"], rb["db_ih"],
              rb["dw_lin"], rb["db_lin"], rb["dh0"], P, T, B, None)
    gpu = {n: (None if v is None else torch.zeros_like(v).cuda()) for n, v in rb.items()}
    gpu["g"] = ll["g"].clone()
    ghn = ll["ghn"].clone()
    ws = torch.zeros(k.gru_bwd_workspace(P, B) // 4 + 4, device="cuda")
    bwd_fn(gpu["g"], ghn, ll["hs"], c(h0), stride, c(w_hh), c(w_lin), c(dpred), c(dh_last), c(dhs), gpu["db_hh"], gpu["db_ih"],
           gpu["dw_lin"], gpu["db_lin"], gpu["dh0"], P, T, B, ws)
    torch.cuda.synchronize()
    for name in ("g", "db_hh", "db_ih", "dw_lin", "db_lin", "dh0"):
        if rb[name] is not None:
            assert _rel(gpu[name], rb[name]) < 5e-5, (name, _rel(gpu[name], rb[name]))
    assert _rel(ghn, rb["g"][..., 2 * H:] * ref["g"][..., :H]) < 5e-5          # dgh_n = dgi_n * r left in the ghn buffer
    if B % 32 == 0:
        ws2 = torch.zeros(k.gru_dwhh_tc_workspace(P, T, B) // 4 + 4, device="cuda")
        k.gru_dwhh_tc(gpu["g"], ghn, ll["hs"], c(h0), stride, gpu["dw_hh"], P, T, B, ws2)
        torch.cuda.synchronize()
        assert _rel(gpu["dw_hh"], rb["dw_hh"]) < 5e-5


@pytest.mark.parametrize("P,T,B,lin,shared_h0,last", [(3, 10, 256, True, True, False), (2, 5, 64, False, False, True),
                                                       (20, 10, 256, True, True, False), (1, 10, 32, True, True, True),
                                                       (2, 3, 300, True, False, True), (1, 1, 128, True, True, False)])
def test_gru_backward_tensor_core(P, T, B, lin, shared_h0, last):
    """tcgen05 BPTT (crvae_gru_bwd_tc: dgh operand in TMEM, 3xTF32) + crvae_gru_dwhh_tc against the CPU oracle."""
    k, o = _k(), OracleKernels()
    gi = _rand(P, T, B, G, seed=1)
    b_ih, w_hh, b_hh = _rand(P, G, seed=2, scale=0.2), _rand(P, G, H, seed=3, scale=0.125), _rand(P, G, seed=4, scale=0.2)
    h0 = _rand(B, H, seed=5) if shared_h0 else _rand(P, B, H, seed=5)
    stride = 0 if shared_h0 else B * H
    w_lin, b_lin = (_rand(P, H, seed=6, scale=0.2), _rand(P, seed=7)) if lin else (None, None)
    c = lambda t: None if t is None else t.cuda()
    fw = dict(g=gi.clone(), hs=torch.zeros(P, T, B, H), ghn=torch.zeros(P, T, B, H), pred=torch.zeros(P, T, B) if lin else None)
    o.gru_fwd(fw["g"], b_ih, w_hh, b_hh, h0, stride, w_lin, b_lin, fw["hs"], fw["ghn"], fw["pred"], P, T, B, 0)
    dpred = _rand(P, T, B, seed=8) if lin else None
    dh_last = _rand(P, B, H, seed=9, scale=0.1) if last else None
    z = lambda *s: torch.zeros(*s)
    ref = dict(g=fw["g"].clone(), dw_hh=z(P, G, H), db_hh=z(P, G), db_ih=z(P, G), dw_lin=z(P, H) if lin else None,
               db_lin=z(P) if lin else None, dh0=z(P, B, H))
    o.gru_bwd(ref["g"], fw["ghn"], fw["hs"], h0, stride, w_hh, w_lin, dpred, dh_last, None, ref["dw_hh"], ref["db_hh"], ref["db_ih"],
              ref["dw_lin"], ref["db_lin"], ref["dh0"], P, T, B, None)
    gpu = {n: (None if v is None else torch.zeros_like(v).cuda()) for n, v in ref.items()}
    gpu["g"] = fw["g"].clone().cuda()
    ghn = fw["ghn"].clone().cuda()
    ws = torch.zeros(k.gru_bwd_workspace(P, B) // 4 + 4, device="cuda")
    k.gru_bwd_tc(gpu["g"], ghn, c(fw["hs"]), c(h0), stride, c(w_hh), c(w_lin), c(dpred), c(dh_last), gpu["db_hh"], gpu["db_ih"],
                 gpu["dw_lin"], gpu["db_lin"], gpu["dh0"], P, T, B, ws)
    torch.cuda.synchronize()
    for name in ("g", "db_hh", "db_ih", "dw_lin", "db_lin", "dh0"):
        if ref[name] is not None:
            assert _rel(gpu[name], ref[name]) < 5e-5, (name, _rel(gpu[name], ref[name]))
    if B % 32 == 0:
        ws2 = torch.zeros(k.gru_dwhh_tc_workspace(P, T, B) // 4 + 4, device="cuda")
        k.gru_dwhh_tc(gpu["g"], ghn, c(fw["hs"]), c(h0), stride, gpu["dw_hh"], P, T, B, ws2)
        torch.cuda.synchronize()
        assert _rel(gpu["dw_hh"], ref["dw_hh"]) < 5e-5


@pytest.mark.parametrize("P,T,B,lin,shared_h0", [(3, 10, 256, True, True), (2, 5, 64, False, False), (20, 10, 256, True, True),
                                                  (1, 10, 32, True, True)])
def test_gru_backward_deferred_dwhh_tensor_core(P, T, B, lin, shared_h0):
    """BPTT with dW_hh deferred to a tcgen05 GEMM (crvae_gru_bwd_deferred + crvae_gru_dwhh_tc) against the oracle and
    against the in-kernel FFMA accumulation."""
    k, o = _k(), OracleKernels()
    gi = _rand(P, T, B, G, seed=1)
    b_ih, w_hh, b_hh = _rand(P, G, seed=2, scale=0.2), _rand(P, G, H, seed=3, scale=0.125), _rand(P, G, seed=4, scale=0.2)
    h0 = _rand(B, H, seed=5) if shared_h0 else _rand(P, B, H, seed=5)
    stride = 0 if shared_h0 else B * H
    w_lin, b_lin = (_rand(P, H, seed=6, scale=0.2), _rand(P, seed=7)) if lin else (None, None)
    c = lambda t: None if t is None else t.cuda()
    fw = dict(g=gi.clone(), hs=torch.zeros(P, T, B, H), ghn=torch.zeros(P, T, B, H), pred=torch.zeros(P, T, B) if lin else None)
    o.gru_fwd(fw["g"], b_ih, w_hh, b_hh, h0, stride, w_lin, b_lin, fw["hs"], fw["ghn"], fw["pred"], P, T, B, 0)
    dpred = _rand(P, T, B, seed=8) if lin else None
    dh_last, dhs = _rand(P, B, H, seed=9, scale=0.1), _rand(P, T, B, H, seed=10, scale=0.1)
    z = lambda *s: torch.zeros(*s)
    ref = dict(g=fw["g"].clone(), dw_hh=z(P, G, H), db_hh=z(P, G), db_ih=z(P, G), dw_lin=z(P, H) if lin else None,
               db_lin=z(P) if lin else None, dh0=z(P, B, H))
    o.gru_bwd(ref["g"], fw["ghn"], fw["hs"], h0, stride, w_hh, w_lin, dpred, dh_last, dhs, ref["dw_hh"], ref["db_hh"], ref["db_ih"],
              ref["dw_lin"], ref["db_lin"], ref["dh0"], P, T, B, None)
    gpu = {n: (None if v is None else torch.zeros_like(v).cuda()) for n, v in ref.items()}
    gpu["g"] = fw["g"].clone().cuda()
    ghn = fw["ghn"].clone().cuda()
    ws = torch.zeros(k.gru_bwd_workspace(P, B) // 4 + 4, device="cuda")
    k.gru_bwd_deferred(gpu["g"], ghn, c(fw["hs"]), c(h0), stride, c(w_hh), c(w_lin), c(dpred), c(dh_last), c(dhs), gpu["db_hh"],
                       gpu["db_ih"], gpu["dw_lin"], gpu["db_lin"], gpu["dh0"], P, T, B, ws)
    ws2 = torch.zeros(k.gru_dwhh_tc_workspace(P, T, B) // 4 + 4, device="cuda")
    k.gru_dwhh_tc(gpu["g"], ghn, c(fw["hs"]), c(h0), stride, gpu["dw_hh"], P, T, B, ws2)
    torch.cuda.synchronize()
    for name, v in ref.items():
        if v is not None:
            assert _rel(gpu[name], v) < 5e-5, (name, _rel(gpu[name], v))
    # dgh_n left in the ghn buffer = dgi_n * r
    r_ = fw["g"][..., :H]
    assert _rel(ghn, ref["g"][..., 2 * H:] * r_) < 5e-5


@pytest.mark.parametrize("form", [0, 1])
@pytest.mark.parametrize("B", [64, 256])
def test_latent_fwd_bwd(form, B):
    k, o = _k(), OracleKernels()
    lat, eps = _rand(B, 2 * H, seed=1, scale=0.5), _rand(B, H, seed=2)
    z_ref, kl_ref = torch.zeros(B, H), torch.zeros(1)
    o.latent_fwd(lat, eps, z_ref, kl_ref, B, form)
    z, kl = torch.zeros(B, H, device="cuda"), torch.zeros(1, device="cuda")
    k.latent_fwd(lat.cuda(), eps.cuda(), z, kl, B, form)
    assert _rel(z, z_ref) < 1e-6 and _rel(kl, kl_ref) < 1e-6
    P = 9
    dh0, extra = _rand(P, B, H, seed=3), _rand(B, H, seed=4)
    for ex in (None, extra):
        d_ref, dz_ref = torch.zeros(B, 2 * H), torch.zeros(B, H)
        o.latent_bwd(dh0, P, ex, lat, eps, 0.1, form, d_ref, dz_ref, B)
        d, dz = torch.zeros(B, 2 * H, device="cuda"), torch.zeros(B, H, device="cuda")
        k.latent_bwd(dh0.cuda(), P, None if ex is None else ex.cuda(), lat.cuda(), eps.cuda(), 0.1, form, d, dz, B)
        assert _rel(d, d_ref) < 1e-5 and _rel(dz, dz_ref) < 1e-6
    # two-stage form used across ranks: head sum only, then pointwise on the reduced dz
    dz = torch.zeros(B, H, device="cuda")
    k.latent_bwd(dh0.cuda(), P, None, None, None, 0.0, form, None, dz, B)
    d2 = torch.zeros(B, 2 * H, device="cuda")
    k.latent_bwd(None, 0, dz, lat.cuda(), eps.cuda(), 0.1, form, d2, None, B)
    d_ref = torch.zeros(B, 2 * H)
    o.latent_bwd(dh0, P, None, lat, eps, 0.1, form, d_ref, None, B)
    assert _rel(d2, d_ref) < 1e-5


def test_mse_fwd_bwd():
    k, o = _k(), OracleKernels()
    P, T, B = 11, 10, 256
    pred, tgt = _rand(P, T, B, seed=1), _rand(P, T, B, seed=2)
    r = dict(sse=torch.zeros(P), dp=torch.zeros(P, T, B), err=torch.zeros(P, T, B))
    o.mse_fwd_bwd(pred, tgt, r["sse"], r["dp"], r["err"], P, T, B)
    g = {n: torch.zeros_like(v).cuda() for n, v in r.items()}
    k.mse_fwd_bwd(pred.cuda(), tgt.cuda(), g["sse"], g["dp"], g["err"], P, T, B)
    assert _rel(g["sse"], r["sse"]) < 1e-6
    assert torch.equal(g["err"].cpu(), r["err"])
    assert _rel(g["dp"], r["dp"]) < 1e-6
    out = torch.zeros(1, device="cuda")
    k.dot_small(g["sse"], P, 1.0 / (T * B), out)
    assert abs(float(out) - float(r["sse"].sum() / (T * B))) < 1e-5


@pytest.mark.parametrize("P,K", [(4, 10), (100, 100), (3, 33)])
@pytest.mark.parametrize("masked", [False, True])
def test_gd_prox_gc_bit_exact(P, K, masked):
    """GD + prox on identical tensors: same zero pattern as torch's prox_update (:308-314) and
    weights equal to 1 ulp-level (the column norm is the only non-bit-identical intermediate)."""
    k = _k()
    lam, lr = 0.1, 5e-2
    thr = np.float32(lam * lr)
    gen = torch.Generator().manual_seed(P * 1000 + K)
    w = torch.randn(P, G, K, generator=gen) * 0.05
    # adversarial columns: norms at thr*(1 +- eps), exactly zero, far above/below
    unit = w / torch.norm(w, dim=1, keepdim=True)
    scales = torch.tensor([1 + 1e-3, 1 - 1e-3, 1 + 1e-5, 1 - 1e-5, 0.0, 0.3, 3.0, 1 + 1e-6, 1 - 1e-6])[: min(K, 9)]
    w[:, :, : len(scales)] = unit[:, :, : len(scales)] * scales * float(thr)
    dw = torch.zeros(P, G, K)
    dw[:, :, len(scales):] = torch.randn(P, G, K - len(scales), generator=gen) * 0.01
    mask = (torch.rand(P, K, generator=gen) < 0.7).to(torch.uint8) if masked else None
    if masked:
        w = w * mask[:, None, :].float()
    w_ref = w - np.float32(lr) * dw
    if masked:
        w_ref = w_ref * mask[:, None, :].float()
    w_ref = O.prox_update(w_ref, lam, lr)
    wg, cn = w.clone().cuda(), torch.zeros(P, K, device="cuda")
    k.gd_prox_gc(wg, dw.cuda(), None if mask is None else mask.cuda(), cn, P, K, float(np.float32(lr)), float(thr), True)
    nz_ref = torch.norm(w_ref, dim=1) > 0
    assert torch.equal((cn > 0).cpu(), nz_ref)                                  # GC decision bit-exact
    assert torch.equal((wg != 0).any(1).cpu(), nz_ref)
    assert _rel(wg, w_ref) < 1e-5
    assert _rel(cn, torch.norm(w_ref, dim=1)) < 1e-5
    # norms-only and prox-only entry forms
    cn2 = torch.zeros(P, K, device="cuda")
    before = wg.clone()
    k.gd_prox_gc(wg, None, None if mask is None else mask.cuda(), cn2, P, K, 0.0, 0.0, False)
    assert torch.equal(wg, before) and torch.equal(cn2, cn)


def test_gd_step_bit_exact():
    k = _k()
    th, g = _rand(100003, seed=1), _rand(100003, seed=2)
    ref = th.clone()
    ref -= np.float32(0.05) * g
    t = th.clone().cuda()
    k.gd_step(t, g.cuda(), th.numel(), float(np.float32(0.05)))
    assert torch.equal(t.cpu(), ref)


def test_adam_matches_torch():
    k = _k()
    n = 50001
    p0 = _rand(n, seed=1)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    th, m, v = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 8):
        g = _rand(n, seed=10 + step)
        ref.grad = g.clone()
        opt.step()
        k.adam_step(th, g.cuda(), m, v, n, 1e-3, 0.9, 0.999, 1e-8, step)
        assert float((th.cpu() - ref.detach()).abs().max()) < 2e-7


def test_sumsq_axpy():
    k = _k()
    x = _rand(192 * 64 * 7, seed=1)
    out = torch.zeros(1, device="cuda")
    k.sumsq(x.cuda(), x.numel(), out)
    assert abs(float(out) - float((x.double() ** 2).sum())) < 1e-6 * float((x.double() ** 2).sum())
    y = _rand(1000, seed=2)
    yg = y.clone().cuda()
    k.axpy(yg, x[:1000].cuda(), 1000, 0.5)
    assert _rel(yg, y + 0.5 * x[:1000]) < 1e-6


@pytest.mark.parametrize("B,K", [(48, 10), (256, 10), (7, 1), (33, 32)])
def test_cs_divergence_fwd_bwd(B, K):
    """Cauchy-Schwarz divergence head (CR-CS-RAE.py:124-163) forward + backward vs the oracle."""
    k = _k()
    lat, pm, pl = _rand(B, 2 * H, seed=1, scale=0.3), _rand(K, H, seed=2, scale=0.3), _rand(K, H, seed=3, scale=0.2)
    # (D_CS >= 0 by Cauchy-Schwarz: the clamp(min=0) only acts on rounding noise when q == p exactly, where the
    #  pass/clamp decision is legitimately ambiguous between implementations -- not constructed here)
    cs_ref, dl_ref, dm_ref, dv_ref = O.cs_head(lat, pm, pl, 0.1)
    ws = torch.zeros(k.cs_div_workspace(B, K) // 4 + 4, device="cuda")
    cs, dl = torch.zeros(1, device="cuda"), torch.zeros(B, 2 * H, device="cuda")
    dm, dv = torch.zeros(K, H, device="cuda"), torch.zeros(K, H, device="cuda")
    k.cs_div_fwd_bwd(lat.cuda(), pm.cuda(), pl.cuda(), B, K, 0.1, cs, dl, dm, dv, ws)
    torch.cuda.synchronize()
    assert abs(float(cs) - float(cs_ref)) < 2e-5 * max(1.0, abs(float(cs_ref)))
    scale = float(dl_ref.abs().max())
    assert float((dl.cpu() - dl_ref).abs().max()) < 1e-4 * scale
    assert float((dm.cpu() - dm_ref).abs().max()) < 1e-4 * float(dm_ref.abs().max())
    assert float((dv.cpu() - dv_ref).abs().max()) < 1e-4 * float(dv_ref.abs().max())


@pytest.mark.parametrize("B,kl_form", [(256, 1), (100, 0), (16, 1), (300, 1)])
def test_latent_head_fused(B, kl_form):
    """Fused fc_mu|fc_std + reparameterisation + KL (forward) and its three gradient products (backward) against fp64 torch."""
    k = _k()
    hT, w, b, eps = _rand(B, H, seed=1), _rand(2 * H, H, seed=2, scale=0.125), _rand(2 * H, seed=3, scale=0.1), _rand(B, H, seed=4)
    c = lambda t: t.cuda()
    lat, z, kl = torch.zeros(B, 2 * H, device="cuda"), torch.zeros(B, H, device="cuda"), torch.zeros(1, device="cuda")
    ws = torch.zeros(k.latent_head_workspace(B) // 4 + 4, device="cuda")
    for _ in range(2):      # twice: the completion counter must reset itself
        k.latent_head_fwd(c(hT), c(w), c(b), c(eps), lat, z, kl, B, kl_form, ws)
    torch.cuda.synchronize()
    lat_ref = hT.double() @ w.double().T + b.double()
    mu, lv = lat_ref[:, :H], lat_ref[:, H:]
    z_ref = mu + torch.exp(0.5 * lv) * eps.double()
    kl_ref = (-0.5 * (1 + mu - lv * lv - torch.exp(mu))).sum() / B if kl_form == 1 else (-0.5 * (1 + lv - mu * mu - torch.exp(lv))).sum() / B
    assert _rel(lat, lat_ref) < 2e-6 and _rel(z, z_ref) < 2e-6
    assert abs(float(kl) - float(kl_ref)) < 2e-6 * abs(float(kl_ref)) + 1e-6
    dlat = _rand(B, 2 * H, seed=5)
    dW, db, dhT = torch.zeros(2 * H, H, device="cuda"), torch.zeros(2 * H, device="cuda"), torch.zeros(B, H, device="cuda")
    k.latent_head_bwd(c(dlat), c(hT), c(w), dW, db, dhT, B)
    torch.cuda.synchronize()
    assert _rel(dW, dlat.double().T @ hT.double()) < 2e-6
    assert _rel(db, dlat.double().sum(0)) < 2e-6
    assert _rel(dhT, dlat.double() @ w.double()) < 2e-6


@pytest.mark.parametrize("B,p,head_lo,P", [(256, 100, 0, 100), (40, 12, 3, 5), (70, 130, 120, 10), (33, 260, 0, 0)])
def test_bind_batch_fused(B, p, head_lo, P):
    """One-pass batch binding (time-major inputs + tf32 splits + per-head targets) against the torch reshapes."""
    k = _k()
    Te = Td = 10
    X = _rand(B, Te + Td, p, seed=1).cuda()
    z = lambda *s: torch.zeros(*s, device="cuda")
    enc, eh, el, dec, dh, dl = z(Te, B, p), z(Te, B, p), z(Te, B, p), z(Td, B, p), z(Td, B, p), z(Td, B, p)
    tgt = z(max(P, 1), Td, B)
    k.bind_batch(X, enc, eh, el, dec, dh, dl, tgt, B, p, Te, Td, head_lo, P)
    torch.cuda.synchronize()
    assert torch.equal(enc, X[:, :Te].transpose(0, 1))
    assert torch.equal(dec[1:], X[:, Te:-1].transpose(0, 1)) and float(dec[0].abs().sum()) == 0
    assert torch.equal(eh + el, enc) and torch.equal(dh + dl, dec)
    assert float((eh.view(torch.int32) & 0x1FFF).abs().sum()) == 0 and float((dh.view(torch.int32) & 0x1FFF).abs().sum()) == 0
    if P > 0:
        assert torch.equal(tgt[:P], X[:, Te:, head_lo:head_lo + P].permute(2, 1, 0))
