"""Memory-safety and race evidence for the hand-written kernels WITHOUT compute-sanitizer (the GPU pool refuses the tool,
profiles/r02_sanitizer_unavailable.txt).  Every buffer a kernel touches is a window inside a larger allocation whose
margins hold a canary bit pattern; every kernel runs twice on identical inputs:

  * out-of-bounds WRITES change a canary                       -> detected bit-exactly;
  * out-of-bounds READS reach the canary, a signalling NaN      -> poison the outputs (outputs must be finite);
  * data races / reads of uninitialised shared or tensor memory -> run-to-run differences (outputs must be bit-identical,
    and between the two runs the scratch buffers are refilled with garbage);
  * ragged sizes (B not a multiple of the 128-/16-row tiles, K not a multiple of the 32-wide TMA boxes) exercise the
    partial-tile paths where such bugs live.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
H, G = 64, 192
GUARD = 256                       # floats on each side (1 KB: keeps the 32-byte / 1024-byte alignment of the window)
CANARY = float("nan")


def _k():
    import vae_connexe_b200.lib as L
    return L.Kernels()


class Guarded:
    """A named set of device tensors, each a window in a canary-filled allocation."""

    def __init__(self):
        self.raw, self.win, self.init = {}, {}, {}

    def add(self, name, value=None, shape=None, dtype=torch.float32):
        if value is not None:
            shape, dtype = tuple(value.shape), value.dtype
        n = int(np.prod(shape))
        pad = (-n) % 64
        raw = torch.full((n + pad + 2 * GUARD,), CANARY if dtype == torch.float32 else 0x5A, dtype=dtype, device="cuda")
        self.raw[name] = raw
        self.win[name] = raw[GUARD:GUARD + n].view(*shape)
        self.init[name] = None if value is None else value.to("cuda")
        return self.win[name]

    def reset(self, garbage_seed):
        g = torch.Generator(device="cuda").manual_seed(garbage_seed)
        for name, w in self.win.items():
            if self.init[name] is not None:
                w.copy_(self.init[name])
            elif w.dtype == torch.float32:                      # outputs / scratch: different garbage before every run
                w.copy_(torch.randn(w.shape, generator=g, device="cuda") * 1e3)
            else:
                w.zero_()

    def check_canaries(self):
        for name, raw in self.raw.items():
            n = self.win[name].numel()
            lo, hi = raw[:GUARD], raw[GUARD + n + ((-n) % 64):]
            if raw.dtype == torch.float32:
                assert bool(torch.isnan(lo).all()) and bool(torch.isnan(hi).all()), f"out-of-bounds write around `{name}`"
            else:
                assert bool((lo == 0x5A).all()) and bool((hi == 0x5A).all()), f"out-of-bounds write around `{name}`"

    def __getitem__(self, name):
        return self.win[name]


def _twice(gd: Guarded, launch, outputs):
    snaps = []
    for rep in range(2):
        gd.reset(100 + rep)
        launch()
        torch.cuda.synchronize()
        gd.check_canaries()
        snaps.append({n: gd[n].clone() for n in outputs})
    for n in outputs:
        assert bool(torch.isfinite(snaps[0][n]).all()), f"`{n}` is not finite: a read reached canary / uninitialised memory"
        assert torch.equal(snaps[0][n], snaps[1][n]), f"`{n}` differs between two identical runs (race or uninitialised read)"
    return snaps[0]


def _rand(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*s, generator=g) * scale


@pytest.mark.parametrize("P,T,B,t_skip", [(2, 10, 256, 1), (3, 4, 100, 0), (1, 10, 33, 1), (9, 10, 129, 1)])
def test_guard_recurrent_tensor_core(P, T, B, t_skip):
    k = _k()
    gd = Guarded()
    gd.add("gates", _rand(P, T, B, G, seed=1))
    gd.add("b_ih", _rand(P, G, seed=2, scale=0.2)); gd.add("w_hh", _rand(P, G, H, seed=3, scale=0.125))
    gd.add("b_hh", _rand(P, G, seed=4, scale=0.2)); gd.add("h0", _rand(B, H, seed=5))
    gd.add("w_lin", _rand(P, H, seed=6, scale=0.2)); gd.add("b_lin", _rand(P, seed=7))
    gd.add("hs", shape=(P, T, B, H)); gd.add("ghn", shape=(P, T, B, H)); gd.add("pred", shape=(P, T, B))
    fw = _twice(gd, lambda: k.gru_fwd_tc(gd["gates"], gd["b_ih"], gd["w_hh"], None, gd["b_hh"], gd["h0"], 0, gd["w_lin"], gd["b_lin"],
                                         gd["hs"], gd["ghn"], gd["pred"], P, T, B, t_skip), ("gates", "hs", "ghn", "pred"))
    # BPTT + deferred dW_hh on the forward's outputs
    gb = Guarded()
    gb.add("gates", fw["gates"]); gb.add("ghn", fw["ghn"]); gb.add("hs", fw["hs"]); gb.add("h0", gd["h0"].clone())
    gb.add("w_hh", gd["w_hh"].clone()); gb.add("w_lin", gd["w_lin"].clone()); gb.add("dpred", _rand(P, T, B, seed=8))
    for n, s in (("db_hh", (P, G)), ("db_ih", (P, G)), ("dw_lin", (P, H)), ("db_lin", (P,)), ("dh0", (P, B, H)), ("dw_hh", (P, G, H))):
        gb.add(n, shape=s)
    gb.add("ws", shape=(k.gru_bwd_workspace(P, B) // 4 + 4,))
    gb.add("ws2", shape=(k.gru_dwhh_tc_workspace(P, T, B) // 4 + 4,))

    def bwd():
        k.gru_bwd_tc(gb["gates"], gb["ghn"], gb["hs"], gb["h0"], 0, gb["w_hh"], gb["w_lin"], gb["dpred"], None, gb["db_hh"], gb["db_ih"],
                     gb["dw_lin"], gb["db_lin"], gb["dh0"], P, T, B, gb["ws"])
        if B % 32 == 0:
            k.gru_dwhh_tc(gb["gates"], gb["ghn"], gb["hs"], gb["h0"], 0, gb["dw_hh"], P, T, B, gb["ws2"])
    outs = ("gates", "ghn", "db_hh", "db_ih", "dw_lin", "db_lin", "dh0") + (("dw_hh",) if B % 32 == 0 else ())
    _twice(gb, bwd, outs)


@pytest.mark.parametrize("P,T,B,tile", [(2, 10, 256, 16), (3, 5, 100, 32), (1, 10, 33, 16), (2, 3, 70, 64)])
def test_guard_recurrent_exact(P, T, B, tile):
    k = _k()
    k.set_batch_tile(tile)
    try:
        gd = Guarded()
        gd.add("gates", _rand(P, T, B, G, seed=1))
        gd.add("b_ih", _rand(P, G, seed=2, scale=0.2)); gd.add("w_hh", _rand(P, G, H, seed=3, scale=0.125))
        gd.add("b_hh", _rand(P, G, seed=4, scale=0.2)); gd.add("h0", _rand(P, B, H, seed=5))
        gd.add("w_lin", _rand(P, H, seed=6, scale=0.2)); gd.add("b_lin", _rand(P, seed=7))
        gd.add("hs", shape=(P, T, B, H)); gd.add("ghn", shape=(P, T, B, H)); gd.add("pred", shape=(P, T, B))
        fw = _twice(gd, lambda: k.gru_fwd(gd["gates"], gd["b_ih"], gd["w_hh"], gd["b_hh"], gd["h0"], B * H, gd["w_lin"], gd["b_lin"],
                                          gd["hs"], gd["ghn"], gd["pred"], P, T, B, 1), ("gates", "hs", "ghn", "pred"))
        gb = Guarded()
        gb.add("gates", fw["gates"]); gb.add("ghn", fw["ghn"]); gb.add("hs", fw["hs"]); gb.add("h0", gd["h0"].clone())
        gb.add("w_hh", gd["w_hh"].clone()); gb.add("w_lin", gd["w_lin"].clone()); gb.add("dpred", _rand(P, T, B, seed=8))
        gb.add("dhs", _rand(P, T, B, H, seed=9, scale=0.1))
        for n, s in (("dw_hh", (P, G, H)), ("db_hh", (P, G)), ("db_ih", (P, G)), ("dw_lin", (P, H)), ("db_lin", (P,)), ("dh0", (P, B, H))):
            gb.add(n, shape=s)
        gb.add("ws", shape=(k.gru_bwd_workspace(P, B) // 4 + 4,))
        _twice(gb, lambda: k.gru_bwd(gb["gates"], gb["ghn"], gb["hs"], gb["h0"], B * H, gb["w_hh"], gb["w_lin"], gb["dpred"], None, gb["dhs"],
                                     gb["dw_hh"], gb["db_hh"], gb["db_ih"], gb["dw_lin"], gb["db_lin"], gb["dh0"], P, T, B, gb["ws"]),
               ("gates", "dw_hh", "db_hh", "db_ih", "dw_lin", "db_lin", "dh0"))
    finally:
        k.set_batch_tile(0)


@pytest.mark.parametrize("P,T,B", [(1, 10, 256), (13, 10, 256), (2, 7, 100), (1, 25, 33), (20, 10, 64)])
def test_guard_recurrent_low_latency(P, T, B):
    """The bulk-copy kernels (cp.async.bulk in both directions, slots reused in place): canaries + run-to-run identity."""
    k = _k()
    gd = Guarded()
    gd.add("gates", _rand(P, T, B, G, seed=1))
    gd.add("b_ih", _rand(P, G, seed=2, scale=0.2)); gd.add("w_hh", _rand(P, G, H, seed=3, scale=0.125))
    gd.add("b_hh", _rand(P, G, seed=4, scale=0.2)); gd.add("h0", _rand(P, B, H, seed=5))
    gd.add("w_lin", _rand(P, H, seed=6, scale=0.2)); gd.add("b_lin", _rand(P, seed=7))
    gd.add("hs", shape=(P, T, B, H)); gd.add("ghn", shape=(P, T, B, H)); gd.add("pred", shape=(P, T, B))
    fw = _twice(gd, lambda: k.gru_fwd_ll(gd["gates"], gd["b_ih"], gd["w_hh"], gd["b_hh"], gd["h0"], B * H, gd["w_lin"], gd["b_lin"],
                                         gd["hs"], gd["ghn"], gd["pred"], P, T, B, 0), ("gates", "hs", "ghn", "pred"))
    gb = Guarded()
    gb.add("gates", fw["gates"]); gb.add("ghn", fw["ghn"]); gb.add("hs", fw["hs"]); gb.add("h0", gd["h0"].clone())
    gb.add("w_hh", gd["w_hh"].clone()); gb.add("w_lin", gd["w_lin"].clone()); gb.add("dpred", _rand(P, T, B, seed=8))
    gb.add("dhs", _rand(P, T, B, H, seed=9, scale=0.1))
    for n, s in (("db_hh", (P, G)), ("db_ih", (P, G)), ("dw_lin", (P, H)), ("db_lin", (P,)), ("dh0", (P, B, H))):
        gb.add(n, shape=s)
    gb.add("ws", shape=(k.gru_bwd_workspace(P, B) // 4 + 4,))
    for with_dhs in (False, True):
        _twice(gb, lambda: k.gru_bwd_ll(gb["gates"], gb["ghn"], gb["hs"], gb["h0"], B * H, gb["w_hh"], gb["w_lin"], gb["dpred"], None,
                                        gb["dhs"] if with_dhs else None, gb["db_hh"], gb["db_ih"], gb["dw_lin"], gb["db_lin"], gb["dh0"],
                                        P, T, B, gb["ws"]), ("gates", "ghn", "db_hh", "db_ih", "dw_lin", "db_lin", "dh0"))


@pytest.mark.parametrize("P,T,B,K,t_skip", [(2, 10, 256, 100, 1), (3, 4, 40, 36, 1), (1, 10, 256, 12, 0), (5, 10, 129, 1000, 1), (2, 10, 96, 260, 1)])
def test_guard_projection_tensor_core(P, T, B, K, t_skip):
    k = _k()
    gd = Guarded()
    x = _rand(T, B, K, seed=1)
    x[:t_skip] = 0
    gd.add("x", x); gd.add("x_hi", shape=(T, B, K)); gd.add("x_lo", shape=(T, B, K))
    gd.add("w", _rand(P, G, K, seed=2, scale=0.1)); gd.add("w_hi", shape=(P, G, K)); gd.add("w_lo", shape=(P, G, K))
    gd.add("b", _rand(P, G, seed=3, scale=0.2)); gd.add("gates", shape=(P, T, B, G))

    def fwd():
        k.split_tf32(gd["x"], gd["x_hi"], gd["x_lo"], T * B * K)
        k.split_tf32_gate_rows(gd["w"], gd["w_hi"], gd["w_lo"], P * G, K)
        k.proj_fwd_tc(gd["x_hi"], gd["x_lo"], gd["w_hi"], gd["w_lo"], gd["b"], gd["gates"], P, T, B, K, t_skip)
    out = _twice(gd, fwd, ("x_hi", "x_lo", "w_hi", "w_lo") + (("gates",) if t_skip == 0 else ()))
    if t_skip:                                   # rows of skipped steps are not written by contract
        assert bool(torch.isfinite(gd["gates"][:, t_skip:]).all())
    gw = Guarded()
    gw.add("dg", _rand(P, T, B, G, seed=4)); gw.add("x_hi", out["x_hi"]); gw.add("x_lo", out["x_lo"])
    mask = (torch.rand(P, K, generator=torch.Generator().manual_seed(5)) < 0.6).to(torch.uint8)
    gw.add("mask", mask); gw.add("dw", shape=(P, G, K))
    gw.add("ws", shape=(k.proj_wgrad_tc_workspace(P, T, B, K, t_skip) // 4 + 4,))
    _twice(gw, lambda: k.proj_wgrad_tc(gw["dg"], gw["x_hi"], gw["x_lo"], gw["mask"], gw["dw"], P, T, B, K, t_skip, gw["ws"]), ("dw",))


@pytest.mark.parametrize("P,K", [(4, 10), (100, 100), (3, 33), (2, 1000)])
def test_guard_update_and_loss_kernels(P, K):
    k = _k()
    gd = Guarded()
    gd.add("w", _rand(P, G, K, seed=1, scale=0.05)); gd.add("dw", _rand(P, G, K, seed=2, scale=0.05))
    gd.add("norm", shape=(P, K))
    _twice(gd, lambda: k.gd_prox_gc(gd["w"], gd["dw"], None, gd["norm"], P, K, 0.05, 0.005, 1), ("w", "norm"))
    T, B = 10, 100
    gm = Guarded()
    gm.add("pred", _rand(P, T, B, seed=3)); gm.add("target", _rand(P, T, B, seed=4))
    gm.add("sse", shape=(P,)); gm.add("dpred", shape=(P, T, B)); gm.add("err", shape=(P, T, B))
    _twice(gm, lambda: k.mse_fwd_bwd(gm["pred"], gm["target"], gm["sse"], gm["dpred"], gm["err"], P, T, B), ("sse", "dpred", "err"))


@pytest.mark.parametrize("B", [256, 100, 33])
def test_guard_latent_head(B):
    import vae_connexe_b200.lib as L
    k = _k()
    gd = Guarded()
    gd.add("hT", _rand(B, H, seed=1)); gd.add("lat_w", _rand(2 * H, H, seed=2, scale=0.1)); gd.add("lat_b", _rand(2 * H, seed=3, scale=0.1))
    gd.add("eps", _rand(B, H, seed=4)); gd.add("lat", shape=(B, 2 * H)); gd.add("z", shape=(B, H)); gd.add("kl", shape=(1,))
    gd.add("ws", torch.zeros(k.latent_head_workspace(B) // 4 + 4))      # holds the kernel's self-resetting ticket counter: zero by contract
    out = _twice(gd, lambda: k.latent_head_fwd(gd["hT"], gd["lat_w"], gd["lat_b"], gd["eps"], gd["lat"], gd["z"], gd["kl"], B, L.KL_SWAPPED,
                                               gd["ws"]), ("lat", "z", "kl"))
    gb = Guarded()
    gb.add("dlat", _rand(B, 2 * H, seed=5)); gb.add("hT", gd["hT"].clone()); gb.add("lat_w", gd["lat_w"].clone())
    gb.add("d_w", shape=(2 * H, H)); gb.add("d_b", shape=(2 * H,)); gb.add("dhT", shape=(1, B, H))
    _twice(gb, lambda: k.latent_head_bwd(gb["dlat"], gb["hT"], gb["lat_w"], gb["d_w"], gb["d_b"], gb["dhT"], B), ("d_w", "d_b", "dhT"))
