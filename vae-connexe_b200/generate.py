"""Test-mode autoregressive generation (CRVAE.forward(mode='test') :223-243 / :264-284 and
VRAE4E.forward(mode='test') :171-179): 21 one-step updates in which every head's next input is the vector of ALL
heads' previous outputs.

One generator = ONE CUDA-graph replay.  The 21 steps are captured once per (model, batch size, phase) into a graph of
    projection (tcgen05 3xTF32, T = 1)  ->  one recurrent step (per-head hidden state carried in [P,B,H])
    ->  [all-gather of the step's outputs across head shards]  ->  crvae_gen_scatter
(3 kernels per step; the scatter kernel writes the next input row, the stored sequence and the tf32 split in one pass),
so a call costs one h_0 upload, one replay and no per-step Python.  It works on a head shard: each rank advances its own
heads and the step's outputs are all-gathered (NCCL, captured in the graph), which is the one place of the path that
needs a per-step exchange.

Why not a single persistent kernel with a grid barrier: the per-step product x_t . W_ih^T needs every head's W_ih
(tf32 hi/lo: 154 KB per head at p = 100, 25.6 MB in all, with W_hh) on chip; with <= 148 co-resident CTAs that is more
than the SMs' shared memory, so a persistent kernel would stream the weights from L2 every step exactly like the
per-step GEMM does, and an exact-fp32 FFMA kernel that keeps them resident is compute-bound at > 1.3 ms (8 MFMA per
head and step at the measured 64 FMA/clk/SM).  The captured graph keeps the tensor-core kernels and pays ~2 us per
kernel boundary instead.

The RNG contract is the reference's: one torch.randn(1, B, H) on the CPU default generator per call (:225 / :266 / :173).
"""
from __future__ import annotations

import torch

from . import rec as R
from .engine import G, H

GEN_STEPS = 21          # int(20/1)+1  (:229, :174)


class _CrvaePlan:
    """Static buffers + the captured graph of one (model, B, phase) generator."""

    def __init__(self, model, B: int, phase: int):
        eng, k = model.engine, model.engine.k
        self.model, self.eng, self.k, self.B, self.phase = model, eng, k, B, phase
        P, p, dev = eng.P, eng.p, eng.device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.h0 = z(B, H)
        self.h = [z(max(P, 1), B, H), z(max(P, 1), B, H)]
        self.x = z(1, B, p)
        self.tc = eng.proj_mode == "tc3"
        self.x_hi, self.x_lo = (z(1, B, p), z(1, B, p)) if self.tc else (None, None)
        self.gates, self.ghn, self.pred = z(max(P, 1), 1, B, G), z(max(P, 1), 1, B, H), z(max(P, 1), 1, B)
        self.xg = z(P, 1, B, eng.Kw) if eng.packed else None       # gather-packed heads: every head's own input columns
        self.out = z(B, GEN_STEPS, p)
        self.noise = z(B, GEN_STEPS, p) if phase == 1 else None
        self.world = model.world_size
        self.base, self.rem = divmod(p, self.world)
        self.widest = self.base + (1 if self.rem else 0)
        if self.world > 1:
            self.y_pad = z(self.widest, B)                       # this rank's outputs, padded to the widest shard
            self.y_all = z(self.world, self.widest, B)
        self.graph = None

    # one generation step; `i` selects the ping-pong hidden-state buffers
    def _step(self, i: int):
        eng, k, th, P, p, B = self.eng, self.k, self.eng.theta, self.eng.P, self.eng.p, self.B
        h, h_next = self.h[i & 1], self.h[(i + 1) & 1]
        if P > 0:
            if eng.packed:
                k.gather_cols(self.x, eng.cols, eng.mask_u8, self.xg, P, B, p, eng.Kw)
                k.proj_fwd_packed(self.xg, th["w_ih"], th["b_ih"], self.gates, P, 1, B, eng.Kw, 0)
            elif self.tc:
                k.proj_fwd_tc(self.x_hi, self.x_lo, eng.w_ih_hi, eng.w_ih_lo, th["b_ih"], self.gates, P, 1, B, p, 0)
            else:
                k.proj_fwd(self.x, th["w_ih"], th["b_ih"], self.gates, P, 1, B, p, 0)
            if hasattr(k, "gru_fwd_tc") and P >= 8 and not (R.has_ll(k) and P <= R.LL_MAX_HEADS):
                k.gru_fwd_tc(self.gates, th["b_ih"], th["w_hh"], None, th["b_hh"], h, B * H, th["w_lin"], th["b_lin"],
                             h_next.view(-1, 1, B, H), self.ghn, self.pred, P, 1, B, 0)
            elif R.has_ll(k):
                k.gru_fwd_ll(self.gates, th["b_ih"], th["w_hh"], th["b_hh"], h, B * H, th["w_lin"], th["b_lin"],
                             h_next.view(-1, 1, B, H), self.ghn, self.pred, P, 1, B, 0)
            else:
                k.gru_fwd(self.gates, th["b_ih"], th["w_hh"], th["b_hh"], h, B * H, th["w_lin"], th["b_lin"],
                          h_next.view(-1, 1, B, H), self.ghn, self.pred, P, 1, B, 0)
        y = self.pred.view(-1, B)
        if self.world > 1:                                          # the per-step all-gather of the heads' outputs (:233-236)
            self.y_pad[:P].copy_(y[:P])
            self._allgather()
            y = self.y_all
        # X_t = cat(out_j) over heads, + 0.1 * noise[:, i] in phase 1 (:281-283); it is both stored and fed back (:232)
        k.gen_scatter(y, self.noise, self.x, self.x_hi, self.x_lo, self.out, B, p, i, GEN_STEPS, self.base, self.rem, self.widest, 0.1)

    def _allgather(self):
        import torch.distributed as dist
        from .sharding import _stage_through_host
        group = self.model.group
        if _stage_through_host(self.y_pad, group):                  # gloo test configuration (not capturable)
            host = self.y_pad.cpu()
            parts = [torch.empty_like(host) for _ in range(self.world)]
            dist.all_gather(parts, host, group=group)
            self.y_all.copy_(torch.stack(parts).to(self.y_all.device))
        else:
            dist.all_gather_into_tensor(self.y_all, self.y_pad, group=group)

    def _sequence(self):
        eng, k = self.eng, self.k
        P, B = eng.P, self.B
        self.x.zero_()                                              # X_seq starts as one zero step (:224)
        if self.tc:
            self.x_hi.zero_(); self.x_lo.zero_()
            if P > 0:
                k.split_tf32_gate_rows(eng.theta["w_ih"], eng.w_ih_hi, eng.w_ih_lo, P * G, eng.p)
        self.h[0].copy_(self.h0.unsqueeze(0).expand_as(self.h[0]))  # every head starts from the same h_0 (:227-228)
        for i in range(GEN_STEPS):
            self._step(i)

    def run(self, h0_host: torch.Tensor, noise) -> torch.Tensor:
        dev = self.eng.device
        self.h0.copy_(h0_host.reshape(self.B, H).to(dev, non_blocking=True))
        if self.phase == 1:
            self.noise.copy_(noise.to(dev, torch.float32)[:, :GEN_STEPS, :])
        capturable = dev.type == "cuda" and (self.world == 1 or not self._gloo())
        if not capturable:
            self._sequence()
        else:
            if self.graph is None:
                self._sequence()                                    # warm-up (kernel attributes, allocator) outside the capture
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._sequence()
                self.graph = g
            self.graph.replay()
        return self.out.clone()

    def _gloo(self) -> bool:
        import torch.distributed as dist
        return dist.get_backend(self.model.group) == "gloo"


def crvae_generate(model, X, noise=None, phase=0, h0=None):
    """Returns X_seq (B, 21, p) like the reference.  phase=1 adds 0.1*noise[:, i] to every generated
    step (:281-283).  Draws h_0 ~ N(0,1) of size (1,B,H) on the CPU generator (:225 / :266) unless the caller
    passes the draw it already made (the trainers draw their noise in blocks, train._NoiseFeed)."""
    B = X.shape[0]
    plans = model.__dict__.setdefault("_gen_plans", {})
    key = (B, int(phase))
    plan = plans.get(key)
    if plan is None:
        plan = plans[key] = _CrvaePlan(model, B, int(phase))
    if h0 is None:
        h0 = torch.randn(size=(1, B, H))
    return plan.run(h0, noise)


class _VraePlan:
    def __init__(self, model, B: int):
        eng = model.engine
        self.eng, self.k, self.B = eng, eng.k, B
        p, dev = eng.p, eng.device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.h0, self.h = z(B, H), [z(1, B, H), z(1, B, H)]
        self.x, self.gates, self.ghn, self.y = z(1, B, p), z(1, 1, B, G), z(1, 1, B, H), z(B, p)
        self.out = z(B, GEN_STEPS + 1, p)
        self.graph = None

    def _sequence(self):
        from . import lib as L
        eng, k, th, B, p = self.eng, self.k, self.eng.theta, self.B, self.eng.p
        self.x.zero_()
        self.h[0][0].copy_(self.h0)
        for i in range(GEN_STEPS):
            h, h_next = self.h[i & 1], self.h[(i + 1) & 1]
            k.proj_fwd(self.x, th["dec_w_ih"].view(1, G, p), th["dec_b_ih"], self.gates, 1, 1, B, p, 0)
            R.gru_forward_small(k, self.gates, th["dec_b_ih"], th["dec_w_hh"], th["dec_b_hh"], h, 0, None, None,
                                h_next.view(1, 1, B, H), self.ghn, None, 1, 1, B, 0)
            k.gemm(L.GEMM_NT, 1, B, p, H, h_next, H, 0, th["out_w"], H, 0, self.y, p, 0, th["out_b"], 0)
            self.out[:, i + 1, :].copy_(self.y)
            self.x[0].copy_(self.y)

    def run(self, h0_host):
        dev = self.eng.device
        self.h0.copy_(h0_host.reshape(self.B, H).to(dev, non_blocking=True))
        if dev.type != "cuda":
            self._sequence()
        else:
            if self.graph is None:
                self._sequence()
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._sequence()
                self.graph = g
            self.graph.replay()
        return self.out.clone()


def vrae_generate(model, X):
    """VRAE4E test mode (:171-179): X_seq (B, 22, p): the zero step followed by 21 generated steps."""
    B = X.shape[0]
    plans = model.__dict__.setdefault("_gen_plans", {})
    plan = plans.get(B)
    if plan is None:
        plan = plans[B] = _VraePlan(model, B)
    return plan.run(torch.randn(size=(1, B, H)))                  # :173
