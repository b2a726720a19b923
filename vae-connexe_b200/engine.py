"""Fused multi-head recurrent engine: the storage and kernel orchestration behind CRVAE / VRAE4E.

All p per-variable decoder GRUs of the reference (`networks[i]`, CRVAE_lorenz96.py:200-201) live in
ONE set of fused buffers and are advanced by ONE kernel per stage:

  theta / grad arenas (flat fp32, identical layout; snapshot = one device copy):
    w_ih [P,G,p] | w_hh [P,G,H] | b_ih [P,G] | b_hh [P,G] | w_lin [P,H] | b_lin [P]     heads (shard)
    enc_w_ih [G,p] | enc_w_hh [G,H] | enc_b_ih [G] | enc_b_hh [G] | lat_w [2H,H] | lat_b [2H]   encoder
  (lat_w = [fc_mu.weight ; fc_std.weight], so mu|log_var come out of one GEMM)

  activations for a bound batch (B rows):
    gates [P,Td,B,G]  (gi -> r|z|n -> dgi, reused in place through projection/forward/backward)
    hs, ghn [P,Td,B,H]; pred, dpred [P,Td,B]; dh0 [P,B,H]; encoder twins with P=1, T=Te.

P is the number of heads THIS rank holds (head shard [head_off, head_off+P) of p); the encoder
is replicated.  The only data-path collective is the sum over ranks of dz = sum_heads dh0.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import lib as L
from . import rec as R

H = 64
G = 3 * H
ENC_STEPS = 10   # CRVAE_lorenz96.py:208
DEC_STEPS = 10   # CRVAE_lorenz96.py:119 / :484 (context 20)

HEAD_FIELDS = ("w_ih", "w_hh", "b_ih", "b_hh", "w_lin", "b_lin")
ENC_FIELDS = ("enc_w_ih", "enc_w_hh", "enc_b_ih", "enc_b_hh", "lat_w", "lat_b")


def _f32(x: float) -> float:
    return float(np.float32(x))


class Arena:
    """Flat fp32 buffer with named views."""

    def __init__(self, shapes: Dict[str, tuple], device):
        self.shapes = shapes
        self.offsets = {}
        off = 0
        for k, s in shapes.items():
            off = (off + 3) // 4 * 4                     # keep every field 16-byte aligned
            self.offsets[k] = off
            off += int(np.prod(s))
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.views = {k: self.flat[self.offsets[k]:self.offsets[k] + int(np.prod(s))].view(*s)
                      for k, s in shapes.items()}

    def __getitem__(self, k):
        return self.views[k]

    def like(self):
        return Arena(self.shapes, self.flat.device)


class CRVAEEngine:
    """Kernel-level implementation of CRVAE.forward(mode='train') (:203-221), the trainer's loss
    (:484-489), backward (:497), GD (:498-499) and prox (:502-504) for a head shard."""

    def __init__(self, p: int, mask: np.ndarray, head_off: int = 0, device="cuda", group=None, comm=None, packed=None):
        self.k = L.kernels()
        self.p = int(p)
        self.P = int(mask.shape[0])
        self.head_off = int(head_off)
        self.device = torch.device(device)
        self.group = group
        # comm: the object that sums dz over the head shards (sharding.TorchComm by default when a group is given; tests
        # inject their own to emulate the other ranks on one GPU)
        if comm is None and group is not None:
            from .sharding import TorchComm
            comm = TorchComm(group)
        self.comm = comm
        assert mask.shape == (self.P, self.p)
        self.mask_np = np.ascontiguousarray(mask.astype(bool))
        self.dense = bool(self.mask_np.all())
        P, p_ = self.P, self.p
        # Pruned (phase-2) heads, two storage forms (:115, :200-201, :788-790):
        #   masked-dense  w_ih [P,G,p] with structural zeros and a [P,p] mask (any graph; keeps the tensor-core projection)
        #   gather-packed w_ih [P,G,Kp], Kp = widest head's input count rounded up to 4, with the per-head column lists
        #                 `cols` and the gathered input dec_in_g [P,Td,B,Kp]: weights, gradients, projection flops and
        #                 prox traffic scale with the graph's in-degree, not with p.  Bit-identical to the exact
        #                 masked-dense form.  Chosen by default when every head reads at most a quarter of the series AND
        #                 p >= 256: measured phase-2 iteration, Lorenz-96 ring (k = 4), B = 256 -- p = 1000: 13.5 ms
        #                 masked-dense (768 MB of W_ih) vs 7.4 ms packed (3 MB); p = 100: 1.10 vs 1.23 ms (at K = 100 the
        #                 tensor-core projection on structural zeros still beats the exact FFMA GEMM on K = 4).
        kmax = int(self.mask_np.sum(1).max()) if P > 0 else 0
        if packed is None:
            import os as _os0
            packed = ((not self.dense) and P > 0 and 4 * kmax <= p_ and p_ >= 256
                      and _os0.environ.get("CRVAE_PACKED", "1") != "0")
        self.packed = bool(packed) and hasattr(self.k, "proj_fwd_packed") and not self.dense and P > 0
        if self.packed:
            self.Kw = Kp = max(4, (kmax + 3) // 4 * 4)
            cols = np.zeros((P, Kp), dtype=np.int32)
            pmask = np.zeros((P, Kp), dtype=np.uint8)
            for i in range(P):
                c = np.where(self.mask_np[i])[0]                     # ascending = the reference's np.where order (:115)
                cols[i, :len(c)] = c
                pmask[i, :len(c)] = 1
            self.cols_np, self.pmask_np = cols, pmask.astype(bool)
            self.cols = torch.from_numpy(cols).to(self.device)
            self.mask_u8 = torch.from_numpy(pmask).to(self.device)
        else:
            self.Kw = p_
            self.mask_u8 = None if self.dense else torch.from_numpy(self.mask_np.astype(np.uint8)).to(self.device)
        shapes = {
            "w_ih": (P, G, self.Kw), "w_hh": (P, G, H), "b_ih": (P, G), "b_hh": (P, G), "w_lin": (P, H), "b_lin": (P,),
            "enc_w_ih": (G, p_), "enc_w_hh": (G, H), "enc_b_ih": (G,), "enc_b_hh": (G,),
            "lat_w": (2 * H, H), "lat_b": (2 * H,),
        }
        self.theta = Arena(shapes, self.device)
        self.grad = self.theta.like()
        self.n_wih = P * G * self.Kw
        self.rest_off = self.theta.offsets["w_hh"]
        self.col_norm = torch.zeros(P, self.Kw, dtype=torch.float32, device=self.device)
        self.col_norm_dense = torch.zeros(P, p_, dtype=torch.float32, device=self.device) if self.packed else self.col_norm
        # projection mode: "tc3" = tcgen05 3xTF32 (needs p % 4 == 0 for the TMA row pitch), "exact" = FFMA fp32,
        # "packed" = gather-packed ragged heads (exact FFMA on Kp columns); the encoder (always dense) has its own flag
        import os as _os1
        tc_ok = p_ % 4 == 0 and hasattr(self.k, "proj_wgrad_tc") and _os1.environ.get("CRVAE_PROJ_MODE", "tc3") == "tc3"
        self.enc_tc = tc_ok
        self.proj_mode = "packed" if self.packed else ("tc3" if tc_ok else "exact")
        zl = lambda t: torch.zeros_like(t)
        if self.proj_mode == "tc3":
            self.w_ih_hi, self.w_ih_lo = zl(self.theta["w_ih"]), zl(self.theta["w_ih"])
        if self.enc_tc:
            self.enc_w_hi, self.enc_w_lo = zl(self.theta["enc_w_ih"]), zl(self.theta["enc_w_ih"])
        # recurrence mode: "tc3" = tcgen05 gate GEMM (crvae_gru_fwd_tc, 3xTF32; default once a rank holds enough heads
        # to fill the SMs with 128-row tiles), "exact" = persistent fp32 FFMA kernel (CRVAE_REC_MODE=exact, small shards)
        # "ll" = low-latency 16-row tiles (crvae_gru_fwd_ll / _bwd_ll): shards of few heads, where one step's latency is the cost
        import os as _os
        want = _os.environ.get("CRVAE_REC_MODE", "auto")
        # "mma" = register-resident warp-level MMA kernels (crvae_gru_fwd_mma / _bwd_mma, 3xTF32): 16-row tiles, persistent grid
        if R.has_mma(self.k) and P > 0 and (want == "mma" or (want == "auto" and R.mma_preferred(P))):
            self.rec_mode = "mma"
        elif R.has_ll(self.k) and P > 0 and (want == "ll" or (want == "auto" and P <= R.LL_MAX_HEADS)):
            self.rec_mode = "ll"
        elif want in ("auto", "tc3") and hasattr(self.k, "gru_fwd_tc") and P >= 8:
            self.rec_mode = "tc3"
        else:
            self.rec_mode = "exact"
        self.B = None
        self.bind_serial = 0             # bumped by every bind_batch: lets autograd nodes detect a re-bound batch
        self.kl_form = L.KL_SWAPPED
        self._side = None
        self.use_side_stream = True
        import os as _os2
        self.bwd_mode = _os2.environ.get("CRVAE_BWD_MODE", "defer")     # "defer" (dW_hh on tcgen05) | "fused" (in-kernel FFMA)

    # ------------------------------------------------------------------ batch binding
    def bind_batch(self, X: torch.Tensor):
        """X (B, 20, p) on the device: the (fixed, :470-473) training batch.  Pre-arranges the
        encoder input X[:,0:10] (:208), decoder input [0, X[:,10:19]] (:119) and the per-head
        targets X[:,10:,i] (:484) in the time-major layouts the kernels stream."""
        assert X.dim() == 3 and X.shape[2] == self.p and X.shape[1] == ENC_STEPS + DEC_STEPS
        X = X.to(self.device, torch.float32)
        B = X.shape[0]
        self.bind_serial += 1
        lo, hi = self.head_off, self.head_off + self.P
        if self.B != B:
            self._alloc(B)
            self.enc_in = torch.empty(ENC_STEPS, B, self.p, dtype=torch.float32, device=self.device)
            self.dec_in = torch.zeros(DEC_STEPS, B, self.p, dtype=torch.float32, device=self.device)
            self.target = torch.empty(max(self.P, 1), DEC_STEPS, B, dtype=torch.float32, device=self.device)
        # in-place (re)binding keeps the buffer addresses stable for captured CUDA graphs
        etc, dtc = self.enc_tc, self.proj_mode == "tc3"
        if etc and (getattr(self, "enc_in_hi", None) is None or self.enc_in_hi.shape != self.enc_in.shape):
            self.enc_in_hi, self.enc_in_lo = torch.zeros_like(self.enc_in), torch.zeros_like(self.enc_in)
        if dtc and (getattr(self, "dec_in_hi", None) is None or self.dec_in_hi.shape != self.dec_in.shape):
            self.dec_in_hi, self.dec_in_lo = torch.zeros_like(self.dec_in), torch.zeros_like(self.dec_in)
        if self.packed and (getattr(self, "dec_in_g", None) is None or self.dec_in_g.shape[2] != B):
            self.dec_in_g = torch.zeros(self.P, DEC_STEPS, B, self.Kw, dtype=torch.float32, device=self.device)
        if hasattr(self.k, "bind_batch") and X.is_contiguous():      # one pass: transposes + tf32 splits + targets
            self.k.bind_batch(X, self.enc_in, self.enc_in_hi if etc else None, self.enc_in_lo if etc else None, self.dec_in,
                              self.dec_in_hi if dtc else None, self.dec_in_lo if dtc else None, self.target, B, self.p,
                              ENC_STEPS, DEC_STEPS, lo, self.P)
        else:
            self.enc_in.copy_(X[:, :ENC_STEPS].transpose(0, 1))                               # [Te,B,p]
            self.dec_in[1:].copy_(X[:, ENC_STEPS:-1].transpose(0, 1))                         # [Td,B,p], step 0 stays 0
            if self.P > 0:
                self.target[: self.P].copy_(X[:, ENC_STEPS:, lo:hi].permute(2, 1, 0))         # [P,Td,B]
            if etc:     # the batch is fixed (:470-473): split it into tf32 hi/lo once
                self.k.split_tf32(self.enc_in, self.enc_in_hi, self.enc_in_lo, self.enc_in.numel())
            if dtc:
                self.k.split_tf32(self.dec_in, self.dec_in_hi, self.dec_in_lo, self.dec_in.numel())
        if self.packed:     # every head's own input columns (:115), gathered once per batch
            self.k.gather_cols(self.dec_in, self.cols, self.mask_u8, self.dec_in_g, self.P, DEC_STEPS * B, self.p, self.Kw)

    def _alloc(self, B: int):
        P, dev = self.P, self.device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.B = B
        if self.comm is not None and self.group is not None:      # the fused peer-memory dz all-reduce needs the batch size
            from .sharding import make_dz_comm
            self.comm = make_dz_comm(self.k, self.group, B, H, dev, self.comm)
        self.gates = z(max(P, 1), DEC_STEPS, B, G)
        self.hs = z(max(P, 1), DEC_STEPS, B, H)
        self.ghn = z(max(P, 1), DEC_STEPS, B, H)
        self.pred = z(max(P, 1), DEC_STEPS, B)
        self.dpred = z(max(P, 1), DEC_STEPS, B)
        self.err = z(max(P, 1), DEC_STEPS, B)
        self.err_tbp = z(DEC_STEPS, B, self.p)       # residual for all series, time-major (VRAE4E input)
        self.dh0 = z(max(P, 1), B, H)
        self.sse = z(max(P, 1))
        self.enc_gates = z(1, ENC_STEPS, B, G)
        self.enc_hs = z(1, ENC_STEPS, B, H)
        self.enc_ghn = z(1, ENC_STEPS, B, H)
        self.enc_dh0 = z(1, B, H)
        self.h0_zero = z(B, H)
        self.lat = z(B, 2 * H)
        self.dlat = z(B, 2 * H)
        self.zlat = z(B, H)
        self.eps = z(B, H)
        self.eps_next = z(B, H)      # staging slot: the next forward's noise (backward still needs self.eps)
        self.dhT = z(1, B, H)
        self.dz_part = z(B, H)
        self.ones_B = torch.ones(B, 1, dtype=torch.float32, device=dev)
        self.kl = z(1)
        self.loss = z(1)
        self.scratch = z(4)
        k = self.k
        nbytes = max(k.gru_bwd_workspace(max(P, 1), B), k.gru_bwd_workspace(1, B))
        self.ws_gru = torch.zeros(nbytes // 4 + 4, dtype=torch.float32, device=dev)
        self.ws_gru_enc = torch.zeros(k.gru_bwd_workspace(1, B) // 4 + 4, dtype=torch.float32, device=dev)
        nbytes = max(k.proj_wgrad_workspace(max(P, 1), DEC_STEPS, B, self.p),
                     k.proj_wgrad_workspace(1, ENC_STEPS, B, self.p))
        if self.packed:
            nbytes = max(nbytes, k.proj_wgrad_packed_workspace(P, DEC_STEPS, B, self.Kw, self.p))
        self.ws_wgrad = torch.zeros(nbytes // 4 + 4, dtype=torch.float32, device=dev)
        self.ws_wgrad_dec = torch.zeros(nbytes // 4 + 4, dtype=torch.float32, device=dev)   # side-stream twin
        self.ws_lat = None
        if hasattr(k, "latent_head_fwd"):
            self.ws_lat = torch.zeros(k.latent_head_workspace(B) // 4 + 4, dtype=torch.float32, device=dev)
        n = R.dwhh_workspace(k, 1, ENC_STEPS, B)
        self.ws_dwhh_enc = torch.zeros(n, dtype=torch.float32, device=dev) if n else None
        self.ws_wgrad_tc = self.ws_dwhh = self.ws_wgrad_tc_enc = None
        if hasattr(k, "proj_wgrad_tc_workspace"):
            self.ws_wgrad_tc_enc = torch.zeros(k.proj_wgrad_tc_workspace(1, ENC_STEPS, B, self.p, 0) // 4 + 4, dtype=torch.float32, device=dev)
        if hasattr(k, "proj_wgrad_tc_workspace") and P > 0:     # split-reduction partials of the tensor-core gradient GEMMs
            if self.proj_mode == "tc3":
                self.ws_wgrad_tc = torch.zeros(k.proj_wgrad_tc_workspace(P, DEC_STEPS, B, self.p, 1) // 4 + 4, dtype=torch.float32, device=dev)
            self.ws_dwhh = torch.zeros(k.gru_dwhh_tc_workspace(P, DEC_STEPS, B) // 4 + 4, dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ forward
    def forward(self, eps: Optional[torch.Tensor] = None, want_err: bool = False):
        """Everything of :204-221 + :484-486 for the bound batch.  eps (B,H) device tensor (the
        reference's CPU-generator draw, already uploaded); None = reuse self.eps."""
        k, th, B, P, p_ = self.k, self.theta, self.B, self.P, self.p
        if eps is not None:
            self.eps.copy_(eps.reshape(B, H), non_blocking=True)
        # The decoder heads' input projection does not depend on the encoder.  The (latency-bound,
        # replicated) encoder chain runs on a HIGH-PRIORITY side stream so its few CTAs are scheduled
        # as soon as an SM frees up, while the big projection GEMM fills the machine from the main stream.
        side = self._fork()
        with self._on(side):
            # encoder GRU (gru_left, :208) -> h_T
            self._project(self.enc_in, "enc", th["enc_w_ih"], th["enc_b_ih"], self.enc_gates, 1, ENC_STEPS, 0)
            R.gru_forward_small(k, self.enc_gates, th["enc_b_ih"], th["enc_w_hh"], th["enc_b_hh"], self.h0_zero, 0, None, None,
                                self.enc_hs, self.enc_ghn, None, 1, ENC_STEPS, B, 0)
            hT = self.enc_hs[0, ENC_STEPS - 1]
            # [mu | log_var] = h_T [fc_mu ; fc_std]^T + b (:210-211); z = mu + exp(.5 lv) eps (:213-216); KL (:486)
            if self.ws_lat is not None:     # fused: one launch instead of GEMM + pointwise on the latency-bound chain
                k.latent_head_fwd(hT, th["lat_w"], th["lat_b"], self.eps, self.lat, self.zlat, self.kl, B, self.kl_form, self.ws_lat)
            else:
                k.gemm(L.GEMM_NT, 1, B, 2 * H, H, hT, H, 0, th["lat_w"], H, 0, self.lat, 2 * H, 0, th["lat_b"], 0)
                k.latent_fwd(self.lat, self.eps, self.zlat, self.kl, B, self.kl_form)
        if P > 0:
            self._project(self.dec_in, "dec", th["w_ih"], th["b_ih"], self.gates, P, DEC_STEPS, 1)
        self._join(side)
        # decoder heads (:218-219 -> GRU.forward :114-121): recurrence (+Linear), MSE
        if P > 0:
            if self.rec_mode == "tc3":
                k.gru_fwd_tc(self.gates, th["b_ih"], th["w_hh"], None, th["b_hh"], self.zlat, 0, th["w_lin"],
                             th["b_lin"], self.hs, self.ghn, self.pred, P, DEC_STEPS, B, 1)
            elif self.rec_mode in ("ll", "mma"):
                fwd = k.gru_fwd_mma if self.rec_mode == "mma" else k.gru_fwd_ll
                fwd(self.gates, th["b_ih"], th["w_hh"], th["b_hh"], self.zlat, 0, th["w_lin"], th["b_lin"],
                    self.hs, self.ghn, self.pred, P, DEC_STEPS, B, 1)
            else:
                k.gru_fwd(self.gates, th["b_ih"], th["w_hh"], th["b_hh"], self.zlat, 0, th["w_lin"], th["b_lin"],
                          self.hs, self.ghn, self.pred, P, DEC_STEPS, B, 1)
            k.mse_fwd_bwd(self.pred, self.target, self.sse, self.dpred, self.err if want_err else None,
                          P, DEC_STEPS, B)
            k.dot_small(self.sse, P, 1.0 / (DEC_STEPS * B), self.loss)
        else:
            self.loss.zero_()

    # ------------------------------------------------------------------ side stream (fork / join; capturable)
    def _fork(self):
        if self.device.type != "cuda" or not self.use_side_stream:
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device, priority=-1)      # higher than the default stream
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._side.wait_event(ev)
        return self._side

    def _on(self, side):
        import contextlib
        return torch.cuda.stream(side) if side is not None else contextlib.nullcontext()

    def _join(self, side):
        if side is None:
            return
        ev = torch.cuda.Event()
        ev.record(side)
        torch.cuda.current_stream().wait_event(ev)

    def _project(self, x, which, w, b, gates, P, T, t_skip):
        """gates = b + x . w^T for all heads / timesteps: tcgen05 3xTF32 GEMM or exact FFMA GEMM."""
        k = self.k
        if which == "dec" and self.packed:
            k.proj_fwd_packed(self.dec_in_g, w, b, gates, P, T, self.B, self.Kw, t_skip)
            return
        if (self.enc_tc if which == "enc" else self.proj_mode == "tc3"):
            x_hi, x_lo = (self.enc_in_hi, self.enc_in_lo) if which == "enc" else (self.dec_in_hi, self.dec_in_lo)
            w_hi, w_lo = (self.enc_w_hi, self.enc_w_lo) if which == "enc" else (self.w_ih_hi, self.w_ih_lo)
            k.split_tf32_gate_rows(w, w_hi, w_lo, P * G, self.p)      # weights change every iteration; rows permuted for the epilogue
            k.proj_fwd_tc(x_hi, x_lo, w_hi, w_lo, b, gates, P, T, self.B, self.p, t_skip)
        else:
            k.proj_fwd(x, w, b, gates, P, T, self.B, self.p, t_skip)

    def residual(self) -> torch.Tensor:
        """error = X[:,10:,:] - pred for ALL series as [10, B, p] (:599/:639; detached by construction).
        Needs a preceding forward(want_err=True).  On a head shard the rows are all-gathered."""
        err = self.err[: self.P]
        if self.group is not None:
            from .sharding import allgather_rows
            world = torch.distributed.get_world_size(self.group)
            rank = torch.distributed.get_rank(self.group)
            err = allgather_rows(err.reshape(self.P, -1), self.p, rank, world, self.group)
        self.k.transpose(err.reshape(self.p, -1), self.err_tbp, self.p, DEC_STEPS * self.B)
        return self.err_tbp

    def forward_staged(self, want_err: bool = False):
        """forward() on the noise previously staged in self.eps_next (CUDA-graph friendly: the
        copy is part of the captured sequence, after the backward that still reads self.eps)."""
        self.eps.copy_(self.eps_next)
        self.forward(None, want_err)

    # ------------------------------------------------------------------ backward
    def backward(self, beta: float, lam_ridge: float = 0.0, dlat_extra: Optional[torch.Tensor] = None):
        """Gradient of smooth = loss + ridge + beta*KL (:489/:515) into the grad arena (:497)."""
        k, th, g, B, P, p_ = self.k, self.theta, self.grad, self.B, self.P, self.p
        # decoder BPTT.  With enough heads the dW_hh accumulation is deferred to one tcgen05 GEMM per head
        # (crvae_gru_dwhh_tc, issued below next to the projection weight gradient); the BPTT itself runs on tcgen05
        # (crvae_gru_bwd_tc) when the rank holds >= 8 heads, else on the exact FFMA2 kernels.
        defer = (P >= 8 or self.rec_mode in ("ll", "mma")) and B % 32 == 0 and hasattr(k, "gru_dwhh_tc") and self.bwd_mode == "defer"
        if P > 0 and defer and (self.rec_mode in ("ll", "mma") or (self.rec_mode == "tc3" and self._bwd_mma(P))):
            bwd = k.gru_bwd_mma if self._bwd_mma(P) else k.gru_bwd_ll
            bwd(self.gates, self.ghn, self.hs, self.zlat, 0, th["w_hh"], th["w_lin"], self.dpred, None, None,
                g["b_hh"], g["b_ih"], g["w_lin"], g["b_lin"], self.dh0, P, DEC_STEPS, B, self.ws_gru)
        elif P > 0 and defer and self.rec_mode == "tc3" and hasattr(k, "gru_bwd_tc"):
            k.gru_bwd_tc(self.gates, self.ghn, self.hs, self.zlat, 0, th["w_hh"], th["w_lin"], self.dpred, None,
                         g["b_hh"], g["b_ih"], g["w_lin"], g["b_lin"], self.dh0, P, DEC_STEPS, B, self.ws_gru)
        elif P > 0 and defer:
            k.gru_bwd_deferred(self.gates, self.ghn, self.hs, self.zlat, 0, th["w_hh"], th["w_lin"], self.dpred, None, None,
                               g["b_hh"], g["b_ih"], g["w_lin"], g["b_lin"], self.dh0, P, DEC_STEPS, B, self.ws_gru)
        elif P > 0:
            k.gru_bwd(self.gates, self.ghn, self.hs, self.zlat, 0, th["w_hh"], th["w_lin"], self.dpred, None, None,
                      g["w_hh"], g["b_hh"], g["b_ih"], g["w_lin"], g["b_lin"], self.dh0, P, DEC_STEPS, B, self.ws_gru)
        # The latent / encoder backward chain (small, latency-bound) goes to the high-priority side stream;
        # the big projection weight-gradient GEMM, which only needs dgates, stays on the main stream.
        side = self._fork()
        with self._on(side):
            # dz = sum over ALL heads of dh0 (every head's h0 is z, :218)
            self._dz_latent_bwd(beta)
            if dlat_extra is not None:
                k.axpy(self.dlat, dlat_extra, B * 2 * H, 1.0)
            hT = self.enc_hs[0, ENC_STEPS - 1]
            # fc_mu|fc_std: dW = dlat^T hT, db = column sums, dhT = dlat W
            if self.ws_lat is not None:
                k.latent_head_bwd(self.dlat, hT, th["lat_w"], g["lat_w"], g["lat_b"], self.dhT, B)
            else:
                k.gemm(L.GEMM_TN, 1, 2 * H, H, B, self.dlat, 2 * H, 0, hT, H, 0, g["lat_w"], H, 0)
                k.gemm(L.GEMM_TN, 1, 1, 2 * H, B, self.ones_B, 1, 0, self.dlat, 2 * H, 0, g["lat_b"], 2 * H, 0)
                k.gemm(L.GEMM_NN, 1, B, H, 2 * H, self.dlat, 2 * H, 0, th["lat_w"], H, 0, self.dhT, H, 0)
            # encoder BPTT: gradient enters only through h_T
            R.gru_backward_small(k, self.enc_gates, self.enc_ghn, self.enc_hs, self.h0_zero, 0, th["enc_w_hh"], None, None, self.dhT,
                                 None, g["enc_w_hh"].view(1, G, H), g["enc_b_hh"], g["enc_b_ih"], None, None, self.enc_dh0,
                                 1, ENC_STEPS, B, self.ws_gru_enc, self.ws_dwhh_enc)
            if self.enc_tc and self.ws_wgrad_tc_enc is not None:     # tcgen05, reduction split over 16 CTAs
                k.proj_wgrad_tc(self.enc_gates, self.enc_in_hi, self.enc_in_lo, None, g["enc_w_ih"], 1, ENC_STEPS, B, p_, 0,
                                self.ws_wgrad_tc_enc)
            else:
                k.proj_wgrad(self.enc_gates, self.enc_in, None, g["enc_w_ih"], 1, ENC_STEPS, B, p_, 0, self.ws_wgrad)
        if P > 0:
            if defer:
                k.gru_dwhh_tc(self.gates, self.ghn, self.hs, self.zlat, 0, g["w_hh"], P, DEC_STEPS, B, self.ws_dwhh)
            if self.packed:
                k.proj_wgrad_packed(self.gates, self.dec_in_g, self.mask_u8, g["w_ih"], P, DEC_STEPS, B, self.Kw, p_, 1, self.ws_wgrad_dec)
            elif self.proj_mode == "tc3":
                k.proj_wgrad_tc(self.gates, self.dec_in_hi, self.dec_in_lo, self.mask_u8, g["w_ih"], P, DEC_STEPS, B, p_, 1,
                                self.ws_wgrad_tc)
            else:
                k.proj_wgrad(self.gates, self.dec_in, self.mask_u8, g["w_ih"], P, DEC_STEPS, B, p_, 1, self.ws_wgrad_dec)
            if lam_ridge != 0.0:      # d/dW of lam*(|linear.W|^2 + |W_hh|^2), ridge_regularize :321-325
                k.axpy(g["w_hh"], th["w_hh"], P * G * H, 2.0 * lam_ridge)
                k.axpy(g["w_lin"], th["w_lin"], P * H, 2.0 * lam_ridge)
        self._join(side)

    def _dz_latent_bwd(self, beta):
        """dz = sum over ALL heads of dh0 (every head's h0 is z, :218) -> gradient into [mu | log_var].  On a head shard the
        sum crosses ranks: one fused kernel over NVLink peer memory (sharding.SymmComm) or sum -> NCCL all-reduce -> pointwise."""
        k, B, P = self.k, self.B, self.P
        if self.comm is None:
            k.latent_bwd(self.dh0, P, None, self.lat, self.eps, beta, self.kl_form, self.dlat, None, B)
        elif getattr(self.comm, "fused", False):
            hook = getattr(self, "_stage_hook", None)
            if hook is not None:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
            self.comm.latent_bwd(self.dh0 if P > 0 else None, P, self.lat, self.eps, beta, self.kl_form, self.dlat, B)
            if hook is not None:
                e.record(); hook(s, e)
        else:
            k.latent_bwd(self.dh0 if P > 0 else None, P, None, None, None, 0.0, self.kl_form, None, self.dz_part, B)
            self._allreduce_dz()
            k.latent_bwd(None, 0, self.dz_part, self.lat, self.eps, beta, self.kl_form, self.dlat, None, B)

    def _allreduce_dz(self):
        """Sum of dz_part over the head shards (the one data-path collective)."""
        hook = getattr(self, "_stage_hook", None)
        if hook is not None:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
        self.comm.allreduce_dz(self.dz_part)
        if hook is not None:
            e.record()
            hook(s, e)

    # ------------------------------------------------------------------ software-pipelined ("flow") iteration
    # The steady-state iteration is a cycle  pre -> rec -> update -> pre -> ...  with
    #     pre    = encoder chain (-> z) + per-head projection gi          (needs the updated weights)
    #     rec    = decoder recurrence + MSE                              (needs gi and z)
    #     update = BPTT, weight gradients, GD, prox                      (needs rec's activations)
    # The only full barrier of the cycle sits between pre and rec (z depends on every head's dh0 through the encoder
    # update).  flow_body() runs  rec ; update ; pre  as ONE unit in which the heads are cut into groups, each on its own
    # stream:  recurrence(g) -> MSE(g) -> BPTT(g) -> dW_hh(g), dW_ih(g) -> prox(g) -> projection(g).  With 200 tiles on
    # 148 SMs the recurrent kernels leave a tail wave of 52 tiles; grouped, the forward's tail overlaps the BPTT of the
    # first group and the BPTT's tail overlaps the first group's gradient GEMMs, and every group's next projection
    # overlaps the (latency-bound) encoder backward / forward chain on the side stream.  Captured into a CUDA graph by
    # train.Phase1Runner; bit-identical to backward() + step() + forward() (same kernels on the same data).
    def flow_supported(self) -> bool:
        k = self.k
        return (self.device.type == "cuda" and self.P > 0 and self.B is not None and self.B % 32 == 0 and not self.packed
                and self.proj_mode == "tc3" and self.rec_mode in ("tc3", "ll", "mma") and self.bwd_mode == "defer"
                and hasattr(k, "gru_dwhh_tc") and self.ws_lat is not None and self.use_side_stream)

    def _auto_grouped(self) -> bool:
        """Two equal head groups whenever the 128-row recurrent grid is more than one wave but not many: measured at p = 100
        (200 tiles, ms per iteration): one group 0.612, 74+26 0.549, 50+50 0.523, 34+33+33 0.543, 4x25 0.551, 5x20 0.557."""
        tiles = ((self.B or 0) + 127) // 128
        return self.rec_mode == "tc3" and self.P >= 16 and 148 < self.P * tiles <= 4 * 148

    def _bwd_mma(self, n_heads: int) -> bool:
        """The K-split MMA BPTT for a launch over n_heads heads?  Not where the flow runs head groups side by side: its
        persistent grid would take every SM and serialise them (p = 100, two groups of 50: 0.537 ms per iteration against
        0.516 with the tcgen05 BPTT).  One decision for the eager and the captured iteration."""
        import os as _os
        grouped = self._auto_grouped() if _os.environ.get("CRVAE_GROUPS", "auto") == "auto" else True
        return R.mma_bwd_preferred(self.k, n_heads, self.B) and not (self.rec_mode == "tc3" and grouped)

    def _flow_setup(self):
        if getattr(self, "_flow", None) is not None and self._flow["B"] == self.B:
            return self._flow
        import os as _os
        P, B, k, dev = self.P, self.B, self.k, self.device
        spec = _os.environ.get("CRVAE_GROUPS", "auto")
        if spec == "auto":
            sizes = [P // 2, P - P // 2] if self._auto_grouped() else [P]
        else:
            sizes = [int(x) for x in spec.split(",") if x]
            if sum(sizes) != P or min(sizes) <= 0:
                sizes = [P]
        groups, lo = [], 0
        for n in sizes:
            groups.append((lo, lo + n)); lo += n
        zf = lambda n: torch.zeros(n // 4 + 4, dtype=torch.float32, device=dev)
        ws = [dict(gru=zf(k.gru_bwd_workspace(hi - lo, B)), dwhh=zf(k.gru_dwhh_tc_workspace(hi - lo, DEC_STEPS, B)),
                   wgrad=zf(k.proj_wgrad_tc_workspace(hi - lo, DEC_STEPS, B, self.p, 1))) for lo, hi in groups]
        streams = [None] + [torch.cuda.Stream(device=dev) for _ in groups[1:]]
        self._flow = dict(B=B, groups=groups, ws=ws, streams=streams)
        return self._flow

    def _flow_rec_group(self, lo, hi):
        """Decoder recurrence + MSE of heads [lo, hi)."""
        k, th, B = self.k, self.theta, self.B
        n = hi - lo
        sl = slice(lo, hi)
        if self.rec_mode == "tc3":
            k.gru_fwd_tc(self.gates[sl], th["b_ih"][sl], th["w_hh"][sl], None, th["b_hh"][sl], self.zlat, 0, th["w_lin"][sl],
                         th["b_lin"][sl], self.hs[sl], self.ghn[sl], self.pred[sl], n, DEC_STEPS, B, 1)
        else:
            fwd = k.gru_fwd_mma if self.rec_mode == "mma" else k.gru_fwd_ll
            fwd(self.gates[sl], th["b_ih"][sl], th["w_hh"][sl], th["b_hh"][sl], self.zlat, 0, th["w_lin"][sl], th["b_lin"][sl],
                self.hs[sl], self.ghn[sl], self.pred[sl], n, DEC_STEPS, B, 1)
        k.mse_fwd_bwd(self.pred[sl], self.target[sl], self.sse[sl], self.dpred[sl], None, n, DEC_STEPS, B)

    def _flow_bwd_group(self, lo, hi, ws):
        k, th, g, B = self.k, self.theta, self.grad, self.B
        n, sl = hi - lo, slice(lo, hi)
        fn = k.gru_bwd_tc if (self.rec_mode == "tc3" and not self._bwd_mma(n)) else None
        if fn is not None:
            fn(self.gates[sl], self.ghn[sl], self.hs[sl], self.zlat, 0, th["w_hh"][sl], th["w_lin"][sl], self.dpred[sl], None,
               g["b_hh"][sl], g["b_ih"][sl], g["w_lin"][sl], g["b_lin"][sl], self.dh0[sl], n, DEC_STEPS, B, ws["gru"])
        else:
            bwd = k.gru_bwd_mma if self._bwd_mma(n) else k.gru_bwd_ll
            bwd(self.gates[sl], self.ghn[sl], self.hs[sl], self.zlat, 0, th["w_hh"][sl], th["w_lin"][sl], self.dpred[sl], None, None,
                g["b_hh"][sl], g["b_ih"][sl], g["w_lin"][sl], g["b_lin"][sl], self.dh0[sl], n, DEC_STEPS, B, ws["gru"])

    def _flow_wgrads_group(self, lo, hi, ws, lam_ridge):
        """dW_hh, dW_ih (+ ridge) of heads [lo, hi)."""
        k, th, g, B, p_ = self.k, self.theta, self.grad, self.B, self.p
        n, sl = hi - lo, slice(lo, hi)
        mask = None if self.mask_u8 is None else self.mask_u8[sl]
        k.gru_dwhh_tc(self.gates[sl], self.ghn[sl], self.hs[sl], self.zlat, 0, g["w_hh"][sl], n, DEC_STEPS, B, ws["dwhh"])
        k.proj_wgrad_tc(self.gates[sl], self.dec_in_hi, self.dec_in_lo, mask, g["w_ih"][sl], n, DEC_STEPS, B, p_, 1, ws["wgrad"])
        if lam_ridge != 0.0:
            k.axpy(g["w_hh"][sl], th["w_hh"][sl], n * G * H, 2.0 * lam_ridge)
            k.axpy(g["w_lin"][sl], th["w_lin"][sl], n * H, 2.0 * lam_ridge)

    def _flow_prox_proj_group(self, lo, hi, lr, lam, with_proj):
        """GD + prox of w_ih, GD of b_ih, then (with_proj) the NEXT iteration's projection of heads [lo, hi)."""
        k, th, g, B, p_ = self.k, self.theta, self.grad, self.B, self.p
        n, sl = hi - lo, slice(lo, hi)
        mask = None if self.mask_u8 is None else self.mask_u8[sl]
        k.gd_prox_gc(th["w_ih"][sl], g["w_ih"][sl], mask, self.col_norm[sl], n, self.Kw, _f32(lr), _f32(lam * lr), lam > 0)
        k.gd_step(th["b_ih"][sl], g["b_ih"][sl], n * G, _f32(lr))         # the projection adds b_ih: updated per group, not by the big GD
        if with_proj:
            k.split_tf32_gate_rows(th["w_ih"][sl], self.w_ih_hi[sl], self.w_ih_lo[sl], n * G, p_)
            k.proj_fwd_tc(self.dec_in_hi, self.dec_in_lo, self.w_ih_hi[sl], self.w_ih_lo[sl], th["b_ih"][sl], self.gates[sl], n, DEC_STEPS,
                          B, p_, 1)

    def _flow_gd_rest(self, lr):
        """GD on everything but w_ih and the heads' b_ih (both updated per group): w_hh | [b_ih] | b_hh ... encoder."""
        k, th, g = self.k, self.theta, self.grad
        o_whh, o_bih, o_bhh = self.theta.offsets["w_hh"], self.theta.offsets["b_ih"], self.theta.offsets["b_hh"]
        k.gd_step(th.flat[o_whh:o_bih], g.flat[o_whh:o_bih], o_bih - o_whh, _f32(lr))
        k.gd_step(th.flat[o_bhh:], g.flat[o_bhh:], self.theta.numel - o_bhh, _f32(lr))

    def _enc_forward_chain(self):
        k, th, B = self.k, self.theta, self.B
        self._project(self.enc_in, "enc", th["enc_w_ih"], th["enc_b_ih"], self.enc_gates, 1, ENC_STEPS, 0)
        R.gru_forward_small(k, self.enc_gates, th["enc_b_ih"], th["enc_w_hh"], th["enc_b_hh"], self.h0_zero, 0, None, None,
                            self.enc_hs, self.enc_ghn, None, 1, ENC_STEPS, B, 0)
        k.latent_head_fwd(self.enc_hs[0, ENC_STEPS - 1], th["lat_w"], th["lat_b"], self.eps, self.lat, self.zlat, self.kl, B,
                          self.kl_form, self.ws_lat)

    def _enc_backward_chain(self, beta, stage_eps: bool = False):
        """stage_eps: copy the next iteration's noise into place on the auxiliary stream as soon as the last reader of the current
        noise (the dz / latent backward) is done, instead of on the critical path in front of the encoder forward; returns the
        event to wait for (None: not staged)."""
        k, th, g, B, P, p_ = self.k, self.theta, self.grad, self.B, self.P, self.p
        self._dz_latent_bwd(beta)
        ev_eps = None
        if stage_eps and self.device.type == "cuda" and self.use_side_stream:
            cur = torch.cuda.current_stream()
            if getattr(self, "_aux", None) is None:
                self._aux = torch.cuda.Stream(device=self.device, priority=-1)
            ev = torch.cuda.Event(); ev.record(cur)
            self._aux.wait_event(ev)
            with torch.cuda.stream(self._aux):
                self.eps.copy_(self.eps_next)
                ev_eps = torch.cuda.Event(); ev_eps.record(self._aux)
        k.latent_head_bwd(self.dlat, self.enc_hs[0, ENC_STEPS - 1], th["lat_w"], g["lat_w"], g["lat_b"], self.dhT, B)
        dwhh = R.gru_backward_small(k, self.enc_gates, self.enc_ghn, self.enc_hs, self.h0_zero, 0, th["enc_w_hh"], None, None, self.dhT,
                                    None, g["enc_w_hh"].view(1, G, H), g["enc_b_hh"], g["enc_b_ih"], None, None, self.enc_dh0,
                                    1, ENC_STEPS, B, self.ws_gru_enc, self.ws_dwhh_enc, split_dwhh=self.device.type == "cuda" and self.use_side_stream)
        # the two encoder weight-gradient GEMMs only read the BPTT's outputs: side by side on two streams (this chain is the
        # critical path of a head-sharded iteration)
        ev_aux = None
        if dwhh is not None:
            cur = torch.cuda.current_stream()
            if getattr(self, "_aux", None) is None:
                self._aux = torch.cuda.Stream(device=self.device, priority=-1)
            ev = torch.cuda.Event(); ev.record(cur)
            self._aux.wait_event(ev)
            with torch.cuda.stream(self._aux):
                dwhh()
                ev_aux = torch.cuda.Event(); ev_aux.record(self._aux)
        if self.enc_tc and self.ws_wgrad_tc_enc is not None:
            k.proj_wgrad_tc(self.enc_gates, self.enc_in_hi, self.enc_in_lo, None, g["enc_w_ih"], 1, ENC_STEPS, B, p_, 0, self.ws_wgrad_tc_enc)
        else:
            k.proj_wgrad(self.enc_gates, self.enc_in, None, g["enc_w_ih"], 1, ENC_STEPS, B, p_, 0, self.ws_wgrad)
        if ev_aux is not None:
            torch.cuda.current_stream().wait_event(ev_aux)
        return ev_eps

    def flow_pre(self):
        """pre: encoder chain on the staged noise (eps_next) + every head's projection, for the CURRENT weights."""
        fl = self._flow_setup()
        th, k, p_ = self.theta, self.k, self.p
        self.eps.copy_(self.eps_next)
        side = self._fork()
        with self._on(side):
            self._enc_forward_chain()
        k.split_tf32_gate_rows(th["w_ih"], self.w_ih_hi, self.w_ih_lo, self.P * G, p_)
        k.proj_fwd_tc(self.dec_in_hi, self.dec_in_lo, self.w_ih_hi, self.w_ih_lo, th["b_ih"], self.gates, self.P, DEC_STEPS, self.B, p_, 1)
        self._join(side)

    def flow_rec(self):
        """rec alone (completes a forward whose pre ran earlier): recurrences + MSE + loss."""
        fl = self._flow_setup()
        main = torch.cuda.current_stream()
        start = torch.cuda.Event(); start.record(main)
        evs = []
        for (lo, hi), st in zip(fl["groups"], fl["streams"]):
            if st is None:
                self._flow_rec_group(lo, hi)
            else:
                st.wait_event(start)
                with torch.cuda.stream(st):
                    self._flow_rec_group(lo, hi)
                    e = torch.cuda.Event(); e.record(st); evs.append(e)
        for e in evs:
            main.wait_event(e)
        self.k.dot_small(self.sse, self.P, 1.0 / (DEC_STEPS * self.B), self.loss)

    def flow_body(self, lr: float, lam: float, lam_ridge: float, beta: float, with_pre: bool = True):
        """rec ; update ; (pre)  as one software-pipelined unit (see the section comment).  Expects the state left by
        flow_pre(): gi in the gate buffer, z in zlat.  with_pre=False stops after the update (a batch re-bind or a check
        block follows)."""
        fl = self._flow_setup()
        k = self.k
        main = torch.cuda.current_stream()
        side = self._side if self._side is not None else torch.cuda.Stream(device=self.device, priority=-1)
        self._side = side
        start = torch.cuda.Event(); start.record(main)
        side.wait_event(start)
        ev_mse, ev_bwd, ev_grad, ev_done = [], [], [], []
        for gi_, ((lo, hi), st) in enumerate(zip(fl["groups"], fl["streams"])):
            ctx = torch.cuda.stream(st) if st is not None else self._on(None)
            if st is not None:
                st.wait_event(start)
            with ctx:
                cur = st if st is not None else main
                self._flow_rec_group(lo, hi)
                e = torch.cuda.Event(); e.record(cur); ev_mse.append(e)
                self._flow_bwd_group(lo, hi, fl["ws"][gi_])
                e = torch.cuda.Event(); e.record(cur); ev_bwd.append(e)
        # encoder backward chain needs every group's dh0
        with torch.cuda.stream(side):
            for e in ev_mse:
                side.wait_event(e)
            k.dot_small(self.sse, self.P, 1.0 / (DEC_STEPS * self.B), self.loss)       # loss of the forward just completed
            for e in ev_bwd:
                side.wait_event(e)
            ev_eps = self._enc_backward_chain(beta, stage_eps=with_pre)
        for gi_, ((lo, hi), st) in enumerate(zip(fl["groups"], fl["streams"])):
            ctx = torch.cuda.stream(st) if st is not None else self._on(None)
            with ctx:
                cur = st if st is not None else main
                self._flow_wgrads_group(lo, hi, fl["ws"][gi_], lam_ridge)
                e = torch.cuda.Event(); e.record(cur); ev_grad.append(e)
                self._flow_prox_proj_group(lo, hi, lr, lam, with_pre)
                e = torch.cuda.Event(); e.record(cur); ev_done.append(e)
        with torch.cuda.stream(side):
            for e in ev_grad:                      # w_hh gradients of every group; also: every reader of z (dW_hh) is done
                side.wait_event(e)
            self._flow_gd_rest(lr)
            if with_pre:
                if ev_eps is None:
                    self.eps.copy_(self.eps_next)
                else:
                    side.wait_event(ev_eps)
                self._enc_forward_chain()
            e_side = torch.cuda.Event(); e_side.record(side)
        for e in ev_done:
            main.wait_event(e)
        main.wait_event(e_side)

    # ------------------------------------------------------------------ update
    def step(self, lr: float, lam: float):
        """GD on every parameter (:498-499) + group-lasso prox on the heads' w_ih (:502-504)."""
        k = self.k
        if self.P > 0:
            k.gd_prox_gc(self.theta["w_ih"], self.grad["w_ih"], self.mask_u8, self.col_norm, self.P, self.Kw,
                         _f32(lr), _f32(lam * lr), lam > 0)
        n_rest = self.theta.numel - self.rest_off
        k.gd_step(self.theta.flat[self.rest_off:], self.grad.flat[self.rest_off:], n_rest, _f32(lr))

    def prox_only(self, lam: float, lr: float):
        if self.P > 0:
            self.k.gd_prox_gc(self.theta["w_ih"], None, self.mask_u8, self.col_norm, self.P, self.Kw,
                              0.0, _f32(lam * lr), True)

    def column_norms(self) -> torch.Tensor:
        """||w_ih[i][:, j]||_2 for the current weights -- what GC() stacks (:297-299)."""
        if self.P > 0:
            self.k.gd_prox_gc(self.theta["w_ih"], None, self.mask_u8, self.col_norm, self.P, self.Kw, 0.0, 0.0, False)
        if self.packed:     # packed column c of head i is series cols[i][c]
            self.col_norm_dense.zero_()
            self.col_norm_dense.scatter_(1, self.cols.long(), self.col_norm * self.mask_u8.float())
        return self.col_norm_dense

    # ------------------------------------------------------------------ packed <-> dense views of w_ih-shaped tensors
    def unpack_w(self, w: torch.Tensor) -> torch.Tensor:
        """[P,G,Kw] storage -> masked-dense [P,G,p] (identity for a dense / masked-dense engine)."""
        if not self.packed:
            return w
        out = torch.zeros(self.P, G, self.p, dtype=w.dtype, device=w.device)
        idx = self.cols.long().to(w.device)[:, None, :].expand(-1, G, -1)
        out.scatter_(2, idx, w * self.mask_u8.to(w.device).to(w.dtype)[:, None, :])
        return out

    def pack_w(self, dense: torch.Tensor) -> torch.Tensor:
        """masked-dense [P,G,p] -> [P,G,Kw] storage."""
        if not self.packed:
            return dense
        idx = self.cols.long().to(dense.device)[:, None, :].expand(-1, G, -1)
        return dense.gather(2, idx) * self.mask_u8.to(dense.device).to(dense.dtype)[:, None, :]

    def ridge_value(self, lam_ridge: float) -> torch.Tensor:
        """sum_i ridge_regularize(net_i, lam) (:321-325, :488) for this shard, as a device scalar."""
        if lam_ridge == 0.0 or self.P == 0:
            return torch.zeros((), dtype=torch.float32, device=self.device)
        self.k.sumsq(self.theta["w_lin"], self.P * H, self.scratch[0:1])
        self.k.sumsq(self.theta["w_hh"], self.P * G * H, self.scratch[1:2])
        return lam_ridge * (self.scratch[0] + self.scratch[1])

    # ------------------------------------------------------------------ snapshots (deepcopy / restore, :547/:558)
    def snapshot(self) -> torch.Tensor:
        return self.theta.flat.clone()

    def restore(self, snap: torch.Tensor):
        self.theta.flat.copy_(snap)

    def zero_grad(self):
        self.grad.flat.zero_()

    def close(self):
        if self.comm is not None:
            self.comm.close()
