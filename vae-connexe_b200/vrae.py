"""Mirror of the reference's generic VRAE (VRAE.py:105-169, GRU variant) on the fused kernels (config 4):
Encoder GRU(D->H) -> fc_mu / fc_logvar (H->Z) -> z = mu + randn_like(std)*std -> h0 = tanh(fc_z2h z) ->
decoder GRUCell loop with teacher forcing -> output activation(fc_out h).

Teacher forcing 1.0 (the reference's default, VRAE.py:134/:157) makes every decoder input known in advance
(step t reads target[:, t], :82/:97), so the whole decoder is ONE projection GEMM + ONE persistent recurrent
kernel instead of T GRUCell calls with a host sync each (`torch.rand(1).item()`, :95-96).  The per-step
`torch.rand(1)` draws are still consumed so the CPU generator stays in lock step with the reference.
With teacher_forcing_ratio < 1 (VRAE.py:95-100, the schedules of :173-182 used by the reference's own example :196-197) some
steps feed the model's OWN previous output back: the decoder then runs step by step (projection, one recurrent step, output
layer per step; the per-step torch.rand(1) decides, as in the reference) and the backward carries the gradient through the
fed-back inputs (d x_in = d gates . W_ih) -- forward_free / backward_free below.  This path is launch-bound like the
reference's GRUCell loop; the fused path stays the default (ratio 1.0).
LSTM / RNN cell types are out of scope.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn

from . import lib as L
from . import rec as R
from .engine import Arena, G, H

_ACT = {"tanh": 0, "sigmoid": 1, "relu": 2}


class _Engine:
    def __init__(self, D: int, Z: int, act: str, device):
        self.k = L.kernels()
        self.D, self.Z, self.act = D, Z, _ACT.get(act, 3)
        self.device = torch.device(device)
        shapes = {"enc_w_ih": (G, D), "enc_w_hh": (G, H), "enc_b_ih": (G,), "enc_b_hh": (G,),
                  "lat_w": (2 * Z, H), "lat_b": (2 * Z,), "z2h_w": (H, Z), "z2h_b": (H,),
                  "dec_w_ih": (G, D), "dec_w_hh": (G, H), "dec_b_ih": (G,), "dec_b_hh": (G,),
                  "out_w": (D, H), "out_b": (D,), "start_token": (1, D)}
        self.theta = Arena(shapes, self.device)
        self.grad = self.theta.like()
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.theta.flat), torch.zeros_like(self.theta.flat)
        self.adam_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.adam_counter_st = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.n_trainable = self.theta.offsets["start_token"]      # start_token never gets a gradient under TF = 1
        self.shape = None

    @staticmethod
    def _row_splits(rows: int) -> int:
        """Chunks a long row reduction is cut into: the largest power of two <= 512 that divides `rows` and leaves >= 512 rows per chunk."""
        s = 1
        while s < 512 and rows % (2 * s) == 0 and rows // (2 * s) >= 512:
            s *= 2
        return s

    def bind(self, x: torch.Tensor):
        B, T, D = x.shape
        if self.shape != (B, T):
            dev = self.device
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
            self.shape, self.B, self.T = (B, T), B, T
            Z = self.Z
            self.xin = z(T, B, D)
            self.enc_gates, self.dec_gates = z(1, T, B, G), z(1, T, B, G)
            self.enc_hs, self.enc_ghn, self.dec_hs, self.dec_ghn = z(1, T, B, H), z(1, T, B, H), z(1, T, B, H), z(1, T, B, H)
            self.dhs, self.h0_zero = z(1, T, B, H), z(B, H)
            self.lat, self.dlat, self.zlat, self.eps, self.dz = z(B, 2 * Z), z(B, 2 * Z), z(B, Z), z(B, Z), z(1, B, Z)
            self.pre0, self.h0, self.dh0, self.dpre0 = z(B, H), z(B, H), z(1, B, H), z(B, H)
            self.pre, self.recon, self.drecon, self.dpre = z(T, B, D), z(T, B, D), z(T, B, D), z(T, B, D)
            self.dhT, self.enc_dh0 = z(1, B, H), z(1, B, H)
            self.sse, self.kl, self.sse_t = z(1), z(1), z(T)
            self.ones_B, self.ones_TB = torch.ones(B, 1, device=dev), torch.ones(T * B, 1, device=dev)
            S = self._row_splits(T * B)
            self.ones_S, self.part_w, self.part_b = torch.ones(1, S, device=dev), z(S, D, H), z(S, D)
            k = self.k
            self.ws_gru = torch.zeros(k.gru_bwd_workspace(1, B) // 4 + 4, dtype=torch.float32, device=dev)
            n = R.dwhh_workspace(k, 1, T, B)
            self.ws_dwhh = torch.zeros(n, dtype=torch.float32, device=dev) if n else None
            self.ws_wgrad = torch.zeros(k.proj_wgrad_workspace(1, T, B, D) // 4 + 4, dtype=torch.float32, device=dev)
            # free-running (teacher forcing < 1) path: actual decoder inputs, per-step scratch
            self.dec_xin, self.dxin, self.dh_carry, self.dh_step = z(T, B, D), z(B, D), z(1, B, H), z(1, B, H)
            self.t_dw_hh, self.t_db_hh, self.t_db_ih = z(1, G, H), z(G), z(G)
            self.ws_gru1 = torch.zeros(k.gru_bwd_workspace(1, B) // 4 + 4, dtype=torch.float32, device=dev)
        self.xin.copy_(x.transpose(0, 1))
        self.free = None

    def forward(self, eps: torch.Tensor):
        """Teacher-forced forward (VRAE.forward :133-137 with teacher_forcing_ratio = 1)."""
        k, th, B, T, D, Z = self.k, self.theta, self.B, self.T, self.D, self.Z
        self.eps.copy_(eps.reshape(B, Z), non_blocking=True)
        k.proj_fwd(self.xin, th["enc_w_ih"], th["enc_b_ih"], self.enc_gates, 1, T, B, D, 0)
        R.gru_forward_small(k, self.enc_gates, th["enc_b_ih"], th["enc_w_hh"], th["enc_b_hh"], self.h0_zero, 0, None, None,
                            self.enc_hs, self.enc_ghn, None, 1, T, B, 0)                                        # :30
        hT = self.enc_hs[0, T - 1]
        k.gemm(L.GEMM_NT, 1, B, 2 * Z, H, hT, H, 0, th["lat_w"], H, 0, self.lat, 2 * Z, 0, th["lat_b"], 0)        # :34-35
        k.latent_fwd(self.lat, self.eps, self.zlat, self.kl, B, L.KL_STANDARD, Z)                                # :117-121, :144
        k.gemm(L.GEMM_NT, 1, B, H, Z, self.zlat, Z, 0, th["z2h_w"], Z, 0, self.pre0, H, 0, th["z2h_b"], 0)
        k.tanh_fwd(self.pre0, self.h0, B * H)                                                                    # :72
        k.proj_fwd(self.xin, th["dec_w_ih"], th["dec_b_ih"], self.dec_gates, 1, T, B, D, 0)                      # input of step t = x[:, t]
        R.gru_forward_small(k, self.dec_gates, th["dec_b_ih"], th["dec_w_hh"], th["dec_b_hh"], self.h0, 0, None, None,
                            self.dec_hs, self.dec_ghn, None, 1, T, B, 0)                                        # :85-89
        k.gemm(L.GEMM_NT, 1, T * B, D, H, self.dec_hs, H, 0, th["out_w"], H, 0, self.pre, D, 0, th["out_b"], 0)
        k.act_fwd(self.pre, self.recon, T * B * D, self.act)                                                     # :91
        # rec_loss = SSE / batch (:143): sse over everything, d(recon) = 2*(recon - x)/B
        k.mse_fwd_bwd(self.recon, self.xin, self.sse_t, self.drecon, None, T, 1, B * D, 2.0 / B)      # one CTA per timestep
        k.dot_small(self.sse_t, T, 1.0, self.sse)

    def backward(self, beta: float):
        k, th, g, B, T, D, Z = self.k, self.theta, self.grad, self.B, self.T, self.D, self.Z
        TB = T * B
        k.act_bwd(self.drecon, self.recon, self.dpre, TB * D, self.act)
        # fc_out gradients: reductions over T*B rows (524,288 at the config-4 size).  One GEMM would walk them in ONE CTA
        # (16+ ms each); cut into S row chunks as a batched GEMM (S CTAs) + a [1 x S] . [S x D*H] GEMM that adds the partials
        S = self._row_splits(TB)
        if S > 1:
            rows = TB // S
            k.gemm(L.GEMM_TN, S, D, H, rows, self.dpre, D, rows * D, self.dec_hs, H, rows * H, self.part_w, H, D * H)
            k.gemm(L.GEMM_NN, 1, 1, D * H, S, self.ones_S, S, 0, self.part_w, D * H, 0, g["out_w"], D * H, 0)
            k.gemm(L.GEMM_TN, S, 1, D, rows, self.ones_TB, 1, rows, self.dpre, D, rows * D, self.part_b, D, D)
            k.gemm(L.GEMM_NN, 1, 1, D, S, self.ones_S, S, 0, self.part_b, D, 0, g["out_b"], D, 0)
        else:
            k.gemm(L.GEMM_TN, 1, D, H, TB, self.dpre, D, 0, self.dec_hs, H, 0, g["out_w"], H, 0)
            k.gemm(L.GEMM_TN, 1, 1, D, TB, self.ones_TB, 1, 0, self.dpre, D, 0, g["out_b"], D, 0)
        k.gemm(L.GEMM_NN, 1, TB, H, D, self.dpre, D, 0, th["out_w"], H, 0, self.dhs, H, 0)
        R.gru_backward_small(k, self.dec_gates, self.dec_ghn, self.dec_hs, self.h0, 0, th["dec_w_hh"], None, None, None, self.dhs,
                             g["dec_w_hh"].view(1, G, H), g["dec_b_hh"], g["dec_b_ih"], None, None, self.dh0, 1, T, B, self.ws_gru,
                             self.ws_dwhh)
        k.proj_wgrad(self.dec_gates, self.xin, None, g["dec_w_ih"], 1, T, B, D, 0, self.ws_wgrad)
        self._backward_tail(beta)

    def _backward_tail(self, beta: float):
        """From dL/dh0 of the decoder back through tanh(fc_z2h z), the reparameterisation + KL, fc_mu / fc_logvar and the encoder."""
        k, th, g, B, T, D, Z = self.k, self.theta, self.grad, self.B, self.T, self.D, self.Z
        k.tanh_bwd(self.dh0, self.h0, self.dpre0, B * H)
        k.gemm(L.GEMM_TN, 1, H, Z, B, self.dpre0, H, 0, self.zlat, Z, 0, g["z2h_w"], Z, 0)
        k.gemm(L.GEMM_TN, 1, 1, H, B, self.ones_B, 1, 0, self.dpre0, H, 0, g["z2h_b"], H, 0)
        k.gemm(L.GEMM_NN, 1, B, Z, H, self.dpre0, H, 0, th["z2h_w"], Z, 0, self.dz, Z, 0)
        k.latent_bwd(self.dz, 1, None, self.lat, self.eps, beta, L.KL_STANDARD, self.dlat, None, B, Z)
        hT = self.enc_hs[0, T - 1]
        k.gemm(L.GEMM_TN, 1, 2 * Z, H, B, self.dlat, 2 * Z, 0, hT, H, 0, g["lat_w"], H, 0)
        k.gemm(L.GEMM_TN, 1, 1, 2 * Z, B, self.ones_B, 1, 0, self.dlat, 2 * Z, 0, g["lat_b"], 2 * Z, 0)
        k.gemm(L.GEMM_NN, 1, B, H, 2 * Z, self.dlat, 2 * Z, 0, th["lat_w"], H, 0, self.dhT, H, 0)
        R.gru_backward_small(k, self.enc_gates, self.enc_ghn, self.enc_hs, self.h0_zero, 0, th["enc_w_hh"], None, None, self.dhT, None,
                             g["enc_w_hh"].view(1, G, H), g["enc_b_hh"], g["enc_b_ih"], None, None, self.enc_dh0, 1, T, B, self.ws_gru,
                             self.ws_dwhh)
        k.proj_wgrad(self.enc_gates, self.xin, None, g["enc_w_ih"], 1, T, B, D, 0, self.ws_wgrad)

    # ------------------------------------------------------------------ teacher forcing < 1 (VRAE.py:68-102)
    def forward_free(self, eps: torch.Tensor, use_tf, first_from_target: bool):
        """Decoder step by step: use_tf[t] (t < T-1) = the reference's `torch.rand(1).item() < teacher_forcing_ratio` of step t
        (:95-96): True -> x_in[t+1] = target[:, t+1], False -> x_in[t+1] = the model's own output of step t (:97-100)."""
        k, th, B, T, D, Z = self.k, self.theta, self.B, self.T, self.D, self.Z
        self.eps.copy_(eps.reshape(B, Z), non_blocking=True)
        k.proj_fwd(self.xin, th["enc_w_ih"], th["enc_b_ih"], self.enc_gates, 1, T, B, D, 0)
        R.gru_forward_small(k, self.enc_gates, th["enc_b_ih"], th["enc_w_hh"], th["enc_b_hh"], self.h0_zero, 0, None, None,
                            self.enc_hs, self.enc_ghn, None, 1, T, B, 0)
        hT = self.enc_hs[0, T - 1]
        k.gemm(L.GEMM_NT, 1, B, 2 * Z, H, hT, H, 0, th["lat_w"], H, 0, self.lat, 2 * Z, 0, th["lat_b"], 0)
        k.latent_fwd(self.lat, self.eps, self.zlat, self.kl, B, L.KL_STANDARD, Z)
        k.gemm(L.GEMM_NT, 1, B, H, Z, self.zlat, Z, 0, th["z2h_w"], Z, 0, self.pre0, H, 0, th["z2h_b"], 0)
        k.tanh_fwd(self.pre0, self.h0, B * H)
        if first_from_target:
            self.dec_xin[0].copy_(self.xin[0])                                             # :80
        else:
            self.dec_xin[0].copy_(th["start_token"].expand(B, D))                          # :82
        for t in range(T):
            h_prev = self.h0 if t == 0 else self.dec_hs[0, t - 1]
            k.proj_fwd(self.dec_xin[t:t + 1], th["dec_w_ih"].view(1, G, D), th["dec_b_ih"], self.dec_gates[:, t:t + 1], 1, 1, B, D, 0)
            R.gru_forward_small(k, self.dec_gates[:, t:t + 1], th["dec_b_ih"], th["dec_w_hh"], th["dec_b_hh"], h_prev, 0, None, None,
                                self.dec_hs[:, t:t + 1], self.dec_ghn[:, t:t + 1], None, 1, 1, B, 0)   # the GRUCell step, :88
            k.gemm(L.GEMM_NT, 1, B, D, H, self.dec_hs[0, t], H, 0, th["out_w"], H, 0, self.pre[t], D, 0, th["out_b"], 0)
            k.act_fwd(self.pre[t], self.recon[t], B * D, self.act)                                     # :91
            if t < T - 1:
                self.dec_xin[t + 1].copy_(self.xin[t + 1] if use_tf[t] else self.recon[t])
        k.mse_fwd_bwd(self.recon, self.xin, self.sse_t, self.drecon, None, T, 1, B * D, 2.0 / B)
        k.dot_small(self.sse_t, T, 1.0, self.sse)
        self.free = dict(use_tf=list(use_tf), first_from_target=first_from_target)

    def backward_free(self, beta: float):
        """Backward of forward_free: BPTT one step at a time, the gradient of a fed-back input (d x_in[t+1] = dgates[t+1] . W_ih)
        joins d recon[t] wherever step t+1 consumed the model's own output."""
        k, th, g, B, T, D, Z = self.k, self.theta, self.grad, self.B, self.T, self.D, self.Z
        use_tf = self.free["use_tf"]
        g["dec_w_hh"].zero_(); g["dec_b_hh"].zero_(); g["dec_b_ih"].zero_(); g["start_token"].zero_()
        self.dh_carry.zero_()
        fed_back = False                                   # does step t+1 read recon[t]?
        for t in range(T - 1, -1, -1):
            if fed_back:
                k.axpy(self.drecon[t], self.dxin, B * D, 1.0)
            k.act_bwd(self.drecon[t], self.recon[t], self.dpre[t], B * D, self.act)
            # dh_t = dpre_t . out_w + (carry from step t+1)
            k.gemm(L.GEMM_NN, 1, B, H, D, self.dpre[t], D, 0, th["out_w"], H, 0, self.dh_step[0], H, 0)
            k.axpy(self.dh_step, self.dh_carry, B * H, 1.0)
            h_prev = self.h0 if t == 0 else self.dec_hs[0, t - 1]
            k.gru_bwd(self.dec_gates[:, t:t + 1], self.dec_ghn[:, t:t + 1], self.dec_hs[:, t:t + 1], h_prev, 0, th["dec_w_hh"], None, None,
                      self.dh_step, None, self.t_dw_hh, self.t_db_hh, self.t_db_ih, None, None, self.dh_carry, 1, 1, B, self.ws_gru1)
            k.axpy(g["dec_w_hh"], self.t_dw_hh, G * H, 1.0); k.axpy(g["dec_b_hh"], self.t_db_hh, G, 1.0); k.axpy(g["dec_b_ih"], self.t_db_ih, G, 1.0)
            fed_back = t > 0 and not use_tf[t - 1]
            if fed_back or (t == 0 and not self.free["first_from_target"]):
                k.gemm(L.GEMM_NN, 1, B, D, G, self.dec_gates[0, t], G, 0, th["dec_w_ih"], D, 0, self.dxin, D, 0)   # d x_in[t]
            if t == 0 and not self.free["first_from_target"]:
                k.gemm(L.GEMM_TN, 1, 1, D, B, self.ones_B, 1, 0, self.dxin, D, 0, g["start_token"], D, 0)
        TB = T * B
        k.gemm(L.GEMM_TN, 1, D, H, TB, self.dpre, D, 0, self.dec_hs, H, 0, g["out_w"], H, 0)
        k.gemm(L.GEMM_TN, 1, 1, D, TB, self.ones_TB, 1, 0, self.dpre, D, 0, g["out_b"], D, 0)
        k.proj_wgrad(self.dec_gates, self.dec_xin, None, g["dec_w_ih"], 1, T, B, D, 0, self.ws_wgrad)
        self.dh0.copy_(self.dh_carry)
        self._backward_tail(beta)

    def adam_step(self, lr: float):
        # start_token has a gradient only when the decoder started from it (teacher_forcing_ratio <= 0); torch.optim.Adam skips
        # parameters without one
        # (its own step counter: torch keeps the step count per parameter, and this one starts when its first gradient arrives)
        n = self.n_trainable
        self.k.adam_step_dev(self.theta.flat, self.grad.flat, self.exp_avg, self.exp_avg_sq, n, lr, 0.9, 0.999, 1e-8, self.adam_counter)
        if self.free is not None and not self.free["first_from_target"]:
            m = self.theta.numel - n
            self.k.adam_step_dev(self.theta.flat[n:], self.grad.flat[n:], self.exp_avg[n:], self.exp_avg_sq[n:], m, lr, 0.9, 0.999, 1e-8,
                                 self.adam_counter_st)


class VRAE(nn.Module):
    """VRAE(input_dim, hidden_dim=64, latent_dim=2, rnn_type='gru', output_activation='sigmoid') (VRAE.py:105-112)."""

    def __init__(self, input_dim: int, hidden_dim: int = 64, latent_dim: int = 2, rnn_type: str = "gru",
                 output_activation: str = "sigmoid") -> None:
        super().__init__()
        if hidden_dim != H:
            raise ValueError(f"kernels are built for hidden_dim={H}")
        if rnn_type.lower() != "gru":
            raise NotImplementedError("only the GRU cell type is on the accelerated path (SURVEY.md 8(a16))")
        kern = L.kernels()
        dev = torch.device("cuda", torch.cuda.current_device()) if kern.device_type == "cuda" else torch.device("cpu")
        self.input_dim, self.latent_dim, self.output_activation = input_dim, latent_dim, output_activation
        self.engine = e = _Engine(input_dim, latent_dim, output_activation, dev)
        # declaration order of the reference (:23-25, :44-56): encoder.rnn, fc_mu, fc_logvar, fc_z2h, cell, fc_out, start_token
        rnn = nn.GRU(input_dim, H, batch_first=True)
        fc_mu, fc_lv, z2h = nn.Linear(H, latent_dim), nn.Linear(H, latent_dim), nn.Linear(latent_dim, H)
        cell, fc_out = nn.GRUCell(input_dim, H), nn.Linear(H, input_dim)
        start = torch.randn(1, input_dim)
        th = e.theta
        with torch.no_grad():
            th["enc_w_ih"].copy_(rnn.weight_ih_l0); th["enc_w_hh"].copy_(rnn.weight_hh_l0)
            th["enc_b_ih"].copy_(rnn.bias_ih_l0); th["enc_b_hh"].copy_(rnn.bias_hh_l0)
            th["lat_w"][:latent_dim].copy_(fc_mu.weight); th["lat_w"][latent_dim:].copy_(fc_lv.weight)
            th["lat_b"][:latent_dim].copy_(fc_mu.bias); th["lat_b"][latent_dim:].copy_(fc_lv.bias)
            th["z2h_w"].copy_(z2h.weight); th["z2h_b"].copy_(z2h.bias)
            th["dec_w_ih"].copy_(cell.weight_ih); th["dec_w_hh"].copy_(cell.weight_hh)
            th["dec_b_ih"].copy_(cell.bias_ih); th["dec_b_hh"].copy_(cell.bias_hh)
            th["out_w"].copy_(fc_out.weight); th["out_b"].copy_(fc_out.bias); th["start_token"].copy_(start)

    def state_dict(self, *a, **k):
        th, Z = self.engine.theta, self.latent_dim
        sd = {"encoder.rnn.weight_ih_l0": th["enc_w_ih"], "encoder.rnn.weight_hh_l0": th["enc_w_hh"],
              "encoder.rnn.bias_ih_l0": th["enc_b_ih"], "encoder.rnn.bias_hh_l0": th["enc_b_hh"],
              "encoder.fc_mu.weight": th["lat_w"][:Z], "encoder.fc_mu.bias": th["lat_b"][:Z],
              "encoder.fc_logvar.weight": th["lat_w"][Z:], "encoder.fc_logvar.bias": th["lat_b"][Z:],
              "decoder.start_token": th["start_token"],
              "decoder.fc_z2h.weight": th["z2h_w"], "decoder.fc_z2h.bias": th["z2h_b"],
              "decoder.cell.weight_ih": th["dec_w_ih"], "decoder.cell.weight_hh": th["dec_w_hh"],
              "decoder.cell.bias_ih": th["dec_b_ih"], "decoder.cell.bias_hh": th["dec_b_hh"],
              "decoder.fc_out.weight": th["out_w"], "decoder.fc_out.bias": th["out_b"]}
        return {n: t.detach().clone() for n, t in sd.items()}

    def _consume_tf_draws(self, T: int):
        for _ in range(T - 1):          # one torch.rand(1) per non-final step (:95-96), result irrelevant at ratio 1.0
            torch.rand(1)

    def forward(self, x: torch.Tensor, teacher_forcing_ratio: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        e = self.engine
        e.bind(x)
        eps = torch.randn(x.shape[0], self.latent_dim)        # randn_like(std) on the CPU generator (:119)
        if teacher_forcing_ratio >= 1.0:                      # every draw says "teacher forcing": the fused path
            self._consume_tf_draws(x.shape[1])
            e.forward(eps.to(e.device))
        else:                                                 # one torch.rand(1) per non-final step decides (:95-96)
            use_tf = [torch.rand(1).item() < teacher_forcing_ratio for _ in range(x.shape[1] - 1)]
            e.forward_free(eps.to(e.device), use_tf, teacher_forcing_ratio > 0)
        Z = self.latent_dim
        return e.recon.permute(1, 0, 2), e.lat[:, :Z], e.lat[:, Z:]

    @staticmethod
    def loss(recon, x, mu, logvar, beta: float = 1.0):
        """VRAE.loss (:142-147) evaluated with torch ops on the returned tensors (reporting only; the trainer uses
        the fused loss kernels)."""
        rec = torch.nn.functional.mse_loss(recon, x, reduction="sum") / x.size(0)
        kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) / x.size(0)
        return rec + beta * kld, rec, kld

    def generate(self, z: torch.Tensor, seq_len: int) -> torch.Tensor:
        """Free-running decoding from latent codes (:139-141): x_in starts at start_token, then the model's own output."""
        e, k, th = self.engine, self.engine.k, self.engine.theta
        B, D, Z = z.shape[0], self.input_dim, self.latent_dim
        dev = e.device
        zt = z.to(dev, torch.float32).contiguous()
        pre0, h = torch.empty(B, H, device=dev), torch.empty(B, H, device=dev)
        k.gemm(L.GEMM_NT, 1, B, H, Z, zt, Z, 0, th["z2h_w"], Z, 0, pre0, H, 0, th["z2h_b"], 0)
        k.tanh_fwd(pre0, h, B * H)
        x_in = th["start_token"].expand(B, D).contiguous().view(1, B, D)
        gates, ghn = torch.empty(1, 1, B, G, device=dev), torch.empty(1, 1, B, H, device=dev)
        h_next, pre, y = torch.empty(B, H, device=dev), torch.empty(B, D, device=dev), torch.empty(B, D, device=dev)
        outs = torch.empty(B, seq_len, D, device=dev)
        for t in range(seq_len):
            k.proj_fwd(x_in, th["dec_w_ih"].view(1, G, D), th["dec_b_ih"], gates, 1, 1, B, D, 0)
            k.gru_fwd(gates, th["dec_b_ih"], th["dec_w_hh"], th["dec_b_hh"], h, 0, None, None, h_next.view(1, 1, B, H), ghn,
                      None, 1, 1, B, 0)
            k.gemm(L.GEMM_NT, 1, B, D, H, h_next, H, 0, th["out_w"], H, 0, pre, D, 0, th["out_b"], 0)
            k.act_fwd(pre, y, B * D, e.act)
            outs[:, t] = y
            x_in = y.view(1, B, D).clone()
            h, h_next = h_next, h
        return outs

    def sample(self, batch_size: int, seq_len: int, device: str = "cpu") -> torch.Tensor:
        z = torch.randn(batch_size, self.latent_dim)          # (:145)
        return self.generate(z, seq_len)


def train(model: VRAE, data: torch.Tensor, epochs: int = 10, lr: float = 1e-3, beta: float = 1.0,
          teacher_forcing_schedule: Optional[Callable] = None, log: Optional[list] = None) -> None:
    """train() of VRAE.py (:150-169): full-batch Adam, optional teacher-forcing schedule (:173-182)."""
    e = model.engine
    for epoch in range(epochs):
        tf_ratio = teacher_forcing_schedule(epoch) if teacher_forcing_schedule else 1.0
        recon, mu, logvar = model(data, teacher_forcing_ratio=tf_ratio)
        if e.free is not None:
            e.backward_free(beta)
        else:
            e.backward(beta)
        e.adam_step(lr)
        if epoch % 10 == 0:
            rec = float(e.sse) / data.shape[0]
            kld = float(e.kl)
            total = rec + beta * kld
            print(f"Epoch {epoch:3d}/{epochs}  |  Total: {total:.4f}  |  Rec: {rec:.4f}  |  KLD: {kld:.4f}  |  TF: {tf_ratio:.2f}")
            if log is not None:
                log.append(dict(epoch=epoch, total=total, rec=rec, kld=kld))


def exponential_teacher_forcing_schedule(epoch: int, initial_ratio: float = 1.0, decay_rate: float = 0.05) -> float:
    return initial_ratio * (1 - decay_rate) ** epoch


def linear_teacher_forcing_schedule(epoch: int, initial_ratio: float = 1.0, final_ratio: float = 0.0, total_epochs: int = 100) -> float:
    return initial_ratio - (initial_ratio - final_ratio) * (epoch / total_epochs)
