"""Kernel orchestration for VRAE4E, the error-compensation VRAE of phase 2
(CRVAE_lorenz96.py:123-179): encoder GRU -> [mu|log_var] -> reparameterise -> tanh(Linear) ->
decoder GRU (h0 = z) -> Linear(H,p), the MSE + KL loss of the trainer (:599-603 / :639-643), the
hand-written backward and torch.optim.Adam-equivalent update (:565, :612-614).

Same kernels as the CRVAE engine with P = 1 (one encoder "head", one decoder "head"); the decoder's
Linear(H,p) head and its gradients are plain GEMMs around the recurrent kernel (gradient w.r.t. every
h_t enters crvae_gru_bwd through its `dhs` input).  The model is replicated on every rank.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import lib as L
from . import rec as R
from .engine import Arena, G, H

STEPS = 10    # VRAE4E sees the 10-step residual (:152-155, :166)


class VRAE4EEngine:
    def __init__(self, p: int, device="cuda"):
        self.k = L.kernels()
        self.p = int(p)
        self.device = torch.device(device)
        p_ = self.p
        shapes = {
            "enc_w_ih": (G, p_), "enc_w_hh": (G, H), "enc_b_ih": (G,), "enc_b_hh": (G,),
            "lat_w": (2 * H, H), "lat_b": (2 * H,), "hid_w": (H, H), "hid_b": (H,),
            "dec_w_ih": (G, p_), "dec_w_hh": (G, H), "dec_b_ih": (G,), "dec_b_hh": (G,),
            "out_w": (p_, H), "out_b": (p_,),
        }
        self.theta = Arena(shapes, self.device)
        self.grad = self.theta.like()
        self.exp_avg = torch.zeros_like(self.theta.flat)
        self.exp_avg_sq = torch.zeros_like(self.theta.flat)
        self.adam_counter = torch.zeros(1, dtype=torch.int32, device=self.device)   # optimizer step count (device side)
        self.B = None
        self.kl_form = L.KL_SWAPPED

    def _alloc(self, B: int):
        dev, p_ = self.device, self.p
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.B = B
        self.enc_in = z(STEPS, B, p_)
        self.dec_in = z(STEPS, B, p_)
        self.enc_gates, self.dec_gates = z(1, STEPS, B, G), z(1, STEPS, B, G)
        self.enc_hs, self.enc_ghn = z(1, STEPS, B, H), z(1, STEPS, B, H)
        self.dec_hs, self.dec_ghn = z(1, STEPS, B, H), z(1, STEPS, B, H)
        self.dhs = z(1, STEPS, B, H)
        self.h0_zero = z(B, H)
        self.lat, self.dlat = z(B, 2 * H), z(B, 2 * H)
        self.zlat, self.eps, self.eps_next = z(B, H), z(B, H), z(B, H)
        self.pre, self.zh, self.dzh, self.dpre, self.dz = z(B, H), z(B, H), z(1, B, H), z(B, H), z(1, B, H)
        self.dhT, self.enc_dh0 = z(1, B, H), z(1, B, H)
        self.pred, self.dpred = z(STEPS, B, p_), z(STEPS, B, p_)
        self.sse, self.kl, self.loss = z(1), z(1), z(1)
        self.ones_B, self.ones_TB = torch.ones(B, 1, device=dev), torch.ones(STEPS * B, 1, device=dev)
        k = self.k
        self.ws_lat = torch.zeros(k.latent_head_workspace(B) // 4 + 4, dtype=torch.float32, device=dev) if hasattr(k, "latent_head_fwd") else None
        self.ws_gru = torch.zeros(k.gru_bwd_workspace(1, B) // 4 + 4, dtype=torch.float32, device=dev)
        n = R.dwhh_workspace(k, 1, STEPS, B)
        self.ws_dwhh = torch.zeros(n, dtype=torch.float32, device=dev) if n else None
        self.ws_wgrad = torch.zeros(k.proj_wgrad_workspace(1, STEPS, B, p_) // 4 + 4, dtype=torch.float32, device=dev)

    def bind_error(self, err_tbp: torch.Tensor):
        """err_tbp [10, B, p] (time-major residual).  Encoder input = the residual itself (:155);
        decoder input = [0, e_0 .. e_8] (:166); the target of the MSE is the residual (:601)."""
        B = err_tbp.shape[1]
        if self.B != B:
            self._alloc(B)
        self.enc_in.copy_(err_tbp)
        self.dec_in[1:].copy_(err_tbp[:-1])

    def forward(self, eps: Optional[torch.Tensor] = None):
        k, th, B, p_ = self.k, self.theta, self.B, self.p
        if eps is not None:
            self.eps.copy_(eps.reshape(B, H), non_blocking=True)
        k.proj_fwd(self.enc_in, th["enc_w_ih"], th["enc_b_ih"], self.enc_gates, 1, STEPS, B, p_, 0)
        R.gru_forward_small(k, self.enc_gates, th["enc_b_ih"], th["enc_w_hh"], th["enc_b_hh"], self.h0_zero, 0, None, None,
                            self.enc_hs, self.enc_ghn, None, 1, STEPS, B, 0)
        hT = self.enc_hs[0, STEPS - 1]
        if self.ws_lat is not None:      # fused fc_mu|fc_std + reparameterisation + KL (:157-163), one launch
            k.latent_head_fwd(hT, th["lat_w"], th["lat_b"], self.eps, self.lat, self.zlat, self.kl, B, self.kl_form, self.ws_lat)
        else:
            k.gemm(L.GEMM_NT, 1, B, 2 * H, H, hT, H, 0, th["lat_w"], H, 0, self.lat, 2 * H, 0, th["lat_b"], 0)      # :157-158
            k.latent_fwd(self.lat, self.eps, self.zlat, self.kl, B, self.kl_form)                                   # :160-163
        k.gemm(L.GEMM_NT, 1, B, H, H, self.zlat, H, 0, th["hid_w"], H, 0, self.pre, H, 0, th["hid_b"], 0)       # :164
        k.tanh_fwd(self.pre, self.zh, B * H)
        k.proj_fwd(self.dec_in, th["dec_w_ih"], th["dec_b_ih"], self.dec_gates, 1, STEPS, B, p_, 1)
        R.gru_forward_small(k, self.dec_gates, th["dec_b_ih"], th["dec_w_hh"], th["dec_b_hh"], self.zh, 0, None, None,
                            self.dec_hs, self.dec_ghn, None, 1, STEPS, B, 1)                                    # :166
        k.gemm(L.GEMM_NT, 1, STEPS * B, p_, H, self.dec_hs, H, 0, th["out_w"], H, 0, self.pred, p_, 0, th["out_b"], 0)  # :167
        k.mse_fwd_bwd(self.pred, self.enc_in, self.sse, self.dpred, None, 1, STEPS, B * p_)                     # :601
        k.dot_small(self.sse, 1, 1.0 / (STEPS * B * p_), self.loss)

    def forward_staged(self):
        self.eps.copy_(self.eps_next)
        self.forward(None)

    def backward(self, beta_e: float = 1.0, dlat_extra: Optional[torch.Tensor] = None):
        k, th, g, B, p_ = self.k, self.theta, self.grad, self.B, self.p
        TB = STEPS * B
        k.gemm(L.GEMM_TN, 1, p_, H, TB, self.dpred, p_, 0, self.dec_hs, H, 0, g["out_w"], H, 0)
        k.gemm(L.GEMM_TN, 1, 1, p_, TB, self.ones_TB, 1, 0, self.dpred, p_, 0, g["out_b"], p_, 0)
        k.gemm(L.GEMM_NN, 1, TB, H, p_, self.dpred, p_, 0, th["out_w"], H, 0, self.dhs, H, 0)
        R.gru_backward_small(k, self.dec_gates, self.dec_ghn, self.dec_hs, self.zh, 0, th["dec_w_hh"], None, None, None, self.dhs,
                             g["dec_w_hh"].view(1, G, H), g["dec_b_hh"], g["dec_b_ih"], None, None, self.dzh, 1, STEPS, B, self.ws_gru,
                             self.ws_dwhh)
        k.proj_wgrad(self.dec_gates, self.dec_in, None, g["dec_w_ih"], 1, STEPS, B, p_, 1, self.ws_wgrad)
        k.tanh_bwd(self.dzh, self.zh, self.dpre, B * H)
        k.gemm(L.GEMM_TN, 1, H, H, B, self.dpre, H, 0, self.zlat, H, 0, g["hid_w"], H, 0)
        k.gemm(L.GEMM_TN, 1, 1, H, B, self.ones_B, 1, 0, self.dpre, H, 0, g["hid_b"], H, 0)
        k.gemm(L.GEMM_NN, 1, B, H, H, self.dpre, H, 0, th["hid_w"], H, 0, self.dz, H, 0)
        k.latent_bwd(self.dz, 1, None, self.lat, self.eps, beta_e, self.kl_form, self.dlat, None, B)
        if dlat_extra is not None:
            k.axpy(self.dlat, dlat_extra, B * 2 * H, 1.0)
        hT = self.enc_hs[0, STEPS - 1]
        if self.ws_lat is not None:
            k.latent_head_bwd(self.dlat, hT, th["lat_w"], g["lat_w"], g["lat_b"], self.dhT, B)
        else:
            k.gemm(L.GEMM_TN, 1, 2 * H, H, B, self.dlat, 2 * H, 0, hT, H, 0, g["lat_w"], H, 0)
            k.gemm(L.GEMM_TN, 1, 1, 2 * H, B, self.ones_B, 1, 0, self.dlat, 2 * H, 0, g["lat_b"], 2 * H, 0)
            k.gemm(L.GEMM_NN, 1, B, H, 2 * H, self.dlat, 2 * H, 0, th["lat_w"], H, 0, self.dhT, H, 0)
        R.gru_backward_small(k, self.enc_gates, self.enc_ghn, self.enc_hs, self.h0_zero, 0, th["enc_w_hh"], None, None, self.dhT, None,
                             g["enc_w_hh"].view(1, G, H), g["enc_b_hh"], g["enc_b_ih"], None, None, self.enc_dh0, 1, STEPS, B, self.ws_gru,
                             self.ws_dwhh)
        k.proj_wgrad(self.enc_gates, self.enc_in, None, g["enc_w_ih"], 1, STEPS, B, p_, 0, self.ws_wgrad)

    def adam_step(self, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        """optimizer.step() of torch.optim.Adam(vrae.parameters(), lr=1e-3) (:565, :613)."""
        self.k.adam_step_dev(self.theta.flat, self.grad.flat, self.exp_avg, self.exp_avg_sq, self.theta.numel,
                             lr, betas[0], betas[1], eps, self.adam_counter)

    def zero_grad(self):
        self.grad.flat.zero_()
