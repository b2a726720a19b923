"""MixtureCSRAE (reference CSRAE_new.py:113-150): an MLP auto-encoder (Bernoulli likelihood) regularised by the closed-form
Cauchy-Schwarz divergence to a learnable equal-weight GMM prior -- the CS-divergence family beyond CR-CS-RAE.py
(SURVEY.md 8(f4)).

    MixtureCSRAE(input_dim, hidden_dims=(400,), latent_dim=20, K=10, lambda_cs=1.0)
    .forward(x) -> (logits, mu_q, logvar_q);   .loss(x) -> (total, recon, cs)   (CSRAE_new.py:133-150)
    .loss_and_grad(x)                          the same three scalars + every parameter gradient in the grad arena

Kernels: the Linear layers are the exact FFMA GEMM (crvae_gemm_f32) with ReLU / its backward as crvae_act_*, the
reparameterisation is crvae_latent_fwd, the reconstruction term is crvae_bce_logits_fwd_bwd, and the divergence with its
gradients is the CR-CS-RAE kernel crvae_cs_div_fwd_bwd.  That kernel is built for 64 latent dimensions; a smaller latent
space is EMBEDDED: padded dimensions carry mu = 0, var = 1 on both sides, which multiplies all three overlap integrals of
D_CS = -log t1 + 0.5 log t2 + 0.5 log t3 by the same constant -- it cancels exactly -- and the gradients of the padding are
never read.  (latent_dim <= 64, K <= 32.)
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from . import lib as L
from .engine import Arena, H as _DPAD


class MixtureCSRAE(nn.Module):
    def __init__(self, input_dim: int, hidden_dims: Sequence[int] = (400,), latent_dim: int = 20, K: int = 10, lambda_cs: float = 1.0):
        super().__init__()
        if latent_dim > _DPAD or K > 32:
            raise ValueError("the CS-divergence kernel covers latent_dim <= 64 and K <= 32")
        self.k = L.kernels()
        dev = torch.device("cuda", torch.cuda.current_device()) if self.k.device_type == "cuda" else torch.device("cpu")
        self.device = dev
        self.input_dim, self.latent_dim, self.K, self.lambda_cs = int(input_dim), int(latent_dim), int(K), lambda_cs
        self.enc_dims = [self.input_dim, *hidden_dims]
        self.dec_dims = [self.latent_dim, *tuple(hidden_dims)[::-1]]
        shapes = {}
        for i, (a, b) in enumerate(zip(self.enc_dims[:-1], self.enc_dims[1:])):
            shapes[f"enc{i}_w"], shapes[f"enc{i}_b"] = (b, a), (b,)
        shapes["lat_w"], shapes["lat_b"] = (2 * latent_dim, self.enc_dims[-1]), (2 * latent_dim,)     # [mu_head ; logvar_head]
        for i, (a, b) in enumerate(zip(self.dec_dims[:-1], self.dec_dims[1:])):
            shapes[f"dec{i}_w"], shapes[f"dec{i}_b"] = (b, a), (b,)
        shapes["out_w"], shapes["out_b"] = (self.input_dim, self.dec_dims[-1]), (self.input_dim,)
        shapes["prior_mu"], shapes["prior_logvar"] = (K, latent_dim), (K, latent_dim)
        self.theta = Arena(shapes, dev)
        self.grad = self.theta.like()
        # declaration order of the reference (:121-123): encoder (net layers, mu_head, logvar_head), decoder (net, out_head), prior
        enc = [nn.Linear(a, b) for a, b in zip(self.enc_dims[:-1], self.enc_dims[1:])]
        mu_h, lv_h = nn.Linear(self.enc_dims[-1], latent_dim), nn.Linear(self.enc_dims[-1], latent_dim)
        dec = [nn.Linear(a, b) for a, b in zip(self.dec_dims[:-1], self.dec_dims[1:])]
        out_h = nn.Linear(self.dec_dims[-1], self.input_dim)
        pmu = torch.randn(K, latent_dim) * 0.05
        th = self.theta
        with torch.no_grad():
            for i, l in enumerate(enc):
                th[f"enc{i}_w"].copy_(l.weight); th[f"enc{i}_b"].copy_(l.bias)
            th["lat_w"][:latent_dim].copy_(mu_h.weight); th["lat_b"][:latent_dim].copy_(mu_h.bias)
            th["lat_w"][latent_dim:].copy_(lv_h.weight); th["lat_b"][latent_dim:].copy_(lv_h.bias)
            for i, l in enumerate(dec):
                th[f"dec{i}_w"].copy_(l.weight); th[f"dec{i}_b"].copy_(l.bias)
            th["out_w"].copy_(out_h.weight); th["out_b"].copy_(out_h.bias)
            th["prior_mu"].copy_(pmu)
        self.B = None

    def _tensors(self, arena):
        Z, n = self.latent_dim, {}
        for i in range(len(self.enc_dims) - 1):
            n[f"encoder.net.{2 * i}.weight"], n[f"encoder.net.{2 * i}.bias"] = arena[f"enc{i}_w"], arena[f"enc{i}_b"]
        n["encoder.mu_head.weight"], n["encoder.mu_head.bias"] = arena["lat_w"][:Z], arena["lat_b"][:Z]
        n["encoder.logvar_head.weight"], n["encoder.logvar_head.bias"] = arena["lat_w"][Z:], arena["lat_b"][Z:]
        for i in range(len(self.dec_dims) - 1):
            n[f"decoder.net.{2 * i}.weight"], n[f"decoder.net.{2 * i}.bias"] = arena[f"dec{i}_w"], arena[f"dec{i}_b"]
        n["decoder.out_head.weight"], n["decoder.out_head.bias"] = arena["out_w"], arena["out_b"]
        n["prior.mu"], n["prior.logvar"] = arena["prior_mu"], arena["prior_logvar"]
        return n

    def state_dict(self, *a, **kw):
        return {k: v.detach().clone() for k, v in self._tensors(self.theta).items()}

    def grad_dict(self):
        return {k: v.detach().clone() for k, v in self._tensors(self.grad).items()}

    def _alloc(self, B):
        dev, Z, k = self.device, self.latent_dim, self.k
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.B = B
        self.h_enc = [z(B, d) for d in self.enc_dims[1:]]
        self.h_dec = [z(B, d) for d in self.dec_dims[1:]]
        self.d_enc = [z(B, d) for d in self.enc_dims[1:]]
        self.d_dec = [z(B, d) for d in self.dec_dims[1:]]
        self.lat, self.dlat, self.zlat, self.eps, self.dz, self.kl = z(B, 2 * Z), z(B, 2 * Z), z(B, Z), z(B, Z), z(B, Z), z(1)
        self.logits, self.dlogits = z(B, self.input_dim), z(B, self.input_dim)
        self.recon_sum, self.cs_mean = z(1), z(1)
        self.lat64, self.dlat64 = z(B, 2 * _DPAD), z(B, 2 * _DPAD)
        self.pmu64, self.plv64, self.dpmu64, self.dplv64 = z(self.K, _DPAD), z(self.K, _DPAD), z(self.K, _DPAD), z(self.K, _DPAD)
        self.ones_B = torch.ones(B, 1, device=dev)
        self.ws_bce = torch.zeros(k.bce_logits_workspace(B * self.input_dim) // 4 + 4, dtype=torch.float32, device=dev)
        self.ws_cs = torch.zeros(k.cs_div_workspace(B, self.K) // 4 + 4, dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ forward (:133-137)
    def _linear(self, x, w, b, y, relu):
        k, Bn, N, Kd = self.k, x.shape[0], w.shape[0], w.shape[1]
        k.gemm(L.GEMM_NT, 1, Bn, N, Kd, x, Kd, 0, w, Kd, 0, y, N, 0, b, 0)
        if relu:
            k.act_fwd(y, y, Bn * N, 2)

    def forward(self, x: torch.Tensor):
        k, th, Z = self.k, self.theta, self.latent_dim
        x = x.to(self.device, torch.float32).contiguous()
        B = x.shape[0]
        if self.B != B:
            self._alloc(B)
        self.x = x
        h = x
        for i, y in enumerate(self.h_enc):
            self._linear(h, th[f"enc{i}_w"], th[f"enc{i}_b"], y, True)
            h = y
        k.gemm(L.GEMM_NT, 1, B, 2 * Z, h.shape[1], h, h.shape[1], 0, th["lat_w"], h.shape[1], 0, self.lat, 2 * Z, 0, th["lat_b"], 0)
        self.eps.copy_(torch.randn(B, Z).to(self.device, non_blocking=True))          # randn_like(std), :130
        k.latent_fwd(self.lat, self.eps, self.zlat, self.kl, B, L.KL_STANDARD, Z)
        h = self.zlat
        for i, y in enumerate(self.h_dec):
            self._linear(h, th[f"dec{i}_w"], th[f"dec{i}_b"], y, True)
            h = y
        self._linear(h, th["out_w"], th["out_b"], self.logits, False)
        return self.logits, self.lat[:, :Z], self.lat[:, Z:]

    # ------------------------------------------------------------------ loss (+ gradients) (:142-150)
    def loss(self, x: torch.Tensor):
        return self.loss_and_grad(x, want_grad=False)

    def loss_and_grad(self, x: torch.Tensor, want_grad: bool = True):
        k, th, g, Z, K = self.k, self.theta, self.grad, self.latent_dim, self.K
        self.forward(x)
        B, D = self.B, self.input_dim
        # recon = BCE-with-logits(sum) / B
        k.bce_logits_fwd_bwd(self.logits, self.x, self.recon_sum, self.dlogits if want_grad else None, B * D, 1.0 / B, self.ws_bce)
        # CS divergence on the 64-dimension embedding; the kernel's operand roles: lat64[:, :64] = log var_q, lat64[:, 64:] = mu_q
        self.lat64.zero_(); self.pmu64.zero_(); self.plv64.zero_()
        self.lat64[:, :Z].copy_(self.lat[:, Z:]); self.lat64[:, _DPAD:_DPAD + Z].copy_(self.lat[:, :Z])
        self.pmu64[:, :Z].copy_(th["prior_mu"]); self.plv64[:, :Z].copy_(th["prior_logvar"])
        k.cs_div_fwd_bwd(self.lat64, self.pmu64, self.plv64, B, K, float(self.lambda_cs) if want_grad else 0.0, self.cs_mean, self.dlat64,
                         self.dpmu64, self.dplv64, self.ws_cs)
        recon = self.recon_sum[0] / B
        cs = self.cs_mean[0]
        total = recon + self.lambda_cs * cs
        if not want_grad:
            return total, recon, cs
        g["prior_mu"].copy_(self.dpmu64[:, :Z]); g["prior_logvar"].copy_(self.dplv64[:, :Z])
        # decoder backward
        dy, h_in = self.dlogits, (self.h_dec[-1] if self.h_dec else self.zlat)
        self._linear_bwd(dy, h_in, th["out_w"], g["out_w"], g["out_b"], self.d_dec[-1] if self.h_dec else self.dz)
        for i in range(len(self.h_dec) - 1, -1, -1):
            k.act_bwd(self.d_dec[i], self.h_dec[i], self.d_dec[i], B * self.h_dec[i].shape[1], 2)
            h_in = self.h_dec[i - 1] if i > 0 else self.zlat
            self._linear_bwd(self.d_dec[i], h_in, th[f"dec{i}_w"], g[f"dec{i}_w"], g[f"dec{i}_b"], self.d_dec[i - 1] if i > 0 else self.dz)
        # reparameterisation: dmu = dz, dlogvar = dz * eps * 0.5 * std (beta = 0: no KL term), plus the divergence's gradient
        k.latent_bwd(self.dz.view(1, B, Z), 1, None, self.lat, self.eps, 0.0, L.KL_STANDARD, self.dlat, None, B, Z)
        self.dlat[:, :Z].add_(self.dlat64[:, _DPAD:_DPAD + Z]); self.dlat[:, Z:].add_(self.dlat64[:, :Z])
        h_last = self.h_enc[-1]
        Kd = h_last.shape[1]
        self._linear_bwd(self.dlat, h_last, th["lat_w"], g["lat_w"], g["lat_b"], self.d_enc[-1])
        for i in range(len(self.h_enc) - 1, -1, -1):
            k.act_bwd(self.d_enc[i], self.h_enc[i], self.d_enc[i], B * self.h_enc[i].shape[1], 2)
            h_in = self.h_enc[i - 1] if i > 0 else self.x
            self._linear_bwd(self.d_enc[i], h_in, th[f"enc{i}_w"], g[f"enc{i}_w"], g[f"enc{i}_b"], self.d_enc[i - 1] if i > 0 else None)
        return total, recon, cs

    def _linear_bwd(self, dy, x, w, dw, db, dx):
        """y = x w^T + b:  dw = dy^T x, db = column sums of dy, dx = dy w."""
        k, Bn, N, Kd = self.k, dy.shape[0], w.shape[0], w.shape[1]
        k.gemm(L.GEMM_TN, 1, N, Kd, Bn, dy, N, 0, x, Kd, 0, dw, Kd, 0)
        k.gemm(L.GEMM_TN, 1, 1, N, Bn, self.ones_B, 1, 0, dy, N, 0, db, N, 0)
        if dx is not None:
            k.gemm(L.GEMM_NN, 1, Bn, Kd, N, dy, N, 0, w, Kd, 0, dx, Kd, 0)
