"""Family-B CR-VAE (the repository owner's re-derivation, reference CRVAE.py:55-199) on the fused kernels.

    CRVAE(D, H, Z, tau)                 encoder GRU(D->H) -> fc_mu / fc_logsig (H->Z);  z = mu + exp(logsig)*0.5*eps;
                                        h0 = tanh(z2h z);  head p: x_sel = dec_in @ W_in[p] (D x H) -> GRU(H->H) -> fc_out(H->1)
    .granger_matrix(thr)                A[p, d] = ||W_in[p][d, :]||_2 > thr                          (CRVAE.py:126-131)
    .ista_step(lam, lr)                 W <- (W - lr*g) * max(1 - lr*lam/||row||, 0)                 (CRVAE.py:134-150)
    CRVAETrainer(model, lam_l1, lr)     .step_stage1(x) / .step_stage2(x): Adam on everything but W_in, ISTA on W_in in
                                        stage 1, the ErrorVAE (GRU hidden H/2, latent Z/2) in stage 2    (CRVAE.py:153-199)

How it maps onto the Family-A engine's kernels ("the same engine with a two-stage projection"):
  * the two-stage projection (dec_in @ W_in[p]) @ W_ih[p]^T collapses into ONE projection with the effective first-layer
    weight  W_eff[p] = W_ih[p] . W_in[p]^T  [3H x D], rebuilt each step by a small batched GEMM; the projection, the
    recurrent kernels (low-latency 16-row tiles: P = D heads), the per-head Linear(H,1) and the weight-gradient GEMM are the
    Family-A ones; the chain rule back through W_eff is two more batched GEMMs (dW_ih = dW_eff . W_in, dW_in = dW_eff^T . W_ih);
  * the latent head uses KL form CRVAE_KL_LOGSIGMA (log sigma, std = 0.5*exp(logsig));
  * ISTA is crvae_ista_rows (one warp per row of W_in), Adam is crvae_adam_step_dev with one step counter per parameter group
    (torch.optim.Adam skips the ErrorVAE's parameters while they have no gradient, so their step count starts in stage 2);
  * the ErrorVAE's GRUs have hidden size H/2 = 32: they run on the H = 64 kernels EMBEDDED in zero-padded weights (the extra
    units see zero weights, start at zero and stay exactly zero; every gradient of a padding entry is exactly zero, so Adam
    never moves it).
Not reproduced: `generate()` (CRVAE.py:104-123) -- not on the training path; the W_in gradient of stage 2 (the reference
masks it and never applies it: W_in is not in the optimizer and stage 2 does not call ista_step).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import lib as L
from . import rec as R
from .engine import Arena, G, H


def _f32(x: float) -> float:
    return float(np.float32(x))


def _unpad_gate_rows(src: torch.Tensor, h: int, cols: Optional[int] = None) -> torch.Tensor:
    parts = [src[g * H:g * H + h] for g in range(3)]
    out = torch.cat(parts, 0)
    return out[:, :cols] if cols is not None else out


class CRVAE(nn.Module):
    """Mirror of the reference's Family-B CRVAE(D, H, Z, tau) (CRVAE.py:55-131)."""

    def __init__(self, D: int, H_: int, Z: int, tau: int, device: Optional[str] = None):
        super().__init__()
        if int(H_) != H:
            raise ValueError(f"kernels are built for H={H} (CRVAE.py:242 uses H=64); got {H_}")
        self.k = L.kernels()
        self.D, self.H, self.Z, self.tau = int(D), H, int(Z), int(tau)
        self.He, self.Ze = H // 2, self.Z // 2                       # ErrorVAE(D, H//2, Z//2), CRVAE.py:66
        dev = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if self.k.device_type == "cuda" else torch.device("cpu"))
        self.device = dev
        D_, Z_, Ze = self.D, self.Z, self.Ze
        shapes = {
            # main model, Adam group 1
            "enc_w_ih": (G, D_), "enc_w_hh": (G, H), "enc_b_ih": (G,), "enc_b_hh": (G,),
            "lat_w": (2 * Z_, H), "lat_b": (2 * Z_,), "z2h_w": (H, Z_), "z2h_b": (H,),
            "w_ih": (D_, G, H), "w_hh": (D_, G, H), "b_ih": (D_, G), "b_hh": (D_, G), "w_out": (D_, H), "b_out": (D_,),
            # ErrorVAE (hidden 32 embedded in 64), Adam group 2
            "e_enc_w_ih": (G, D_), "e_enc_w_hh": (G, H), "e_enc_b_ih": (G,), "e_enc_b_hh": (G,),
            "e_dec_w_ih": (G, D_), "e_dec_w_hh": (G, H), "e_dec_b_ih": (G,), "e_dec_b_hh": (G,),
            "e_lat_w": (2 * Ze, H), "e_lat_b": (2 * Ze,), "e_z2h_w": (H, Ze), "e_z2h_b": (H,), "e_out_w": (D_, H), "e_out_b": (D_,),
            # ISTA group
            "W_in": (D_, D_, H),
        }
        self.theta = Arena(shapes, dev)
        self.grad = self.theta.like()
        self.off_err = self.theta.offsets["e_enc_w_ih"]
        self.off_win = self.theta.offsets["W_in"]
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.theta.flat), torch.zeros_like(self.theta.flat)
        self.cnt_main = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cnt_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.row_norm = torch.zeros(D_, D_, dtype=torch.float32, device=dev)
        self._init_like_reference()
        self.B = None

    # ------------------------------------------------------------------ init (declaration order of CRVAE.py:58-66)
    def _init_like_reference(self):
        th, D_, Z_, He, Ze = self.theta, self.D, self.Z, self.He, self.Ze
        enc = nn.GRU(D_, H, batch_first=True); fc_mu, fc_ls = nn.Linear(H, Z_), nn.Linear(H, Z_)      # Encoder :8-13
        z2h = nn.Linear(Z_, H)                                                                        # :60
        W_in = [0.01 * torch.randn(D_, H) for _ in range(D_)]                                         # :62-64
        heads = [(nn.GRU(H, H, batch_first=True), nn.Linear(H, 1)) for _ in range(D_)]                # :65, DecoderHead :22-26
        e_enc, e_dec = nn.GRU(D_, He, batch_first=True), nn.GRU(D_, He, batch_first=True)             # ErrorVAE :34-40
        e_mu, e_ls = nn.Linear(He, Ze), nn.Linear(He, Ze)
        e_z2h, e_out = nn.Linear(Ze, He), nn.Linear(He, D_)
        with torch.no_grad():
            th["enc_w_ih"].copy_(enc.weight_ih_l0); th["enc_w_hh"].copy_(enc.weight_hh_l0)
            th["enc_b_ih"].copy_(enc.bias_ih_l0); th["enc_b_hh"].copy_(enc.bias_hh_l0)
            th["lat_w"][:Z_].copy_(fc_mu.weight); th["lat_w"][Z_:].copy_(fc_ls.weight)
            th["lat_b"][:Z_].copy_(fc_mu.bias); th["lat_b"][Z_:].copy_(fc_ls.bias)
            th["z2h_w"].copy_(z2h.weight); th["z2h_b"].copy_(z2h.bias)
            th["W_in"].copy_(torch.stack(W_in))
            th["w_ih"].copy_(torch.stack([g.weight_ih_l0 for g, _ in heads])); th["w_hh"].copy_(torch.stack([g.weight_hh_l0 for g, _ in heads]))
            th["b_ih"].copy_(torch.stack([g.bias_ih_l0 for g, _ in heads])); th["b_hh"].copy_(torch.stack([g.bias_hh_l0 for g, _ in heads]))
            th["w_out"].copy_(torch.stack([l.weight[0] for _, l in heads])); th["b_out"].copy_(torch.stack([l.bias[0] for _, l in heads]))
            self._load_err_vae({"enc.weight_ih_l0": e_enc.weight_ih_l0, "enc.weight_hh_l0": e_enc.weight_hh_l0,
                                "enc.bias_ih_l0": e_enc.bias_ih_l0, "enc.bias_hh_l0": e_enc.bias_hh_l0,
                                "dec.weight_ih_l0": e_dec.weight_ih_l0, "dec.weight_hh_l0": e_dec.weight_hh_l0,
                                "dec.bias_ih_l0": e_dec.bias_ih_l0, "dec.bias_hh_l0": e_dec.bias_hh_l0,
                                "mu.weight": e_mu.weight, "mu.bias": e_mu.bias, "logσ.weight": e_ls.weight, "logσ.bias": e_ls.bias,
                                "z2h.weight": e_z2h.weight, "z2h.bias": e_z2h.bias, "out.weight": e_out.weight, "out.bias": e_out.bias})

    def _load_err_vae(self, sd):
        """Reference-shaped ErrorVAE tensors (hidden He = 32) -> zero-padded H = 64 storage."""
        th, He, Ze = self.theta, self.He, self.Ze
        dev = self.device
        with torch.no_grad():
            for pre in ("enc", "dec"):
                w_ih, w_hh = th[f"e_{pre}_w_ih"], th[f"e_{pre}_w_hh"]
                b_ih, b_hh = th[f"e_{pre}_b_ih"], th[f"e_{pre}_b_hh"]
                w_ih.zero_(); w_hh.zero_(); b_ih.zero_(); b_hh.zero_()
                for g in range(3):
                    w_ih[g * H:g * H + He].copy_(sd[f"{pre}.weight_ih_l0"][g * He:(g + 1) * He].to(dev))
                    w_hh[g * H:g * H + He, :He].copy_(sd[f"{pre}.weight_hh_l0"][g * He:(g + 1) * He].to(dev))
                    b_ih[g * H:g * H + He].copy_(sd[f"{pre}.bias_ih_l0"][g * He:(g + 1) * He].to(dev))
                    b_hh[g * H:g * H + He].copy_(sd[f"{pre}.bias_hh_l0"][g * He:(g + 1) * He].to(dev))
            th["e_lat_w"].zero_()
            th["e_lat_w"][:Ze, :He].copy_(sd["mu.weight"].to(dev)); th["e_lat_w"][Ze:, :He].copy_(sd["logσ.weight"].to(dev))
            th["e_lat_b"][:Ze].copy_(sd["mu.bias"].to(dev)); th["e_lat_b"][Ze:].copy_(sd["logσ.bias"].to(dev))
            th["e_z2h_w"].zero_(); th["e_z2h_b"].zero_()
            th["e_z2h_w"][:He].copy_(sd["z2h.weight"].to(dev)); th["e_z2h_b"][:He].copy_(sd["z2h.bias"].to(dev))
            th["e_out_w"].zero_()
            th["e_out_w"][:, :He].copy_(sd["out.weight"].to(dev)); th["e_out_b"].copy_(sd["out.bias"].to(dev))

    def _err_vae_tensors(self, arena):
        """The ErrorVAE slices of an arena (parameters or gradients) in the reference's shapes."""
        He, Ze = self.He, self.Ze
        out = {}
        for pre in ("enc", "dec"):
            out[f"err_vae.{pre}.weight_ih_l0"] = _unpad_gate_rows(arena[f"e_{pre}_w_ih"], He)
            out[f"err_vae.{pre}.weight_hh_l0"] = _unpad_gate_rows(arena[f"e_{pre}_w_hh"], He, He)
            out[f"err_vae.{pre}.bias_ih_l0"] = _unpad_gate_rows(arena[f"e_{pre}_b_ih"], He)
            out[f"err_vae.{pre}.bias_hh_l0"] = _unpad_gate_rows(arena[f"e_{pre}_b_hh"], He)
        out["err_vae.mu.weight"], out["err_vae.mu.bias"] = arena["e_lat_w"][:Ze, :He], arena["e_lat_b"][:Ze]
        out["err_vae.logσ.weight"], out["err_vae.logσ.bias"] = arena["e_lat_w"][Ze:, :He], arena["e_lat_b"][Ze:]
        out["err_vae.z2h.weight"], out["err_vae.z2h.bias"] = arena["e_z2h_w"][:He], arena["e_z2h_b"][:He]
        out["err_vae.out.weight"], out["err_vae.out.bias"] = arena["e_out_w"][:, :He], arena["e_out_b"]
        return out

    def _tensors(self, arena):
        Z_ = self.Z
        sd = {"encoder.gru.weight_ih_l0": arena["enc_w_ih"], "encoder.gru.weight_hh_l0": arena["enc_w_hh"],
              "encoder.gru.bias_ih_l0": arena["enc_b_ih"], "encoder.gru.bias_hh_l0": arena["enc_b_hh"],
              "encoder.fc_mu.weight": arena["lat_w"][:Z_], "encoder.fc_mu.bias": arena["lat_b"][:Z_],
              "encoder.fc_logsig.weight": arena["lat_w"][Z_:], "encoder.fc_logsig.bias": arena["lat_b"][Z_:],
              "z2h.weight": arena["z2h_w"], "z2h.bias": arena["z2h_b"]}
        for p in range(self.D):
            sd[f"W_in.{p}"] = arena["W_in"][p]
        for p in range(self.D):
            sd[f"heads.{p}.gru.weight_ih_l0"] = arena["w_ih"][p]; sd[f"heads.{p}.gru.weight_hh_l0"] = arena["w_hh"][p]
            sd[f"heads.{p}.gru.bias_ih_l0"] = arena["b_ih"][p]; sd[f"heads.{p}.gru.bias_hh_l0"] = arena["b_hh"][p]
            sd[f"heads.{p}.fc_out.weight"] = arena["w_out"][p:p + 1]; sd[f"heads.{p}.fc_out.bias"] = arena["b_out"][p:p + 1]
        sd.update(self._err_vae_tensors(arena))
        return sd

    def state_dict(self, *a, **kw):
        """Reference-shaped state_dict (same keys, shapes and order as CRVAE.py's model.state_dict())."""
        return {k: v.detach().clone() for k, v in self._tensors(self.theta).items()}

    def grad_dict(self):
        return {k: v.detach().clone() for k, v in self._tensors(self.grad).items()}

    def load_state_dict(self, sd, strict: bool = True):
        th, Z_ = self.theta, self.Z
        dev = self.device
        with torch.no_grad():
            th["enc_w_ih"].copy_(sd["encoder.gru.weight_ih_l0"]); th["enc_w_hh"].copy_(sd["encoder.gru.weight_hh_l0"])
            th["enc_b_ih"].copy_(sd["encoder.gru.bias_ih_l0"]); th["enc_b_hh"].copy_(sd["encoder.gru.bias_hh_l0"])
            th["lat_w"][:Z_].copy_(sd["encoder.fc_mu.weight"]); th["lat_w"][Z_:].copy_(sd["encoder.fc_logsig.weight"])
            th["lat_b"][:Z_].copy_(sd["encoder.fc_mu.bias"]); th["lat_b"][Z_:].copy_(sd["encoder.fc_logsig.bias"])
            th["z2h_w"].copy_(sd["z2h.weight"]); th["z2h_b"].copy_(sd["z2h.bias"])
            for p in range(self.D):
                th["W_in"][p].copy_(sd[f"W_in.{p}"])
                th["w_ih"][p].copy_(sd[f"heads.{p}.gru.weight_ih_l0"]); th["w_hh"][p].copy_(sd[f"heads.{p}.gru.weight_hh_l0"])
                th["b_ih"][p].copy_(sd[f"heads.{p}.gru.bias_ih_l0"]); th["b_hh"][p].copy_(sd[f"heads.{p}.gru.bias_hh_l0"])
                th["w_out"][p].copy_(sd[f"heads.{p}.fc_out.weight"][0]); th["b_out"][p].copy_(sd[f"heads.{p}.fc_out.bias"][0])
            self._load_err_vae({k[len("err_vae."):]: v for k, v in sd.items() if k.startswith("err_vae.")})

    # ------------------------------------------------------------------ buffers
    def _alloc(self, B: int):
        k, dev, D_, T, Z_, Ze = self.k, self.device, self.D, self.tau, self.Z, self.Ze
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.B = B
        self.enc_in, self.dec_in, self.cur = z(T, B, D_), z(T, B, D_), z(T, B, D_)
        self.target = z(D_, T, B)
        self.enc_gates, self.enc_hs, self.enc_ghn = z(1, T, B, G), z(1, T, B, H), z(1, T, B, H)
        self.h0_zero, self.lat, self.dlat, self.zlat, self.eps = z(B, H), z(B, 2 * Z_), z(B, 2 * Z_), z(B, Z_), z(B, Z_)
        self.pre0, self.h0, self.dpre0, self.dzl = z(B, H), z(B, H), z(B, H), z(B, Z_)
        self.w_eff, self.dw_eff = z(D_, G, D_), z(D_, G, D_)
        self.gates, self.hs, self.ghn = z(D_, T, B, G), z(D_, T, B, H), z(D_, T, B, H)
        self.pred, self.dpred, self.err = z(D_, T, B), z(D_, T, B), z(D_, T, B)
        self.dh0, self.dh0_sum = z(D_, B, H), z(B, H)
        self.dhT, self.enc_dh0 = z(1, B, H), z(1, B, H)
        self.sse, self.kl, self.sse1, self.kl_e, self.mse_val = z(D_), z(1), z(1), z(1), z(1)
        self.ones_B, self.ones_TB = torch.ones(B, 1, device=dev), torch.ones(T * B, 1, device=dev)
        # ErrorVAE
        self.e_in = z(T, B, D_)
        self.e_enc_gates, self.e_enc_hs, self.e_enc_ghn = z(1, T, B, G), z(1, T, B, H), z(1, T, B, H)
        self.e_dec_gates, self.e_dec_hs, self.e_dec_ghn = z(1, T, B, G), z(1, T, B, H), z(1, T, B, H)
        self.e_lat, self.e_dlat, self.e_z, self.e_eps = z(B, 2 * Ze), z(B, 2 * Ze), z(B, Ze), z(B, Ze)
        self.e_pre0, self.e_h0, self.e_dpre0, self.e_dz = z(B, H), z(B, H), z(B, H), z(B, Ze)
        self.e_hat, self.recon_tbd, self.dcommon, self.e_dhs = z(T, B, D_), z(T, B, D_), z(T, B, D_), z(1, T, B, H)
        self.e_dh0, self.e_dhT, self.e_enc_dh0 = z(1, B, H), z(1, B, H), z(1, B, H)
        self.ws_gru = torch.zeros(k.gru_bwd_workspace(D_, B) // 4 + 4, dtype=torch.float32, device=dev)
        n = R.dwhh_workspace(k, D_, T, B)
        self.ws_dwhh = torch.zeros(n, dtype=torch.float32, device=dev) if n else None
        self.ws_wgrad = torch.zeros(max(k.proj_wgrad_workspace(D_, T, B, D_), k.proj_wgrad_workspace(1, T, B, D_)) // 4 + 4,
                                    dtype=torch.float32, device=dev)

    def _bind(self, x_past: torch.Tensor, x_cur: torch.Tensor):
        B = x_past.shape[0]
        if self.B != B:
            self._alloc(B)
        xp, xc = x_past.to(self.device, torch.float32), x_cur.to(self.device, torch.float32)
        self.enc_in.copy_(xp.transpose(0, 1))
        self.cur.copy_(xc.transpose(0, 1))
        self.dec_in[0].copy_(xp[:, -1]); self.dec_in[1:].copy_(xc[:, :-1].transpose(0, 1))          # CRVAE.py:80
        self.target.copy_(xc.permute(2, 1, 0))

    # ------------------------------------------------------------------ forward (CRVAE.py:70-101)
    def forward(self, x_past: torch.Tensor, x_cur: torch.Tensor, phase: int = 1):
        k, th, D_, T, Z_ = self.k, self.theta, self.D, self.tau, self.Z
        self._bind(x_past, x_cur)
        B = self.B
        self.phase = int(phase)
        k.proj_fwd(self.enc_in, th["enc_w_ih"], th["enc_b_ih"], self.enc_gates, 1, T, B, D_, 0)
        R.gru_forward_small(k, self.enc_gates, th["enc_b_ih"], th["enc_w_hh"], th["enc_b_hh"], self.h0_zero, 0, None, None,
                            self.enc_hs, self.enc_ghn, None, 1, T, B, 0)
        hT = self.enc_hs[0, T - 1]
        k.gemm(L.GEMM_NT, 1, B, 2 * Z_, H, hT, H, 0, th["lat_w"], H, 0, self.lat, 2 * Z_, 0, th["lat_b"], 0)
        self.eps.copy_(torch.randn(B, Z_).to(self.device, non_blocking=True))                    # randn_like(logsig), :74
        k.latent_fwd(self.lat, self.eps, self.zlat, self.kl, B, L.KL_LOGSIGMA, Z_)
        k.gemm(L.GEMM_NT, 1, B, H, Z_, self.zlat, Z_, 0, th["z2h_w"], Z_, 0, self.pre0, H, 0, th["z2h_b"], 0)
        k.tanh_fwd(self.pre0, self.h0, B * H)                                                      # :78
        # two-stage projection collapsed: W_eff[p] = W_ih[p] . W_in[p]^T  (x_sel = dec_in @ W_in[p], :86; GRU input weights)
        k.gemm(L.GEMM_NT, D_, G, D_, H, th["w_ih"], H, G * H, th["W_in"], H, D_ * H, self.w_eff, D_, G * D_)
        k.proj_fwd(self.dec_in, self.w_eff, th["b_ih"], self.gates, D_, T, B, D_, 0)
        R.gru_forward_small(k, self.gates, th["b_ih"], th["w_hh"], th["b_hh"], self.h0, 0, th["w_out"], th["b_out"],
                            self.hs, self.ghn, self.pred, D_, T, B, 0)                             # heads, :85-90
        recon = self.pred.permute(2, 1, 0)                                                         # [B,tau,D]
        mu, logsig = self.lat[:, :Z_], self.lat[:, Z_:]
        if self.phase == 1:
            return recon, mu, logsig, None, None
        # phase 2 (:96-101): eps = (x_cur - recon).detach(); eps_hat = err_vae(eps); recon_plus = recon + eps_hat
        k.mse_fwd_bwd(self.pred, self.target, self.sse, None, self.err, D_, T, B)
        k.transpose(self.err.view(D_, T * B), self.e_in, D_, T * B)
        self._err_vae_forward()
        k.transpose(self.pred.view(D_, T * B), self.recon_tbd, D_, T * B)
        k.axpy(self.e_hat, self.recon_tbd, T * B * D_, 1.0)                                        # e_hat <- recon_plus
        Ze = self.Ze
        return self.e_hat.permute(1, 0, 2), mu, logsig, self.e_lat[:, :Ze], self.e_lat[:, Ze:]

    def _err_vae_forward(self):
        """ErrorVAE.forward (:47-53) on the residual in self.e_in [tau,B,D]; leaves eps_hat in self.e_hat."""
        k, th, D_, T, B, Ze = self.k, self.theta, self.D, self.tau, self.B, self.Ze
        k.proj_fwd(self.e_in, th["e_enc_w_ih"], th["e_enc_b_ih"], self.e_enc_gates, 1, T, B, D_, 0)
        R.gru_forward_small(k, self.e_enc_gates, th["e_enc_b_ih"], th["e_enc_w_hh"], th["e_enc_b_hh"], self.h0_zero, 0, None, None,
                            self.e_enc_hs, self.e_enc_ghn, None, 1, T, B, 0)
        hT = self.e_enc_hs[0, T - 1]
        k.gemm(L.GEMM_NT, 1, B, 2 * Ze, H, hT, H, 0, th["e_lat_w"], H, 0, self.e_lat, 2 * Ze, 0, th["e_lat_b"], 0)
        self.e_eps.copy_(torch.randn(B, Ze).to(self.device, non_blocking=True))                   # :45
        k.latent_fwd(self.e_lat, self.e_eps, self.e_z, self.kl_e, B, L.KL_LOGSIGMA, Ze)
        k.gemm(L.GEMM_NT, 1, B, H, Ze, self.e_z, Ze, 0, th["e_z2h_w"], Ze, 0, self.e_pre0, H, 0, th["e_z2h_b"], 0)
        k.tanh_fwd(self.e_pre0, self.e_h0, B * H)
        k.proj_fwd(self.e_in, th["e_dec_w_ih"], th["e_dec_b_ih"], self.e_dec_gates, 1, T, B, D_, 0)   # dec(eps, h0), :52
        R.gru_forward_small(k, self.e_dec_gates, th["e_dec_b_ih"], th["e_dec_w_hh"], th["e_dec_b_hh"], self.e_h0, 0, None, None,
                            self.e_dec_hs, self.e_dec_ghn, None, 1, T, B, 0)
        k.gemm(L.GEMM_NT, 1, T * B, D_, H, self.e_dec_hs, H, 0, th["e_out_w"], H, 0, self.e_hat, D_, 0, th["e_out_b"], 0)

    # ------------------------------------------------------------------ losses + backward (CRVAETrainer :161-196)
    def loss_and_backward(self):
        """loss = mse(recon[_plus], x_cur) + kl_main (+ kl_err) with kl = -0.5*mean(1 + 2s - mu^2 - exp(2s)); gradients of
        every parameter into the grad arena.  Returns the loss as a device scalar."""
        k, th, g, D_, T, B, Z_, Ze = self.k, self.theta, self.grad, self.D, self.tau, self.B, self.Z, self.Ze
        n_all = T * B * D_
        if self.phase == 1:
            k.mse_fwd_bwd(self.pred, self.target, self.sse, self.dpred, None, D_, T, B, 2.0 / n_all)      # mean over B*tau*D
            k.dot_small(self.sse, D_, 1.0 / n_all, self.mse_val)
        else:
            k.mse_fwd_bwd(self.e_hat, self.cur, self.sse1, self.dcommon, None, 1, T, B * D_, 2.0 / n_all)
            k.dot_small(self.sse1, 1, 1.0 / n_all, self.mse_val)
            k.transpose(self.dcommon.view(T * B, D_), self.dpred, T * B, D_)                             # d recon = d recon_plus
            self._err_vae_backward()
        # decoder heads
        R.gru_backward_small(k, self.gates, self.ghn, self.hs, self.h0, 0, th["w_hh"], th["w_out"], self.dpred, None, None,
                             g["w_hh"], g["b_hh"], g["b_ih"], g["w_out"], g["b_out"], self.dh0, D_, T, B, self.ws_gru, self.ws_dwhh)
        k.proj_wgrad(self.gates, self.dec_in, None, self.dw_eff, D_, T, B, D_, 0, self.ws_wgrad)
        # chain rule through W_eff = W_ih . W_in^T
        k.gemm(L.GEMM_NN, D_, G, H, D_, self.dw_eff, D_, G * D_, th["W_in"], H, D_ * H, g["w_ih"], H, G * H)
        if self.phase == 1:
            k.gemm(L.GEMM_TN, D_, D_, H, G, self.dw_eff, D_, G * D_, th["w_ih"], H, G * H, g["W_in"], H, D_ * H)
        # h0 = tanh(z2h z) shared by every head
        k.latent_bwd(self.dh0, D_, None, None, None, 0.0, L.KL_LOGSIGMA, None, self.dh0_sum, B, H)
        k.tanh_bwd(self.dh0_sum, self.h0, self.dpre0, B * H)
        k.gemm(L.GEMM_TN, 1, H, Z_, B, self.dpre0, H, 0, self.zlat, Z_, 0, g["z2h_w"], Z_, 0)
        k.gemm(L.GEMM_TN, 1, 1, H, B, self.ones_B, 1, 0, self.dpre0, H, 0, g["z2h_b"], H, 0)
        k.gemm(L.GEMM_NN, 1, B, Z_, H, self.dpre0, H, 0, th["z2h_w"], Z_, 0, self.dzl, Z_, 0)
        # kl = mean over B*Z  ->  beta = 1/Z on the kernel's  mean_b sum_z
        k.latent_bwd(self.dzl.view(1, B, Z_), 1, None, self.lat, self.eps, 1.0 / Z_, L.KL_LOGSIGMA, self.dlat, None, B, Z_)
        hT = self.enc_hs[0, T - 1]
        k.gemm(L.GEMM_TN, 1, 2 * Z_, H, B, self.dlat, 2 * Z_, 0, hT, H, 0, g["lat_w"], H, 0)
        k.gemm(L.GEMM_TN, 1, 1, 2 * Z_, B, self.ones_B, 1, 0, self.dlat, 2 * Z_, 0, g["lat_b"], 2 * Z_, 0)
        k.gemm(L.GEMM_NN, 1, B, H, 2 * Z_, self.dlat, 2 * Z_, 0, th["lat_w"], H, 0, self.dhT, H, 0)
        R.gru_backward_small(k, self.enc_gates, self.enc_ghn, self.enc_hs, self.h0_zero, 0, th["enc_w_hh"], None, None, self.dhT, None,
                             g["enc_w_hh"].view(1, G, H), g["enc_b_hh"], g["enc_b_ih"], None, None, self.enc_dh0, 1, T, B,
                             self.ws_gru, self.ws_dwhh)
        k.proj_wgrad(self.enc_gates, self.enc_in, None, g["enc_w_ih"], 1, T, B, D_, 0, self.ws_wgrad)
        loss = self.mse_val[0] + self.kl[0] / Z_
        if self.phase == 2:
            loss = loss + self.kl_e[0] / Ze
        return loss

    def _err_vae_backward(self):
        k, th, g, D_, T, B, Ze = self.k, self.theta, self.grad, self.D, self.tau, self.B, self.Ze
        TB = T * B
        k.gemm(L.GEMM_TN, 1, D_, H, TB, self.dcommon, D_, 0, self.e_dec_hs, H, 0, g["e_out_w"], H, 0)
        k.gemm(L.GEMM_TN, 1, 1, D_, TB, self.ones_TB, 1, 0, self.dcommon, D_, 0, g["e_out_b"], D_, 0)
        k.gemm(L.GEMM_NN, 1, TB, H, D_, self.dcommon, D_, 0, th["e_out_w"], H, 0, self.e_dhs, H, 0)
        R.gru_backward_small(k, self.e_dec_gates, self.e_dec_ghn, self.e_dec_hs, self.e_h0, 0, th["e_dec_w_hh"], None, None, None, self.e_dhs,
                             g["e_dec_w_hh"].view(1, G, H), g["e_dec_b_hh"], g["e_dec_b_ih"], None, None, self.e_dh0, 1, T, B,
                             self.ws_gru, self.ws_dwhh)
        k.proj_wgrad(self.e_dec_gates, self.e_in, None, g["e_dec_w_ih"], 1, T, B, D_, 0, self.ws_wgrad)
        k.tanh_bwd(self.e_dh0.view(B, H), self.e_h0, self.e_dpre0, B * H)
        k.gemm(L.GEMM_TN, 1, H, Ze, B, self.e_dpre0, H, 0, self.e_z, Ze, 0, g["e_z2h_w"], Ze, 0)
        k.gemm(L.GEMM_TN, 1, 1, H, B, self.ones_B, 1, 0, self.e_dpre0, H, 0, g["e_z2h_b"], H, 0)
        k.gemm(L.GEMM_NN, 1, B, Ze, H, self.e_dpre0, H, 0, th["e_z2h_w"], Ze, 0, self.e_dz, Ze, 0)
        k.latent_bwd(self.e_dz.view(1, B, Ze), 1, None, self.e_lat, self.e_eps, 1.0 / Ze, L.KL_LOGSIGMA, self.e_dlat, None, B, Ze)
        hT = self.e_enc_hs[0, T - 1]
        k.gemm(L.GEMM_TN, 1, 2 * Ze, H, B, self.e_dlat, 2 * Ze, 0, hT, H, 0, g["e_lat_w"], H, 0)
        k.gemm(L.GEMM_TN, 1, 1, 2 * Ze, B, self.ones_B, 1, 0, self.e_dlat, 2 * Ze, 0, g["e_lat_b"], 2 * Ze, 0)
        k.gemm(L.GEMM_NN, 1, B, H, 2 * Ze, self.e_dlat, 2 * Ze, 0, th["e_lat_w"], H, 0, self.e_dhT, H, 0)
        R.gru_backward_small(k, self.e_enc_gates, self.e_enc_ghn, self.e_enc_hs, self.h0_zero, 0, th["e_enc_w_hh"], None, None, self.e_dhT, None,
                             g["e_enc_w_hh"].view(1, G, H), g["e_enc_b_hh"], g["e_enc_b_ih"], None, None, self.e_enc_dh0, 1, T, B,
                             self.ws_gru, self.ws_dwhh)
        k.proj_wgrad(self.e_enc_gates, self.e_in, None, g["e_enc_w_ih"], 1, T, B, D_, 0, self.ws_wgrad)

    # ------------------------------------------------------------------ graph readout + ISTA (CRVAE.py:126-150)
    def granger_matrix(self, thr: float = 1e-6) -> torch.Tensor:
        D_ = self.D
        self.k.ista_rows(self.theta["W_in"], None, self.row_norm, D_ * D_, H, 0.0, 0.0, False)
        return (self.row_norm > thr).float()

    def ista_step(self, lam: float, lr: float):
        D_ = self.D
        self.k.ista_rows(self.theta["W_in"], self.grad["W_in"], self.row_norm, D_ * D_, H, _f32(lr), _f32(lr * lam), True)
        self.grad["W_in"].zero_()                                                                  # :150


class CRVAETrainer:
    """Mirror of CRVAETrainer (CRVAE.py:153-199): Adam(lr) on every parameter except W_in, ISTA(lam_l1, lr) on W_in."""

    def __init__(self, model: CRVAE, λ_l1: float = 5e-2, lr: float = 1e-3):
        self.m, self.lr, self.λ = model, lr, λ_l1

    def _adam(self, lo: int, hi: int, counter):
        m = self.m
        m.k.adam_step_dev(m.theta.flat[lo:hi], m.grad.flat[lo:hi], m.exp_avg[lo:hi], m.exp_avg_sq[lo:hi], hi - lo,
                          self.lr, 0.9, 0.999, 1e-8, counter)

    def step_stage1(self, x_batch: torch.Tensor) -> float:
        m = self.m
        x_past, x_cur = torch.split(x_batch, m.tau, dim=1)                                         # :163
        m.forward(x_past, x_cur, phase=1)
        loss = m.loss_and_backward()
        m.ista_step(self.λ, self.lr)                                                               # :171, before opt.step
        self._adam(0, m.off_err, m.cnt_main)                                                       # the ErrorVAE has no gradient yet: Adam skips it
        return float(loss)

    def step_stage2(self, x_batch: torch.Tensor) -> float:
        m = self.m
        x_past, x_cur = torch.split(x_batch, m.tau, dim=1)
        m.forward(x_past, x_cur, phase=2)
        loss = m.loss_and_backward()
        self._adam(0, m.off_err, m.cnt_main)                                                       # W_in: gradient masked, never applied (:189-196)
        self._adam(m.off_err, m.off_win, m.cnt_err)
        return float(loss)
