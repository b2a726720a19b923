"""TEST INFRASTRUCTURE ONLY: a checker backend with the same call surface as
vae-connexe_b200/lib.py:Kernels, implemented on CPU tensors with the oracle's formulas
(oracle/crvae_oracle.py).  Installed through lib.set_test_backend() by the `not gpu` tests so the
HOST logic (engine orchestration, trainers, RNG order, head sharding over gloo) can be exercised
and compared with the golden vectors on a machine without a GPU.  Never used by the product path.
"""
from __future__ import annotations

import math

import torch

from oracle import crvae_oracle as O

H, G = 64, 192


class OracleKernels:
    device_type = "cpu"

    def __init__(self):
        self.launches = 0

    def launch_count(self):
        return self.launches

    def reset_launch_count(self):
        self.launches = 0

    def set_batch_tile(self, rows):
        pass

    # -- GEMM ------------------------------------------------------------------------------------
    def gemm(self, form, batch, M, N, K, A, lda, sA, Bm, ldb, sB, Cm, ldc, sC, bias=None, sBias=0, accumulate=False):
        a, b, c = A.reshape(-1), Bm.reshape(-1), Cm.reshape(-1)
        for i in range(batch):
            ao, bo, co = a.storage_offset() + i * sA, b.storage_offset() + i * sB, c.storage_offset() + i * sC
            if form == 0:      # NT
                Am = torch.as_strided(a, (M, K), (lda, 1), ao); Bt = torch.as_strided(b, (N, K), (ldb, 1), bo).t()
            elif form == 1:    # TN
                Am = torch.as_strided(a, (K, M), (lda, 1), ao).t(); Bt = torch.as_strided(b, (K, N), (ldb, 1), bo)
            else:              # NN
                Am = torch.as_strided(a, (M, K), (lda, 1), ao); Bt = torch.as_strided(b, (K, N), (ldb, 1), bo)
            out = Am @ Bt
            if bias is not None:
                out = out + bias.reshape(-1)[i * sBias:i * sBias + N]
            Cv = torch.as_strided(c, (M, N), (ldc, 1), co)
            if accumulate:
                Cv += out
            else:
                Cv.copy_(out)
        self.launches += 1

    # -- projection ------------------------------------------------------------------------------
    def proj_fwd(self, x, w_ih, b_ih, gates, P, T, B, K, t_skip):
        w = w_ih.reshape(P, G, K); x = x.reshape(T, B, K)
        gi = torch.einsum("tbk,pgk->ptbg", x[t_skip:], w) + b_ih.reshape(P, 1, 1, G)
        gates.reshape(-1)[: P * T * B * G].view(P, T, B, G)[:, t_skip:] = gi
        self.launches += 1

    def proj_wgrad_workspace(self, P, T, B, K):
        return 16

    def proj_wgrad(self, dgates, x, mask, dw_ih, P, T, B, K, t_skip, ws):
        dg = dgates.reshape(-1)[: P * T * B * G].view(P, T, B, G); x = x.reshape(T, B, K)
        dw = torch.einsum("ptbg,tbk->pgk", dg[:, t_skip:], x[t_skip:])
        if mask is not None:
            dw = dw * mask.reshape(P, 1, K).to(dw.dtype)
        dw_ih.reshape(P, G, K).copy_(dw)
        self.launches += 1

    # -- gather-packed ragged heads -----------------------------------------------------------------
    def gather_cols(self, x, cols, mask, xg, P, rows, K, Kp):
        xv = x.reshape(rows, K)
        g = xv[:, cols.long().reshape(P, Kp)]                      # [rows, P, Kp]
        g = g.permute(1, 0, 2)
        if mask is not None:
            g = g * mask.reshape(P, 1, Kp).to(g.dtype)
        xg.reshape(-1)[: P * rows * Kp].view(P, rows, Kp).copy_(g)
        self.launches += 1

    def proj_fwd_packed(self, xg, w_ih, b_ih, gates, P, T, B, Kp, t_skip):
        w = w_ih.reshape(P, G, Kp); x = xg.reshape(P, T, B, Kp)
        gi = torch.einsum("ptbk,pgk->ptbg", x[:, t_skip:], w) + b_ih.reshape(P, 1, 1, G)
        gates.reshape(-1)[: P * T * B * G].view(P, T, B, G)[:, t_skip:] = gi
        self.launches += 1

    def proj_wgrad_packed_workspace(self, P, T, B, Kp, K_dense):
        return 16

    def proj_wgrad_packed(self, dgates, xg, mask, dw_ih, P, T, B, Kp, K_dense, t_skip, ws):
        dg = dgates.reshape(-1)[: P * T * B * G].view(P, T, B, G); x = xg.reshape(P, T, B, Kp)
        dw = torch.einsum("ptbg,ptbk->pgk", dg[:, t_skip:], x[:, t_skip:])
        if mask is not None:
            dw = dw * mask.reshape(P, 1, Kp).to(dw.dtype)
        dw_ih.reshape(P, G, Kp).copy_(dw)
        self.launches += 1

    # -- recurrence ------------------------------------------------------------------------------
    def gru_fwd(self, gates, b_ih, w_hh, b_hh, h0, h0_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip):
        gv = gates.reshape(-1)[: P * T * B * G].view(P, T, B, G)
        gi = gv.clone()
        gi[:, :t_skip] = b_ih.reshape(P, 1, 1, G)
        h0v = h0.reshape(B, H) if h0_stride == 0 else h0.reshape(P, B, H)
        hs_, r, z, n, g = O.gru_forward(gi, h0v, w_hh.reshape(P, G, H), b_hh.reshape(P, G))
        gv.copy_(torch.cat([r, z, n], -1))
        hs.reshape(-1)[: P * T * B * H].view(P, T, B, H).copy_(hs_[:, 1:])
        ghn.reshape(-1)[: P * T * B * H].view(P, T, B, H).copy_(g)
        if w_lin is not None:
            pr = torch.einsum("ptbh,ph->ptb", hs_[:, 1:], w_lin.reshape(P, H)) + b_lin.reshape(P, 1, 1)
            pred.reshape(-1)[: P * T * B].view(P, T, B).copy_(pr)
        self.launches += 1

    def gru_bwd_workspace(self, P, B):
        return 16

    def gru_bwd(self, gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, dw_hh, db_hh, db_ih, dw_lin,
                db_lin, dh0, P, T, B, ws):
        gv = gates.reshape(-1)[: P * T * B * G].view(P, T, B, G)
        r, z, n = gv[..., :H].clone(), gv[..., H:2 * H].clone(), gv[..., 2 * H:].clone()
        hsv = hs.reshape(-1)[: P * T * B * H].view(P, T, B, H)
        h0v = (h0.reshape(1, B, H).expand(P, B, H) if h0_stride == 0 else h0.reshape(P, B, H))
        hs_full = torch.cat([h0v[:, None], hsv], 1)
        dh_out = torch.zeros(P, T, B, H)
        if w_lin is not None:
            dp = dpred.reshape(-1)[: P * T * B].view(P, T, B)
            dh_out = dh_out + dp[..., None] * w_lin.reshape(P, 1, 1, H)
        if dhs is not None:
            dh_out = dh_out + dhs.reshape(-1)[: P * T * B * H].view(P, T, B, H)
        dl = None if dh_last is None else dh_last.reshape(-1)[: P * B * H].view(P, B, H)
        dgi, dW, dbh, d0 = O.gru_backward(dh_out, hs_full, r, z, n,
                                          ghn.reshape(-1)[: P * T * B * H].view(P, T, B, H),
                                          w_hh.reshape(P, G, H), dh_last=dl)
        gv.copy_(dgi)
        dw_hh.reshape(P, G, H).copy_(dW); db_hh.reshape(P, G).copy_(dbh); db_ih.reshape(P, G).copy_(dgi.sum((1, 2)))
        if w_lin is not None:
            dw_lin.reshape(P, H).copy_(torch.einsum("ptb,ptbh->ph", dp, hsv))
            db_lin.reshape(P).copy_(dp.sum((1, 2)))
        dh0.reshape(-1)[: P * B * H].view(P, B, H).copy_(d0)
        self.launches += 2

    # -- latent ----------------------------------------------------------------------------------
    @staticmethod
    def _kl_terms(mu, lv, form):
        if form == 1:
            return 1 + mu - lv * lv - torch.exp(mu)
        return 1 + lv - mu * mu - torch.exp(lv)

    def latent_fwd(self, lat, eps, z, kl_out, B, kl_form, Z=64):
        H = Z
        mu, lv = lat[:, :H], lat[:, H:]
        if kl_form == 2:                   # Family-B: lv holds log sigma (CRVAE.py:72-75, :169)
            z.copy_(mu + torch.exp(lv) * 0.5 * eps.reshape(B, H))
            kl_out[0] = (-0.5 * (1 + 2 * lv - mu * mu - torch.exp(2 * lv))).sum(-1).mean(0)
            self.launches += 1
            return
        z.copy_(mu + torch.exp(0.5 * lv) * eps.reshape(B, H))
        kl_out[0] = (-0.5 * self._kl_terms(mu, lv, kl_form)).sum(-1).mean(0)
        self.launches += 1

    def latent_bwd(self, dh0, P, dz_extra, lat, eps, beta, kl_form, dlat, dz_out, B, Z=64):
        H = Z
        dz = torch.zeros(B, H)
        if P > 0:
            dz = dh0.reshape(-1)[: P * B * H].view(P, B, H).sum(0)
        if dz_out is not None:
            dz_out.copy_(dz)
        self.launches += 1
        if dlat is None:
            return
        if dz_extra is not None:
            dz = dz + dz_extra
        mu, lv = lat[:, :H], lat[:, H:]
        if kl_form == 2:
            dlat[:, :H] = dz + beta * mu / B
            dlat[:, H:] = dz * eps.reshape(B, H) * 0.5 * torch.exp(lv) + beta * (-(1 - torch.exp(2 * lv)) / B)
            return
        if kl_form == 1:
            dkm, dkl = -0.5 * (1 - torch.exp(mu)) / B, lv / B
        else:
            dkm, dkl = mu / B, -0.5 * (1 - torch.exp(lv)) / B
        dlat[:, :H] = dz + beta * dkm
        dlat[:, H:] = dz * eps.reshape(B, H) * 0.5 * torch.exp(0.5 * lv) + beta * dkl

    def mse_fwd_bwd(self, pred, target, sse, dpred, err, P, T, B, dscale=0.0):
        pr = pred.reshape(-1)[: P * T * B].view(P, T, B); tg = target.reshape(P, T, B)
        d = pr - tg
        sse.reshape(-1)[:P] = (d * d).sum((1, 2))
        if dpred is not None:
            dpred.reshape(-1)[: P * T * B].view(P, T, B).copy_((dscale if dscale > 0 else 2.0 / (T * B)) * d)
        if err is not None:
            err.reshape(-1)[: P * T * B].view(P, T, B).copy_(tg - pr)
        self.launches += 1

    # -- updates ---------------------------------------------------------------------------------
    def gd_step(self, theta, grad, n, lr):
        theta.reshape(-1)[:n] -= lr * grad.reshape(-1)[:n]
        self.launches += 1

    def gd_prox_gc(self, w_ih, dw_ih, mask, col_norm, P, K, lr, thr, do_prox):
        w = w_ih.reshape(P, G, K)
        if dw_ih is not None:
            w -= lr * dw_ih.reshape(P, G, K)
        if mask is not None:
            w *= mask.reshape(P, 1, K).to(w.dtype)
        if do_prox:
            norm = torch.norm(w, dim=1, keepdim=True)
            w.copy_((w / torch.clamp(norm, min=thr)) * torch.clamp(norm - thr, min=0.0))
        if col_norm is not None:
            col_norm.reshape(-1)[: P * K].view(P, K).copy_(torch.norm(w, dim=1))
        self.launches += 1

    def adam_step(self, theta, grad, m, v, n, lr, b1, b2, eps, step):
        th, g = theta.reshape(-1)[:n], grad.reshape(-1)[:n]
        mm, vv = m.reshape(-1)[:n], v.reshape(-1)[:n]
        mm.lerp_(g, 1 - b1)
        vv.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        denom = (vv.sqrt() / math.sqrt(bc2)).add_(eps)
        th.addcdiv_(mm, denom, value=-(lr / bc1))
        self.launches += 1

    def adam_step_dev(self, theta, grad, m, v, n, lr, b1, b2, eps, counter):
        self.adam_step(theta, grad, m, v, n, lr, b1, b2, eps, int(counter[0]) + 1)
        counter[0] += 1

    def cs_div_workspace(self, B, K):
        return 16

    def cs_div_fwd_bwd(self, lat, prior_mu, prior_logvar, B, K, scale_loss, cs_mean, dlat, dprior_mu, dprior_logvar, ws):
        cs, dl, dm, dv = O.cs_head(lat.reshape(B, 2 * H), prior_mu.reshape(K, H), prior_logvar.reshape(K, H), scale_loss)
        cs_mean[0] = cs
        dlat.reshape(B, 2 * H).copy_(dl); dprior_mu.reshape(K, H).copy_(dm); dprior_logvar.reshape(K, H).copy_(dv)
        self.launches += 3

    def tanh_fwd(self, x, y, n):
        y.reshape(-1)[:n] = torch.tanh(x.reshape(-1)[:n]); self.launches += 1

    def tanh_bwd(self, dy, y, dx, n):
        dx.reshape(-1)[:n] = dy.reshape(-1)[:n] * (1 - y.reshape(-1)[:n] ** 2); self.launches += 1

    def act_fwd(self, x, y, n, kind):
        v = x.reshape(-1)[:n]
        y.reshape(-1)[:n] = [torch.tanh, torch.sigmoid, torch.relu, lambda t: t][kind](v); self.launches += 1

    def act_bwd(self, dy, y, dx, n, kind):
        v = y.reshape(-1)[:n]
        gy = [1 - v * v, v * (1 - v), (v > 0).float(), torch.ones_like(v)][kind]
        dx.reshape(-1)[:n] = dy.reshape(-1)[:n] * gy; self.launches += 1

    def transpose(self, src, dst, rows, cols):
        dst.reshape(-1)[: rows * cols].view(cols, rows).copy_(src.reshape(-1)[: rows * cols].view(rows, cols).t())
        self.launches += 1

    def bce_logits_workspace(self, n):
        return 16

    def bce_logits_fwd_bwd(self, logits, x, sum_out, dlogits, n, dscale, ws):
        l, t = logits.reshape(-1)[:n], x.reshape(-1)[:n]
        sum_out[0] = torch.nn.functional.binary_cross_entropy_with_logits(l, t, reduction="sum")
        if dlogits is not None:
            dlogits.reshape(-1)[:n] = (torch.sigmoid(l) - t) * dscale
        self.launches += 1

    def ista_rows(self, w, dw, row_norm, rows, cols, lr, thr, do_prox):
        wv = w.reshape(-1)[: rows * cols].view(rows, cols)
        tmp = wv if dw is None else wv - torch.tensor(lr, dtype=torch.float32) * dw.reshape(-1)[: rows * cols].view(rows, cols)
        nu = torch.norm(tmp, dim=1, keepdim=True)
        shrink = torch.clamp(1 - torch.tensor(thr, dtype=torch.float32) / nu, min=0.) if do_prox else torch.ones_like(nu)
        shrink = torch.nan_to_num(shrink, nan=0.0)
        if do_prox or dw is not None:
            wv.copy_(tmp * shrink)
        if row_norm is not None:
            row_norm.reshape(-1)[:rows] = (nu * shrink).reshape(-1)
        self.launches += 1

    def gen_scatter(self, y, noise, x, x_hi, x_lo, out, B, p, t, steps, base, rem, widest, scale):
        yv = y.reshape(-1, widest, B)
        cols = []
        for j in range(p):
            if j < rem * (base + 1):
                r, l = divmod(j, base + 1)
            else:
                r, l = divmod(j - rem * (base + 1), base)
                r += rem
            cols.append(yv[r, l])
        v = torch.stack(cols, 1)                                   # [B, p]
        if noise is not None:
            v = v + torch.tensor(scale, dtype=torch.float32) * noise[:, t, :]
        x.reshape(B, p).copy_(v)
        out[:, t, :] = v
        self.launches += 1

    def sumsq(self, x, n, out):
        out[0] = (x.reshape(-1)[:n] ** 2).sum()
        self.launches += 1

    def dot_small(self, x, n, scale, out):
        out[0] = x.reshape(-1)[:n].sum() * scale
        self.launches += 1

    def axpy(self, y, x, n, alpha):
        y.reshape(-1)[:n] += alpha * x.reshape(-1)[:n]
        self.launches += 1
