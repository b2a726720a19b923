"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the CR-VAE training hot path.

A CPU restatement (torch CPU tensors, explicit formulas, hand-derived backward -- no nn.GRU, no
autograd) of the algorithm the reference runs in /root/reference/CRVAE_lorenz96.py.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file;
the product path (vae-connexe_b200/) never does.

Where the arithmetic lives: the reference delegates to PyTorch (unpinned; torch 2.11.0 here):
nn.GRU (CRVAE_lorenz96.py:104,:119,:192,:208,:133,:142), nn.Linear (:106,:120,:195-196),
torch.norm (:297,:311), autograd (:497) and optim.Adam (:565).  The formulas below restate
PyTorch's published GRU / Linear / MSE / Adam definitions at those call sites.

Parity pin: the reference has no tests or golden vectors of its own, so this
oracle is pinned against outputs of the reference itself, run in the build container through
oracle/ref_loader.py and committed under tests/golden/ (generator: tests/golden/make_golden.py).
tests/test_oracle_golden.py checks every function here against those fixtures.

Fused parameter layout (P = heads held, p = number of series = projection depth K, H hidden,
G = 3H with gate rows ordered [r; z; n] as in nn.GRU.weight_*):
  enc_w_ih [G,p]  enc_w_hh [G,H]  enc_b_ih [G]  enc_b_hh [G]      (gru_left, :192)
  mu_w [H,H] mu_b [H] std_w [H,H] std_b [H]                       (fc_mu / fc_std, :195-196)
  w_ih [P,G,p]  w_hh [P,G,H]  b_ih [P,G]  b_hh [P,G]              (networks[i].gru, :104)
  w_lin [P,H]  b_lin [P]                                          (networks[i].linear, :106)
  mask [P,p] bool -- head i reads column j iff connection[j, i] != 0 (:115, :201); masked-out
                     entries of w_ih are structural zeros (ragged phase-2 heads, :788-790).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

ENC_STEPS = 10   # CRVAE_lorenz96.py:208  X[:,1:11,:] of the zero-prepended window
DEC_STEPS = 10   # CRVAE_lorenz96.py:119  [zero step, X'[:,11:-1,:]] for context=20
ENC_KEYS = ("enc_w_ih", "enc_w_hh", "enc_b_ih", "enc_b_hh", "mu_w", "mu_b", "std_w", "std_b")
HEAD_KEYS = ("w_ih", "w_hh", "b_ih", "b_hh", "w_lin", "b_lin")
PARAM_KEYS = ENC_KEYS + HEAD_KEYS


# ----------------------------------------------------------------------------------------------
# state_dict <-> fused layout
# ----------------------------------------------------------------------------------------------
def connection_mask(connection: np.ndarray) -> np.ndarray:
    """mask[i, j] = head i reads input column j.  The reference indexes COLUMN i of
    `connection` (CRVAE_lorenz96.py:115 `np.where(connection!=0)` on `self.connection[:,i]`,
    :218), i.e. the transpose of the GC convention -- reproduced on purpose."""
    return (np.asarray(connection) != 0).T.copy()


def params_from_state_dict(sd: Dict[str, Tensor], connection: np.ndarray,
                           dtype=torch.float32) -> Params:
    """Fuse a reference-shaped CRVAE state_dict (keys as in SURVEY.md 8(a3))."""
    mask = connection_mask(connection)
    P, p = mask.shape
    G = sd["gru_left.weight_hh_l0"].shape[0]
    H = G // 3
    out: Params = {
        "enc_w_ih": sd["gru_left.weight_ih_l0"], "enc_w_hh": sd["gru_left.weight_hh_l0"],
        "enc_b_ih": sd["gru_left.bias_ih_l0"], "enc_b_hh": sd["gru_left.bias_hh_l0"],
        "mu_w": sd["fc_mu.weight"], "mu_b": sd["fc_mu.bias"],
        "std_w": sd["fc_std.weight"], "std_b": sd["fc_std.bias"],
    }
    w_ih = torch.zeros(P, G, p, dtype=dtype)
    for i in range(P):
        cols = np.where(mask[i])[0]
        w_ih[i][:, cols] = sd[f"networks.{i}.gru.weight_ih_l0"].to(dtype)
    out["w_ih"] = w_ih
    out["w_hh"] = torch.stack([sd[f"networks.{i}.gru.weight_hh_l0"] for i in range(P)])
    out["b_ih"] = torch.stack([sd[f"networks.{i}.gru.bias_ih_l0"] for i in range(P)])
    out["b_hh"] = torch.stack([sd[f"networks.{i}.gru.bias_hh_l0"] for i in range(P)])
    out["w_lin"] = torch.stack([sd[f"networks.{i}.linear.weight"][0] for i in range(P)])
    out["b_lin"] = torch.stack([sd[f"networks.{i}.linear.bias"][0] for i in range(P)])
    out = {k: v.detach().to(dtype).clone().contiguous() for k, v in out.items()}
    out["mask"] = torch.from_numpy(mask)
    assert out["w_hh"].shape == (P, G, H)
    return out


def state_dict_from_params(prm: Params) -> Dict[str, Tensor]:
    """Inverse of params_from_state_dict (ragged heads get their packed columns back)."""
    sd = {
        "gru_left.weight_ih_l0": prm["enc_w_ih"], "gru_left.weight_hh_l0": prm["enc_w_hh"],
        "gru_left.bias_ih_l0": prm["enc_b_ih"], "gru_left.bias_hh_l0": prm["enc_b_hh"],
        "fc_mu.weight": prm["mu_w"], "fc_mu.bias": prm["mu_b"],
        "fc_std.weight": prm["std_w"], "fc_std.bias": prm["std_b"],
    }
    mask = prm["mask"].numpy()
    for i in range(mask.shape[0]):
        cols = np.where(mask[i])[0]
        sd[f"networks.{i}.gru.weight_ih_l0"] = prm["w_ih"][i][:, cols]
        sd[f"networks.{i}.gru.weight_hh_l0"] = prm["w_hh"][i]
        sd[f"networks.{i}.gru.bias_ih_l0"] = prm["b_ih"][i]
        sd[f"networks.{i}.gru.bias_hh_l0"] = prm["b_hh"][i]
        sd[f"networks.{i}.linear.weight"] = prm["w_lin"][i][None, :]
        sd[f"networks.{i}.linear.bias"] = prm["b_lin"][i][None]
    return {k: v.clone() for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------
# input preparation (CRVAE_lorenz96.py:332-350, :205-208, :119, :484)
# ----------------------------------------------------------------------------------------------
def arrange_input(data: Tensor, context: int) -> Tuple[Tensor, Tensor]:
    """Window n = data[n:n+context]; target is the same shifted by one (:332-350)."""
    assert context >= 1 and isinstance(context, int)
    n = len(data) - context
    idx = torch.arange(n)[:, None] + torch.arange(context)[None, :]
    return data[idx].to(torch.float32), data[idx + 1].to(torch.float32)


def split_window(X: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """X (B, 20, p) -> encoder input (Te,B,p) = X[:,0:10] (:208 on the zero-prepended window),
    decoder input (Td,B,p) = [0, X[:,10:19]] (:119), target (Td,B,p) = X[:,10:20] (:484)."""
    enc_in = X[:, 0:ENC_STEPS].transpose(0, 1).contiguous()
    dec_in = torch.cat([torch.zeros_like(X[:, 0:1]), X[:, ENC_STEPS:-1]], 1).transpose(0, 1).contiguous()
    target = X[:, ENC_STEPS:].transpose(0, 1).contiguous()
    return enc_in, dec_in, target


# ----------------------------------------------------------------------------------------------
# GRU cell math (nn.GRU definition; operation order verified bit-exact vs ATen CPU)
# ----------------------------------------------------------------------------------------------
def gru_forward(gi: Tensor, h0: Tensor, w_hh: Tensor, b_hh: Tensor):
    """Batched-over-heads GRU recurrence.
    gi [P,T,B,G] input projections (bias included), h0 [P,B,H] or [B,H] (shared),
    w_hh [P,G,H], b_hh [P,G].  Returns hs [P,T+1,B,H] and the saved gates r,z,n,ghn [P,T,B,H]."""
    P, T, B, G = gi.shape
    H = G // 3
    if h0.dim() == 2:
        h0 = h0.unsqueeze(0).expand(P, B, H)
    hs = [h0]
    rs, zs, ns, ghns = [], [], [], []
    h = h0
    w_hh_t = w_hh.transpose(1, 2)
    for t in range(T):
        gh = torch.bmm(h, w_hh_t) + b_hh[:, None, :]
        g = gi[:, t]
        r = torch.sigmoid(g[..., :H] + gh[..., :H])
        z = torch.sigmoid(g[..., H:2 * H] + gh[..., H:2 * H])
        ghn = gh[..., 2 * H:]
        n = torch.tanh(g[..., 2 * H:] + r * ghn)
        h = (h - n) * z + n
        hs.append(h); rs.append(r); zs.append(z); ns.append(n); ghns.append(ghn)
    st = lambda xs: torch.stack(xs, 1)
    return st(hs), st(rs), st(zs), st(ns), st(ghns)


def gru_backward(dh_out: Tensor, hs: Tensor, r: Tensor, z: Tensor, n: Tensor, ghn: Tensor,
                 w_hh: Tensor, dh_last: Optional[Tensor] = None):
    """Hand-derived BPTT.  dh_out [P,T,B,H] = dL/d(h_t output of step t).
    Returns dgi [P,T,B,G], dw_hh [P,G,H], db_hh [P,G], dh0 [P,B,H]."""
    P, T, B, H = r.shape
    dgi = torch.empty(P, T, B, 3 * H, dtype=r.dtype)
    dw_hh = torch.zeros_like(w_hh)
    db_hh = torch.zeros(P, 3 * H, dtype=r.dtype)
    dh = torch.zeros(P, B, H, dtype=r.dtype) if dh_last is None else dh_last.clone()
    for t in range(T - 1, -1, -1):
        dh = dh + dh_out[:, t]
        h_prev = hs[:, t]
        dn = dh * (1 - z[:, t])
        dz = dh * (h_prev - n[:, t])
        da_n = dn * (1 - n[:, t] * n[:, t])
        dr = da_n * ghn[:, t]
        da_r = dr * r[:, t] * (1 - r[:, t])
        da_z = dz * z[:, t] * (1 - z[:, t])
        dgi[:, t] = torch.cat([da_r, da_z, da_n], -1)
        dgh = torch.cat([da_r, da_z, da_n * r[:, t]], -1)
        dw_hh += torch.bmm(dgh.transpose(1, 2), h_prev)
        db_hh += dgh.sum(1)
        dh = dh * z[:, t] + torch.bmm(dgh, w_hh)
    return dgi, dw_hh, db_hh, dh


# ----------------------------------------------------------------------------------------------
# CRVAE forward / loss / backward  (CRVAE_lorenz96.py:203-221, :484-489, :497)
# ----------------------------------------------------------------------------------------------
def crvae_forward(prm: Params, X: Tensor, eps: Tensor) -> Dict[str, Tensor]:
    """X (B,20,p); eps (B,H) is the N(0,1) draw of :214.  Returns every activation."""
    enc_in, dec_in, target = split_window(X.to(prm["w_hh"].dtype))
    eps = eps.to(enc_in.dtype).reshape(-1, eps.shape[-1])
    # encoder GRU over the first 10 real steps, h0 = 0 (:207-208)
    gi_e = (enc_in @ prm["enc_w_ih"].t() + prm["enc_b_ih"]).unsqueeze(0)
    h0 = torch.zeros(X.shape[0], prm["enc_w_hh"].shape[1], dtype=enc_in.dtype)
    ehs, er, ez, en, eghn = gru_forward(gi_e, h0, prm["enc_w_hh"][None], prm["enc_b_hh"][None])
    hT = ehs[0, -1]
    mu = hT @ prm["mu_w"].t() + prm["mu_b"]            # :210
    log_var = hT @ prm["std_w"].t() + prm["std_b"]     # :211
    sigma = torch.exp(0.5 * log_var)                   # :213
    zlat = mu + sigma * eps                            # :216
    # decoder heads: masked-dense projection (exact zeros add exactly 0.0), h0 = z (:218)
    w_ih = prm["w_ih"] * prm["mask"][:, None, :].to(enc_in.dtype)
    gi = torch.einsum("tbk,pgk->ptbg", dec_in, w_ih) + prm["b_ih"][:, None, None, :]
    hs, r, z, n, ghn = gru_forward(gi, zlat, prm["w_hh"], prm["b_hh"])
    pred = torch.einsum("ptbh,ph->ptb", hs[:, 1:], prm["w_lin"]) + prm["b_lin"][:, None, None]  # :120
    return dict(enc_in=enc_in, dec_in=dec_in, target=target, eps=eps, ehs=ehs, er=er, ez=ez, en=en,
                eghn=eghn, hT=hT, mu=mu, log_var=log_var, sigma=sigma, zlat=zlat, gi=gi, hs=hs, r=r,
                z=z, n=n, ghn=ghn, pred=pred)


def crvae_loss(prm: Params, act: Dict[str, Tensor], lam_ridge: float, beta: float,
               head_slice: Optional[slice] = None):
    """loss = sum_i MSE(pred_i, X[:,10:,i]) (:484); ridge (:321-325, :488); 'mmd' (:486).
    NOTE the reference unpacks `pred, mu, log_var = crvae(X)` (:482) while forward returns
    `pred, log_var, mu` (:221): inside the trainer the two names are swapped, so the KL term
    actually evaluated is  mean_b sum_h -0.5*(1 + mu - log_var**2 - exp(mu)).  Reproduced here."""
    pred, target = act["pred"], act["target"]
    P = pred.shape[0]
    tgt = target.permute(2, 0, 1)                      # [p,Td,B]
    if head_slice is not None:
        tgt = tgt[head_slice]
    elif tgt.shape[0] != P:
        raise ValueError("head_slice required for a head shard")
    diff = pred - tgt
    loss = (diff * diff).mean(dim=(1, 2)).sum()
    a, s = act["mu"], act["log_var"]
    kl = (-0.5 * (1 + a - s * s - torch.exp(a))).sum(-1).mean(0)
    ridge = lam_ridge * ((prm["w_lin"] ** 2).sum() + (prm["w_hh"] ** 2).sum())
    smooth = loss + ridge + beta * kl
    return dict(loss=loss, kl=kl, ridge=ridge, smooth=smooth, diff=diff)


def crvae_backward(prm: Params, act: Dict[str, Tensor], lossd: Dict[str, Tensor],
                   lam_ridge: float, beta: float, dz_extra: Optional[Tensor] = None) -> Params:
    """Gradient of `smooth` w.r.t. every parameter (what :497 produces)."""
    P, Td, B = act["pred"].shape
    dpred = 2.0 * lossd["diff"] / (B * Td)
    dh_out = dpred[..., None] * prm["w_lin"][:, None, None, :]
    dgi, dw_hh, db_hh, dh0 = gru_backward(dh_out, act["hs"], act["r"], act["z"], act["n"],
                                          act["ghn"], prm["w_hh"])
    g: Params = {}
    g["w_ih"] = torch.einsum("ptbg,tbk->pgk", dgi, act["dec_in"]) * prm["mask"][:, None, :].to(dgi.dtype)
    g["b_ih"] = dgi.sum((1, 2))
    g["w_hh"] = dw_hh + 2 * lam_ridge * prm["w_hh"]
    g["b_hh"] = db_hh
    g["w_lin"] = torch.einsum("ptb,ptbh->ph", dpred, act["hs"][:, 1:]) + 2 * lam_ridge * prm["w_lin"]
    g["b_lin"] = dpred.sum((1, 2))
    dzlat = dh0.sum(0)                                  # every head's h0 is z (:218)
    if dz_extra is not None:
        dzlat = dzlat + dz_extra
    g["dzlat"] = dzlat
    a, s = act["mu"], act["log_var"]
    dmu = dzlat + beta * (-0.5 * (1 - torch.exp(a))) / B          # swapped-KL quirk, see crvae_loss
    dlv = dzlat * act["eps"] * 0.5 * act["sigma"] + beta * s / B
    g["mu_w"] = dmu.t() @ act["hT"]; g["mu_b"] = dmu.sum(0)
    g["std_w"] = dlv.t() @ act["hT"]; g["std_b"] = dlv.sum(0)
    dhT = dmu @ prm["mu_w"] + dlv @ prm["std_w"]
    Te = act["er"].shape[1]
    zero = torch.zeros(1, Te, B, dhT.shape[-1], dtype=dhT.dtype)
    dgi_e, dw, db, _ = gru_backward(zero, act["ehs"], act["er"], act["ez"], act["en"], act["eghn"],
                                    prm["enc_w_hh"][None], dh_last=dhT[None])
    g["enc_w_ih"] = torch.einsum("tbg,tbk->gk", dgi_e[0], act["enc_in"])
    g["enc_b_ih"] = dgi_e[0].sum((0, 1))
    g["enc_w_hh"] = dw[0]; g["enc_b_hh"] = db[0]
    return g


# ----------------------------------------------------------------------------------------------
# update rules
# ----------------------------------------------------------------------------------------------
def gd_step(prm: Params, grads: Params, lr: float) -> None:
    """param.data -= lr * param.grad for every parameter (:498-499); product rounded first."""
    for k in PARAM_KEYS:
        prm[k] -= lr * grads[k]


def prox_update(w_ih: Tensor, lam: float, lr: float) -> Tensor:
    """Group-lasso soft threshold on input columns of every head (:308-314):
    W[:,j] <- W[:,j] / max(nu, lam*lr) * max(nu - lr*lam, 0),  nu = ||W[:,j]||_2."""
    norm = torch.norm(w_ih, dim=-2, keepdim=True)
    return (w_ih / torch.clamp(norm, min=(lam * lr))) * torch.clamp(norm - (lr * lam), min=0.0)


def gc_matrix(w_ih: Tensor, threshold: bool = True) -> Tensor:
    """CRVAE.GC (:286-304): row i = column norms of head i's weight_ih_l0."""
    gc = torch.norm(w_ih, dim=-2)
    return (torch.abs(gc) > 0).int() if threshold else gc


def prox_margin(w_ih: Tensor, lam: float, lr: float, mask: Optional[Tensor] = None) -> float:
    """min_j | ||W[:,j]|| - lr*lam | / (lr*lam): how close any column is to the prox threshold."""
    nu = torch.norm(w_ih.double(), dim=-2)
    m = (nu - lr * lam).abs() / (lr * lam)
    if mask is not None:
        m = m[mask]
    return float(m.min())


def adam_step(prm: Dict[str, Tensor], grads: Dict[str, Tensor], state: Dict[str, Dict[str, Tensor]],
              step: int, lr: float = 1e-3, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam defaults (:565), single-tensor formulation; `step` counts from 1."""
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    for k, g in grads.items():
        st = state.setdefault(k, {"m": torch.zeros_like(g), "v": torch.zeros_like(g)})
        st["m"].lerp_(g, 1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        prm[k].addcdiv_(st["m"], denom, value=-(lr / bc1))


# ----------------------------------------------------------------------------------------------
# one steady-state phase-1 iteration and the free-running trainer (:457-560)
# ----------------------------------------------------------------------------------------------
def phase1_iteration(prm: Params, X: Tensor, eps: Tensor, lr: float, lam: float,
                     lam_ridge: float, beta: float = 0.1):
    """forward(eps) -> loss -> backward -> GD -> prox.  (The reference orders its loop
    backward/GD/prox/forward, :495-515; one call here = the forward that ends iteration k-1 plus
    the update that opens iteration k -- the same sequence of states.)"""
    act = crvae_forward(prm, X, eps)
    lossd = crvae_loss(prm, act, lam_ridge, beta)
    grads = crvae_backward(prm, act, lossd, lam_ridge, beta)
    gd_step(prm, grads, lr)
    if lam > 0:
        prm["w_ih"] = prox_update(prm["w_ih"], lam, lr)
    return act, lossd, grads


def draw_eps(B: int, H: int) -> Tensor:
    """The reference draws its reparameterisation noise on the CPU default generator as
    torch.randn(size=(1,B,H)) (:214-215); same call, same stream position."""
    return torch.randn(size=(1, B, H))[0]


def train_phase1(prm: Params, X_series: Tensor, context: int, lr: float, max_iter: int,
                 lam: float = 0.0, lam_ridge: float = 0.0, check_every: int = 50,
                 batch_size: int = 256, log: Optional[List[dict]] = None,
                 resume_it: Optional[int] = None, idx: Optional[np.ndarray] = None) -> Params:
    """Free-running restatement of train_phase1 (:457-560) incl. the RNG draw order (one
    np.random.randint for the fixed batch :470; one randn(1,B,H) per forward :214; in each check
    block one more forward draw :522 and one generation draw :225) and best-checkpoint selection
    (:544-547, :558).  Returns the restored (best) parameters; `prm` is updated in place."""
    beta = 0.1                                               # :475
    wins = torch.cat([arrange_input(x, context)[0] for x in X_series], 0)
    if idx is None:
        idx = np.random.randint(len(wins), size=(batch_size,))   # :470
    X = wins[idx]
    B, H = X.shape[0], prm["enc_w_hh"].shape[1]
    P = prm["w_hh"].shape[0]
    best_loss, best, best_it = np.inf, None, None
    start = 0
    if resume_it is None:
        act = crvae_forward(prm, X, draw_eps(B, H))          # :482
        lossd = crvae_loss(prm, act, lam_ridge, beta)
    else:
        start = resume_it
    for it in range(start, max_iter):
        if resume_it is None or it > resume_it:
            grads = crvae_backward(prm, act, lossd, lam_ridge, beta)   # :497
            gd_step(prm, grads, lr)                                    # :498-499
            if lam > 0:
                prm["w_ih"] = prox_update(prm["w_ih"], lam, lr)        # :502-504
        act = crvae_forward(prm, X, draw_eps(B, H))                # :508
        lossd = crvae_loss(prm, act, lam_ridge, beta)
        if it % check_every == 0:                                  # :518
            act_t = crvae_forward(prm, X, draw_eps(B, H))          # :522
            l_t = crvae_loss(prm, act_t, lam_ridge, beta)
            mean_loss = float((l_t["loss"] + l_t["ridge"]) / P)    # :530-533
            usage = float(100 * gc_matrix(prm["w_ih"]).float().mean())
            if log is not None:
                log.append(dict(it=it, mean_loss=mean_loss, kl=float(lossd["kl"]), usage=usage))
            if mean_loss < best_loss:                              # :544-547
                best_loss, best_it = mean_loss, it
                best = {k: v.clone() for k, v in prm.items()}
            draw_eps(B, H)                                         # :550 -> :225 (result unused)
    if best is not None:                                           # :558
        for k in best:
            prm[k] = best[k]
    prm["_best_it"] = torch.tensor(-1 if best_it is None else best_it)
    return prm


# ----------------------------------------------------------------------------------------------
# VRAE4E -- the error-compensation VRAE of phase 2  (CRVAE_lorenz96.py:123-179, :599-603, :639-643)
# parameters: enc_w_ih [G,p] enc_w_hh [G,H] enc_b_ih enc_b_hh [G]   (gru_left :133)
#             mu_w mu_b std_w std_b (fc_mu/fc_std :136-137)  hid_w [H,H] hid_b [H] (linear_hidden :139)
#             dec_w_ih [G,p] dec_w_hh [G,H] dec_b_ih dec_b_hh [G]   (gru :142)   out_w [p,H] out_b [p] (linear :144)
# ----------------------------------------------------------------------------------------------
VRAE_KEYS = ("enc_w_ih", "enc_w_hh", "enc_b_ih", "enc_b_hh", "mu_w", "mu_b", "std_w", "std_b", "hid_w", "hid_b",
             "dec_w_ih", "dec_w_hh", "dec_b_ih", "dec_b_hh", "out_w", "out_b")
_VRAE_SD = ("gru_left.weight_ih_l0", "gru_left.weight_hh_l0", "gru_left.bias_ih_l0", "gru_left.bias_hh_l0",
            "fc_mu.weight", "fc_mu.bias", "fc_std.weight", "fc_std.bias", "linear_hidden.weight", "linear_hidden.bias",
            "gru.weight_ih_l0", "gru.weight_hh_l0", "gru.bias_ih_l0", "gru.bias_hh_l0", "linear.weight", "linear.bias")


def vrae_params_from_state_dict(sd: Dict[str, Tensor], dtype=torch.float32) -> Params:
    return {k: sd[s].detach().to(dtype).clone().contiguous() for k, s in zip(VRAE_KEYS, _VRAE_SD)}


def vrae_state_dict_from_params(prm: Params) -> Dict[str, Tensor]:
    return {s: prm[k].clone() for k, s in zip(VRAE_KEYS, _VRAE_SD)}


def vrae_forward(prm: Params, err: Tensor, eps: Tensor) -> Dict[str, Tensor]:
    """err (B,10,p) = the detached residual (:599); eps (B,H) the draw of :161."""
    dt = prm["enc_w_hh"].dtype
    err = err.to(dt)
    B = err.shape[0]
    Hh = prm["enc_w_hh"].shape[1]
    eps = eps.to(dt).reshape(B, Hh)
    enc_in = err.transpose(0, 1).contiguous()                                             # X'[:,1:] (:155)
    dec_in = torch.cat([torch.zeros_like(err[:, :1]), err[:, :-1]], 1).transpose(0, 1).contiguous()   # X'[:,:-1] (:166)
    gi_e = (enc_in @ prm["enc_w_ih"].t() + prm["enc_b_ih"]).unsqueeze(0)
    ehs, er, ez, en, eghn = gru_forward(gi_e, torch.zeros(B, Hh, dtype=dt), prm["enc_w_hh"][None], prm["enc_b_hh"][None])
    hT = ehs[0, -1]
    mu = hT @ prm["mu_w"].t() + prm["mu_b"]                  # :157
    log_var = hT @ prm["std_w"].t() + prm["std_b"]           # :158
    sigma = torch.exp(0.5 * log_var)
    zlat = mu + sigma * eps                                  # :160-163
    zh = torch.tanh(zlat @ prm["hid_w"].t() + prm["hid_b"])  # :164
    gi_d = (dec_in @ prm["dec_w_ih"].t() + prm["dec_b_ih"]).unsqueeze(0)
    dhs, dr, dz, dn, dghn = gru_forward(gi_d, zh, prm["dec_w_hh"][None], prm["dec_b_hh"][None])
    pred = dhs[0, 1:] @ prm["out_w"].t() + prm["out_b"]      # [Td,B,p]  (:167)
    return dict(enc_in=enc_in, dec_in=dec_in, eps=eps, ehs=ehs, er=er, ez=ez, en=en, eghn=eghn, hT=hT, mu=mu,
                log_var=log_var, sigma=sigma, zlat=zlat, zh=zh, dhs=dhs, dr=dr, dz=dz, dn=dn, dghn=dghn, pred=pred,
                target=enc_in)


def vrae_loss(act: Dict[str, Tensor], beta_e: float = 1.0):
    """loss_e = MSE(pred_e, error) (:601); KL with the same swapped names as the CRVAE trainer (:600-602)."""
    diff = act["pred"] - act["target"]
    loss = (diff * diff).mean()
    a, s = act["mu"], act["log_var"]
    kl = (-0.5 * (1 + a - s * s - torch.exp(a))).sum(-1).mean(0)
    return dict(loss=loss, kl=kl, smooth=loss + beta_e * kl, diff=diff)


def vrae_backward(prm: Params, act: Dict[str, Tensor], lossd: Dict[str, Tensor], beta_e: float = 1.0) -> Params:
    Td, B, p = act["pred"].shape
    dpred = 2.0 * lossd["diff"] / (Td * B * p)
    g: Params = {}
    hs_out = act["dhs"][0, 1:]
    g["out_w"] = torch.einsum("tbp,tbh->ph", dpred, hs_out)
    g["out_b"] = dpred.sum((0, 1))
    dh_out = (dpred @ prm["out_w"])[None]
    dgi_d, dw, db, dh0 = gru_backward(dh_out, act["dhs"], act["dr"], act["dz"], act["dn"], act["dghn"], prm["dec_w_hh"][None])
    g["dec_w_ih"] = torch.einsum("tbg,tbk->gk", dgi_d[0], act["dec_in"])
    g["dec_b_ih"] = dgi_d[0].sum((0, 1))
    g["dec_w_hh"], g["dec_b_hh"] = dw[0], db[0]
    dpre = dh0[0] * (1 - act["zh"] * act["zh"])
    g["hid_w"] = dpre.t() @ act["zlat"]
    g["hid_b"] = dpre.sum(0)
    dzlat = dpre @ prm["hid_w"]
    a, s = act["mu"], act["log_var"]
    dmu = dzlat + beta_e * (-0.5 * (1 - torch.exp(a))) / B
    dlv = dzlat * act["eps"] * 0.5 * act["sigma"] + beta_e * s / B
    g["mu_w"] = dmu.t() @ act["hT"]; g["mu_b"] = dmu.sum(0)
    g["std_w"] = dlv.t() @ act["hT"]; g["std_b"] = dlv.sum(0)
    dhT = dmu @ prm["mu_w"] + dlv @ prm["std_w"]
    Te = act["er"].shape[1]
    zero = torch.zeros(1, Te, B, dhT.shape[-1], dtype=dhT.dtype)
    dgi_e, dw, db, _ = gru_backward(zero, act["ehs"], act["er"], act["ez"], act["en"], act["eghn"], prm["enc_w_hh"][None],
                                    dh_last=dhT[None])
    g["enc_w_ih"] = torch.einsum("tbg,tbk->gk", dgi_e[0], act["enc_in"])
    g["enc_b_ih"] = dgi_e[0].sum((0, 1))
    g["enc_w_hh"], g["enc_b_hh"] = dw[0], db[0]
    return g


def crvae_error(act: Dict[str, Tensor]) -> Tensor:
    """error = X[:,10:,:] - stack(pred)[...,0].permute(1,2,0)  (B,10,p), detached (:599/:639)."""
    return (act["target"] - act["pred"].permute(1, 2, 0)).permute(1, 0, 2).contiguous()


def phase2_iteration(prm: Params, vprm: Params, adam_state: dict, step: int, X: Tensor, eps_c: Tensor, eps_e: Tensor,
                     lr: float, lam: float = 0.0, lam_ridge: float = 0.0):
    """forward both models -> losses -> Adam on the VRAE (:611-614) -> GD (+prox) on the CRVAE (:616-623).
    beta = beta_e = 1 (:582-583).  `step` counts Adam steps from 1."""
    act = crvae_forward(prm, X, eps_c)
    lossd = crvae_loss(prm, act, lam_ridge, 1.0)
    err = crvae_error(act)
    vact = vrae_forward(vprm, err, eps_e)
    vloss = vrae_loss(vact, 1.0)
    vgrads = vrae_backward(vprm, vact, vloss, 1.0)
    if lam == 0:
        adam_step(vprm, vgrads, adam_state, step)
    grads = crvae_backward(prm, act, lossd, lam_ridge, 1.0)
    gd_step(prm, grads, lr)
    if lam > 0:
        prm["w_ih"] = prox_update(prm["w_ih"], lam, lr)
    return dict(act=act, lossd=lossd, grads=grads, err=err, vact=vact, vloss=vloss, vgrads=vgrads)


# ----------------------------------------------------------------------------------------------
# CS-RAE variant: Cauchy-Schwarz divergence to a learnable GMM prior  (CR-CS-RAE.py:107-163, :568-582)
# ----------------------------------------------------------------------------------------------
def gaussian_overlap(mu1: Tensor, var1: Tensor, mu2: Tensor, var2: Tensor) -> Tensor:
    """N(mu1 | mu2, var1 + var2) for diagonal covariances, evaluated as exp(log-density) (CR-CS-RAE.py:124-134)."""
    var_sum = var1 + var2
    diff = mu1 - mu2
    D = mu1.size(-1)
    log_norm = -0.5 * D * math.log(2 * math.pi) - 0.5 * var_sum.log().sum(dim=-1)
    log_exp = -0.5 * (diff.pow(2) / var_sum).sum(dim=-1)
    return (log_norm + log_exp).exp()


def cs_divergence_gmm(mu_q: Tensor, var_q: Tensor, mu_p: Tensor, var_p: Tensor) -> Tensor:
    """D_CS(q || p), Gaussian q vs equal-weight GMM p: exp -> mean -> log, clamp(min=0) (CR-CS-RAE.py:137-163)."""
    D = mu_q.size(-1)
    term1 = gaussian_overlap(mu_q.unsqueeze(1), var_q.unsqueeze(1), mu_p.unsqueeze(0), var_p.unsqueeze(0)).mean(dim=1)
    term2 = gaussian_overlap(mu_p.unsqueeze(1), var_p.unsqueeze(1), mu_p.unsqueeze(0), var_p.unsqueeze(0)).mean()
    term3 = (-0.5 * D * math.log(2 * math.pi) - 0.5 * (2 * var_q).log().sum(dim=-1)).exp()
    return (-term1.log() + 0.5 * term2.log() + 0.5 * term3.log()).clamp(min=0)


def cs_head(lat: Tensor, prior_mu: Tensor, prior_logvar: Tensor, lambda_cs: float):
    """The trainer's CS term (:568-582) on lat = [fc_mu out | fc_std out] with its swapped unpacking
    (mu_q = fc_std out, var_q = exp(fc_mu out)).  Returns mean D_CS and the gradients of lambda_cs * mean
    w.r.t. lat, prior_mu, prior_logvar.  (Gradients through torch autograd of the restated formula: this
    small head is exactly what the reference differentiates.)"""
    Hh = lat.shape[1] // 2
    lat_ = lat.detach().clone().requires_grad_(True)
    pm = prior_mu.detach().clone().requires_grad_(True)
    pl = prior_logvar.detach().clone().requires_grad_(True)
    mu_q, var_q = lat_[:, Hh:], torch.exp(lat_[:, :Hh])
    cs = cs_divergence_gmm(mu_q, var_q, pm, pl.exp()).mean()
    (lambda_cs * cs).backward()
    z = lambda t: torch.zeros_like(t) if t.grad is None else t.grad
    return cs.detach(), z(lat_), z(pm), z(pl)


# ----------------------------------------------------------------------------------------------
# Generic VRAE of VRAE.py (config 4): Encoder :11-36, Decoder :38-102 (GRU variant, teacher forcing 1.0),
# VRAE.loss :142-147.  Parameters: enc_w_ih [G,D] enc_w_hh [G,H] enc_b_ih enc_b_hh [G] (encoder.rnn),
# mu_w [Z,H] mu_b [Z] lv_w [Z,H] lv_b [Z] (fc_mu / fc_logvar), z2h_w [H,Z] z2h_b [H] (fc_z2h),
# dec_w_ih [G,D] dec_w_hh [G,H] dec_b_ih dec_b_hh [G] (decoder.cell, an nn.GRUCell), out_w [D,H] out_b [D] (fc_out)
# ----------------------------------------------------------------------------------------------
GVRAE_KEYS = ("enc_w_ih", "enc_w_hh", "enc_b_ih", "enc_b_hh", "mu_w", "mu_b", "lv_w", "lv_b", "z2h_w", "z2h_b",
              "dec_w_ih", "dec_w_hh", "dec_b_ih", "dec_b_hh", "out_w", "out_b")
_GVRAE_SD = ("encoder.rnn.weight_ih_l0", "encoder.rnn.weight_hh_l0", "encoder.rnn.bias_ih_l0", "encoder.rnn.bias_hh_l0",
             "encoder.fc_mu.weight", "encoder.fc_mu.bias", "encoder.fc_logvar.weight", "encoder.fc_logvar.bias",
             "decoder.fc_z2h.weight", "decoder.fc_z2h.bias", "decoder.cell.weight_ih", "decoder.cell.weight_hh",
             "decoder.cell.bias_ih", "decoder.cell.bias_hh", "decoder.fc_out.weight", "decoder.fc_out.bias")


def gvrae_params_from_state_dict(sd: Dict[str, Tensor], dtype=torch.float32) -> Params:
    return {k: sd[s].detach().to(dtype).clone().contiguous() for k, s in zip(GVRAE_KEYS, _GVRAE_SD)}


def _act(x: Tensor, kind: str) -> Tensor:
    return {"sigmoid": torch.sigmoid, "tanh": torch.tanh, "relu": torch.relu}.get(kind, lambda t: t)(x)


def _act_grad(y: Tensor, kind: str) -> Tensor:
    if kind == "sigmoid":
        return y * (1 - y)
    if kind == "tanh":
        return 1 - y * y
    if kind == "relu":
        return (y > 0).to(y.dtype)
    return torch.ones_like(y)


def gvrae_forward(prm: Params, x: Tensor, eps: Tensor, act: str = "tanh") -> Dict[str, Tensor]:
    """x (B,T,D); eps (B,Z) the randn_like draw of VRAE.reparameterize (:117-121).  Teacher forcing 1.0:
    the decoder cell's input at step t is x[:, t] (:82-97)."""
    dt = prm["enc_w_hh"].dtype
    x = x.to(dt)
    B, T, D = x.shape
    Hh = prm["enc_w_hh"].shape[1]
    xin = x.transpose(0, 1).contiguous()                                                   # [T,B,D]
    gi_e = (xin @ prm["enc_w_ih"].t() + prm["enc_b_ih"]).unsqueeze(0)
    ehs, er, ez, en, eghn = gru_forward(gi_e, torch.zeros(B, Hh, dtype=dt), prm["enc_w_hh"][None], prm["enc_b_hh"][None])
    hT = ehs[0, -1]
    mu = hT @ prm["mu_w"].t() + prm["mu_b"]
    logvar = hT @ prm["lv_w"].t() + prm["lv_b"]
    std = torch.exp(0.5 * logvar)
    z = mu + eps.to(dt) * std
    h0 = torch.tanh(z @ prm["z2h_w"].t() + prm["z2h_b"])                                   # :72
    gi_d = (xin @ prm["dec_w_ih"].t() + prm["dec_b_ih"]).unsqueeze(0)
    dhs, dr, dz, dn, dghn = gru_forward(gi_d, h0, prm["dec_w_hh"][None], prm["dec_b_hh"][None])
    recon = _act(dhs[0, 1:] @ prm["out_w"].t() + prm["out_b"], act)                        # [T,B,D]  (:91)
    return dict(xin=xin, eps=eps.to(dt), ehs=ehs, er=er, ez=ez, en=en, eghn=eghn, hT=hT, mu=mu, logvar=logvar, std=std, z=z,
                h0=h0, dhs=dhs, dr=dr, dz=dz, dn=dn, dghn=dghn, recon=recon)


def gvrae_loss(a: Dict[str, Tensor], beta: float = 1.0):
    """VRAE.loss (:142-147): SSE / batch + beta * KLD / batch (standard KL, no name swap here)."""
    B = a["mu"].shape[0]
    diff = a["recon"] - a["xin"]
    rec = (diff * diff).sum() / B
    kld = -0.5 * torch.sum(1 + a["logvar"] - a["mu"].pow(2) - a["logvar"].exp()) / B
    return dict(total=rec + beta * kld, rec=rec, kld=kld, diff=diff)


def gvrae_backward(prm: Params, a: Dict[str, Tensor], l: Dict[str, Tensor], beta: float = 1.0, act: str = "tanh") -> Params:
    B = a["mu"].shape[0]
    drecon = 2.0 * l["diff"] / B
    dpre = drecon * _act_grad(a["recon"], act)
    g: Params = {}
    hs_out = a["dhs"][0, 1:]
    g["out_w"] = torch.einsum("tbd,tbh->dh", dpre, hs_out)
    g["out_b"] = dpre.sum((0, 1))
    dh_out = (dpre @ prm["out_w"])[None]
    dgi_d, dw, db, dh0 = gru_backward(dh_out, a["dhs"], a["dr"], a["dz"], a["dn"], a["dghn"], prm["dec_w_hh"][None])
    g["dec_w_ih"] = torch.einsum("tbg,tbk->gk", dgi_d[0], a["xin"])
    g["dec_b_ih"] = dgi_d[0].sum((0, 1))
    g["dec_w_hh"], g["dec_b_hh"] = dw[0], db[0]
    dpre0 = dh0[0] * (1 - a["h0"] * a["h0"])
    g["z2h_w"] = dpre0.t() @ a["z"]
    g["z2h_b"] = dpre0.sum(0)
    dzl = dpre0 @ prm["z2h_w"]
    dmu = dzl + beta * a["mu"] / B
    dlv = dzl * a["eps"] * 0.5 * a["std"] + beta * (-0.5 * (1 - torch.exp(a["logvar"]))) / B
    g["mu_w"] = dmu.t() @ a["hT"]; g["mu_b"] = dmu.sum(0)
    g["lv_w"] = dlv.t() @ a["hT"]; g["lv_b"] = dlv.sum(0)
    dhT = dmu @ prm["mu_w"] + dlv @ prm["lv_w"]
    T = a["er"].shape[1]
    zero = torch.zeros(1, T, B, dhT.shape[-1], dtype=dhT.dtype)
    dgi_e, dw, db, _ = gru_backward(zero, a["ehs"], a["er"], a["ez"], a["en"], a["eghn"], prm["enc_w_hh"][None], dh_last=dhT[None])
    g["enc_w_ih"] = torch.einsum("tbg,tbk->gk", dgi_e[0], a["xin"])
    g["enc_b_ih"] = dgi_e[0].sum((0, 1))
    g["enc_w_hh"], g["enc_b_hh"] = dw[0], db[0]
    return g
