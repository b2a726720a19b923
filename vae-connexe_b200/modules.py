"""Drop-in mirrors of the reference's model classes (CRVAE_lorenz96.py:97-304) on top of the fused
engine: same constructor signatures, attribute names, parameter order / state_dict keys, forward
return conventions and GC() readout -- the storage is the engine's fused arenas and every number is
produced by libcrvae_b200.so kernels.
"""
from __future__ import annotations

import copy
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import lib as L
from .engine import CRVAEEngine, H as _H, G as _G, ENC_STEPS, DEC_STEPS
from .sharding import head_range


def _require_hidden(hidden: int):
    if int(hidden) != _H:
        raise ValueError(f"vae-connexe_b200 kernels are built for hidden={_H} (CRVAE_lorenz96.py:768); got {hidden}")


def _view_param(t: torch.Tensor, g: torch.Tensor) -> nn.Parameter:
    p = nn.Parameter(t, requires_grad=True)      # shares storage with the arena view
    p.grad = g
    return p


class _GRUParams(nn.Module):
    """Attribute surface of nn.GRU that the reference touches (:104-105, :310, :325)."""

    def __init__(self, w_ih, w_hh, b_ih, b_hh, gw_ih, gw_hh, gb_ih, gb_hh):
        super().__init__()
        self.weight_ih_l0 = _view_param(w_ih, gw_ih)
        self.weight_hh_l0 = _view_param(w_hh, gw_hh)
        self.bias_ih_l0 = _view_param(b_ih, gb_ih)
        self.bias_hh_l0 = _view_param(b_hh, gb_hh)

    def flatten_parameters(self):   # cuDNN weight-packing hint in the reference; nothing to do here
        return None


class _LinearParams(nn.Module):
    def __init__(self, w, b, gw, gb):
        super().__init__()
        self.weight = _view_param(w, gw)
        self.bias = _view_param(b, gb)


class GRU(nn.Module):
    """One decoder head (reference class GRU, :97-121) as a *view* into the fused engine.
    For a pruned head the reference stores a packed (3H, k_i) weight_ih_l0: with the engine's gather-packed storage the
    view IS that (3H, k_i) matrix; with masked-dense storage it is the (3H, p) row block with structural zeros."""

    def __init__(self, owner: "CRVAE", local_idx: int):
        super().__init__()
        eng = owner.engine
        th, g = eng.theta, eng.grad
        i = local_idx
        self.p = int(eng.mask_np[i].sum())
        self.hidden = _H
        w_ih, gw_ih = (th["w_ih"][i][:, :self.p], g["w_ih"][i][:, :self.p]) if eng.packed else (th["w_ih"][i], g["w_ih"][i])
        self.gru = _GRUParams(w_ih, th["w_hh"][i], th["b_ih"][i], th["b_hh"][i],
                              gw_ih, g["w_hh"][i], g["b_ih"][i], g["b_hh"][i])
        self.linear = _LinearParams(th["w_lin"][i:i + 1], th["b_lin"][i:i + 1], g["w_lin"][i:i + 1], g["b_lin"][i:i + 1])
        self._owner = [owner]          # list: keep nn.Module from registering the parent as a child
        self._local_idx = i

    def init_hidden(self, batch):
        return torch.zeros(1, batch, self.hidden, device=self.gru.weight_ih_l0.device)


class _TrainFn(torch.autograd.Function):
    """Custom autograd bridge: forward runs the fused kernels; backward receives d(pred),
    d(log_var), d(mu) from whatever loss the caller built in torch, runs the hand-written BPTT and
    leaves the parameter gradients in the engine's grad arena (== every Parameter's .grad)."""

    @staticmethod
    def forward(ctx, anchor, owner, eps_dev):
        eng = owner.engine
        eng.forward(eps_dev)
        owner._fwd_serial += 1
        owner._bwd_ran = False
        ctx.owner, ctx.serial, ctx.eps = owner, owner._fwd_serial, eng.eps.clone()
        ctx.bound_X, ctx.bind_serial = owner._bound_X, eng.bind_serial     # the batch these activations belong to
        B = eng.B
        lat = eng.lat.clone()
        return eng.pred[:eng.P].clone(), lat[:, _H:].reshape(1, B, _H), lat[:, :_H].reshape(1, B, _H)

    @staticmethod
    def backward(ctx, dpred, dlog_var, dmu):
        owner = ctx.owner
        eng = owner.engine
        if owner._fwd_serial != ctx.serial or eng.bind_serial != ctx.bind_serial:
            # activations were overwritten by a later forward and / or another batch was bound since: recompute on
            # the batch and the noise this graph node was built from
            if eng.bind_serial != ctx.bind_serial:
                owner._bind(ctx.bound_X)
            eng.forward(ctx.eps)
            owner._fwd_serial += 1
        B = eng.B
        if dpred is not None:
            eng.dpred[:eng.P].copy_(dpred)
        else:
            eng.dpred.zero_()
        extra = torch.zeros(B, 2 * _H, dtype=torch.float32, device=eng.device)
        if dmu is not None:
            extra[:, :_H] = dmu.reshape(B, _H)
        if dlog_var is not None:
            extra[:, _H:] = dlog_var.reshape(B, _H)
        eng.backward(beta=0.0, lam_ridge=0.0, dlat_extra=extra)
        owner._sync_ragged_grads()
        from .functional import apply_ridge_grad
        for i, alpha in owner._ridge_pending:          # ridge_regularize nodes that ran before this one
            apply_ridge_grad(eng, i, alpha)
        owner._ridge_pending.clear()
        owner._bwd_ran = True
        return None, None, None


class CRVAE(nn.Module):
    """Mirror of reference class CRVAE (:181-304): CRVAE(num_series, connection, hidden).

    rank / world_size / group select a head shard (heads [lo, hi) of num_series live on this rank;
    the encoder is replicated).  Initialisation draws from the global CPU generator in the
    reference's declaration order (gru_left, fc_mu, fc_std, then (GRU, Linear) per head) by
    constructing the same torch modules, so torch.manual_seed(s) gives the reference's weights."""

    def __init__(self, num_series, connection, hidden, rank: int = 0, world_size: int = 1, group=None,
                 device: Optional[str] = None, _init: bool = True, comm=None, packed=None):
        super().__init__()
        _require_hidden(hidden)
        kern = L.kernels()                 # raises without libcrvae_b200.so + a B200: there is no CPU path
        if device is not None:
            self.device = torch.device(device)
        elif kern.device_type == "cuda":
            self.device = torch.device("cuda", torch.cuda.current_device())
        else:
            self.device = torch.device(kern.device_type)
        self.p = int(num_series)
        self.hidden = int(hidden)
        self.connection = connection
        conn = np.asarray(connection)
        assert conn.shape == (self.p, self.p)
        self.rank, self.world_size, self.group = rank, world_size, group
        lo, hi = head_range(self.p, rank, world_size)
        self.head_lo, self.head_hi = lo, hi
        full_mask = (conn != 0).T                       # head i reads column j iff connection[j, i] != 0 (:115, :201)
        # packed: None = automatic (gather-packed storage when every head reads <= p/4 series), True / False = forced
        self.engine = CRVAEEngine(self.p, full_mask[lo:hi], head_off=lo, device=self.device,
                                  group=group if world_size > 1 else None, comm=comm if world_size > 1 else None, packed=packed)
        if _init:
            self._init_like_reference(full_mask, lo, hi)
        eng = self.engine
        th, g = eng.theta, eng.grad
        self.gru_left = _GRUParams(th["enc_w_ih"], th["enc_w_hh"], th["enc_b_ih"], th["enc_b_hh"],
                                   g["enc_w_ih"], g["enc_w_hh"], g["enc_b_ih"], g["enc_b_hh"])
        self.fc_mu = _LinearParams(th["lat_w"][:_H], th["lat_b"][:_H], g["lat_w"][:_H], g["lat_b"][:_H])
        self.fc_std = _LinearParams(th["lat_w"][_H:], th["lat_b"][_H:], g["lat_w"][_H:], g["lat_b"][_H:])
        self._register_extra()    # hook: modules a variant declares between fc_std and the heads (parameters() order)
        self.networks = nn.ModuleList([GRU(self, i) for i in range(hi - lo)])
        self._anchor = torch.zeros(1, device=self.device, requires_grad=True)
        self._fwd_serial = 0
        self._bound_X = None
        self._pinned = None
        self._bwd_ran = False
        self._ridge_pending = []

    # ------------------------------------------------------------------ init / state
    def _init_like_reference(self, full_mask, lo, hi):
        eng = self.engine
        th = eng.theta
        enc = nn.GRU(self.p, _H, batch_first=True)                        # :192
        fc_mu, fc_std = nn.Linear(_H, _H), nn.Linear(_H, _H)              # :195-196
        self._init_extra()        # hook: variants that declare more parameters before the heads (CR-CS-RAE's prior)
        with torch.no_grad():
            th["enc_w_ih"].copy_(enc.weight_ih_l0); th["enc_w_hh"].copy_(enc.weight_hh_l0)
            th["enc_b_ih"].copy_(enc.bias_ih_l0); th["enc_b_hh"].copy_(enc.bias_hh_l0)
            th["lat_w"][:_H].copy_(fc_mu.weight); th["lat_w"][_H:].copy_(fc_std.weight)
            th["lat_b"][:_H].copy_(fc_mu.bias); th["lat_b"][_H:].copy_(fc_std.bias)
            P = hi - lo
            w_ih = torch.zeros(P, _G, eng.Kw)
            w_hh, b_ih, b_hh = torch.zeros(P, _G, _H), torch.zeros(P, _G), torch.zeros(P, _G)
            w_lin, b_lin = torch.zeros(P, _H), torch.zeros(P)
            for i in range(self.p):                                       # :200-201, every head draws in order
                cols = np.where(full_mask[i])[0]
                gru = nn.GRU(len(cols), _H, batch_first=True)             # :104
                lin = nn.Linear(_H, 1)                                    # :106
                if lo <= i < hi:
                    j = i - lo
                    if eng.packed:
                        w_ih[j][:, :len(cols)] = gru.weight_ih_l0
                    else:
                        w_ih[j][:, cols] = gru.weight_ih_l0
                    w_hh[j], b_ih[j], b_hh[j] = gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0
                    w_lin[j], b_lin[j] = lin.weight[0], lin.bias[0]
            for name, val in (("w_ih", w_ih), ("w_hh", w_hh), ("b_ih", b_ih), ("b_hh", b_hh), ("w_lin", w_lin),
                              ("b_lin", b_lin)):
                th[name].copy_(val)

    def _init_extra(self):
        return None

    def _register_extra(self):
        return None

    def _clone_args(self):
        return (self.p, copy.deepcopy(self.connection), self.hidden)

    def _copy_extra_to(self, new):
        return None

    def state_dict(self, *args, **kwargs):
        """Reference-shaped state_dict (keys of SURVEY 8(a3); pruned heads packed to (3H, k_i));
        head indices are global."""
        eng, th = self.engine, self.engine.theta
        sd = {"gru_left.weight_ih_l0": th["enc_w_ih"], "gru_left.weight_hh_l0": th["enc_w_hh"],
              "gru_left.bias_ih_l0": th["enc_b_ih"], "gru_left.bias_hh_l0": th["enc_b_hh"],
              "fc_mu.weight": th["lat_w"][:_H], "fc_mu.bias": th["lat_b"][:_H],
              "fc_std.weight": th["lat_w"][_H:], "fc_std.bias": th["lat_b"][_H:]}
        for j in range(eng.P):
            i = self.head_lo + j
            cols = torch.from_numpy(np.where(eng.mask_np[j])[0]).to(self.device)
            sd[f"networks.{i}.gru.weight_ih_l0"] = th["w_ih"][j][:, :len(cols)] if eng.packed else th["w_ih"][j].index_select(1, cols)
            sd[f"networks.{i}.gru.weight_hh_l0"] = th["w_hh"][j]
            sd[f"networks.{i}.gru.bias_ih_l0"] = th["b_ih"][j]
            sd[f"networks.{i}.gru.bias_hh_l0"] = th["b_hh"][j]
            sd[f"networks.{i}.linear.weight"] = th["w_lin"][j:j + 1]
            sd[f"networks.{i}.linear.bias"] = th["b_lin"][j:j + 1]
        return {k: v.detach().clone() for k, v in sd.items()}

    def load_state_dict(self, sd, strict: bool = True):
        eng, th = self.engine, self.engine.theta
        with torch.no_grad():
            th["enc_w_ih"].copy_(sd["gru_left.weight_ih_l0"]); th["enc_w_hh"].copy_(sd["gru_left.weight_hh_l0"])
            th["enc_b_ih"].copy_(sd["gru_left.bias_ih_l0"]); th["enc_b_hh"].copy_(sd["gru_left.bias_hh_l0"])
            th["lat_w"][:_H].copy_(sd["fc_mu.weight"]); th["lat_w"][_H:].copy_(sd["fc_std.weight"])
            th["lat_b"][:_H].copy_(sd["fc_mu.bias"]); th["lat_b"][_H:].copy_(sd["fc_std.bias"])
            th["w_ih"].zero_()
            for j in range(eng.P):
                i = self.head_lo + j
                cols = torch.from_numpy(np.where(eng.mask_np[j])[0]).to(self.device)
                if eng.packed:
                    th["w_ih"][j][:, :len(cols)].copy_(sd[f"networks.{i}.gru.weight_ih_l0"])
                else:
                    th["w_ih"][j].index_copy_(1, cols, sd[f"networks.{i}.gru.weight_ih_l0"].to(self.device))
                th["w_hh"][j].copy_(sd[f"networks.{i}.gru.weight_hh_l0"])
                th["b_ih"][j].copy_(sd[f"networks.{i}.gru.bias_ih_l0"])
                th["b_hh"][j].copy_(sd[f"networks.{i}.gru.bias_hh_l0"])
                th["w_lin"][j].copy_(sd[f"networks.{i}.linear.weight"][0])
                th["b_lin"][j].copy_(sd[f"networks.{i}.linear.bias"][0])

    def zero_grad(self, set_to_none: bool = False):
        self.engine.zero_grad()            # .grad tensors are views of the grad arena: keep them
        self._bwd_ran = False
        self._ridge_pending.clear()

    def close(self):
        """Drop everything that holds communicator state: the generator's captured CUDA graphs (they contain the per-step
        all-gather on a head shard) and the dz communicator.  Call before torch.distributed.destroy_process_group():
        tearing the communicator down while a captured graph still references its kernels hangs."""
        plans = self.__dict__.get("_gen_plans")
        if plans:
            for plan in plans.values():
                plan.graph = None
            plans.clear()
        self.engine.close()

    def _sync_ragged_grads(self):
        return None                        # masked-dense: gradients of structural zeros are already masked

    def to(self, *args, **kwargs):
        dev = args[0] if args else kwargs.get("device")
        if dev is not None and torch.device(dev).type != self.device.type:
            raise L.CrvaeLibraryError("vae-connexe_b200 modules live on the GPU only")
        return self

    def __deepcopy__(self, memo):
        """deepcopy(crvae) is how the reference snapshots its best model (:547): a device-side copy
        of the fused parameter arena into a fresh engine."""
        new = type(self)(*self._clone_args(), rank=self.rank, world_size=self.world_size, group=self.group,
                         device=str(self.device), _init=False, comm=self.engine.comm, packed=self.engine.packed)   # no init draws: the caller's generator is untouched
        new.engine.theta.flat.copy_(self.engine.theta.flat)
        new.engine.grad.flat.copy_(self.engine.grad.flat)
        self._copy_extra_to(new)
        return new

    # ------------------------------------------------------------------ forward
    def _bind(self, X: torch.Tensor):
        """Always re-bind (one small kernel).  A cache keyed on the tensor's address would be wrong: the caching allocator
        hands a freshly indexed temporary (crvae(X_all[idx])) the address of the previous one, and the trainers re-bind
        the engine behind the module's back."""
        self.engine.bind_batch(X)
        self._bound_X = X

    def _draw_eps(self, B: int) -> torch.Tensor:
        """torch.randn(size=mu.size()) on the CPU default generator (:214), then to the device (:215)."""
        eps = torch.randn(size=(1, B, _H))
        if self.device.type != "cuda":
            return eps
        if self._pinned is None or self._pinned.shape != eps.shape:
            self._pinned = torch.empty_like(eps).pin_memory()
        torch.cuda.current_stream().synchronize()   # the staging buffer may still feed the previous upload
        self._pinned.copy_(eps)
        return self._pinned.to(self.device, non_blocking=True)

    def forward(self, X, noise=None, mode="train", phase=0):
        if mode == "train":                       # phase 0 and 1 are identical in train mode (:206-221, :247-262)
            self._bind(X)
            eps = self._draw_eps(X.shape[0])
            pred, log_var, mu = _TrainFn.apply(self._anchor, self, eps[0])
            # reference returns a python list of p tensors (B, 10, 1) (:218-221)
            pred_list = list(pred.permute(0, 2, 1).unsqueeze(-1).unbind(0))
            return pred_list, log_var, mu
        if mode == "test":
            from .generate import crvae_generate
            return crvae_generate(self, X, noise, phase)
        raise ValueError(mode)

    # ------------------------------------------------------------------ GC readout (:286-304)
    def GC(self, threshold=True):
        """(p x p): entry (i, j) = whether series j Granger-causes series i = ||weight_ih_l0[i][:, j]||_2 > 0.
        On a head shard the rows are all-gathered so every rank returns the full matrix."""
        norms = self.engine.column_norms()[: self.engine.P].clone()
        if self.world_size > 1:
            from .sharding import allgather_rows
            norms = allgather_rows(norms, self.p, self.rank, self.world_size, self.group)
        if threshold:
            return (torch.abs(norms) > 0).int()
        return norms


class _VraeTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, owner, eps_dev):
        eng = owner.engine
        eng.forward(eps_dev)
        owner._fwd_serial += 1
        ctx.owner, ctx.serial, ctx.eps = owner, owner._fwd_serial, eng.eps.clone()
        ctx.err = eng.enc_in.clone()
        B = eng.B
        lat = eng.lat.clone()
        return eng.pred.clone(), lat[:, _H:].reshape(1, B, _H), lat[:, :_H].reshape(1, B, _H)

    @staticmethod
    def backward(ctx, dpred, dlog_var, dmu):
        owner = ctx.owner
        eng = owner.engine
        if owner._fwd_serial != ctx.serial:
            eng.bind_error(ctx.err)
            eng.forward(ctx.eps)
            owner._fwd_serial += 1
        B = eng.B
        if dpred is not None:
            eng.dpred.copy_(dpred)
        else:
            eng.dpred.zero_()
        extra = torch.zeros(B, 2 * _H, dtype=torch.float32, device=eng.device)
        if dmu is not None:
            extra[:, :_H] = dmu.reshape(B, _H)
        if dlog_var is not None:
            extra[:, _H:] = dlog_var.reshape(B, _H)
        eng.backward(beta_e=0.0, dlat_extra=extra)
        return None, None, None


class VRAE4E(nn.Module):
    """Mirror of reference class VRAE4E (:123-179), the error-compensation VRAE of phase 2:
    VRAE4E(num_series, hidden); forward(X, mode) -> (pred, log_var, mu) in train mode (:169)."""

    def __init__(self, num_series, hidden, device: Optional[str] = None, _init: bool = True):
        super().__init__()
        _require_hidden(hidden)
        from .vrae_engine import VRAE4EEngine
        kern = L.kernels()
        if device is not None:
            self.device = torch.device(device)
        elif kern.device_type == "cuda":
            self.device = torch.device("cuda", torch.cuda.current_device())
        else:
            self.device = torch.device(kern.device_type)
        self.p, self.hidden = int(num_series), int(hidden)
        self.engine = VRAE4EEngine(self.p, self.device)
        th, g = self.engine.theta, self.engine.grad
        if _init:
            # declaration order of the reference (:133-144): gru_left, fc_mu, fc_std, linear_hidden, gru, linear
            enc = nn.GRU(self.p, _H, batch_first=True)
            fc_mu, fc_std, hid = nn.Linear(_H, _H), nn.Linear(_H, _H), nn.Linear(_H, _H)
            dec = nn.GRU(self.p, _H, batch_first=True)
            out = nn.Linear(_H, self.p)
            with torch.no_grad():
                th["enc_w_ih"].copy_(enc.weight_ih_l0); th["enc_w_hh"].copy_(enc.weight_hh_l0)
                th["enc_b_ih"].copy_(enc.bias_ih_l0); th["enc_b_hh"].copy_(enc.bias_hh_l0)
                th["lat_w"][:_H].copy_(fc_mu.weight); th["lat_w"][_H:].copy_(fc_std.weight)
                th["lat_b"][:_H].copy_(fc_mu.bias); th["lat_b"][_H:].copy_(fc_std.bias)
                th["hid_w"].copy_(hid.weight); th["hid_b"].copy_(hid.bias)
                th["dec_w_ih"].copy_(dec.weight_ih_l0); th["dec_w_hh"].copy_(dec.weight_hh_l0)
                th["dec_b_ih"].copy_(dec.bias_ih_l0); th["dec_b_hh"].copy_(dec.bias_hh_l0)
                th["out_w"].copy_(out.weight); th["out_b"].copy_(out.bias)
        self.gru_left = _GRUParams(th["enc_w_ih"], th["enc_w_hh"], th["enc_b_ih"], th["enc_b_hh"],
                                   g["enc_w_ih"], g["enc_w_hh"], g["enc_b_ih"], g["enc_b_hh"])
        self.fc_mu = _LinearParams(th["lat_w"][:_H], th["lat_b"][:_H], g["lat_w"][:_H], g["lat_b"][:_H])
        self.fc_std = _LinearParams(th["lat_w"][_H:], th["lat_b"][_H:], g["lat_w"][_H:], g["lat_b"][_H:])
        self.linear_hidden = _LinearParams(th["hid_w"], th["hid_b"], g["hid_w"], g["hid_b"])
        self.gru = _GRUParams(th["dec_w_ih"], th["dec_w_hh"], th["dec_b_ih"], th["dec_b_hh"],
                              g["dec_w_ih"], g["dec_w_hh"], g["dec_b_ih"], g["dec_b_hh"])
        self.linear = _LinearParams(th["out_w"], th["out_b"], g["out_w"], g["out_b"])
        self._anchor = torch.zeros(1, device=self.device, requires_grad=True)
        self._fwd_serial = 0
        self._pinned = None

    def init_hidden(self, batch):
        return torch.zeros(1, batch, self.hidden, device=self.device)

    def zero_grad(self, set_to_none: bool = False):
        self.engine.zero_grad()

    def to(self, *args, **kwargs):
        return self

    def __deepcopy__(self, memo):
        new = VRAE4E(self.p, self.hidden, device=str(self.device), _init=False)
        new.engine.theta.flat.copy_(self.engine.theta.flat)
        new.engine.exp_avg.copy_(self.engine.exp_avg); new.engine.exp_avg_sq.copy_(self.engine.exp_avg_sq)
        new.engine.adam_counter.copy_(self.engine.adam_counter)
        return new

    _draw_eps = CRVAE._draw_eps

    def forward(self, X, mode="train"):
        if mode == "train":
            self.engine.bind_error(X.transpose(0, 1))          # (B,10,p) -> time-major
            eps = self._draw_eps(X.shape[0])
            pred, log_var, mu = _VraeTrainFn.apply(self._anchor, self, eps[0])
            return pred.permute(1, 0, 2), log_var, mu            # (B,10,p), (1,B,H), (1,B,H)  (:169)
        if mode == "test":
            from .generate import vrae_generate
            return vrae_generate(self, X)
        raise ValueError(mode)
