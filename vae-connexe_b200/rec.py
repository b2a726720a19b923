"""Choice of the recurrent kernel for a (heads, batch) shape.

Three implementations of the same GRU recurrence / BPTT live in the library (include/crvae_b200.h):
  * tcgen05  (crvae_gru_fwd_tc / crvae_gru_bwd_tc)   128-row tiles, throughput: many heads per GPU;
  * low-latency (crvae_gru_fwd_ll / crvae_gru_bwd_ll) 16-row tiles, bulk-copy slab ring: few heads per GPU, the
    replicated encoder, VRAE4E, the long sequences of VRAE.py -- shapes bound by the latency of one step;
  * exact FFMA (crvae_gru_fwd / crvae_gru_bwd)         the general fallback (any B, in-kernel dW_hh).
The low-latency BPTT defers dW_hh to crvae_gru_dwhh_tc, which needs B % 32 == 0.
"""
from __future__ import annotations

import os

LL_MAX_HEADS = int(os.environ.get("CRVAE_LL_MAX_HEADS", "40"))    # decoder shards up to this many heads take the low-latency kernels
LL_ENABLED = os.environ.get("CRVAE_LL", "1") != "0"


def has_ll(k) -> bool:
    return LL_ENABLED and hasattr(k, "gru_fwd_ll") and hasattr(k, "gru_bwd_ll")


def gru_forward_small(k, gates, b_ih, w_hh, b_hh, h0, h0_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip):
    """Forward recurrence of a small head set (encoder, VRAE4E, generic VRAE: P = 1)."""
    fn = k.gru_fwd_ll if has_ll(k) else k.gru_fwd
    fn(gates, b_ih, w_hh, b_hh, h0, h0_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip)


def dwhh_workspace(k, P, T, B):
    """Workspace (floats) the deferred dW_hh GEMM of a small head set needs, or 0 when the in-kernel path is used."""
    if has_ll(k) and hasattr(k, "gru_dwhh_tc") and B % 32 == 0:
        return k.gru_dwhh_tc_workspace(P, T, B) // 4 + 4
    return 0


def gru_backward_small(k, gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, dw_hh, db_hh, db_ih, dw_lin, db_lin,
                       dh0, P, T, B, ws, ws_dwhh):
    """BPTT of a small head set: low-latency kernel + tcgen05 dW_hh GEMM when the batch allows, else the exact kernel."""
    if has_ll(k) and hasattr(k, "gru_dwhh_tc") and B % 32 == 0 and ws_dwhh is not None:
        k.gru_bwd_ll(gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws)
        k.gru_dwhh_tc(gates, ghn, hs, h0, h0_stride, dw_hh, P, T, B, ws_dwhh)
    else:
        k.gru_bwd(gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, dw_hh, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws)
