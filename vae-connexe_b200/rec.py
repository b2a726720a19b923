"""Choice of the recurrent kernel for a (heads, batch) shape.

Four implementations of the same GRU recurrence / BPTT live in the library (include/crvae_b200.h):
  * warp-level MMA (crvae_gru_fwd_mma / crvae_gru_bwd_mma) 16-row tiles, W_hh in registers, 3xTF32: small head shards, the
    encoder, VRAE4E, VRAE.py -- the default for the latency-bound shapes;
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


# warp-level MMA kernels (crvae_gru_fwd_mma: warp-specialised, TMA-plumbed; crvae_gru_bwd_mma), 3xTF32 like the tcgen05 kernels.
# CRVAE_MMA: "small" (default) = the shapes bound by the latency of one step: decoder shards up to MMA_MAX_HEADS heads, the
# replicated encoder, VRAE4E, VRAE.py; "1" = every decoder shape; "0" = off (exact-fp32 low-latency kernels instead).
# Measured on one B200 (tools/time_gru_ll.py, B = 256, T = 10, forward): P = 1: ll 31.6, mma 23.6 us; P = 13: ll 50.2, mma 35.7,
# tcgen05 63.5; P = 25: ll 72.7, mma 62.4, tcgen05 64.5; P = 50: mma 90.9, tcgen05 73.7; P = 100: mma 171, tcgen05 138.
MMA_MODE = os.environ.get("CRVAE_MMA", "small")
MMA_MAX_HEADS = int(os.environ.get("CRVAE_MMA_MAX_HEADS", "26"))


def has_mma(k) -> bool:
    return MMA_MODE != "0" and hasattr(k, "gru_fwd_mma") and hasattr(k, "gru_bwd_mma")


def mma_preferred(P: int) -> bool:
    """Decoder shards of P heads take the MMA kernels (auto mode)."""
    return MMA_MODE == "1" or (MMA_MODE == "small" and P <= MMA_MAX_HEADS)


# BPTT: the warp-specialised K-split MMA kernel is the fastest BPTT from one head up to ~100 heads (B = 256, T = 10, us:
# P = 1: ll 42.0, mma 29.5; P = 13: ll 66.4, mma 41.8, tcgen05 64; P = 50: mma 105.5, tcgen05 110; P = 100: mma 193.5, tcgen05 201.7);
# beyond that the tcgen05 kernel's many full waves win (p = 1000: 149 us per 100 heads).
MMA_BWD_MAX_TILES = int(os.environ.get("CRVAE_MMA_BWD_MAX_TILES", "1700"))


def mma_bwd_preferred(k, P: int, B: int) -> bool:
    return has_mma(k) and (MMA_MODE == "1" or P * ((B + 15) // 16) <= MMA_BWD_MAX_TILES)


def has_ll(k) -> bool:
    return LL_ENABLED and hasattr(k, "gru_fwd_ll") and hasattr(k, "gru_bwd_ll")


def gru_forward_small(k, gates, b_ih, w_hh, b_hh, h0, h0_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip):
    """Forward recurrence of a small head set (encoder, VRAE4E, generic VRAE: P = 1)."""
    fn = k.gru_fwd_mma if has_mma(k) else (k.gru_fwd_ll if has_ll(k) else k.gru_fwd)
    fn(gates, b_ih, w_hh, b_hh, h0, h0_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip)


def dwhh_workspace(k, P, T, B):
    """Workspace (floats) the deferred dW_hh GEMM of a small head set needs, or 0 when the in-kernel path is used."""
    if (has_ll(k) or has_mma(k)) and hasattr(k, "gru_dwhh_tc") and B % 32 == 0:
        return k.gru_dwhh_tc_workspace(P, T, B) // 4 + 4
    return 0


def gru_backward_small(k, gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, dw_hh, db_hh, db_ih, dw_lin, db_lin,
                       dh0, P, T, B, ws, ws_dwhh, split_dwhh=False):
    """BPTT of a small head set: MMA / low-latency kernel + tcgen05 dW_hh GEMM when the batch allows, else the exact kernel.
    split_dwhh=True: when dW_hh is a separate GEMM it is NOT launched; a callable that launches it is returned instead, so the
    caller can put it on another stream next to the projection weight gradient (both only read the BPTT's outputs)."""
    if (has_ll(k) or has_mma(k)) and hasattr(k, "gru_dwhh_tc") and B % 32 == 0 and ws_dwhh is not None:
        bwd = k.gru_bwd_mma if mma_bwd_preferred(k, P, B) else k.gru_bwd_ll
        bwd(gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws)
        dwhh = lambda: k.gru_dwhh_tc(gates, ghn, hs, h0, h0_stride, dw_hh, P, T, B, ws_dwhh)
        if split_dwhh:
            return dwhh
        dwhh()
    else:
        k.gru_bwd(gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, dw_hh, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws)
    return None
