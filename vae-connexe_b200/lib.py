"""ctypes binding of libcrvae_b200.so (the C ABI declared in include/crvae_b200.h).

The product path has no CPU fallback: if the shared library cannot be loaded, or the current
device is not a B200-class (sm_100) GPU, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import build as _build

_c_void_p, _c_int, _c_i64, _c_float, _c_double, _c_size_t = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/crvae_b200.h declares
SIGNATURES = {
    "crvae_abi_version": (_c_int, []),
    "crvae_last_error": (C.c_char_p, []),
    "crvae_launch_count": (C.c_uint64, []),
    "crvae_launch_count_reset": (None, []),
    "crvae_check_device": (_c_int, []),
    "crvae_gemm_f32": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_int, _c_i64, _c_void_p, _c_int,
                                _c_i64, _c_void_p, _c_int, _c_i64, _c_void_p, _c_i64, _c_int, _c_void_p]),
    "crvae_proj_fwd": (_c_int, [_c_void_p] * 4 + [_c_int] * 5 + [_c_void_p]),
    "crvae_proj_fwd_tc": (_c_int, [_c_void_p] * 6 + [_c_int] * 5 + [_c_void_p]),
    "crvae_proj_wgrad_tc_workspace": (_c_size_t, [_c_int] * 5),
    "crvae_proj_wgrad_tc": (_c_int, [_c_void_p] * 5 + [_c_int] * 5 + [_c_void_p, _c_void_p]),
    "crvae_bind_batch": (_c_int, [_c_void_p] * 8 + [_c_int] * 6 + [_c_void_p]),
    "crvae_split_tf32": (_c_int, [_c_void_p] * 3 + [_c_i64, _c_void_p]),
    "crvae_split_tf32_gate_rows": (_c_int, [_c_void_p] * 3 + [_c_i64, _c_int, _c_void_p]),
    "crvae_proj_wgrad_workspace": (_c_size_t, [_c_int] * 4),
    "crvae_proj_wgrad": (_c_int, [_c_void_p] * 4 + [_c_int] * 5 + [_c_void_p, _c_void_p]),
    "crvae_gru_fwd": (_c_int, [_c_void_p] * 5 + [_c_i64] + [_c_void_p] * 5 + [_c_int] * 4 + [_c_void_p]),
    "crvae_gru_fwd_tc": (_c_int, [_c_void_p] * 6 + [_c_i64] + [_c_void_p] * 5 + [_c_int] * 4 + [_c_void_p]),
    "crvae_gru_bwd_workspace": (_c_size_t, [_c_int] * 2),
    "crvae_gru_bwd": (_c_int, [_c_void_p] * 4 + [_c_i64] + [_c_void_p] * 11 + [_c_int] * 3 + [_c_void_p, _c_void_p]),
    "crvae_gru_bwd_deferred": (_c_int, [_c_void_p] * 4 + [_c_i64] + [_c_void_p] * 10 + [_c_int] * 3 + [_c_void_p, _c_void_p]),
    "crvae_gru_bwd_tc": (_c_int, [_c_void_p] * 4 + [_c_i64] + [_c_void_p] * 9 + [_c_int] * 3 + [_c_void_p, _c_void_p]),
    "crvae_gru_fwd_ll": (_c_int, [_c_void_p] * 5 + [_c_i64] + [_c_void_p] * 5 + [_c_int] * 4 + [_c_void_p]),
    "crvae_gru_bwd_ll": (_c_int, [_c_void_p] * 4 + [_c_i64] + [_c_void_p] * 10 + [_c_int] * 3 + [_c_void_p, _c_void_p]),
    "crvae_gru_fwd_mma": (_c_int, [_c_void_p] * 5 + [_c_i64] + [_c_void_p] * 5 + [_c_int] * 4 + [_c_void_p]),
    "crvae_gru_bwd_mma": (_c_int, [_c_void_p] * 4 + [_c_i64] + [_c_void_p] * 10 + [_c_int] * 3 + [_c_void_p, _c_void_p]),
    "crvae_gru_dwhh_tc_workspace": (_c_size_t, [_c_int] * 3),
    "crvae_gru_dwhh_tc": (_c_int, [_c_void_p] * 4 + [_c_i64, _c_void_p] + [_c_int] * 3 + [_c_void_p, _c_void_p]),
    "crvae_latent_fwd": (_c_int, [_c_void_p] * 4 + [_c_int, _c_int, _c_int, _c_void_p]),
    "crvae_latent_head_workspace": (_c_size_t, [_c_int]),
    "crvae_latent_head_fwd": (_c_int, [_c_void_p] * 7 + [_c_int, _c_int, _c_void_p, _c_void_p]),
    "crvae_latent_head_bwd": (_c_int, [_c_void_p] * 6 + [_c_int, _c_void_p]),
    "crvae_latent_bwd": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_float, _c_int, _c_void_p,
                                  _c_void_p, _c_int, _c_int, _c_void_p]),
    "crvae_mse_fwd_bwd": (_c_int, [_c_void_p] * 5 + [_c_int] * 3 + [_c_float, _c_void_p]),
    "crvae_gd_step": (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_float, _c_void_p]),
    "crvae_gd_prox_gc": (_c_int, [_c_void_p] * 4 + [_c_int, _c_int, _c_float, _c_float, _c_int, _c_void_p]),
    "crvae_adam_step": (_c_int, [_c_void_p] * 4 + [_c_i64] + [_c_double] * 4 + [_c_int, _c_void_p]),
    "crvae_adam_step_dev": (_c_int, [_c_void_p] * 4 + [_c_i64] + [_c_double] * 4 + [_c_void_p, _c_void_p]),
    "crvae_cs_div_workspace": (_c_size_t, [_c_int, _c_int]),
    "crvae_cs_div_fwd_bwd": (_c_int, [_c_void_p] * 3 + [_c_int, _c_int, _c_float] + [_c_void_p] * 6),
    "crvae_tanh_fwd": (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_void_p]),
    "crvae_tanh_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_i64, _c_void_p]),
    "crvae_act_fwd": (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_int, _c_void_p]),
    "crvae_act_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_i64, _c_int, _c_void_p]),
    "crvae_transpose": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "crvae_gather_cols": (_c_int, [_c_void_p] * 4 + [_c_int, _c_i64, _c_int, _c_int, _c_void_p]),
    "crvae_proj_fwd_packed": (_c_int, [_c_void_p] * 4 + [_c_int] * 5 + [_c_void_p]),
    "crvae_proj_wgrad_packed_workspace": (_c_size_t, [_c_int] * 5),
    "crvae_proj_wgrad_packed": (_c_int, [_c_void_p] * 4 + [_c_int] * 6 + [_c_void_p, _c_void_p]),
    "crvae_dz_allreduce_bytes": (_c_size_t, [_c_int] * 3),
    "crvae_dz_allreduce_latent_bwd": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_float, _c_int,
                                              _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "crvae_bce_logits_workspace": (_c_size_t, [_c_i64]),
    "crvae_bce_logits_fwd_bwd": (_c_int, [_c_void_p] * 4 + [_c_i64, _c_float, _c_void_p, _c_void_p]),
    "crvae_ista_rows": (_c_int, [_c_void_p] * 3 + [_c_i64, _c_int, _c_float, _c_float, _c_int, _c_void_p]),
    "crvae_gen_scatter": (_c_int, [_c_void_p] * 6 + [_c_int] * 7 + [_c_float, _c_void_p]),
    "crvae_sumsq": (_c_int, [_c_void_p, _c_i64, _c_void_p, _c_void_p]),
    "crvae_dot_small": (_c_int, [_c_void_p, _c_int, _c_float, _c_void_p, _c_void_p]),
    "crvae_axpy": (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_float, _c_void_p]),
    "crvae_debug_set_batch_tile": (None, [_c_int]),
}

GEMM_NT, GEMM_TN, GEMM_NN = 0, 1, 2
KL_STANDARD, KL_SWAPPED, KL_LOGSIGMA = 0, 1, 2

_lib: Optional[C.CDLL] = None


class CrvaeLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the in-tree shared library (building it first when nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    # build() is a cheap source-digest compare when the library is up to date; it rebuilds a stale one (sources or the
    # header edited since) and falls back to the shipped binary where there is no nvcc
    path = _build.build(force=bool(os.environ.get("CRVAE_REBUILD")))
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.crvae_abi_version() != 1:
        raise CrvaeLibraryError("libcrvae_b200.so ABI version mismatch")
    _lib = lib
    return lib


def ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise CrvaeLibraryError("libcrvae_b200 takes CUDA device pointers only (no CPU path exists)")
    if t.data_ptr() % 4:
        raise CrvaeLibraryError("misaligned tensor")
    if t.dtype not in (torch.float32, torch.uint8, torch.int32):
        raise CrvaeLibraryError(f"unsupported dtype {t.dtype}")
    if not t.is_contiguous():
        raise CrvaeLibraryError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class Kernels:
    """Thin checked wrappers: raise on any non-zero return code (no silent fallback)."""

    device_type = "cuda"

    def __init__(self):
        if not torch.cuda.is_available():
            raise CrvaeLibraryError("vae-connexe_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = load()
        rc = self.lib.crvae_check_device()
        if rc != 0:
            raise CrvaeLibraryError(self.lib.crvae_last_error().decode())

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise CrvaeLibraryError(f"{what} failed (rc={rc}): {self.lib.crvae_last_error().decode()}")

    def launch_count(self) -> int:
        return int(self.lib.crvae_launch_count())

    def reset_launch_count(self):
        self.lib.crvae_launch_count_reset()

    def set_batch_tile(self, rows: int):
        self.lib.crvae_debug_set_batch_tile(int(rows))

    def gemm(self, form, batch, M, N, K, A, lda, sA, Bm, ldb, sB, Cm, ldc, sC, bias=None, sBias=0, accumulate=False):
        self._ck(self.lib.crvae_gemm_f32(form, batch, M, N, K, ptr(A), lda, sA, ptr(Bm), ldb, sB, ptr(Cm), ldc, sC,
                                         ptr(bias), sBias, int(accumulate), stream_ptr()), "crvae_gemm_f32")

    def proj_fwd(self, x, w_ih, b_ih, gates, P, T, B, K, t_skip):
        self._ck(self.lib.crvae_proj_fwd(ptr(x), ptr(w_ih), ptr(b_ih), ptr(gates), P, T, B, K, t_skip, stream_ptr()),
                 "crvae_proj_fwd")

    def proj_fwd_tc(self, x_hi, x_lo, w_hi, w_lo, b_ih, gates, P, T, B, K, t_skip):
        self._ck(self.lib.crvae_proj_fwd_tc(ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), ptr(b_ih), ptr(gates), P, T, B, K,
                                            t_skip, stream_ptr()), "crvae_proj_fwd_tc")

    def proj_wgrad_tc_workspace(self, P, T, B, K, t_skip) -> int:
        return int(self.lib.crvae_proj_wgrad_tc_workspace(P, T, B, K, t_skip))

    def proj_wgrad_tc(self, dgates, x_hi, x_lo, mask, dw_ih, P, T, B, K, t_skip, ws=None):
        self._ck(self.lib.crvae_proj_wgrad_tc(ptr(dgates), ptr(x_hi), ptr(x_lo), ptr(mask), ptr(dw_ih), P, T, B, K, t_skip,
                                              ptr(ws), stream_ptr()), "crvae_proj_wgrad_tc")

    def bind_batch(self, X, enc_in, enc_hi, enc_lo, dec_in, dec_hi, dec_lo, target, B, p, Te, Td, head_lo, P):
        self._ck(self.lib.crvae_bind_batch(ptr(X), ptr(enc_in), ptr(enc_hi), ptr(enc_lo), ptr(dec_in), ptr(dec_hi), ptr(dec_lo),
                                           ptr(target), B, p, Te, Td, head_lo, P, stream_ptr()), "crvae_bind_batch")

    def split_tf32_gate_rows(self, src, hi, lo, rows, cols):
        self._ck(self.lib.crvae_split_tf32_gate_rows(ptr(src), ptr(hi), ptr(lo), rows, cols, stream_ptr()), "crvae_split_tf32_gate_rows")

    def split_tf32(self, src, hi, lo, n):
        self._ck(self.lib.crvae_split_tf32(ptr(src), ptr(hi), ptr(lo), n, stream_ptr()), "crvae_split_tf32")

    def proj_wgrad_workspace(self, P, T, B, K) -> int:
        return int(self.lib.crvae_proj_wgrad_workspace(P, T, B, K))

    def proj_wgrad(self, dgates, x, mask, dw_ih, P, T, B, K, t_skip, ws):
        self._ck(self.lib.crvae_proj_wgrad(ptr(dgates), ptr(x), ptr(mask), ptr(dw_ih), P, T, B, K, t_skip, ptr(ws),
                                           stream_ptr()), "crvae_proj_wgrad")

    def gru_fwd(self, gates, b_ih, w_hh, b_hh, h0, h0_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip):
        self._ck(self.lib.crvae_gru_fwd(ptr(gates), ptr(b_ih), ptr(w_hh), ptr(b_hh), ptr(h0), h0_stride, ptr(w_lin),
                                        ptr(b_lin), ptr(hs), ptr(ghn), ptr(pred), P, T, B, t_skip, stream_ptr()),
                 "crvae_gru_fwd")

    def gru_fwd_tc(self, gates, b_ih, w_hh_hi, w_hh_lo, b_hh, h0, h0_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip):
        self._ck(self.lib.crvae_gru_fwd_tc(ptr(gates), ptr(b_ih), ptr(w_hh_hi), ptr(w_hh_lo), ptr(b_hh), ptr(h0), h0_stride,
                                           ptr(w_lin), ptr(b_lin), ptr(hs), ptr(ghn), ptr(pred), P, T, B, t_skip,
                                           stream_ptr()), "crvae_gru_fwd_tc")

    def gru_bwd_workspace(self, P, B) -> int:
        return int(self.lib.crvae_gru_bwd_workspace(P, B))

    def gru_bwd(self, gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, dw_hh, db_hh, db_ih, dw_lin,
                db_lin, dh0, P, T, B, ws):
        self._ck(self.lib.crvae_gru_bwd(ptr(gates), ptr(ghn), ptr(hs), ptr(h0), h0_stride, ptr(w_hh), ptr(w_lin),
                                        ptr(dpred), ptr(dh_last), ptr(dhs), ptr(dw_hh), ptr(db_hh), ptr(db_ih), ptr(dw_lin),
                                        ptr(db_lin), ptr(dh0), P, T, B, ptr(ws), stream_ptr()), "crvae_gru_bwd")

    def gru_bwd_deferred(self, gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, db_hh, db_ih, dw_lin, db_lin,
                         dh0, P, T, B, ws):
        self._ck(self.lib.crvae_gru_bwd_deferred(ptr(gates), ptr(ghn), ptr(hs), ptr(h0), h0_stride, ptr(w_hh), ptr(w_lin),
                                                 ptr(dpred), ptr(dh_last), ptr(dhs), ptr(db_hh), ptr(db_ih), ptr(dw_lin),
                                                 ptr(db_lin), ptr(dh0), P, T, B, ptr(ws), stream_ptr()), "crvae_gru_bwd_deferred")

    def gru_bwd_tc(self, gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, db_hh, db_ih, dw_lin, db_lin, dh0, P, T, B, ws):
        self._ck(self.lib.crvae_gru_bwd_tc(ptr(gates), ptr(ghn), ptr(hs), ptr(h0), h0_stride, ptr(w_hh), ptr(w_lin), ptr(dpred),
                                           ptr(dh_last), ptr(db_hh), ptr(db_ih), ptr(dw_lin), ptr(db_lin), ptr(dh0), P, T, B,
                                           ptr(ws), stream_ptr()), "crvae_gru_bwd_tc")

    def gru_fwd_ll(self, gates, b_ih, w_hh, b_hh, h0, h0_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip):
        self._ck(self.lib.crvae_gru_fwd_ll(ptr(gates), ptr(b_ih), ptr(w_hh), ptr(b_hh), ptr(h0), h0_stride, ptr(w_lin),
                                           ptr(b_lin), ptr(hs), ptr(ghn), ptr(pred), P, T, B, t_skip, stream_ptr()),
                 "crvae_gru_fwd_ll")

    def gru_bwd_ll(self, gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, db_hh, db_ih, dw_lin, db_lin,
                   dh0, P, T, B, ws):
        self._ck(self.lib.crvae_gru_bwd_ll(ptr(gates), ptr(ghn), ptr(hs), ptr(h0), h0_stride, ptr(w_hh), ptr(w_lin),
                                           ptr(dpred), ptr(dh_last), ptr(dhs), ptr(db_hh), ptr(db_ih), ptr(dw_lin),
                                           ptr(db_lin), ptr(dh0), P, T, B, ptr(ws), stream_ptr()), "crvae_gru_bwd_ll")

    def gru_fwd_mma(self, gates, b_ih, w_hh, b_hh, h0, h0_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip):
        self._ck(self.lib.crvae_gru_fwd_mma(ptr(gates), ptr(b_ih), ptr(w_hh), ptr(b_hh), ptr(h0), h0_stride, ptr(w_lin),
                                            ptr(b_lin), ptr(hs), ptr(ghn), ptr(pred), P, T, B, t_skip, stream_ptr()),
                 "crvae_gru_fwd_mma")

    def gru_bwd_mma(self, gates, ghn, hs, h0, h0_stride, w_hh, w_lin, dpred, dh_last, dhs, db_hh, db_ih, dw_lin, db_lin,
                    dh0, P, T, B, ws):
        self._ck(self.lib.crvae_gru_bwd_mma(ptr(gates), ptr(ghn), ptr(hs), ptr(h0), h0_stride, ptr(w_hh), ptr(w_lin),
                                            ptr(dpred), ptr(dh_last), ptr(dhs), ptr(db_hh), ptr(db_ih), ptr(dw_lin),
                                            ptr(db_lin), ptr(dh0), P, T, B, ptr(ws), stream_ptr()), "crvae_gru_bwd_mma")

    def gru_dwhh_tc_workspace(self, P, T, B) -> int:
        return int(self.lib.crvae_gru_dwhh_tc_workspace(P, T, B))

    def gru_dwhh_tc(self, dgates, dghn, hs, h0, h0_stride, dw_hh, P, T, B, ws=None):
        self._ck(self.lib.crvae_gru_dwhh_tc(ptr(dgates), ptr(dghn), ptr(hs), ptr(h0), h0_stride, ptr(dw_hh), P, T, B,
                                            ptr(ws), stream_ptr()), "crvae_gru_dwhh_tc")

    def latent_fwd(self, lat, eps, z, kl_out, B, kl_form, Z=64):
        self._ck(self.lib.crvae_latent_fwd(ptr(lat), ptr(eps), ptr(z), ptr(kl_out), B, Z, kl_form, stream_ptr()),
                 "crvae_latent_fwd")

    def latent_head_workspace(self, B) -> int:
        return int(self.lib.crvae_latent_head_workspace(B))

    def latent_head_fwd(self, hT, lat_w, lat_b, eps, lat, z, kl_out, B, kl_form, ws):
        self._ck(self.lib.crvae_latent_head_fwd(ptr(hT), ptr(lat_w), ptr(lat_b), ptr(eps), ptr(lat), ptr(z), ptr(kl_out), B, kl_form,
                                                ptr(ws), stream_ptr()), "crvae_latent_head_fwd")

    def latent_head_bwd(self, dlat, hT, lat_w, d_lat_w, d_lat_b, dhT, B):
        self._ck(self.lib.crvae_latent_head_bwd(ptr(dlat), ptr(hT), ptr(lat_w), ptr(d_lat_w), ptr(d_lat_b), ptr(dhT), B, stream_ptr()),
                 "crvae_latent_head_bwd")

    def latent_bwd(self, dh0, P, dz_extra, lat, eps, beta, kl_form, dlat, dz_out, B, Z=64):
        self._ck(self.lib.crvae_latent_bwd(ptr(dh0), P, ptr(dz_extra), ptr(lat), ptr(eps), float(beta), kl_form,
                                           ptr(dlat), ptr(dz_out), B, Z, stream_ptr()), "crvae_latent_bwd")

    def mse_fwd_bwd(self, pred, target, sse, dpred, err, P, T, B, dscale=0.0):
        self._ck(self.lib.crvae_mse_fwd_bwd(ptr(pred), ptr(target), ptr(sse), ptr(dpred), ptr(err), P, T, B, float(dscale),
                                            stream_ptr()), "crvae_mse_fwd_bwd")

    def gd_step(self, theta, grad, n, lr):
        self._ck(self.lib.crvae_gd_step(ptr(theta), ptr(grad), n, float(lr), stream_ptr()), "crvae_gd_step")

    def gd_prox_gc(self, w_ih, dw_ih, mask, col_norm, P, K, lr, thr, do_prox):
        self._ck(self.lib.crvae_gd_prox_gc(ptr(w_ih), ptr(dw_ih), ptr(mask), ptr(col_norm), P, K, float(lr),
                                           float(thr), int(do_prox), stream_ptr()), "crvae_gd_prox_gc")

    def adam_step(self, theta, grad, m, v, n, lr, b1, b2, eps, step):
        self._ck(self.lib.crvae_adam_step(ptr(theta), ptr(grad), ptr(m), ptr(v), n, lr, b1, b2, eps, step,
                                          stream_ptr()), "crvae_adam_step")

    def adam_step_dev(self, theta, grad, m, v, n, lr, b1, b2, eps, counter):
        self._ck(self.lib.crvae_adam_step_dev(ptr(theta), ptr(grad), ptr(m), ptr(v), n, lr, b1, b2, eps, ptr(counter),
                                              stream_ptr()), "crvae_adam_step_dev")

    def cs_div_workspace(self, B, K) -> int:
        return int(self.lib.crvae_cs_div_workspace(B, K))

    def cs_div_fwd_bwd(self, lat, prior_mu, prior_logvar, B, K, scale_loss, cs_mean, dlat, dprior_mu, dprior_logvar, ws):
        self._ck(self.lib.crvae_cs_div_fwd_bwd(ptr(lat), ptr(prior_mu), ptr(prior_logvar), B, K, float(scale_loss), ptr(cs_mean),
                                               ptr(dlat), ptr(dprior_mu), ptr(dprior_logvar), ptr(ws), stream_ptr()),
                 "crvae_cs_div_fwd_bwd")

    def tanh_fwd(self, x, y, n):
        self._ck(self.lib.crvae_tanh_fwd(ptr(x), ptr(y), n, stream_ptr()), "crvae_tanh_fwd")

    def tanh_bwd(self, dy, y, dx, n):
        self._ck(self.lib.crvae_tanh_bwd(ptr(dy), ptr(y), ptr(dx), n, stream_ptr()), "crvae_tanh_bwd")

    def act_fwd(self, x, y, n, kind):
        self._ck(self.lib.crvae_act_fwd(ptr(x), ptr(y), n, kind, stream_ptr()), "crvae_act_fwd")

    def act_bwd(self, dy, y, dx, n, kind):
        self._ck(self.lib.crvae_act_bwd(ptr(dy), ptr(y), ptr(dx), n, kind, stream_ptr()), "crvae_act_bwd")

    def transpose(self, src, dst, rows, cols):
        self._ck(self.lib.crvae_transpose(ptr(src), ptr(dst), rows, cols, stream_ptr()), "crvae_transpose")

    def gather_cols(self, x, cols, mask, xg, P, rows, K, Kp):
        self._ck(self.lib.crvae_gather_cols(ptr(x), ptr(cols), ptr(mask), ptr(xg), P, rows, K, Kp, stream_ptr()), "crvae_gather_cols")

    def proj_fwd_packed(self, xg, w_ih, b_ih, gates, P, T, B, Kp, t_skip):
        self._ck(self.lib.crvae_proj_fwd_packed(ptr(xg), ptr(w_ih), ptr(b_ih), ptr(gates), P, T, B, Kp, t_skip, stream_ptr()),
                 "crvae_proj_fwd_packed")

    def proj_wgrad_packed_workspace(self, P, T, B, Kp, K_dense) -> int:
        return int(self.lib.crvae_proj_wgrad_packed_workspace(P, T, B, Kp, K_dense))

    def proj_wgrad_packed(self, dgates, xg, mask, dw_ih, P, T, B, Kp, K_dense, t_skip, ws):
        self._ck(self.lib.crvae_proj_wgrad_packed(ptr(dgates), ptr(xg), ptr(mask), ptr(dw_ih), P, T, B, Kp, K_dense, t_skip, ptr(ws),
                                                  stream_ptr()), "crvae_proj_wgrad_packed")

    def dz_allreduce_bytes(self, B, Z, world) -> int:
        return int(self.lib.crvae_dz_allreduce_bytes(B, Z, world))

    def dz_allreduce_latent_bwd(self, dh0, P, peer_ptrs, rank, world, lat, eps, beta, kl_form, dlat, dz_out, B, Z=64):
        arr = (C.c_void_p * world)(*[int(x) for x in peer_ptrs])
        self._ck(self.lib.crvae_dz_allreduce_latent_bwd(ptr(dh0), P, arr, rank, world, ptr(lat), ptr(eps), float(beta), kl_form,
                                                        ptr(dlat), ptr(dz_out), B, Z, stream_ptr()), "crvae_dz_allreduce_latent_bwd")

    def bce_logits_workspace(self, n) -> int:
        return int(self.lib.crvae_bce_logits_workspace(n))

    def bce_logits_fwd_bwd(self, logits, x, sum_out, dlogits, n, dscale, ws):
        self._ck(self.lib.crvae_bce_logits_fwd_bwd(ptr(logits), ptr(x), ptr(sum_out), ptr(dlogits), n, float(dscale), ptr(ws),
                                                   stream_ptr()), "crvae_bce_logits_fwd_bwd")

    def ista_rows(self, w, dw, row_norm, rows, cols, lr, thr, do_prox):
        self._ck(self.lib.crvae_ista_rows(ptr(w), ptr(dw), ptr(row_norm), rows, cols, float(lr), float(thr), int(do_prox), stream_ptr()),
                 "crvae_ista_rows")

    def gen_scatter(self, y, noise, x, x_hi, x_lo, out, B, p, t, steps, base, rem, widest, scale):
        self._ck(self.lib.crvae_gen_scatter(ptr(y), ptr(noise), ptr(x), ptr(x_hi), ptr(x_lo), ptr(out), B, p, t, steps, base, rem,
                                            widest, float(scale), stream_ptr()), "crvae_gen_scatter")

    def sumsq(self, x, n, out):
        self._ck(self.lib.crvae_sumsq(ptr(x), n, ptr(out), stream_ptr()), "crvae_sumsq")

    def dot_small(self, x, n, scale, out):
        self._ck(self.lib.crvae_dot_small(ptr(x), n, float(scale), ptr(out), stream_ptr()), "crvae_dot_small")

    def axpy(self, y, x, n, alpha):
        self._ck(self.lib.crvae_axpy(ptr(y), ptr(x), n, float(alpha), stream_ptr()), "crvae_axpy")


_kernels: Optional[Kernels] = None


def kernels() -> Kernels:
    """The process-wide kernel table.  Always the CUDA library; tests/ may install a checker
    backend with set_test_backend() to exercise the host logic on a machine without a GPU."""
    global _kernels
    if _kernels is None:
        _kernels = Kernels()
    return _kernels


def set_test_backend(backend) -> None:
    """TEST HOOK ONLY (tests/cpu_backend.py): replace the kernel table.  Nothing in the package
    calls this; with no backend installed every entry point needs libcrvae_b200.so + a B200."""
    global _kernels
    _kernels = backend
