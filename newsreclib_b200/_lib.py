"""ctypes binding of the C ABI declared in ``include/nrl.h``.

The shared library is built in-tree (``python __graft_entry__.py`` or
``newsreclib_b200.build.build()``) as ``newsreclib_b200/libnrl_b200.so``.  There is no CPU
fallback: importing an op without the library raises, and every non-zero status from the
library raises ``RuntimeError`` carrying ``nrl_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnrl_b200.so")

c_ll_p = C.POINTER(C.c_longlong)
c_f_p = C.POINTER(C.c_float)
c_i_p = C.POINTER(C.c_int)


class BlockParams(C.Structure):
    """``nrl_block_params`` / ``nrl_block_grads`` (same layout: seven float pointers)."""

    _fields_ = [
        ("in_proj_weight", C.c_void_p),
        ("in_proj_bias", C.c_void_p),
        ("out_proj_weight", C.c_void_p),
        ("out_proj_bias", C.c_void_p),
        ("add_weight", C.c_void_p),
        ("add_bias", C.c_void_p),
        ("add_query", C.c_void_p),
    ]


class CnnParams(C.Structure):
    """``nrl_cnn_params`` / ``nrl_cnn_grads`` (five float pointers)."""

    _fields_ = [
        ("cnn_weight", C.c_void_p),
        ("cnn_bias", C.c_void_p),
        ("add_weight", C.c_void_p),
        ("add_bias", C.c_void_p),
        ("add_query", C.c_void_p),
    ]


class CnnDims(C.Structure):
    _fields_ = [("embed_dim", C.c_int), ("num_filters", C.c_int), ("window", C.c_int), ("query_dim", C.c_int)]


MAX_RANKS = 16
FLAG_BYTES = 512
IPC_HANDLE_BYTES = 64


class PeerSet(C.Structure):
    """``nrl_peer_set``: the peer-mapped flat parameter / gradient buffers and flag blocks of every rank."""

    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("params", C.c_void_p * MAX_RANKS),
                ("grads", C.c_void_p * MAX_RANKS), ("flags", C.c_void_p * MAX_RANKS),
                ("bitmaps", C.c_void_p * MAX_RANKS)]


class Dims(C.Structure):
    _fields_ = [("embed_dim", C.c_int), ("num_heads", C.c_int), ("query_dim", C.c_int)]


TFM_EMBED_FIELDS = ("word", "pos", "type0", "ln_g", "ln_b")
TFM_LAYER_FIELDS = ("q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "ao_w", "ao_b", "ln1_g", "ln1_b", "i_w", "i_b", "o_w",
                    "o_b", "ln2_g", "ln2_b")


class TfmEmbed(C.Structure):
    """``nrl_tfm_embed_params`` / ``nrl_tfm_embed_grads`` (five float pointers)."""

    _fields_ = [(n, C.c_void_p) for n in TFM_EMBED_FIELDS]


class TfmLayer(C.Structure):
    """``nrl_tfm_layer_params`` / ``nrl_tfm_layer_grads`` (sixteen float pointers)."""

    _fields_ = [(n, C.c_void_p) for n in TFM_LAYER_FIELDS]


class TfmDims(C.Structure):
    _fields_ = [("hidden", C.c_int), ("heads", C.c_int), ("intermediate", C.c_int), ("num_layers", C.c_int),
                ("vocab", C.c_int), ("max_pos", C.c_int), ("pad_idx", C.c_int), ("ln_eps", C.c_float),
                ("hidden_dropout", C.c_float), ("attn_dropout", C.c_float), ("position_mode", C.c_int)]


PREC_BF16X3 = 0
PREC_BF16 = 1

# name -> (restype, argtypes); every symbol include/nrl.h declares
_VP, _LL, _I, _F, _ULL, _SZ = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_ulonglong, C.c_size_t
_BP = C.POINTER(BlockParams)
_CP = C.POINTER(CnnParams)
SIGNATURES = {
    "nrl_version": (C.c_char_p, []),
    "nrl_last_error": (C.c_char_p, []),
    "nrl_launch_count": (_LL, []),
    "nrl_profile_start": (_I, [_VP]),
    "nrl_profile_stop": (_I, [_VP, _I, _VP, _I]),
    "nrl_news_encoder_ws_bytes": (_SZ, [_LL, _I, Dims]),
    "nrl_news_encoder_fwd": (_I, [_VP, _LL, _I, _VP, _LL, _BP, Dims, _F, _I, _ULL, _VP, _VP, _SZ, _I, _VP]),
    "nrl_news_encoder_bwd": (_I, [_VP, _LL, _I, _LL, _BP, Dims, _F, _I, _ULL, _VP, _BP, _VP, _VP, _SZ, _I, _VP]),
    "nrl_user_encoder_ws_bytes": (_SZ, [_I, _I, Dims]),
    "nrl_user_encoder_fwd": (_I, [_VP, _I, _I, _BP, Dims, _I, _VP, _VP, _SZ, _I, _VP]),
    "nrl_user_encoder_bwd": (_I, [_I, _I, _BP, Dims, _I, _VP, _BP, _VP, _VP, _SZ, _I, _VP]),
    "nrl_additive_ws_bytes": (_SZ, [_LL, _I, _I, _I]),
    "nrl_additive_fwd": (_I, [_VP, _LL, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _SZ, _I, _VP]),
    "nrl_additive_bwd": (_I, [_VP, _LL, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _I, _VP]),
    "nrl_cnn_encoder_ws_bytes": (_SZ, [_LL, _I, CnnDims]),
    "nrl_cnn_encoder_fwd": (_I, [_VP, _LL, _I, _VP, _LL, _CP, CnnDims, _F, _I, _ULL, _VP, _VP, _SZ, _I, _VP]),
    "nrl_cnn_encoder_bwd": (_I, [_VP, _LL, _I, _LL, _CP, CnnDims, _F, _I, _ULL, _VP, _CP, _VP, _VP, _SZ, _I, _VP]),
    "nrl_linear_encoder_ws_bytes": (_SZ, [_LL, _I, _I]),
    "nrl_linear_encoder_fwd": (_I, [_VP, _LL, _VP, _LL, _I, _VP, _VP, _I, _F, _I, _ULL, _VP, _VP, _SZ, _I, _VP]),
    "nrl_linear_encoder_bwd": (_I, [_VP, _LL, _LL, _I, _VP, _I, _F, _I, _ULL, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _I, _VP]),
    "nrl_plm_head_fwd": (_I, [_VP, _I, _I, _BP, Dims, _I, _F, _I, _ULL, _VP, _VP, _SZ, _I, _VP]),
    "nrl_plm_head_bwd": (_I, [_I, _I, _BP, Dims, _I, _F, _I, _ULL, _VP, _BP, _VP, _VP, _SZ, _I, _VP]),
    "nrl_segment_offsets": (_I, [_VP, _LL, _I, _VP, _VP]),
    "nrl_to_dense_fwd": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "nrl_to_dense_bwd": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "nrl_gather_rows": (_I, [_VP, _LL, _I, _VP, _LL, _VP, _VP]),
    "nrl_late_fusion_fwd": (_I, [_VP, _VP, _I, _I, _VP, _VP]),
    "nrl_late_fusion_bwd": (_I, [_VP, _VP, _I, _I, _VP, _VP]),
    "nrl_score_fwd": (_I, [_VP, _VP, _VP, _I, _I, _I, _VP, _VP]),
    "nrl_score_bwd": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _VP, _VP, _VP]),
    "nrl_ce_soft_fwd": (_I, [_VP, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP]),
    "nrl_ce_soft_bwd": (_I, [_VP, _VP, _VP, _I, _I, _VP, _F, _VP, _VP]),
    "nrl_rank_metrics": (_I, [_VP, _VP, _VP, _I, C.POINTER(C.c_int), _I, _VP, _VP, _VP]),
    "nrl_supcon_fwd": (_I, [_VP, _VP, _VP, _I, _I, _F, _VP, _F, _VP, _VP, _VP, _VP]),
    "nrl_supcon_bwd": (_I, [_VP, _VP, _VP, _I, _I, _F, _VP, _VP, _VP, _F, _I, _VP, _VP]),
    "nrl_adam_step": (_I, [_VP, _VP, _VP, _VP, _LL, _F, _F, _F, _F, _LL, _F, _VP]),
    "nrl_adam_step_zero_grad": (_I, [_VP, _VP, _VP, _VP, _LL, _F, _F, _F, _F, _LL, _F, _VP]),
    "nrl_embedding_gather": (_I, [_VP, _LL, _VP, _LL, _I, _VP, _VP, _VP, _VP]),
    "nrl_device_status": (_I, [C.POINTER(C.c_int), _VP]),
    "nrl_peer_alloc": (_I, [_SZ, C.POINTER(C.c_void_p), C.c_char_p]),
    "nrl_peer_free": (_I, [_VP]),
    "nrl_peer_open": (_I, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "nrl_peer_close": (_I, [_VP]),
    "nrl_exchange_adam_step": (_I, [C.POINTER(PeerSet), _VP, _VP, _LL, _F, _F, _F, _F, _LL, _ULL, _F, _I, _ULL, _LL, _I, _I,
                                    _VP]),
    "nrl_exchange_status": (_I, [_VP, C.POINTER(_ULL), _VP]),
    "nrl_nrms_ws_bytes": (_SZ, [_LL, _LL, _I, _I, _I, _I, Dims]),
    "nrl_nrms_step": (_I, [_VP, _VP, _VP, _VP, _VP, _LL, _LL, _I, _I, _I, _I, _VP, _LL, _BP, _BP, Dims,
                           _I, _F, _I, _ULL, _VP, _VP, _I, _BP, _BP, _VP, _VP, _SZ, _I, _VP]),
    "nrl_nrms_step_host": (_I, [_VP, _VP, _VP, _VP, _VP, _LL, _LL, _I, _I, _I, _I, _VP, _LL, _BP, _BP, Dims,
                                _I, _F, _I, _ULL, _VP, _VP, _I, _BP, _BP, _VP, _VP, _SZ, _I, _VP]),
    "nrl_nrms_step_bwd": (_I, [_VP, _VP, _LL, _LL, _I, _I, _I, _I, _LL, _BP, _BP, Dims, _I, _F, _I, _ULL, _BP, _BP, _VP,
                               _VP, _SZ, _I, _VP]),
    "nrl_nrms_step_host_begin": (_I, [_VP, _VP, _VP, _VP, _VP, _LL, _LL, _I, _I, _I, _I, _VP, _LL, _BP, _BP, Dims,
                                      _I, _F, _I, _ULL, _VP, _VP, _I, _BP, _BP, _VP, _VP, _SZ, _I, _VP, _VP, _VP,
                                      C.POINTER(C.c_void_p)]),
    "nrl_nrms_step_host_end": (_I, [_VP]),
    "nrl_tfm_wpack_bytes": (_SZ, [TfmDims]),
    "nrl_tfm_pack_weights": (_I, [_VP, _I, _I, TfmDims, _VP, _SZ, _I, _VP]),
    "nrl_tfm_ws_bytes": (_SZ, [_LL, _I, TfmDims, _I]),
    "nrl_tfm_encoder_fwd": (_I, [_VP, _VP, _I, _I, _VP, _VP, TfmDims, _I, _ULL, _VP, _VP, _I, _VP, _SZ, _I, _VP]),
    "nrl_tfm_encoder_bwd": (_I, [_VP, _VP, _I, _I, _VP, _VP, TfmDims, _I, _ULL, _VP, _VP, _VP, _VP, _VP, _SZ, _I, _VP]),
    "nrl_tfm_attn_dropout_mask": (_I, [_VP, _I, _I, _I, _I, _I, _ULL, _F, _VP]),
    "nrl_tfm_hidden_dropout_mask": (_I, [_VP, _LL, _I, _I, _ULL, _F, _VP]),
    "nrl_dropout_mask": (_I, [_VP, _LL, _ULL, _I, _F, _VP]),
    "nrl_gemm_test_ws_bytes": (_SZ, [_I, _I, _I]),
    "nrl_gemm_test": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _VP, _SZ, _VP]),
    "nrl_gemm_test_planes": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _VP, _SZ, _VP]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the library and bind every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: newsreclib_b200 has no CPU fallback. Build the sm_100a "
            "library first with `python __graft_entry__.py` (needs nvcc)."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().nrl_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {status}: {msg}")
