"""ctypes binding of libb200med.so (C ABI declared in include/b200med.h).

There is NO CPU fallback: if the shared library has not been built, importing any op raises.
Build it with ``python -m multimodal_error_detection_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MED_LIB") or os.path.join(_PKG, "libb200med.so")      # B200MED_LIB: another build of the same ABI (A/B timing)

F32, BF16, F16 = 0, 1, 2


class StreamDesc(C.Structure):
    """Mirror of ``b200med_stream_desc`` (include/b200med.h)."""
    _fields_ = [("table", C.c_void_p), ("mean", C.c_void_p), ("stdv", C.c_void_p), ("out", C.c_void_p),
                ("dim", C.c_int32), ("table_dtype", C.c_int32), ("out_dtype", C.c_int32), ("out_ld", C.c_int32),
                ("out_col", C.c_int32), ("stat_rows", C.c_int32), ("exact_div", C.c_int32), ("table_rows", C.c_int32)]


_p, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol of include/b200med.h
SIGNATURES = {
    "b200med_version": (C.c_int, []),
    "b200med_last_error": (C.c_char_p, []),
    "b200med_launch_count": (_i64, []),
    "b200med_set_sm_limit": (C.c_int, [_i32]),
    "b200med_set_pdl": (C.c_int, [_i32]),
    "b200med_window_count": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p, _p]),
    "b200med_window_fill": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "b200med_powerset": (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    "b200med_gather_norm": (C.c_int, [C.POINTER(StreamDesc), _i32, _p, _i64, _i32, _i32, _p]),
    "b200med_gather_last_variant": (C.c_int, []),
    "b200med_gather_linear_bf16": (C.c_int, [_p, _i64, _p, _p, _p, _i64, _i32, _p, _p, _i32, _p, _p, _i32, _i32, _p]),
    "b200med_standardise_rows": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "b200med_linear_fwd_f32": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "b200med_gemm_f32_ws_bytes": (_i64, [_i64, _i64, _i64]),
    "b200med_gemm_f32": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _p, _i64, _i32, _p, _p]),
    "b200med_linear_bwd_data_f32": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _p]),
    "b200med_linear_bwd_weight_ws_bytes": (_i64, [_i64, _i32, _i32]),
    "b200med_linear_bwd_weight_f32": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _p]),
    "b200med_gemm_bf16_ws_bytes": (_i64, [_i64, _i64, _i64, _i32]),
    "b200med_gemm_bf16_pick_split": (_i32, [_i64, _i64, _i64, _i32]),
    "b200med_gemm_bf16": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32,
                                    _i32, _i32, _p, _p]),
    "b200med_has_tcgen05": (C.c_int, []),
    "b200med_colsum": (C.c_int, [_p, _i32, _p, _i64, _i32, _i64, _p, _p]),
    "b200med_colsum_ws_bytes": (_i64, [_i64, _i32]),
    "b200med_cast_f32_to_bf16": (C.c_int, [_p, _p, _i64, _p]),
    "b200med_relu_cast_f32_to_bf16": (C.c_int, [_p, _p, _i64, _p]),
    "b200med_split_bf16x3": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _i32, _p]),
    "b200med_split_bf16x6": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "b200med_cast_bf16_to_f32": (C.c_int, [_p, _p, _i64, _p]),
    "b200med_lstm_pack_inputs": (C.c_int, [_p, _p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "b200med_lstm_pack_parts": (C.c_int, [_p, _i32, _p, _i64, _i32, _p, _p, _i32, _p, _p, _i64, _i64, _i32, _i32, _i32, _i32, _p]),
    "b200med_lstm_pack_parts_bf16": (C.c_int, [_p, _i32, _p, _i64, _i32, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _p]),
    "b200med_lstm_unpack_dx_bf16": (C.c_int, [_p, _p, _i64, _i64, _i32, _i32, _i32, _p]),
    "b200med_lstm_rec_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i32, _p, _i64, _i64, _i32, _i32, _f, _p,
                                       C.c_uint64, _p]),
    "b200med_lstm_rec_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i32, _p, _i64, _i64, _i32, _i32, _f, _p, C.c_uint64, _p]),
    "b200med_lstm_pack_weights2": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _p, _p, _p]),
    "b200med_lstm_rec2_fwd": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _p, _i32, _p, _i64, _i64, _i32, _f, _p, C.c_uint64, _p]),
    "b200med_lstm_pack_weights2_bwd": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p]),
    "b200med_lstm_rec2_bwd": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, _p, _i64, _i64, _i32, _f, _p, C.c_uint64, _p]),
    "b200med_lstm_unpack_grads2": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p, _p]),
    "b200med_lstm_unpack_dx": (C.c_int, [_p, _p, _i64, _i64, _i32, _i32, _i32, _i32, _p]),
    "b200med_zero_cols_bf16": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p]),
    "b200med_lstm_cell_fwd": (C.c_int, [_p, _p, _p, _p, _i32, _p, _i32, _p, _i64, _i32, _f, _p, C.c_uint64, _p]),
    "b200med_lstm_cell_bwd": (C.c_int, [_p, _p, _p, _p, _i32, _p, _i32, _p, _i32, _p, _i64, _i32, _f, _p, C.c_uint64, _p]),
    "b200med_lstm_pack_f32": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p]),
    "b200med_lstm_unpack_f32": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _i32, _p]),
    "b200med_lstm_cell_fwd_f32": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _f, _p, C.c_uint64, _p]),
    "b200med_lstm_cell_bwd_f32": (C.c_int, [_p, _p, _p, _p, _i32, _p, _p, _i32, _p, _i64, _i32, _f, _p, C.c_uint64, _p]),
    "b200med_bn_ws_bytes": (_i64, [_i64, _i32]),
    "b200med_bn_fwd": (C.c_int, [_p, _i64, _i32, _p, _p, _f, _f, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "b200med_bn_bwd": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p, _i32, _p, _p, _p, _p, _p]),
    "b200med_tail_supported": (_i32, [_i32, _i32, _i32]),
    "b200med_tail_slabs": (_i64, [_i64]),
    "b200med_tail_fwd_hidden": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _p, _p]),
    "b200med_tail_fwd_out": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _p]),
    "b200med_tail_bwd_out": (C.c_int, [_p, _i32, _p, _p, _i64, _i32, _p, _p, _p, _p]),
    "b200med_tail_bwd_hidden": (C.c_int, [_p, _p, _i32, _p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p,
                                          _p]),
    "b200med_pool_drop_fwd": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _f, _p, C.c_uint64, _p]),
    "b200med_pool_drop_bwd": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _f, _p, C.c_uint64, _p]),
    "b200med_conv_pack": (C.c_int, [_p, _p, _p, _i32, _i32, _p]),
    "b200med_conv_unpack_grad": (C.c_int, [_p, _p, _i32, _i32, _p]),
    "b200med_transpose_last2": (C.c_int, [_p, _p, _i64, _i32, _i32, _p]),
    "b200med_concat2": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _p]),
    "b200med_slice_cols": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p]),
    "b200med_take_rows": (C.c_int, [_p, _p, _p, _i64, _i32, _p]),
    "b200med_loss_ws_bytes": (_i64, [_i64]),
    "b200med_bce_logits": (C.c_int, [_p, _p, _i64, _f, _f, _p, _p, _p, _p, _p, _i32, _p, _p]),
    "b200med_ce_logits": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _f, _p, _p, _p, _p, _i32, _i32, _p, _i32,
                                    _i32, _p, _p]),
    "b200med_ce_frame": (C.c_int, [_p, _p, _i32, _i64, _f, _p, _p, _p, _p, _i32, _p, _p]),
    "b200med_tcn_slots": (_i32, [_i64]),
    "b200med_tcn_pack": (C.c_int, [_p, _i32, _p, _p]),
    "b200med_tcn_layer_fwd": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _p, _p, _f, C.c_uint64, _p, C.c_uint64, _p]),
    "b200med_tcn_layer_bwd_hidden": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i64, _i32, _i32, _p, _p, _f, C.c_uint64, _p,
                                               C.c_uint64, _p]),
    "b200med_tcn_layer_bwd_input": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _p, _p, _p]),
    "b200med_tcn_reduce_grads": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "b200med_tcn_out_fwd": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _p]),
    "b200med_tcn_out_bwd": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _p]),
    "b200med_tcn_softmax_fwd": (C.c_int, [_p, _p, _i64, _i32, _p]),
    "b200med_tcn_softmax_bwd": (C.c_int, [_p, _p, _p, _i64, _i32, _p]),
    "b200med_tcn_stage_fwd": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _i32, _p, _p, _i32, _i64, _i32, _p, _p, _p, C.c_uint64, _p,
                                        C.c_uint64, _i32, _p, _p, _p, _p, _p, _p]),
    "b200med_tcn_stage_bwd_ws_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "b200med_tcn_stage_bwd": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _i32, _i32, _i64, _i32, _p, _p, _p, C.c_uint64, _p,
                                        C.c_uint64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "b200med_tcn_stage_fwd_bf16": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _i32, _p, _p, _i32, _i64, _i32, _p, _p, _p, _p, _p, _p, _p,
                                             _p, _p]),
    "b200med_adam_advance": (C.c_int, [_p, _f, _f, _p]),
    "b200med_multi_copy_f32": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _f, _f, _p]),
    "b200med_adam_step": (C.c_int, [_p, _p, _p, _p, _i64, _p, _f, _f, _f, _f, _f, _p]),
    "b200med_window_vote": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p]),
    "b200med_soft_vote": (C.c_int, [_p, _p, _p, _i64, _p, _p, _i32, _p, _p]),
    "b200med_cascade": (C.c_int, [_p, _p, _i64, _p, _p]),
    "b200med_confusion": (C.c_int, [_p, _p, _i64, _i32, _p, _i32, _p]),
    "b200med_roc_auc_ws_bytes": (_i64, [_i64]),
    "b200med_roc_auc": (C.c_int, [_p, _p, _i64, _p, _p, _p, _p]),
    "b200med_peer_alloc": (C.c_int, [_i64, C.POINTER(C.c_void_p)]),
    "b200med_peer_free": (C.c_int, [_p]),
    "b200med_peer_export": (C.c_int, [_p, _p]),
    "b200med_peer_import": (C.c_int, [_p, C.POINTER(C.c_void_p)]),
    "b200med_peer_close": (C.c_int, [_p]),
    "b200med_peer_flag_bytes": (_i64, []),
    "b200med_peer_allreduce_f32": (C.c_int, [_p, _p, _i32, _i32, _i64, _p]),
}

_lib = None


class B200MedError(RuntimeError):
    pass


def load():
    """Load libb200med.so once; raise loudly if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200MedError(
            f"{LIB_PATH} not found: the CUDA extension has not been built. Run "
            "`python -m multimodal_error_detection_b200.build` (needs nvcc); there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    if os.environ.get("B200MED_PDL", "1") == "0":      # A/B switch of the programmatic dependent launches (scripts, bench)
        lib.b200med_set_pdl(0)
    _lib = lib
    return lib


def call(name: str, *args):
    """Call an int-returning entry point; map error codes to Python exceptions the way the
    reference surfaces them (ValueError for bad arguments, RuntimeError otherwise)."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.b200med_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{name}: {msg}")
        raise B200MedError(f"{name} failed ({rc}): {msg}")
    return rc


def launch_count() -> int:
    return int(load().b200med_launch_count())
