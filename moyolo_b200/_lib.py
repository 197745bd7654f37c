"""ctypes binding of libmoyolo_b200.so (the C ABI declared in include/moyolo_b200.h).

There is no CPU or PyTorch fallback: if the shared object is missing the import of any op fails
loudly with instructions to build it.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libmoyolo_b200.so"
_lib = None

# status codes / enums (mirror include/moyolo_b200.h)
OK, ERR_BAD_ARG, ERR_BAD_SHAPE, ERR_UNSUPPORTED, ERR_CUDA, ERR_ALIGNMENT = range(6)
F32, BF16, F64 = 0, 1, 2
SOFTMAX, SOFTMAX_PLUS1 = 0, 1
EPI_NONE, EPI_RELU = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2

_p, _i, _l, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

class FrameSubmit(C.Structure):
    """moyolo_frame_submit_t (include/moyolo_b200.h)."""
    _fields_ = [("copy_stream", _p), ("main_stream", _p), ("main_stream_valid", _i), ("sync_inputs", _i),
                ("ev_slot_free", _p), ("ev_scratch", _p), ("ev_copy", _p), ("ev_done", _p), ("graph_exec", _p),
                ("n_inputs", _i), ("n_outputs", _i),
                ("in_src", _p * 4), ("in_dst", _p * 4), ("in_bytes", _l * 4),
                ("out_src", _p * 4), ("out_dst", _p * 4), ("out_bytes", _l * 4),
                ("out_stream", _p), ("out_stream_valid", _i), ("ev_graph", _p),
                ("vp_stream", _p), ("vp_valid", _i), ("ev_tail_prev", _p), ("ev_vp", _p),
                ("vp_x", _p), ("vp_ldx", _l), ("vp_w", _p), ("vp_bias", _p), ("vp_y", _p), ("vp_ldy", _l),
                ("vp_M", _l), ("vp_N", _i), ("vp_max_ctas", _i), ("pre_graph_exec", _p)]


class DecoderLayerWeights(C.Structure):
    """moyolo_decoder_layer_weights_t (include/moyolo_b200.h)."""
    _fields_ = [(n, _p) for n in ("wqkv", "wo", "woff", "wout", "w1", "w2", "wb1", "wb2", "bqkv", "bo", "boff", "bout",
                                  "b1", "b2", "bb1", "bb2", "wb3", "bb3", "ln1_w", "ln1_b", "ln2_w", "ln2_b", "ln3_w",
                                  "ln3_b")]


class DecoderCluster(C.Structure):
    """moyolo_decoder_cluster_t (include/moyolo_b200.h)."""
    _fields_ = [("n_layers", _i), ("d_model", _i), ("n_heads", _i), ("d_ffn", _i), ("n_levels", _i), ("n_points", _i),
                ("layers", DecoderLayerWeights * 8),
                ("x_in", _p), ("pos", _p), ("refer0", _p), ("x_out", _p), ("x_lp_out", _p), ("refer_out", _p * 8),
                ("kv", _p), ("values", _p), ("value_batch_stride", _l), ("value_pos_stride", _l),
                ("value_shapes", C.c_int32 * 16), ("softmax_mode", _i), ("row_offsets", _p), ("n_seq", _i),
                ("rows_pad", _l), ("rows_per_tile", _i), ("grid_barrier", _p), ("reset_barrier", _i), ("status", _p),
                ("score_w", _p), ("score_b", _p), ("nc", _i), ("logits", _p), ("scores", _p), ("labels", _p),
                ("eps", _f), ("profile", _p)]


# name -> (restype, argtypes); must list every symbol include/moyolo_b200.h declares
SIGNATURES = {
    "moyolo_version": (_i, []),
    "moyolo_last_error": (C.c_char_p, []),
    "moyolo_device_supported": (_i, []),
    "moyolo_launch_count": (C.c_uint64, []),
    "moyolo_decoder_cluster_forward": (_i, [_p, _p]),
    "moyolo_decoder_cluster_limits": (_i, [_i, _p, _p]),
    "moyolo_table_pack": (_i, [_p, _l, _p, _p, _l, _p, _p]),
    "moyolo_table_merge": (_i, [_p, _i, _l, _p, _p, _p]),
    "moyolo_split_bf16x3": (_i, [_p, _l, _p, _l, _l, _i, _i, _p]),
    "moyolo_fsqm_update": (_i, [_i, _i, _f, _f, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _i, _p, _p, _p, _p]),
    "moyolo_msda_sampled_forward": (_i, [_p, _i, _l, _l, _p, _i, _i, _l, _i, _i, _i, _p, _p, _i, _l, _p, _p, _l, _p]),
    "moyolo_msda_sampled_backward": (_i, [_p, _i, _l, _l, _p, _i, _i, _l, _i, _i, _i, _p, _p, _i, _p, _l, _l, _p, _p,
                                          _p, _p, _p]),
    "moyolo_msda_fused_forward": (_i, [_p, _i, _l, _l, _p, _i, _i, _l, _i, _i, _i, _p, _l, _p, _l, _p, _i, _i, _i,
                                       _l, _p, _p, _l, _p]),
    "moyolo_msda_fused_forward_headmajor": (_i, [_p, _i, _l, _l, _l, _p, _i, _i, _l, _i, _i, _i, _p, _l, _p, _l, _p, _i, _i,
                                                 _i, _l, _p, _p, _l, _p]),
    "moyolo_msda_proj_fused_forward": (_i, [_p, _i, _l, _l, _p, _i, _i, _l, _i, _i, _i, _p, _l, _p, _p, _p, _i, _i, _i,
                                            _l, _p, _p, _l, _p]),
    "moyolo_linear": (_i, [_p, _l, _p, _p, _p, _l, _l, _i, _i, _i, _i, _i, _p, _i, _p]),
    "moyolo_linear_tall": (_i, [_p, _l, _p, _p, _p, _l, _l, _i, _i, _p, _i, _p]),
    "moyolo_linear_dual": (_i, [_p, _l, _p, _l, _i, _p, _p, _p, _l, _l, _i, _i, _i, _p]),
    "moyolo_linear_add_layernorm": (_i, [_p, _l, _p, _p, _p, _p, _p, _f, _l, _i, _i, _p, _p, _p, _p, _p]),
    "moyolo_linear_add_layernorm_scores": (_i, [_p, _l, _p, _p, _p, _p, _p, _f, _l, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p,
                                                _p]),
    "moyolo_ffn_add_layernorm": (_i, [_p, _l, _p, _p, _p, _p, _p, _i, _p, _p, _p, _f, _l, _i, _p, _p, _p, _p, _p]),
    "moyolo_self_attention": (_i, [_p, _l, _p, _l, _p, _l, _p, _l, _i, _i, _p, _p, _p, _i, _i, _p, _p]),
    "moyolo_add_layernorm": (_i, [_p, _p, _p, _p, _f, _l, _i, _p, _p, _p, _p, _i, _p]),
    "moyolo_add_cast": (_i, [_p, _p, _p, _i, _l, _p]),
    "moyolo_box_refine": (_i, [_p, _l, _i, _p, _p, _p, _p, _l, _i, _p]),
    "moyolo_score_head": (_i, [_p, _l, _i, _p, _p, _p, _p, _p, _l, _i, _i, _p, _p]),
    "moyolo_enc_output_scores": (_i, [_p, _l, _p, _p, _p, _p, _f, _p, _p, _p, _i, _l, _p, _p, _p, _p, _p]),
    "moyolo_topk": (_i, [_p, _l, _i, _i, _i, _p, _p, _p]),
    "moyolo_select_gather": (_i, [_p, _p, _p, _i, _i, _l, _i, _i, _p, _p, _i, _p, _p]),
    "moyolo_anchor_box": (_i, [_p, _l, _i, _p, _p, _p, _p, _i, _l, _f, _f, _p, _l, _i, _p]),
    "moyolo_anchor_invalid": (_i, [_p, _i, _l, _f, _f, _p, _p]),
    "moyolo_mask_rows": (_i, [_p, _p, _l, _p, _l, _i, _p]),
    "moyolo_sigmoid": (_i, [_p, _p, _l, _p]),
    "moyolo_inverse_sigmoid": (_i, [_p, _p, _l, _p]),
    "moyolo_pos2posemb": (_i, [_p, _p, _l, _i, _i, _f, _p]),
    "moyolo_linear_k4_relu": (_i, [_p, _p, _p, _p, _i, _l, _i, _p]),
    "moyolo_track_workspace_bytes": (_l, [_l]),
    "moyolo_track_assign": (_i, [_p, _p, _p, _p, _p, _l, _f, _f, _i, _f, _p, _p]),
    "moyolo_track_compact": (_i, [_p, _l, _p, _p, _p, _p, _p, _i, _p]),
    "moyolo_track_assign_batched": (_i, [_p, _p, _p, _p, _p, _p, _i, _l, _f, _f, _i, _f, _p, _p, _p]),
    "moyolo_frame_assemble": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _l,
                                   _i, _f, _p, _p, _p, _p, _i, _p]),
    "moyolo_frame_compact": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                  _p, _p, _i, _i, _f, _p]),
    "moyolo_track_suppress_batched": (_i, [_p, _p, _p, _p, _i, _l, _f, _p, _p, _p]),
    "moyolo_frame_assign_compact": (_i, [_i, _i, _i, _l, _p, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p, _p, _p, _p, _p,
                                         _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _f, _p]),
    "moyolo_frame_writeback": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "moyolo_frame_submit": (_i, [C.POINTER(FrameSubmit)]),
    "moyolo_event_create": (_p, []),
    "moyolo_event_destroy": (_i, [_p]),
    "moyolo_event_record": (_i, [_p, _p]),
    "moyolo_stream_wait_event": (_i, [_p, _p]),
    "moyolo_frame_emit": (_i, [_i, _l, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _p, _p]),
}


def lib_path() -> Path:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing: the CUDA extension is the only implementation of this path "
                "(no CPU/PyTorch fallback). Build it with `python -m moyolo_b200.build`.")
        handle = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().moyolo_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a C status to the exception type the reference raises for the same condition."""
    if rc == OK:
        return
    msg = last_error()
    if rc in (ERR_BAD_ARG, ERR_BAD_SHAPE, ERR_ALIGNMENT):
        raise ValueError(msg)
    raise RuntimeError(f"moyolo_b200 status {rc}: {msg}")
