"""ctypes binding of the C ABI declared in include/jen1_b200.h.

There is no fallback: if the shared library is missing the import of the product path fails with an
explicit error (build it with `python -m jen1_b200.build` or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JEN1_B200_LIB: A/B a variant build (scripts/build_variant.py) without touching the in-tree library
LIB_PATH = os.environ.get("JEN1_B200_LIB") or os.path.join(_HERE, "_C", "libjen1_b200.so")
MAX_LEVELS = 16

DTYPE_F32, DTYPE_BF16 = 0, 1
OBJECTIVES = {"noise": 0, "x0": 1, "v": 2}


class Jen1ModelDesc(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("channels", C.c_int32), ("num_layers", C.c_int32),
        ("multipliers", C.c_int32 * (MAX_LEVELS + 1)), ("factors", C.c_int32 * MAX_LEVELS),
        ("num_blocks", C.c_int32 * MAX_LEVELS), ("attentions", C.c_int32 * (MAX_LEVELS + 1)),
        ("resnet_groups", C.c_int32), ("context_channels", C.c_int32), ("context_features_multiplier", C.c_int32),
        ("context_embedding_features", C.c_int32), ("context_embedding_max_length", C.c_int32),
        ("attention_heads", C.c_int32), ("attention_multiplier", C.c_int32), ("use_skip_scale", C.c_int32),
    ]


class Jen1CodecDesc(C.Structure):
    _fields_ = [
        ("channels", C.c_int32), ("dimension", C.c_int32), ("n_filters", C.c_int32), ("n_ratios", C.c_int32),
        ("ratios", C.c_int32 * 8), ("kernel_size", C.c_int32), ("last_kernel_size", C.c_int32),
        ("residual_kernel_size", C.c_int32), ("compress", C.c_int32), ("lstm_layers", C.c_int32), ("eps", C.c_float),
    ]


# every symbol include/jen1_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "jen1_engine_create": (C.c_int, [C.POINTER(Jen1ModelDesc), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "jen1_engine_destroy": (None, [C.c_void_p]),
    "jen1_last_error": (C.c_char_p, [C.c_void_p]),
    "jen1_engine_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "jen1_engine_finalize": (C.c_int, [C.c_void_p]),
    "jen1_engine_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "jen1_engine_reserve": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "jen1_engine_set_context": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "jen1_engine_set_timesteps": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "jen1_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_int,
                                    C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "jen1_sample_begin": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p]),
    "jen1_sample_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jen1_attention_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p]),
    "jen1_engine_launch_count": (C.c_int64, [C.c_void_p]),
    "jen1_engine_weight_bytes": (C.c_int64, [C.c_void_p]),
    "jen1_engine_umma_launch_count": (C.c_int64, [C.c_void_p]),
    "jen1_engine_umma_attn_launch_count": (C.c_int64, [C.c_void_p]),
    "jen1_engine_fused_transformer_launch_count": (C.c_int64, [C.c_void_p]),
    "jen1_engine_debug_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "jen1_codec_create": (C.c_int, [C.POINTER(Jen1CodecDesc), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "jen1_codec_create_encoder": (C.c_int, [C.POINTER(Jen1CodecDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "jen1_codec_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "jen1_codec_quantize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "jen1_codec_destroy": (None, [C.c_void_p]),
    "jen1_codec_last_error": (C.c_char_p, [C.c_void_p]),
    "jen1_codec_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "jen1_codec_finalize": (C.c_int, [C.c_void_p]),
    "jen1_codec_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "jen1_codec_reserve": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "jen1_codec_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "jen1_codec_launch_count": (C.c_int64, [C.c_void_p]),
    "jen1_codec_weight_bytes": (C.c_int64, [C.c_void_p]),
    "jen1_codec_tf32_launch_count": (C.c_int64, [C.c_void_p]),
    "jen1_codec_lstm_tc_launch_count": (C.c_int64, [C.c_void_p]),
    "jen1_codec_hop": (C.c_int, [C.c_void_p]),
    "jen1_codec_lstm_cluster": (C.c_int, [C.c_void_p]),
}

_lib = None


def load():
    """Load the engine library (once) and attach the prototypes; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "jen1_b200: CUDA engine library not found at %s -- build it with `python -m jen1_b200.build`. "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
