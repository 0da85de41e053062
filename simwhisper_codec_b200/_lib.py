"""ctypes binding of libswc.so (C ABI declared in include/swc.h).

The library is built in-tree by `make -C simwhisper_codec_b200/csrc` (or `__graft_entry__.build()`).
There is no fallback: if the shared object is missing, importing a compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SWC_LIB_PATH: another build of the same library (development A/B runs); the product path is the in-tree libswc.so
LIB_PATH = os.environ.get("SWC_LIB_PATH") or os.path.join(_HERE, "libswc.so")

PRECISION = {"fp32": 0, "bf16": 1, "bf16x3": 2}
KCLASS = ("gemm_tcgen05", "gemm_simt_fp32", "attention", "layernorm", "dwconv7_ln", "aa_snake", "other")
STAGE = {"mel": 0, "encoder": 1, "downsample": 2, "quantizer": 3, "upsample": 4, "decoder": 5, "vocos": 6,
         "tokenize": 7, "detokenize": 8, "forward": 9}

_p, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t

# name -> (restype, argtypes); every symbol include/swc.h declares
SIGNATURES = {
    "swc_version": (_i, []),
    "swc_last_error": (C.c_char_p, []),
    "swc_model_create": (_i, [C.POINTER(_p), _i]),
    "swc_model_set_tensor": (_i, [_p, C.c_char_p, _p, _i, C.POINTER(_i64), _i]),
    "swc_model_pack": (_i, [_p]),
    "swc_model_packed_numel": (_i64, [_p, C.c_char_p]),
    "swc_model_get_packed": (_i, [_p, C.c_char_p, _p, _i64]),
    "swc_model_finalize": (_i, [_p, _i]),
    "swc_model_destroy": (None, [_p]),
    "swc_workspace_bytes": (_sz, [_p, _i, _i, _i]),
    "swc_mel": (_i, [_p, _p, _i64, _i, _p, _i, _p, _p, _p, _sz, _p]),
    "swc_encoder": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _sz, _p]),
    "swc_encoder_hidden": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "swc_downsample": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _sz, _p]),
    "swc_quantize": (_i, [_p, _p, _p, _i, _i, _p, _p, _p]),
    "swc_dequantize": (_i, [_p, _p, _i, _p, _i, _i, _p, _p]),
    "swc_upsample": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _sz, _p]),
    "swc_decoder": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _sz, _p]),
    "swc_vocos": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _sz, _p]),
    "swc_tokenize": (_i, [_p, _p, _i64, _i, _p, _i, _p, _p, _p, _p, _sz, _p]),
    "swc_detokenize": (_i, [_p, _p, _i, _p, _i, _i, _p, _p, _p, _sz, _p]),
    "swc_max_ragged": (_i, []),
    "swc_tokenize_ragged": (_i, [_p, _p, _i64, _i, _p, C.POINTER(_i64), _i, _p, _p, _p, _p, _sz, _p]),
    "swc_detokenize_ragged": (_i, [_p, _p, _i, _p, C.POINTER(_i64), _i, _i, _p, _p, _p, _sz, _p]),
    "swc_forward": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "swc_profile": (None, [_i]),
    "swc_profile_read": (_i, [C.POINTER(C.c_double), C.POINTER(_i64), _i]),
    "swc_test_gemm": (_i, [_i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "swc_set_gemm_variant": (None, [_i]),
    "swc_test_attention": (_i, [_i, _p, _p, _p, _i, _i, _i, _p]),
    "swc_debug_attn_trace": (_i, [_i, _p, _i]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(the SimWhisper-Codec B200 path has no Python/CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().swc_last_error().decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {last_error()}")
