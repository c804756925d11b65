"""ctypes binding of libqasr_b200.so (include/qasr_b200.h).  No fallback: if the library is missing
or cannot be loaded this raises, it never routes to a CPU or PyTorch implementation."""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QASR_B200_LIB") or os.path.join(HERE, "libqasr_b200.so")

QASR_F32, QASR_BF16, QASR_F16 = 0, 1, 2
ABI_VERSION = 1


class QasrConfig(C.Structure):
    _fields_ = [
        ("d_model", C.c_int32),
        ("encoder_layers", C.c_int32),
        ("encoder_attention_heads", C.c_int32),
        ("encoder_ffn_dim", C.c_int32),
        ("output_dim", C.c_int32),
        ("n_window", C.c_int32),
        ("n_window_infer", C.c_int32),
        ("downsample_hidden_size", C.c_int32),
        ("num_mel_bins", C.c_int32),
        ("max_source_positions", C.c_int32),
        ("max_chunks", C.c_int32),
        ("max_tokens", C.c_int32),
        ("flags", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/qasr_b200.h declares
_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
SIGNATURES = {
    "qasr_abi_version": (C.c_int, []),
    "qasr_last_error": (C.c_char_p, []),
    "qasr_create": (C.c_int, [C.POINTER(QasrConfig), C.c_int, C.POINTER(_P)]),
    "qasr_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int, _I64P, C.c_int]),
    "qasr_finalize": (C.c_int, [_P]),
    "qasr_workspace_bytes": (C.c_size_t, [_P]),
    "qasr_token_len": (C.c_int64, [C.c_int64]),
    "qasr_logmel": (C.c_int, [_P, _P, _I64P, C.c_int, _P, C.c_int64, _I64P, _P]),
    "qasr_encode": (C.c_int, [_P, _P, C.c_int, C.c_int64, _I64P, C.c_int, _P, _I64P, _P]),
    "qasr_encode_scatter": (C.c_int, [_P, _P, C.c_int, C.c_int64, _I64P, C.c_int, _P, C.c_int64, _P, _I64P, _P]),
    "qasr_encode_pcm": (C.c_int, [_P, _P, _I64P, C.c_int, _P, _I64P, _P]),
    "qasr_encode_pcm_host": (C.c_int, [_P, _P, _I64P, C.c_int, _P, C.c_int64, _I64P, _P]),
    "qasr_submit_pcm_host": (C.c_int, [_P, _P, _I64P, C.c_int, _P, C.c_int64, _I64P, _P, C.POINTER(C.c_uint64)]),
    "qasr_wait": (C.c_int, [_P, C.c_uint64]),
    "qasr_poll": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_int)]),
    "qasr_pipe_times": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "qasr_submit_clips_host": (C.c_int, [_P, _P, _I64P, _I64P, _I64P, C.c_int, _P, C.c_int64, _I64P, _P, C.POINTER(C.c_uint64)]),
    "qasr_logmel_host": (C.c_int, [_P, _P, _I64P, C.c_int, _P, _I64P, _P]),
    "qasr_resample_len": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "qasr_resample_pcm16": (C.c_int, [_P, _P, _I64P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, _P, C.c_int64, _I64P, _P]),
    "qasr_resample_f32_len": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "qasr_resample_f32": (C.c_int, [_P, _P, _I64P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), _P, C.c_int64, _I64P, _P]),
    "qasr_resample_f32_taps": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int)]),
    "qasr_split_audio": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double, _I64P, C.c_int, C.POINTER(C.c_int), _P]),
    "qasr_ws_window": (C.c_int, [_P, _P, _I64P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_int, C.c_int, _P, C.c_int64,
                                 _I64P, _P]),
    "qasr_destroy": (None, [_P]),
    "qasr_pool_create": (C.c_int, [C.POINTER(QasrConfig), C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "qasr_pool_size": (C.c_int, [_P]),
    "qasr_pool_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int, _I64P, C.c_int]),
    "qasr_pool_finalize": (C.c_int, [_P]),
    "qasr_pool_workspace_bytes": (C.c_size_t, [_P]),
    "qasr_pool_submit": (C.c_int, [_P, _P, _I64P, C.c_int, _P, C.c_int64, _I64P, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "qasr_pool_collect": (C.c_int, [_P, C.c_uint64]),
    "qasr_gelu_table": (C.c_int, [C.POINTER(C.c_float), C.c_int]),
    "qasr_mel_plan": (C.c_int64, [_I64P, C.c_int, C.c_int, _I64P, C.c_int64]),
    "qasr_pool_plan": (C.c_int, [_I64P, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "qasr_pool_stats": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int, C.c_int]),
    "qasr_pool_set_sharding": (C.c_int, [_P, C.c_int]),
    "qasr_pool_plan_mode": (C.c_int, [_I64P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "qasr_pool_destroy": (None, [_P]),
    "qasr_launch_count": (C.c_uint64, [_P]),
    "qasr_graph_stats": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)]),
    "qasr_profile_enable": (C.c_int, [_P, C.c_int]),
    "qasr_profile_read": (C.c_int, [_P, C.c_char_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int)]),
    "qasr_debug_read": (C.c_int, [_P, C.c_char_p, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "qasr_debug_gemm": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "qasr_debug_quant_fp8": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int, _P]),
    "qasr_debug_gemm_fp8": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "qasr_debug_layernorm": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "qasr_debug_attention": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, _P]),
    "qasr_debug_attention_tc": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, C.c_int, _P]),
}

_lib = None


class QasrError(RuntimeError):
    pass


def load_library(path: str | None = None) -> C.CDLL:
    """Load the C-ABI library and bind every declared symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise QasrError(
            f"{p} not found: build it with `python -m qwen3_asr_b200.build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for this backend."
        )
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.qasr_abi_version()
    if got != ABI_VERSION:
        raise QasrError(f"{p}: ABI version {got}, expected {ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib


def check(lib: C.CDLL, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.qasr_last_error()
        raise QasrError(f"{what} failed ({rc}): {msg.decode(errors='replace') if msg else '?'}")
