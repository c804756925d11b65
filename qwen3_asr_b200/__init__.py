"""qwen3-asr_b200: B200-native (sm_100a) log-mel frontend + Qwen3-ASR audio-encoder backend.

A drop-in third encoder backend for jaaacki/qwen3-asr's encoder-backend hook (src/server.py, next to
the ONNX_ENCODER_PATH / TRT_ENCODER_PATH slots).  Hand-written CUDA behind a C ABI
(include/qasr_b200.h, libqasr_b200.so); this package is the thin Python host side.
"""

from ._lib import QasrError, load_library  # noqa: F401
from .encoder import B200AudioEncoder, EncoderOutput, make_config, sinusoid_table  # noqa: F401

from .pool import B200EncoderPool  # noqa: F401
from .prefrontend import B200PreFrontend  # noqa: F401

__all__ = ["B200AudioEncoder", "B200EncoderPool", "B200PreFrontend", "EncoderOutput", "QasrError", "load_library", "make_config", "sinusoid_table"]
