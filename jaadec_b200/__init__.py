"""jaadec_b200 -- B200-native batched AAC decode engine behind JAAD's per-frame decode API.

The CUDA library (csrc/, built in-tree as _build/libjaadb200.so) is the product;
this package is the host-side mirror of the reference interface plus a ctypes
binding of the C ABI in include/jaadb200.h.  There is no CPU decode path here.
"""
from .engine import (Engine, Batch, EngineError, PCM_S16LE, PCM_S16BE, PCM_F32_PLANAR, FLAG_PROFILE, FLAG_DEBUG_TAPS, FLAG_PULSE_ISO,
                     FRAME_DESC_DTYPE, FRAME_RESULT_DTYPE, TNS_JAAD, TNS_ISO, CONTAINER_ADTS, CONTAINER_MP4)

__all__ = ["Engine", "Batch", "EngineError", "PCM_S16LE", "PCM_S16BE", "PCM_F32_PLANAR", "FLAG_PROFILE",
           "FLAG_DEBUG_TAPS", "FLAG_PULSE_ISO", "FRAME_DESC_DTYPE", "FRAME_RESULT_DTYPE", "TNS_JAAD", "TNS_ISO", "CONTAINER_ADTS", "CONTAINER_MP4"]
