"""ctypes binding of the C ABI declared in include/jaadb200.h.

The product path has no CPU fallback: if the CUDA library is missing or cannot be
loaded this module raises, it never substitutes another decoder.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

# symbols include/jaadb200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "jaadb_abi_version", "jaadb_status_string", "jaadb_last_error", "jaadb_engine_create", "jaadb_engine_destroy",
    "jaadb_stream_open_asc", "jaadb_stream_open_asc_sbr", "jaadb_stream_open_adts", "jaadb_stream_close", "jaadb_stream_get_info", "jaadb_decode",
    "jaadb_batch_create", "jaadb_batch_pcm_bytes", "jaadb_batch_upload", "jaadb_batch_decode", "jaadb_batch_sync",
    "jaadb_batch_download", "jaadb_batch_timings", "jaadb_batch_destroy", "jaadb_batch_tap", "jaadb_batch_tap_sbr", "jaadb_batch_tap_ps",
    "jaadb_adts_index", "jaadb_adts_index_many", "jaadb_mp4_index", "jaadb_mp4_index_many", "jaadb_probe_sbr", "jaadb_probe_sbr_asc",
    "jaadb_frames_interleave", "jaadb_decode_containers",
]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_streams", C.c_uint32), ("pcm_format", C.c_int32), ("tns_mode", C.c_int32),
                ("flags", C.c_uint32), ("chunk_frames", C.c_uint32), ("sbr_tile_frames", C.c_uint32), ("k2_segment_frames", C.c_uint32)]


class FrameDesc(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("nbytes", C.c_uint32), ("stream_id", C.c_int32)]


class FrameResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("channels", C.c_uint16), ("sample_length", C.c_uint16), ("sample_rate", C.c_uint32),
                ("pcm_bytes", C.c_uint32)]


class StreamInfo(C.Structure):
    _fields_ = [("profile", C.c_int32), ("sf_index", C.c_int32), ("channel_config", C.c_int32), ("channels", C.c_int32),
                ("sample_rate", C.c_int32), ("sample_length", C.c_int32), ("sbr", C.c_int32), ("reserved", C.c_int32)]


class Timings(C.Structure):
    _fields_ = [("parse_ms", C.c_float), ("filterbank_ms", C.c_float), ("sbr_ms", C.c_float), ("total_ms", C.c_float),
                ("launches", C.c_uint32), ("reserved", C.c_uint32 * 3)]


class AdtsInfo(C.Structure):
    _fields_ = [("profile", C.c_int32), ("sf_index", C.c_int32), ("channel_config", C.c_int32), ("sample_rate", C.c_int32),
                ("n_frames", C.c_uint64)]


class Mp4Track(C.Structure):
    _fields_ = [("asc", C.c_uint8 * 64), ("asc_bytes", C.c_uint32), ("track_id", C.c_int32), ("timescale", C.c_uint32),
                ("channel_count", C.c_uint32), ("sample_size_bits", C.c_uint32), ("sample_rate", C.c_uint32),
                ("object_type", C.c_uint32), ("max_bitrate", C.c_uint32), ("avg_bitrate", C.c_uint32),
                ("reserved", C.c_uint32), ("duration", C.c_uint64), ("n_frames", C.c_uint64)]


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if os.environ.get("JAADB200_LIB"):
        path = os.environ["JAADB200_LIB"]   # tuning experiments only: another build of the same sources (tools/build_variants.sh)
    elif build_if_missing and _build.stale():
        path = _build.build()
    if not os.path.exists(path):
        raise RuntimeError("libjaadb200.so is missing (%s): build it with __graft_entry__.build(); "
                           "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    vp, u8p = C.c_void_p, C.c_void_p
    lib.jaadb_abi_version.restype = C.c_int
    lib.jaadb_status_string.restype = C.c_char_p
    lib.jaadb_status_string.argtypes = [C.c_int32]
    lib.jaadb_last_error.restype = C.c_char_p
    lib.jaadb_last_error.argtypes = [vp]
    lib.jaadb_engine_create.argtypes = [C.POINTER(Options), C.POINTER(vp)]
    lib.jaadb_engine_destroy.argtypes = [vp]
    lib.jaadb_engine_destroy.restype = None
    lib.jaadb_stream_open_asc.argtypes = [vp, u8p, C.c_uint32, C.POINTER(C.c_int32)]
    lib.jaadb_stream_open_asc_sbr.argtypes = [vp, u8p, C.c_uint32, C.c_int32, C.POINTER(C.c_int32)]
    lib.jaadb_stream_open_adts.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    lib.jaadb_stream_close.argtypes = [vp, C.c_int32]
    lib.jaadb_probe_sbr.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, u8p, C.c_uint32, C.POINTER(C.c_int32)]
    lib.jaadb_probe_sbr_asc.argtypes = [vp, u8p, C.c_uint32, u8p, C.c_uint32, C.POINTER(C.c_int32)]
    lib.jaadb_stream_get_info.argtypes = [vp, C.c_int32, C.POINTER(StreamInfo)]
    lib.jaadb_decode.argtypes = [vp, u8p, C.c_uint64, vp, C.c_uint32, vp, C.c_uint64, vp, vp]
    lib.jaadb_batch_create.argtypes = [vp, vp, C.c_uint32, C.c_uint64, vp, C.POINTER(vp)]
    lib.jaadb_batch_pcm_bytes.restype = C.c_uint64
    lib.jaadb_batch_pcm_bytes.argtypes = [vp]
    lib.jaadb_batch_upload.argtypes = [vp, u8p, C.c_uint64]
    lib.jaadb_batch_decode.argtypes = [vp]
    lib.jaadb_batch_sync.argtypes = [vp]
    lib.jaadb_batch_download.argtypes = [vp, vp, C.c_uint64, vp]
    lib.jaadb_batch_timings.argtypes = [vp, C.POINTER(Timings)]
    lib.jaadb_batch_destroy.argtypes = [vp]
    lib.jaadb_batch_destroy.restype = None
    lib.jaadb_batch_tap.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp]
    lib.jaadb_batch_tap_sbr.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint32]
    for fn in (lib.jaadb_adts_index, lib.jaadb_mp4_index):
        fn.restype = C.c_int64
        fn.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_int32, vp, C.c_uint64, vp]
    for fn in (lib.jaadb_adts_index_many, lib.jaadb_mp4_index_many):
        fn.restype = C.c_int64
        fn.argtypes = [u8p, vp, C.c_uint32, vp, vp, C.c_uint64, vp, vp, C.c_uint32]
    lib.jaadb_decode_containers.restype = C.c_int64
    lib.jaadb_decode_containers.argtypes = [vp, C.c_int32, u8p, vp, C.c_uint32, vp, vp, C.c_uint64, vp, C.c_uint64, vp, C.c_uint32]
    lib.jaadb_frames_interleave.restype = C.c_int64
    lib.jaadb_frames_interleave.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32]
    _lib = lib
    return lib
