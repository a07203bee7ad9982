"""Test-only ctypes binding to an FFmpeg libavcodec that happens to ship inside the image (opencv-python-headless vendors
one): an AAC decoder written independently of JAAD, used to cross-check the oracle's float stages (tests/
test_oracle_vs_libavcodec_cpu.py).  Never imported by the product; every user skips when the library is absent.

Only the stable leading fields of AVPacket / AVFrame are touched:
  AVPacket { AVBufferRef *buf; int64_t pts, dts; uint8_t *data; int size; ... }
  AVFrame  { uint8_t *data[8]; int linesize[8]; uint8_t **extended_data; int width, height; int nb_samples; int format; ... }
The `aac` decoder takes ADTS frames as they are (header included) and returns planar float in [-1, 1).
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

_LIBDIR = os.path.join(os.path.dirname(np.__file__), "..", "opencv_python_headless.libs")
AV_CODEC_ID_AAC = 86018
AVERROR_EAGAIN = -11
_libs = None


def _load():
    global _libs
    if _libs is not None:
        return _libs
    d = os.path.abspath(_LIBDIR)

    def one(pat, required=True):
        hits = sorted(glob.glob(os.path.join(d, pat)))
        if not hits:
            if required:
                raise OSError("no " + pat)
            return None
        return C.CDLL(hits[0], mode=C.RTLD_GLOBAL)

    try:
        for dep in ("libdrm*", "libcrypto*", "libssl*"):
            one(dep, required=False)
        util = one("libavutil*")
        for dep in ("libswresample*", "libaom*", "libvpx*"):
            one(dep, required=False)
        codec = one("libavcodec*")
    except OSError:
        _libs = False
        return _libs
    codec.avcodec_find_decoder.restype = C.c_void_p
    codec.avcodec_find_decoder.argtypes = [C.c_int]
    codec.avcodec_alloc_context3.restype = C.c_void_p
    codec.avcodec_alloc_context3.argtypes = [C.c_void_p]
    codec.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    codec.avcodec_free_context.argtypes = [C.POINTER(C.c_void_p)]
    codec.av_packet_alloc.restype = C.c_void_p
    codec.av_packet_free.argtypes = [C.POINTER(C.c_void_p)]
    codec.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    codec.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    util.av_frame_alloc.restype = C.c_void_p
    util.av_frame_free.argtypes = [C.POINTER(C.c_void_p)]
    util.av_frame_unref.argtypes = [C.c_void_p]
    util.av_log_set_level.argtypes = [C.c_int]
    util.av_log_set_level(-8)   # quiet
    _libs = (codec, util)
    return _libs


def available() -> bool:
    l = _load()
    return bool(l) and bool(l[0].avcodec_find_decoder(AV_CODEC_ID_AAC))


class _Packet(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("data", C.c_void_p), ("size", C.c_int)]


class _Frame(C.Structure):
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p), ("width", C.c_int),
                ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int)]


class AacDecoder:
    """libavcodec's native `aac` decoder fed with whole ADTS frames; decode() returns float32 [channels, samples] scaled to
    JAAD's +-32768 range, or None while the decoder has nothing to give."""

    def __init__(self, channels: int):
        codec, util = _load()
        self._c, self._u = codec, util
        self.channels = channels
        self._ctx = C.c_void_p(codec.avcodec_alloc_context3(codec.avcodec_find_decoder(AV_CODEC_ID_AAC)))
        if codec.avcodec_open2(self._ctx, None, None) < 0:
            raise RuntimeError("avcodec_open2 failed")
        self._pkt = C.c_void_p(codec.av_packet_alloc())
        self._frm = C.c_void_p(util.av_frame_alloc())

    def decode(self, adts_frame: np.ndarray):
        buf = np.zeros(len(adts_frame) + 64, np.uint8)   # AV_INPUT_BUFFER_PADDING_SIZE
        buf[: len(adts_frame)] = adts_frame
        pkt = _Packet.from_address(self._pkt.value)
        pkt.data, pkt.size = buf.ctypes.data, len(adts_frame)
        rc = self._c.avcodec_send_packet(self._ctx, self._pkt)
        pkt.data, pkt.size = None, 0
        if rc < 0:
            raise RuntimeError("avcodec_send_packet: %d" % rc)
        rc = self._c.avcodec_receive_frame(self._ctx, self._frm)
        if rc == AVERROR_EAGAIN:
            return None
        if rc < 0:
            raise RuntimeError("avcodec_receive_frame: %d" % rc)
        fr = _Frame.from_address(self._frm.value)
        assert fr.format == 8, "expected planar float (AV_SAMPLE_FMT_FLTP), got %d" % fr.format
        n = fr.nb_samples
        out = np.stack([np.ctypeslib.as_array(C.cast(fr.data[c], C.POINTER(C.c_float)), shape=(n,)).copy() for c in range(self.channels)])
        self._u.av_frame_unref(self._frm)
        return out * np.float32(32768.0)

    def close(self):
        if getattr(self, "_ctx", None):
            self._u.av_frame_free(C.byref(self._frm))
            self._c.av_packet_free(C.byref(self._pkt))
            self._c.avcodec_free_context(C.byref(self._ctx))
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
