"""Container indexers: ADTS streams and MP4 files -> frame tables for Engine.decode.

Host-side mirror of the two demultiplexers JAAD's Main feeds the decoder from
(src/main/java/net/sourceforge/jaad/Main.java:52-111):
  adts/ADTSDemultiplexer.java:26-74   -> adts_index / adts_index_many
  mp4/.../api/Track.java:90-172       -> mp4_index / mp4_index_many
Both run in the native library (csrc/container_index.cpp); nothing is parsed in Python.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import FRAME_DESC_DTYPE, EngineError


def _u8(data) -> np.ndarray:
    a = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, np.uint8)
    return a


def _one(fn, info_cls, data, stream_id, blob_offset):
    lib = _lib.load()
    a = _u8(data)
    info = info_cls()
    n = getattr(lib, fn)(a.ctypes.data, a.nbytes, blob_offset, stream_id, None, 0, C.byref(info))
    if n < 0:
        raise EngineError("%s failed: %d" % (fn, n))
    frames = np.zeros(n, FRAME_DESC_DTYPE)
    if n:
        m = getattr(lib, fn)(a.ctypes.data, a.nbytes, blob_offset, stream_id, frames.ctypes.data, n, None)
        assert m == n
    return frames, info


def adts_index(data, stream_id: int = 0, blob_offset: int = 0):
    """Frame table of one ADTS stream and the header fields of its first frame (ADTSDemultiplexer)."""
    return _one("jaadb_adts_index", _lib.AdtsInfo, data, stream_id, blob_offset)


def mp4_index(data, stream_id: int = 0, blob_offset: int = 0):
    """Frame table (decoding-time order) and track description of the first AAC track of an MP4 file."""
    frames, t = _one("jaadb_mp4_index", _lib.Mp4Track, data, stream_id, blob_offset)
    return frames, t


def asc_of(track: _lib.Mp4Track) -> bytes:
    return bytes(track.asc[: track.asc_bytes])


def _many(fn, info_cls, blob, begin, stream_ids, threads, out=None):
    lib = _lib.load()
    a = _u8(blob)
    begin = np.ascontiguousarray(begin, np.uint64)
    n_streams = len(begin) - 1
    ids = None if stream_ids is None else np.ascontiguousarray(stream_ids, np.int32)
    first = np.zeros(n_streams + 1, np.uint64)
    infos = (info_cls * max(n_streams, 1))()
    f = getattr(lib, fn)
    idp = None if ids is None else ids.ctypes.data
    if out is not None:
        # a caller-owned table (steady-state use: the frame count of a batch is known): one call counts and fills
        total = f(a.ctypes.data, begin.ctypes.data, n_streams, idp, out.ctypes.data, len(out), first.ctypes.data, infos, threads)
        if total < 0 or total > len(out):
            raise EngineError("%s failed: %d (table holds %d rows)" % (fn, total, len(out)))
        return out[:total], first.astype(np.int64), infos
    total = f(a.ctypes.data, begin.ctypes.data, n_streams, idp, None, 0, first.ctypes.data, infos, threads)
    if total < 0:
        raise EngineError("%s failed: %d" % (fn, total))
    frames = np.zeros(total, FRAME_DESC_DTYPE)
    if total:
        f(a.ctypes.data, begin.ctypes.data, n_streams, idp, frames.ctypes.data, total, first.ctypes.data, infos, threads)
    return frames, first.astype(np.int64), list(infos)[:n_streams]


def adts_index_many(blob, stream_begin, stream_ids=None, threads: int = 0, out=None):
    """Index n ADTS streams stored back to back in `blob` (stream s = blob[begin[s]:begin[s+1]]) on host threads.

    Returns (frames stream-major, first_frame[n+1], infos).  `out`: a FRAME_DESC_DTYPE array to fill instead of a new one."""
    return _many("jaadb_adts_index_many", _lib.AdtsInfo, blob, stream_begin, stream_ids, threads, out)


def mp4_index_many(blob, file_begin, stream_ids=None, threads: int = 0, out=None):
    return _many("jaadb_mp4_index_many", _lib.Mp4Track, blob, file_begin, stream_ids, threads, out)


def interleave(frames: np.ndarray, first_frame: np.ndarray, out: np.ndarray | None = None, threads: int = 0) -> np.ndarray:
    """Reorder a stream-major frame table frame-major (frame 0 of every stream, frame 1 of every stream ...), the
    order a live batch of concurrent streams arrives in.  Per-stream order is preserved, so both decode identically."""
    lib = _lib.load()
    frames = np.ascontiguousarray(frames, FRAME_DESC_DTYPE)
    first = np.ascontiguousarray(first_frame, np.uint64)
    out = np.empty(len(frames), FRAME_DESC_DTYPE) if out is None else out
    n = lib.jaadb_frames_interleave(frames.ctypes.data, first.ctypes.data, len(first) - 1, out.ctypes.data, threads)
    if n != len(frames):
        raise EngineError("jaadb_frames_interleave failed: %d" % n)
    return out
