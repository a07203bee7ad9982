"""Engine / Batch: thin Python objects over the C ABI (include/jaadb200.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FrameDesc, FrameResult, Options, StreamInfo, Timings

PCM_S16LE, PCM_S16BE, PCM_F32_PLANAR = 0, 1, 2
FLAG_PROFILE, FLAG_DEBUG_TAPS, FLAG_PULSE_ISO = 1, 2, 4
TNS_JAAD, TNS_ISO = 0, 1
CONTAINER_ADTS, CONTAINER_MP4 = 0, 1

FRAME_DESC_DTYPE = np.dtype([("offset", "<u8"), ("nbytes", "<u4"), ("stream_id", "<i4")])
FRAME_RESULT_DTYPE = np.dtype([("status", "<i4"), ("channels", "<u2"), ("sample_length", "<u2"), ("sample_rate", "<u4"),
                               ("pcm_bytes", "<u4")])


class EngineError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data


class Engine:
    def __init__(self, device: int = 0, max_streams: int = 4096, pcm_format: int = PCM_S16LE, flags: int = 0,
                 chunk_frames: int = 0, sbr_tile_frames: int = 0, tns_mode: int = TNS_JAAD, k2_segment_frames: int = 0):
        self._lib = _lib.load()
        opts = Options(device, max_streams, pcm_format, tns_mode, flags, chunk_frames, sbr_tile_frames, k2_segment_frames)
        h = C.c_void_p()
        rc = self._lib.jaadb_engine_create(C.byref(opts), C.byref(h))
        if rc != 0:
            raise EngineError("jaadb_engine_create failed: %d (no CUDA device? there is no CPU fallback)" % rc)
        self._h = h
        self.pcm_format = pcm_format
        self.flags = flags

    def close(self):
        if getattr(self, "_h", None):
            self._lib.jaadb_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError("%s failed: %d (%s)" % (what, rc, self._lib.jaadb_last_error(self._h).decode()))

    def open_adts(self, profile: int, sf_index: int, channel_config: int, expect_sbr: int = 0) -> int:
        sid = C.c_int32(-1)
        self._check(self._lib.jaadb_stream_open_adts(self._h, profile, sf_index, channel_config, expect_sbr, C.byref(sid)), "stream_open_adts")
        return sid.value

    def probe_sbr(self, profile: int, sf_index: int, channel_config: int, frame) -> int:
        """What expect_sbr should be for an LC-signalled stream, judged by its first frame: 0 = plain AAC-LC, 1 = SBR
        payload present, 2 = SBR + parametric stereo (JAAD's implicit signalling, sbr/SBR.java:98-101)."""
        buf = np.ascontiguousarray(frame, np.uint8)
        out = C.c_int32(0)
        self._check(self._lib.jaadb_probe_sbr(self._h, profile, sf_index, channel_config, _ptr(buf), buf.nbytes, C.byref(out)),
                    "probe_sbr")
        return int(out.value)

    def probe_sbr_asc(self, asc: bytes, frame) -> int:
        """probe_sbr for a stream described by an AudioSpecificConfig (MP4 tracks); feeds open_asc(asc, expect_sbr)."""
        a = np.frombuffer(bytes(asc), np.uint8).copy()
        buf = np.ascontiguousarray(frame, np.uint8)
        out = C.c_int32(0)
        self._check(self._lib.jaadb_probe_sbr_asc(self._h, _ptr(a), a.nbytes, _ptr(buf), buf.nbytes, C.byref(out)), "probe_sbr_asc")
        return int(out.value)

    def open_asc(self, asc: bytes, expect_sbr: int = 0) -> int:
        """Decoder.create(byte[] asc).  expect_sbr > 0 (see probe_sbr): the frames carry SBR (2: + PS) the ASC does not
        signal -- JAAD then runs its down-sampled SBR tool, 1024 samples per frame at the core rate (SURVEY A-20)."""
        sid = C.c_int32(-1)
        buf = np.frombuffer(bytes(asc), np.uint8).copy()
        self._check(self._lib.jaadb_stream_open_asc_sbr(self._h, buf.ctypes.data, len(buf), expect_sbr, C.byref(sid)), "stream_open_asc")
        return sid.value

    def close_stream(self, sid: int):
        self._check(self._lib.jaadb_stream_close(self._h, sid), "stream_close")

    def stream_info(self, sid: int) -> StreamInfo:
        info = StreamInfo()
        self._check(self._lib.jaadb_stream_get_info(self._h, sid, C.byref(info)), "stream_get_info")
        return info

    def decode(self, blob: np.ndarray, frames: np.ndarray, pcm_out: np.ndarray | None = None, pcm_offsets: np.ndarray | None = None,
               results: np.ndarray | None = None):
        """One-call decode with host buffers. Returns (pcm uint8 array, results structured array).

        `pcm_out` / `results` may be passed in to be reused across calls (every entry is overwritten)."""
        blob = np.ascontiguousarray(blob, np.uint8)
        frames = np.ascontiguousarray(frames, FRAME_DESC_DTYPE)
        if results is None:
            results = np.zeros(len(frames), FRAME_RESULT_DTYPE)
        assert results.dtype == FRAME_RESULT_DTYPE and len(results) == len(frames) and results.flags.c_contiguous
        if pcm_out is None:
            need = self._packed_bytes(frames) if pcm_offsets is None else int(pcm_offsets.max(initial=0)) + 8 * 2048 * 4
            pcm_out = np.zeros(need, np.uint8)
        if pcm_offsets is not None:
            pcm_offsets = np.ascontiguousarray(pcm_offsets, np.uint64)
        self._check(self._lib.jaadb_decode(self._h, _ptr(blob), blob.nbytes, _ptr(frames), len(frames), _ptr(pcm_out),
                                           pcm_out.nbytes, _ptr(pcm_offsets), _ptr(results)), "decode")
        return pcm_out, results

    def decode_ptr(self, blob_ptr: int, blob_bytes: int, frames: np.ndarray, pcm_ptr: int, pcm_capacity: int,
                   pcm_offsets: np.ndarray | None = None, results: np.ndarray | None = None):
        """jaadb_decode on raw addresses: `blob_ptr` / `pcm_ptr` may point to host memory or to memory of the engine's GPU
        (e.g. torch tensors' data_ptr()).  With a device pcm_ptr the kernels write the PCM in place and nothing but the
        frame table and the per-frame results crosses PCIe.  Returns the results array."""
        frames = np.ascontiguousarray(frames, FRAME_DESC_DTYPE)
        if results is None:
            results = np.zeros(len(frames), FRAME_RESULT_DTYPE)
        if pcm_offsets is not None:
            pcm_offsets = np.ascontiguousarray(pcm_offsets, np.uint64)
        self._check(self._lib.jaadb_decode(self._h, blob_ptr, blob_bytes, _ptr(frames), len(frames), pcm_ptr, pcm_capacity,
                                           _ptr(pcm_offsets), _ptr(results)), "decode")
        return results

    def decode_containers(self, kind: int, blob: np.ndarray, stream_begin: np.ndarray, stream_ids, pcm_out, results: np.ndarray,
                          frames_out: np.ndarray | None = None, threads: int = 0, pcm_capacity: int | None = None) -> int:
        """Container bytes in, PCM out (jaadb_decode_containers): `blob` holds ADTS streams (kind CONTAINER_ADTS) or MP4 files
        (CONTAINER_MP4) back to back; they are indexed on host threads while the bytes are on their way to the GPU, then
        decoded in frame-major order.  `pcm_out`: a numpy array, or a device address (int) together with pcm_capacity.
        Returns the number of frames; results[:n] (and frames_out[:n]) are filled."""
        blob = np.ascontiguousarray(blob, np.uint8)
        begin = np.ascontiguousarray(stream_begin, np.uint64)
        ids = None if stream_ids is None else np.ascontiguousarray(stream_ids, np.int32)
        assert results.dtype == FRAME_RESULT_DTYPE and results.flags.c_contiguous
        if isinstance(pcm_out, np.ndarray):
            pcm_ptr, cap = pcm_out.ctypes.data, pcm_out.nbytes
        else:
            pcm_ptr, cap = pcm_out, int(pcm_capacity)
        n = self._lib.jaadb_decode_containers(self._h, kind, _ptr(blob), _ptr(begin), len(begin) - 1, _ptr(ids), pcm_ptr, cap,
                                              _ptr(results), len(results), _ptr(frames_out), threads)
        if n < 0:
            raise EngineError("decode_containers failed: %d (%s)" % (n, self._lib.jaadb_last_error(self._h).decode()))
        return int(n)

    def packed_bytes(self, frames) -> int:
        """Size of the PCM buffer jaadb_decode fills for `frames` when pcm_offsets is None."""
        return self._packed_bytes(np.ascontiguousarray(frames, FRAME_DESC_DTYPE))

    def _packed_bytes(self, frames) -> int:
        per = 4 if self.pcm_format == PCM_F32_PLANAR else 2
        total = 0
        ids, counts = np.unique(frames["stream_id"], return_counts=True)
        for sid, n in zip(ids, counts):
            i = self.stream_info(int(sid))
            total += int(n) * i.channels * i.sample_length * per
        return total

    def batch(self, frames: np.ndarray, blob_bytes: int, pcm_offsets: np.ndarray | None = None) -> "Batch":
        return Batch(self, frames, blob_bytes, pcm_offsets)


class Batch:
    """Staged decode: create -> upload -> decode (device resident) -> download."""

    def __init__(self, engine: Engine, frames: np.ndarray, blob_bytes: int, pcm_offsets=None):
        self.engine = engine
        self._lib = engine._lib
        self.frames = np.ascontiguousarray(frames, FRAME_DESC_DTYPE)
        if pcm_offsets is not None:
            pcm_offsets = np.ascontiguousarray(pcm_offsets, np.uint64)
        h = C.c_void_p()
        engine._check(self._lib.jaadb_batch_create(engine._h, _ptr(self.frames), len(self.frames), blob_bytes, _ptr(pcm_offsets), C.byref(h)), "batch_create")
        self._h = h
        self.pcm_bytes = int(self._lib.jaadb_batch_pcm_bytes(h))

    def upload(self, blob: np.ndarray):
        self.engine._check(self._lib.jaadb_batch_upload(self._h, blob.ctypes.data, blob.nbytes), "batch_upload")

    def decode(self):
        self.engine._check(self._lib.jaadb_batch_decode(self._h), "batch_decode")

    def sync(self):
        self.engine._check(self._lib.jaadb_batch_sync(self._h), "batch_sync")

    def download(self, pcm_out: np.ndarray | None = None, want_results: bool = True, want_pcm: bool = True):
        """PCM and per-frame results of the decoded batch.  want_pcm=False fetches the results only (jaadb_batch_download with
        a NULL pcm_out): status / channels / sizes of every frame without moving the PCM off the device."""
        if pcm_out is None and want_pcm:
            pcm_out = np.zeros(self.pcm_bytes, np.uint8)
        results = np.zeros(len(self.frames), FRAME_RESULT_DTYPE) if want_results else None
        self.engine._check(self._lib.jaadb_batch_download(self._h, _ptr(pcm_out), 0 if pcm_out is None else pcm_out.nbytes, _ptr(results)),
                           "batch_download")
        return pcm_out, results

    def timings(self) -> Timings:
        t = Timings()
        self.engine._check(self._lib.jaadb_batch_timings(self._h, C.byref(t)), "batch_timings")
        return t

    def tap(self, frame: int, ch: int, want_spec: bool = True):
        q = np.zeros(1024, np.int16)
        sf = np.zeros(120, np.int16)
        cb = np.zeros(120, np.uint8)
        spec = np.zeros(1024, np.float32) if want_spec else None
        info = np.zeros(16, np.int32)
        ms = np.zeros(128, np.uint8)
        self.engine._check(self._lib.jaadb_batch_tap(self._h, frame, ch, _ptr(q), _ptr(sf), _ptr(cb), _ptr(spec), _ptr(info), _ptr(ms)), "batch_tap")
        return dict(q=q, sfidx=sf, sfbcb=cb, spec=spec, info=info, msused=ms)

    SBR_FRAME_DTYPE = np.dtype([
        ("E_orig", "<f4", (5, 64)), ("Q_div", "<f4", (2, 8)), ("Q_div2", "<f4", (2, 8)), ("f_table_res", "u1", (2, 64)),
        ("f_table_noise", "u1", (8,)), ("f_table_lim", "i1", (64,)), ("table_map_k_to_g", "u1", (64,)),
        ("bs_add_harmonic", "u1", (64,)), ("bs_add_harmonic_prev", "u1", (64,)), ("patchNoSubbands", "u1", (8,)),
        ("patchStartSubband", "i1", (8,)), ("t_E", "u1", (6,)), ("t_Q", "u1", (3,)), ("f", "u1", (6,)), ("bs_invf_mode", "u1", (5,)),
        ("mode", "u1"), ("reset", "u1"), ("L_E", "u1"), ("L_Q", "u1"), ("kx", "u1"), ("M", "u1"), ("N_high", "u1"), ("N_low", "u1"),
        ("N_Q", "u1"), ("N_L", "u1"), ("kx_prev", "u1"), ("M_prev", "u1"), ("noPatches", "u1"), ("limiter_gains", "u1"),
        ("interpol_freq", "u1"), ("smoothing_mode", "u1"), ("add_harmonic_flag_prev", "u1"), ("l_A", "i1"),
        ("prevEnvIsShort", "i1"), ("frame_status", "u1"), ("ord", "<u4"), ("back", "<u4"), ("fwd", "<u4"),
        ("back_ps", "<u4"), ("fwd_ps", "<u4"), ("pad", "u1", (12,))])

    def tap_sbr(self, frame: int, ch: int):
        """SBR record of frame `frame`, channel `ch` (None when the stream carries no SBR)."""
        out = np.zeros(1, self.SBR_FRAME_DTYPE)
        rc = self._lib.jaadb_batch_tap_sbr(self._h, frame, ch, out.ctypes.data, out.nbytes)
        if rc < 0:
            raise EngineError("batch_tap_sbr failed: %d" % rc)
        return out[0] if rc > 0 else None

    PS_FRAME_DTYPE = np.dtype([("use_ps", "u1"), ("num_env", "u1"), ("border", "u1", (6,)), ("iid_mode", "i1"), ("icc_mode", "i1"),
                               ("iid", "i1", (5, 20)), ("icc", "i1", (5, 20)), ("nr_ipdopd_par", "u1"), ("enable_ipdopd", "u1"),
                               ("pad", "u1", (12,)), ("ipd", "i1", (5, 17)), ("pad2", "u1", (11,))])

    def tap_ps(self, frame: int):
        """Parametric-stereo parameters of frame `frame` after ps_data_decode (None when the stream carries no PS)."""
        assert self.PS_FRAME_DTYPE.itemsize == 320
        out = np.zeros(1, self.PS_FRAME_DTYPE)
        self._lib.jaadb_batch_tap_ps.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        rc = self._lib.jaadb_batch_tap_ps(self._h, frame, out.ctypes.data, out.nbytes)
        if rc < 0:
            raise EngineError("batch_tap_ps failed: %d" % rc)
        return out[0] if rc > 0 else None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.jaadb_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
