"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of the C++ restatement of JAAD (oracle/*.hpp).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
module; the product package (jaadec_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libjaad_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.jo_open_adts.restype = C.c_void_p
        _lib.jo_open_adts.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        _lib.jo_open_asc.restype = C.c_void_p
        _lib.jo_open_asc.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        _lib.jo_close.argtypes = [C.c_void_p]
        _lib.jo_set_tns_mode.argtypes = [C.c_void_p, C.c_int]
        _lib.jo_set_pulse_mode.argtypes = [C.c_void_p, C.c_int]
        _lib.jo_decode_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.jo_tap_ics.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.jo_tap_msused.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.jo_tap_sbr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.jo_adts_index.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.jo_decode_streams.restype = C.c_double
        _lib.jo_decode_streams.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return _lib


class AACError(Exception):
    def __init__(self, status):
        super().__init__("oracle status %d" % status)
        self.status = status


class Decoder:
    """Mirror of net.sourceforge.jaad.aac.Decoder (create + decodeFrame)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def create_adts(cls, profile: int, sf_index: int, chan_cfg: int) -> "Decoder":
        st = C.c_int(0)
        h = lib().jo_open_adts(profile, sf_index, chan_cfg, C.byref(st))
        if not h:
            raise AACError(st.value)
        return cls(h)

    @classmethod
    def create_asc(cls, asc: bytes) -> "Decoder":
        st = C.c_int(0)
        buf = np.frombuffer(asc, np.uint8).copy()
        h = lib().jo_open_asc(buf.ctypes.data, len(buf), C.byref(st))
        if not h:
            raise AACError(st.value)
        return cls(h)

    def set_tns_mode(self, mode: int) -> "Decoder":
        """0 = JAAD (TNS data parsed and ignored, tools/TNS.java:63-68), 1 = the ISO 14496-3 4.6.9 all-pole filter."""
        lib().jo_set_tns_mode(self._h, int(mode))
        return self

    def set_pulse_mode(self, mode: int) -> "Decoder":
        """0 = JAAD (pulse_data parsed and ignored, syntax/ICStream.java:17), 1 = the pulses of ISO 14496-3 4.6.3.3 applied."""
        lib().jo_set_pulse_mode(self._h, int(mode))
        return self

    def close(self):
        if self._h:
            lib().jo_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode_frame(self, frame: np.ndarray, big_endian: bool = False):
        """Returns dict(status, channels, sample_length, sample_rate, f32 [C,L], s16 [L,C])."""
        frame = np.ascontiguousarray(frame, np.uint8)
        meta = np.zeros(4, np.int32)
        f32 = np.zeros(8 * 2048, np.float32)
        s16 = np.zeros(8 * 2048, np.int16)
        lib().jo_decode_frame(self._h, frame.ctypes.data, len(frame), f32.ctypes.data, s16.ctypes.data, int(big_endian), meta.ctypes.data)
        st, ch, ln, sr = (int(x) for x in meta)
        out = dict(status=st, channels=ch, sample_length=ln, sample_rate=sr)
        if st == 0:
            out["f32"] = f32[: ch * ln].reshape(ch, ln).copy()
            out["s16"] = s16[: ch * ln].reshape(ln, ch).copy()
        return out

    def saw_sbr_payload(self) -> bool:
        """The parse of the frame just decoded reached a fill element that claims an SBR payload (for the fuzz tools: in a stream
        opened without SBR, JAAD creates an SBR object on the spot, the engine needs that decision at stream_open)."""
        lib().jo_saw_sbr_payload.argtypes = [C.c_void_p]
        return bool(lib().jo_saw_sbr_payload(self._h))

    def tap_ics(self, el: int, ch: int):
        q = np.zeros(1024, np.int16)
        sf = np.zeros(120, np.int16)
        cb = np.zeros(120, np.uint8)
        spec = np.zeros(1024, np.float32)
        info = np.zeros(16, np.int32)
        r = lib().jo_tap_ics(self._h, el, ch, q.ctypes.data, sf.ctypes.data, cb.ctypes.data, spec.ctypes.data, info.ctypes.data)
        if r < 0:
            return None
        return dict(type=r, q=q, sfidx=sf, sfbcb=cb, spec=spec, info=info)

    def tap_sbr(self, el: int, ch: int):
        """SBR state of the frame just decoded (None if the element carries no SBR): dict(ints[480], extra[16],
        e_orig[5,64] float32, q_div[2,64] float32); ints uses the generator's truth layout."""
        out = np.zeros(480 + 16 + 320 + 128, np.int32)
        if lib().jo_tap_sbr(self._h, el, ch, out.ctypes.data) < 0:
            return None
        return dict(ints=out[:480].copy(), extra=out[480:496].copy(), e_orig=out[496:816].view(np.float32).reshape(5, 64).copy(),
                    q_div=out[816:944].view(np.float32).reshape(2, 64).copy())

    def tap_ps(self, el: int = 0):
        """PS state after the frame just decoded: dict(num_env, border[6], iid[5,34], icc[5,34], iid_mode, icc_mode, ipd[5,17],
        nr_ipdopd_par (Extension.nr_par), enable_ipdopd (ExtData.enabled))."""
        out = np.zeros(440, np.int32)
        lib().jo_tap_ps.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        if lib().jo_tap_ps(self._h, el, out.ctypes.data) < 0:
            return None
        return dict(num_env=int(out[0]), border=out[1:7].copy(), iid=out[8:178].reshape(5, 34).copy(),
                    icc=out[178:348].reshape(5, 34).copy(), iid_mode=int(out[348]), icc_mode=int(out[349]), raw=out[:348].copy(),
                    ipd=out[350:435].reshape(5, 17).copy(), nr_ipdopd_par=int(out[435]), enable_ipdopd=int(out[436]))

    def tap_msused(self, el: int):
        ms = np.zeros(128, np.uint8)
        if lib().jo_tap_msused(self._h, el, ms.ctypes.data) < 0:
            return None
        return ms


def qmf_roundtrip(x: np.ndarray):
    """x [n_frames*1024] float32 -> (X [n_frames*32, 32] complex64, pcm [n_frames*2048] float32) through JAAD's 32-band
    analysis and 64-band synthesis banks."""
    x = np.ascontiguousarray(x, np.float32)
    n = len(x) // 1024
    X = np.zeros((n * 32, 32, 2), np.float32)
    pcm = np.zeros(n * 2048, np.float32)
    L = lib()
    L.jo_qmf_roundtrip.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    if L.jo_qmf_roundtrip(x.ctypes.data, n, X.ctypes.data, pcm.ctypes.data) != 0:
        raise RuntimeError("oracle built without SBR")
    return X[..., 0] + 1j * X[..., 1], pcm


def qmf_roundtrip32(x: np.ndarray):
    """x [n_frames*1024] float32 -> pcm [n_frames*1024] float32 through JAAD's 32-band analysis bank and its 32-band
    (down-sampled) synthesis bank (sbr/SynthesisFilterbank32.java)."""
    x = np.ascontiguousarray(x, np.float32)
    n = len(x) // 1024
    pcm = np.zeros(n * 1024, np.float32)
    L = lib()
    L.jo_qmf_roundtrip32.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    if L.jo_qmf_roundtrip32(x.ctypes.data, n, pcm.ctypes.data) != 0:
        raise RuntimeError("oracle built without SBR")
    return pcm


def adts_index(data: np.ndarray, max_frames: int = 1 << 20):
    data = np.ascontiguousarray(data, np.uint8)
    offs = np.zeros(max_frames, np.int64)
    sizes = np.zeros(max_frames, np.int32)
    hdr = np.zeros(3, np.int32)
    n = lib().jo_adts_index(data.ctypes.data, len(data), offs.ctypes.data, sizes.ctypes.data, max_frames, hdr.ctypes.data)
    return offs[:n].copy(), sizes[:n].copy(), tuple(int(x) for x in hdr)


def decode_streams(blob, first, offsets, sizes, *, hdr=None, asc=None, threads=1, pcm_out=None, pcm_first=None, big_endian=False):
    """CPU baseline: decode many independent streams with a pool of `threads` workers. Returns (seconds, samples, errors)."""
    blob = np.ascontiguousarray(blob, np.uint8)
    first = np.ascontiguousarray(first, np.int64)
    offsets = np.ascontiguousarray(offsets, np.int64).ravel()
    sizes = np.ascontiguousarray(sizes, np.int32).ravel()
    n_streams = len(first) - 1
    samples = C.c_int64(0)
    errors = C.c_int64(0)
    hdr_a = np.asarray(hdr if hdr is not None else (0, 0, 0), np.int32)
    asc_a = np.frombuffer(asc, np.uint8).copy() if asc is not None else np.zeros(1, np.uint8)
    sec = lib().jo_decode_streams(
        n_streams, blob.ctypes.data, first.ctypes.data, offsets.ctypes.data, sizes.ctypes.data,
        0 if asc is None else 1, hdr_a.ctypes.data, asc_a.ctypes.data, 0 if asc is None else len(asc_a), threads,
        pcm_out.ctypes.data if pcm_out is not None else None,
        np.ascontiguousarray(pcm_first, np.int64).ctypes.data if pcm_first is not None else None,
        int(big_endian), C.byref(samples), C.byref(errors))
    return sec, samples.value, errors.value
