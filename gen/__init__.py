"""Synthetic AAC bitstream generator (ctypes front-end of gen/aacgen.cpp).

Produces random valid ADTS / raw-frame AAC streams of the BASELINE.json
configurations together with the generator's own ground truth (SURVEY.md §8d,
Appendix C).  Test and benchmark input only -- not product code.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libaacgen.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
        for f in os.listdir(_HERE) if f.endswith((".cpp", ".inc"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB


class _Cfg(C.Structure):
    _fields_ = [
        ("sf_index", C.c_int32), ("chan_cfg", C.c_int32), ("n_frames", C.c_int32), ("target_bytes", C.c_int32),
        ("long_only", C.c_int32), ("adts", C.c_int32), ("p_transient", C.c_float), ("p_common_window", C.c_float),
        ("p_tns", C.c_float), ("p_is", C.c_float), ("ms_mode", C.c_int32), ("sbr_mode", C.c_int32),
        ("target_rms", C.c_float), ("sbr_quirk", C.c_int32), ("sbr_downsampled", C.c_int32), ("p_pns", C.c_float), ("tns_mild", C.c_int32), ("ps_ext", C.c_float), ("ps_iso", C.c_int32), ("p_pulse", C.c_float), ("pulse_wild", C.c_int32), ("p_drc", C.c_float),
    ]


class _Truth(C.Structure):
    _fields_ = [("q", C.c_void_p), ("sfidx", C.c_void_p), ("sfbcb", C.c_void_p), ("info", C.c_void_p), ("msused", C.c_void_p),
                ("sbr", C.c_void_p), ("ps", C.c_void_p), ("tns", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.jg_generate.restype = C.c_int64
        _lib.jg_generate.argtypes = [C.POINTER(_Cfg), C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(_Truth)]
        _lib.jg_generate_many.restype = C.c_int64
        _lib.jg_generate_many.argtypes = [C.POINTER(_Cfg), C.c_uint64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.jg_ics_per_frame.argtypes = [C.c_int]
        _lib.jg_elements_per_frame.argtypes = [C.c_int]
    return _lib


@dataclass
class GenConfig:
    sf_index: int = 3
    chan_cfg: int = 2
    n_frames: int = 16
    target_bytes: int = 341
    long_only: bool = False
    adts: bool = True
    p_transient: float = 0.1
    p_common_window: float = 0.8
    p_tns: float = 0.3
    p_is: float = 0.05
    ms_mode: int = 1
    sbr_mode: int = 0
    target_rms: float = 2500.0
    sbr_quirk: bool = False   # also emit coupled SBR frames only the reference's parser reads (see gen/aacgen_sbr.inc)
    sbr_downsampled: bool = False   # SBR band tables for the core rate: streams JAAD opens from an ASC without SBR signalling
    p_pns: float = 0.0        # share of the scalefactor bands coded as perceptual noise (codebook 13)
    tns_mild: bool = False    # TNS filters an ISO decoder can apply: order <= 12 (long) / 7 (short), small coefficients
    ps_ext: float = 0.0       # probability that a parametric-stereo header enables the IPD/OPD extension
    ps_iso: bool = False      # PS modes restricted to the ones JAAD decodes the way ISO/IEC 14496-3 says (10-band, type-A mixing)
    p_pulse: float = 0.0      # probability that a long-window ICS carries pulse_data (truth q = what an ISO decoder reconstructs)
    pulse_wild: bool = False  # pulses may also land past max_sfb (decoders must leave those alone; FFmpeg does not)
    p_drc: float = 0.0        # probability that a frame ends with a dynamic_range_info fill element (JAAD parses and drops it)

    def c(self) -> _Cfg:
        return _Cfg(self.sf_index, self.chan_cfg, self.n_frames, self.target_bytes, int(self.long_only), int(self.adts),
                    self.p_transient, self.p_common_window, self.p_tns, self.p_is, self.ms_mode, self.sbr_mode,
                    self.target_rms, int(self.sbr_quirk), int(self.sbr_downsampled), self.p_pns, int(self.tns_mild), self.ps_ext, int(self.ps_iso), self.p_pulse, int(self.pulse_wild), self.p_drc)


# BASELINE.json configurations.  Seeds: 0xAAC0 + 1000*config + stream_id.
def config(n: int, **over) -> GenConfig:
    base = {
        1: dict(sf_index=4, chan_cfg=2, n_frames=431, target_bytes=372, long_only=True, p_transient=0.0, p_tns=0.0, p_is=0.0, ms_mode=0),
        2: dict(sf_index=3, chan_cfg=2, n_frames=469, target_bytes=341),
        3: dict(sf_index=6, chan_cfg=2, n_frames=235, target_bytes=341, sbr_mode=1),
        4: dict(sf_index=6, chan_cfg=1, n_frames=235, target_bytes=171, sbr_mode=2),
        5: dict(sf_index=3, chan_cfg=6, n_frames=469, target_bytes=1024, adts=False),
    }[n]
    base.update(over)
    return GenConfig(**base)


def seed_for(config_no: int, stream_id: int) -> int:
    return 0xAAC0 + 1000 * config_no + stream_id


@dataclass
class Stream:
    data: np.ndarray      # uint8 bytes (ADTS file or concatenated raw frames)
    offsets: np.ndarray   # int64 payload offset of each frame in data
    sizes: np.ndarray     # int32 payload size
    truth: dict | None


def generate(cfg: GenConfig, seed: int, with_truth: bool = False) -> Stream:
    L = lib()
    cap = cfg.n_frames * (cfg.target_bytes * 3 + 4096)
    out = np.zeros(cap, np.uint8)
    offs = np.zeros(cfg.n_frames, np.int64)
    sizes = np.zeros(cfg.n_frames, np.int32)
    truth = None
    tp = None
    if with_truth:
        nics = L.jg_ics_per_frame(cfg.chan_cfg)
        nel = L.jg_elements_per_frame(cfg.chan_cfg)
        truth = dict(
            q=np.zeros((cfg.n_frames, nics, 1024), np.int16),
            sfidx=np.zeros((cfg.n_frames, nics, 120), np.int16),
            sfbcb=np.zeros((cfg.n_frames, nics, 120), np.uint8),
            info=np.zeros((cfg.n_frames, nics, 16), np.int32),
            msused=np.zeros((cfg.n_frames, nel, 128), np.uint8),
        )
        sbr_p = None
        if cfg.sbr_mode:
            truth["sbr"] = np.zeros((cfg.n_frames, nics, L.jg_sbr_truth_ints()), np.int32)
            sbr_p = truth["sbr"].ctypes.data
        ps_p = None
        if cfg.sbr_mode > 1:
            truth["ps"] = np.zeros((cfg.n_frames, L.jg_ps_truth_ints()), np.int32)
            ps_p = truth["ps"].ctypes.data
        truth["tns"] = np.zeros((cfg.n_frames, nics, 600), np.int32)
        t = _Truth(*(truth[k].ctypes.data for k in ("q", "sfidx", "sfbcb", "info", "msused")), sbr_p, ps_p, truth["tns"].ctypes.data)
        tp = C.byref(t)
    cc = cfg.c()
    n = L.jg_generate(C.byref(cc), seed, out.ctypes.data, cap, offs.ctypes.data, sizes.ctypes.data, tp)
    if n < 0:
        raise RuntimeError("generator failed: %d" % n)
    return Stream(out[:n].copy(), offs, sizes, truth)


def generate_many(cfg: GenConfig, base_seed: int, n_streams: int, threads: int | None = None):
    """Returns (blob uint8, offsets [S,F] int64 absolute, sizes [S,F] int32).  Streams live at s*stride."""
    L = lib()
    stride = cfg.n_frames * (cfg.target_bytes * 2 + 64) + 4096
    stride = (stride + 255) // 256 * 256
    blob = np.zeros(n_streams * stride, np.uint8)
    offs = np.zeros((n_streams, cfg.n_frames), np.int64)
    sizes = np.zeros((n_streams, cfg.n_frames), np.int32)
    sb = np.zeros(n_streams, np.int64)
    cc = cfg.c()
    rc = L.jg_generate_many(C.byref(cc), base_seed, n_streams, blob.ctypes.data, stride, offs.ctypes.data,
                            sizes.ctypes.data, sb.ctypes.data, threads or min(32, os.cpu_count() or 8))
    if rc < 0:
        raise RuntimeError("generator failed: %d" % rc)
    # compact: streams back to back (the stride layout only exists so threads can write independently)
    starts = np.concatenate([[0], np.cumsum(sb)[:-1]]).astype(np.int64)
    out = np.empty(int(sb.sum()), np.uint8)
    for s in range(n_streams):
        out[starts[s]: starts[s] + sb[s]] = blob[s * stride: s * stride + sb[s]]
    offs += (starts - np.arange(n_streams, dtype=np.int64) * stride)[:, None]
    return out, offs, sizes, sb
