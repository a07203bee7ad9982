"""Shared helpers of the parity tests: build synthetic workloads, run the oracle over them."""
from __future__ import annotations

import numpy as np

import gen
import oracle
from jaadec_b200 import FRAME_DESC_DTYPE


class Workload:
    """n_streams independent streams of one generator config, frames interleaved round-robin (frame-major)."""

    def __init__(self, cfg: gen.GenConfig, n_streams: int, base_seed: int, with_truth: bool = True, asc: bytes | None = None):
        self.cfg = cfg
        self.n_streams = n_streams
        self.asc = asc
        self.streams = [gen.generate(cfg, base_seed + s, with_truth=with_truth) for s in range(n_streams)]
        sizes = [len(s.data) for s in self.streams]
        self.base = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        self.blob = np.concatenate([s.data for s in self.streams])
        self.hdr = (2, cfg.sf_index, cfg.chan_cfg)

    def frame_table(self, stream_ids, frame_lo=0, frame_hi=None):
        """Frame descriptors for frames [lo, hi) of every stream, frame-major order."""
        hi = self.cfg.n_frames if frame_hi is None else frame_hi
        rows = []
        index = []
        for f in range(frame_lo, hi):
            for s in range(self.n_streams):
                st = self.streams[s]
                rows.append((self.base[s] + st.offsets[f], st.sizes[f], stream_ids[s]))
                index.append((s, f))
        return np.array(rows, FRAME_DESC_DTYPE), index

    def oracle_decoders(self):
        if self.asc is not None:
            return [oracle.Decoder.create_asc(self.asc) for _ in range(self.n_streams)]
        return [oracle.Decoder.create_adts(*self.hdr) for _ in range(self.n_streams)]

    def frame_bytes(self, s, f):
        st = self.streams[s]
        return st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]]


def f32_bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_float_bits(a, b):
    """Bit-exact float comparison that treats +0.0 and -0.0 as different and NaNs by pattern."""
    return np.array_equal(f32_bits(a), f32_bits(b))
