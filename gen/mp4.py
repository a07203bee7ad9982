"""Minimal ISO base media (MP4) writer for synthetic AAC tracks -- test-vector generator, not a product path.

Writes what JAAD's mp4 package needs to find an AAC track (mp4/src/main/java/net/sourceforge/jaad/mp4/api/Track.java:
90-172): ftyp, moov(mvhd, trak(tkhd, mdia(mdhd, hdlr, minf(smhd, dinf, stbl(stsd(mp4a(esds)), stts, stsc, stsz,
stco|co64))))), mdat.  The layout knobs exist to exercise the sample-table arithmetic: chunks of varying length,
64-bit chunk offsets, a 64-bit mdat size, a decoy video-handler track before the audio track, free boxes, moov
before or after mdat, padding between chunks.  The writer returns the ground-truth (offset, size) of every sample,
known by construction.
"""
from __future__ import annotations

import struct

import numpy as np


def box(kind: bytes, body: bytes, large: bool = False) -> bytes:
    if large:
        return struct.pack(">I4sQ", 1, kind, 16 + len(body)) + body
    return struct.pack(">I4s", 8 + len(body), kind) + body


def full(kind: bytes, version: int, flags: int, body: bytes) -> bytes:
    return box(kind, struct.pack(">I", (version << 24) | flags) + body)


def descriptor(tag: int, body: bytes, long_size: bool = False) -> bytes:
    n = len(body)
    if long_size:   # 4-byte continued size, as most muxers write it
        size = bytes([0x80 | ((n >> 21) & 0x7F), 0x80 | ((n >> 14) & 0x7F), 0x80 | ((n >> 7) & 0x7F), n & 0x7F])
    else:
        assert n < 128
        size = bytes([n])
    return bytes([tag]) + size + body


def esds(asc: bytes, long_size: bool, avg_bitrate: int = 128000) -> bytes:
    dsi = descriptor(5, asc, long_size)
    dcd = descriptor(4, struct.pack(">BB", 0x40, 0x15) + b"\x00\x18\x00" + struct.pack(">II", 2 * avg_bitrate, avg_bitrate) + dsi,
                     long_size)
    sl = descriptor(6, b"\x02", long_size)
    es = descriptor(3, struct.pack(">HB", 1, 0) + dcd + sl, long_size)
    return full(b"esds", 0, 0, es)


def _trak(track_id, handler, timescale, duration, stbl_body, media_header):
    tkhd = full(b"tkhd", 0, 7, struct.pack(">IIIII", 0, 0, track_id, 0, duration) + b"\0" * 8 + struct.pack(">hhhh", 0, 0, 0x100, 0)
                + struct.pack(">9i", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000) + struct.pack(">II", 0, 0))
    mdhd = full(b"mdhd", 0, 0, struct.pack(">IIIIHH", 0, 0, timescale, duration, 0x55C4, 0))
    hdlr = full(b"hdlr", 0, 0, struct.pack(">I4s", 0, handler) + b"\0" * 12 + b"handler\0")
    dinf = box(b"dinf", full(b"dref", 0, 0, struct.pack(">I", 1) + full(b"url ", 0, 1, b"")))
    minf = box(b"minf", media_header + dinf + box(b"stbl", stbl_body))
    return box(b"trak", tkhd + box(b"mdia", mdhd + hdlr + minf))


def write_mp4(frames, asc: bytes, sample_rate: int, channels: int, *, chunk_pattern=(4,), co64: bool = False,
              large_mdat: bool = False, decoy_track: bool = False, moov_first: bool = True, free_boxes: bool = False,
              chunk_gap: int = 0, long_descriptors: bool = True, frame_duration: int = 1024):
    """Returns (file bytes, offsets[int64], sizes[int32]) for the samples in `frames`: a list of bytes-like, or -- the fast
    path for the benchmark's thousands of files -- a pair (data uint8 array, sizes) of samples stored back to back."""
    packed = isinstance(frames, tuple)
    if packed:
        data_in, sizes = np.ascontiguousarray(frames[0], np.uint8), np.asarray(frames[1], np.int32)
        n = len(sizes)
        assert chunk_gap == 0 and int(sizes.sum()) == len(data_in)
    else:
        n = len(frames)
        sizes = np.array([len(f) for f in frames], np.int32)
    # chunking: cycle through chunk_pattern
    chunks, i, k = [], 0, 0
    while i < n:
        c = min(chunk_pattern[k % len(chunk_pattern)], n - i)
        chunks.append((i, c))
        i += c
        k += 1
    # stsc runs: (first_chunk, samples_per_chunk, 1) whenever the count changes
    stsc_rows = []
    for ci, (_, c) in enumerate(chunks):
        if not stsc_rows or stsc_rows[-1][1] != c:
            stsc_rows.append((ci + 1, c, 1))

    def build(chunk_offsets):
        mp4a = box(b"mp4a", b"\0" * 6 + struct.pack(">H", 1) + b"\0" * 8 + struct.pack(">HHHH", channels, 16, 0, 0)
                   + struct.pack(">HH", sample_rate & 0xFFFF, 0) + esds(asc, long_descriptors))
        stsd = full(b"stsd", 0, 0, struct.pack(">I", 1) + mp4a)
        stts = full(b"stts", 0, 0, struct.pack(">III", 1, n, frame_duration))
        stsc = full(b"stsc", 0, 0, struct.pack(">I", len(stsc_rows)) + b"".join(struct.pack(">III", *r) for r in stsc_rows))
        stsz = full(b"stsz", 0, 0, struct.pack(">II", 0, n) + sizes.astype(">u4").tobytes())
        if co64:
            stco = full(b"co64", 0, 0, struct.pack(">I", len(chunks)) + np.asarray(chunk_offsets, ">u8").tobytes())
        else:
            stco = full(b"stco", 0, 0, struct.pack(">I", len(chunks)) + np.asarray(chunk_offsets, ">u4").tobytes())
        stbl = stsd + stts + stsc + stsz + stco
        if free_boxes:
            stbl = stsd + box(b"free", b"\0" * 5) + stts + stsc + box(b"sgpd", b"\0" * 12) + stsz + stco
        audio = _trak(2 if decoy_track else 1, b"soun", sample_rate, n * frame_duration, stbl, full(b"smhd", 0, 0, b"\0" * 4))
        tracks = audio
        if decoy_track:   # a video-handler track whose sample table must not be picked up
            vstbl = (full(b"stsd", 0, 0, struct.pack(">I", 0)) + full(b"stts", 0, 0, struct.pack(">I", 0))
                     + full(b"stsc", 0, 0, struct.pack(">I", 0)) + full(b"stsz", 0, 0, struct.pack(">II", 0, 0))
                     + full(b"stco", 0, 0, struct.pack(">I", 0)))
            tracks = _trak(1, b"vide", 90000, 0, vstbl, full(b"vmhd", 0, 1, b"\0" * 8)) + audio
        mvhd = full(b"mvhd", 0, 0, struct.pack(">IIII", 0, 0, sample_rate, n * frame_duration) + struct.pack(">IH", 0x10000, 0x100)
                    + b"\0" * 10 + struct.pack(">9i", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000) + b"\0" * 24
                    + struct.pack(">I", 3))
        return box(b"moov", mvhd + tracks + (box(b"udta", box(b"free", b"xyz")) if free_boxes else b""))

    ftyp = box(b"ftyp", b"M4A \0\0\0\0M4A mp42isom")
    pre = ftyp + (box(b"free", b"\0" * 11) if free_boxes else b"")
    # mdat body: chunks back to back with an optional gap of junk bytes between them
    if packed:
        sample_rel = np.concatenate([[0], np.cumsum(sizes, dtype=np.int64)[:-1]]) if n else np.zeros(0, np.int64)
        rel = [int(sample_rel[s0]) for (s0, _) in chunks]
        body = data_in.tobytes()
    else:
        body = bytearray()
        rel = []          # chunk offsets relative to the mdat body
        sample_rel = np.zeros(n, np.int64)
        for (s0, c) in chunks:
            rel.append(len(body))
            for j in range(s0, s0 + c):
                sample_rel[j] = len(body)
                body += bytes(frames[j])
            body += b"\xAA" * chunk_gap
    mdat_hdr = 16 if large_mdat else 8
    moov_len = len(build([0] * len(chunks)))   # offsets have a fixed width, so the length does not depend on them
    base = len(pre) + (moov_len if moov_first else 0) + mdat_hdr
    moov = build([base + r for r in rel])
    assert len(moov) == moov_len
    mdat = box(b"mdat", bytes(body), large=large_mdat)
    data = pre + (moov + mdat if moov_first else mdat + moov)
    return np.frombuffer(data, np.uint8).copy(), sample_rel + base, sizes
