"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product (jaadec_b200/).

Restatement of the part of JAAD's MP4 demultiplexer that feeds the AAC decoder (paths relative to
/root/reference/mp4/src/main/java/net/sourceforge/jaad/mp4/):
  boxes/BoxFactory.java:319-363      box header (32/64-bit size, uuid), children inside the parent's span
  api/Movie.java:55-70               handler type 'soun' -> AudioTrack
  api/Track.java:90-152              stsz + stco|co64 + stsc + stts -> frames, sorted by timestamp
  api/Track.java:155-172, od/*.java  esds -> ESDescriptor -> DecoderConfigDescriptor -> DecoderSpecificInfo
Pure Python, for small files.  Pinned by gen/mp4.py's by-construction sample positions (tests/test_demux_cpu.py);
JAAD ships no MP4 fixtures and no JVM is available: parity unpinned against the real JAAD.
"""
from __future__ import annotations

import struct

CONTAINERS = {b"moov", b"trak", b"mdia", b"minf", b"stbl", b"dinf", b"udta", b"edts"}


def boxes(d: bytes, start: int, end: int):
    p = start
    while p + 8 <= end:
        size, kind = struct.unpack_from(">I4s", d, p)
        body = p + 8
        if size == 1:
            size = struct.unpack_from(">Q", d, body)[0]
            body += 8
        elif size == 0:
            size = end - p
        if kind == b"uuid":
            body += 16
        if size < body - p or p + size > end:
            return
        yield kind, body, p + size
        p += size


def child(d, start, end, kind):
    for k, b, e in boxes(d, start, end):
        if k == kind:
            return b, e
    return None


def descriptors(d, start, end):
    p = start
    while p + 2 <= end:
        tag = d[p]
        p += 1
        size = 0
        while True:
            b = d[p]
            p += 1
            size = (size << 7) | (b & 0x7F)
            if not b & 0x80:
                break
        yield tag, p, min(end, p + size)
        p += size


def parse_track(d: bytes):
    """Returns (asc bytes, [(offset, size)...] in decoding-time order) of the first 'soun'/mp4a track, or None."""
    d = bytes(d)
    moov = child(d, 0, len(d), b"moov")
    if not moov:
        return None
    for kind, tb, te in boxes(d, *moov):
        if kind != b"trak":
            continue
        mdia = child(d, tb, te, b"mdia")
        hdlr = mdia and child(d, *mdia, b"hdlr")
        if not hdlr or d[hdlr[0] + 8: hdlr[0] + 12] != b"soun":
            continue
        minf = child(d, *mdia, b"minf")
        stbl = minf and child(d, *minf, b"stbl")
        stsd = stbl and child(d, *stbl, b"stsd")
        if not stsd:
            continue
        entry = next(boxes(d, stsd[0] + 8, stsd[1]), None)
        if not entry or entry[0] != b"mp4a":
            continue
        es = child(d, entry[1] + 28, entry[2], b"esds")
        asc = None
        for tag, b, e in descriptors(d, es[0] + 4, es[1]):
            if tag != 3:
                continue
            flags = d[b + 2]
            q = b + 3 + (2 if flags & 0x80 else 0)
            if flags & 0x40:
                q += 1 + d[q]
            for tag2, b2, e2 in descriptors(d, q, e):
                if tag2 == 4:
                    for tag3, b3, e3 in descriptors(d, b2 + 13, e2):
                        if tag3 == 5:
                            asc = d[b3:e3]
            break
        stsz = child(d, *stbl, b"stsz")[0]
        stco = child(d, *stbl, b"stco")
        width = 4
        if not stco:
            stco, width = child(d, *stbl, b"co64"), 8
        stsc = child(d, *stbl, b"stsc")[0]
        stts = child(d, *stbl, b"stts")[0]
        fixed, n = struct.unpack_from(">II", d, stsz + 4)
        sizes = [fixed] * n if fixed else list(struct.unpack_from(">%dI" % n, d, stsz + 12))
        n_chunks = struct.unpack_from(">I", d, stco[0] + 4)[0]
        chunk_off = list(struct.unpack_from(">%d%s" % (n_chunks, "I" if width == 4 else "Q"), d, stco[0] + 8))
        n_runs = struct.unpack_from(">I", d, stsc + 4)[0]
        runs = [struct.unpack_from(">III", d, stsc + 8 + 12 * i) for i in range(n_runs)]
        n_tt = struct.unpack_from(">I", d, stts + 4)[0]
        times, t = [], 0
        for i in range(n_tt):
            cnt, delta = struct.unpack_from(">II", d, stts + 8 + 8 * i)
            for _ in range(cnt):
                times.append(t)
                t += delta
        frames, cur = [], 0
        for i, (first, per, _) in enumerate(runs):
            last = runs[i + 1][0] - 1 if i + 1 < n_runs else n_chunks
            for j in range(first - 1, last):
                off = chunk_off[j]
                for _ in range(per):
                    frames.append((times[cur], off, sizes[cur]))
                    off += sizes[cur]
                    cur += 1
        frames.sort(key=lambda f: f[0])   # stable, like Collections.sort
        return asc, [(o, s) for _, o, s in frames]
    return None
