"""Container indexers (native, host side of the C ABI) against the generator's by-construction frame positions and
against the oracle's restatements of JAAD's ADTSDemultiplexer / MP4 Track.  No GPU needed: pure host logic."""
from __future__ import annotations

import numpy as np
import pytest

import gen
import oracle
from gen import mp4 as genmp4
from oracle import mp4 as omp4
from jaadec_b200 import demux, EngineError


def adts_header(payload: int, *, crc: bool, blocks: int = 0, profile=2, sf=4, ch=2) -> bytes:
    extra = 0
    if crc:
        extra = 2 + ((2 * blocks + 2 + 2 * blocks) if blocks else 0)
    flen = payload + 7 + (2 if crc else 0)   # JAAD subtracts 9 (or 7): the block CRC table is NOT part of that figure
    h = bytearray(7)
    h[0] = 0xFF
    h[1] = 0xF0 | (0 if crc else 1)
    h[2] = ((profile - 1) << 6) | (sf << 2) | (ch >> 2)
    h[3] = ((ch & 3) << 6) | (flen >> 11)
    h[4] = (flen >> 3) & 0xFF
    h[5] = ((flen & 7) << 5) | 0x1F
    h[6] = 0xFC | blocks
    return bytes(h) + b"\x5A" * extra


@pytest.mark.parametrize("cfg_no", [1, 2, 3])
def test_adts_index_matches_generator(cfg_no):
    cfg = gen.config(cfg_no, n_frames=17)
    st = gen.generate(cfg, gen.seed_for(cfg_no, 5))
    frames, info = demux.adts_index(st.data, stream_id=9, blob_offset=1000)
    assert len(frames) == cfg.n_frames == info.n_frames
    assert np.array_equal(frames["offset"], np.asarray(st.offsets) + 1000)
    assert np.array_equal(frames["nbytes"], st.sizes)
    assert (frames["stream_id"] == 9).all()
    assert (info.profile, info.sf_index, info.channel_config) == (2, cfg.sf_index, cfg.chan_cfg)
    off, sz, hdr = oracle.adts_index(st.data)
    assert np.array_equal(off, frames["offset"] - 1000) and np.array_equal(sz, frames["nbytes"])
    assert tuple(hdr) == (info.profile, info.sf_index, info.channel_config)


def test_adts_resync_crc_and_truncation():
    rng = np.random.default_rng(3)
    pieces, want = [], []
    pos = 0

    def put(b):
        nonlocal pos
        pieces.append(b)
        pos += len(b)

    # junk before the first sync, including a lone 0xFF and 0xFF 0xFF runs (the second byte is pushed back and re-read)
    put(bytes([0x00, 0xFF, 0x12, 0xFF, 0xFF]))
    for i in range(12):
        payload = int(rng.integers(1, 700))
        crc = i % 3 == 1
        blocks = 2 if i == 4 else 0
        h = adts_header(payload, crc=crc, blocks=blocks)
        put(h)
        want.append((pos, payload))
        # payload bytes avoid 0xFF so that they never look like a sync word
        put(bytes(rng.integers(0, 0xF0, payload, dtype=np.uint8)))
        if i % 4 == 2:
            put(bytes(rng.integers(0, 0xF0, int(rng.integers(1, 40)), dtype=np.uint8)))   # junk between frames
    data = np.frombuffer(b"".join(pieces), np.uint8)
    frames, info = demux.adts_index(data)
    off, sz, _ = oracle.adts_index(data)
    assert [(int(o), int(s)) for o, s in zip(frames["offset"], frames["nbytes"])] == want
    assert np.array_equal(off, frames["offset"]) and np.array_equal(sz, frames["nbytes"])
    # cut in the middle of the last payload: that frame disappears (EOFException), the rest stays
    cut = data[: want[-1][0] + want[-1][1] - 1]
    f2, _ = demux.adts_index(cut)
    assert len(f2) == len(want) - 1
    assert len(oracle.adts_index(cut)[0]) == len(want) - 1
    # more than 6144 bytes without a sync word end the stream
    gap = np.concatenate([data, np.zeros(7000, np.uint8), data])
    assert len(demux.adts_index(gap)[0]) == len(want) == len(oracle.adts_index(gap)[0])
    short_gap = np.concatenate([data, np.zeros(6000, np.uint8), data])
    assert len(demux.adts_index(short_gap)[0]) == 2 * len(want) == len(oracle.adts_index(short_gap)[0])
    # empty and garbage input
    assert len(demux.adts_index(np.zeros(0, np.uint8))[0]) == 0
    assert len(demux.adts_index(np.full(100, 0xFF, np.uint8))[0]) == len(oracle.adts_index(np.full(100, 0xFF, np.uint8))[0])


def test_adts_index_many_threads():
    cfg = gen.config(2, n_frames=9)
    streams = [gen.generate(cfg, 400 + s) for s in range(13)]
    streams.insert(5, gen.generate(gen.config(1, n_frames=4), 1))    # ragged: a shorter stream of another config
    blob = np.concatenate([s.data for s in streams] + [np.zeros(0, np.uint8)])
    begin = np.concatenate([[0], np.cumsum([len(s.data) for s in streams])])
    ids = np.arange(100, 100 + len(streams))
    for threads in (1, 4, 0):
        frames, first, infos = demux.adts_index_many(blob, begin, ids, threads=threads)
        assert first[-1] == len(frames) == sum(len(s.offsets) for s in streams)
        for s, st in enumerate(streams):
            f = frames[first[s]: first[s + 1]]
            assert np.array_equal(f["offset"], np.asarray(st.offsets) + begin[s])
            assert np.array_equal(f["nbytes"], st.sizes)
            assert (f["stream_id"] == ids[s]).all()
            assert infos[s].n_frames == len(st.offsets)
    inter = demux.interleave(frames, first)
    assert len(inter) == len(frames)
    assert list(inter["stream_id"][: len(streams)]) == list(ids)          # frame 0 of every stream first
    for s in range(len(streams)):                                            # per-stream order is kept
        assert np.array_equal(inter[inter["stream_id"] == ids[s]], frames[first[s]: first[s + 1]])


MP4_LAYOUTS = [
    dict(),
    dict(chunk_pattern=(3, 1, 5), co64=True, large_mdat=True),
    dict(decoy_track=True, moov_first=False, free_boxes=True, chunk_gap=7, long_descriptors=False),
    dict(chunk_pattern=(1,), moov_first=False),
    dict(chunk_pattern=(1000,)),
]


def _mp4_case(n_frames=23, seed=77):
    cfg = gen.config(5, n_frames=n_frames)
    st = gen.generate(cfg, seed)
    frames = [st.data[o: o + s].tobytes() for o, s in zip(st.offsets, st.sizes)]
    return frames, bytes([0x11, 0xB0])


@pytest.mark.parametrize("layout", MP4_LAYOUTS)
def test_mp4_index_matches_writer(layout):
    frames, asc = _mp4_case()
    data, off, sz = genmp4.write_mp4(frames, asc, 48000, 6, **layout)
    f, t = demux.mp4_index(data, stream_id=4, blob_offset=64)
    assert demux.asc_of(t) == asc
    assert np.array_equal(f["offset"], off + 64) and np.array_equal(f["nbytes"], sz) and (f["stream_id"] == 4).all()
    assert (t.channel_count, t.sample_rate, t.timescale, t.n_frames) == (6, 48000, 48000, len(frames))
    assert t.track_id == (2 if layout.get("decoy_track") else 1)
    assert t.duration == 1024 * len(frames) and t.object_type == 0x40
    o_asc, o_frames = omp4.parse_track(data.tobytes())
    assert o_asc == asc and [a for a, _ in o_frames] == list(off) and [b for _, b in o_frames] == list(sz)
    for j, fr in enumerate(frames):
        assert data[off[j]: off[j] + sz[j]].tobytes() == fr


def test_mp4_errors_and_truncation():
    frames, asc = _mp4_case(9)
    data, off, sz = genmp4.write_mp4(frames, asc, 48000, 6)
    with pytest.raises(EngineError):
        demux.mp4_index(np.zeros(64, np.uint8))                      # no moov
    with pytest.raises(EngineError):
        demux.mp4_index(data[:200])                                  # moov cut short
    cut = data[: off[-1] + sz[-1] - 1]                               # last sample incomplete: the stream ends before it
    f, t = demux.mp4_index(cut)
    assert len(f) == len(frames) - 1 == t.n_frames
    # a movie with only a non-audio track has no AAC track
    vid = bytearray(data.tobytes())
    i = vid.find(b"soun")
    vid[i: i + 4] = b"vide"
    with pytest.raises(EngineError):
        demux.mp4_index(bytes(vid))


def test_mp4_index_many():
    files, truth = [], []
    for s in range(7):
        frames, asc = _mp4_case(5 + s, seed=900 + s)
        data, off, sz = genmp4.write_mp4(frames, asc, 48000, 6, **MP4_LAYOUTS[s % len(MP4_LAYOUTS)])
        files.append(data)
        truth.append((off, sz))
    blob = np.concatenate(files)
    begin = np.concatenate([[0], np.cumsum([len(f) for f in files])])
    frames, first, tracks = demux.mp4_index_many(blob, begin, threads=3)
    assert first[-1] == len(frames)
    for s, (off, sz) in enumerate(truth):
        f = frames[first[s]: first[s + 1]]
        assert np.array_equal(f["offset"], off + begin[s]) and np.array_equal(f["nbytes"], sz) and (f["stream_id"] == s).all()
        assert demux.asc_of(tracks[s]) == bytes([0x11, 0xB0])
