"""End-to-end container path on the GPU: MP4 files / ADTS streams -> native indexer -> engine -> PCM, against the
oracle decoding the generator's raw frames (the way JAAD's Main walks a Track or an ADTSDemultiplexer)."""
import numpy as np
import pytest

import gen
import oracle
from gen import mp4 as genmp4
from jaadec_b200 import Engine, demux, PCM_S16LE

pytestmark = pytest.mark.gpu


def test_mp4_files_decode_like_raw_frames():
    asc = bytes([0x11, 0xB0])    # AAC-LC, 48 kHz, 5.1
    cfg = gen.config(5, n_frames=14, p_transient=0.3)
    layouts = [dict(), dict(chunk_pattern=(3, 1, 5), co64=True, large_mdat=True),
               dict(decoy_track=True, moov_first=False, free_boxes=True, chunk_gap=7, long_descriptors=False)]
    streams = [gen.generate(cfg, 3100 + s) for s in range(len(layouts))]
    files = []
    for st, kw in zip(streams, layouts):
        raw = [st.data[o: o + n].tobytes() for o, n in zip(st.offsets, st.sizes)]
        files.append(genmp4.write_mp4(raw, asc, 48000, 6, **kw)[0])
    blob = np.concatenate(files)
    begin = np.concatenate([[0], np.cumsum([len(f) for f in files])])
    eng = Engine(max_streams=8, pcm_format=PCM_S16LE)
    # index first to learn each file's AudioSpecificConfig, open the streams, then index with the engine's ids
    _, _, tracks = demux.mp4_index_many(blob, begin)
    ids = [eng.open_asc(demux.asc_of(t)) for t in tracks]
    frames, first, _ = demux.mp4_index_many(blob, begin, ids)
    frames = demux.interleave(frames, first)
    pcm, res = eng.decode(blob, frames)
    assert (res["status"] == 0).all()
    per = 6 * 1024 * 2
    pcm = np.frombuffer(pcm, np.int16).reshape(len(frames), 1024, 6)
    assert pcm.nbytes == len(frames) * per
    decs = [oracle.Decoder.create_asc(asc) for _ in streams]
    k = 0
    for f in range(cfg.n_frames):
        for s, st in enumerate(streams):
            r = decs[s].decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
            assert r["status"] == 0 and frames["stream_id"][k] == ids[s]
            assert np.array_equal(pcm[k], r["s16"]), (s, f)
            k += 1


def test_adts_streams_indexed_natively():
    cfg = gen.config(2, n_frames=11, p_transient=0.3)
    streams = [gen.generate(cfg, 5200 + s) for s in range(5)]
    blob = np.concatenate([s.data for s in streams])
    begin = np.concatenate([[0], np.cumsum([len(s.data) for s in streams])])
    eng = Engine(max_streams=8, pcm_format=PCM_S16LE)
    _, _, infos = demux.adts_index_many(blob, begin)
    ids = [eng.open_adts(i.profile, i.sf_index, i.channel_config) for i in infos]
    frames, first, _ = demux.adts_index_many(blob, begin, ids)
    pcm, res = eng.decode(blob, frames)             # stream-major order this time
    assert (res["status"] == 0).all()
    pcm = np.frombuffer(pcm, np.int16).reshape(len(frames), 1024, 2)
    for s, st in enumerate(streams):
        dec = oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg)
        for f in range(cfg.n_frames):
            r = dec.decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
            assert np.array_equal(pcm[first[s] + f], r["s16"]), (s, f)


def _asc(fields):
    v = n = 0
    for val, w in fields:
        v = (v << w) | val
        n += w
    pad = (8 - n % 8) % 8
    return (v << pad).to_bytes((n + pad) // 8, "big")


def test_mp4_he_aac_tracks_with_explicit_signalling():
    """HE-AAC v1 / v2 in MP4: the AudioSpecificConfig in esds says SBR (object type 5) or PS (29) with the extension
    sampling rate (DecoderConfig.java:175-254); frames are raw_data_blocks.  Indexed natively, opened from the ASC."""
    cases = [
        (_asc([(5, 5), (6, 4), (2, 4), (3, 4), (2, 5), (0, 3)]), gen.config(3, n_frames=12, adts=False), 2, 48000),
        (_asc([(29, 5), (6, 4), (1, 4), (3, 4), (2, 5), (0, 3)]), gen.config(4, n_frames=12, adts=False), 1, 48000),
    ]
    eng = Engine(max_streams=8, pcm_format=PCM_S16LE, sbr_tile_frames=5)
    files, streams, ascs = [], [], []
    for k, (asc, cfg, core_ch, rate) in enumerate(cases):
        for j in range(2):
            st = gen.generate(cfg, 6400 + 10 * k + j)
            raw = [st.data[o: o + n].tobytes() for o, n in zip(st.offsets, st.sizes)]
            files.append(genmp4.write_mp4(raw, asc, rate, core_ch, chunk_pattern=(2, 5), frame_duration=2048)[0])
            streams.append(st)
            ascs.append(asc)
    blob = np.concatenate(files)
    begin = np.concatenate([[0], np.cumsum([len(f) for f in files])])
    _, _, tracks = demux.mp4_index_many(blob, begin)
    assert [demux.asc_of(t) for t in tracks] == ascs
    ids = [eng.open_asc(demux.asc_of(t)) for t in tracks]
    for sid in ids:
        info = eng.stream_info(sid)
        assert (info.channels, info.sample_length, info.sample_rate, info.sbr != 0) == (2, 2048, 48000, True)
    frames, first, _ = demux.mp4_index_many(blob, begin, ids)
    pcm, res = eng.decode(blob, frames)
    assert (res["status"] == 0).all()
    pcm = np.frombuffer(pcm, np.int16).reshape(len(frames), 2048, 2)
    for s, st in enumerate(streams):
        dec = oracle.Decoder.create_asc(ascs[s])
        for f in range(len(st.offsets)):
            r = dec.decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
            assert r["status"] == 0
            assert np.array_equal(pcm[first[s] + f], r["s16"]), (s, f)


def test_mp4_he_aac_tracks_with_implicit_signalling_run_the_downsampled_tool():
    """HE-AAC in MP4 whose esds only says AAC-LC at the core rate: JAAD meets the SBR payload with the output rate already
    fixed by the ASC and runs its down-sampled SBR tool (SURVEY A-20: 1024 samples per frame at the core rate, band tables
    for the core rate).  The probe looks at the track's first sample; an LC-only track next to them stays plain."""
    cases = [
        (_asc([(2, 5), (6, 4), (2, 4), (0, 3)]), gen.GenConfig(sf_index=6, chan_cfg=2, n_frames=14, target_bytes=341, sbr_mode=1, adts=False, sbr_downsampled=True), 2, 1),
        (_asc([(2, 5), (6, 4), (1, 4), (0, 3)]), gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=14, target_bytes=171, sbr_mode=2, adts=False, sbr_downsampled=True), 1, 2),
        (_asc([(2, 5), (6, 4), (2, 4), (0, 3)]), gen.GenConfig(sf_index=6, chan_cfg=2, n_frames=14, target_bytes=300, adts=False), 2, 0),
    ]
    eng = Engine(max_streams=8, pcm_format=PCM_S16LE, sbr_tile_frames=4)
    files, streams, ascs, want = [], [], [], []
    for k, (asc, cfg, core_ch, mode) in enumerate(cases):
        st = gen.generate(cfg, 7100 + k)
        raw = [st.data[o: o + n].tobytes() for o, n in zip(st.offsets, st.sizes)]
        files.append(genmp4.write_mp4(raw, asc, 24000, core_ch, chunk_pattern=(3, 4), frame_duration=1024)[0])
        streams.append(st)
        ascs.append(asc)
        want.append(mode)
    blob = np.concatenate(files)
    begin = np.concatenate([[0], np.cumsum([len(f) for f in files])])
    frames0, first0, tracks = demux.mp4_index_many(blob, begin)
    ids = []
    for s, t in enumerate(tracks):
        o, n = int(frames0["offset"][first0[s]]), int(frames0["nbytes"][first0[s]])
        mode = eng.probe_sbr_asc(demux.asc_of(t), blob[o:o + n])
        assert mode == want[s]
        ids.append(eng.open_asc(demux.asc_of(t), expect_sbr=mode))
    for sid in ids:
        info = eng.stream_info(sid)
        assert (info.channels, info.sample_length, info.sample_rate) == (2, 1024, 24000)
    frames, first, _ = demux.mp4_index_many(blob, begin, ids)
    pcm, res = eng.decode(blob, frames)
    assert (res["status"] == 0).all()
    pcm = np.frombuffer(pcm, np.int16).reshape(len(frames), 1024, 2)
    for s, st in enumerate(streams):
        dec = oracle.Decoder.create_asc(ascs[s])
        for f in range(len(st.offsets)):
            r = dec.decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
            assert r["status"] == 0 and r["sample_length"] == 1024
            assert np.array_equal(pcm[first[s] + f], r["s16"]), (s, f)
    eng.close()


def test_probe_sbr_tells_what_an_adts_stream_carries():
    """ADTS says AAC-LC for HE-AAC streams too (implicit signalling): the engine looks at the first frame."""
    eng = Engine(max_streams=4, pcm_format=PCM_S16LE)
    cases = [
        (gen.config(2, n_frames=2), 0),
        (gen.config(3, n_frames=2), 1),
        (gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=2, target_bytes=171, sbr_mode=1), 1),
        (gen.config(4, n_frames=2), 2),
        (gen.config(5, n_frames=2, adts=True), 0),
    ]
    for k, (cfg, want) in enumerate(cases):
        st = gen.generate(cfg, 8800 + k)
        frames, info = demux.adts_index(st.data)
        f0 = st.data[int(frames["offset"][0]): int(frames["offset"][0]) + int(frames["nbytes"][0])]
        assert eng.probe_sbr(info.profile, info.sf_index, info.channel_config, f0) == want, (k, want)
    # the probe leaves no stream behind and the table is still usable
    ids = [eng.open_adts(2, 3, 2) for _ in range(4)]
    assert sorted(ids) == [0, 1, 2, 3]
    eng.close()


@pytest.mark.parametrize("kind", ["adts", "mp4"])
def test_decode_containers_one_call(kind):
    """jaadb_decode_containers: container bytes in, PCM out -- host-side indexing overlapped with the upload, frames in
    frame-major order.  Same frame table and the same PCM bytes as indexing + interleaving + jaadb_decode done by hand, for
    ragged streams, with the PCM in host memory and in device memory; a table that is too small is refused."""
    import torch
    from jaadec_b200 import CONTAINER_ADTS, CONTAINER_MP4, FRAME_RESULT_DTYPE, FRAME_DESC_DTYPE, EngineError
    lens = [9, 14, 5, 14]
    if kind == "adts":
        cfgs = [gen.config(2, n_frames=n, p_transient=0.3) for n in lens]
        streams = [gen.generate(c, 8100 + i) for i, c in enumerate(cfgs)]
        files = [s.data for s in streams]
    else:
        cfgs = [gen.config(5, n_frames=n, p_transient=0.3) for n in lens]
        streams = [gen.generate(c, 8200 + i) for i, c in enumerate(cfgs)]
        files = [genmp4.write_mp4((s.data, s.sizes), bytes([0x11, 0xB0]), 48000, 6, chunk_pattern=(3, 2))[0] for s in streams]
    begin = np.concatenate([[0], np.cumsum([len(f) for f in files])]).astype(np.uint64)
    blob = np.concatenate(files)
    ck = CONTAINER_ADTS if kind == "adts" else CONTAINER_MP4

    def open_all(e):
        if kind == "adts":
            return [e.open_adts(2, cfgs[0].sf_index, cfgs[0].chan_cfg) for _ in lens]
        return [e.open_asc(bytes([0x11, 0xB0])) for _ in lens]

    e1, e2, e3 = (Engine(max_streams=8, pcm_format=PCM_S16LE, chunk_frames=16) for _ in range(3))
    ids = np.asarray(open_all(e1), np.int32)
    assert list(open_all(e2)) == list(ids) and list(open_all(e3)) == list(ids)
    # by hand
    fr, first, _ = (demux.adts_index_many if kind == "adts" else demux.mp4_index_many)(blob, begin, ids)
    fr = demux.interleave(fr, first)
    want_pcm, want_res = e1.decode(blob, fr)
    assert (want_res["status"] == 0).all() and len(fr) == sum(lens)
    # one call, host PCM
    n_max = sum(lens)
    res = np.zeros(n_max, FRAME_RESULT_DTYPE)
    tbl = np.zeros(n_max, FRAME_DESC_DTYPE)
    pcm = np.zeros(want_pcm.nbytes, np.uint8)
    assert e2.decode_containers(ck, blob, begin, ids, pcm, res, frames_out=tbl) == n_max
    assert np.array_equal(tbl, fr) and np.array_equal(res, want_res) and np.array_equal(pcm, want_pcm)
    # one call, device PCM
    d_pcm = torch.zeros(want_pcm.nbytes, dtype=torch.uint8, device="cuda")
    res3 = np.zeros(n_max, FRAME_RESULT_DTYPE)
    assert e3.decode_containers(ck, blob, begin, ids, d_pcm.data_ptr(), res3, pcm_capacity=d_pcm.numel()) == n_max
    torch.cuda.synchronize()
    assert np.array_equal(d_pcm.cpu().numpy(), want_pcm) and np.array_equal(res3, want_res)
    with pytest.raises(EngineError):
        e3.decode_containers(ck, blob, begin, ids, pcm, np.zeros(n_max - 1, FRAME_RESULT_DTYPE))
    for e in (e1, e2, e3):
        e.close()
