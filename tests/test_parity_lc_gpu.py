"""GPU parity tests for the AAC-LC path: CUDA engine (through the C ABI) vs the CPU oracle.

Bit-exact everywhere: quantised coefficients, scalefactors, sections, the dequantised
spectrum after M/S + intensity stereo, float PCM and int16 PCM (BASELINE.json asks for
bit-exact integers and <= 1e-5 of full scale on float PCM; the engine keeps JAAD's float
operation graph, so the float stage is bit-exact too and is tested as such).
"""
import numpy as np
import pytest

import gen
from helpers import Workload, same_float_bits
from jaadec_b200 import Engine, FLAG_DEBUG_TAPS, PCM_F32_PLANAR, PCM_S16BE, PCM_S16LE

pytestmark = pytest.mark.gpu

CASES = [
    # (label, generator config, streams)
    ("c1_long_only_44k", gen.config(1, n_frames=24), 3),
    ("c2_mixed_48k", gen.config(2, n_frames=40, p_transient=0.3), 6),
    ("c2_no_common", gen.config(2, n_frames=16, p_common_window=0.0, p_transient=0.3), 2),
    ("mono_24k", gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=20, target_bytes=171, p_transient=0.3), 3),
    ("c5_51_48k", gen.config(5, n_frames=16, adts=True, p_transient=0.3), 3),
    ("sf_8k", gen.GenConfig(sf_index=11, chan_cfg=2, n_frames=12, target_bytes=300, p_transient=0.3), 2),
    ("sf_96k", gen.GenConfig(sf_index=0, chan_cfg=2, n_frames=12, target_bytes=400, p_transient=0.3), 2),
]


def run_oracle(wl):
    decs = wl.oracle_decoders()
    out = {}
    for f in range(wl.cfg.n_frames):
        for s in range(wl.n_streams):
            r = decs[s].decode_frame(wl.frame_bytes(s, f))
            taps = []
            if r["status"] == 0:
                el = 0
                while True:
                    t = decs[s].tap_ics(el, 0)
                    if t is None:
                        break
                    taps.append(t)
                    t2 = decs[s].tap_ics(el, 1)
                    if t2 is not None:
                        taps.append(t2)
                    el += 1
            r["taps"] = taps
            out[(s, f)] = r
    return out


@pytest.mark.parametrize("label,cfg,n_streams", CASES, ids=[c[0] for c in CASES])
def test_lc_bit_exact(label, cfg, n_streams):
    wl = Workload(cfg, n_streams, base_seed=gen.seed_for(2, 100))
    ref = run_oracle(wl)
    eng = Engine(max_streams=64, pcm_format=PCM_F32_PLANAR, flags=FLAG_DEBUG_TAPS)
    ids = [eng.open_adts(*wl.hdr) for _ in range(n_streams)]
    frames, index = wl.frame_table(ids)
    b = eng.batch(frames, wl.blob.nbytes)
    b.upload(wl.blob)
    b.decode()
    pcm, res = b.download()
    info0 = eng.stream_info(ids[0])
    ch, ln = info0.channels, info0.sample_length
    per = ch * ln * 4
    n_checked = 0
    for i, (s, f) in enumerate(index):
        r = ref[(s, f)]
        assert res["status"][i] == r["status"], (label, s, f)
        assert r["status"] == 0
        assert (res["channels"][i], res["sample_length"][i], res["sample_rate"][i]) == (r["channels"], r["sample_length"], r["sample_rate"])
        # integer stage + dequantised spectrum vs oracle AND vs the generator's own ground truth
        truth = wl.streams[s].truth
        for c, t in enumerate(r["taps"]):
            g = b.tap(i, c)
            assert np.array_equal(g["q"], t["q"]), (label, s, f, c, "q")
            assert np.array_equal(g["q"], truth["q"][f, c]), (label, s, f, c, "q vs generator")
            assert np.array_equal(g["sfbcb"], t["sfbcb"]), (label, s, f, c, "sfbcb")
            assert np.array_equal(g["sfidx"], t["sfidx"]), (label, s, f, c, "sfidx")
            assert np.array_equal(g["sfidx"], truth["sfidx"][f, c]), (label, s, f, c, "sfidx vs generator")
            assert np.array_equal(g["info"][[1, 2, 4, 5]], t["info"][[1, 2, 4, 5]]), (label, s, f, c, "info")
            assert np.array_equal(g["info"][6:16], t["info"][6:16]), (label, s, f, c, "groups/ms")
            assert same_float_bits(g["spec"], t["spec"]), (label, s, f, c, "spectrum")
        got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(ch, ln)
        assert same_float_bits(got, r["f32"]), (label, s, f, "float pcm", np.abs(got - r["f32"]).max())
        n_checked += 1
    assert n_checked == len(index)
    b.close()
    eng.close()


@pytest.mark.parametrize("fmt,big", [(PCM_S16LE, False), (PCM_S16BE, True)])
def test_s16_identical_and_state_carries_across_calls(fmt, big):
    cfg = gen.config(2, n_frames=30, p_transient=0.3)
    wl = Workload(cfg, 4, base_seed=777, with_truth=False)
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=16, pcm_format=fmt)
    ids = [eng.open_adts(*wl.hdr) for _ in range(4)]
    # three calls of uneven size: the overlap / window-shape state must carry over
    for lo, hi in ((0, 1), (1, 17), (17, 30)):
        frames, index = wl.frame_table(ids, lo, hi)
        pcm, res = eng.decode(wl.blob, frames)
        per = 2 * 1024 * 2
        for i, (s, f) in enumerate(index):
            r = decs[s].decode_frame(wl.frame_bytes(s, f), big_endian=big)
            assert res["status"][i] == 0 and r["status"] == 0
            assert res["pcm_bytes"][i] == per
            got = pcm[i * per:(i + 1) * per].view(np.int16).reshape(1024, 2)
            assert np.array_equal(got, r["s16"]), (s, f)
    eng.close()


def test_bad_frames_do_not_poison_the_batch():
    cfg = gen.config(2, n_frames=12, p_transient=0.3)
    wl = Workload(cfg, 3, base_seed=4242, with_truth=False)
    blob = wl.blob.copy()
    frames, index = wl.frame_table([0, 1, 2])
    frames = frames.copy()
    # stream 1 frame 4: truncated (EOS); stream 2 frame 6: 3 bytes (EOS at the ADIF peek);
    # stream 0 frame 5: corrupt the first bytes (whatever status the oracle reports)
    mutate = {}
    for i, (s, f) in enumerate(index):
        if (s, f) == (1, 4):
            frames["nbytes"][i] //= 2
        if (s, f) == (2, 6):
            frames["nbytes"][i] = 3
        if (s, f) == (0, 5):
            o = int(frames["offset"][i])
            blob[o:o + 6] = [0x21, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF]
        mutate[(s, f)] = (int(frames["offset"][i]), int(frames["nbytes"][i]))
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=8, pcm_format=PCM_F32_PLANAR)
    ids = [eng.open_adts(*wl.hdr) for _ in range(3)]
    assert ids == [0, 1, 2]
    pcm, res = eng.decode(blob, frames)
    per = 2 * 1024 * 4
    statuses = []
    for i, (s, f) in enumerate(index):
        o, n = mutate[(s, f)]
        r = decs[s].decode_frame(blob[o:o + n])
        statuses.append(r["status"])
        assert res["status"][i] == r["status"], (s, f, res["status"][i], r["status"])
        if r["status"] == 0:
            got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(2, 1024)
            assert same_float_bits(got, r["f32"]), (s, f)
        else:
            assert res["pcm_bytes"][i] == 0
    assert sum(1 for x in statuses if x != 0) >= 2
    eng.close()


def test_asc_open_5_1_raw_frames():
    cfg = gen.config(5, n_frames=10, p_transient=0.3)  # raw frames, as MP4 samples
    asc = bytes([0x11, 0xB0])
    wl = Workload(cfg, 2, base_seed=31337, with_truth=False, asc=asc)
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=4, pcm_format=PCM_S16LE)
    ids = [eng.open_asc(asc) for _ in range(2)]
    info = eng.stream_info(ids[0])
    assert (info.channels, info.sample_rate, info.sample_length) == (6, 48000, 1024)
    frames, index = wl.frame_table(ids)
    pcm, res = eng.decode(wl.blob, frames)
    per = 6 * 1024 * 2
    for i, (s, f) in enumerate(index):
        r = decs[s].decode_frame(wl.frame_bytes(s, f))
        assert res["status"][i] == 0 and r["status"] == 0
        got = pcm[i * per:(i + 1) * per].view(np.int16).reshape(1024, 6)
        assert np.array_equal(got, r["s16"]), (s, f)
    eng.close()


GOLDEN_CASES = ["lc_c1_long_44k", "lc_c2_mixed_48k", "lc_mono_24k", "lc_c5_51_raw", "sbr_c3_stereo", "sbr_mono", "ps_c4_mono",
                "sbr_ds_stereo", "ps_ds_mono", "lc_pns_48k", "ps_ipdopd_mono", "lc_drc_48k"]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_engine_reproduces_committed_golden(name):
    """CUDA engine vs tests/golden/*.npz (made by tests/golden/make_golden.py): int16 PCM identical, float PCM
    bit-identical (SHA-256 of the float bits), quantised coefficients and scalefactors equal to the generator truth."""
    import hashlib
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    asc = g["asc"].tobytes()
    n_streams = int(g["frame_stream"].max()) + 1
    frames = np.zeros(len(g["frame_offset"]), dtype=[("offset", "<u8"), ("nbytes", "<u4"), ("stream_id", "<i4")])
    frames["offset"], frames["nbytes"] = g["frame_offset"], g["frame_nbytes"]
    for fmt in (PCM_S16LE, PCM_F32_PLANAR):
        eng = Engine(max_streams=8, pcm_format=fmt, flags=FLAG_DEBUG_TAPS)
        sbr = int(g["sbr"][0]) if "sbr" in g.files else 0
        ids = [eng.open_asc(asc, expect_sbr=sbr) if len(asc) else eng.open_adts(*[int(x) for x in g["hdr"]], expect_sbr=sbr) for _ in range(n_streams)]
        frames["stream_id"] = np.asarray(ids)[g["frame_stream"]]
        b = eng.batch(frames, g["blob"].nbytes)
        b.upload(g["blob"])
        b.decode()
        pcm, res = b.download()
        assert (res["status"] == 0).all()
        n, ln, ch = g["s16"].shape
        if fmt == PCM_S16LE:
            assert np.array_equal(pcm.view(np.int16).reshape(n, ln, ch), g["s16"])
        else:
            assert hashlib.sha256(pcm.tobytes()).digest() == g["f32_sha256"].tobytes()
            seen = [0] * n_streams
            for i, s in enumerate(g["frame_stream"]):
                for c in range(g["truth_q"].shape[2]):
                    t = b.tap(i, c, want_spec=False)
                    assert np.array_equal(t["q"], g["truth_q"][s, seen[s], c])
                    assert np.array_equal(t["sfidx"], g["truth_sfidx"][s, seen[s], c])
                seen[s] += 1
        b.close()
        eng.close()


def test_one_call_decode_pipelines_chunks_and_honours_pcm_offsets():
    """jaadb_decode cuts the frame array into chunks (here 10 frames, so chunks straddle frames of all streams and
    leave a remainder) and overlaps PCM download with the next chunk; caller-placed, non-monotonic PCM offsets take
    the single-range path.  Both must equal the oracle frame by frame."""
    cfg = gen.config(2, n_frames=23, p_transient=0.3)
    wl = Workload(cfg, 5, base_seed=2024, with_truth=False)
    decs = wl.oracle_decoders()
    per = 2 * 1024 * 2
    ref = {}
    for f in range(cfg.n_frames):
        for s in range(5):
            ref[(s, f)] = decs[s].decode_frame(wl.frame_bytes(s, f))["s16"]
    for mode in ("packed", "reversed"):
        eng = Engine(max_streams=8, pcm_format=PCM_S16LE, chunk_frames=10)
        ids = [eng.open_adts(*wl.hdr) for _ in range(5)]
        frames, index = wl.frame_table(ids)
        n = len(frames)
        offs = None
        if mode == "reversed":
            offs = (np.arange(n)[::-1] * per).astype(np.uint64)
        pcm, res = eng.decode(wl.blob, frames, pcm_out=np.zeros(n * per, np.uint8), pcm_offsets=offs)
        assert (res["status"] == 0).all() and (res["pcm_bytes"] == per).all()
        for i, (s, f) in enumerate(index):
            o = i * per if offs is None else int(offs[i])
            got = pcm[o:o + per].view(np.int16).reshape(1024, 2)
            assert np.array_equal(got, ref[(s, f)]), (mode, s, f)
        eng.close()


PNS_CASES = [
    ("stereo_48k", gen.config(2, n_frames=36, p_transient=0.3, p_pns=0.2), 4),
    ("stereo_48k_no_common", gen.config(2, n_frames=16, p_transient=0.4, p_common_window=0.0, p_pns=0.5), 2),
    ("mono_24k", gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=20, target_bytes=171, p_transient=0.3, p_pns=0.3), 3),
    ("c5_51_48k", gen.config(5, n_frames=12, adts=True, p_transient=0.3, p_pns=0.15), 2),
    ("sf_8k", gen.GenConfig(sf_index=11, chan_cfg=2, n_frames=12, target_bytes=300, p_transient=0.3, p_pns=0.3), 2),
]


@pytest.mark.parametrize("label,cfg,n_streams", PNS_CASES, ids=[c[0] for c in PNS_CASES])
@pytest.mark.parametrize("segment", [0, 3])
def test_pns_bit_exact_in_one_batch_and_across_calls(label, cfg, n_streams, segment):
    """Perceptual noise substitution (codebook 13, ICStream.java:241-257): every stream owns a generator seeded like
    JAAD's static one, i.e. the oracle decoding that stream alone.  All streams in one batch (the frames of a stream are
    parsed in parallel and the generator state of each frame comes from a prefix over the draw counts), then the same
    streams again in three calls -- the generator state is carried in the stream's state."""
    wl = Workload(cfg, n_streams, base_seed=gen.seed_for(2, 500))
    assert sum(int((s.truth["sfbcb"] == 13).sum()) for s in wl.streams) > 50
    ref = run_oracle(wl)
    for calls in ([(0, cfg.n_frames)], [(0, 1), (1, 9), (9, cfg.n_frames)]):
        eng = Engine(max_streams=64, pcm_format=PCM_F32_PLANAR, flags=FLAG_DEBUG_TAPS, k2_segment_frames=segment)
        ids = [eng.open_adts(*wl.hdr) for _ in range(n_streams)]
        info0 = eng.stream_info(ids[0])
        ch, ln = info0.channels, info0.sample_length
        per = ch * ln * 4
        for lo, hi in calls:
            frames, index = wl.frame_table(ids, lo, hi)
            b = eng.batch(frames, wl.blob.nbytes)
            b.upload(wl.blob)
            b.decode()
            pcm, res = b.download()
            for i, (s, f) in enumerate(index):
                r = ref[(s, f)]
                assert res["status"][i] == 0 and r["status"] == 0, (label, s, f, res["status"][i])
                truth = wl.streams[s].truth
                for c, t in enumerate(r["taps"]):
                    g = b.tap(i, c)
                    assert np.array_equal(g["sfbcb"], t["sfbcb"]), (label, s, f, c, "sfbcb")
                    assert np.array_equal(g["sfidx"], t["sfidx"]), (label, s, f, c, "sfidx")
                    assert np.array_equal(g["sfidx"], truth["sfidx"][f, c]), (label, s, f, c, "sfidx vs generator")
                    assert np.array_equal(g["q"], truth["q"][f, c]), (label, s, f, c, "q vs generator")
                    assert same_float_bits(g["spec"], t["spec"]), (label, s, f, c, "spectrum", np.abs(g["spec"] - t["spec"]).max())
                got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(ch, ln)
                assert same_float_bits(got, r["f32"]), (label, s, f, "float pcm")
            b.close()
        eng.close()


TNS_ISO_CASES = [
    ("stereo_48k", gen.config(2, n_frames=30, p_transient=0.3, p_tns=0.7, tns_mild=True), 4),
    ("stereo_48k_wild", gen.config(2, n_frames=16, p_transient=0.3, p_tns=1.0), 2),      # orders up to 20, any coefficient
    ("mono_24k", gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=20, target_bytes=171, p_transient=0.3, p_tns=0.7, tns_mild=True), 2),
    ("c5_51_48k", gen.config(5, n_frames=10, adts=True, p_transient=0.3, p_tns=0.7, tns_mild=True), 2),
    ("sf_8k", gen.GenConfig(sf_index=11, chan_cfg=2, n_frames=12, target_bytes=300, p_transient=0.3, p_tns=0.7, tns_mild=True), 2),
]


@pytest.mark.parametrize("label,cfg,n_streams", TNS_ISO_CASES, ids=[c[0] for c in TNS_ISO_CASES])
def test_tns_iso_mode_matches_the_oracle(label, cfg, n_streams):
    """JAADB_TNS_ISO: the all-pole filter of 14496-3 4.6.9.3 between the stereo tools and the filterbank.  The oracle's
    ISO restatement does the same binary32 operations in the same order, so spectrum and PCM are bit-identical (the
    1e-5-of-full-scale budget of BASELINE.json is checked as well); JAAD mode on the same engine build stays JAAD's."""
    from jaadec_b200 import TNS_ISO
    wl = Workload(cfg, n_streams, base_seed=gen.seed_for(2, 900))
    assert sum(int(s.truth["tns"][:, :, 0].sum()) for s in wl.streams) > 5
    decs = [d.set_tns_mode(1) for d in wl.oracle_decoders()]
    eng = Engine(max_streams=16, pcm_format=PCM_F32_PLANAR, flags=FLAG_DEBUG_TAPS, tns_mode=TNS_ISO)
    ids = [eng.open_adts(*wl.hdr) for _ in range(n_streams)]
    frames, index = wl.frame_table(ids)
    b = eng.batch(frames, wl.blob.nbytes)
    b.upload(wl.blob)
    b.decode()
    pcm, res = b.download()
    info0 = eng.stream_info(ids[0])
    ch, ln = info0.channels, info0.sample_length
    per = ch * ln * 4
    n_filtered = 0
    for i, (s, f) in enumerate(index):
        r = decs[s].decode_frame(wl.frame_bytes(s, f))
        assert res["status"][i] == 0 and r["status"] == 0
        taps, el = [], 0
        while True:
            t = decs[s].tap_ics(el, 0)
            if t is None:
                break
            taps.append(t)
            t2 = decs[s].tap_ics(el, 1)
            if t2 is not None:
                taps.append(t2)
            el += 1
        assert len(taps) == wl.streams[s].truth["q"].shape[1]
        for c, t in enumerate(taps):
            g = b.tap(i, c)
            assert same_float_bits(g["spec"], t["spec"]), (label, s, f, c, "spectrum after TNS")
            n_filtered += int(wl.streams[s].truth["tns"][f, c, 0])
        got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(ch, ln)
        ok = np.isfinite(r["f32"])
        assert np.array_equal(ok, np.isfinite(got))
        assert np.abs(got[ok] - r["f32"][ok]).max(initial=0.0) <= 1e-5 * 32768.0 * max(1.0, np.abs(r["f32"][ok]).max(initial=0.0) / 32768.0)
        assert same_float_bits(got, r["f32"]), (label, s, f, "float pcm")
    assert n_filtered > 5
    b.close()
    eng.close()


SEGMENT_CASES = [("c1_long_only_44k", gen.config(1, n_frames=40), 1), ("c2_mixed_48k", gen.config(2, n_frames=33, p_transient=0.3), 3),
                 ("mono_24k", gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=21, target_bytes=171, p_transient=0.3), 2),
                 ("c5_51_48k", gen.config(5, n_frames=14, adts=True, p_transient=0.3), 2)]


@pytest.mark.parametrize("label,cfg,n_streams", SEGMENT_CASES, ids=[c[0] for c in SEGMENT_CASES])
@pytest.mark.parametrize("segment", [1, 2, 5])
def test_segmented_filterbank_is_bit_identical(label, cfg, n_streams, segment):
    """Few streams: the filterbank kernel cuts a stream's frames into segments, one CTA each; a segment re-runs the
    frame before it for the overlap it starts from.  Same float PCM bits as the oracle, also across calls and around a
    frame that fails (its channels keep the overlap of the frame before)."""
    wl = Workload(cfg, n_streams, base_seed=gen.seed_for(2, 1300), with_truth=False)
    blob = wl.blob.copy()
    bad = (0, min(7, cfg.n_frames - 2))
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=8, pcm_format=PCM_F32_PLANAR, k2_segment_frames=segment)
    ids = [eng.open_adts(*wl.hdr) for _ in range(n_streams)]
    info0 = eng.stream_info(ids[0])
    per = info0.channels * info0.sample_length * 4
    for lo, hi in ((0, 11), (11, cfg.n_frames)):
        frames, index = wl.frame_table(ids, lo, hi)
        frames = frames.copy()
        for i, (s, f) in enumerate(index):
            if (s, f) == bad:
                frames["nbytes"][i] //= 2
        pcm, res = eng.decode(blob, frames)
        for i, (s, f) in enumerate(index):
            r = decs[s].decode_frame(blob[int(frames["offset"][i]): int(frames["offset"][i]) + int(frames["nbytes"][i])])
            assert res["status"][i] == r["status"], (label, s, f)
            if r["status"] == 0:
                got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(info0.channels, info0.sample_length)
                assert same_float_bits(got, r["f32"]), (label, segment, s, f)
            else:
                assert (s, f) == bad and res["pcm_bytes"][i] == 0 and not pcm[i * per:(i + 1) * per].any()
    eng.close()


@pytest.mark.parametrize("cfg_no,n_frames", [(2, 30), (3, 20), (4, 20)])
def test_device_resident_blob_and_pcm(cfg_no, n_frames):
    """jaadb_decode with the blob and the PCM buffer in device memory (torch tensors): the kernels write the PCM in place
    (no staging buffer, no PCIe transfer of PCM); same bytes as the host-buffer call and as the oracle, chunked
    (chunks of 16 frames) and across two calls."""
    import torch
    cfg = gen.config(cfg_no, n_frames=n_frames, p_transient=0.3)
    wl = Workload(cfg, 4, base_seed=gen.seed_for(cfg_no, 321), with_truth=False)
    decs = wl.oracle_decoders()
    engs = [Engine(max_streams=8, pcm_format=PCM_S16LE, chunk_frames=16) for _ in range(2)]
    ids = [[e.open_adts(*wl.hdr, expect_sbr=cfg.sbr_mode) for _ in range(4)] for e in engs]
    info = engs[0].stream_info(ids[0][0])
    per = info.channels * info.sample_length * 2
    d_blob = torch.from_numpy(wl.blob).cuda()
    for lo, hi in ((0, 7), (7, n_frames)):
        frames, index = wl.frame_table(ids[0], lo, hi)
        n = len(frames)
        d_pcm = torch.full((n * per,), 0x55, dtype=torch.uint8, device="cuda")
        res_d = engs[0].decode_ptr(d_blob.data_ptr(), wl.blob.nbytes, frames, d_pcm.data_ptr(), d_pcm.numel())
        torch.cuda.synchronize()
        got_d = d_pcm.cpu().numpy()
        frames_h, _ = wl.frame_table(ids[1], lo, hi)
        got_h, res_h = engs[1].decode(wl.blob, frames_h)
        assert np.array_equal(res_d["status"], res_h["status"]) and (res_d["status"] == 0).all()
        assert np.array_equal(got_d, got_h)
        for i, (s, f) in enumerate(index):
            r = decs[s].decode_frame(wl.frame_bytes(s, f))
            assert np.array_equal(got_d[i * per:(i + 1) * per].view(np.int16).reshape(info.sample_length, info.channels), r["s16"]), (s, f)
    for e in engs:
        e.close()


PULSE_CASES = [
    ("stereo_48k", gen.config(2, n_frames=30, p_transient=0.3, p_pulse=0.8, pulse_wild=True), 4),
    ("stereo_48k_pns_tns", gen.config(2, n_frames=16, p_transient=0.2, p_pulse=0.9, pulse_wild=True, p_pns=0.2, p_tns=0.5, tns_mild=True), 2),
    ("c1_44k_long_only", gen.config(1, n_frames=20, p_pulse=1.0), 2),
    ("mono_24k", gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=20, target_bytes=171, p_transient=0.3, p_pulse=0.8, pulse_wild=True), 2),
    ("c5_51_48k", gen.config(5, n_frames=10, adts=True, p_transient=0.3, p_pulse=0.8, pulse_wild=True), 2),
    ("sf_8k", gen.GenConfig(sf_index=11, chan_cfg=2, n_frames=12, target_bytes=300, p_transient=0.3, p_pulse=0.8, pulse_wild=True), 2),
]


@pytest.mark.parametrize("label,cfg,n_streams", PULSE_CASES, ids=[c[0] for c in PULSE_CASES])
@pytest.mark.parametrize("iso", [True, False])
def test_pulse_data_jaad_and_iso_mode(label, cfg, n_streams, iso):
    """pulse_data (ICStream.java:148-170).  JAAD parses it and never applies it (ICStream.java:17): the default mode does the
    same and must match the oracle's default.  JAADB_FLAG_PULSE_ISO adds the pulses to the quantised coefficients (ISO/IEC
    14496-3 4.6.3.3), like the oracle's pulseMode 1 (pinned against FFmpeg in tests/test_oracle_vs_libavcodec_cpu.py): the
    quantised coefficients then equal the generator's ground truth, spectrum and PCM are bit-identical to the oracle's.
    The generator also drops pulses on bands without spectral data and past max_sfb; both sides leave those alone."""
    from jaadec_b200 import FLAG_PULSE_ISO, TNS_ISO
    wl = Workload(cfg, n_streams, base_seed=gen.seed_for(2, 1100))
    decs = [d.set_pulse_mode(1 if iso else 0).set_tns_mode(1 if iso else 0) for d in wl.oracle_decoders()]
    eng = Engine(max_streams=16, pcm_format=PCM_F32_PLANAR, flags=FLAG_DEBUG_TAPS | (FLAG_PULSE_ISO if iso else 0),
                 tns_mode=TNS_ISO if iso else 0)
    ids = [eng.open_adts(*wl.hdr) for _ in range(n_streams)]
    frames, index = wl.frame_table(ids)
    b = eng.batch(frames, wl.blob.nbytes)
    b.upload(wl.blob)
    b.decode()
    pcm, res = b.download()
    info0 = eng.stream_info(ids[0])
    ch, ln = info0.channels, info0.sample_length
    per = ch * ln * 4
    n_pulsed = 0
    for i, (s, f) in enumerate(index):
        r = decs[s].decode_frame(wl.frame_bytes(s, f))
        assert res["status"][i] == 0 and r["status"] == 0
        taps, el = [], 0
        while True:
            t = decs[s].tap_ics(el, 0)
            if t is None:
                break
            taps.append(t)
            t2 = decs[s].tap_ics(el, 1)
            if t2 is not None:
                taps.append(t2)
            el += 1
        truth = wl.streams[s].truth
        assert len(taps) == truth["q"].shape[1]
        for c, t in enumerate(taps):
            g = b.tap(i, c)
            assert np.array_equal(g["q"], t["q"]), (label, s, f, c, "q vs oracle")
            if iso:
                assert np.array_equal(g["q"], truth["q"][f, c]), (label, s, f, c, "q vs generator")
            else:
                n_pulsed += int((g["q"] != truth["q"][f, c]).sum())
            assert same_float_bits(g["spec"], t["spec"]), (label, s, f, c, "spectrum")
        got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(ch, ln)
        assert same_float_bits(got, r["f32"]), (label, s, f, "float pcm")
    assert iso or n_pulsed > 10     # the streams do carry pulses: without them the transmitted coefficients are not the truth
    b.close()
    eng.close()


@pytest.mark.parametrize("lanes_log2", [0, 1, 3, 5])
def test_parse_kernel_frames_per_warp(lanes_log2, monkeypatch):
    """K1 gives a small batch fewer frames per warp (down to one: the latency of a frame parsed alone instead of the lock-step
    time of 32) and a batch that fills the GPU all 32 lanes; the mapping is by batch size.  JAADB_K1_LANES_LOG2 forces it, so
    that every mapping sees the same streams: statuses, quantised coefficients and PCM identical to the oracle for each."""
    monkeypatch.setenv("JAADB_K1_LANES_LOG2", str(lanes_log2))
    cfg = gen.config(5, n_frames=12, adts=True, p_transient=0.3, p_pns=0.1, p_pulse=0.3, p_drc=0.5)
    wl = Workload(cfg, 7, base_seed=gen.seed_for(2, 1500))       # 84 frames: not a multiple of any warp load
    ref = run_oracle(wl)
    eng = Engine(max_streams=16, pcm_format=PCM_F32_PLANAR, flags=FLAG_DEBUG_TAPS)
    ids = [eng.open_adts(*wl.hdr) for _ in range(7)]
    frames, index = wl.frame_table(ids)
    b = eng.batch(frames, wl.blob.nbytes)
    b.upload(wl.blob)
    b.decode()
    pcm, res = b.download()
    per = 6 * 1024 * 4
    for i, (s, f) in enumerate(index):
        r = ref[(s, f)]
        assert res["status"][i] == 0 and r["status"] == 0
        for c, t in enumerate(r["taps"]):
            g = b.tap(i, c)
            assert np.array_equal(g["q"], t["q"]) and np.array_equal(g["sfidx"], t["sfidx"]), (lanes_log2, s, f, c)
        assert same_float_bits(pcm[i * per:(i + 1) * per].view(np.float32).reshape(6, 1024), r["f32"]), (lanes_log2, s, f)
    b.close()
    eng.close()
