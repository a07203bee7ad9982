"""GPU parity tests for HE-AAC v1 (AAC-LC core + SBR): CUDA engine through the C ABI vs the CPU oracle.

K1 parses the core, K3 the SBR payload (sequential per element: headers, delta-time coding), K2 produces the core PCM,
K4 runs QMF analysis -> HF generation -> HF adjustment -> QMF synthesis.  Float PCM is compared bit for bit: the SBR
kernels keep JAAD's operation order exactly like the AAC-LC ones.
"""
import numpy as np
import pytest

import gen
from helpers import Workload, same_float_bits
from jaadec_b200 import Engine, PCM_F32_PLANAR, PCM_S16BE, PCM_S16LE

pytestmark = pytest.mark.gpu

CASES = [
    ("c3_stereo_sbr", gen.config(3, n_frames=45), 6),
    ("c3_stereo_sbr_jaad_coupling", gen.config(3, n_frames=45, sbr_quirk=True), 6),
    ("mono_sbr", gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=45, target_bytes=171, sbr_mode=1), 4),
    ("stereo_sbr_32k", gen.GenConfig(sf_index=8, chan_cfg=2, n_frames=30, target_bytes=300, sbr_mode=1), 3),
    # HE-AAC v2: mono core + SBR + parametric stereo (hybrid filterbank, decorrelator, mixing, two synthesis banks)
    ("c4_mono_sbr_ps", gen.config(4, n_frames=45), 8),
    # other core rates: other master / derived band tables, patch layouts and limiter tables
    ("stereo_sbr_22k", gen.GenConfig(sf_index=7, chan_cfg=2, n_frames=30, target_bytes=320, sbr_mode=1), 3),
    ("mono_sbr_ps_16k", gen.GenConfig(sf_index=8, chan_cfg=1, n_frames=30, target_bytes=150, sbr_mode=2), 3),
    ("mono_sbr_ps_32k", gen.GenConfig(sf_index=5, chan_cfg=1, n_frames=30, target_bytes=200, sbr_mode=2), 3),
    # the IPD/OPD extension of parametric stereo (ps/Extension.java, the phase rotation of ps/PSImpl.java:488-660)
    ("c4_mono_sbr_ps_ipdopd", gen.config(4, n_frames=60, ps_ext=0.7), 8),
    ("mono_sbr_ps_ipdopd_16k", gen.GenConfig(sf_index=8, chan_cfg=1, n_frames=40, target_bytes=170, sbr_mode=2, ps_ext=1.0), 3),
]


# the SBR stages walk tiles of frames (0: one tile sized to the workspace budget); every tiling must give the same bits
@pytest.mark.parametrize("tile", [0, 1, 7])
@pytest.mark.parametrize("label,cfg,n_streams", CASES, ids=[c[0] for c in CASES])
def test_sbr_float_pcm_bit_exact(label, cfg, n_streams, tile):
    wl = Workload(cfg, n_streams, base_seed=gen.seed_for(3, 40), with_truth=False)
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=16, pcm_format=PCM_F32_PLANAR, sbr_tile_frames=tile)
    ids = [eng.open_adts(*wl.hdr, expect_sbr=cfg.sbr_mode) for _ in range(n_streams)]
    info = eng.stream_info(ids[0])
    assert (info.channels, info.sample_length) == (2, 2048)
    frames, index = wl.frame_table(ids)
    pcm, res = eng.decode(wl.blob, frames)
    per = 2 * 2048 * 4
    worst = 0.0
    for i, (s, f) in enumerate(index):
        r = decs[s].decode_frame(wl.frame_bytes(s, f))
        assert res["status"][i] == r["status"] == 0, (label, s, f, res["status"][i], r["status"])
        assert (res["channels"][i], res["sample_length"][i], res["sample_rate"][i]) == (r["channels"], r["sample_length"], r["sample_rate"])
        assert res["pcm_bytes"][i] == per
        got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(2, 2048)
        if not same_float_bits(got, r["f32"]):
            worst = max(worst, float(np.abs(got - r["f32"]).max()))
            bad = np.argwhere(got.view(np.uint32) != r["f32"].view(np.uint32))
            raise AssertionError((label, s, f, "float pcm differs", worst, bad[:5].tolist()))
    eng.close()


@pytest.mark.parametrize("cfg_no", [3, 4])
@pytest.mark.parametrize("fmt,big", [(PCM_S16LE, False), (PCM_S16BE, True)])
def test_sbr_s16_and_state_across_calls(fmt, big, cfg_no):
    cfg = gen.config(cfg_no, n_frames=50)
    wl = Workload(cfg, 3, base_seed=4711, with_truth=False)
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=8, pcm_format=fmt, chunk_frames=16)
    ids = [eng.open_adts(*wl.hdr, expect_sbr=cfg.sbr_mode) for _ in range(3)]
    per = 2 * 2048 * 2
    for lo, hi in ((0, 1), (1, 22), (22, 50)):   # headers arrive at frames 0, 20, 40: state and tables carry across calls
        frames, index = wl.frame_table(ids, lo, hi)
        pcm, res = eng.decode(wl.blob, frames)
        for i, (s, f) in enumerate(index):
            r = decs[s].decode_frame(wl.frame_bytes(s, f), big_endian=big)
            assert res["status"][i] == 0 and r["status"] == 0
            got = pcm[i * per:(i + 1) * per].view(np.int16).reshape(2048, 2)
            assert np.array_equal(got, r["s16"]), (s, f)
    eng.close()


@pytest.mark.parametrize("tile", [0, 1, 3])
def test_sbr_frames_without_payload_are_upsampled(tile):
    """A frame that carries no SBR fill element takes SBR.upsample (sample 1 keeps the core value) and leaves the
    QMF state alone; bad frames yield no PCM and the stream continues."""
    cfg = gen.config(3, n_frames=12)
    wl = Workload(cfg, 2, base_seed=99, with_truth=False)
    lc = Workload(gen.GenConfig(sf_index=6, chan_cfg=2, n_frames=12, target_bytes=300), 2, base_seed=199, with_truth=False)
    # stream 0: frames 5 and 6 replaced by plain AAC-LC frames of the same layout (no fill element)
    blob = np.concatenate([wl.blob, lc.blob])
    frames, index = wl.frame_table([0, 1])
    frames = frames.copy()
    repl = {}
    for i, (s, f) in enumerate(index):
        if s == 0 and f in (5, 6):
            st = lc.streams[0]
            frames["offset"][i] = len(wl.blob) + lc.base[0] + st.offsets[f]
            frames["nbytes"][i] = st.sizes[f]
            repl[(s, f)] = lc.frame_bytes(0, f)
        if s == 1 and f == 4:
            frames["nbytes"][i] //= 3
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=4, pcm_format=PCM_F32_PLANAR, sbr_tile_frames=tile)
    ids = [eng.open_adts(*wl.hdr, expect_sbr=1) for _ in range(2)]
    assert ids == [0, 1]
    pcm, res = eng.decode(blob, frames)
    per = 2 * 2048 * 4
    n_bad = 0
    for i, (s, f) in enumerate(index):
        data = repl.get((s, f))
        if data is None:
            o, n = int(frames["offset"][i]), int(frames["nbytes"][i])
            data = blob[o:o + n]
        r = decs[s].decode_frame(data)
        assert res["status"][i] == r["status"], (s, f, res["status"][i], r["status"])
        if r["status"] == 0:
            got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(2, 2048)
            assert same_float_bits(got, r["f32"]), (s, f)
        else:
            n_bad += 1
            assert res["pcm_bytes"][i] == 0
    assert n_bad >= 1
    eng.close()


@pytest.mark.parametrize("tile", [0, 2, 3])
def test_sbr_ragged_streams_and_calls(tile):
    """Streams of different lengths and kinds (HE-AAC v1 stereo, v1 mono, v2 mono+PS) in one engine, decoded in several
    calls of uneven size: runs end inside tiles, tiles start inside runs, and every piece of carried state (QMF history,
    Xsbr rows, v-vectors of both PS banks, decorrelator delay lines) crosses call and tile boundaries."""
    kinds = [
        (gen.config(3, n_frames=31), 1),
        (gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=17, target_bytes=171, sbr_mode=1), 1),
        (gen.config(4, n_frames=26), 2),
        (gen.config(4, n_frames=9), 2),
        (gen.config(3, n_frames=5), 1),
    ]
    streams = [gen.generate(cfg, 7700 + i) for i, (cfg, _) in enumerate(kinds)]
    base = np.concatenate([[0], np.cumsum([len(s.data) for s in streams])]).astype(np.int64)
    blob = np.concatenate([s.data for s in streams])
    eng = Engine(max_streams=8, pcm_format=PCM_F32_PLANAR, sbr_tile_frames=tile)
    ids = [eng.open_adts(2, cfg.sf_index, cfg.chan_cfg, expect_sbr=mode) for cfg, mode in kinds]
    import oracle
    decs = [oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg) for cfg, _ in kinds]
    per = 2 * 2048 * 4
    from jaadec_b200 import FRAME_DESC_DTYPE
    for lo, hi in ((0, 4), (4, 5), (5, 16), (16, 40)):
        rows, index = [], []
        for f in range(lo, hi):
            for s, (cfg, _) in enumerate(kinds):
                if f < cfg.n_frames:
                    rows.append((base[s] + streams[s].offsets[f], streams[s].sizes[f], ids[s]))
                    index.append((s, f))
        if not rows:
            continue
        pcm, res = eng.decode(blob, np.array(rows, FRAME_DESC_DTYPE))
        for i, (s, f) in enumerate(index):
            st = streams[s]
            r = decs[s].decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
            assert res["status"][i] == r["status"] == 0, (s, f, res["status"][i], r["status"])
            got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(2, 2048)
            assert same_float_bits(got, r["f32"]), (tile, s, f)
    eng.close()


@pytest.mark.parametrize("cfg_no", [3, 4])
def test_sbr_long_streams(cfg_no):
    """The benchmark's stream length (235 frames, a header every 20): no drift, no rare-path surprises."""
    cfg = gen.config(cfg_no, n_frames=235)
    wl = Workload(cfg, 3, base_seed=gen.seed_for(cfg_no, 7), with_truth=False)
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=4, pcm_format=PCM_S16LE)
    ids = [eng.open_adts(*wl.hdr, expect_sbr=cfg.sbr_mode) for _ in range(3)]
    frames, index = wl.frame_table(ids)
    pcm, res = eng.decode(wl.blob, frames)
    per = 2 * 2048 * 2
    for i, (s, f) in enumerate(index):
        r = decs[s].decode_frame(wl.frame_bytes(s, f))
        assert res["status"][i] == r["status"] == 0, (s, f)
        assert np.array_equal(pcm[i * per:(i + 1) * per].view(np.int16).reshape(2048, 2), r["s16"]), (s, f)
    eng.close()


def test_ps_parameters_equal_generator_truth():
    """Integer stage of parametric stereo on the GPU (K3: ps_data syntax, Huffman, delta decoding in frequency / time with
    JAAD's stride quirk, envelope borders): the parameters the mixing stage receives are the generator's ground truth --
    no decoder involved -- and the oracle's, for every frame."""
    n_checked, modes = 0, set()
    for seed in range(12):
        cfg = gen.config(4, n_frames=20)
        st = gen.generate(cfg, gen.seed_for(4, 600 + seed), with_truth=True)
        dec = oracle_decoder(cfg)
        eng = Engine(max_streams=2, pcm_format=PCM_S16LE, sbr_tile_frames=6)
        sid = eng.open_adts(2, cfg.sf_index, cfg.chan_cfg, expect_sbr=2)
        frames = np.zeros(cfg.n_frames, dtype=[("offset", "<u8"), ("nbytes", "<u4"), ("stream_id", "<i4")])
        frames["offset"], frames["nbytes"], frames["stream_id"] = st.offsets, st.sizes, sid
        b = eng.batch(frames, st.data.nbytes)
        b.upload(st.data)
        b.decode()
        _, res = b.download()
        assert (res["status"] == 0).all()
        for f in range(cfg.n_frames):
            r = dec.decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
            assert r["status"] == 0
            g, t, tr = b.tap_ps(f), dec.tap_ps(0), st.truth["ps"][f]
            assert g is not None and g["use_ps"] == 1
            ne = int(g["num_env"])
            assert ne == tr[0] == t["num_env"], (seed, f)
            assert np.array_equal(g["border"][:ne + 1], tr[1:2 + ne]) and np.array_equal(g["border"][:ne + 1], t["border"][:ne + 1]), (seed, f)
            assert np.array_equal(g["iid"][:ne], tr[8:178].reshape(5, 34)[:ne, :20]), (seed, f)
            assert np.array_equal(g["icc"][:ne], tr[178:348].reshape(5, 34)[:ne, :20]), (seed, f)
            assert np.array_equal(g["iid"][:ne], t["iid"][:ne, :20]) and np.array_equal(g["icc"][:ne], t["icc"][:ne, :20]), (seed, f)
            modes.add((int(g["iid_mode"]), int(g["icc_mode"])))
            n_checked += 1
        b.close()
        eng.close()
    assert n_checked == 12 * 20 and len(modes) >= 3


def oracle_decoder(cfg):
    import oracle
    return oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg)


def test_ps_ipdopd_parameters_equal_generator_truth_and_state_crosses_calls():
    """The IPD/OPD extension: K3 reads ps_extension (ps/Extension.java:40-59, ExtData.java:17-25), decodes the phase
    indices modulo 8 (PDMode) and hands ps_mix_phase the number of parameter bands that rotate; the values equal the
    generator's ground truth and the oracle's.  Then the same streams in three calls: the phase history (PDData.prev,
    phase_hist) and the imaginary parts of the previous mixing matrices cross call boundaries -- float PCM bit-exact."""
    n_rot = 0
    for seed in range(6):
        cfg = gen.config(4, n_frames=40, ps_ext=0.8)
        st = gen.generate(cfg, gen.seed_for(4, 900 + seed), with_truth=True)
        dec = oracle_decoder(cfg)
        eng = Engine(max_streams=2, pcm_format=PCM_F32_PLANAR, sbr_tile_frames=5)
        sid = eng.open_adts(2, cfg.sf_index, cfg.chan_cfg, expect_sbr=2)
        per = 2 * 2048 * 4
        for lo, hi in ((0, 3), (3, 19), (19, 40)):
            n = hi - lo
            frames = np.zeros(n, dtype=[("offset", "<u8"), ("nbytes", "<u4"), ("stream_id", "<i4")])
            frames["offset"], frames["nbytes"], frames["stream_id"] = st.offsets[lo:hi], st.sizes[lo:hi], sid
            b = eng.batch(frames, st.data.nbytes)
            b.upload(st.data)
            b.decode()
            pcm, res = b.download()
            assert (res["status"] == 0).all()
            for i, f in enumerate(range(lo, hi)):
                r = dec.decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
                assert r["status"] == 0
                g, t, tr = b.tap_ps(i), dec.tap_ps(0), st.truth["ps"][f]
                assert g is not None and g["use_ps"] == 1
                assert int(g["nr_ipdopd_par"]) == tr[433] == t["nr_ipdopd_par"], (seed, f)
                if tr[433]:
                    assert np.array_equal(g["ipd"], tr[348:433].reshape(5, 17)) and np.array_equal(g["ipd"], t["ipd"]), (seed, f)
                    n_rot += 1
                got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(2, 2048)
                assert same_float_bits(got, r["f32"]), (seed, f, float(np.abs(got - r["f32"]).max()))
            b.close()
        eng.close()
    assert n_rot > 60


@pytest.mark.gpu
@pytest.mark.parametrize("cfg_no", [3, 4])
@pytest.mark.parametrize("segment", [2, 7])
def test_sbr_streams_with_a_segmented_core_filterbank(cfg_no, segment):
    """K2 cuts the runs of small batches into segments decoded by different CTAs (a segment re-runs the frame before it for
    the overlap it starts from).  For SBR streams the core PCM it hands to the QMF stages must not change: float PCM of the
    whole stream bit-identical to the oracle, in one call and across calls."""
    cfg = gen.config(cfg_no, n_frames=23)
    wl = Workload(cfg, 3, base_seed=990 + cfg_no, with_truth=False)
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=8, pcm_format=PCM_F32_PLANAR, k2_segment_frames=segment)
    ids = [eng.open_adts(*wl.hdr, expect_sbr=cfg.sbr_mode) for _ in range(3)]
    per = 2 * 2048 * 4
    for lo, hi in ((0, 9), (9, 23)):
        frames, index = wl.frame_table(ids, lo, hi)
        pcm, res = eng.decode(wl.blob, frames)
        for i, (s, f) in enumerate(index):
            r = decs[s].decode_frame(wl.frame_bytes(s, f))
            assert res["status"][i] == 0 and r["status"] == 0
            got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(2, 2048)
            assert same_float_bits(got, r["f32"]), (s, f)
    eng.close()
