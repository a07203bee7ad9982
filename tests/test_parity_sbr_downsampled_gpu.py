"""GPU parity tests for JAAD's down-sampled SBR tool (SURVEY A-20): a stream opened from an AudioSpecificConfig that
does not signal SBR, whose frames carry SBR (and PS) payloads.  JAAD creates the tool at the first payload with the output
rate already fixed by the ASC, so it runs the 32-band synthesis bank (sbr/SynthesisFilterbank32.java) with band tables
for the CORE rate and delivers 1024 samples per frame.  CUDA engine through the C ABI vs the CPU oracle, bit for bit.
"""
import numpy as np
import pytest

import gen
from helpers import Workload, same_float_bits
from jaadec_b200 import Engine, PCM_F32_PLANAR, PCM_S16BE, PCM_S16LE

pytestmark = pytest.mark.gpu


def asc_lc(sf_index: int, chan_cfg: int) -> bytes:
    """AudioSpecificConfig: AOT 2 (AAC-LC), sampling frequency index, channel configuration, GASpecificConfig = 000."""
    v = (2 << 11) | (sf_index << 7) | (chan_cfg << 3)
    return bytes([v >> 8, v & 0xFF])


def ds_cfg(sf_index, chan_cfg, sbr_mode, n_frames, target_bytes):
    return gen.GenConfig(sf_index=sf_index, chan_cfg=chan_cfg, n_frames=n_frames, target_bytes=target_bytes, sbr_mode=sbr_mode,
                         adts=False, sbr_downsampled=True)


CASES = [
    ("stereo_24k", ds_cfg(6, 2, 1, 40, 341), 5),
    ("mono_24k", ds_cfg(6, 1, 1, 40, 171), 4),
    ("mono_ps_24k", ds_cfg(6, 1, 2, 40, 171), 6),
    ("stereo_48k_core", ds_cfg(3, 2, 1, 30, 341), 3),     # a core rate JAAD could not have doubled into a table rate
    ("mono_ps_22k", ds_cfg(7, 1, 2, 30, 160), 3),
]


@pytest.mark.parametrize("tile", [0, 1, 7])
@pytest.mark.parametrize("label,cfg,n_streams", CASES, ids=[c[0] for c in CASES])
def test_downsampled_sbr_float_pcm_bit_exact(label, cfg, n_streams, tile):
    asc = asc_lc(cfg.sf_index, cfg.chan_cfg)
    wl = Workload(cfg, n_streams, base_seed=gen.seed_for(3, 900), with_truth=False, asc=asc)
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=16, pcm_format=PCM_F32_PLANAR, sbr_tile_frames=tile)
    first = wl.frame_bytes(0, 0)
    assert eng.probe_sbr_asc(asc, first) == cfg.sbr_mode
    ids = [eng.open_asc(asc, expect_sbr=cfg.sbr_mode) for _ in range(n_streams)]
    info = eng.stream_info(ids[0])
    assert (info.channels, info.sample_length, info.sample_rate) == (2, 1024, gen_rate(cfg.sf_index))
    frames, index = wl.frame_table(ids)
    pcm, res = eng.decode(wl.blob, frames)
    per = 2 * 1024 * 4
    for i, (s, f) in enumerate(index):
        r = decs[s].decode_frame(wl.frame_bytes(s, f))
        assert res["status"][i] == r["status"] == 0, (label, s, f, res["status"][i], r["status"])
        assert (res["channels"][i], res["sample_length"][i], res["sample_rate"][i]) == (r["channels"], r["sample_length"], r["sample_rate"])
        assert res["pcm_bytes"][i] == per
        got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(2, 1024)
        if not same_float_bits(got, r["f32"]):
            bad = np.argwhere(got.view(np.uint32) != r["f32"].view(np.uint32))
            raise AssertionError((label, s, f, "float pcm differs", float(np.abs(got - r["f32"]).max()), bad[:5].tolist()))
    eng.close()


def gen_rate(sf_index):
    return [96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000][sf_index]


@pytest.mark.parametrize("mode,chan_cfg", [(1, 2), (2, 1)])
@pytest.mark.parametrize("fmt,big", [(PCM_S16LE, False), (PCM_S16BE, True)])
def test_downsampled_sbr_s16_state_across_calls_and_mixed_batch(fmt, big, mode, chan_cfg):
    """int16 output, state carried across calls, and a batch that mixes the two synthesis banks (an ADTS stream of the
    same content decodes up-sampled next to the ASC-opened one)."""
    cfg = ds_cfg(6, chan_cfg, mode, 36, 341 if chan_cfg == 2 else 171)
    asc = asc_lc(6, chan_cfg)
    wl = Workload(cfg, 2, base_seed=31337, with_truth=False, asc=asc)
    up_cfg = gen.GenConfig(sf_index=6, chan_cfg=chan_cfg, n_frames=36, target_bytes=cfg.target_bytes, sbr_mode=mode, adts=False)
    up = Workload(up_cfg, 1, base_seed=555, with_truth=False)
    decs = wl.oracle_decoders() + up.oracle_decoders()
    eng = Engine(max_streams=8, pcm_format=fmt, chunk_frames=16)
    ids = [eng.open_asc(asc, expect_sbr=mode) for _ in range(2)]
    up_id = eng.open_adts(*up.hdr, expect_sbr=mode)
    blob = np.concatenate([wl.blob, up.blob])
    for lo, hi in ((0, 1), (1, 22), (22, 36)):
        fa, ia = wl.frame_table(ids, lo, hi)
        fb, ib = up.frame_table([up_id], lo, hi)
        fb = fb.copy()
        fb["offset"] += len(wl.blob)
        # interleave: one frame of every stream at a time
        rows, index = [], []
        na = len(ids)
        for k in range(hi - lo):
            for j in range(na):
                rows.append(fa[k * na + j]); index.append(ia[k * na + j])
            rows.append(fb[k]); index.append((2, ib[k][1]))
        frames = np.array(rows, fa.dtype)
        pcm, res = eng.decode(blob, frames)
        off = 0
        for i, (s, f) in enumerate(index):
            data = wl.frame_bytes(s, f) if s < 2 else up.frame_bytes(0, f)
            r = decs[s].decode_frame(data, big_endian=big)
            assert res["status"][i] == 0 and r["status"] == 0
            L = 1024 if s < 2 else 2048
            assert res["sample_length"][i] == L == r["sample_length"]
            n = 2 * L * 2
            got = pcm[off:off + n].view(np.int16).reshape(L, 2)
            assert np.array_equal(got, r["s16"]), (s, f)
            off += n
    eng.close()


def test_downsampled_frames_without_payload_keep_the_core_pcm():
    """No SBR payload in a frame of a down-sampled stream: the element's buffers have the core's length, so JAAD neither
    runs the tool nor SBR.upsample -- the core PCM goes out as it is (SCE.java:122-133) and the QMF state stays."""
    cfg = ds_cfg(6, 2, 1, 12, 341)
    asc = asc_lc(6, 2)
    wl = Workload(cfg, 2, base_seed=77, with_truth=False, asc=asc)
    lc = Workload(gen.GenConfig(sf_index=6, chan_cfg=2, n_frames=12, target_bytes=300, adts=False), 1, base_seed=78, with_truth=False)
    blob = np.concatenate([wl.blob, lc.blob])
    frames, index = wl.frame_table([0, 1])
    frames = frames.copy()
    repl = {}
    for i, (s, f) in enumerate(index):
        if s == 0 and f in (4, 5):
            st = lc.streams[0]
            frames["offset"][i] = len(wl.blob) + st.offsets[f]
            frames["nbytes"][i] = st.sizes[f]
            repl[(s, f)] = lc.frame_bytes(0, f)
    decs = wl.oracle_decoders()
    eng = Engine(max_streams=4, pcm_format=PCM_F32_PLANAR, sbr_tile_frames=3)
    ids = [eng.open_asc(asc, expect_sbr=1) for _ in range(2)]
    assert ids == [0, 1]
    pcm, res = eng.decode(blob, frames)
    per = 2 * 1024 * 4
    for i, (s, f) in enumerate(index):
        data = repl.get((s, f), wl.frame_bytes(s, f))
        r = decs[s].decode_frame(data)
        assert res["status"][i] == r["status"] == 0, (s, f)
        got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(2, 1024)
        assert same_float_bits(got, r["f32"]), (s, f)
    eng.close()
