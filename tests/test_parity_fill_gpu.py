"""Fill elements on the GPU: dynamic_range_info (extension type 11) and padding behind the audio elements.

JAAD parses dynamic range info into an object nobody reads (SyntacticElements.java:181-183,216-224 -> syntax/DRC.java:31-83);
a stream that carries it decodes exactly like the stream without it.  The engine used to answer such frames with
JAADB_ST_UNSUPPORTED_ELEMENT; K1 now walks the element the way DRC.decode does, because the walk can still end the frame
(a read past the element's sub-stream, JAAD's seven-entry excludeMask taking a second group of flags).
"""
import dataclasses

import numpy as np
import pytest

import gen
from helpers import Workload, same_float_bits
from jaadec_b200 import Engine, FRAME_DESC_DTYPE, PCM_F32_PLANAR
from test_oracle_cpu import DRC_CASES, _fil, _splice_before_end

pytestmark = pytest.mark.gpu

CASES = [
    ("c2_stereo", gen.config(2, n_frames=24, p_transient=0.3), 4, 0),
    ("mono_24k", gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=16, target_bytes=171, p_transient=0.3), 2, 0),
    ("c5_51", gen.config(5, n_frames=10, adts=True, p_transient=0.3), 2, 0),
    ("c3_sbr", gen.config(3, n_frames=24), 3, 1),
    ("c4_sbr_ps", gen.config(4, n_frames=24), 3, 2),
]


@pytest.mark.parametrize("label,cfg,n_streams,sbr", CASES, ids=[c[0] for c in CASES])
def test_streams_with_dynamic_range_info_decode_like_the_streams_without(label, cfg, n_streams, sbr):
    wl0 = Workload(cfg, n_streams, base_seed=gen.seed_for(2, 1300), with_truth=False)
    wl1 = Workload(dataclasses.replace(cfg, p_drc=0.7), n_streams, base_seed=gen.seed_for(2, 1300), with_truth=False)
    assert wl1.blob.nbytes > wl0.blob.nbytes
    decs = wl1.oracle_decoders()
    out = []
    for wl in (wl0, wl1):
        eng = Engine(max_streams=16, pcm_format=PCM_F32_PLANAR)
        ids = [eng.open_adts(*wl.hdr, expect_sbr=sbr) for _ in range(n_streams)]
        frames, index = wl.frame_table(ids)
        pcm, res = eng.decode(wl.blob, frames)
        assert (res["status"] == 0).all(), (label, res["status"])
        out.append((pcm, res, index))
        eng.close()
    (p0, r0, _), (p1, r1, index) = out
    assert np.array_equal(r0["pcm_bytes"], r1["pcm_bytes"])
    assert np.array_equal(p0.view(np.uint32), p1.view(np.uint32))                  # the fill elements change nothing
    per = int(r1["pcm_bytes"][0])
    for i, (s, f) in enumerate(index):                                             # ... and the oracle says the same
        r = decs[s].decode_frame(wl1.frame_bytes(s, f))
        assert r["status"] == 0
        got = p1[i * per:(i + 1) * per].view(np.float32).reshape(r["f32"].shape)
        assert same_float_bits(got, r["f32"]), (label, s, f)


def test_dynamic_range_info_that_ends_the_frame():
    """The hand-made elements of tests/test_oracle_cpu.py (valid ones, a second group of excluded-channel flags = JAAD's
    ArrayIndexOutOfBoundsException, a payload that stops short = EOSException), spliced into the frames of one stream each.
    Status per frame as the oracle reports it; the good frames around them bit-identical."""
    import oracle
    cfg = gen.config(2, n_frames=6, p_transient=0.3)
    elements = [_fil("1011" + bits) for _, bits, _ in DRC_CASES] + [_fil("1011" + "0" "0" "1" "0111" "0000" + "0" * 8)]
    blobs, rows, plan = [], [], []
    eng = Engine(max_streams=16, pcm_format=PCM_F32_PLANAR)
    off = 0
    for k, el in enumerate(elements):
        st = gen.generate(cfg, gen.seed_for(2, 1400 + k))
        sid = eng.open_adts(2, cfg.sf_index, cfg.chan_cfg)
        dec = oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg)
        for f in range(cfg.n_frames):
            fr = st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]]
            if f % 2 == 1:
                fr = _splice_before_end(fr, el)
            blobs.append(fr)
            rows.append((off, len(fr), sid))
            off += len(fr)
            plan.append((k, f, dec.decode_frame(fr)))
    frames = np.array(rows, FRAME_DESC_DTYPE)
    pcm, res = eng.decode(np.concatenate(blobs), frames)
    seen = set()
    per = 2 * 1024 * 4                                                             # packed layout: one slot per frame
    for i, (k, f, r) in enumerate(plan):
        assert res["status"][i] == r["status"], (k, f, res["status"][i], r["status"])
        seen.add(int(r["status"]))
        if r["status"] == 0:
            assert res["pcm_bytes"][i] == per
            got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(r["f32"].shape)
            assert same_float_bits(got, r["f32"]), (k, f)
        else:
            assert res["pcm_bytes"][i] == 0
    assert seen == {0, 1, 13}
    eng.close()
