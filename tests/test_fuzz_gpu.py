"""Corrupted streams on the GPU: bit flips, truncations and random bursts in a large share of the frames.  Every frame
must end with the status the oracle reports (and, when it still decodes, with bit-identical PCM); a bad frame never
aborts the batch, and later frames of the stream must still match -- i.e. the state a failing frame leaves behind
(window shapes updated before the exception, overlap untouched ...) is JAAD's.

Known, documented deviations are counted separately by tools/fuzz_gpu.py: frames that address
element objects the stream does not own (JAAD decodes them against fresh objects, the engine reports JAADB_ST_LAYOUT),
elements outside the engine's scope (CCE / PCE: both fail, JAAD possibly with a later error), and an SBR payload
showing up in a stream that was opened without SBR."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fuzz_gpu  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg_no,streams,frames,seed,p", [
    (2, 32, 24, 2, 0.4),      # AAC-LC stereo
    (2, 64, 16, 7, 0.5),
    (1, 16, 24, 9, 0.4),      # long windows only
    (5, 16, 16, 8, 0.4),      # 5.1: several elements per frame
])
def test_corrupted_lc_streams_match_the_oracle(cfg_no, streams, frames, seed, p):
    r = fuzz_gpu.run(cfg_no, streams, frames, seed, p, verbose=False)
    assert r["mutated"] > frames and sum(v for k, v in r["oracle_statuses"].items() if k != 0) > 10
    assert r["bad_status"] == [], r["bad_status"]
    assert r["bad_pcm"] == [], r["bad_pcm"]


@pytest.mark.parametrize("cfg_no,seed,tile,ds", [(3, 5, 0, False), (4, 6, 3, False), (3, 15, 5, True), (4, 16, 0, True)])
def test_corrupted_sbr_streams_do_not_derail_the_engine(cfg_no, seed, tile, ds):
    """HE-AAC: the same exercise, also for ASC-opened streams on the down-sampled SBR tool.  Index errors inside JAAD's SBR /
    PS tools (uncaught ArrayIndexOutOfBoundsExceptions) are reported as JAADB_ST_ARRAY_BOUNDS like the oracle does and end the
    comparison of that stream (tools/fuzz_gpu.py says why); a frame or two per thousand mutated ones may still end differently
    (foreign element objects); everything else must match bit for bit."""
    r = fuzz_gpu.run(cfg_no, 24, 24, seed, 0.3, tile=tile, verbose=False, downsampled=ds)
    assert r["mutated"] > 100
    assert r["bad_pcm"] == [], r["bad_pcm"]
    # (engine JAADB_ST_LAYOUT: a damaged tag / id addresses element objects the stream does not own, JAAD decodes them afresh)
    assert all(g == 11 for (_, _, g, _) in r["bad_status"]) and len(r["bad_status"]) <= 1, r["bad_status"]


@pytest.mark.parametrize("cfg_no,seed", [(3, 23), (4, 42)])
def test_fuzz_regressions(cfg_no, seed):
    """Runs that once showed PCM off the oracle: an SBR payload error that has to win over a later
    core error, JAAD's never-cleared E_orig / Q_div / E_curr arrays read through a limiter table that outlived a header
    change, and an element object decoded twice in one frame.  What may remain are the documented status deviations: index
    errors inside JAAD's SBR / PS tools (oracle 13) that the engine does not see coming, and elements outside the engine's
    scope (engine 10).  Since the end of round 2 the two index errors the sweeps kept finding -- get_S_mapped with an odd
    N_high, parametric-stereo indices past the tables -- are reported by the engine too: these runs have no deviation left."""
    r = fuzz_gpu.run(cfg_no, 48, 32, seed, 0.3, verbose=False)
    assert r["mutated"] > 300
    assert r["bad_pcm"] == [], r["bad_pcm"]
    assert r["bad_status"] == [], r["bad_status"]


@pytest.mark.parametrize("cfg_no,seed,over,iso", [
    (2, 61, dict(p_drc=0.9), False),                       # dynamic range info + padding behind the audio elements
    (2, 62, dict(p_drc=0.9, p_pulse=0.7), False),          # ... and pulse data, parsed and dropped (JAAD's mode)
    (2, 63, dict(p_drc=0.5, p_pulse=0.9, pulse_wild=True), True),   # pulses applied (JAADB_FLAG_PULSE_ISO / oracle pulseMode 1)
    (5, 64, dict(p_drc=0.9, p_pulse=0.5), False),          # 5.1: the fill elements follow four audio elements
    (3, 65, dict(p_drc=0.9), False),                       # HE-AAC: a fill element that is not the SBR payload
])
def test_corrupted_fill_elements_and_pulse_data(cfg_no, seed, over, iso):
    """The side paths K1 walks out of line (drc_parse, pulse_apply) under damage: a dynamic-range-info element that is cut
    short or whose flags are flipped must end the frame exactly where JAAD's DRC.decode does (EOS inside the fill element's
    sub-stream, the second group of excluded-channel flags), and the frames after it must still match."""
    r = fuzz_gpu.run(cfg_no, 32, 24, seed, 0.4, verbose=False, gen_over=over, pulse_iso=iso)
    assert r["mutated"] > 200
    if cfg_no == 3:
        assert len(r["bad_status"]) + len(r["bad_pcm"]) <= 3, (r["bad_status"], r["bad_pcm"])
    else:
        assert r["bad_status"] == [], r["bad_status"]
        assert r["bad_pcm"] == [], r["bad_pcm"]
