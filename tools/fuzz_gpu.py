#!/usr/bin/env python3
"""Fuzz aid: flip random bits in a share of the frames, decode on the GPU and with the oracle, compare status and PCM.

    python tools/fuzz_gpu.py <config 2|3|4|5> <streams> <frames> <seed> [p_corrupt]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gen, oracle
from helpers import Workload
from jaadec_b200 import Engine, PCM_F32_PLANAR


def run(cfg_no, n, nf, seed, p_corrupt=0.25, tile=0, verbose=True, downsampled=False, gen_over=None, pulse_iso=False):
    """gen_over: generator knobs on top of the BASELINE config (e.g. p_drc, p_pulse: fill elements with dynamic range info,
    pulse data); pulse_iso: JAADB_FLAG_PULSE_ISO on the engine against the oracle's pulseMode 1."""
    over = dict(gen_over or {})
    cfg = gen.config(cfg_no, n_frames=nf, adts=True, **over) if cfg_no == 5 else gen.config(cfg_no, n_frames=nf, **over)
    asc = None
    if downsampled:
        # opened from an AAC-LC ASC at the core rate, SBR / PS implicit: JAAD's down-sampled SBR tool (SURVEY A-20)
        cfg.adts, cfg.sbr_downsampled = False, True
        v = (2 << 11) | (cfg.sf_index << 7) | (cfg.chan_cfg << 3)
        asc = bytes([v >> 8, v & 0xFF])
    wl = Workload(cfg, n, base_seed=seed, with_truth=False, asc=asc)
    rng = np.random.default_rng(seed)
    blob = wl.blob.copy()
    frames, index = wl.frame_table(list(range(n)))
    frames = frames.copy()
    n_mut = 0
    for i, (s, f) in enumerate(index):
        if rng.random() < p_corrupt:
            o, nb = int(frames["offset"][i]), int(frames["nbytes"][i])
            kind = rng.integers(0, 4)
            if kind == 0:      # a few bit flips
                for _ in range(int(rng.integers(1, 4))):
                    b = int(rng.integers(0, nb * 8))
                    blob[o + b // 8] ^= 1 << (7 - b % 8)
            elif kind == 1:    # truncate
                frames["nbytes"][i] = int(rng.integers(1, nb))
            elif kind == 2:    # a burst of random bytes
                a = int(rng.integers(0, nb)); e = min(nb, a + int(rng.integers(1, 16)))
                blob[o + a:o + e] = rng.integers(0, 256, e - a, dtype=np.uint8)
            else:              # flip inside the first 8 bytes (element headers, ics_info, section data)
                b = int(rng.integers(0, min(nb, 8) * 8))
                blob[o + b // 8] ^= 1 << (7 - b % 8)
            n_mut += 1
    decs = [d.set_pulse_mode(1 if pulse_iso else 0) for d in wl.oracle_decoders()]
    from jaadec_b200 import FLAG_PULSE_ISO
    eng = Engine(max_streams=n, pcm_format=PCM_F32_PLANAR, sbr_tile_frames=tile, flags=FLAG_PULSE_ISO if pulse_iso else 0)
    if asc is not None:
        ids = [eng.open_asc(asc, expect_sbr=cfg.sbr_mode) for _ in range(n)]
    else:
        ids = [eng.open_adts(*wl.hdr, expect_sbr=cfg.sbr_mode) for _ in range(n)]
    pcm, res = eng.decode(blob, frames)
    info = eng.stream_info(ids[0])
    per = info.channels * info.sample_length * 4
    offs = np.concatenate([[0], np.cumsum(np.full(len(frames), per))])
    bad_status, bad_pcm, n_err, dead = [], [], 0, set()
    n_foreign = n_unsupported = n_sbr_switch = 0
    hist = {}
    for i, (s, f) in enumerate(index):
        o, nb = int(frames["offset"][i]), int(frames["nbytes"][i])
        r = decs[s].decode_frame(blob[o:o + nb])
        hist[r["status"]] = hist.get(r["status"], 0) + 1
        if s in dead:
            continue
        if cfg.sbr_mode == 0 and decs[s].saw_sbr_payload():
            # a damaged fill element that claims an SBR payload in a stream opened without SBR: JAAD creates an SBR object for
            # the element on the spot, parses the bytes (and usually dies of their end: EOS), and from a payload that does
            # parse on it delivers 2048-sample frames (sbr/SBR.java:98-101); the engine takes that decision at stream_open
            # (include/jaadb200.h) and reads over the element.  The two decoders are different decoders from here on (JAAD's
            # element keeps the SBR object even when the parse died), whatever this frame's two statuses say.
            n_sbr_switch += 1
            dead.add(s)
            continue
        if res["status"][i] == 11 and r["status"] == 0:
            # an element_instance_tag (or element type) the stream did not use before: JAAD decodes the frame against other
            # element objects; the engine reports JAADB_ST_LAYOUT.  Both leave the stream's own objects alone.
            n_foreign += 1
            continue
        if res["status"][i] == 10 and r["status"] != 0:
            # CCE / PCE / gain control: the engine stops at the element, JAAD parses on and dies of something else
            n_unsupported += 1
            continue
        if res["status"][i] != r["status"]:
            bad_status.append((s, f, int(res["status"][i]), r["status"]))
            dead.add(s)   # the two decoders' states have diverged: stop comparing this stream
            continue
        if r["status"] == 13 and cfg.sbr_mode:
            # a Java ArrayIndexOutOfBoundsException (inside the SBR / PS tools in almost every case: get_S_mapped with an odd
            # N_high, parametric-stereo indices past the tables).  Nothing in JAAD catches it -- Decoder.decodeFrame swallows
            # the EOSException only -- so it ends the decode in the middle of the frame's processing (QMF analysis and HF
            # adjustment done, synthesis and the end-of-frame bookkeeping not).  The engine reports the same status and
            # leaves the frame out as a whole; there is no JAAD behaviour "after" to compare with.
            n_err += 1
            dead.add(s)
            continue
        if r["status"] != 0:
            n_err += 1
            continue
        got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(info.channels, info.sample_length)
        if r["f32"].shape != got.shape:
            # an SBR payload in a stream opened without SBR: JAAD switches the stream to 2048-sample output on the spot
            # (sbr/SBR.java:98-101); the engine needs that decision at stream_open (include/jaadb200.h)
            n_sbr_switch += 1
            dead.add(s)
            continue
        ref = np.ascontiguousarray(r["f32"], np.float32)
        # (NaN payloads are the platform's, not the algorithm's: a NaN is a NaN -- damaged SBR data can divide 0 by 0)
        same = (got.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(got) & np.isnan(ref))
        if not same.all():
            bad_pcm.append((s, f))
            dead.add(s)
    if verbose: print("config %d: %d frames, %d mutated, oracle statuses %s" % (cfg_no, len(index), n_mut, dict(sorted(hist.items()))))
    if verbose: print("frames with foreign elements (engine: LAYOUT, JAAD: other objects):", n_foreign)
    if verbose: print("frames failing in both with another code after an unsupported element:", n_unsupported)
    if verbose: print("streams JAAD switched to SBR output after a corrupted frame:", n_sbr_switch)
    if verbose: print("status mismatches:", bad_status[:10], "total", len(bad_status))
    if verbose: print("pcm mismatches:", bad_pcm[:10], "total", len(bad_pcm))
    eng.close()
    return dict(frames=len(index), mutated=n_mut, oracle_statuses=hist, foreign=n_foreign, unsupported=n_unsupported,
                sbr_switch=n_sbr_switch, bad_status=bad_status, bad_pcm=bad_pcm)


if __name__ == "__main__":
    r = run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]) if len(sys.argv) > 5 else 0.25,
            tile=int(os.environ.get("TILE", "0")))
    sys.exit(1 if (r["bad_status"] or r["bad_pcm"]) else 0)
