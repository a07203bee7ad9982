#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz: committed input/output vectors of the AAC decode path.

JAAD ships no golden vectors, known-answer tests or fixtures for this path
(the only test plays an external file to the sound card), and
there is no JVM in the build image, so these vectors are produced by the
committed generator (gen/) and the C++ restatement of JAAD (oracle/).  They pin
(a) the generator's bitstreams byte for byte, (b) the oracle's output against
accidental drift, and (c) the CUDA engine on the GPU box, where neither
/root/reference nor a rebuild of history is available.

Each case stores the compressed frames themselves (small), the frame table,
the int16 PCM of every frame, a SHA-256 of the float PCM bits, and per-ICS
integer ground truth straight from the generator (independent of any decoder).

    python tests/golden/make_golden.py            # rewrite the fixtures
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import gen      # noqa: E402
import oracle   # noqa: E402


def cases():
    return {
        # name: (generator config, n_streams, asc or None)
        "lc_c1_long_44k": (gen.config(1, n_frames=6), 2, None),
        "lc_c2_mixed_48k": (gen.config(2, n_frames=12, p_transient=0.35), 3, None),
        "lc_mono_24k": (gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=8, target_bytes=171, p_transient=0.3), 2, None),
        "lc_c5_51_raw": (gen.config(5, n_frames=5, p_transient=0.3), 2, bytes([0x11, 0xB0])),
        # HE-AAC v1: 24 kHz core + SBR, header at frame 0 and (possibly changed) at frame 20
        "sbr_c3_stereo": (gen.config(3, n_frames=24), 2, None),
        "sbr_mono": (gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=22, target_bytes=171, sbr_mode=1), 2, None),
        # HE-AAC v2: mono core + SBR + parametric stereo
        "ps_c4_mono": (gen.config(4, n_frames=24), 2, None),
        # down-sampled SBR (SURVEY A-20): opened from an AAC-LC ASC at the core rate, SBR / PS arrive implicitly
        "sbr_ds_stereo": (gen.GenConfig(sf_index=6, chan_cfg=2, n_frames=22, target_bytes=341, sbr_mode=1, adts=False,
                                        sbr_downsampled=True), 2, bytes([0x13, 0x10])),
        "ps_ds_mono": (gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=22, target_bytes=171, sbr_mode=2, adts=False,
                                     sbr_downsampled=True), 2, bytes([0x13, 0x08])),
        # round 2: perceptual noise substitution (every stream decoded by its own Decoder = alone in a fresh JVM, which is
        # how tools/jaad_verify runs JAAD: one process per file) and the IPD/OPD extension of parametric stereo
        "lc_pns_48k": (gen.config(2, n_frames=12, p_transient=0.35, p_pns=0.25), 3, None),
        "ps_ipdopd_mono": (gen.config(4, n_frames=30, ps_ext=0.8), 2, None),
        # fill elements behind the audio elements: dynamic range info (JAAD parses it and drops it, syntax/DRC.java) + padding
        "lc_drc_48k": (gen.config(2, n_frames=12, p_transient=0.3, p_drc=0.7), 2, None),
    }


def build_case(name, cfg, n_streams, asc):
    seed0 = gen.seed_for(9, sum(name.encode()) % 500)
    streams = [gen.generate(cfg, seed0 + s, with_truth=True) for s in range(n_streams)]
    blob = np.concatenate([s.data for s in streams])
    base = np.concatenate([[0], np.cumsum([len(s.data) for s in streams])]).astype(np.int64)
    rows = []
    s16, f32_sha, status = [], hashlib.sha256(), []
    decs = [oracle.Decoder.create_asc(asc) if asc else oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg)
            for _ in range(n_streams)]
    for f in range(cfg.n_frames):
        for s in range(n_streams):
            st = streams[s]
            rows.append((base[s] + st.offsets[f], st.sizes[f], s))
            r = decs[s].decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])
            status.append(r["status"])
            assert r["status"] == 0, (name, s, f, r["status"])
            s16.append(r["s16"])
            f32_sha.update(np.ascontiguousarray(r["f32"], np.float32).tobytes())
    extra = {}
    if cfg.sbr_mode:
        extra["truth_sbr"] = np.stack([s.truth["sbr"] for s in streams])
    if cfg.sbr_mode > 1:
        extra["truth_ps"] = np.stack([s.truth["ps"] for s in streams])
    out = dict(
        sbr=np.array([cfg.sbr_mode], np.int32),
        blob=blob,
        frame_offset=np.array([r[0] for r in rows], np.int64),
        frame_nbytes=np.array([r[1] for r in rows], np.int32),
        frame_stream=np.array([r[2] for r in rows], np.int32),
        hdr=np.array([2, cfg.sf_index, cfg.chan_cfg], np.int32),
        asc=np.frombuffer(asc or b"", np.uint8),
        adts=np.array([int(cfg.adts)], np.int32),
        s16=np.stack(s16),                                      # [frames, 1024, channels]
        f32_sha256=np.frombuffer(f32_sha.digest(), np.uint8),
        truth_q=np.stack([s.truth["q"] for s in streams]),      # [streams, frames, ics, 1024]
        truth_sfidx=np.stack([s.truth["sfidx"] for s in streams]),
        truth_sfbcb=np.stack([s.truth["sfbcb"] for s in streams]),
        truth_info=np.stack([s.truth["info"] for s in streams]),
        **extra,
    )
    return out


def main():
    only = set(sys.argv[1:])
    for name, (cfg, n, asc) in cases().items():
        if only and name not in only:
            continue
        data = build_case(name, cfg, n, asc)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print("%-20s %7d bytes blob, %3d frames -> %s (%d bytes)" % (name, data["blob"].nbytes, len(data["frame_offset"]),
                                                                     os.path.basename(path), os.path.getsize(path)))


if __name__ == "__main__":
    main()
