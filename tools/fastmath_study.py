#!/usr/bin/env python3
"""What bit-exactness costs: the engine built with FMA contraction allowed (nvcc --fmad=true, the compiler's default) against
the product build (--fmad=false = Java's float semantics), on BASELINE config 2 (and config 3 for the SBR stages).

    JAADB200_LIB=<variant.so> python tools/fastmath_study.py [--config 2] [--streams 64]

Decodes `streams` generator streams with the loaded library in F32 mode and compares every frame with the oracle:
max abs error in JAAD's +-32768 float domain (BASELINE.json's budget is 1e-5 of full scale = 0.328), the number of int16
samples that round differently, and the kernel time of a 4096-stream batch.  One JSON line on stdout.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gen      # noqa: E402
import oracle   # noqa: E402
from jaadec_b200 import Engine, FLAG_PROFILE, FRAME_DESC_DTYPE, PCM_F32_PLANAR, PCM_S16LE  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--bench-streams", type=int, default=4096)
    a = ap.parse_args()
    n_frames = {2: 469, 3: 235, 4: 235}[a.config]
    cfg = gen.config(a.config, n_frames=n_frames)
    blob, offs, sizes, _ = gen.generate_many(cfg, gen.seed_for(a.config, 0), a.streams)
    S, F = offs.shape
    eng = Engine(max_streams=S, pcm_format=PCM_F32_PLANAR)
    ids = [eng.open_adts(2, cfg.sf_index, cfg.chan_cfg, expect_sbr=cfg.sbr_mode) for _ in range(S)]
    fr = np.zeros(S * F, FRAME_DESC_DTYPE)
    fr["offset"], fr["nbytes"], fr["stream_id"] = offs.T.reshape(-1), sizes.T.reshape(-1), np.tile(np.asarray(ids, np.int32), F)
    pcm, res = eng.decode(blob, fr)
    info = eng.stream_info(ids[0])
    ch, ln = info.channels, info.sample_length
    got = pcm.view(np.float32).reshape(F, S, ch, ln)
    eng.close()
    max_err, n_flip, n_samples, n_bits = 0.0, 0, 0, 0
    for s in range(S):
        dec = oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg)
        for f in range(F):
            r = dec.decode_frame(blob[offs[s, f]: offs[s, f] + sizes[s, f]])
            ref = r["f32"]
            g = got[f, s]
            max_err = max(max_err, float(np.abs(g.astype(np.float64) - ref).max()))
            # Java Math.round + clamp on both
            q = lambda x: np.clip(np.floor(x.astype(np.float64) + 0.5), -32768, 32767)
            n_flip += int((q(g) != q(ref)).sum())
            n_bits += int((g.view(np.uint32) != ref.view(np.uint32)).sum())
            n_samples += g.size
    # kernel time of the benchmark batch
    cfgb = gen.config(a.config, n_frames=n_frames)
    blob, offs, sizes, _ = gen.generate_many(cfgb, gen.seed_for(a.config, 0), a.bench_streams)
    S, F = offs.shape
    eng = Engine(max_streams=S, pcm_format=PCM_S16LE, flags=FLAG_PROFILE)
    ids = [eng.open_adts(2, cfg.sf_index, cfg.chan_cfg, expect_sbr=cfg.sbr_mode) for _ in range(S)]
    fr = np.zeros(S * F, FRAME_DESC_DTYPE)
    fr["offset"], fr["nbytes"], fr["stream_id"] = offs.T.reshape(-1), sizes.T.reshape(-1), np.tile(np.asarray(ids, np.int32), F)
    b = eng.batch(fr, blob.nbytes)
    b.upload(blob)
    ms = []
    for i in range(5):
        b.decode()
        t = b.timings()
        if i >= 2:
            ms.append((t.parse_ms, t.filterbank_ms, t.sbr_ms, t.total_ms))
    ms = np.mean(np.array(ms), axis=0)
    print(json.dumps({"lib": os.environ.get("JAADB200_LIB", "product build (--fmad=false)"), "config": a.config, "frames_checked": a.streams * n_frames,
                      "max_abs_err": max_err, "budget_1e-5_of_full_scale": 0.32768, "int16_mismatches": n_flip, "float_bit_mismatches": n_bits,
                      "samples": n_samples, "k1_k3_ms": float(ms[0]), "k2_ms": float(ms[1]), "k4_k5_ms": float(ms[2]), "step_ms": float(ms[3])}))


if __name__ == "__main__":
    main()
