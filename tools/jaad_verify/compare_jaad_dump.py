#!/usr/bin/env python3
"""Compares JaadDump's output (real JAAD on a JVM) with the committed golden fixtures: per case, the int16 PCM of every frame
(Math.round + clamp, as SampleBuffer.accept does) must equal tests/golden/<case>.npz["s16"], and the SHA-256 over the float
bits must equal ["f32_sha256"].  A case that passes turns "parity unpinned" into "pinned against JAAD" for everything the
fixture covers; exits non-zero if any case fails."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def read_dump(path):
    raw = np.fromfile(path, np.uint8)
    pos, frames = 0, []
    while pos + 16 <= len(raw):
        status, ch, ln, rate = np.frombuffer(raw[pos:pos + 16].tobytes(), "<i4")
        pos += 16
        if status != 0:
            frames.append(None)
            continue
        n = int(ch) * int(ln)
        frames.append(np.frombuffer(raw[pos:pos + 4 * n].tobytes(), "<f4").reshape(int(ch), int(ln)))
        pos += 4 * n
    return frames


def java_round_s16(x):
    return np.clip(np.floor(x.astype(np.float64) + 0.5), -32768, 32767).astype(np.int16)


def main():
    work = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_work")
    manifest = json.load(open(os.path.join(work, "manifest.json")))
    cases, bad = {}, 0
    for m in manifest:
        cases.setdefault(m["case"], []).append(m)
    for case, ms in sorted(cases.items()):
        g = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
        dumps = {m["stream"]: read_dump(os.path.join(work, os.path.splitext(m["file"])[0] + ".dump")) for m in ms}
        sha = hashlib.sha256()
        seen = {s: 0 for s in dumps}
        ok_s16, n_frames, worst = True, 0, 0.0
        for i, s in enumerate(g["frame_stream"]):       # the fixture's frame order: frame-major over the streams
            s = int(s)
            fr = dumps[s][seen[s]] if seen[s] < len(dumps[s]) else None
            seen[s] += 1
            if fr is None:
                ok_s16 = False
                continue
            sha.update(np.ascontiguousarray(fr, "<f4").tobytes())
            want = g["s16"][i]                           # [samples, channels]
            got = java_round_s16(fr).T
            if got.shape != want.shape or not np.array_equal(got, want):
                ok_s16 = False
                if got.shape == want.shape:
                    worst = max(worst, float(np.abs(got.astype(np.int32) - want).max()))
            n_frames += 1
        ok_f32 = sha.digest() == g["f32_sha256"].tobytes()
        print("%-18s %3d frames  int16 PCM %s  float bits %s%s" % (case, n_frames, "IDENTICAL" if ok_s16 else "DIFFERENT",
                                                                   "IDENTICAL" if ok_f32 else "different",
                                                                   "" if ok_s16 else "  (max int16 difference %g)" % worst))
        bad += (not ok_s16) or (not ok_f32)
    print("parity against JAAD: %s" % ("PINNED for every fixture" if bad == 0 else "%d fixture(s) differ" % bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
