#!/usr/bin/env python3
"""Writes every stream of the committed golden fixtures (tests/golden/*.npz) as a file JAAD's own front-ends read: ADTS
streams as <case>_s<k>.aac, raw-frame (AudioSpecificConfig) streams as <case>_s<k>.mp4 (gen/mp4.py).  A manifest.json lists
them.  Part of the JVM-host verification kit (tools/jaad_verify/run.sh)."""
import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gen import mp4 as genmp4  # noqa: E402

SF = [96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000]


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_work")
    os.makedirs(out_dir, exist_ok=True)
    manifest = []
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        name = os.path.splitext(os.path.basename(path))[0]
        g = np.load(path)
        blob, off, nb, sid = g["blob"], g["frame_offset"], g["frame_nbytes"], g["frame_stream"]
        asc = g["asc"].tobytes()
        hdr = [int(x) for x in g["hdr"]]
        for s in range(int(sid.max()) + 1):
            rows = np.nonzero(sid == s)[0]
            if len(asc) == 0:
                # the generator wrote 7-byte ADTS headers in front of every payload: the stream is one contiguous byte range
                lo, hi = int(off[rows[0]]) - 7, int(off[rows[-1]] + nb[rows[-1]])
                data, fn = blob[lo:hi], "%s_s%d.aac" % (name, s)
            else:
                raw = np.concatenate([blob[int(off[r]): int(off[r]) + int(nb[r])] for r in rows])
                channels = {1: 1, 2: 2, 6: 6}.get(hdr[2], 2)
                data = genmp4.write_mp4((raw, nb[rows]), asc, SF[hdr[1]], channels)[0]
                fn = "%s_s%d.mp4" % (name, s)
            data.tofile(os.path.join(out_dir, fn))
            manifest.append({"case": name, "stream": s, "file": fn, "frames": int(len(rows))})
    with open(os.path.join(out_dir, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("%d files in %s" % (len(manifest), out_dir))


if __name__ == "__main__":
    main()
