"""The JVM-host verification kit (tools/jaad_verify) must not rot: export every golden stream as an .aac / .mp4 file, decode
the files the way JaadDump does -- here with the oracle standing in for JAAD, fed through the product's host-side container
indexers -- write dumps in JaadDump's format and run the kit's comparer, which must report every fixture identical.

On a host with a JVM, `tools/jaad_verify/run.sh <JAAD classpath>` does the same with the real reference and is the command
that turns "parity unpinned" into "pinned" (README.md)."""
import json
import os
import subprocess
import sys

import numpy as np

import oracle
from jaadec_b200 import demux

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KIT = os.path.join(ROOT, "tools", "jaad_verify")


def test_kit_round_trip(tmp_path):
    work = str(tmp_path)
    subprocess.check_call([sys.executable, os.path.join(KIT, "export_streams.py"), work], stdout=subprocess.DEVNULL)
    manifest = json.load(open(os.path.join(work, "manifest.json")))
    assert len(manifest) >= 19 and {m["file"].rsplit(".", 1)[1] for m in manifest} == {"aac", "mp4"}
    for m in manifest:
        data = np.fromfile(os.path.join(work, m["file"]), np.uint8)
        if m["file"].endswith(".aac"):
            frames, info = demux.adts_index(data)                      # ADTSDemultiplexer
            dec = oracle.Decoder.create_adts(info.profile, info.sf_index, info.channel_config)
        else:
            frames, track = demux.mp4_index(data)                      # MP4Container / Track
            dec = oracle.Decoder.create_asc(demux.asc_of(track))
        assert len(frames) == m["frames"]
        with open(os.path.join(work, os.path.splitext(m["file"])[0] + ".dump"), "wb") as out:
            for r in frames:
                res = dec.decode_frame(data[int(r["offset"]): int(r["offset"]) + int(r["nbytes"])])
                assert res["status"] == 0
                out.write(np.array([0, res["channels"], res["sample_length"], res["sample_rate"]], "<i4").tobytes())
                out.write(res["f32"].astype("<f4").tobytes())
    rc = subprocess.run([sys.executable, os.path.join(KIT, "compare_jaad_dump.py"), work], capture_output=True, text=True)
    assert rc.returncode == 0 and "PINNED for every fixture" in rc.stdout, rc.stdout + rc.stderr
    # and the comparer does notice a difference: flip one float of one dump
    victim = os.path.join(work, os.path.splitext(manifest[0]["file"])[0] + ".dump")
    raw = np.fromfile(victim, np.uint8)
    raw[16 + 400] ^= 0x40
    raw.tofile(victim)
    rc = subprocess.run([sys.executable, os.path.join(KIT, "compare_jaad_dump.py"), work], capture_output=True, text=True)
    assert rc.returncode == 1 and "differ" in rc.stdout
