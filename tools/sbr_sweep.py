#!/usr/bin/env python3
"""Debug aid: sweep SBR seeds on the GPU and report which header settings break bit-exactness."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gen, oracle
from helpers import Workload
from jaadec_b200 import Engine, PCM_F32_PLANAR

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mono = len(sys.argv) > 2 and sys.argv[2] in ("mono", "ps")
ps = len(sys.argv) > 2 and sys.argv[2] == "ps"
cfg = gen.GenConfig(sf_index=6, chan_cfg=1, n_frames=24, target_bytes=171, sbr_mode=2 if ps else 1) if mono else gen.config(3, n_frames=24, sbr_quirk=True)
wl = Workload(cfg, n, base_seed=int(sys.argv[3]) if len(sys.argv) > 3 else 70000, with_truth=False)
decs = wl.oracle_decoders()
eng = Engine(max_streams=n, pcm_format=PCM_F32_PLANAR, sbr_tile_frames=int(os.environ.get('TILE', '0')))
ids = [eng.open_adts(*wl.hdr, expect_sbr=cfg.sbr_mode) for _ in range(n)]
frames, index = wl.frame_table(ids)
pcm, res = eng.decode(wl.blob, frames)
per = 2 * 2048 * 4
failed = {}
for i, (s, f) in enumerate(index):
    r = decs[s].decode_frame(wl.frame_bytes(s, f))
    if s in failed:
        continue
    t = decs[s].tap_sbr(0, 0)
    if res["status"][i] != r["status"]:
        failed[s] = (f, "status", int(res["status"][i]), r["status"])
        continue
    got = pcm[i * per:(i + 1) * per].view(np.float32).reshape(2, 2048)
    if not np.array_equal(got.view(np.uint32), r["f32"].view(np.uint32)):
        bad = np.argwhere(got.view(np.uint32) != r["f32"].view(np.uint32))
        failed[s] = (f, "pcm", float(np.abs(got - r["f32"]).max()), "first bad", bad[0].tolist(), "nbad", len(bad),
                     "ints", t["ints"][:18].tolist(), "extra", t["extra"].tolist())
for s in range(n):
    t = decs[s].tap_sbr(0, 0)
    print(s, "FAIL" if s in failed else "ok", failed.get(s, ""), "" if s in failed else t["extra"].tolist())
print("failed", len(failed), "of", n)
