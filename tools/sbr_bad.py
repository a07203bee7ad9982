#!/usr/bin/env python3
"""Debug aid: list frames of the config-3 bench workload whose status differs from 0 on the GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gen
from jaadec_b200 import Engine, PCM_S16LE, FRAME_DESC_DTYPE
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = gen.config(3)
blob, offs, sizes, sb = gen.generate_many(cfg, gen.seed_for(3, 0), S)
F = offs.shape[1]
eng = Engine(max_streams=S, pcm_format=PCM_S16LE)
ids = [eng.open_adts(2, 6, 2, expect_sbr=1) for _ in range(S)]
fr = np.zeros(S * F, FRAME_DESC_DTYPE)
fr["offset"] = offs.T.reshape(-1); fr["nbytes"] = sizes.T.reshape(-1); fr["stream_id"] = np.tile(np.asarray(ids, np.int32), F)
b = eng.batch(fr, blob.nbytes); b.upload(blob); b.decode(); b.sync(); b.decode()
pcm, res = b.download()
bad = np.nonzero(res["status"])[0]
print("bad", len(bad), "of", S * F)
seen = {}
for i in bad:
    s, f = int(i % S), int(i // S)
    seen.setdefault(s, []).append((f, int(res["status"][i])))
for s, lst in list(seen.items())[:12]:
    print("stream", s, "seed", gen.seed_for(3, s), lst[:6], "n", len(lst))
    f0 = lst[0][0]
    g = b.tap_sbr(f0 * S + s, 0)
    print("   rec", {k: (g[k].tolist() if hasattr(g[k], "tolist") else g[k]) for k in ("mode","reset","L_E","kx","M","N_high","N_low","N_Q","N_L","noPatches","t_E","f")})
