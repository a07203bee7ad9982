"""Which frames of the (valid) config-4 bench workload does the engine fail, and with what status?"""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from jaadec_b200 import Engine, PCM_S16LE, FLAG_DEBUG_TAPS
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wl = bench.Workload(4, S, 235, 0)
eng = Engine(device=0, max_streams=S, pcm_format=PCM_S16LE, flags=FLAG_DEBUG_TAPS)
ids = np.asarray(wl.open_streams(eng), np.int32)
frames = wl.index(ids)
b = eng.batch(frames, wl.blob.nbytes)
b.upload(wl.blob); b.decode()
_, res = b.download(want_results=True)
bad = np.nonzero(res["status"] != 0)[0]
print("bad", len(bad), "statuses", np.unique(res["status"][bad], return_counts=True))
for i in bad[:12]:
    sid = int(frames["stream_id"][i]); s = int(np.nonzero(ids == sid)[0][0])
    f = int((frames["stream_id"][:i] == sid).sum())
    ps = b.tap_ps(int(i)); sb = b.tap_sbr(int(i), 0)
    print("stream", s, "frame", f, "status", int(res["status"][i]),
          "| sbr: N_high", None if sb is None else int(sb["N_high"]), "L_E", None if sb is None else int(sb["L_E"]), "f", None if sb is None else sb["f"].tolist(), "kx", None if sb is None else int(sb["kx"]), "M", None if sb is None else int(sb["M"]),
          "noPatches", None if sb is None else int(sb["noPatches"]), "pNoSb", None if sb is None else sb["patchNoSubbands"].tolist(), "pStart", None if sb is None else sb["patchStartSubband"].tolist(),
          "| ps:", None if ps is None else dict(num_env=int(ps["num_env"]), iid_mode=int(ps["iid_mode"]), icc_mode=int(ps["icc_mode"]), iid=ps["iid"][:int(ps["num_env"])].tolist(), icc=ps["icc"][:int(ps["num_env"])].tolist()))
