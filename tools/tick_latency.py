"""Latency of one tick of a live batch through the one-call API: S streams x ONE frame per jaadb_decode call, host buffers
(pinned), frames of consecutive ticks fed in order -- the shape a drop-in behind N x Decoder.decodeFrame sees.

    python tools/tick_latency.py [streams=4096] [ticks=64]          (JAADB_K1_LANES_LOG2=5 forces the dense parse mapping)
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from jaadec_b200 import Engine, FRAME_RESULT_DTYPE, PCM_S16LE  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
wl = bench.Workload(2, S, T, 0)
eng = Engine(device=0, max_streams=S, pcm_format=PCM_S16LE)
ids = np.asarray(wl.open_streams(eng), np.int32)
frames = wl.index(ids)                       # frame-major: tick t = frames[t * S:(t + 1) * S]
# what arrives in one tick: the S frames of that tick, packed (a caller hands over the bytes it received, not the whole files)
ticks = []
for t in range(T):
    tick = frames[t * S:(t + 1) * S].copy()
    sizes = tick["nbytes"].astype(np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    buf = torch.empty(int(sizes.sum()), dtype=torch.uint8, pin_memory=True)
    b = buf.numpy()
    for i in range(S):
        o = int(tick["offset"][i])
        b[starts[i]:starts[i] + sizes[i]] = wl.blob[o:o + sizes[i]]
    tick["offset"] = starts
    ticks.append((buf, tick))
pcm = torch.empty(S * 1024 * 2 * 2, dtype=torch.uint8, pin_memory=True)
res = np.zeros(S, FRAME_RESULT_DTYPE)
ms = []
for buf, tick in ticks:
    t0 = time.perf_counter()
    eng.decode_ptr(buf.data_ptr(), buf.numel(), tick, pcm.data_ptr(), pcm.numel(), results=res)
    ms.append((time.perf_counter() - t0) * 1e3)
    assert (res["status"] == 0).all()
ms = np.array(ms[8:])
print("tick of %d streams x 1 frame (%.1f ms of audio each): median %.3f ms, p90 %.3f ms, min %.3f ms per jaadb_decode call -> %.0f x realtime"
      % (S, 1024 / 48.0, np.median(ms), np.percentile(ms, 90), ms.min(), S * 1024 / 48000.0 / (np.median(ms) * 1e-3)))
