"""Does the parse (K1, 73 % of the issue slots alone) overlap with the filterbank (K2, 64 %) when both are in flight?

Two engines on one GPU, half of config 2 each (2048 streams x 469 frames), resident batches.  (a) one after the other on
one host thread with a sync in between; (b) both launched back to back (each engine has its own CUDA stream, so K1 of the
second batch is eligible while K2 of the first runs); (c) the single 4096-stream batch bench.py times.  Wall clock around
launch + sync, best of N (25 ms steps, the launch overhead is microseconds)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from jaadec_b200 import Engine, PCM_S16LE  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = 8
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wls = [bench.Workload(2, S // parts, 469, i * (S // parts)) for i in range(parts)]
engs, batches = [], []
for wl in wls:
    eng = Engine(device=0, max_streams=wl.n_streams, pcm_format=PCM_S16LE)
    ids = np.asarray(wl.open_streams(eng), np.int32)
    frames = wl.index(ids)
    b = eng.batch(frames, wl.blob.nbytes)
    b.upload(wl.blob)
    b.sync()
    engs.append(eng)
    batches.append(b)


def timed(fn):
    best = 1e9
    for _ in range(N):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def sequential():
    for b in batches:
        b.decode()
        b.sync()


def concurrent():
    for b in batches:
        b.decode()
    for b in batches:
        b.sync()


for _ in range(3):
    sequential()
print("parts %d x %d streams: sequential %.2f ms, back-to-back on %d CUDA streams %.2f ms" % (parts, S // parts, timed(sequential), parts, timed(concurrent)))
