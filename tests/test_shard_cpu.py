"""Host-side multi-GPU logic on CPU: stream sharding and the timing reduction, world_size 2 over gloo."""
import hashlib
import os
import sys

import numpy as np
import pytest

from jaadec_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 4096, 16384, 16385):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = shard.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                cover.extend(range(lo, hi))
                for u in (lo, hi - 1):
                    if lo < hi:
                        assert shard.owner_of(u, n, world) == r
            assert cover == list(range(n))
            sizes = [shard.shard_range(n, r, world) for r in range(world)]
            assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import gen
    import oracle
    from jaadec_b200 import shard as sh
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_streams, cfg = 5, gen.config(2, n_frames=4)
        lo, hi = sh.shard_range(n_streams, rank, world)
        digests = {}
        for s in range(lo, hi):   # every rank decodes only the streams it owns (oracle stands in for the GPU here)
            st = gen.generate(cfg, gen.seed_for(2, s))
            dec = oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg)
            h = hashlib.sha256()
            for f in range(cfg.n_frames):
                h.update(dec.decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])["s16"].tobytes())
            digests[s] = h.hexdigest()
        gathered = [None] * world
        dist.all_gather_object(gathered, digests)
        units, secs = sh.aggregate(float(hi - lo), 1.0 + rank)
        q.put((rank, gathered, units, secs))
    finally:
        dist.destroy_process_group()


def test_two_ranks_cover_all_streams_and_reduce_timing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import gen
    import oracle
    cfg = gen.config(2, n_frames=4)
    for rank, gathered, units, secs in out:
        merged = {}
        for d in gathered:
            assert not (set(d) & set(merged)), "a stream was decoded by two ranks"
            merged.update(d)
        assert sorted(merged) == [0, 1, 2, 3, 4]
        assert units == 5.0 and secs == 2.0   # sum of units, max of seconds
    # and the sharded result equals the single-process one
    st = gen.generate(cfg, gen.seed_for(2, 3))
    dec = oracle.Decoder.create_adts(2, cfg.sf_index, cfg.chan_cfg)
    h = hashlib.sha256()
    for f in range(cfg.n_frames):
        h.update(dec.decode_frame(st.data[st.offsets[f]: st.offsets[f] + st.sizes[f]])["s16"].tobytes())
    assert out[0][1][0].get(3, out[0][1][1].get(3)) == h.hexdigest()
