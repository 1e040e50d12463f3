"""CPU tests of the N>1 path: two gloo ranks shard the work exactly as the GPU ranks do (stripes of the
all-vs-all result, ranges of findClosest queries), compute their shard with the CPU oracle, exchange results
with the collectives of kmers/reps_b200/sharding.py, and must reproduce the single-rank answer."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import kmers.reps_b200 as kr
        from kmers.reps_b200 import sharding, synth
        from oracle import kr_oracle as ko
        o = ko.get()
        res, off, _ = synth.generate(99, n, 12)
        b = o.batch(res=res, off=off, K=8)
        full = o.sim_matrix(b, ident_inf=True)
        for layout in (kr.LAYOUT_BLOCKTRI, kr.LAYOUT_FULL, kr.LAYOUT_TRIANGLE):
            if layout == kr.LAYOUT_BLOCKTRI:
                whole = sharding.blocktri_from_full(full)
            elif layout == kr.LAYOUT_FULL:
                whole = full.ravel()
            else:
                whole = full[np.triu_indices(n, 1)]
            r0, r1, elems, first = sharding.my_stripe(n, layout, rank, world)
            assert r0 % 256 == 0 and (r1 % 256 == 0 or r1 == n)
            local = torch.from_numpy(np.ascontiguousarray(whole[first:first + elems]).astype(np.int32))
            got = sharding.allgather_stripes(local, n, layout, dist)
            assert got.shape[0] == whole.shape[0] and (got.numpy() == whole).all()
        # findClosest: replicated database, queries split by rank, results gathered back in query order
        nrep = n // 2
        reps = o.batch(res=res[: int(off[nrep])], off=off[: nrep + 1], K=8)
        nq = n - nrep
        lo, hi = sharding.shard_range(nq, rank, world)
        qoff = off[nrep + lo: nrep + hi + 1] - off[nrep + lo]
        qs = o.batch(res=res[int(off[nrep + lo]): int(off[nrep + hi])], off=qoff, K=8)
        bs, bi = o.find_closest(reps, qs)
        local = torch.from_numpy(np.stack([bs, bi], 1).astype(np.int64))
        allq = sharding.allgather_ranges(local, nq, dist).numpy()
        if rank == 0:
            wq = o.batch(res=res[int(off[nrep]):], off=off[nrep:] - off[nrep], K=8)
            ws, wi = o.find_closest(reps, wq)
            assert (allq[:, 0] == ws).all() and (allq[:, 1] == wi).all()
        # greedy build: replicas of the database on every rank, scoring of each candidate batch split by rank
        class Replica:                                   # the oracle stands in for this rank's GPU replica
            def __init__(self):
                self.idx = []
        rep = Replica()
        seqs = synth.to_strings(res, off)

        def score_slice(lo, hi):
            if hi <= lo:
                return np.zeros(0, np.int32)
            return o.find_closest(o.batch([seqs[i] for i in rep.idx], K=8), o.batch(seqs[lo:hi], K=8))[0]

        def survivor_matrix(idx):
            return o.sim_matrix(o.batch([seqs[i] for i in idx], K=8), ident_inf=True)

        for batch in (64, 257):
            rep.idx = []
            added = sharding.greedy_build_sharded(n, 50, batch, dist, score_slice, survivor_matrix, lambda idx: rep.idx.extend(int(i) for i in idx))
            _, want, _ = o.greedy_build(b, 50, ident_inf=True)
            assert (added == want).all() and rep.idx == list(np.nonzero(want)[0])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, "FAIL %r" % (e,)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_paths_reassemble(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, 700, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    results = sorted(q.get(timeout=5) for _ in range(world))
    assert results == [(r, "ok") for r in range(world)], results


def test_shard_range_partition():
    from kmers.reps_b200 import sharding
    for n in (0, 1, 7, 10, 1000):
        for w in (1, 2, 3, 8):
            cuts = [sharding.shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_block_shards_cut_at_tiles():
    from kmers.reps_b200 import sharding
    for n in (1, 255, 256, 257, 100000, 1480):
        for w in (1, 2, 3, 4, 8):
            cuts = sharding.block_shards(n, w)
            assert cuts[0] == 0 and cuts[-1] == n and len(cuts) == w + 1
            assert all(a <= b for a, b in zip(cuts, cuts[1:]))
            assert all(c % 256 == 0 for c in cuts[:-1])          # inner shards hold whole 256-sequence blocks
    cuts = sharding.block_shards(100000, 8)
    sizes = [b - a for a, b in zip(cuts, cuts[1:])]
    assert max(sizes) - min(sizes) <= 256 + 160
