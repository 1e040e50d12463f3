"""TEST INFRASTRUCTURE: runs the product's device + host code on the CPU emulator (tests/emul/cuda_emul.h) and
compares with the oracle.  Started by tests/test_emul_device_code.py in a subprocess with KR_LIB_PATH pointing at
tests/emul/_build/libkmersreps_emul.so (the product sources compiled by g++ with -DKR_EMUL).

    python tests/emul/run_emul_checks.py <case>
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import kmers.reps_b200 as kr                      # noqa: E402
from kmers.reps_b200 import binding, synth        # noqa: E402
from oracle import kr_oracle as ko                # noqa: E402

assert "libkmersreps_emul" in kr.lib_path(), "these checks are for the emulator build only"
O = ko.get()


def family(seed, n, nfam, lo=None, hi=None):
    res, off, _ = synth.generate(seed, n, nfam)
    seqs = synth.to_strings(res, off)
    if lo is not None:                                # trim to keep the emulated kernels quick
        rng = np.random.default_rng(seed)
        seqs = [s[: int(rng.integers(lo, hi))] for s in seqs]
    return seqs


def blocktri(full):
    n = full.shape[0]
    return np.concatenate([full[i, 256 * (i // 256):] for i in range(n)])


def cells(arr):
    return {(int(r), int(c), int(v)) for r, c, v, _ in arr}


def want_cells(dense, min_sim, triangle, row0=0):
    ii, jj = np.nonzero(dense >= min_sim)
    return {(int(i) + row0, int(j), int(dense[i, j])) for i, j in zip(ii, jj) if not triangle or j > i + row0}


PATHS = (kr.FORCE_JOIN, kr.FORCE_INDEX, 0)      # tile join pinned, index scan pinned, the library's own choice


def case_matrix_layouts():
    """encode + (block lists + TMA-staged tile join | sorted index + warp scan): every layout / element type, two blocks,
    the equal-sequence rule incl. k-mer-less equal sequences"""
    seqs = family(11, 300, 6, 60, 140) + ["MKV", "", "MKVLAAGIVG", "MKVLAAGIVG", "MKV", ""]
    seqs[7] = seqs[290]                               # equal sequences in different blocks
    n = len(seqs)
    for ident in (kr.IDENT_IS_INFINITY, 0):
        want = O.sim_matrix(O.batch(seqs, K=8), ident_inf=bool(ident), threads=4)
        for path in PATHS:
            c = kr.Context(device=0, K=8, flags=ident | path)
            s = c.encode(seqs)
            assert (s.sizes() == O.batch(seqs, K=8).sizes()).all()
            assert (c.sim_matrix(s, elem=kr.ELEM_I32, layout=kr.LAYOUT_FULL) == want).all(), (ident, path)
            assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_TRIANGLE) == want[np.triu_indices(n, 1)]).all(), (ident, path)
            assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI) == blocktri(want)).all(), (ident, path)
            hi = c.sim_matrix(s, elem=kr.ELEM_I32, layout=kr.LAYOUT_BLOCKTRI, row_begin=256, row_end=n)
            assert (hi == blocktri(want)[-len(hi):]).all(), (ident, path)
            lo = c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_FULL, row_begin=0, row_end=256)
            assert (lo == want[:256]).all(), (ident, path)
            c.close()


def case_k10_rectangular_pairs():
    """32-bit low words (K = 10, three pipeline stages) in the join; 50-bit keys in the index; rectangular matrix,
    explicit pairs, key decoding"""
    seqs = family(12, 90, 3, 50, 120)
    oa, ob = O.batch(seqs[:40], K=10), O.batch(seqs[30:], K=10)
    full = O.sim_matrix(oa, ob, ident_inf=False)
    for path in PATHS:
        c = kr.Context(device=0, K=10, flags=path)
        a, b = c.encode(seqs[:40]), c.encode(seqs[30:])
        assert (c.sim_matrix(a, b) == full).all(), path
        ia, ib = np.arange(40, dtype=np.uint32), np.arange(40, dtype=np.uint32)[::-1].copy()
        sim, ident = c.sim_pairs(a, ia, b, ib)
        assert (sim == full[ia, ib]).all()
        assert sorted(a.kmers(3)) == sorted(ko.py_kmers(seqs[3], 10))
        c.close()
    # K = 11: 55-bit keys leave 9 id bits -- a set of 90 fits the index, a database (ids sized for growth) does not
    c = kr.Context(device=0, K=11, flags=kr.IDENT_IS_INFINITY)
    s = c.encode(seqs)
    w11 = O.sim_matrix(O.batch(seqs, K=11), ident_inf=True)
    assert (c.sim_matrix(s) == w11).all()
    db = c.db(); db.append(seqs[:60])
    sim, first, _ = db.find_closest(c.encode(seqs[60:]))
    bs, bi = O.find_closest(O.batch(seqs[:60], K=11), O.batch(seqs[60:], K=11))
    assert (sim == bs).all() and (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all()
    c.close()


def case_find_closest_and_build():
    """ARGMAX join and index scan (pinned and chosen), ranks / ties, greedy build with database compaction and the
    two-level index (main + tail) growing batch by batch"""
    seqs = family(13, 330, 8, 60, 130)
    reps, qs = seqs[:270], seqs[270:] + [seqs[5], "MKV", ""]
    reps[20] = "MKV"
    ob, oq = O.batch(reps, K=8), O.batch(qs, K=8)
    rng = np.random.default_rng(3)
    rank = rng.permutation(len(reps)).astype(np.uint32)
    order = np.argsort(rank)
    bs, bi = O.find_closest(ob, oq, order=order.astype(np.int32), ident_inf=True)
    bs2, bl2 = O.find_closest(ob, oq, order=order.astype(np.int32), ident_inf=True, tie_last=True)
    for path in PATHS:
        c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | path)
        db = c.db()
        db.append(reps[:100]); db.append(reps[100:])
        db.set_ranks(rank)
        sim, first, last = db.find_closest(c.encode(qs))
        assert (sim == bs).all(), path
        assert (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all(), path
        assert (last.astype(np.int64) == np.where(bs2 > 0, bl2.astype(np.int64), kr.NO_REP)).all(), path
        assert (db.any_ge(c.encode(qs), 25) == (bs >= 25)).all()
        c.close()
    cand = list(seqs)
    cand[50:50] = ["MKV", "MKV", "", "MKVLAAGIVG", "MKVLAAGIVG"]
    for thr in (30, 10):
        _, want, _ = O.greedy_build(O.batch(cand, K=8), thr, ident_inf=True)
        for path in (kr.FORCE_JOIN, 0):
            for batch in (40, 4096):                                        # 40: many small segments -> db_compact, tail rebuilds
                c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | path)
                db = c.db()
                a0 = db.greedy_build(c.encode(cand[:120]), thr, batch=batch)     # fills the database ...
                a1 = db.greedy_build(c.encode(cand[120:]), thr, batch=batch)     # ... which the second call extends
                assert (np.concatenate([a0, a1]) == want).all(), (thr, path, batch)
                assert db.size == int(want.sum())
                # the grown database answers like a fresh one
                sim, first, _ = db.find_closest(c.encode(qs))
                acc = [cand[i] for i in np.nonzero(want)[0]]
                bs3, bi3 = O.find_closest(O.batch(acc, K=8), oq, ident_inf=True)
                assert (sim == bs3).all() and (first.astype(np.int64) == np.where(bs3 > 0, bi3.astype(np.int64), kr.NO_REP)).all(), (thr, path, batch)
                c.close()
    _, w0, _ = O.greedy_build(O.batch(cand[:60], K=8), 0, ident_inf=True)   # threshold 0: the first candidate represents everything
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
    assert (c.db().greedy_build(c.encode(cand[:60]), 0, batch=16) == w0).all() and w0.sum() == 1
    c.close()


def case_index_crowded():
    """the exactness paths of the index: a k-mer shared by 5,000 sequences (a build prefix larger than the shared-memory
    sort, a posting list longer than a score table), queries that touch more ids than a table holds (4 / 16 passes),
    for ARGMAX, sparse cells and dense cells"""
    rng = np.random.default_rng(21)
    alpha = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    rnd = lambda n: alpha[rng.integers(0, 20, n)].tobytes().decode()
    hub = "MKVLAAGI"
    reps = [rnd(24) for _ in range(900)] + [rnd(8) + hub + rnd(8) for _ in range(5000)]
    fam = [reps[i][:12] + rnd(6) + reps[i][18:] for i in (5, 17, 3300, 4000)]          # near copies: clear best matches
    qs = fam + [rnd(10) + hub + rnd(10), hub, reps[77], rnd(30), reps[3200][8:16] + rnd(5)]
    order = rng.permutation(len(reps)).astype(np.uint32)
    ob, oq = O.batch(reps, K=8), O.batch(qs, K=8)
    bs, bi = O.find_closest(ob, oq, order=np.argsort(order).astype(np.int32), ident_inf=True, threads=4)
    bs2, bl2 = O.find_closest(ob, oq, order=np.argsort(order).astype(np.int32), ident_inf=True, tie_last=True, threads=4)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | kr.FORCE_INDEX)
    db = c.db(); db.append(reps); db.set_ranks(order)
    Q = c.encode(qs)
    sim, first, last = db.find_closest(Q)
    assert (sim == bs).all()
    assert (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all()
    assert (last.astype(np.int64) == np.where(bs2 > 0, bl2.astype(np.int64), kr.NO_REP)).all()
    assert bs[5] == 1 and bi[5] != bl2[5]                                           # the hub query ties over 5,000 representatives
    # the same crowd as a matrix: 9 query rows x 5,900 columns, dense and sparse
    R = c.encode(reps)
    want = O.sim_matrix(oq, ob, ident_inf=True, threads=4)
    assert (c.sim_matrix(Q, R, elem=kr.ELEM_U16) == want).all()
    got = c.sim_matrix_sparse(Q, R, min_sim=1, layout=kr.LAYOUT_FULL, cap=1 << 16)
    assert len(cells(got)) == len(got) and cells(got) == want_cells(want, 1, False)
    c.close()


def case_sparse_matrix():
    """kr_sim_matrix_sparse through both paths: thresholds, too little room (retried), row stripes, rectangular"""
    seqs = family(15, 300, 6, 60, 120) + ["MKV", "", "MKV"]
    seqs[5] = seqs[280] = "MKV"
    want = O.sim_matrix(O.batch(seqs, K=8), ident_inf=True, threads=4)
    for path in PATHS:
        c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | path)
        s = c.encode(seqs)
        for min_sim in (1, 20, 9999):
            got = c.sim_matrix_sparse(s, min_sim=min_sim, layout=kr.LAYOUT_TRIANGLE)
            assert len(cells(got)) == len(got) and cells(got) == want_cells(want, min_sim, True), (path, min_sim)
        got = c.sim_matrix_sparse(s, min_sim=1, layout=kr.LAYOUT_TRIANGLE, cap=16)       # too little room: retried
        assert cells(got) == want_cells(want, 1, True), path
        got = c.sim_matrix_sparse(s, min_sim=1, layout=kr.LAYOUT_TRIANGLE, row_begin=256, row_end=len(seqs))
        assert cells(got) == want_cells(want[256:], 1, True, row0=256), path
        q = c.encode(seqs[:290])
        got = c.sim_matrix_sparse(q, s, min_sim=10, layout=kr.LAYOUT_FULL, row_begin=256, row_end=290)
        assert cells(got) == want_cells(want[256:290], 10, False, row0=256), path
        # packed 8-byte cells: the same sets of cells
        unp = lambda a: {(int(r), int(cc), int(v)) for r, cc, v in zip(*kr.unpack_cells(a))}
        for min_sim in (1, 20, 9999):
            got = c.sim_matrix_sparse64(s, min_sim=min_sim, layout=kr.LAYOUT_TRIANGLE)
            assert len(unp(got)) == len(got) and unp(got) == want_cells(want, min_sim, True), (path, min_sim)
        assert unp(c.sim_matrix_sparse64(s, min_sim=1, layout=kr.LAYOUT_TRIANGLE, cap=16)) == want_cells(want, 1, True), path
        got = c.sim_matrix_sparse64(q, s, min_sim=10, layout=kr.LAYOUT_FULL, row_begin=256, row_end=290)
        assert unp(got) == want_cells(want[256:290], 10, False, row0=256), path
        c.close()


def case_encode_warp():
    """K1, one warp per sequence (counting sort; sorting network for crowded buckets and 513..1024 k-mers): the in-memory path for very long
    sequences, repeats, all K; keys come back sorted and distinct"""
    rng = np.random.default_rng(9)
    alpha = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX*-", dtype=np.uint8)
    lens = [0, 1, 7, 8, 9, 40, 70, 71, 72, 73, 135, 136, 137, 263, 264, 265, 300, 519, 520, 521, 800, 1031, 1032, 1033, 1100, 2500]
    seqs = [alpha[rng.integers(0, len(alpha), n)].tobytes().decode() for n in lens]
    seqs += ["A" * 50, "AC" * 200, "ACDEFGHIK" * 130, seqs[10], seqs[24]]            # repeats (few distinct k-mers), duplicates
    seqs += family(17, 40, 3, 200, 420)
    for K in (8, 3, 11, 1, 2):
        ob = O.batch(seqs, K=K)
        want = O.sim_matrix(ob, ident_inf=True, threads=4)
        c = kr.Context(device=0, K=K, flags=kr.IDENT_IS_INFINITY)
        s = c.encode(seqs)
        assert (s.sizes() == ob.sizes()).all(), K
        for i in range(len(seqs)):
            k = s.keys(i)
            assert (np.diff(k.astype(np.int64)) > 0).all(), (K, i)                 # sorted, distinct
            if i % 4 == 0:
                assert sorted(s.kmers(i)) == sorted(ko.py_kmers(seqs[i], K)), (K, i)
        assert (c.sim_matrix(s, elem=kr.ELEM_U16) == want).all(), K                # equal-sequence rule: hashes agree
        c.close()
    short = [q for q in seqs if len(q) <= 500]                                     # batch served by the 16-register build
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
    s = c.encode(short)
    assert (s.sizes() == O.batch(short, K=8).sizes()).all()
    assert (c.sim_matrix(s, elem=kr.ELEM_U16) == O.sim_matrix(O.batch(short, K=8), ident_inf=True, threads=4)).all()
    c.close()
    c = kr.Context(device=0, K=8, flags=0)
    try:
        c.encode(["MKVLAAGIVG", "MKVLAbGIVGAAAA"])
        raise AssertionError("a residue outside the alphabet must fail")
    except kr.KrError as e:
        assert e.code == -4 and "sequence 1 position 5" in str(e)
    try:      # in a later piece of a pipelined host encode the sequence number is still the batch's
        c.encode(["MKVLAAGIVG"] * 7 + ["MKVLAbGIVGAAAA", "MKV"])
        raise AssertionError("a residue outside the alphabet must fail")
    except kr.KrError as e:
        assert e.code == -4 and "sequence 7 position 5" in str(e)
    bad = np.array([0, 10, 5, 20], dtype=np.uint64)                                # offsets that go backwards
    h = kr.binding.C.c_void_p()
    assert c.lib.kr_encode(c.h, np.zeros(32, np.uint8).ctypes.data, bad.ctypes.data, 3, kr.binding.C.byref(h)) == -1
    assert list(c.encode(["MKVLAAGIVG"] * 5).sizes()) == [3] * 5                  # the context is still usable
    c.close()


def case_subchunks():
    """key-range chunks longer than one pipeline stage (2040 entries): several (A sub-chunk x B sub-chunk) jobs"""
    rng = np.random.default_rng(5)
    alpha = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    roots = [alpha[rng.integers(0, 20, 900)] for _ in range(4)]
    seqs = []
    for i in range(262):
        s = roots[i % 4].copy()
        pos = rng.integers(0, 900, 20 + 15 * (i % 5))
        s[pos] = alpha[rng.integers(0, 20, len(pos))]
        seqs.append(s[: int(rng.integers(700, 901))].tobytes().decode())
    assert np.mean([len(s) for s in seqs]) * 256 / 64 > 2200
    want = O.sim_matrix(O.batch(seqs, K=8), ident_inf=True, threads=8)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
    s = c.encode(seqs)
    assert (c.sim_matrix(s, elem=kr.ELEM_U16) == want).all()
    db = c.db()
    db.append(seqs[:250])
    sim, first, _ = db.find_closest(c.encode(seqs[250:]))
    bs, bi = O.find_closest(O.batch(seqs[:250], K=8), O.batch(seqs[250:], K=8))
    assert (sim == bs).all() and (first == bi).all()
    c.close()


def _clone(bl):
    c = kr.BlockLists()
    for name, _ in kr.BlockLists._fields_:
        setattr(c, name, getattr(bl, name))
    return c


def case_sharded_import():
    """export / import of block lists (the N > 1 encode): two and three shards re-assembled into one set"""
    seqs = family(14, 700, 9, 40, 90)
    res, off = kr.flat(seqs)
    want = O.sim_matrix(O.batch(seqs, K=8), ident_inf=True, threads=4)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
    from kmers.reps_b200 import sharding
    for world in (2, 3):
        cuts = sharding.block_shards(len(seqs), world)
        parts, shards = [], []
        for k in range(world):
            lo, hi = cuts[k], cuts[k + 1]
            p = c.encode_device(res.ctypes.data, off[lo: hi + 1])          # "device" memory is host memory here
            parts.append(p)
            bl = c.export_blocklists(p)
            assert bl.entries <= c.blocklist_capacity(off[lo: hi + 1])
            shards.append(bl)
        full = c.import_blocklists(shards, res.ctypes.data, off)
        got = c.sim_matrix(full, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI)
        assert (got == blocktri(want)).all(), world
    # shards that break the 16-byte rules of the TMA copies are refused (KR_ERR_ARG), nothing is launched on them
    import ctypes
    bad = [_clone(x) for x in shards]
    bad[1].hi = shards[1].hi + 4
    try:
        c.import_blocklists(bad, res.ctypes.data, off)
        raise AssertionError("misaligned hi accepted")
    except kr.KrError as e:
        assert e.code == -1 and "aligned" in str(e), e
    nb = shards[1].nblk + 1
    base = np.ctypeslib.as_array((ctypes.c_uint64 * nb).from_address(shards[1].blk_base)).copy()
    base += 4
    bad = [_clone(x) for x in shards]
    bad[1].blk_base = base.ctypes.data
    try:
        c.import_blocklists(bad, res.ctypes.data, off)
        raise AssertionError("misaligned block start accepted")
    except kr.KrError as e:
        assert e.code == -1 and "layout" in str(e), e
    full = c.import_blocklists(shards, res.ctypes.data, off)
    assert (c.sim_matrix(full, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI) == blocktri(want)).all()
    c.close()


def case_api_misc():
    """the rest of the ABI on small inputs: row stripes of a FULL matrix over three blocks, gathers of gathers, append_set,
    any_ge, the deferred alphabet error of the device encode, statistics"""
    seqs = family(20, 600, 9, 40, 90)
    ob = O.batch(seqs, K=8)
    want = O.sim_matrix(ob, ident_inf=True, threads=4)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
    s = c.encode(seqs)
    for (r0, r1) in ((0, 256), (256, 600), (512, 600)):
        assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_FULL, row_begin=r0, row_end=r1) == want[r0:r1]).all()
        tri = c.sim_matrix(s, elem=kr.ELEM_I32, layout=kr.LAYOUT_TRIANGLE, row_begin=r0, row_end=r1)
        assert (tri == np.concatenate([want[i, i + 1:] for i in range(r0, r1)])).all()
    idx = np.array([599, 3, 3, 258, 0, 511, 300], dtype=np.uint32)
    g1 = s.gather(idx)
    g2 = g1.gather(np.array([6, 0, 2], dtype=np.uint32))
    assert (g1.sizes() == ob.sizes()[idx]).all()
    assert (c.sim_matrix(g1, g2) == want[np.ix_(idx, idx[[6, 0, 2]])]).all()
    assert (c.sim_matrix(g2, s)[1] == want[599]).all()
    db = c.db()
    db.append_set(s, np.arange(100, 400, dtype=np.uint32))
    db.append_set(g1, np.array([0, 4], dtype=np.uint32))                # 599 and 0
    assert db.size == 302
    rep_rows = list(range(100, 400)) + [599, 0]
    sim, first, last = db.find_closest(s)
    sub = want[:, rep_rows]
    assert (sim == sub.max(1)).all()
    assert (first == np.where(sub.max(1) > 0, sub.argmax(1), kr.NO_REP)).all()
    assert (db.any_ge(s, 40) == (sub.max(1) >= 40)).all()
    st = c.stats()
    assert st["kernel_launches"] > 10 and st["pairs"] > 0 and st["join_launches"] > 0
    # a residue outside the alphabet in a device-side encode is reported at the next call that synchronises
    bad, boff = kr.flat(["MKVLAAGIVGLAAA", "MKVLA#GIVGAAAA"])
    sd = c.encode_device(bad.ctypes.data, boff)
    try:
        c.synchronize()
        raise AssertionError("expected the alphabet error")
    except kr.KrError as e:
        assert e.code == -4 and "sequence 1 position 5" in str(e)
    sd.free()
    assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_FULL, row_begin=512, row_end=600) == want[512:]).all()   # still usable
    c.close()


def case_set_ops():
    """kr_seqset_from_keys: the key lists left by SequenceKmers.retainAll / removeAll become a set again; sizes, keys,
    similarities (pairs, matrix through both paths, findClosest) equal plain Python set arithmetic; the equal-sequence rule
    follows the unchanged sequences when they are given"""
    seqs = family(41, 60, 3, 40, 90) + ["MKV", ""]
    seqs[7] = seqs[3]
    K = 8
    sets = [ko.py_kmers(s, K) for s in seqs]
    for path in PATHS:
        c = kr.Context(device=0, K=K, flags=kr.IDENT_IS_INFINITY | path)
        S = c.encode(seqs)
        keys = [S.keys(i) for i in range(len(seqs))]
        other = set(int(k) for k in keys[1])
        kept = [np.array([k for k in ks if int(k) in other], dtype=np.uint64) for ks in keys]        # retainAll(seqs[1])
        gone = [np.array([k for k in ks if int(k) not in other], dtype=np.uint64) for ks in keys]     # removeAll(seqs[1])
        for lists, op in ((kept, set.__and__), (gone, set.__sub__)):
            want_sets = [op(a, sets[1]) for a in sets]
            T = c.from_keys(lists, seqs)
            assert list(T.sizes()) == [len(w) for w in want_sets]
            for i in (0, 1, 3, 5):
                assert sorted(T.kmers(i)) == sorted(want_sets[i])
            # INFINITY needs equal sequences AND equal k-mer sets (the device only looks at pairs whose count equals both sizes):
            # an object whose set was really changed is no longer "identical" to the plain object of the same sequence
            def expect(rows, cols):
                raw = np.array([[len(a & b) for b in cols] for a in rows], dtype=np.int32)
                inf = np.array([[seqs[i] == seqs[j] and rows[i] == cols[j] for j in range(len(cols))] for i in range(len(rows))])
                return np.where(inf, kr.INFINITY, raw)
            assert (c.sim_matrix(T, elem=kr.ELEM_I32) == expect(want_sets, want_sets)).all(), path
            mixed = expect(want_sets, sets)                                                             # mutated rows against plain columns
            assert (c.sim_matrix(T, S, elem=kr.ELEM_I32) == mixed).all(), path
            sim, ident = c.sim_pairs(T, [0, 3, 2], S, [2, 7, 1])
            assert list(sim) == [int(mixed[0, 2]), int(mixed[3, 7]), int(mixed[2, 1])]
            db = c.db(); db.append_set(T, np.arange(len(seqs), dtype=np.uint32))
            s2, f2, _ = db.find_closest(S)
            col = mixed.T                                                                               # rows = plain queries, columns = mutated reps
            assert (s2 == col.max(axis=1)).all() and all(col[q, f2[q]] == s2[q] for q in range(len(seqs)) if s2[q] > 0), path
        U = c.from_keys(kept)                                                                           # without sequences: equal only to itself
        m = c.sim_matrix(U, elem=kr.ELEM_I32)
        want = np.array([[len(a & b) for b in [x & sets[1] for x in sets]] for a in [x & sets[1] for x in sets]], dtype=np.int32)
        np.fill_diagonal(want, kr.INFINITY)
        assert (m == want).all(), path
        for bad in ([np.array([5, 5], np.uint64)], [np.array([9, 3], np.uint64)], [np.array([1 << 40], np.uint64)]):
            try:
                c.from_keys(bad)
                raise AssertionError("unsorted / out-of-range keys accepted")
            except kr.KrError as e:
                assert e.code == -1
        c.close()


def case_multi():
    """several contexts behind one kr_multi handle (KR_EMUL_DEVICES "GPUs", one worker thread each): matrices in every
    layout, sparse cells with too little room per slice, rectangular, findClosest with ranks, any_ge, the greedy build with
    the scoring split -- all equal to the oracle, i.e. to the single-GPU calls"""
    ndev = int(os.environ.get("KR_EMUL_DEVICES", "1"))
    devs = list(range(ndev)) if ndev > 1 else [0, 0, 0]           # one device listed three times: three contexts on it
    seqs = family(31, 800, 10, 50, 110) + ["MKV", "", "MKVLAAGIVG", "MKVLAAGIVG", "MKV"]
    seqs[9] = seqs[700]
    n = len(seqs)
    ob = O.batch(seqs, K=8)
    want = O.sim_matrix(ob, ident_inf=True, threads=4)
    for path in PATHS:
        m = kr.MultiContext(devs, K=8, flags=kr.IDENT_IS_INFINITY | path)
        assert m.ndev == len(devs)
        assert (m.sim_matrix(seqs, elem=kr.ELEM_I32, layout=kr.LAYOUT_FULL) == want).all(), path
        tri = m.sim_matrix(seqs, elem=kr.ELEM_U16, layout=kr.LAYOUT_TRIANGLE)
        assert (tri == np.concatenate([want[i, i + 1:] for i in range(n)])).all(), path
        assert (m.sim_matrix(seqs, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI) == blocktri(want)).all(), path
        got = m.sim_matrix_sparse(seqs, min_sim=1, layout=kr.LAYOUT_TRIANGLE, cap=1 << 16)
        assert len(cells(got)) == len(got) and cells(got) == want_cells(want, 1, True), path
        full = want_cells(want, 20, False)
        got = m.sim_matrix_sparse(seqs, min_sim=20, layout=kr.LAYOUT_FULL, cap=len(full))     # exact room: slices overflow, the total fits
        assert cells(got) == full, path
        try:
            m.lib.kr_multi_sim_matrix_sparse                                                   # too little room in total: refused with the count
            c = np.zeros((8, 4), np.uint32); cnt = kr.binding.C.c_uint64(0)
            r, o = kr.flat(seqs)
            rc = m.lib.kr_multi_sim_matrix_sparse(m.h, r.ctypes.data, o.ctypes.data, n, None, None, 0, 20, kr.LAYOUT_FULL, c.ctypes.data, 8, kr.binding.C.byref(cnt))
            assert rc == -1 and cnt.value == len(full)
        finally:
            pass
        # rectangular: 150 query rows against 500 columns
        qa, qb = seqs[600:750], seqs[:500]
        wr = O.sim_matrix(O.batch(qa, K=8), O.batch(qb, K=8), ident_inf=True, threads=4)
        assert (m.sim_matrix(qa, b_seqs=qb, elem=kr.ELEM_I32) == wr).all(), path
        got = m.sim_matrix_sparse(qa, b_seqs=qb, min_sim=3, layout=kr.LAYOUT_FULL)
        assert cells(got) == want_cells(wr, 3, False), path
        # database replicated, queries split
        reps, qs = seqs[:500], seqs[500:]
        rank = np.random.default_rng(4).permutation(len(reps)).astype(np.uint32)
        order = np.argsort(rank).astype(np.int32)
        bs, bi = O.find_closest(O.batch(reps, K=8), O.batch(qs, K=8), order=order, ident_inf=True)
        bs2, bl2 = O.find_closest(O.batch(reps, K=8), O.batch(qs, K=8), order=order, ident_inf=True, tie_last=True)
        db = m.db()
        db.append(reps[:200]); db.append(reps[200:])
        db.set_ranks(rank)
        assert db.size == 500
        sim, first, last = db.find_closest(qs)
        assert (sim == bs).all() and (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all(), path
        assert (last.astype(np.int64) == np.where(bs2 > 0, bl2.astype(np.int64), kr.NO_REP)).all(), path
        assert (db.any_ge(30, qs) == (bs >= 30)).all()
        sim1, _, _ = db.find_closest(qs[:2])                         # fewer queries than contexts
        assert (sim1 == bs[:2]).all()
        db.free()
        # greedy build, scoring split over the contexts
        for thr, batch in ((30, 64), (10, 0)):
            _, wa, _ = O.greedy_build(ob, thr, ident_inf=True)
            db = m.db()
            a0 = db.greedy_build(thr, seqs[:300], batch=batch)
            a1 = db.greedy_build(thr, seqs[300:], batch=batch)
            assert (np.concatenate([a0, a1]) == wa).all() and db.size == int(wa.sum()), (path, thr)
            db.free()
        st = m.stats()
        assert st["kernel_launches"] > 0 and st["pairs"] > 0
        m.close()
    try:
        kr.MultiContext([0, 99], K=8)
        raise AssertionError("device 99 accepted")
    except kr.KrError as e:
        assert "device 99" in str(e)


CASES = {k[5:]: v for k, v in globals().items() if k.startswith("case_")}

if __name__ == "__main__":
    name = sys.argv[1]
    CASES[name]()
    print("OK", name)
