"""TEST INFRASTRUCTURE: runs the product's device + host code on the CPU emulator (tests/emul/cuda_emul.h) and
compares with the oracle.  Started by tests/test_emul_device_code.py in a subprocess with KR_LIB_PATH pointing at
tests/emul/_build/libkmersreps_emul.so (the product sources compiled by g++ with -DKR_EMUL -DKR_EXPERIMENTAL).

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


def case_matrix_layouts():
    """encode + block lists + TMA-staged tile join: every layout / element type, two blocks, the equal-sequence rule"""
    seqs = family(11, 300, 6, 60, 140) + ["MKV", "", "MKVLAAGIVG", "MKVLAAGIVG"]
    seqs[7] = seqs[290]                               # equal sequences in different blocks
    for ident in (kr.IDENT_IS_INFINITY, 0):
        want = O.sim_matrix(O.batch(seqs, K=8), ident_inf=bool(ident), threads=4)
        c = kr.Context(device=0, K=8, flags=ident)
        s = c.encode(seqs)
        assert (s.sizes() == O.batch(seqs, K=8).sizes()).all()
        assert (c.sim_matrix(s, elem=kr.ELEM_I32, layout=kr.LAYOUT_FULL) == want).all()
        assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_TRIANGLE) == want[np.triu_indices(len(seqs), 1)]).all()
        assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI) == blocktri(want)).all()
        hi = c.sim_matrix(s, elem=kr.ELEM_I32, layout=kr.LAYOUT_BLOCKTRI, row_begin=256, row_end=len(seqs))
        assert (hi == blocktri(want)[-len(hi):]).all()
        c.close()


def case_k10_rectangular_pairs():
    """32-bit low words (K = 10, three pipeline stages), rectangular matrix, explicit pairs, key decoding"""
    seqs = family(12, 90, 3, 50, 120)
    c = kr.Context(device=0, K=10, flags=0)
    a, b = c.encode(seqs[:40]), c.encode(seqs[30:])
    oa, ob = O.batch(seqs[:40], K=10), O.batch(seqs[30:], K=10)
    assert (c.sim_matrix(a, b) == O.sim_matrix(oa, ob, ident_inf=False)).all()
    ia, ib = np.arange(40, dtype=np.uint32), np.arange(40, dtype=np.uint32)[::-1].copy()
    sim, ident = c.sim_pairs(a, ia, b, ib)
    full = O.sim_matrix(oa, ob, ident_inf=False)
    assert (sim == full[ia, ib]).all()
    assert sorted(a.kmers(3)) == sorted(ko.py_kmers(seqs[3], 10))
    c.close()


def case_find_closest_and_build():
    """ARGMAX join, inverted index (forced), ranks / ties, greedy build with database compaction"""
    seqs = family(13, 330, 8, 60, 130)
    reps, qs = seqs[:270], seqs[270:] + [seqs[5], "MKV"]
    ob, oq = O.batch(reps, K=8), O.batch(qs, K=8)
    rng = np.random.default_rng(3)
    rank = rng.permutation(len(reps)).astype(np.uint32)
    order = np.argsort(rank)
    bs, bi = O.find_closest(ob, oq, order=order.astype(np.int32), ident_inf=True)
    for force in (kr.FORCE_JOIN, kr.FORCE_INDEX):
        c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | force)
        db = c.db()
        db.append(reps[:100]); db.append(reps[100:])
        db.set_ranks(rank)
        sim, first, last = db.find_closest(c.encode(qs))
        assert (sim == bs).all(), force
        assert (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all(), force
        c.close()
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
    db = c.db()
    added = db.greedy_build(c.encode(seqs), 30, batch=40)      # many small segments: db_compact runs
    _, want, _ = O.greedy_build(O.batch(seqs, K=8), 30, ident_inf=True)
    assert (added == want).all() and db.size == int(want.sum())
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


def case_exp_sparse_join():
    """EXPERIMENTAL: sparse result from the join's epilogue filter"""
    seqs = family(15, 300, 6, 60, 120)
    seqs[5] = seqs[280] = "MKV"
    want = O.sim_matrix(O.batch(seqs, K=8), ident_inf=True, threads=4)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
    s = c.encode(seqs)
    for min_sim in (1, 20, 9999):
        got = c.sim_matrix_sparse(s, min_sim=min_sim, layout=kr.LAYOUT_TRIANGLE)
        assert len(cells(got)) == len(got) and cells(got) == want_cells(want, min_sim, True)
    got = c.sim_matrix_sparse(s, min_sim=1, layout=kr.LAYOUT_TRIANGLE, cap=16)       # too little room: retried
    assert cells(got) == want_cells(want, 1, True)
    q = c.encode(seqs[:290])
    got = c.sim_matrix_sparse(q, s, min_sim=10, layout=kr.LAYOUT_FULL, row_begin=256, row_end=290)
    assert cells(got) == want_cells(want[256:290], 10, False, row0=256)
    c.close()


def case_exp_pairs():
    """EXPERIMENTAL: posting-list all-vs-all (dense, every layout) and the pair-table sparse result, end to end"""
    seqs = family(16, 300, 6, 60, 120) + ["MKV", "", "MKVLAAGIVG", "MKVLAAGIVG", "MKV"]
    seqs[7] = seqs[290]
    n = len(seqs)
    for ident in (kr.IDENT_IS_INFINITY, 0):
        want = O.sim_matrix(O.batch(seqs, K=8), ident_inf=bool(ident), threads=4)
        for mode in (binding.EXP_PAIRS, binding.EXP_PAIRS_AUTO, binding.EXP_PAIRS | binding.EXP_BUCKETS, binding.EXP_PAIRS_AUTO | binding.EXP_BUCKETS):
            if os.environ.get("KR_EMUL_ONLY_BUCKETS") and not (mode & binding.EXP_BUCKETS):
                continue
            c = kr.Context(device=0, K=8, flags=ident | mode)
            s = c.encode(seqs)
            assert (c.sim_matrix(s, elem=kr.ELEM_I32, layout=kr.LAYOUT_FULL) == want).all()
            assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_TRIANGLE) == want[np.triu_indices(n, 1)]).all()
            assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI) == blocktri(want)).all()
            lo = c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI, row_begin=0, row_end=256)
            hi = c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI, row_begin=256, row_end=n)
            assert (np.concatenate([lo, hi]) == blocktri(want)).all()
            if ident:
                for min_sim in (1, 15):
                    got = c.sim_matrix_sparse(s, min_sim=min_sim, layout=kr.LAYOUT_TRIANGLE)
                    assert len(cells(got)) == len(got) and cells(got) == want_cells(want, min_sim, True)
                got = c.sim_matrix_sparse(s, min_sim=1, layout=kr.LAYOUT_TRIANGLE, row_begin=256, row_end=n)
                assert cells(got) == want_cells(want[256:], 1, True, row0=256)
                db = c.db()
                added = db.greedy_build(s, 30, batch=64)                 # survivor matrices through the same path
                _, w2, _ = O.greedy_build(O.batch(seqs, K=8), 30, ident_inf=True)
                assert (added == w2).all()
            c.close()


def case_exp_encode_warp():
    """EXPERIMENTAL: K1 with one warp per sequence (register bitonic network): every register budget, the in-memory path
    for very long sequences, repeats, all K"""
    rng = np.random.default_rng(9)
    alpha = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX*-", dtype=np.uint8)
    lens = [0, 1, 7, 8, 9, 40, 70, 71, 72, 73, 135, 136, 137, 263, 264, 265, 300, 519, 520, 521, 800, 1031, 1032, 1033, 1100, 2500]
    seqs = [alpha[rng.integers(0, len(alpha), n)].tobytes().decode() for n in lens]
    seqs += ["A" * 50, "AC" * 200, "ACDEFGHIK" * 130, seqs[10], seqs[24]]            # repeats (few distinct k-mers), duplicates
    seqs += family(17, 40, 3, 200, 420)
    for K in (8, 3, 11):
        ob = O.batch(seqs, K=K)
        want = O.sim_matrix(ob, ident_inf=True, threads=4)
        c = kr.Context(device=0, K=K, flags=kr.IDENT_IS_INFINITY | binding.EXP_ENCODE)
        s = c.encode(seqs)
        assert (s.sizes() == ob.sizes()).all(), K
        ref = kr.Context(device=0, K=K, flags=kr.IDENT_IS_INFINITY)
        s0 = ref.encode(seqs)
        for i in range(len(seqs)):
            assert (s.keys(i) == s0.keys(i)).all(), (K, i)
        assert (c.sim_matrix(s, elem=kr.ELEM_U16) == want).all(), K                # equal-sequence rule: same hashes
        c.close(); ref.close()
    short = [q for q in seqs if len(q) <= 500]                                     # batch served by the 16-register build
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | binding.EXP_ENCODE)
    s = c.encode(short)
    assert (s.sizes() == O.batch(short, K=8).sizes()).all()
    assert (c.sim_matrix(s, elem=kr.ELEM_U16) == O.sim_matrix(O.batch(short, K=8), ident_inf=True, threads=4)).all()
    c.close()
    c = kr.Context(device=0, K=8, flags=binding.EXP_ENCODE)
    try:
        c.encode(["MKVLAAGIVG", "MKVLAbGIVGAAAA"])
        raise AssertionError("a residue outside the alphabet must fail")
    except kr.KrError as e:
        assert e.code == -4 and "sequence 1 position 5" in str(e)
    c.close()


def case_exp_idxq_warp():
    """EXPERIMENTAL: K5 with one warp per query -- dense per-warp scores (small database), per-warp hash (large database,
    short posting lists), and the multi-pass path of a query that touches more representatives than the hash holds"""
    rng = np.random.default_rng(21)
    alpha = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    rnd = lambda n: alpha[rng.integers(0, 20, n)].tobytes().decode()
    # (a) dense: a family-structured database of 300, ties and ranks, identical and k-mer-less queries
    seqs = family(18, 340, 8, 60, 130)
    reps, qs = seqs[:300], seqs[300:] + [seqs[5], "MKV", ""]
    reps[20] = "MKV"
    rank = rng.permutation(len(reps)).astype(np.uint32)
    bs, bi = O.find_closest(O.batch(reps, K=8), O.batch(qs, K=8), order=np.argsort(rank).astype(np.int32), ident_inf=True)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | kr.FORCE_INDEX | binding.EXP_IDXQ)
    db = c.db(); db.append(reps); db.set_ranks(rank)
    sim, first, last = db.find_closest(c.encode(qs))
    assert (sim == bs).all() and (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all()
    bs2, bl2 = O.find_closest(O.batch(reps, K=8), O.batch(qs, K=8), order=np.argsort(rank).astype(np.int32), ident_inf=True, tie_last=True)
    assert (last.astype(np.int64) == np.where(bs2 > 0, bl2.astype(np.int64), kr.NO_REP)).all()
    c.close()
    # (b) hash: 4,300 short representatives (> IDXQ2_DENSE_MAX), 1,150 of which share one 8-mer with the hub queries
    hub = "MKVLAAGI"
    reps = [rnd(24) for _ in range(3150)] + [rnd(8) + hub + rnd(8) for _ in range(1150)]
    fam = [reps[i][:12] + rnd(6) + reps[i][18:] for i in (5, 17, 3300, 4000)]          # near copies: clear best matches
    qs = fam + [rnd(10) + hub + rnd(10), hub, reps[77], rnd(30), reps[3200][8:16] + rnd(5)]
    order = rng.permutation(len(reps)).astype(np.uint32)
    bs, bi = O.find_closest(O.batch(reps, K=8), O.batch(qs, K=8), order=np.argsort(order).astype(np.int32), ident_inf=True, threads=4)
    bs2, bl2 = O.find_closest(O.batch(reps, K=8), O.batch(qs, K=8), order=np.argsort(order).astype(np.int32), ident_inf=True, tie_last=True, threads=4)
    for flags in (binding.EXP_IDXQ, 0):                                              # 0: the CTA kernel, same expectations
        c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | kr.FORCE_INDEX | flags)
        db = c.db(); db.append(reps); db.set_ranks(order)
        sim, first, last = db.find_closest(c.encode(qs))
        assert (sim == bs).all(), flags
        assert (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all(), flags
        assert (last.astype(np.int64) == np.where(bs2 > 0, bl2.astype(np.int64), kr.NO_REP)).all(), flags
        c.close()
    assert bs[5] == 1 and bi[5] != bl2[5], (bs, bi, bl2)                             # the hub query ties over 1,150 representatives


def case_exp_gidx_build():
    """EXPERIMENTAL: greedy build through the growing index -- against the oracle and the default build, starting from a
    non-empty database, with k-mer-less and duplicate candidates, several batch sizes, and a candidate whose k-mers
    reach more representatives than the per-warp table holds (multi-pass path)"""
    rng = np.random.default_rng(31)
    alpha = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    rnd = lambda n: alpha[rng.integers(0, 20, n)].tobytes().decode()
    seqs = family(19, 420, 9, 60, 130)
    seqs[50:50] = ["MKV", "MKV", "", "MKVLAAGIVG", "MKVLAAGIVG"]
    first, rest = seqs[:120], seqs[120:] + ["MKV", seqs[3]]
    for thr in (30, 10):
        _, w0, _ = O.greedy_build(O.batch(first + rest, K=8), thr, ident_inf=True)
        for batch in (37, 4096):
            c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | binding.EXP_GIDX)
            db = c.db()
            a0 = db.greedy_build(c.encode(first), thr, batch=batch)             # fills the database ...
            a1 = db.greedy_build(c.encode(rest), thr, batch=batch)              # ... which the second call must index first
            assert (np.concatenate([a0, a1]) == w0).all(), (thr, batch)
            assert db.size == int(w0.sum())
            c.close()
    hub = "MKVLAAGI"
    crowd = [rnd(9) + hub + rnd(9) for _ in range(1150)]                        # pairwise similarity 1: all become representatives
    cands = crowd + [rnd(5) + hub + rnd(30), crowd[7][:20] + rnd(4), rnd(40)]
    _, want, _ = O.greedy_build(O.batch(cands, K=8), 5, ident_inf=True)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | binding.EXP_GIDX)
    db = c.db()
    got = db.greedy_build(c.encode(cands), 5, batch=400)
    assert (got == want).all() and want.sum() > 1100 and not want[1151]      # > 1024 representatives on the hub k-mer's list
    c.close()


def dna_family(seed, n, length):
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"acgt", dtype=np.uint8)
    roots = [alpha[rng.integers(0, 4, length)] for _ in range(4)]
    seqs = []
    for i in range(n):
        s = roots[i % 4].copy()
        pos = rng.integers(0, length, 2 + 6 * (i % 5))
        s[pos] = alpha[rng.integers(0, 4, len(pos))]
        t = s[: int(rng.integers(length * 2 // 3, length + 1))].tobytes().decode()
        if i % 7 == 3:
            t = ko.py_dna_revcomp(t)                       # the other strand of a relative: same k-mer set
        if i % 5 == 1:
            t = t.upper()                                 # case does not matter
        if i % 11 == 4:
            t = t[:40] + "n" + t[41:]
        seqs.append(t)
    return seqs


def case_exp_dna():
    """EXPERIMENTAL: DnaKmers contexts -- K-mers of both strands through the unchanged join / index / posting-list paths"""
    seqs = dna_family(51, 60, 420) + ["acgtacgtac", "", "ACGTACGTAC", "acgu" * 12, "tttttttttttttttttttttttttttt", "aaaaaaaaaaaaaaaaaaaaaaaaaaaa"]
    seqs[9] = seqs[2].upper()                              # equal sequences up to case: INFINITY
    for K in (18, 15, 7):
        want = np.array([[ko.py_dna_similarity(a, b, K, True) for b in seqs] for a in seqs], dtype=np.int64)
        sizes = np.array([len(ko.py_dna_kmers(q, K)) for q in seqs])
        for extra in (0, binding.EXP_PAIRS | binding.EXP_BUCKETS):
            c = kr.Context(device=0, K=K, flags=kr.IDENT_IS_INFINITY | binding.ALPHABET_DNA | extra)
            s = c.encode(seqs)
            assert (s.sizes() == sizes).all(), (K, extra)
            assert (c.sim_matrix(s, elem=kr.ELEM_I32) == want).all(), (K, extra)
            assert sorted(s.kmers(0)) == sorted(ko.py_dna_kmers(seqs[0], K))
            db = c.db()
            db.append(seqs[:40])
            sim, first, _ = db.find_closest(c.encode(seqs[40:]))
            sub = want[40:, :40]
            assert (sim == sub.max(1)).all() and (first == np.where(sub.max(1) > 0, sub.argmax(1), kr.NO_REP)).all()
            c.close()
    assert want[-2, -1] == len(ko.py_dna_kmers("t" * 28, 7)) == 2      # poly-t and poly-a are each other's reverse complement
    c = kr.Context(device=0, K=15, flags=binding.ALPHABET_DNA)
    try:
        c.encode(["acgtacgtacgtacgtacgt", "acgtacgtrcgtacgtacgt"])
        raise AssertionError("an ambiguity code must fail")
    except kr.KrError as e:
        assert e.code == -4 and "sequence 1 position 8" in str(e)
    c.close()
    for badK in (4, 19):
        try:
            kr.Context(device=0, K=badK, flags=binding.ALPHABET_DNA)
            raise AssertionError("K outside [5,18] must fail")
        except kr.KrError:
            pass


CASES = {k[5:]: v for k, v in globals().items() if k.startswith("case_")}

if __name__ == "__main__":
    name = sys.argv[1]
    CASES[name]()
    print("OK", name)
