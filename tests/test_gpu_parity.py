"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI of
libkmersreps_cuda.so (ctypes) and is compared bit-for-bit with the CPU oracle on the same inputs:
the reference's fixtures (rep50.db, VapbProt.fa, GTO seeds, rickettsia duplicates) and seeded
synthetic families, including the edge cases the domain has (empty / shorter-than-K sequences,
identical sequences, ragged last blocks, X-rich sequences, every K the packing supports)."""
import os

import numpy as np
import pytest

import kmers.reps_b200 as kr
from kmers.reps_b200 import synth
from oracle import kr_oracle as ko

pytestmark = pytest.mark.gpu
# "device" tensors: CUDA -- or host memory when KR_LIB_PATH names the CPU emulator build of the library (tools/emul_gpu_suite.sh
# runs this file against it so that the tests' own expectations are exercised before they meet a GPU)
EMUL = "libkmersreps_emul" in kr.lib_path()
DEV = "cpu" if EMUL else "cuda"


# B200-sized inputs of the at-scale tests; tools/emul_gpu_suite.sh shrinks them (KR_TEST_SCALE=<divisor>) so that the tests' own
# logic -- index formulas, invariants -- runs on the emulator before it meets a GPU
SCALE = max(1, int(os.environ.get("KR_TEST_SCALE", "1")))


def dsync():
    """torch fills its tensors on ITS stream; the library works on its own non-blocking stream: finish torch's work first"""
    if not EMUL:
        import torch
        torch.cuda.synchronize()


# every test that takes `ctx` / `ctx_raw` runs three times: the library's own path choice, the tile join pinned, and the
# sorted index + warp scan pinned (include/kmersreps.h, KR_FORCE_*): both device paths must equal the oracle
PATHS = {"auto": 0, "join": kr.FORCE_JOIN, "index": kr.FORCE_INDEX}


@pytest.fixture(scope="module", params=list(PATHS))
def ctx(request):
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | PATHS[request.param])
    yield c
    c.close()


@pytest.fixture(scope="module", params=list(PATHS))
def ctx_raw(request):
    c = kr.Context(device=0, K=8, flags=PATHS[request.param])
    yield c
    c.close()


def tri_from_full(m):
    iu = np.triu_indices(m.shape[0], 1)
    return m[iu]


def test_kat_k10(expected):
    k = expected["kat_k10"]
    c = kr.Context(device=0, K=10, flags=kr.IDENT_IS_INFINITY)
    s = c.encode([k["p1"], k["p2"], k["p3"]])
    assert list(s.sizes()) == [41, 41, 20]
    m = c.sim_matrix(s)
    assert m[1, 2] == 3 and m[2, 1] == 3 and m[0, 1] == 0 and m[0, 0] == kr.INFINITY
    sim, ident = c.sim_pairs(s, [1, 2, 0, 0], s, [2, 1, 1, 0])
    assert list(sim) == [3, 3, 0, kr.INFINITY] and list(ident) == [0, 0, 0, 1]
    assert kr.distance(3, 41, 20) == k["dist_p2_p3"] and kr.distance(0, 41, 41) == 1.0 and kr.distance(9999, 1, 1) == 0.0
    # SequenceKmers.iterator(): decoded keys are exactly the distinct substrings
    assert sorted(s.kmers(2)) == sorted(ko.py_kmers(k["p3"], 10))
    c.close()


def test_vapb_matrix_all_layouts(ctx, vapb, oracle, expected):
    seqs = [s for (_, _, s) in vapb]
    b = oracle.batch(seqs, K=8)
    want = oracle.sim_matrix(b, ident_inf=True)
    want_raw = oracle.sim_matrix(b, ident_inf=False)
    ctx_raw = kr.Context(device=0, K=8, flags=ctx.flags & ~kr.IDENT_IS_INFINITY)
    s = ctx.encode(seqs)
    assert (s.sizes() == b.sizes()).all()
    n = len(seqs)
    for elem in (kr.ELEM_I32, kr.ELEM_U16):
        full = ctx.sim_matrix(s, elem=elem, layout=kr.LAYOUT_FULL)
        assert (full == want).all()
        tri = ctx.sim_matrix(s, elem=elem, layout=kr.LAYOUT_TRIANGLE)
        assert (tri == tri_from_full(want)).all()
        bt = ctx.sim_matrix(s, elem=elem, layout=kr.LAYOUT_BLOCKTRI)
        parts = []
        for i in range(n):
            parts.append(want[i, 256 * (i // 256):])
        assert (bt == np.concatenate(parts)).all()
    sr = ctx_raw.encode(seqs)
    raw = ctx_raw.sim_matrix(sr)
    assert (raw == want_raw).all()
    e = expected["config2"]
    t = tri_from_full(raw)
    assert int(t.sum()) == e["sum_sim"] and int((t > 0).sum()) == e["nonzero_pairs"] and int(t.max()) == e["max_raw_sim"]
    assert int((tri_from_full(full) == kr.INFINITY).sum()) == e["identical_pairs"] == 11
    # the same matrix as sparse cells (j > i, sim >= 1): the 1,610 non-zero pairs, INFINITY cells included
    cells = ctx.sim_matrix_sparse(s, min_sim=1, layout=kr.LAYOUT_TRIANGLE)
    ii, jj = np.nonzero(np.triu(want, 1))
    assert {(int(r), int(c), int(v)) for r, c, v, _ in cells} == {(int(i), int(j), int(want[i, j])) for i, j in zip(ii, jj)}
    assert len(cells) == e["nonzero_pairs"] == 1610
    ctx_raw.close()


def test_rep50_matrix_and_stripes(ctx, rep50, oracle, expected):
    seqs = [s for (_, _, s) in rep50["reps"]]
    b = oracle.batch(seqs, K=8)
    want = oracle.sim_matrix(b, ident_inf=True, threads=8)
    s = ctx.encode(seqs)
    assert (s.sizes() == b.sizes()).all()
    assert int(s.sizes().sum()) == expected["rep50"]["total_kmers"]
    full = ctx.sim_matrix(s, elem=kr.ELEM_U16)
    assert (full == want).all()
    t = tri_from_full(full.astype(np.int64))
    assert t.max() == 49 and int((t == 49).sum()) == 92 and int((t > 0).sum()) == expected["rep50"]["nonzero_pairs"]
    # row stripes (the multi-GPU sharding unit) reassemble to the same matrix, in every layout
    for layout in (kr.LAYOUT_FULL, kr.LAYOUT_TRIANGLE, kr.LAYOUT_BLOCKTRI):
        whole = ctx.sim_matrix(s, elem=kr.ELEM_U16, layout=layout).ravel()
        for nsh in (2, 3):
            rb = kr.matrix_stripes(len(seqs), len(seqs), layout, nsh)
            parts = [ctx.sim_matrix(s, elem=kr.ELEM_U16, layout=layout, row_begin=int(rb[k]), row_end=int(rb[k + 1])).ravel()
                     for k in range(nsh)]
            assert (np.concatenate(parts) == whole).all()


def test_rectangular_matrix_and_pairs(ctx, rep50, gto_seeds, rickettsia, oracle):
    reps = [s for (_, _, s) in rep50["reps"]][:700]
    qs = [s for (_, _, s) in gto_seeds] + [p for (_, p) in rickettsia]
    A, B = ctx.encode(qs), ctx.encode(reps)
    oa, ob = oracle.batch(qs, K=8), oracle.batch(reps, K=8)
    want = oracle.sim_matrix(oa, ob, ident_inf=True, threads=4)
    got = ctx.sim_matrix(A, B)
    assert got.shape == want.shape and (got == want).all()
    rng = np.random.default_rng(5)
    ia = rng.integers(0, len(qs), 5000)
    ib = rng.integers(0, len(reps), 5000)
    sim, ident = ctx.sim_pairs(A, ia, B, ib)
    assert (sim == want[ia, ib]).all() and (ident == (want[ia, ib] == kr.INFINITY)).all()
    # duplicate-heavy real set against itself: 66 duplicate records -> INFINITY cells
    R = ctx.encode([p for (_, p) in rickettsia])
    orr = oracle.batch([p for (_, p) in rickettsia], K=8)
    assert (ctx.sim_matrix(R) == oracle.sim_matrix(orr, ident_inf=True)).all()


def test_config1_find_closest(ctx, rep50, gto_seeds, oracle, expected):
    reps = [s for (_, _, s) in rep50["reps"]]
    db = ctx.db()
    db.append(reps[:1000])
    db.append(reps[1000:])                  # two segments
    assert db.size == 1480
    q = ctx.encode([s for (_, _, s) in gto_seeds])
    sim, first, last = db.find_closest(q)
    for i, e in enumerate(expected["config1"]):
        assert sim[i] == e["sim"] and first[i] == e["rep_index"] == last[i]
        sz = q.sizes()[i]
        assert sz == e["size"]
    # rep50.db queried with its own members: each finds itself with INFINITY
    qq = ctx.encode(reps)
    sim, first, last = db.find_closest(qq)
    assert (sim == kr.INFINITY).all() and (first == np.arange(1480)).all() and (last == first).all()
    # any_ge is the checkSimilarity of the build loop
    assert db.any_ge(q, 50).all() and not db.any_ge(q, 400).any()


def test_find_closest_ties_and_ranks(ctx_raw, oracle):
    # duplicates in the database create exact ties; the rank permutation decides first/last
    res, off, _ = synth.generate(77, 900, 9, dup_per_10k=2500)
    seqs = synth.to_strings(res, off)
    reps, qs = seqs[:600], seqs[600:]
    rng = np.random.default_rng(3)
    order = rng.permutation(len(reps)).astype(np.int32)       # order[k] = rep visited k-th
    rank = np.empty(len(reps), dtype=np.uint32)
    rank[order] = np.arange(len(reps), dtype=np.uint32)
    db = ctx_raw.db()
    db.append(reps)
    db.set_ranks(rank)
    q = ctx_raw.encode(qs)
    sim, first, last = db.find_closest(q)
    orp, oq = oracle.batch(reps, K=8), oracle.batch(qs, K=8)
    bs, bi = oracle.find_closest(orp, oq, order=order, ident_inf=False)
    bs2, bi2 = oracle.find_closest(orp, oq, order=order, ident_inf=False, tie_last=True)
    assert (sim == bs).all() and (sim == bs2).all()
    assert (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all()
    assert (last.astype(np.int64) == np.where(bs2 > 0, bi2.astype(np.int64), kr.NO_REP)).all()
    assert (first != last).any()            # the case is exercised


def test_edge_cases_and_all_k(oracle):
    seqs = ["", "A", "ACDEFGH", "ACDEFGHI", "ACDEFGHIK", "AAAAAAAAAAAAAAAAAAAA", "ACDEFGHI", "XXXXXXXXXXXXXXXXXXXXXXXXX",
            "MKV*LL-AAGT*", "BJOUZXBJOUZXBJOUZX", "ACDEFGH", ""]
    res, off, _ = synth.generate(5, 300, 5)
    seqs += synth.to_strings(res, off)
    for K, flags in [(K, f | pth) for K in (1, 2, 3, 5, 8, 9, 11) for f in (0, kr.IDENT_IS_INFINITY) for pth in PATHS.values()]:
        if True:
            c = kr.Context(device=0, K=K, flags=flags)
            s = c.encode(seqs)
            b = oracle.batch(seqs, K=K)
            assert (s.sizes() == b.sizes()).all()
            want = oracle.sim_matrix(b, ident_inf=bool(flags & kr.IDENT_IS_INFINITY))
            assert (c.sim_matrix(s) == want).all(), (K, flags)
            for i in (2, 3, 5, 8, 9):
                assert sorted(s.kmers(i)) == sorted(ko.py_kmers(seqs[i], K))
            c.close()
    # empty batch and a batch of only empty sequences
    c = kr.Context(device=0, K=8)
    e = c.encode([])
    assert e.count == 0 and c.sim_matrix(e).size == 0
    z = c.encode(["", "AC", ""])
    assert list(z.sizes()) == [0, 0, 0]
    m = c.sim_matrix(z)
    assert m[0, 2] == kr.INFINITY and m[0, 1] == 0 and m[1, 1] == kr.INFINITY
    c.close()


def test_error_paths():
    c = kr.Context(device=0, K=8)
    with pytest.raises(kr.KrError) as ei:
        c.encode(["ACDEFGHIK", "ACDEFGHIKacdef"])
    assert ei.value.code == -4 and "sequence 1 position 9" in str(ei.value) and "'a'" in str(ei.value)
    # the context stays usable after an error
    s = c.encode(["ACDEFGHIKL"])
    assert list(s.sizes()) == [3]
    with pytest.raises(kr.KrError):
        kr.Context(device=0, K=12)
    with pytest.raises(kr.KrError):
        kr.Context(device=0, K=0)
    with pytest.raises(kr.KrError):
        c.sim_matrix(s, layout=kr.LAYOUT_TRIANGLE, row_begin=3)
    other = kr.Context(device=0, K=8)
    with pytest.raises(kr.KrError):
        other.sim_matrix(s)
    other.close()
    # device-pointer entry points refuse pointers their stores could not take (a misaligned store is a device fault, not a status)
    import torch
    buf = torch.zeros(4096, dtype=torch.uint8, device=DEV)
    dsync()
    base = buf.data_ptr()
    with pytest.raises(kr.KrError):
        c.sim_matrix_sparse_device(s, None, base + 8, 16)
    with pytest.raises(kr.KrError):
        c.sim_matrix_device(s, None, base + 1, elem=kr.ELEM_U16)
    with pytest.raises(kr.KrError):
        c.sim_matrix_device(s, None, base + 2, elem=kr.ELEM_I32)
    db = c.db(); db.append(["ACDEFGHIKLMN"])
    with pytest.raises(kr.KrError):
        db.find_closest_device(s, base + 2)
    c.sim_matrix_device(s, None, base + 2, elem=kr.ELEM_U16)          # element-aligned is enough for a dense result
    c.synchronize()
    c.close()


def test_long_sequences(oracle):
    # longer than the shared-memory sort capacity (2048 keys) and than one pipeline stage per chunk
    rng = np.random.default_rng(9)
    alpha = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    base = alpha[rng.integers(0, 20, 9000)]
    seqs = []
    for i in range(6):
        s = base.copy()
        pos = rng.integers(0, len(s), 300 * i)
        s[pos] = alpha[rng.integers(0, 20, len(pos))]
        seqs.append(s[: 9000 - 500 * i].tobytes().decode())
    seqs.append("ACDEFGHIKL" * 500)          # 5000 residues, 10 distinct k-mers
    c = kr.Context(device=0, K=8)
    s = c.encode(seqs)
    b = oracle.batch(seqs, K=8)
    assert (s.sizes() == b.sizes()).all()
    assert (c.sim_matrix(s) == oracle.sim_matrix(b, ident_inf=True)).all()
    c.close()


def test_synthetic_families_vs_oracle(ctx, oracle):
    res, off, fam = synth.generate(0x5EED0003, 3000, 60)
    s = ctx.encode(res=res, off=off)
    b = oracle.batch(res=res, off=off, K=8)
    want = oracle.sim_matrix(b, ident_inf=True, threads=8)
    bt = ctx.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI)
    parts = [want[i, 256 * (i // 256):] for i in range(3000)]
    assert (bt == np.concatenate(parts)).all()
    # encode from a device-resident residue buffer gives the same set
    import torch
    d = torch.from_numpy(res).to(DEV)
    s2 = ctx.encode_device(d.data_ptr(), off)
    assert (s2.sizes() == s.sizes()).all()
    sim, _ = ctx.sim_pairs(s, np.arange(2999), s2, np.arange(1, 3000))
    assert (sim == want[np.arange(2999), np.arange(1, 3000)]).all()


def test_greedy_build_matches_oracle(oracle, rep50, rickettsia):
    res, off, _ = synth.generate(0x5EED0004, 6000, 150, xrun_per_10k=30)
    b = oracle.batch(res=res, off=off, K=8)
    for flags, path in [(f, pth) for f in (kr.IDENT_IS_INFINITY, 0) for pth in (0, kr.FORCE_JOIN)]:
        n, added, _ = oracle.greedy_build(b, 50, ident_inf=bool(flags))
        for batch in (0, 256, 1000):
            c = kr.Context(device=0, K=8, flags=flags | path)
            cands = c.encode(res=res, off=off)
            db = c.db()
            got = db.greedy_build(cands, 50, batch=batch)
            assert db.size == n and (got == added).all(), (flags, batch)
            c.close()
    # the reference's own database rebuilt from its own sequences keeps all 1480 (stingy-add invariant)
    c = kr.Context(device=0, K=8)
    reps = [s for (_, _, s) in rep50["reps"]]
    db = c.db()
    added = db.greedy_build(c.encode(reps), 50, batch=500)
    assert added.all() and db.size == 1480
    # duplicates: 100 PheS of which 34 distinct
    seqs = [p for (_, p) in rickettsia]
    n, want, _ = oracle.greedy_build(oracle.batch(seqs, K=8), 50)
    db2 = c.db()
    assert (db2.greedy_build(c.encode(seqs), 50, batch=16) == want).all() and db2.size == n
    # building on top of an existing database
    db3 = c.db()
    db3.append(reps[:700])
    got = db3.greedy_build(c.encode(reps), 50, batch=300)
    assert not got[:700].any() and got[700:].all() and db3.size == 1480
    c.close()


# ---- sorted-index path of findClosest (K4/K5) against the tile join and the oracle ---------------------------
def test_index_path_matches_join_and_oracle(rep50, gto_seeds, rickettsia, oracle, expected):
    reps = [s for (_, _, s) in rep50["reps"]]
    qs = [s for (_, _, s) in gto_seeds] + [p for (_, p) in rickettsia] + reps[:300] + ["", "ACD", "MKV"]
    orp, oq = oracle.batch(reps, K=8), oracle.batch(qs, K=8)
    for flags in (kr.IDENT_IS_INFINITY, 0):
        bs, bi = oracle.find_closest(orp, oq, ident_inf=bool(flags), threads=8)
        want_first = np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)
        for force in (kr.FORCE_INDEX, kr.FORCE_JOIN):
            c = kr.Context(device=0, K=8, flags=flags | force)
            db = c.db()
            db.append(reps[:900])
            db.append(reps[900:])
            sim, first, last = db.find_closest(c.encode(qs))
            assert (sim == bs).all(), (flags, force)
            assert (first.astype(np.int64) == want_first).all(), (flags, force)
            c.close()
    for i, e in enumerate(expected["config1"]):
        assert bs[i] == e["sim"]


def test_index_path_ties_ranks_and_short_sequences(oracle):
    res, off, _ = synth.generate(78, 1500, 15, dup_per_10k=2500)
    seqs = synth.to_strings(res, off)
    reps = seqs[:1000] + ["", "ACDE", "ACDE", "MKVLAAG"]          # representatives shorter than K
    qs = seqs[1000:] + ["ACDE", "", "MKVLAAG", "MKVLAAH", seqs[5]]
    rng = np.random.default_rng(4)
    order = rng.permutation(len(reps)).astype(np.int32)
    rank = np.empty(len(reps), dtype=np.uint32)
    rank[order] = np.arange(len(reps), dtype=np.uint32)
    orp, oq = oracle.batch(reps, K=8), oracle.batch(qs, K=8)
    for flags in (0, kr.IDENT_IS_INFINITY):
        bs, bi = oracle.find_closest(orp, oq, order=order, ident_inf=bool(flags))
        bs2, bi2 = oracle.find_closest(orp, oq, order=order, ident_inf=bool(flags), tie_last=True)
        for force in (kr.FORCE_INDEX, kr.FORCE_JOIN):
            c = kr.Context(device=0, K=8, flags=flags | force)
            db = c.db()
            db.append(reps)
            db.set_ranks(rank)
            sim, first, last = db.find_closest(c.encode(qs))
            assert (sim == bs).all() and (sim == bs2).all(), (flags, force)
            assert (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all(), (flags, force)
            assert (last.astype(np.int64) == np.where(bs2 > 0, bi2.astype(np.int64), kr.NO_REP)).all(), (flags, force)
            c.close()
    assert (first != last).any()


def test_index_path_large_database_and_batch(oracle):
    # 60,000 representatives (an index of 17.6 M entries) and 20,000 queries: index scan against the tile join, cell by cell
    res, off, _ = synth.generate(0x5EED0005, 80000 // SCALE, 1600 // SCALE)
    nrep = 60000 // SCALE
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
    db = c.db()
    db.append(res=res[: int(off[nrep])], off=off[: nrep + 1])
    q = c.encode(res=res[int(off[nrep]):], off=off[nrep:] - off[nrep])
    sim, first, last = db.find_closest(q)
    assert c.stats()["scan_launches"] >= 1 and c.stats()["join_launches"] == 0
    cj = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | kr.FORCE_JOIN)
    dbj = cj.db()
    dbj.append(res=res[: int(off[nrep])], off=off[: nrep + 1])
    simj, firstj, lastj = dbj.find_closest(cj.encode(res=res[int(off[nrep]):], off=off[nrep:] - off[nrep]))
    assert (sim == simj).all() and (first == firstj).all() and (last == lastj).all()
    # and a sample of them against the oracle
    pick = np.random.default_rng(8).choice(q.count, min(150, q.count), replace=False)
    orp = oracle.batch(res=res[: int(off[nrep])], off=off[: nrep + 1], K=8)
    qsel = [res[int(off[nrep + i]): int(off[nrep + i + 1])].tobytes().decode() for i in pick]
    bs, bi = oracle.find_closest(orp, oracle.batch(qsel, K=8), threads=16)
    assert (sim[pick] == bs).all()
    assert (first[pick].astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all()
    c.close()
    cj.close()


def test_greedy_build_properties_at_scale():
    """Sizes the oracle cannot reach: the domain's own invariants.  (a) stingy-add: no two representatives reach
    the threshold; (b) every rejected candidate is represented; (c) the result does not depend on the batch size;
    (d) all-vs-all at this size agrees with the independent per-pair kernel and is symmetric."""
    import torch
    n, thr = 120000 // SCALE, 50
    res, off, _ = synth.generate(0x5EED0004, n, n // 50, xrun_per_10k=10)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
    cands = c.encode(res=res, off=off)
    db = c.db()
    added = db.greedy_build(cands, thr)
    db2 = c.db()
    assert (db2.greedy_build(cands, thr, batch=1500) == added).all()              # (c)
    rep_idx = np.nonzero(added)[0].astype(np.uint32)
    R = len(rep_idx)
    assert db.size == R and 5000 // SCALE < R < n
    reps = cands.gather(rep_idx)
    ne = kr.matrix_elems(R, R, kr.LAYOUT_FULL)
    out = torch.empty(ne, dtype=torch.int16, device=DEV)
    c.sim_matrix_device(reps, None, out.data_ptr(), elem=kr.ELEM_U16, layout=kr.LAYOUT_FULL)
    c.synchronize()
    m = out.view(R, R)
    assert bool((m == m.T).all())                                                  # (d) symmetry
    assert bool((torch.diagonal(m) == kr.INFINITY).all())
    offdiag = m.clone()
    offdiag.fill_diagonal_(0)
    assert int(offdiag.max()) < thr                                                # (a)
    rng = np.random.default_rng(2)
    ia, ib = rng.integers(0, R, 20000 // SCALE), rng.integers(0, R, 20000 // SCALE)
    sim, _ = c.sim_pairs(reps, ia, reps, ib)
    got = m[torch.from_numpy(ia).to(DEV), torch.from_numpy(ib).to(DEV)].cpu().numpy().astype(np.int32)
    assert (sim == got).all()                                                      # (d) two kernels agree
    del m, offdiag, out
    rej = np.nonzero(added == 0)[0].astype(np.uint32)
    best, first, _ = db.find_closest(cands.gather(rej))
    assert (best >= thr).all()                                                     # (b)
    c.close()


def test_allvsall_full_size_index_equals_join():
    """BASELINE configs[2] at full size (100,000 sequences, 5e9 pairs): the index scan's dense result equals the tile
    join's, cell for cell, on the device; the sparse cells are exactly the non-zero cells of the strict triangle; and
    row stripes of the index path reassemble to the whole (the multi-GPU sharding unit)."""
    import torch
    n = 100000 // SCALE
    res, off, _ = synth.generate(0x5EED0003, n, n // 50)
    d_res = torch.from_numpy(res).to(DEV)
    ne = kr.matrix_elems(n, n, kr.LAYOUT_BLOCKTRI)
    outs = {}
    for name, flag in (("join", kr.FORCE_JOIN), ("index", kr.FORCE_INDEX), ("auto", 0)):
        c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | flag)
        s = c.encode_device(d_res.data_ptr(), off)
        out = torch.full((ne + 8,), -1, dtype=torch.int16, device=DEV)
        dsync()
        c.sim_matrix_device(s, None, out.data_ptr(), elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI)
        c.synchronize()
        st = c.stats()
        if name == "auto" and SCALE == 1:
            assert st["scan_launches"] >= 1 and st["join_launches"] == 0      # a family-structured batch belongs to the index
        if name == "index":
            # stripes (4 ranks' worth) reassemble to the whole
            rb = kr.matrix_stripes(n, n, kr.LAYOUT_BLOCKTRI, 4)
            pos = 0
            for k in range(4):
                cnt = kr.matrix_elems(n, n, kr.LAYOUT_BLOCKTRI, int(rb[k]), int(rb[k + 1]))
                part = torch.full((cnt + 8,), -1, dtype=torch.int16, device=DEV)
                dsync()
                c.sim_matrix_device(s, None, part.data_ptr(), elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI, row_begin=int(rb[k]), row_end=int(rb[k + 1]))
                c.synchronize()
                assert torch.equal(part[:cnt], out[pos:pos + cnt]), k
                pos += cnt
                del part
            assert pos == ne
            cap = (1 << 23) // SCALE
            cells = torch.zeros((cap, 4), dtype=torch.int32, device=DEV)
            dsync()
            cnt = c.sim_matrix_sparse_device(s, None, cells.data_ptr(), cap, min_sim=1, layout=kr.LAYOUT_TRIANGLE)
            c.synchronize()
            outs["cells"] = cells[:cnt].cpu().numpy()
        outs[name] = out[:ne]
        s.free()
        c.close()
    assert torch.equal(outs["join"], outs["index"]) and torch.equal(outs["join"], outs["auto"])
    # sparse cells against the dense result: same count of non-zero strict-triangle cells, and every cell's value
    cells = outs["cells"]
    i, j, v = cells[:, 0].astype(np.int64), cells[:, 1].astype(np.int64), cells[:, 2]
    assert (j > i).all() and len(np.unique(i * n + j)) == len(cells)
    I = i // 256
    idx = 256 * (I * n - 128 * I * (I - 1)) + (i - 256 * I) * (n - 256 * I) + (j - 256 * I)
    dense = outs["join"]
    assert (dense[torch.from_numpy(idx).to(DEV)].cpu().numpy().astype(np.int32) == v).all()
    diag_i = np.arange(n, dtype=np.int64); DI = diag_i // 256
    didx = 256 * (DI * n - 128 * DI * (DI - 1)) + (diag_i - 256 * DI) * (n - 256 * DI) + (diag_i - 256 * DI)
    nz_total = int((dense != 0).sum().item())
    # non-zero cells of the block triangle = the diagonal (INFINITY) + the strict triangle's non-zero cells + the
    # below-diagonal cells inside the 256 x 256 diagonal blocks (mirror images of cells with j > i in the same block)
    same_block = int(((i // 256) == (j // 256)).sum())
    assert bool((dense[torch.from_numpy(didx).to(DEV)] == kr.INFINITY).all())
    assert nz_total == n + len(cells) + same_block


def test_greedy_build_sharded_driver_on_one_gpu(oracle):
    """sharding.greedy_build_gpu (the N>1 build: scoring split by rank, identical in-order sweep on every replica)
    with a one-rank gloo group: every GPU call it makes, against kr_greedy_build and the oracle"""
    import torch.distributed as dist
    from kmers.reps_b200 import sharding, synth
    res, off, _ = synth.generate(0x5EED0007, 3000, 40)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
    cands = c.encode(res=res, off=off)
    own = not dist.is_initialized()
    if own:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29671", rank=0, world_size=1)
    try:
        db = c.db()
        added = sharding.greedy_build_gpu(c, db, cands, 50, dist, batch=512)
    finally:
        if own:
            dist.destroy_process_group()
    db2 = c.db()
    assert (added == db2.greedy_build(cands, 50)).all() and db.size == db2.size == int(added.sum())
    _, want, _ = oracle.greedy_build(oracle.batch(res=res, off=off, K=8), 50, ident_inf=True)
    assert (added == want).all()
    c.close()


def _clone(bl):
    c = kr.BlockLists()
    for name, _ in kr.BlockLists._fields_:
        setattr(c, name, getattr(bl, name))
    return c


def _export_shards(c, d_res, off, nshards):
    """encode the batch in `nshards` slices cut at block boundaries and export each slice's block lists (one GPU, no NCCL)"""
    from kmers.reps_b200 import sharding
    cuts = sharding.block_shards(len(off) - 1, nshards)
    parts, shards = [], []
    for k in range(nshards):
        lo, hi = cuts[k], cuts[k + 1]
        p = c.encode_device(d_res.data_ptr() + int(off[lo]), off[lo: hi + 1] - off[lo])
        bl = c.export_blocklists(p)
        assert bl.entries <= c.blocklist_capacity(off[lo: hi + 1])
        parts.append(p)
        shards.append(bl)
    return parts, shards


def test_shard_import_single_process(ctx, oracle):
    """the N-GPU block-list exchange without NCCL: 2, 4 and 8 artificial shards (ragged last block) are imported on ONE
    GPU and joined; every cell equals the oracle's and the plain encode's.  This is the path whose misaligned shard
    starts faulted in round 1: it must be exercised (under compute-sanitizer too) before any multi-GPU run."""
    import torch
    res, off, _ = synth.generate(0x5EED0011, 3000 + 37, 60)
    seqs = synth.to_strings(res, off)
    want = oracle.sim_matrix(oracle.batch(seqs, K=8), ident_inf=True, threads=8)
    d_res = torch.from_numpy(res).to(DEV)
    plain = ctx.sim_matrix(ctx.encode(res=res, off=off), elem=kr.ELEM_U16, layout=kr.LAYOUT_FULL)
    assert (plain == want).all()
    for nshards in (2, 4, 8):
        parts, shards = _export_shards(ctx, d_res, off, nshards)
        full = ctx.import_blocklists(shards, d_res.data_ptr(), off)
        got = ctx.sim_matrix(full, elem=kr.ELEM_U16, layout=kr.LAYOUT_FULL)
        assert (got == want).all(), nshards
        bt = ctx.sim_matrix(full, elem=kr.ELEM_I32, layout=kr.LAYOUT_BLOCKTRI, row_begin=1024, row_end=len(seqs))
        assert (bt == np.concatenate([want[i, 256 * (i // 256):] for i in range(1024, len(seqs))])).all(), nshards
        full.free()
        for p in parts:
            p.free()


def test_shard_import_rejects_bad_layout(ctx):
    """a shard whose arrays break the 16-byte rules of the TMA copies is refused with KR_ERR_ARG; nothing is launched on it"""
    import torch
    res, off, _ = synth.generate(0x5EED0012, 600, 12)
    d_res = torch.from_numpy(res).to(DEV)
    parts, shards = _export_shards(ctx, d_res, off, 2)
    bad = [_clone(s) for s in shards]
    bad[1].hi = shards[1].hi + 4                                  # misaligned array start
    with pytest.raises(kr.KrError) as e:
        ctx.import_blocklists(bad, d_res.data_ptr(), off)
    assert e.value.code == -1 and "aligned" in str(e.value)
    # misaligned block start inside an aligned array: caught by the device-side layout check
    from kmers.reps_b200 import sharding
    nb = shards[1].nblk + 1
    base = sharding._view(shards[1].blk_base, nb * 8, d_res.device).view(torch.int64).clone()
    base += 4
    if not EMUL:
        torch.cuda.synchronize()
    bad = [_clone(s) for s in shards]
    bad[1].blk_base = base.data_ptr()
    with pytest.raises(kr.KrError) as e:
        ctx.import_blocklists(bad, d_res.data_ptr(), off)
    assert e.value.code == -1 and "layout" in str(e.value)
    # the context is still healthy afterwards
    ok = ctx.import_blocklists(shards, d_res.data_ptr(), off)
    assert ctx.sim_matrix(ok, elem=kr.ELEM_U16).shape == (600, 600)
    for p in parts:
        p.free()


def _gpu_list():
    """every GPU of the box -- or, on a one-GPU box, that GPU listed twice (two contexts, two worker threads on it)"""
    if EMUL:
        return [0, 0]
    import torch
    n = torch.cuda.device_count()
    return list(range(n)) if n > 1 else [0, 0]


@pytest.mark.parametrize("path", list(PATHS))
def test_multi_handle_equals_single_gpu_and_oracle(path, rep50, gto_seeds, oracle):
    """kr_multi (include/kmersreps.h, last section): ONE handle over several contexts with internal sharding gives what the
    single-GPU calls and the oracle give: rep50.db all-vs-all in every layout, sparse cells, rectangular, findClosest of the
    GTO seeds (BASELINE configs[0]) with HashMap ranks, and the greedy rebuild of rep50.db (keeps all 1480)."""
    seqs = [r[2] for r in rep50["reps"]]
    n = len(seqs)
    b = oracle.batch(seqs, K=8)
    want = oracle.sim_matrix(b, ident_inf=True, threads=8)
    m = kr.MultiContext(_gpu_list(), K=8, flags=kr.IDENT_IS_INFINITY | PATHS[path])
    assert (m.sim_matrix(seqs, elem=kr.ELEM_I32, layout=kr.LAYOUT_FULL) == want).all()
    assert (m.sim_matrix(seqs, elem=kr.ELEM_U16, layout=kr.LAYOUT_TRIANGLE) == tri_from_full(want)).all()
    from kmers.reps_b200 import sharding
    assert (m.sim_matrix(seqs, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI) == sharding.blocktri_from_full(want)).all()
    got = m.sim_matrix_sparse(seqs, min_sim=40, layout=kr.LAYOUT_TRIANGLE)
    ii, jj = np.nonzero(np.triu(want, 1) >= 40)
    assert {(int(r), int(c), int(v)) for r, c, v, _ in got} == {(int(i), int(j), int(want[i, j])) for i, j in zip(ii, jj)}
    qs = [s for (_, _, s) in gto_seeds]
    wr = oracle.sim_matrix(oracle.batch(qs, K=8), b, ident_inf=True, threads=8)
    assert (m.sim_matrix(qs, b_seqs=seqs, elem=kr.ELEM_I32) == wr).all()
    db = m.db()
    db.append(seqs)
    bs, bi = oracle.find_closest(b, oracle.batch(qs, K=8), ident_inf=True)
    sim, first, _ = db.find_closest(qs)
    assert (sim == bs).all() and (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all()
    single = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | PATHS[path])
    sdb = single.db(); sdb.append(seqs)
    s1, f1, l1 = sdb.find_closest(single.encode(seqs[:300]))
    s2, f2, l2 = db.find_closest(seqs[:300])
    assert (s1 == s2).all() and (f1 == f2).all() and (l1 == l2).all()
    single.close()
    db.free()
    db = m.db()
    added = db.greedy_build(50, seqs, batch=400)
    assert added.all() and db.size == n          # rep50.db is a threshold-50 database: rebuilding it keeps every representative
    db.free()
    assert m.stats()["kernel_launches"] > 0
    m.close()


def test_sets_from_key_lists(ctx, rep50):
    """kr_seqset_from_keys (SequenceKmers.retainAll / removeAll, TargetProcessor.java:166,184): the merged key lists become a
    set again; sizes, decoded k-mers, matrix and findClosest equal Python set arithmetic on substrings"""
    seqs = [s for (_, _, s) in rep50["reps"]][:400]
    sets = [ko.py_kmers(s, 8) for s in seqs]
    S = ctx.encode(seqs)
    keys = [S.keys(i) for i in range(len(seqs))]
    pool = set(int(k) for i in (0, 1, 2, 3, 4, 5) for k in keys[i])
    pool_str = set().union(*[sets[i] for i in range(6)])
    kept = [np.array([k for k in ks if int(k) in pool], dtype=np.uint64) for ks in keys]
    gone = [np.array([k for k in ks if int(k) not in pool], dtype=np.uint64) for ks in keys]
    for lists, want_sets in ((kept, [a & pool_str for a in sets]), (gone, [a - pool_str for a in sets])):
        T = ctx.from_keys(lists, seqs)
        assert list(T.sizes()) == [len(w) for w in want_sets]
        assert sorted(T.kmers(7)) == sorted(want_sets[7])
        raw = np.array([[len(a & b) for b in sets] for a in want_sets], dtype=np.int32)
        inf = np.array([[seqs[i] == seqs[j] and want_sets[i] == sets[j] for j in range(len(seqs))] for i in range(len(seqs))])
        assert (ctx.sim_matrix(T, S, elem=kr.ELEM_I32) == np.where(inf, kr.INFINITY, raw)).all()
        db = ctx.db()
        db.append_set(T, np.arange(len(seqs), dtype=np.uint32))
        sim, first, _ = db.find_closest(S)
        col = np.where(inf, kr.INFINITY, raw).T
        assert (sim == col.max(axis=1)).all()
        assert all(col[q, first[q]] == sim[q] and (col[q, :first[q]] < sim[q]).all() for q in range(len(seqs)) if sim[q] > 0)
        db.free()
    with pytest.raises(kr.KrError):
        ctx.from_keys([np.array([7, 7], np.uint64)])
