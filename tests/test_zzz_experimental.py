"""GPU tests of EXPERIMENTAL entry points.  They exist only in a library built with -DKR_EXPERIMENTAL
(kmers/reps_b200/libkmersreps_cuda_exp.so, never the default library) and were written while GPU access was closed,
so they are skipped unless KR_LIB_PATH points at such a build:

    KR_LIB_PATH=$PWD/kmers/reps_b200/libkmersreps_cuda_exp.so python -m pytest tests -m gpu -q
"""
import numpy as np
import pytest

import kmers.reps_b200 as kr
from kmers.reps_b200 import binding, synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not binding.has_experimental(), reason="default library: no experimental entry points")]


def cells_of(dense, min_sim, triangle):
    n, m = dense.shape
    out = set()
    ii, jj = np.nonzero(dense >= min_sim)
    for i, j in zip(ii, jj):
        if not triangle or j > i:
            out.add((int(i), int(j), int(dense[i, j])))
    return out


@pytest.mark.parametrize("via", [0, binding.EXP_PAIRS, binding.EXP_PAIRS | binding.EXP_BUCKETS])
@pytest.mark.parametrize("n,nfam", [(700, 12), (1500, 300)])
def test_sparse_matrix_equals_dense(oracle, n, nfam, via):
    """kr_sim_matrix_sparse through the join's epilogue filter (via = 0) and through posting lists + pair table"""
    res, off, _ = synth.generate(0x5EED0008, n, nfam)
    seqs = synth.to_strings(res, off)
    seqs[5] = seqs[9] = "MKV"                                            # equal sequences without any k-mer
    seqs[11] = ""
    res, off = kr.flat(seqs)
    want = oracle.sim_matrix(oracle.batch(res=res, off=off, K=8), ident_inf=True, threads=4)
    assert want[5, 9] == 9999
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | via)
    s = c.encode(res=res, off=off)
    for min_sim in (1, 50, 9999):
        got = c.sim_matrix_sparse(s, min_sim=min_sim, layout=kr.LAYOUT_TRIANGLE)
        cells = {(int(r), int(cc), int(v)) for r, cc, v, _ in got}
        assert len(cells) == len(got)                                   # no cell twice
        assert cells == cells_of(want, min_sim, True)
    # too little room: the count comes back, the binding retries with enough
    got = c.sim_matrix_sparse(s, min_sim=1, layout=kr.LAYOUT_TRIANGLE, cap=16)
    assert {(int(r), int(cc), int(v)) for r, cc, v, _ in got} == cells_of(want, 1, True)
    # rectangular, FULL layout, a row stripe
    q = c.encode(res=res[: int(off[300])], off=off[:301])
    got = c.sim_matrix_sparse(q, s, min_sim=20, layout=kr.LAYOUT_FULL, row_begin=256, row_end=300)
    assert {(int(r), int(cc), int(v)) for r, cc, v, _ in got} == {t for t in cells_of(want[:300], 20, False) if t[0] >= 256}
    c.close()


def blocktri(full):
    n = full.shape[0]
    return np.concatenate([full[i, 256 * (i // 256):] for i in range(n)])


@pytest.mark.parametrize("flags", [binding.EXP_PAIRS, binding.EXP_PAIRS_AUTO, binding.EXP_PAIRS | binding.EXP_BUCKETS, binding.EXP_PAIRS_AUTO | binding.EXP_BUCKETS])
def test_posting_list_all_vs_all_equals_oracle(oracle, vapb, rep50, flags):
    """the posting-list path (kr_pairs.cuh) must give the join's matrix bit for bit: every layout, element type,
    row stripe, with and without the equal-sequence rule; VapbProt.fa holds 11 identical-sequence pairs"""
    sets = [[s for (_, _, s) in vapb], [s for (_, _, s) in rep50["reps"][:700]]]
    res, off, _ = synth.generate(0x5EED0009, 1300, 25)
    sets.append(synth.to_strings(res, off))
    for seqs in sets:
        n = len(seqs)
        ob = oracle.batch(seqs, K=8)
        for ident in (kr.IDENT_IS_INFINITY, 0):
            want = oracle.sim_matrix(ob, ident_inf=bool(ident), threads=4)
            c = kr.Context(device=0, K=8, flags=ident | flags)
            s = c.encode(seqs)
            assert (c.sim_matrix(s, elem=kr.ELEM_I32, layout=kr.LAYOUT_FULL) == want).all()
            assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_FULL) == want).all()
            assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_TRIANGLE) == want[np.triu_indices(n, 1)]).all()
            assert (c.sim_matrix(s, elem=kr.ELEM_I32, layout=kr.LAYOUT_BLOCKTRI) == blocktri(want)).all()
            if n > 256:
                lo = c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI, row_begin=0, row_end=256)
                hi = c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI, row_begin=256, row_end=n)
                assert (np.concatenate([lo, hi]) == blocktri(want)).all()
            c.close()


def test_posting_list_path_in_greedy_build_and_short_sequences(oracle):
    res, off, _ = synth.generate(0x5EED000A, 2500, 40)
    seqs = synth.to_strings(res, off) + ["MKV", "", "MKVLAAGIVG", "MKVLAAGIVG", "MKV"]
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | binding.EXP_PAIRS)
    s = c.encode(seqs)
    ob = oracle.batch(seqs, K=8)
    assert (c.sim_matrix(s, elem=kr.ELEM_U16) == oracle.sim_matrix(ob, ident_inf=True, threads=4)).all()
    db = c.db()
    added = db.greedy_build(s, 50, batch=600)
    _, want, _ = oracle.greedy_build(ob, 50, ident_inf=True)
    assert (added == want).all()
    c.close()


def test_warp_per_sequence_encode_equals_cta_encode(oracle, rep50, vapb):
    """EXP_ENCODE (K1 with one warp per sequence) against the default K1 and the oracle: real and synthetic sets, very
    long sequences (in-memory path), all K"""
    rng = np.random.default_rng(4)
    alpha = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX*-", dtype=np.uint8)
    longs = [alpha[rng.integers(0, len(alpha), n)].tobytes().decode() for n in (519, 520, 1031, 1032, 1040, 5000, 30000)]
    res, off, _ = synth.generate(0x5EED000B, 5000, 80)
    for seqs in ([s for (_, _, s) in rep50["reps"]], [s for (_, _, s) in vapb] + longs + ["", "MKV", "A" * 400], synth.to_strings(res, off)):
        for K in (8, 5, 11):
            c = kr.Context(device=0, K=K, flags=kr.IDENT_IS_INFINITY | binding.EXP_ENCODE)
            ref = kr.Context(device=0, K=K, flags=kr.IDENT_IS_INFINITY)
            s, s0 = c.encode(seqs), ref.encode(seqs)
            assert (s.sizes() == s0.sizes()).all() and (s.sizes() == oracle.batch(seqs, K=K).sizes()).all()
            for i in rng.choice(len(seqs), min(len(seqs), 200), replace=False):
                assert (s.keys(int(i)) == s0.keys(int(i))).all()
            if len(seqs) <= 1500 and K == 8:
                assert (c.sim_matrix(s, elem=kr.ELEM_U16) == ref.sim_matrix(s0, elem=kr.ELEM_U16)).all()
            c.close(); ref.close()


def test_warp_per_query_index_equals_cta_index(oracle, rep50, gto_seeds, rickettsia):
    """EXP_IDXQ (K5 with one warp per query): dense per-warp scores on rep50.db, the per-warp hash on a 60,000-rep
    synthetic database (two slabs), against the default index kernel and the oracle"""
    reps = [s for (_, _, s) in rep50["reps"]]
    qs = [s for (_, _, s) in gto_seeds] + [p for (_, p) in rickettsia] + reps[:50] + ["MKV", ""]
    rng = np.random.default_rng(8)
    rank = rng.permutation(len(reps)).astype(np.uint32)
    out = []
    for flags in (binding.EXP_IDXQ, 0):
        c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | kr.FORCE_INDEX | flags)
        db = c.db(); db.append(reps); db.set_ranks(rank)
        out.append(db.find_closest(c.encode(qs)))
        c.close()
    for a, b in zip(*out):
        assert (a == b).all()
    bs, bi = oracle.find_closest(oracle.batch(reps, K=8), oracle.batch(qs, K=8), order=np.argsort(rank).astype(np.int32), ident_inf=True, threads=8)
    assert (out[0][0] == bs).all() and (out[0][1].astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all()
    res, off, _ = synth.generate(0x5EED000C, 90000, 1500)
    out = []
    for flags in (binding.EXP_IDXQ, 0):
        c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | kr.FORCE_INDEX | flags)
        db = c.db(); db.append(res=res[: int(off[60000])], off=off[:60001])
        out.append(db.find_closest(c.encode(res=res[int(off[60000]):], off=off[60000:] - off[60000])))
        c.close()
    for a, b in zip(*out):
        assert (a == b).all()


def test_greedy_build_through_the_growing_index(oracle, rep50, rickettsia):
    """EXP_GIDX: kr_greedy_build scoring candidates through the appendable index, against the default build and the
    oracle: synthetic families at two thresholds and batch sizes, rep50.db's own sequences (all kept), duplicates"""
    res, off, _ = synth.generate(0x5EED000D, 30000, 500, xrun_per_10k=10)
    seqs = synth.to_strings(res, off)
    for thr, batch in ((50, 0), (100, 1000)):
        outs = []
        for flags in (binding.EXP_GIDX, 0):
            c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | flags)
            db = c.db()
            a0 = db.greedy_build(c.encode(seqs[:5000]), thr, batch=batch)
            a1 = db.greedy_build(c.encode(seqs[5000:]), thr, batch=batch)
            outs.append(np.concatenate([a0, a1]))
            c.close()
        assert (outs[0] == outs[1]).all()
    _, want, _ = oracle.greedy_build(oracle.batch(seqs[:6000], K=8), 50, ident_inf=True)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | binding.EXP_GIDX)
    assert (c.db().greedy_build(c.encode(seqs[:6000]), 50) == want).all()
    reps = [s for (_, _, s) in rep50["reps"]] + [p for (_, p) in rickettsia] + ["MKV", "MKV", ""]
    _, want, _ = oracle.greedy_build(oracle.batch(reps, K=8), 50, ident_inf=True)
    assert (c.db().greedy_build(c.encode(reps), 50, batch=300) == want).all() and want[:1480].all()
    c.close()


def test_dna_kmers_both_strands():
    """ALPHABET_DNA contexts (DnaKmers: K-mers of the lower-cased sequence and of its reverse complement) against the
    Python restatement: SSU-sized sequences at K = 18 and 15, other strand / other case / 'n' / RNA 'u', the join and the
    posting-list path, and the alphabet error"""
    from oracle import kr_oracle as ko
    rng = np.random.default_rng(6)
    alpha = np.frombuffer(b"acgt", dtype=np.uint8)
    roots = [alpha[rng.integers(0, 4, 1500)] for _ in range(5)]
    seqs = []
    for i in range(300):
        s = roots[i % 5].copy()
        pos = rng.integers(0, 1500, 5 + 25 * (i % 6))
        s[pos] = alpha[rng.integers(0, 4, len(pos))]
        t = s[: int(rng.integers(1100, 1501))].tobytes().decode()
        t = ko.py_dna_revcomp(t) if i % 7 == 3 else t
        t = t.upper() if i % 5 == 1 else t
        t = t[:300] + "nn" + t[302:] if i % 11 == 4 else t
        seqs.append(t)
    seqs += ["acgtacgtacgtacgtacgtacgt", "", "ACGU" * 10, seqs[3].upper()]
    for K in (18, 15):
        sets = [ko.py_dna_kmers(q, K) for q in seqs]
        want = np.array([[9999 if seqs[a].lower() == seqs[b].lower() else len(sets[a] & sets[b]) for b in range(len(seqs))] for a in range(len(seqs))])
        for extra in (0, binding.EXP_PAIRS | binding.EXP_BUCKETS):
            c = kr.Context(device=0, K=K, flags=kr.IDENT_IS_INFINITY | binding.ALPHABET_DNA | extra)
            s = c.encode(seqs)
            assert (s.sizes() == [len(x) for x in sets]).all()
            assert (c.sim_matrix(s, elem=kr.ELEM_U16) == want).all()
            c.close()
    c = kr.Context(device=0, K=15, flags=binding.ALPHABET_DNA)
    with pytest.raises(kr.KrError):
        c.encode(["acgtacgtacgtacgtacgt", "acgtacgtrcgtacgtacgt"])
    c.close()


def test_seqcomp_command(tmp_path, rep50):
    """`kmers_reps seqcomp` (SeqCompProcessor.java:105-137): protein and SSU-rRNA k-mer distances of every genome pair,
    byte for byte against the restatement.  The host tool finds the library through LD_LIBRARY_PATH, so this needs the
    experimental build to be the one named libkmersreps_cuda.so there (tools/exp_gpu_check.sh arranges that)."""
    import os
    import subprocess
    from oracle import kr_oracle as ko
    from oracle import ref_commands as rc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = os.path.join(root, "kmers", "reps_b200", "host", "kmers_reps")
    rng = np.random.default_rng(12)
    alpha = np.frombuffer(b"acgt", dtype=np.uint8)
    roots = [alpha[rng.integers(0, 4, 900)] for _ in range(3)]
    rows = []
    for i, (fid, name, prot) in enumerate(rep50["reps"][:40]):
        d = roots[i % 3].copy()
        pos = rng.integers(0, 900, 3 + 20 * (i % 4))
        d[pos] = alpha[rng.integers(0, 4, len(pos))]
        dna = d.tobytes().decode()
        dna = ko.py_dna_revcomp(dna) if i % 5 == 2 else (dna.upper() if i % 5 == 1 else dna)
        rows.append((ko.genome_id_of(fid), name, prot, dna))
    rows.append((ko.genome_id_of(rep50["reps"][40][0]), "copy", rows[3][2], rows[3][3]))     # equal to genome 3 in both sequences
    tbl = tmp_path / "seqs.tbl"
    tbl.write_text("genome_id\tgenome_name\tseed_protein\tssu_rna\n" + "".join("%s\t%s\t%s\t%s\n" % r for r in rows))
    p = subprocess.run([tool, "seqcomp", "-k", "15", str(tbl)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()
    want = rc.seqcomp(rows, K=8, dna_k=15)
    assert p.stdout.decode() == want
    lines = want.split("\n")[1:-1]
    assert len(lines) == len(rows) * (len(rows) + 1) // 2 and any(ln.endswith("0.0000\t0.0000") for ln in lines)


def test_ssureps_command(tmp_path):
    """`kmers_reps ssureps` (SsuRepGenomeProcessor.java:104-149): closest SSU-rRNA representative by k-mer distance, byte for
    byte against the restatement (needs the experimental build as the host tools' libkmersreps_cuda.so)"""
    import json
    import os
    import subprocess
    from oracle import kr_oracle as ko
    from oracle import ref_commands as rc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = os.path.join(root, "kmers", "reps_b200", "host", "kmers_reps")
    rng = np.random.default_rng(14)
    alpha = np.frombuffer(b"acgt", dtype=np.uint8)
    roots = [alpha[rng.integers(0, 4, 1200)] for _ in range(4)]

    def variant(i, rate):
        d = roots[i % 4].copy()
        pos = rng.integers(0, 1200, int(1200 * rate))
        d[pos] = alpha[rng.integers(0, 4, len(pos))]
        return d[: int(rng.integers(900, 1201))].tobytes().decode()

    table = [("%d.1" % (500 + i), "rep %d" % i, variant(i, 0.01 * (1 + i % 3))) for i in range(12)]
    genomes = [("%d.2" % (7000 + 3 * i), "input %d" % i, "" if i == 4 else variant(i, 0.02 + 0.02 * (i % 4))) for i in range(15)]
    genomes[6] = (genomes[6][0], genomes[6][1], ko.py_dna_revcomp(table[2][2]).upper())     # the other strand of a representative: distance 0 only if equal
    genomes[9] = (genomes[9][0], genomes[9][1], "a" * 400)                                  # shares nothing: no row
    gdir = tmp_path / "gtos"
    gdir.mkdir()
    for (gid, name, ssu) in genomes:
        json.dump({"id": gid, "scientific_name": name, "domain": "Bacteria", "ssu_rrna": ssu, "features": []}, open(gdir / (gid + ".gto"), "w"))
    tbl = tmp_path / "ssu.tbl"
    tbl.write_text("genome_id\tgenome_name\tssu_rna\n" + "".join("%s\t%s\t%s\n" % t for t in table))
    p = subprocess.run([tool, "ssureps", "-v", str(tbl), str(gdir)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()
    want = rc.ssureps(table, sorted(genomes, key=lambda g: g[0]), K=18)
    assert p.stdout.decode() == want
    assert "15 genomes processed.  1 missing SSUs. 1 outliers." in p.stderr.decode()
