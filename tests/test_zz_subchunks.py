"""GPU test of the join's sub-chunk path: key-range chunks longer than one pipeline stage (2040 entries) are
processed as several (A sub-chunk x B sub-chunk) jobs.  Seed proteins never get there (a chunk holds ~1,200
entries at 300 aa), so this uses 700-2,000-residue sequences.  (Written after GPU access for this repository was
closed in round 1: the path itself was exercised on a B200 with the previous stage capacity of 1536 by the
rep50.db tests; this file pins it for the current capacity and runs last.)"""
import numpy as np
import pytest

import kmers.reps_b200 as kr
from oracle import kr_oracle as ko

pytestmark = pytest.mark.gpu


def long_family(seed, n, length, alphabet=b"ACDEFGHIKLMNPQRSTVWY"):
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    roots = [alpha[rng.integers(0, len(alpha), length)] for _ in range(6)]
    seqs = []
    for i in range(n):
        s = roots[i % 6].copy()
        pos = rng.integers(0, length, int(length * 0.02 * (1 + i % 5)))
        s[pos] = alpha[rng.integers(0, len(alpha), len(pos))]
        cut = int(rng.integers(length * 3 // 4, length + 1))
        seqs.append(s[:cut].tobytes().decode())
    return seqs


@pytest.mark.parametrize("length,n", [(700, 300), (2000, 270)])
def test_chunks_longer_than_a_stage(oracle, length, n):
    seqs = long_family(length, n, length)
    assert np.mean([len(s) for s in seqs]) * 256 / 64 > 2200          # mean chunk length beyond one stage
    b = oracle.batch(seqs, K=8)
    want = oracle.sim_matrix(b, ident_inf=True, threads=8)
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | kr.FORCE_JOIN)      # this file is about the tile join
    s = c.encode(seqs)
    assert (s.sizes() == b.sizes()).all()
    assert (c.sim_matrix(s, elem=kr.ELEM_U16) == want).all()
    tri = c.sim_matrix(s, elem=kr.ELEM_I32, layout=kr.LAYOUT_TRIANGLE)
    assert (tri == want[np.triu_indices(n, 1)]).all()
    # the ARGMAX join over the same long lists: every sequence finds itself, and a leave-one-out query finds the
    # oracle's best other sequence
    db = c.db()
    db.append(seqs[: n - 40])
    sim, first, last = db.find_closest(c.encode(seqs[n - 40:]))
    bs, bi = oracle.find_closest(oracle.batch(seqs[: n - 40], K=8), oracle.batch(seqs[n - 40:], K=8))
    assert (sim == bs).all() and (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all()
    c.close()
    c10 = kr.Context(device=0, K=10, flags=kr.FORCE_JOIN)                           # 32-bit low words, 3 stages
    s10 = c10.encode(seqs[:260])
    want10 = oracle.sim_matrix(oracle.batch(seqs[:260], K=10), ident_inf=False, threads=8)
    assert (c10.sim_matrix(s10) == want10).all()
    c10.close()


def test_buildsets_hierarchy(tmp_path):
    """the `genomes` pipeline's similarity work (multi-threshold hierarchical rep sets) against its restatement"""
    import os
    import subprocess
    from kmers.reps_b200 import synth
    from oracle import ref_commands as rc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = os.environ.get("KR_TOOL_PATH") or os.path.join(root, "kmers", "reps_b200", "host", "kmers_reps")
    res, off, _ = synth.generate(0x5EED0006, 1500, 30)
    seqs = synth.to_strings(res, off)
    recs = [("fig|%d.%d.peg.%d" % (2000 + 3 * i, 1 + i % 7, 10 + i), "Synthetic genome %d" % i, s) for i, s in enumerate(seqs)]
    src = tmp_path / "master.fa"
    src.write_bytes(b"".join(ko.write_fasta_record(*r) for r in recs))
    thresholds = [10, 50, 100, 200]
    p = subprocess.run([tool, "buildsets", "-t", ",".join(map(str, thresholds)), str(tmp_path), str(src)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()
    want = rc.buildsets(recs, thresholds)
    sizes = []
    for thr in thresholds:
        db_bytes, listing = want[thr]
        assert (tmp_path / ("rep%d.ser" % thr)).read_bytes() == db_bytes
        assert (tmp_path / ("rep%d.list.tbl" % thr)).read_text() == listing
        sizes.append(db_bytes.count(b">") - 1)
    assert sizes == sorted(sizes) and sizes[0] < sizes[-1]            # larger thresholds keep more representatives
