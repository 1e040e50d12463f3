#!/usr/bin/env python
"""Randomised comparison of the library (emulator build: KR_LIB_PATH=tests/emul/_build/libkmersreps_emul.so) with the
oracle: random batch shapes (empty / shorter-than-K / very long sequences, families, duplicates, block and chunk
boundaries), every K, both equal-sequence policies, all layouts and element types, row stripes, both findClosest paths,
the greedy build, and -- with --exp -- the experimental paths.  TEST INFRASTRUCTURE; CPU only.

    KR_LIB_PATH=$PWD/tests/emul/_build/libkmersreps_emul.so KR_EMUL_SMS=3 python tools/emul_fuzz.py --seconds 600 [--exp] [--seed N]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmers.reps_b200 as kr                      # noqa: E402
from kmers.reps_b200 import binding, sharding     # noqa: E402
from oracle import kr_oracle as ko                # noqa: E402

assert "emul" in kr.lib_path(), "point KR_LIB_PATH at the emulator build"
O = ko.get()
ALPHA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", dtype=np.uint8)


def batch(rng):
    n = int(rng.choice([1, 2, 5, 40, 255, 256, 257, 300, 513, 600], p=[.05, .05, .1, .2, .1, .1, .1, .15, .1, .05]))
    nfam = int(rng.integers(1, 8))
    small = rng.random() < 0.3                       # few distinct residues: heavy k-mer sharing, long posting lists
    alpha = ALPHA[:4] if small else ALPHA
    base = int(rng.choice([12, 40, 120, 320]))
    roots = [alpha[rng.integers(0, len(alpha), base + int(rng.integers(0, base)))] for _ in range(nfam)]
    seqs = []
    for i in range(n):
        r = rng.random()
        if r < 0.03:
            seqs.append("")
        elif r < 0.08:
            seqs.append(alpha[rng.integers(0, len(alpha), int(rng.integers(1, 12)))].tobytes().decode())
        elif r < 0.10 and n <= 300:
            seqs.append(alpha[rng.integers(0, len(alpha), int(rng.integers(1500, 3500)))].tobytes().decode())
        elif r < 0.16 and seqs:
            seqs.append(seqs[int(rng.integers(0, len(seqs)))])
        else:
            s = roots[int(rng.integers(0, nfam))].copy()
            k = int(len(s) * rng.choice([0.0, 0.02, 0.1, 0.3]))
            if k:
                s[rng.integers(0, len(s), k)] = alpha[rng.integers(0, len(alpha), k)]
            seqs.append(s[: int(rng.integers(max(1, len(s) // 2), len(s) + 1))].tobytes().decode())
    return seqs


def one(rng):
    seqs = batch(rng)
    n = len(seqs)
    K = int(rng.choice([1, 2, 3, 5, 8, 8, 8, 9, 10, 11]))
    ident = bool(rng.random() < 0.7)
    flags = kr.IDENT_IS_INFINITY if ident else 0
    flags |= int(rng.choice([0, kr.FORCE_JOIN, kr.FORCE_INDEX, kr.FORCE_INDEX]))      # the library's choice, or a pinned device path
    desc = "n=%d K=%d ident=%d flags=0x%x" % (n, K, ident, flags)
    ob = O.batch(seqs, K=K)
    want = O.sim_matrix(ob, ident_inf=ident, threads=4)
    c = kr.Context(device=0, K=K, flags=flags)
    s = c.encode(seqs)
    assert (s.sizes() == ob.sizes()).all(), desc
    layout = int(rng.choice([kr.LAYOUT_FULL, kr.LAYOUT_TRIANGLE, kr.LAYOUT_BLOCKTRI]))
    elem = int(rng.choice([kr.ELEM_I32, kr.ELEM_U16]))
    nblk = (n + 255) // 256
    b0 = int(rng.integers(0, nblk))
    b1 = int(rng.integers(b0 + 1, nblk + 1))
    r0, r1 = 256 * b0, min(n, 256 * b1)
    got = c.sim_matrix(s, elem=elem, layout=layout, row_begin=r0, row_end=r1)
    if layout == kr.LAYOUT_FULL:
        exp_cells = want[r0:r1]
    elif layout == kr.LAYOUT_TRIANGLE:
        exp_cells = np.concatenate([want[i, i + 1:] for i in range(r0, r1)]) if r1 > r0 else np.zeros(0)
    else:
        exp_cells = np.concatenate([want[i, 256 * (i // 256):] for i in range(r0, r1)])
    assert (np.asarray(got).ravel() == np.asarray(exp_cells).ravel()).all(), desc + " layout=%d elem=%d rows[%d,%d)" % (layout, elem, r0, r1)
    if rng.random() < 0.5 and n > 1:
        ms = int(rng.choice([1, 2, 5]))
        cells = c.sim_matrix_sparse(s, min_sim=ms, layout=kr.LAYOUT_TRIANGLE, cap=64, row_begin=r0, row_end=r1)
        got_cells = {(int(a), int(b), int(v)) for a, b, v, _ in cells}
        ii, jj = np.nonzero(np.triu(want, 1) >= ms)
        want_cells = {(int(i), int(j), int(want[i, j])) for i, j in zip(ii, jj) if r0 <= i < r1}
        assert len(got_cells) == len(cells) and got_cells == want_cells, desc + " sparse min_sim=%d rows[%d,%d)" % (ms, r0, r1)
    # findClosest: a random split into database and queries, random ranks, both device paths
    if n >= 2:
        nrep = int(rng.integers(1, n))
        rank = rng.permutation(nrep).astype(np.uint32)
        bs, bi = O.find_closest(O.batch(seqs[:nrep], K=K), O.batch(seqs[nrep:], K=K), order=np.argsort(rank).astype(np.int32), ident_inf=ident, threads=4)
        bs2, bl = O.find_closest(O.batch(seqs[:nrep], K=K), O.batch(seqs[nrep:], K=K), order=np.argsort(rank).astype(np.int32), ident_inf=ident, tie_last=True, threads=4)
        path = int(rng.choice([kr.FORCE_JOIN, kr.FORCE_INDEX]))
        c2 = kr.Context(device=0, K=K, flags=(flags & ~(kr.FORCE_JOIN | kr.FORCE_INDEX)) | path)
        db = c2.db()
        cut = int(rng.integers(0, nrep + 1))
        db.append(seqs[:cut]) if cut else None
        db.append(seqs[cut:nrep]) if cut < nrep else None
        db.set_ranks(rank)
        sim, first, last = db.find_closest(c2.encode(seqs[nrep:]))
        assert (sim == bs).all(), desc + " findClosest path=%d" % path
        assert (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all(), desc + " first, path=%d" % path
        assert (last.astype(np.int64) == np.where(bs2 > 0, bl.astype(np.int64), kr.NO_REP)).all(), desc + " last, path=%d" % path
        c2.close()
    # greedy build
    thr = int(rng.choice([1, 5, 20, 50]))
    _, w, _ = O.greedy_build(ob, thr, ident_inf=ident)
    added = c.db().greedy_build(s, thr, batch=int(rng.choice([0, 7, 64, 300])))
    assert (added == w).all(), desc + " greedy thr=%d" % thr
    c.close()
    # the same through one kr_multi handle over 1..3 contexts (internal sharding)
    if rng.random() < 0.3:
        nd = int(rng.integers(1, 4))
        m = kr.MultiContext([0] * nd, K=K, flags=flags)
        got = m.sim_matrix(seqs, elem=elem, layout=layout)
        if layout == kr.LAYOUT_FULL:
            exp_all = want
        elif layout == kr.LAYOUT_TRIANGLE:
            exp_all = np.concatenate([want[i, i + 1:] for i in range(n)]) if n > 1 else np.zeros(0)
        else:
            exp_all = np.concatenate([want[i, 256 * (i // 256):] for i in range(n)])
        assert (np.asarray(got).ravel() == np.asarray(exp_all).ravel()).all(), desc + " multi x%d layout=%d" % (nd, layout)
        if n > 1:
            cells = m.sim_matrix_sparse(seqs, min_sim=2, layout=kr.LAYOUT_TRIANGLE, cap=max(1, int(rng.choice([4, 4096]))))
            ii, jj = np.nonzero(np.triu(want, 1) >= 2)
            assert {(int(a), int(b), int(v)) for a, b, v, _ in cells} == {(int(i), int(j), int(want[i, j])) for i, j in zip(ii, jj)}, desc + " multi sparse"
            nrep = int(rng.integers(1, n))
            bs, bi = O.find_closest(O.batch(seqs[:nrep], K=K), O.batch(seqs[nrep:], K=K), ident_inf=ident, threads=4)
            db = m.db(); db.append(seqs[:nrep])
            sim, first, _ = db.find_closest(seqs[nrep:])
            assert (sim == bs).all() and (first.astype(np.int64) == np.where(bs > 0, bi.astype(np.int64), kr.NO_REP)).all(), desc + " multi findClosest x%d" % nd
        added = m.db().greedy_build(thr, seqs, batch=int(rng.choice([0, 7, 64])))
        assert (added == w).all(), desc + " multi greedy x%d thr=%d" % (nd, thr)
        m.close()
    return desc


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    t0, k = time.time(), 0
    while time.time() - t0 < a.seconds:
        state = rng.bit_generator.state
        try:
            d = one(rng)
        except Exception:
            print("FAILED at case %d; generator state:\n%r" % (k, state), flush=True)
            raise
        k += 1
        if k % 10 == 0:
            print("%d cases, %.0f s, last: %s" % (k, time.time() - t0, d), flush=True)
    print("OK %d cases" % k)
