#!/usr/bin/env python
"""A small pass over every experimental device path, meant to run on the GPU box under compute-sanitizer BEFORE anything
larger does (tools/exp_gpu_check.sh):

    KR_LIB_PATH=.../libkmersreps_cuda_exp.so compute-sanitizer --tool memcheck  python tools/exp_sanitize.py
    KR_LIB_PATH=.../libkmersreps_cuda_exp.so compute-sanitizer --tool racecheck python tools/exp_sanitize.py

An out-of-bounds access or a shared-memory race then shows up as a sanitizer report on a few hundred sequences instead of
as a device fault on a large run.  Results are compared with the default paths of the same library (GPU against GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmers.reps_b200 as kr                      # noqa: E402
from kmers.reps_b200 import binding, synth        # noqa: E402

assert binding.has_experimental(), "point KR_LIB_PATH at libkmersreps_cuda_exp.so"
res, off, _ = synth.generate(0x5EED0031, 700, 60)
seqs = synth.to_strings(res, off) + ["MKV", "", "MKV"]
ref = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)
s0 = ref.encode(seqs)
m0 = ref.sim_matrix(s0, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI)
a0 = ref.db().greedy_build(s0, 50, batch=200)
for name, flags in (("encode", binding.EXP_ENCODE), ("pairs", binding.EXP_PAIRS), ("pairs+buckets", binding.EXP_PAIRS | binding.EXP_BUCKETS),
                    ("gidx", binding.EXP_GIDX)):
    c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | flags)
    s = c.encode(seqs)
    assert (s.sizes() == s0.sizes()).all(), name
    assert (c.sim_matrix(s, elem=kr.ELEM_U16, layout=kr.LAYOUT_BLOCKTRI) == m0).all(), name
    assert (c.db().greedy_build(s, 50, batch=200) == a0).all(), name
    cells = c.sim_matrix_sparse(s, min_sim=1, layout=kr.LAYOUT_TRIANGLE)
    full = ref.sim_matrix(s0, elem=kr.ELEM_I32)
    assert len(cells) == int((np.triu(full, 1) > 0).sum()), name
    c.close()
    print("ok", name, flush=True)
# index query: dense per-warp table (small database) and per-warp hash (> 4096 representatives)
for nrep in (500, 4500):
    r2, o2, _ = synth.generate(0x5EED0032, nrep + 300, 200)
    q2 = [x[:100] for x in synth.to_strings(r2, o2)]
    out = []
    for flags in (binding.EXP_IDXQ, 0):
        c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY | kr.FORCE_INDEX | flags)
        db = c.db()
        db.append(q2[:nrep])
        out.append(db.find_closest(c.encode(q2[nrep:])))
        c.close()
    assert all((a == b).all() for a, b in zip(*out)), nrep
    print("ok idxq", nrep, flush=True)
# DNA contexts
rng = np.random.default_rng(3)
alpha = np.frombuffer(b"acgt", dtype=np.uint8)
dna = [alpha[rng.integers(0, 4, int(rng.integers(300, 1600)))].tobytes().decode() for _ in range(40)]
c = kr.Context(device=0, K=18, flags=kr.IDENT_IS_INFINITY | binding.ALPHABET_DNA)
d = c.encode(dna + [dna[0].upper()])
m = c.sim_matrix(d, elem=kr.ELEM_I32)
assert m[0, 40] == 9999 and (np.diag(m) == 9999).all()
c.close()
ref.close()
print("ALL OK")
