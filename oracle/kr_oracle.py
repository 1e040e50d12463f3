"""TEST INFRASTRUCTURE ONLY -- loader for oracle/kr_oracle.c plus a pure-Python mirror.

PARITY UNPINNED (see the header of kr_oracle.c): the hot-path classes of the reference live
in un-vendored org.theseed:shared:1.0.0 / sequence:1.0.0 and there is no JVM here, so this
restatement is anchored on the reference's data files, not on outputs of the Java code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (kmers/reps_b200) never does.

Two independent restatements live here so that they can check one another:
  * `py_*`   : the 5-line set-of-substrings model (pure Python; small cases only),
  * `Oracle` : ctypes binding of kr_oracle.c (hash set of substrings probed key by key,
               the data-structure choice of the reference's HashSet<String>).
Reference call sites each function follows are cited in kr_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from decimal import Decimal, ROUND_HALF_UP

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libkr_oracle.so")
INFINITY = 9999


# ----------------------------------------------------------------------------- pure Python
def py_kmers(seq: str, K: int = 8) -> set:
    """ProteinKmers(String): all distinct length-K substrings (SequenceInfo.java:47)."""
    return {seq[i:i + K] for i in range(len(seq) - K + 1)}


def py_similarity(a: str, b: str, K: int = 8, ident_inf: bool = True) -> int:
    """SequenceKmers.similarity (DistanceMatrixProcessor.java:76)."""
    if ident_inf and a == b:
        return INFINITY
    return len(py_kmers(a, K) & py_kmers(b, K))


def py_distance(sim: int, size_a: int, size_b: int) -> float:
    """SequenceKmers.distance(sim, a, b) (ScoreCouplingReporter.java:72-73)."""
    if sim == INFINITY:
        return 0.0
    if sim == 0:
        return 1.0
    return 1.0 - (sim / float(size_a + size_b - sim))


def java_string_hash(s: str) -> int:
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return h


def py_java_hashmap_order(keys, initial_cap: int = 0):
    """Pure-Python java.util.HashMap<String,V> iteration order (no tree bins)."""
    cap, thr, size = 0, 0, 0
    tab = None
    init = 16
    if initial_cap > 0:
        init = 1
        while init < initial_cap:
            init <<= 1
    seen = {}

    def resize():
        nonlocal cap, thr, tab
        ocap = cap
        nt = [[] for _ in range(2 * ocap)]
        for j in range(ocap):
            for (h, i) in tab[j]:
                nt[j + ocap if (h & ocap) else j].append((h, i))
        tab = nt
        cap = 2 * ocap
        thr = thr * 2 if ocap >= 16 else int(cap * 0.75)

    for i, k in enumerate(keys):
        if k in seen:
            continue
        seen[k] = i
        h = java_string_hash(k)
        h ^= h >> 16
        if tab is None:
            cap = init
            thr = int(cap * 0.75)
            tab = [[] for _ in range(cap)]
        b = tab[h & (cap - 1)]
        depth = len(b)
        b.append((h, i))
        if depth >= 8:
            if cap < 64:
                resize()
            else:
                raise ValueError("tree bin would form; not emulated")
        size += 1
        if size > thr:
            resize()
    return [i for b in (tab or []) for (_, i) in b]


def genome_id_of(fid: str) -> str:
    """Feature.genomeOf / RepGenome.getGenomeId: the digits.digits part of fig|G.peg.N."""
    import re
    m = re.match(r"fig\|(\d+\.\d+)\.", fid)
    return m.group(1) if m else fid


# ----------------------------------------------------------------------------- FASTA / rep DB
def read_fasta(data: bytes):
    """FastaInputStream: records (label, comment, sequence); a header with no sequence
    lines yields an empty sequence (the rep-DB header record, data/rep50.db:1)."""
    recs = []
    label = comment = None
    parts = []
    for line in data.decode("ascii").split("\n"):
        if line.endswith("\r"):
            line = line[:-1]
        if line.startswith(">"):
            if label is not None:
                recs.append((label, comment, "".join(parts)))
            hdr = line[1:]
            sp = hdr.split(None, 1)
            label = sp[0] if sp else ""
            comment = sp[1].strip() if len(sp) > 1 else ""
            parts = []
        elif label is not None and line.strip():
            parts.append(line.strip())
    if label is not None:
        recs.append((label, comment, "".join(parts)))
    return recs


def write_fasta_record(label: str, comment: str, seq: str, width: int = 60) -> bytes:
    out = [">" + label + ((" " + comment) if comment else "")]
    for i in range(0, len(seq), width):
        out.append(seq[i:i + width])
    return ("\n".join(out) + "\n").encode("ascii")


def parse_repdb(data: bytes):
    """RepGenomeDb.load: record 0 = '>Rep<thr>,K=<K> <protName>'; the rest are reps."""
    import re
    recs = read_fasta(data)
    m = re.match(r"Rep(\d+),K=(\d+)$", recs[0][0])
    if not m:
        raise ValueError("bad rep DB header: %r" % (recs[0][0],))
    return dict(threshold=int(m.group(1)), K=int(m.group(2)), prot_name=recs[0][1], reps=recs[1:])


def save_repdb(threshold: int, K: int, prot_name: str, reps) -> bytes:
    """RepGenomeDb.save: header record without sequence, then the reps in the given order."""
    out = [(">Rep%d,K=%d %s\n" % (threshold, K, prot_name)).encode("ascii")]
    for (fid, name, seq) in reps:
        out.append(write_fasta_record(fid, name, seq))
    return b"".join(out)


# ----------------------------------------------------------------------------- Java formatting
def java_fixed(x: float, width: int, prec: int) -> str:
    """String.format("%<width>.<prec>f"): Java rounds the shortest repr decimal HALF_UP."""
    d = Decimal(repr(float(x))).quantize(Decimal(1).scaleb(-prec), rounding=ROUND_HALF_UP)
    return format(d, "f").rjust(width)


# ----------------------------------------------------------------------------- C binding
def build_lib(force: bool = False) -> str:
    src = os.path.join(HERE, "kr_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB_PATH


def flat(seqs):
    """list[str|bytes] -> (uint8 residues, uint64 offsets[S+1])"""
    bs = [s.encode("ascii") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    res = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return res, off


class Batch:
    def __init__(self, oracle, seqs=None, K=8, res=None, off=None):
        self.o = oracle
        if res is None:
            res, off = flat(seqs)
        # keep the residue buffer alive and NUL-pad so an empty batch has a valid pointer
        self.res = np.concatenate([np.ascontiguousarray(res, dtype=np.uint8), np.zeros(1, np.uint8)])
        self.off = np.ascontiguousarray(off, dtype=np.uint64)
        self.n = len(self.off) - 1
        self.K = K
        self.h = oracle.lib.kro_batch_new(self.res.ctypes.data, self.off.ctypes.data, self.n, K)

    def sizes(self):
        out = np.zeros(self.n, dtype=np.int32)
        self.o.lib.kro_batch_sizes(self.h, out.ctypes.data)
        return out

    def __del__(self):
        try:
            self.o.lib.kro_batch_free(self.h)
        except Exception:
            pass


class Oracle:
    def __init__(self):
        self.lib = lib = C.CDLL(build_lib())
        vp, i32, u64p = C.c_void_p, C.c_int, C.c_void_p
        lib.kro_batch_new.restype = vp
        lib.kro_batch_new.argtypes = [vp, u64p, i32, i32]
        lib.kro_batch_free.argtypes = [vp]
        lib.kro_batch_sizes.argtypes = [vp, vp]
        lib.kro_sim_rows.argtypes = [vp, vp, i32, i32, i32, vp]
        lib.kro_sim_triangle_rows.argtypes = [vp, i32, i32, i32, vp]
        lib.kro_sim_matrix_mt.argtypes = [vp, vp, i32, i32, vp]
        lib.kro_all_pairs_count_mt.restype = C.c_uint64
        lib.kro_all_pairs_count_mt.argtypes = [vp, i32, i32, vp]
        lib.kro_find_closest.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp]
        lib.kro_find_closest_mt.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]
        lib.kro_greedy_build.restype = i32
        lib.kro_greedy_build.argtypes = [vp, i32, i32, vp, vp]
        lib.kro_distance.restype = C.c_double
        lib.kro_distance.argtypes = [i32, i32, i32]
        lib.kro_similarity.restype = i32
        lib.kro_similarity.argtypes = [C.c_char_p, i32, C.c_char_p, i32, i32, i32]
        lib.kro_java_hashmap_order.restype = i32
        lib.kro_java_hashmap_order.argtypes = [vp, i32, i32, vp]

    def batch(self, seqs=None, K=8, res=None, off=None) -> Batch:
        return Batch(self, seqs, K, res, off)

    def similarity(self, a: str, b: str, K=8, ident_inf=True) -> int:
        ab, bb = a.encode(), b.encode()
        return self.lib.kro_similarity(ab, len(ab), bb, len(bb), K, int(ident_inf))

    def distance(self, sim, na, nb) -> float:
        return self.lib.kro_distance(int(sim), int(na), int(nb))

    def sim_matrix(self, A: Batch, B: Batch | None = None, ident_inf=True, threads=1):
        B = B or A
        out = np.zeros((A.n, B.n), dtype=np.int32)
        if A.n and B.n:
            self.lib.kro_sim_matrix_mt(A.h, B.h, int(ident_inf), threads, out.ctypes.data)
        return out

    def all_pairs_timed(self, A: Batch, ident_inf=True, threads=1):
        cs = C.c_int64(0)
        n = self.lib.kro_all_pairs_count_mt(A.h, int(ident_inf), threads, C.byref(cs))
        return int(n), int(cs.value)

    def find_closest(self, reps: Batch, queries: Batch, order=None, ident_inf=True, tie_last=False, threads=1):
        bs = np.zeros(queries.n, dtype=np.int32)
        bi = np.zeros(queries.n, dtype=np.int32)
        op = None
        if order is not None:
            order = np.ascontiguousarray(order, dtype=np.int32)
            op = order.ctypes.data
        self.lib.kro_find_closest_mt(reps.h, op, queries.h, int(ident_inf), int(tie_last), threads,
                                     bs.ctypes.data, bi.ctypes.data)
        return bs, bi

    def greedy_build(self, cands: Batch, threshold: int, ident_inf=True):
        added = np.zeros(cands.n, dtype=np.uint8)
        pairs = C.c_uint64(0)
        n = self.lib.kro_greedy_build(cands.h, threshold, int(ident_inf), added.ctypes.data, C.byref(pairs))
        return n, added, int(pairs.value)

    def java_hashmap_order(self, keys, initial_cap=0):
        arr = (C.c_char_p * len(keys))(*[k.encode("ascii") for k in keys])
        out = np.zeros(max(len(keys), 1), dtype=np.int32)
        n = self.lib.kro_java_hashmap_order(C.cast(arr, C.c_void_p), len(keys), initial_cap, out.ctypes.data)
        if n < 0:
            raise ValueError("tree bin would form; not emulated")
        return out[:n].copy()


_ORACLE = None


def get() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
