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


class _HNode:
    """HashMap.Node / HashMap.TreeNode (tree fields used only while the node sits in a tree bin)"""
    __slots__ = ("hash", "key", "idx", "next", "prev", "parent", "left", "right", "red", "tree")

    def __init__(self, h, key, idx):
        self.hash, self.key, self.idx = h, key, idx
        self.next = self.prev = self.parent = self.left = self.right = None
        self.red = self.tree = False


class JavaHashMap:
    """java.util.HashMap<String,V> restated from OpenJDK's HashMap.java (putVal, resize, treeifyBin, TreeNode.treeify /
    putTreeVal / split / untreeify / removeTreeNode / balanceInsertion / balanceDeletion / moveRootToFront) to the
    extent iteration order depends on it.  TEST INFRASTRUCTURE.  String keys are Comparable and distinct keys never
    compare equal, so tieBreakOrder (identity hash) is never reached and the order is deterministic."""
    TREEIFY, UNTREEIFY, MIN_TREEIFY_CAP = 8, 6, 64

    def __init__(self, initial_cap=0):
        self.tab, self.size, self.thr, self.nodes = None, 0, 0, {}
        self.init_cap = 16
        self.explicit = initial_cap > 0
        if initial_cap > 0:
            c = 1
            while c < initial_cap:
                c <<= 1
            self.init_cap = c

    @staticmethod
    def spread(key):
        h = java_string_hash(key)
        h ^= h >> 16
        return h - (1 << 32) if h & 0x80000000 else h          # Java int

    # --- table -------------------------------------------------------------------------------
    def _resize(self):
        old = self.tab
        if old is None:
            cap = self.init_cap
            self.thr = int(cap * 0.75)
            self.tab = [None] * cap
            return
        ocap = len(old)
        ncap = ocap * 2
        self.thr = self.thr * 2 if ocap >= 16 else int(ncap * 0.75)
        new = [None] * ncap
        self.tab = new
        for j in range(ocap):
            e = old[j]
            if e is None:
                continue
            old[j] = None
            if e.next is None:
                new[e.hash & (ncap - 1)] = e
            elif e.tree:
                self._split(e, new, j, ocap)
            else:
                lo_h = lo_t = hi_h = hi_t = None
                while e is not None:
                    nxt = e.next
                    if (e.hash & ocap) == 0:
                        if lo_t is None:
                            lo_h = e
                        else:
                            lo_t.next = e
                        lo_t = e
                    else:
                        if hi_t is None:
                            hi_h = e
                        else:
                            hi_t.next = e
                        hi_t = e
                    e = nxt
                if lo_t is not None:
                    lo_t.next = None
                    new[j] = lo_h
                if hi_t is not None:
                    hi_t.next = None
                    new[j + ocap] = hi_h

    def put(self, key, idx):
        if key in self.nodes:
            return False
        if self.tab is None:
            self._resize()
        tab = self.tab
        h = self.spread(key)
        x = _HNode(h, key, idx)
        self.nodes[key] = x
        i = h & (len(tab) - 1)
        p = tab[i]
        if p is None:
            tab[i] = x
        elif p.tree:
            self._put_tree_val(p, x)
        else:
            count = 0
            while p.next is not None:
                p = p.next
                count += 1
            p.next = x
            if count >= self.TREEIFY - 1:
                self._treeify_bin(h)
        self.size += 1
        if self.size > self.thr:
            self._resize()
        return True

    def remove(self, key, movable=False):
        """movable=False: Iterator.remove();  movable=True: Map.remove(key)"""
        node = self.nodes.pop(key, None)
        if node is None:
            return
        tab = self.tab
        i = node.hash & (len(tab) - 1)
        if node.tree:
            self._remove_tree_node(node, movable)
        elif tab[i] is node:
            tab[i] = node.next
        else:
            p = tab[i]
            while p.next is not node:
                p = p.next
            p.next = node.next
        self.size -= 1

    def order(self):
        out = []
        for e in (self.tab or []):
            while e is not None:
                out.append(e.idx)
                e = e.next
        return out

    def tree_bins(self):
        return sum(1 for e in (self.tab or []) if e is not None and e.tree)

    # --- tree bins -----------------------------------------------------------------------------
    @staticmethod
    def _dir(x, p):
        if p.hash > x.hash:
            return -1
        if p.hash < x.hash:
            return 1
        return -1 if x.key < p.key else 1                      # String.compareTo on ASCII keys

    def _treeify_bin(self, h):
        tab = self.tab
        if len(tab) < self.MIN_TREEIFY_CAP:
            self._resize()
            return
        e = tab[h & (len(tab) - 1)]
        tl = None
        hd = e
        while e is not None:
            e.tree = True
            e.prev = tl
            e.parent = e.left = e.right = None
            e.red = False
            tl = e
            e = e.next
        self._treeify(hd)

    def _treeify(self, head):
        root = None
        x = head
        while x is not None:
            nxt = x.next
            x.left = x.right = None
            if root is None:
                x.parent = None
                x.red = False
                root = x
            else:
                p = root
                while True:
                    d = self._dir(x, p)
                    xp = p
                    p = p.left if d <= 0 else p.right
                    if p is None:
                        x.parent = xp
                        if d <= 0:
                            xp.left = x
                        else:
                            xp.right = x
                        root = self._balance_insertion(root, x)
                        break
            x = nxt
        self._move_root_to_front(root)

    def _untreeify(self, head):
        q = head
        while q is not None:
            q.tree = False
            q.prev = q.parent = q.left = q.right = None
            q.red = False
            q = q.next
        return head

    def _put_tree_val(self, first, x):
        root = first
        while root.parent is not None:
            root = root.parent
        p = root
        while True:
            d = self._dir(x, p)
            xp = p
            p = p.left if d <= 0 else p.right
            if p is None:
                xpn = xp.next
                x.tree = True
                x.next = xpn
                if d <= 0:
                    xp.left = x
                else:
                    xp.right = x
                xp.next = x
                x.parent = x.prev = xp
                if xpn is not None:
                    xpn.prev = x
                self._move_root_to_front(self._balance_insertion(root, x))
                return

    def _move_root_to_front(self, root):
        tab = self.tab
        if root is None or not tab:
            return
        i = root.hash & (len(tab) - 1)
        first = tab[i]
        if root is not first:
            tab[i] = root
            rp, rn = root.prev, root.next
            if rn is not None:
                rn.prev = rp
            if rp is not None:
                rp.next = rn
            if first is not None:
                first.prev = root
            root.next = first
            root.prev = None

    @staticmethod
    def _rotate_left(root, p):
        if p is not None and p.right is not None:
            r = p.right
            rl = p.right = r.left
            if rl is not None:
                rl.parent = p
            pp = r.parent = p.parent
            if pp is None:
                root = r
                r.red = False
            elif pp.left is p:
                pp.left = r
            else:
                pp.right = r
            r.left = p
            p.parent = r
        return root

    @staticmethod
    def _rotate_right(root, p):
        if p is not None and p.left is not None:
            l = p.left
            lr = p.left = l.right
            if lr is not None:
                lr.parent = p
            pp = l.parent = p.parent
            if pp is None:
                root = l
                l.red = False
            elif pp.right is p:
                pp.right = l
            else:
                pp.left = l
            l.right = p
            p.parent = l
        return root

    def _balance_insertion(self, root, x):
        x.red = True
        while True:
            xp = x.parent
            if xp is None:
                x.red = False
                return x
            xpp = xp.parent
            if not xp.red or xpp is None:
                return root
            xppl = xpp.left
            if xp is xppl:
                xppr = xpp.right
                if xppr is not None and xppr.red:
                    xppr.red = False
                    xp.red = False
                    xpp.red = True
                    x = xpp
                else:
                    if x is xp.right:
                        x = xp
                        root = self._rotate_left(root, x)
                        xp = x.parent
                        xpp = None if xp is None else xp.parent
                    if xp is not None:
                        xp.red = False
                        if xpp is not None:
                            xpp.red = True
                            root = self._rotate_right(root, xpp)
            else:
                if xppl is not None and xppl.red:
                    xppl.red = False
                    xp.red = False
                    xpp.red = True
                    x = xpp
                else:
                    if x is xp.left:
                        x = xp
                        root = self._rotate_right(root, x)
                        xp = x.parent
                        xpp = None if xp is None else xp.parent
                    if xp is not None:
                        xp.red = False
                        if xpp is not None:
                            xpp.red = True
                            root = self._rotate_left(root, xpp)

    def _balance_deletion(self, root, x):
        red = lambda n: n is not None and n.red
        while True:
            if x is None or x is root:
                return root
            xp = x.parent
            if xp is None:
                x.red = False
                return x
            if x.red:
                x.red = False
                return root
            xpl = xp.left
            if xpl is x:
                xpr = xp.right
                if xpr is not None and xpr.red:
                    xpr.red = False
                    xp.red = True
                    root = self._rotate_left(root, xp)
                    xp = x.parent
                    xpr = None if xp is None else xp.right
                if xpr is None:
                    x = xp
                else:
                    sl, sr = xpr.left, xpr.right
                    if not red(sr) and not red(sl):
                        xpr.red = True
                        x = xp
                    else:
                        if not red(sr):
                            if sl is not None:
                                sl.red = False
                            xpr.red = True
                            root = self._rotate_right(root, xpr)
                            xp = x.parent
                            xpr = None if xp is None else xp.right
                        if xpr is not None:
                            xpr.red = False if xp is None else xp.red
                            sr = xpr.right
                            if sr is not None:
                                sr.red = False
                        if xp is not None:
                            xp.red = False
                            root = self._rotate_left(root, xp)
                        x = root
            else:
                if xpl is not None and xpl.red:
                    xpl.red = False
                    xp.red = True
                    root = self._rotate_right(root, xp)
                    xp = x.parent
                    xpl = None if xp is None else xp.left
                if xpl is None:
                    x = xp
                else:
                    sl, sr = xpl.left, xpl.right
                    if not red(sl) and not red(sr):
                        xpl.red = True
                        x = xp
                    else:
                        if not red(sl):
                            if sr is not None:
                                sr.red = False
                            xpl.red = True
                            root = self._rotate_left(root, xpl)
                            xp = x.parent
                            xpl = None if xp is None else xp.left
                        if xpl is not None:
                            xpl.red = False if xp is None else xp.red
                            sl = xpl.left
                            if sl is not None:
                                sl.red = False
                        if xp is not None:
                            xp.red = False
                            root = self._rotate_right(root, xp)
                        x = root

    def _remove_tree_node(self, p, movable):
        tab = self.tab
        i = p.hash & (len(tab) - 1)
        first = root = tab[i]
        succ, pred = p.next, p.prev
        if pred is None:
            tab[i] = first = succ
        else:
            pred.next = succ
        if succ is not None:
            succ.prev = pred
        if first is None:
            return
        while root.parent is not None:
            root = root.parent
        if movable and (root.right is None or root.left is None or root.left.left is None):
            tab[i] = self._untreeify(first)
            return
        pl, pr = p.left, p.right
        if pl is not None and pr is not None:
            s = pr
            while s.left is not None:
                s = s.left
            s.red, p.red = p.red, s.red
            sr, pp = s.right, p.parent
            if s is pr:
                p.parent = s
                s.right = p
            else:
                sp = s.parent
                p.parent = sp
                if sp is not None:
                    if s is sp.left:
                        sp.left = p
                    else:
                        sp.right = p
                s.right = pr
                if pr is not None:
                    pr.parent = s
            p.left = None
            p.right = sr
            if sr is not None:
                sr.parent = p
            s.left = pl
            if pl is not None:
                pl.parent = s
            s.parent = pp
            if pp is None:
                root = s
            elif p is pp.left:
                pp.left = s
            else:
                pp.right = s
            replacement = sr if sr is not None else p
        elif pl is not None:
            replacement = pl
        elif pr is not None:
            replacement = pr
        else:
            replacement = p
        if replacement is not p:
            pp = replacement.parent = p.parent
            if pp is None:
                root = replacement
                root.red = False
            elif p is pp.left:
                pp.left = replacement
            else:
                pp.right = replacement
            p.left = p.right = p.parent = None
        r = root if p.red else self._balance_deletion(root, replacement)
        if replacement is p:
            pp = p.parent
            p.parent = None
            if pp is not None:
                if p is pp.left:
                    pp.left = None
                elif p is pp.right:
                    pp.right = None
        if movable:
            self._move_root_to_front(r)

    def _split(self, b, new, index, bit):
        lo_h = lo_t = hi_h = hi_t = None
        lc = hc = 0
        e = b
        while e is not None:
            nxt = e.next
            e.next = None
            if (e.hash & bit) == 0:
                e.prev = lo_t
                if lo_t is None:
                    lo_h = e
                else:
                    lo_t.next = e
                lo_t = e
                lc += 1
            else:
                e.prev = hi_t
                if hi_t is None:
                    hi_h = e
                else:
                    hi_t.next = e
                hi_t = e
                hc += 1
            e = nxt
        if lo_h is not None:
            if lc <= self.UNTREEIFY:
                new[index] = self._untreeify(lo_h)
            else:
                new[index] = lo_h
                if hi_h is not None:
                    self._treeify(lo_h)
        if hi_h is not None:
            if hc <= self.UNTREEIFY:
                new[index + bit] = self._untreeify(hi_h)
            else:
                new[index + bit] = hi_h
                if lo_h is not None:
                    self._treeify(hi_h)


def py_java_hashmap_order(keys, initial_cap: int = 0, removals=None):
    """Pure-Python java.util.HashMap<String,V> iteration order after put()ting `keys` in order (tree bins included).
    `removals`: optional {position: [(key, movable), ...]} applied after the put of keys[position]."""
    m = JavaHashMap(initial_cap)
    for i, k in enumerate(keys):
        m.put(k, i)
        if removals and i in removals:
            for key, movable in removals[i]:
                m.remove(key, movable)
    return m.order()


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
        if n < 0:                                   # a tree bin formed: the C fast path stops there
            return np.array(py_java_hashmap_order(keys, initial_cap), dtype=np.int32)
        return out[:n].copy()


_ORACLE = None


def get() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
