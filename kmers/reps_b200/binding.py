"""ctypes binding of include/kmersreps.h (one Python method per C entry point)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
INFINITY = 9999
NO_REP = 0xFFFFFFFF
IDENT_IS_INFINITY = 1
TIE_LAST = 2
FORCE_JOIN = 4
FORCE_INDEX = 8
ELEM_I32, ELEM_U16 = 0, 1
LAYOUT_FULL, LAYOUT_TRIANGLE, LAYOUT_BLOCKTRI = 0, 1, 2
TILE = 256

# every symbol include/kmersreps.h declares (tests check that the library exports them all)
SYMBOLS = [
    "kr_create", "kr_destroy", "kr_last_error", "kr_kmer_size", "kr_build_info", "kr_synchronize", "kr_encode", "kr_encode_device",
    "kr_seqset_free", "kr_seqset_count", "kr_seqset_export_blocklists", "kr_seqset_import_blocklists", "kr_blocklist_capacity", "kr_seqset_sizes", "kr_seqset_keys", "kr_key_decode", "kr_seqset_gather", "kr_seqset_from_keys",
    "kr_sim_pairs", "kr_distance", "kr_matrix_elems", "kr_matrix_stripes", "kr_sim_matrix", "kr_sim_matrix_device",
    "kr_sim_matrix_sparse", "kr_sim_matrix_sparse_device", "kr_sim_matrix_sparse64", "kr_sim_matrix_sparse64_device",
    "kr_db_create", "kr_db_free", "kr_db_size", "kr_db_append", "kr_db_append_set", "kr_db_set_ranks",
    "kr_db_find_closest", "kr_db_find_closest_device", "kr_db_any_ge", "kr_greedy_build", "kr_stats_get",
    "kr_stats_reset", "kr_timer_begin", "kr_timer_end", "kr_host_alloc", "kr_host_free",
    "kr_multi_create", "kr_multi_destroy", "kr_multi_ndev", "kr_multi_ctx", "kr_multi_last_error", "kr_multi_sim_matrix",
    "kr_multi_sim_matrix_sparse", "kr_multi_db_create", "kr_multi_db_free", "kr_multi_db_size", "kr_multi_db_append",
    "kr_multi_db_set_ranks", "kr_multi_db_find_closest", "kr_multi_db_any_ge", "kr_multi_greedy_build", "kr_multi_stats_get",
]


class KrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("kmersreps error %d: %s" % (code, msg))
        self.code = code


class MatrixDesc(C.Structure):
    _fields_ = [("elem", C.c_uint32), ("layout", C.c_uint32), ("row_begin", C.c_uint32), ("row_end", C.c_uint32)]


class BlockLists(C.Structure):
    _fields_ = [("nseq", C.c_uint32), ("nblk", C.c_uint32), ("entries", C.c_uint64), ("lo_size", C.c_uint32), ("nchunk", C.c_uint32),
                ("hi", C.c_void_p), ("lo", C.c_void_p), ("chunk_off", C.c_void_p), ("chunk_cnt", C.c_void_p), ("blk_base", C.c_void_p),
                ("meta", C.c_void_p)]

    def arrays(self):
        """name -> (device pointer, bytes) of every exchanged array"""
        return {"hi": (self.hi, self.entries * 4), "lo": (self.lo, self.entries * self.lo_size),
                "chunk_off": (self.chunk_off, self.nblk * (self.nchunk + 1) * 4), "chunk_cnt": (self.chunk_cnt, self.nblk * self.nchunk * 4),
                "blk_base": (self.blk_base, (self.nblk + 1) * 8), "meta": (self.meta, self.nseq * 24)}


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("pairs", C.c_uint64), ("encode_ms", C.c_double),
                ("blocklist_ms", C.c_double), ("join_ms", C.c_double), ("other_ms", C.c_double),
                ("join_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("index_ms", C.c_double), ("scan_ms", C.c_double), ("scan_launches", C.c_uint64), ("index_builds", C.c_uint64),
                ("index_entries", C.c_uint64), ("index_walk", C.c_uint64), ("fill_ms", C.c_double), ("fill_launches", C.c_uint64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def lib_path() -> str:
    # KR_LIB_PATH lets a tuning run point at another build of the same library (never at a fallback)
    return os.environ.get("KR_LIB_PATH") or os.path.join(HERE, "libkmersreps_cuda.so")


_LIB = None


def load_library():
    """Load libkmersreps_cuda.so; raise (never fall back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise KrError(-3, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
    lib.kr_create.argtypes = [C.c_int, C.c_int, u32, C.POINTER(vp)]
    lib.kr_destroy.argtypes = [vp]
    lib.kr_destroy.restype = None
    lib.kr_last_error.argtypes = [vp]
    lib.kr_last_error.restype = C.c_char_p
    lib.kr_kmer_size.argtypes = [vp]
    lib.kr_build_info.argtypes = []
    lib.kr_build_info.restype = C.c_char_p
    lib.kr_synchronize.argtypes = [vp]
    lib.kr_encode.argtypes = [vp, vp, vp, u32, C.POINTER(vp)]
    lib.kr_encode_device.argtypes = [vp, vp, vp, u32, C.POINTER(vp)]
    lib.kr_seqset_free.argtypes = [vp]
    lib.kr_seqset_free.restype = None
    lib.kr_seqset_count.argtypes = [vp]
    lib.kr_seqset_count.restype = u32
    lib.kr_blocklist_capacity.argtypes = [vp, vp, u32]
    lib.kr_blocklist_capacity.restype = u64
    lib.kr_seqset_export_blocklists.argtypes = [vp, vp, C.POINTER(BlockLists)]
    lib.kr_seqset_import_blocklists.argtypes = [vp, C.POINTER(BlockLists), u32, vp, vp, u32, C.POINTER(vp)]
    lib.kr_seqset_sizes.argtypes = [vp, vp]
    lib.kr_seqset_keys.argtypes = [vp, u32, vp, u32, C.POINTER(u32)]
    lib.kr_key_decode.argtypes = [vp, u64, C.c_char_p]
    lib.kr_seqset_gather.argtypes = [vp, vp, vp, u32, C.POINTER(vp)]
    lib.kr_seqset_from_keys.argtypes = [vp, vp, vp, vp, vp, u32, C.POINTER(vp)]
    lib.kr_sim_pairs.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp]
    lib.kr_distance.argtypes = [i32, u32, u32]
    lib.kr_distance.restype = C.c_double
    lib.kr_matrix_elems.argtypes = [u32, u32, u32, u32, u32]
    lib.kr_matrix_elems.restype = u64
    lib.kr_matrix_stripes.argtypes = [u32, u32, u32, u32, vp]
    lib.kr_sim_matrix.argtypes = [vp, vp, vp, C.POINTER(MatrixDesc), vp]
    lib.kr_sim_matrix_device.argtypes = [vp, vp, vp, C.POINTER(MatrixDesc), vp]
    lib.kr_db_create.argtypes = [vp, C.POINTER(vp)]
    lib.kr_db_free.argtypes = [vp]
    lib.kr_db_free.restype = None
    lib.kr_db_size.argtypes = [vp]
    lib.kr_db_size.restype = u32
    lib.kr_db_append.argtypes = [vp, vp, vp, vp, u32]
    lib.kr_db_append_set.argtypes = [vp, vp, vp, vp, u32]
    lib.kr_db_set_ranks.argtypes = [vp, vp, vp, u32]
    lib.kr_db_find_closest.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.kr_db_find_closest_device.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.kr_db_any_ge.argtypes = [vp, vp, vp, i32, vp]
    lib.kr_greedy_build.argtypes = [vp, vp, vp, i32, u32, vp]
    lib.kr_stats_get.argtypes = [vp, C.POINTER(Stats)]
    lib.kr_stats_reset.argtypes = [vp]
    lib.kr_timer_begin.argtypes = [vp]
    lib.kr_timer_end.argtypes = [vp, C.POINTER(C.c_double)]
    lib.kr_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.kr_host_free.argtypes = [vp, vp]
    lib.kr_host_free.restype = None
    lib.kr_sim_matrix_sparse.argtypes = [vp, vp, vp, C.POINTER(MatrixDesc), i32, vp, u64, C.POINTER(u64)]
    lib.kr_sim_matrix_sparse_device.argtypes = [vp, vp, vp, C.POINTER(MatrixDesc), i32, vp, u64, C.POINTER(u64)]
    lib.kr_sim_matrix_sparse64.argtypes = [vp, vp, vp, C.POINTER(MatrixDesc), i32, vp, u64, C.POINTER(u64)]
    lib.kr_sim_matrix_sparse64_device.argtypes = [vp, vp, vp, C.POINTER(MatrixDesc), i32, vp, u64, C.POINTER(u64)]
    lib.kr_multi_create.argtypes = [vp, C.c_int, C.c_int, u32, C.POINTER(vp)]
    lib.kr_multi_destroy.argtypes = [vp]
    lib.kr_multi_destroy.restype = None
    lib.kr_multi_ndev.argtypes = [vp]
    lib.kr_multi_ctx.argtypes = [vp, C.c_int]
    lib.kr_multi_ctx.restype = vp
    lib.kr_multi_last_error.argtypes = [vp]
    lib.kr_multi_last_error.restype = C.c_char_p
    lib.kr_multi_sim_matrix.argtypes = [vp, vp, vp, u32, vp, vp, u32, u32, u32, vp]
    lib.kr_multi_sim_matrix_sparse.argtypes = [vp, vp, vp, u32, vp, vp, u32, i32, u32, vp, u64, C.POINTER(u64)]
    lib.kr_multi_db_create.argtypes = [vp, C.POINTER(vp)]
    lib.kr_multi_db_free.argtypes = [vp]
    lib.kr_multi_db_free.restype = None
    lib.kr_multi_db_size.argtypes = [vp]
    lib.kr_multi_db_size.restype = u32
    lib.kr_multi_db_append.argtypes = [vp, vp, vp, vp, u32]
    lib.kr_multi_db_set_ranks.argtypes = [vp, vp, vp, u32]
    lib.kr_multi_db_find_closest.argtypes = [vp, vp, vp, vp, u32, vp, vp, vp]
    lib.kr_multi_db_any_ge.argtypes = [vp, vp, vp, vp, u32, i32, vp]
    lib.kr_multi_greedy_build.argtypes = [vp, vp, vp, vp, u32, i32, u32, vp]
    lib.kr_multi_stats_get.argtypes = [vp, C.POINTER(Stats)]
    _LIB = lib
    return lib


def build_info() -> str:
    return load_library().kr_build_info().decode()


def flat(seqs):
    """list[str|bytes] -> (uint8 residues, uint64 offsets[S+1])"""
    bs = [s.encode("ascii") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    res = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return res, off


def distance(sim, size_a, size_b) -> float:
    return load_library().kr_distance(int(sim), int(size_a), int(size_b))


def matrix_elems(n_rows, n_cols, layout, row_begin=0, row_end=None) -> int:
    return int(load_library().kr_matrix_elems(n_rows, n_cols, layout, row_begin, n_rows if row_end is None else row_end))


def matrix_stripes(n_rows, n_cols, layout, nshards):
    out = np.zeros(nshards + 1, dtype=np.uint32)
    rc = load_library().kr_matrix_stripes(n_rows, n_cols, layout, nshards, out.ctypes.data)
    if rc:
        raise KrError(rc, "kr_matrix_stripes")
    return out


def _ptr(a):
    return None if a is None else a.ctypes.data


class Context:
    """One GPU, one K (the static ProteinKmers.setKmerSize of the reference)."""

    def __init__(self, device=0, K=8, flags=IDENT_IS_INFINITY):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.kr_create(device, K, flags, C.byref(h))
        if rc:
            raise KrError(rc, self.lib.kr_last_error(None).decode())
        self.h = h
        self.K = K
        self.flags = flags

    def _ck(self, rc):
        if rc:
            raise KrError(rc, self.lib.kr_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            for p in getattr(self, "_pinned", []):
                self.lib.kr_host_free(self.h, p)
            self._pinned = []
            self.lib.kr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._ck(self.lib.kr_synchronize(self.h))

    def encode(self, seqs=None, res=None, off=None) -> "SeqSet":
        if res is None:
            res, off = flat(seqs)
        res = np.ascontiguousarray(res, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        h = C.c_void_p()
        pad = res if res.size else np.zeros(1, np.uint8)
        self._ck(self.lib.kr_encode(self.h, pad.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(h)))
        return SeqSet(self, h)

    def encode_device(self, d_res_ptr: int, off) -> "SeqSet":
        off = np.ascontiguousarray(off, dtype=np.uint64)
        h = C.c_void_p()
        self._ck(self.lib.kr_encode_device(self.h, d_res_ptr, off.ctypes.data, len(off) - 1, C.byref(h)))
        return SeqSet(self, h)

    def from_keys(self, key_lists, seqs=None) -> "SeqSet":
        """a set from explicit sorted key lists (the result of retainAll / removeAll); seqs = the objects' sequences"""
        koff = np.zeros(len(key_lists) + 1, dtype=np.uint64)
        if key_lists:
            koff[1:] = np.cumsum([len(k) for k in key_lists], dtype=np.uint64)
        keys = np.concatenate([np.asarray(k, dtype=np.uint64) for k in key_lists]) if key_lists and int(koff[-1]) else np.zeros(1, np.uint64)
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        res = off = None
        if seqs is not None:
            res, off = flat(seqs)
            if not res.size:
                res = np.zeros(1, np.uint8)
        h = C.c_void_p()
        self._ck(self.lib.kr_seqset_from_keys(self.h, keys.ctypes.data, koff.ctypes.data, _ptr(res), _ptr(off), len(key_lists), C.byref(h)))
        return SeqSet(self, h)

    def blocklist_capacity(self, off) -> int:
        off = np.ascontiguousarray(off, dtype=np.uint64)
        return int(self.lib.kr_blocklist_capacity(self.h, off.ctypes.data, len(off) - 1))

    def export_blocklists(self, s: "SeqSet") -> BlockLists:
        b = BlockLists()
        self._ck(self.lib.kr_seqset_export_blocklists(self.h, s.h, C.byref(b)))
        return b

    def import_blocklists(self, shards, d_res_ptr: int, off) -> "SeqSet":
        off = np.ascontiguousarray(off, dtype=np.uint64)
        arr = (BlockLists * len(shards))(*shards)
        h = C.c_void_p()
        self._ck(self.lib.kr_seqset_import_blocklists(self.h, arr, len(shards), d_res_ptr, off.ctypes.data, len(off) - 1, C.byref(h)))
        return SeqSet(self, h)

    def key_decode(self, key: int) -> str:
        buf = C.create_string_buffer(self.K)
        self._ck(self.lib.kr_key_decode(self.h, int(key), buf))
        return buf.raw.decode()

    def sim_pairs(self, A: "SeqSet", ia, B: "SeqSet", ib):
        ia = np.ascontiguousarray(ia, dtype=np.uint32)
        ib = np.ascontiguousarray(ib, dtype=np.uint32)
        n = len(ia)
        sim = np.zeros(n, dtype=np.int32)
        ident = np.zeros(n, dtype=np.uint8)
        self._ck(self.lib.kr_sim_pairs(self.h, A.h, ia.ctypes.data, B.h, ib.ctypes.data, n, sim.ctypes.data, ident.ctypes.data))
        return sim, ident

    def sim_matrix(self, A: "SeqSet", B: "SeqSet" = None, elem=ELEM_I32, layout=LAYOUT_FULL, row_begin=0, row_end=None, out=None):
        n_rows = A.count
        n_cols = A.count if B is None else B.count
        row_end = n_rows if row_end is None else row_end
        d = MatrixDesc(elem, layout, row_begin, row_end)
        ne = matrix_elems(n_rows, n_cols, layout, row_begin, row_end)
        if out is None:
            out = np.zeros(ne, dtype=np.uint16 if elem == ELEM_U16 else np.int32)
        self._ck(self.lib.kr_sim_matrix(self.h, A.h, None if B is None else B.h, C.byref(d), out.ctypes.data))
        if layout == LAYOUT_FULL:
            return out.reshape(row_end - row_begin, n_cols)
        return out

    def sim_matrix_sparse(self, A: "SeqSet", B: "SeqSet" = None, min_sim=1, layout=LAYOUT_TRIANGLE, row_begin=0, row_end=None, cap=1 << 22, out=None):
        """the cells (row, col, sim, 0) with sim >= min_sim, unordered, as an (n, 4) uint32 array"""
        row_end = A.count if row_end is None else row_end
        d = MatrixDesc(ELEM_I32, layout, row_begin, row_end)
        while True:
            cells = out if out is not None else np.zeros((cap, 4), dtype=np.uint32)
            n = C.c_uint64(0)
            rc = self.lib.kr_sim_matrix_sparse(self.h, A.h, None if B is None else B.h, C.byref(d), int(min_sim), cells.ctypes.data, len(cells), C.byref(n))
            if rc == 0:
                return cells[: n.value]
            if n.value <= len(cells) or out is not None:
                self._ck(rc)
            cap = int(n.value)                     # the library reported how many cells qualify: retry with room

    def sim_matrix_sparse64(self, A: "SeqSet", B: "SeqSet" = None, min_sim=1, layout=LAYOUT_TRIANGLE, row_begin=0, row_end=None, cap=1 << 22, out=None):
        """the same cells packed into one uint64 each (row << 40 | col << 16 | sim), as a uint64 array; unpack_cells() splits them"""
        row_end = A.count if row_end is None else row_end
        d = MatrixDesc(ELEM_I32, layout, row_begin, row_end)
        while True:
            cells = out if out is not None else np.zeros(cap, dtype=np.uint64)
            n = C.c_uint64(0)
            rc = self.lib.kr_sim_matrix_sparse64(self.h, A.h, None if B is None else B.h, C.byref(d), int(min_sim), cells.ctypes.data, len(cells), C.byref(n))
            if rc == 0:
                return cells[: n.value]
            if n.value <= len(cells) or out is not None:
                self._ck(rc)
            cap = int(n.value)

    def sim_matrix_sparse64_device(self, A: "SeqSet", B, d_cells_ptr: int, cap: int, min_sim=1, layout=LAYOUT_TRIANGLE, row_begin=0, row_end=None) -> int:
        row_end = A.count if row_end is None else row_end
        d = MatrixDesc(ELEM_I32, layout, row_begin, row_end)
        n = C.c_uint64(0)
        self._ck(self.lib.kr_sim_matrix_sparse64_device(self.h, A.h, None if B is None else B.h, C.byref(d), int(min_sim), d_cells_ptr, int(cap), C.byref(n)))
        return int(n.value)

    def sim_matrix_sparse_device(self, A: "SeqSet", B, d_cells_ptr: int, cap: int, min_sim=1, layout=LAYOUT_TRIANGLE, row_begin=0, row_end=None) -> int:
        """same, cells left in device memory (16 bytes each); returns how many qualify"""
        row_end = A.count if row_end is None else row_end
        d = MatrixDesc(ELEM_I32, layout, row_begin, row_end)
        n = C.c_uint64(0)
        self._ck(self.lib.kr_sim_matrix_sparse_device(self.h, A.h, None if B is None else B.h, C.byref(d), int(min_sim), d_cells_ptr, int(cap), C.byref(n)))
        return int(n.value)

    def sim_matrix_device(self, A, B, d_out_ptr: int, elem=ELEM_U16, layout=LAYOUT_FULL, row_begin=0, row_end=None):
        row_end = A.count if row_end is None else row_end
        d = MatrixDesc(elem, layout, row_begin, row_end)
        self._ck(self.lib.kr_sim_matrix_device(self.h, A.h, None if B is None else B.h, C.byref(d), d_out_ptr))

    def db(self) -> "RepDb":
        h = C.c_void_p()
        self._ck(self.lib.kr_db_create(self.h, C.byref(h)))
        return RepDb(self, h)

    def stats(self) -> dict:
        s = Stats()
        self._ck(self.lib.kr_stats_get(self.h, C.byref(s)))
        return s.asdict()

    def stats_reset(self):
        self._ck(self.lib.kr_stats_reset(self.h))

    def timer_begin(self):
        self._ck(self.lib.kr_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_double(0)
        self._ck(self.lib.kr_timer_end(self.h, C.byref(ms)))
        return ms.value

    def host_alloc(self, nbytes: int, dtype=np.uint8):
        """pinned host array (freed when the context closes or by host_free)"""
        p = C.c_void_p()
        self._ck(self.lib.kr_host_alloc(self.h, nbytes, C.byref(p)))
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype)
        self._pinned = getattr(self, "_pinned", []) + [p]
        return arr


class SeqSet:
    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        self.count = int(ctx.lib.kr_seqset_count(h))

    def free(self):
        if self.h:
            self.ctx.lib.kr_seqset_free(self.h)
            self.h = None
        self._keep = None   # buffers an imported set was built from

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass

    def sizes(self):
        out = np.zeros(self.count, dtype=np.uint32)
        self.ctx._ck(self.ctx.lib.kr_seqset_sizes(self.h, out.ctypes.data))
        return out

    def keys(self, i: int):
        n = C.c_uint32(0)
        self.ctx.lib.kr_seqset_keys(self.h, i, None, 0, C.byref(n))
        out = np.zeros(max(n.value, 1), dtype=np.uint64)
        self.ctx._ck(self.ctx.lib.kr_seqset_keys(self.h, i, out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value]

    def kmers(self, i: int):
        """SequenceKmers.iterator(): the distinct k-mers as strings"""
        return [self.ctx.key_decode(k) for k in self.keys(i)]

    def gather(self, idx) -> "SeqSet":
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        h = C.c_void_p()
        self.ctx._ck(self.ctx.lib.kr_seqset_gather(self.ctx.h, self.h, idx.ctypes.data, len(idx), C.byref(h)))
        return SeqSet(self.ctx, h)


class RepDb:
    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    def free(self):
        if self.h:
            self.ctx.lib.kr_db_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass

    @property
    def size(self):
        return int(self.ctx.lib.kr_db_size(self.h))

    def append(self, seqs=None, res=None, off=None):
        if res is None:
            res, off = flat(seqs)
        res = np.ascontiguousarray(res, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        pad = res if res.size else np.zeros(1, np.uint8)
        self.ctx._ck(self.ctx.lib.kr_db_append(self.ctx.h, self.h, pad.ctypes.data, off.ctypes.data, len(off) - 1))

    def append_set(self, s: SeqSet, idx):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        self.ctx._ck(self.ctx.lib.kr_db_append_set(self.ctx.h, self.h, s.h, idx.ctypes.data, len(idx)))

    def set_ranks(self, rank):
        rank = np.ascontiguousarray(rank, dtype=np.uint32)
        self.ctx._ck(self.ctx.lib.kr_db_set_ranks(self.ctx.h, self.h, rank.ctypes.data, len(rank)))

    def find_closest(self, q: SeqSet):
        sim = np.zeros(q.count, dtype=np.int32)
        first = np.zeros(q.count, dtype=np.uint32)
        last = np.zeros(q.count, dtype=np.uint32)
        self.ctx._ck(self.ctx.lib.kr_db_find_closest(self.ctx.h, self.h, q.h, sim.ctypes.data, first.ctypes.data, last.ctypes.data))
        return sim, first, last

    def find_closest_device(self, q: SeqSet, d_sim: int, d_first: int = None, d_last: int = None):
        self.ctx._ck(self.ctx.lib.kr_db_find_closest_device(self.ctx.h, self.h, q.h, d_sim, d_first, d_last))

    def any_ge(self, cands: SeqSet, threshold: int):
        out = np.zeros(cands.count, dtype=np.uint8)
        self.ctx._ck(self.ctx.lib.kr_db_any_ge(self.ctx.h, self.h, cands.h, threshold, out.ctypes.data))
        return out

    def greedy_build(self, cands: SeqSet, threshold: int, batch: int = 0):
        added = np.zeros(cands.count, dtype=np.uint8)
        self.ctx._ck(self.ctx.lib.kr_greedy_build(self.ctx.h, self.h, cands.h, threshold, batch, added.ctypes.data))
        return added


def unpack_cells(cells64):
    """packed cells (kr_sim_matrix_sparse64) -> (row, col, sim) uint32 arrays"""
    c = np.asarray(cells64, dtype=np.uint64)
    return (c >> np.uint64(40)).astype(np.uint32), ((c >> np.uint64(16)) & np.uint64(0xFFFFFF)).astype(np.uint32), (c & np.uint64(0xFFFF)).astype(np.uint32)


def _flat_args(seqs, res, off):
    if res is None:
        res, off = flat(seqs)
    res = np.ascontiguousarray(res, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    return (res if res.size else np.zeros(1, np.uint8)), off


class MultiContext:
    """Several GPUs behind one handle (kr_multi): host buffers in, host buffers out, sharding inside the library."""

    def __init__(self, devices, K=8, flags=IDENT_IS_INFINITY):
        self.lib = load_library()
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.kr_multi_create(devs, len(devices), K, flags, C.byref(h))
        if rc:
            raise KrError(rc, self.lib.kr_multi_last_error(None).decode())
        self.h, self.K, self.flags = h, K, flags
        self.ndev = int(self.lib.kr_multi_ndev(h))

    def _ck(self, rc):
        if rc:
            raise KrError(rc, self.lib.kr_multi_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.kr_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sim_matrix(self, seqs=None, res=None, off=None, b_seqs=None, b_res=None, b_off=None, elem=ELEM_I32, layout=LAYOUT_FULL):
        ra, oa = _flat_args(seqs, res, off)
        na = len(oa) - 1
        rect = b_seqs is not None or b_off is not None
        rb, ob = _flat_args(b_seqs, b_res, b_off) if rect else (None, None)
        nb = len(ob) - 1 if rect else 0
        ncols = nb if rect else na
        out = np.zeros(matrix_elems(na, ncols, layout), dtype=np.uint16 if elem == ELEM_U16 else np.int32)
        self._ck(self.lib.kr_multi_sim_matrix(self.h, ra.ctypes.data, oa.ctypes.data, na, _ptr(rb), _ptr(ob), nb, elem, layout, out.ctypes.data))
        return out.reshape(na, ncols) if layout == LAYOUT_FULL else out

    def sim_matrix_sparse(self, seqs=None, res=None, off=None, b_seqs=None, b_res=None, b_off=None, min_sim=1, layout=LAYOUT_TRIANGLE, cap=1 << 20):
        ra, oa = _flat_args(seqs, res, off)
        rect = b_seqs is not None or b_off is not None
        rb, ob = _flat_args(b_seqs, b_res, b_off) if rect else (None, None)
        while True:
            cells = np.zeros((cap, 4), dtype=np.uint32)
            n = C.c_uint64(0)
            rc = self.lib.kr_multi_sim_matrix_sparse(self.h, ra.ctypes.data, oa.ctypes.data, len(oa) - 1, _ptr(rb), _ptr(ob), len(ob) - 1 if rect else 0,
                                                     int(min_sim), layout, cells.ctypes.data, cap, C.byref(n))
            if rc == 0:
                return cells[: n.value]
            if n.value <= cap:
                self._ck(rc)
            cap = int(n.value)

    def db(self) -> "MultiDb":
        h = C.c_void_p()
        self._ck(self.lib.kr_multi_db_create(self.h, C.byref(h)))
        return MultiDb(self, h)

    def stats(self) -> dict:
        s = Stats()
        self._ck(self.lib.kr_multi_stats_get(self.h, C.byref(s)))
        return s.asdict()


class MultiDb:
    def __init__(self, m: MultiContext, h):
        self.m, self.h = m, h

    def free(self):
        if self.h:
            self.m.lib.kr_multi_db_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.m.h:
                self.free()
        except Exception:
            pass

    @property
    def size(self):
        return int(self.m.lib.kr_multi_db_size(self.h))

    def append(self, seqs=None, res=None, off=None):
        r, o = _flat_args(seqs, res, off)
        self.m._ck(self.m.lib.kr_multi_db_append(self.m.h, self.h, r.ctypes.data, o.ctypes.data, len(o) - 1))

    def set_ranks(self, rank):
        rank = np.ascontiguousarray(rank, dtype=np.uint32)
        self.m._ck(self.m.lib.kr_multi_db_set_ranks(self.m.h, self.h, rank.ctypes.data, len(rank)))

    def find_closest(self, seqs=None, res=None, off=None):
        r, o = _flat_args(seqs, res, off)
        n = len(o) - 1
        sim, first, last = np.zeros(n, np.int32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        self.m._ck(self.m.lib.kr_multi_db_find_closest(self.m.h, self.h, r.ctypes.data, o.ctypes.data, n, sim.ctypes.data, first.ctypes.data, last.ctypes.data))
        return sim, first, last

    def any_ge(self, threshold, seqs=None, res=None, off=None):
        r, o = _flat_args(seqs, res, off)
        out = np.zeros(len(o) - 1, dtype=np.uint8)
        self.m._ck(self.m.lib.kr_multi_db_any_ge(self.m.h, self.h, r.ctypes.data, o.ctypes.data, len(o) - 1, int(threshold), out.ctypes.data))
        return out

    def greedy_build(self, threshold, seqs=None, res=None, off=None, batch=0):
        r, o = _flat_args(seqs, res, off)
        added = np.zeros(len(o) - 1, dtype=np.uint8)
        self.m._ck(self.m.lib.kr_multi_greedy_build(self.m.h, self.h, r.ctypes.data, o.ctypes.data, len(o) - 1, int(threshold), int(batch), added.ctypes.data))
        return added
