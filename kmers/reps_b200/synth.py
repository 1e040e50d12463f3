"""ctypes wrapper of synth.c: the synthetic workloads of BASELINE.json configs 3-5."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        p = os.path.join(HERE, "libkrsynth.so")
        if not os.path.exists(p):
            raise RuntimeError(p + " is missing: run __graft_entry__.build()")
        _LIB = C.CDLL(p)
        _LIB.krs_capacity.restype = C.c_uint64
        _LIB.krs_capacity.argtypes = [C.c_uint32]
        _LIB.krs_generate.restype = C.c_uint64
        _LIB.krs_generate.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64,
                                      C.c_void_p, C.c_void_p]
    return _LIB


def generate(seed: int, nseq: int, nfam: int, dup_per_10k: int = 50, xrun_per_10k: int = 0):
    """-> (residues uint8[total], offsets uint64[nseq+1], family uint32[nseq])"""
    lib = _lib()
    cap = int(lib.krs_capacity(nseq))
    res = np.empty(max(cap, 1), dtype=np.uint8)
    off = np.zeros(nseq + 1, dtype=np.uint64)
    fam = np.zeros(max(nseq, 1), dtype=np.uint32)
    total = lib.krs_generate(seed, nseq, nfam, dup_per_10k, xrun_per_10k, res.ctypes.data, cap, off.ctypes.data, fam.ctypes.data)
    if nseq and total == 0:
        raise RuntimeError("krs_generate: capacity too small")
    return res[:total].copy(), off, fam[:nseq]


def to_strings(res, off):
    b = res.tobytes()
    return [b[int(off[i]):int(off[i + 1])].decode("ascii") for i in range(len(off) - 1)]
