"""kmers.reps_b200 -- B200 (sm_100a) implementation of the protein k-mer similarity hot path of
SEEDtk/kmers.reps, behind the C ABI declared in include/kmersreps.h.

This Python package is only the ctypes binding used by tests/ and bench.py; the product is
libkmersreps_cuda.so (csrc/) plus the C++ host mirror of the reference's Java classes (host/).
There is no CPU fallback: importing works anywhere, but creating a Context without the built
library or without a CUDA device raises.
"""
from .binding import (  # noqa: F401
    INFINITY, NO_REP, IDENT_IS_INFINITY, TIE_LAST, FORCE_JOIN, FORCE_INDEX, ELEM_I32, ELEM_U16, LAYOUT_FULL, LAYOUT_TRIANGLE,
    LAYOUT_BLOCKTRI, TILE, Context, SeqSet, RepDb, KrError, BlockLists, lib_path, load_library, distance,
    matrix_elems, matrix_stripes, flat, MultiContext, MultiDb, unpack_cells,
)
