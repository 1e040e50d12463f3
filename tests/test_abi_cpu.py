"""CPU tests of the boundary: the library loads, exports every symbol include/kmersreps.h declares,
refuses to run without a GPU (no CPU fallback), and its pure-host helpers agree with the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

import kmers.reps_b200 as kr
from kmers.reps_b200.binding import SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = kr.load_library()
    hdr = open(os.path.join(ROOT, "include", "kmersreps.h")).read()
    declared = set(re.findall(r"\b(kr_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"kr_status"}
    assert declared == set(SYMBOLS), declared ^ set(SYMBOLS)
    for s in SYMBOLS:
        assert hasattr(lib, s), s


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(kr.KrError) as ei:
        kr.Context(device=0, K=8)
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "kmers")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".c")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "kr_oracle" not in txt and "kro_" not in txt, os.path.join(dp, f)


def test_distance_matches_oracle(oracle):
    for (s, a, b) in [(0, 10, 20), (3, 41, 20), (9999, 5, 5), (49, 343, 350), (1, 1, 1), (221, 337, 316)]:
        assert kr.distance(s, a, b) == oracle.distance(s, a, b)


def test_matrix_layout_arithmetic():
    for n in (1, 255, 256, 257, 1000, 1480):
        assert kr.matrix_elems(n, n, kr.LAYOUT_TRIANGLE) == n * (n - 1) // 2
        assert kr.matrix_elems(n, 77, kr.LAYOUT_FULL) == n * 77
        bt = sum(n - 256 * (i // 256) for i in range(n))
        assert kr.matrix_elems(n, n, kr.LAYOUT_BLOCKTRI) == bt
        for layout in (0, 1, 2):
            for nsh in (1, 2, 3, 8):
                rb = kr.matrix_stripes(n, n, layout, nsh)
                assert rb[0] == 0 and rb[-1] == n and (np.diff(rb.astype(np.int64)) >= 0).all()
                assert all(int(x) % 256 == 0 or int(x) == n for x in rb)
                tot = sum(kr.matrix_elems(n, n, layout, int(rb[k]), int(rb[k + 1])) for k in range(nsh))
                assert tot == kr.matrix_elems(n, n, layout)
    # balance at the bench size: 8 stripes of the 100k block-triangle within 3% of each other
    rb = kr.matrix_stripes(100000, 100000, kr.LAYOUT_BLOCKTRI, 8)
    cells = [kr.matrix_elems(100000, 100000, kr.LAYOUT_BLOCKTRI, int(rb[k]), int(rb[k + 1])) for k in range(8)]
    assert max(cells) / min(cells) < 1.03


def test_synth_is_deterministic():
    from kmers.reps_b200 import synth
    a = synth.generate(123, 500, 10)
    b = synth.generate(123, 500, 10)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all()
    import hashlib
    assert hashlib.sha256(a[0].tobytes()).hexdigest()[:16] == hashlib.sha256(b[0].tobytes()).hexdigest()[:16]


def test_header_is_plain_c(tmp_path):
    """include/kmersreps.h is the drop-in boundary: it must compile as C99 (no C++ in the signatures), warnings as errors"""
    src = tmp_path / "t.c"
    src.write_text('#include "kmersreps.h"\nint main(void) { kr_multi *m = 0; kr_ctx *c = 0; (void)m; (void)c; return KR_CELL_SIM(5ull) == 5 ? 0 : 1; }\n')
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert p.returncode == 0, p.stdout.decode()
