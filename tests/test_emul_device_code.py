"""CPU execution of the product's device and host code on the functional CUDA emulator in tests/emul/ (fibers for
threads, mbarrier / bulk-copy / shuffle semantics, cp.async.bulk alignment rules checked).  The sources under
kmers/reps_b200/csrc are compiled unchanged by g++ with -DKR_EMUL into tests/emul/_build/libkmersreps_emul.so, and
tests/emul/run_emul_checks.py drives them through the ctypes binding against the oracle, one subprocess per case.

This is a logic check for a container without a GPU (it found nothing the B200 runs had not, and it is what vouches
for code written after GPU access closed); it is test infrastructure, never a product path: the product library is
still the nvcc build, and kr_create still fails without a device."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")
LIB = os.path.join(EMUL, "_build", "libkmersreps_emul.so")
CSRC = os.path.join(ROOT, "kmers", "reps_b200", "csrc")


@pytest.fixture(scope="module")
def emul_lib():
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(EMUL, "cuda_emul.h"), os.path.join(ROOT, "include", "kmersreps.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-DKR_EMUL", "-DKR_EXPERIMENTAL",
                        "-I" + EMUL, "-Wno-unknown-pragmas", "-o", LIB, os.path.join(CSRC, "kr_api.cu")], check=True, timeout=600)
    return LIB


@pytest.mark.parametrize("order,case", [("reverse", "matrix_layouts"), ("reverse", "subchunks"), ("random", "find_closest_and_build"),
                                        ("reverse", "exp_pairs"), ("random", "exp_idxq_warp"), ("reverse", "exp_encode_warp")])
def test_device_code_under_other_thread_orders(emul_lib, order, case):
    """between synchronisation points an emulated thread runs undisturbed, so the order in which a CTA's threads take
    their turn decides which side of a missing barrier wins; last-to-first and shuffled orders must give the same results"""
    env = dict(os.environ, KR_LIB_PATH=emul_lib, KR_EMUL_SMS="2", KR_EMUL_ORDER=order)
    p = subprocess.run([sys.executable, os.path.join(EMUL, "run_emul_checks.py"), case], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=900)
    assert p.returncode == 0 and p.stdout.decode().strip().endswith("OK " + case), p.stdout.decode()[-3000:]


def build_variant(name, *defines):
    out = os.path.join(EMUL, "_build", name)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(EMUL, "cuda_emul.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-DKR_EMUL", "-DKR_EXPERIMENTAL", *defines,
                        "-I" + EMUL, "-Wno-unknown-pragmas", "-o", out, os.path.join(CSRC, "kr_api.cu")], check=True, timeout=600)
    return out


def test_bucket_grouping_sub_passes_on_the_emulator(emul_lib):
    """the bucketed posting-list path with a 256-slot shared-memory table: every bucket holds more distinct keys than
    fit, so each one is grouped in several sub-passes (the exactness path a real run almost never takes)"""
    lib = build_variant("libkmersreps_emul_bk256.so", "-DKR_BK_SLOTS=256")
    env = dict(os.environ, KR_LIB_PATH=lib, KR_EMUL_SMS="2", KR_EMUL_ONLY_BUCKETS="1")
    p = subprocess.run([sys.executable, os.path.join(EMUL, "run_emul_checks.py"), "exp_pairs"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=900)
    assert p.returncode == 0 and p.stdout.decode().strip().endswith("OK exp_pairs"), p.stdout.decode()[-3000:]


@pytest.mark.parametrize("case", ["matrix_layouts", "k10_rectangular_pairs", "find_closest_and_build", "subchunks", "sharded_import", "api_misc",
                                  "exp_sparse_join", "exp_pairs", "exp_encode_warp", "exp_idxq_warp", "exp_gidx_build", "exp_dna"])
def test_device_code_on_the_emulator(emul_lib, case):
    env = dict(os.environ, KR_LIB_PATH=emul_lib, KR_EMUL_SMS="2")
    p = subprocess.run([sys.executable, os.path.join(EMUL, "run_emul_checks.py"), case], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=900)
    assert p.returncode == 0 and p.stdout.decode().strip().endswith("OK " + case), p.stdout.decode()[-3000:]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_paths_end_to_end_on_the_emulator(emul_lib, world):
    """gloo ranks, one emulator context each, run sharding.encode_sharded (block-list all-gather + import) and
    sharding.encode_sharded_keys (key-list all-gather + import, posting-list path) and compute their result stripes"""
    env = dict(os.environ, KR_LIB_PATH=emul_lib, KR_EMUL_SMS="2")
    p = subprocess.run([sys.executable, os.path.join(EMUL, "run_emul_sharded.py"), str(world)], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, timeout=1500)
    assert p.returncode == 0 and p.stdout.decode().strip().endswith("OK sharded %d" % world), p.stdout.decode()[-3000:]


def test_host_tools_on_the_emulator(emul_lib, tmp_path):
    """the C++ host layer and its commands (tests/test_host_tools.py, byte-level output checks) with the emulator build
    standing in for libkmersreps_cuda.so: the same test file the GPU box runs, here as a logic check of host + device code"""
    d = tmp_path / "as_cuda"
    d.mkdir()
    os.symlink(emul_lib, d / "libkmersreps_cuda.so")
    env = dict(os.environ, KR_LIB_PATH=str(d / "libkmersreps_cuda.so"), LD_LIBRARY_PATH=str(d), KR_EMUL_SMS="2")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_host_tools.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "neighbors or classify or distances or group or error_behaviour or load_report or gtoclass or gtoreps"], env=env, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1200)
    assert p.returncode == 0, p.stdout.decode()[-3000:]


def test_seqcomp_on_the_emulator(emul_lib, tmp_path):
    """the seqcomp command (protein + DNA k-mer distances; DnaKmers needs the experimental encoder) end to end"""
    d = tmp_path / "as_cuda"
    d.mkdir()
    os.symlink(emul_lib, d / "libkmersreps_cuda.so")
    env = dict(os.environ, KR_LIB_PATH=str(d / "libkmersreps_cuda.so"), LD_LIBRARY_PATH=str(d), KR_EMUL_SMS="2")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_zzz_experimental.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "seqcomp or ssureps"], env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1200)
    assert p.returncode == 0 and " passed" in p.stdout.decode(), p.stdout.decode()[-3000:]


def test_v2_selfcheck_switches_off_a_wrong_kernel(emul_lib, tmp_path):
    """-DKR_DEFAULTS_V2 makes the experimental paths every context's default, guarded by a start-up self-check that runs each
    of them once against the path it replaces (GPU result against GPU result).  Built here with -DKR_FAULT_INJECT (the
    posting-list kernel drops some pairs on purpose): the check must say so, switch that path off, and results stay exact;
    the same build without the fault stays silent."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import numpy as np, kmers.reps_b200 as kr\n"
            "from kmers.reps_b200 import synth\n"
            "from oracle import kr_oracle as ko\n"
            "c = kr.Context(device=0, K=8, flags=kr.IDENT_IS_INFINITY)\n"
            "res, off, _ = synth.generate(5, 700, 140)\n"          # sparse similarity: the AUTO dispatch would pick the posting lists
            "s = c.encode(res=res, off=off)\n"
            "want = ko.get().sim_matrix(ko.get().batch(res=res, off=off, K=8), ident_inf=True, threads=4)\n"
            "assert (c.sim_matrix(s, elem=kr.ELEM_U16) == want).all()\n"
            "print('RESULT OK')\n") % ROOT
    for fault, expect_msg in ((True, True), (False, False)):
        lib = build_variant("libkmersreps_emul_v2%s.so" % ("_fault" if fault else ""), "-DKR_DEFAULTS_V2", *(["-DKR_FAULT_INJECT"] if fault else []))
        env = dict(os.environ, KR_LIB_PATH=lib, KR_EMUL_SMS="2", KR_V2_SELFCHECK_N="210")
        env.pop("KR_V2_SELFCHECK", None)
        p = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
        assert p.returncode == 0 and "RESULT OK" in p.stdout.decode(), p.stderr.decode()[-3000:]
        assert ("posting-list all-vs-all path disagrees" in p.stderr.decode()) == expect_msg, p.stderr.decode()[-2000:]


def test_smoke_entry_point_on_the_emulator(emul_lib):
    """__graft_entry__.smoke() -- what the driver runs on the GPU box before the bench -- as a logic check"""
    env = dict(os.environ, KR_LIB_PATH=emul_lib, KR_EMUL_SMS="2")
    p = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], env=env, cwd=ROOT, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, timeout=900)
    assert p.returncode == 0 and "smoke ok" in p.stdout.decode(), p.stdout.decode()[-2000:]
