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
        subprocess.run(["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-DKR_EMUL",
                        "-I" + EMUL, "-Wno-unknown-pragmas", "-o", LIB, os.path.join(CSRC, "kr_api.cu")], check=True, timeout=600)
    return LIB


@pytest.mark.parametrize("order,case", [("reverse", "matrix_layouts"), ("reverse", "subchunks"), ("random", "find_closest_and_build"),
                                        ("reverse", "sparse_matrix"), ("random", "index_crowded"), ("reverse", "encode_warp")])
def test_device_code_under_other_thread_orders(emul_lib, order, case):
    """between synchronisation points an emulated thread runs undisturbed, so the order in which a CTA's threads take
    their turn decides which side of a missing barrier wins; last-to-first and shuffled orders must give the same results"""
    env = dict(os.environ, KR_LIB_PATH=emul_lib, KR_EMUL_SMS="2", KR_EMUL_ORDER=order)
    p = subprocess.run([sys.executable, os.path.join(EMUL, "run_emul_checks.py"), case], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=900)
    assert p.returncode == 0 and p.stdout.decode().strip().endswith("OK " + case), p.stdout.decode()[-3000:]


@pytest.mark.parametrize("seed,case", [("1", "index_crowded"), ("7", "find_closest_and_build"), ("3", "sparse_matrix")])
def test_device_code_with_preempted_atomics(emul_lib, seed, case):
    """lanes of a warp run concurrently on hardware; KR_EMUL_PREEMPT suspends a thread pseudo-randomly just before an atomic
    executes, so the lock-free score table of the index scan (read slot, compare-and-swap, count) and the pair counters are
    exercised with other lanes getting in between"""
    env = dict(os.environ, KR_LIB_PATH=emul_lib, KR_EMUL_SMS="2", KR_EMUL_PREEMPT=seed)
    p = subprocess.run([sys.executable, os.path.join(EMUL, "run_emul_checks.py"), case], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=900)
    assert p.returncode == 0 and p.stdout.decode().strip().endswith("OK " + case), p.stdout.decode()[-3000:]


def build_variant(name, *defines):
    out = os.path.join(EMUL, "_build", name)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(EMUL, "cuda_emul.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-DKR_EMUL", *defines,
                        "-I" + EMUL, "-Wno-unknown-pragmas", "-o", out, os.path.join(CSRC, "kr_api.cu")], check=True, timeout=600)
    return out


def test_index_overflow_passes_on_the_emulator(emul_lib):
    """the index scan with a 64-slot score table (-DKR_IXQ_H=64): nearly every query touches more ids than fit, so results
    come from the 4 / 16 / 64-pass path (the exactness path a real run rarely takes), for ARGMAX, sparse and dense cells"""
    lib = build_variant("libkmersreps_emul_h64.so", "-DKR_IXQ_H=64")
    for case in ("find_closest_and_build", "sparse_matrix", "matrix_layouts"):
        env = dict(os.environ, KR_LIB_PATH=lib, KR_EMUL_SMS="2")
        p = subprocess.run([sys.executable, os.path.join(EMUL, "run_emul_checks.py"), case], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           timeout=1500)
        assert p.returncode == 0 and p.stdout.decode().strip().endswith("OK " + case), p.stdout.decode()[-3000:]


def test_pipelined_host_encode_on_the_emulator(emul_lib):
    """kr_encode of a host batch in four pieces (copy stream + per-piece encoder launches): a build whose threshold is 256 bytes
    takes that path for every batch of the ordinary cases -- sizes, keys, the alphabet error's sequence number, matrices"""
    lib = build_variant("libkmersreps_emul_pipe.so", "-DKR_ENC_PIPE_MIN=256")
    for case in ("encode_warp", "api_misc", "matrix_layouts"):
        env = dict(os.environ, KR_LIB_PATH=lib, KR_EMUL_SMS="2")
        p = subprocess.run([sys.executable, os.path.join(EMUL, "run_emul_checks.py"), case], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           timeout=1500)
        assert p.returncode == 0 and p.stdout.decode().strip().endswith("OK " + case), p.stdout.decode()[-3000:]


@pytest.mark.parametrize("case", ["matrix_layouts", "k10_rectangular_pairs", "find_closest_and_build", "subchunks", "sharded_import", "api_misc",
                                  "sparse_matrix", "index_crowded", "encode_warp", "set_ops"])
def test_device_code_on_the_emulator(emul_lib, case):
    env = dict(os.environ, KR_LIB_PATH=emul_lib, KR_EMUL_SMS="2")
    p = subprocess.run([sys.executable, os.path.join(EMUL, "run_emul_checks.py"), case], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=900)
    assert p.returncode == 0 and p.stdout.decode().strip().endswith("OK " + case), p.stdout.decode()[-3000:]


@pytest.mark.parametrize("devices", ["3", "1"])
def test_multi_gpu_handle_on_the_emulator(emul_lib, devices):
    """kr_multi: three emulated devices, and one device listed three times; one worker thread per context"""
    env = dict(os.environ, KR_LIB_PATH=emul_lib, KR_EMUL_SMS="2", KR_EMUL_DEVICES=devices)
    p = subprocess.run([sys.executable, os.path.join(EMUL, "run_emul_checks.py"), "multi"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=1500)
    assert p.returncode == 0 and p.stdout.decode().strip().endswith("OK multi"), p.stdout.decode()[-3000:]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_paths_end_to_end_on_the_emulator(emul_lib, world):
    """gloo ranks, one emulator context each: the block-list exchange (all-gather + import) of the tile-join path, and the
    default scheme (every rank encodes and indexes everything, scans its row stripe) through the join and the index, with
    the dense stripes and the sparse cells all-gathered"""
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


def test_smoke_entry_point_on_the_emulator(emul_lib):
    """__graft_entry__.smoke() -- what the driver runs on the GPU box before the bench -- as a logic check"""
    env = dict(os.environ, KR_LIB_PATH=emul_lib, KR_EMUL_SMS="2")
    p = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], env=env, cwd=ROOT, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, timeout=900)
    assert p.returncode == 0 and "smoke ok" in p.stdout.decode(), p.stdout.decode()[-2000:]
