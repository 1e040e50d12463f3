"""CPU check of the EXPERIMENTAL posting-list all-vs-all kernels (kmers/reps_b200/csrc/kr_pairs.cuh): they use only
global atomics, so tests/emul/pairs_emul.cpp executes them thread by thread with a small shim and compares every cell
of every layout / element type / row stripe with a brute-force count."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pairs_kernels_on_the_cpu_shim(tmp_path):
    exe = str(tmp_path / "pairs_emul")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fsanitize=address,undefined", "-o", exe, os.path.join(ROOT, "tests", "emul", "pairs_emul.cpp")],
                   check=True, timeout=300)
    for seed in ("1",):
        p = subprocess.run([exe, seed], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
        assert p.returncode == 0 and p.stdout.decode().startswith("OK"), p.stdout.decode()
