#!/usr/bin/env python
"""TEST INFRASTRUCTURE, CPU only: how sensitive is the race / stream-order detector (tools/emul_race.sh) on the PRODUCT's
own code?  Each mutation below removes one synchronisation from a scratch copy of the sources -- the emulated run still
computes the right values (most missing barriers do not get that far: they change the emulated result and the functional
tests catch them) -- and the detector must report it.  Nothing in the repository is modified.

    python tools/emul_race_seeded.py          (about two minutes; exit 0 = every seeded bug was reported)
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-x", "c++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "--param", "tsan-distinguish-volatile=1", "-pthread", "-DKR_EMUL",
         "-DKR_EMUL_TSAN", "-DKR_ENC_PIPE_MIN=256", "-Itests/emul", "-Wno-unknown-pragmas"]

MUTATIONS = [
    # (what, file, text to remove, replacement, identity granularity)
    ("K1 counting sort: no __syncwarp between the scatter into bucket order and the insertion sort",
     "kmers/reps_b200/csrc/kr_encode.cuh",
     "    __syncwarp();\n    // buckets with more than one key: insertion sort by the lane that owns the bucket",
     "    // buckets with more than one key: insertion sort by the lane that owns the bucket", "thread"),
    ("dense matrix: the scan no longer waits for the zero-fill on the second stream",
     "kmers/reps_b200/csrc/kr_api.cu",
     "        if (filled) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_fill[1], 0));", "", "warp"),
    ("pipelined host encode: the encoder of a piece does not wait for the piece's copy on the copy stream",
     "kmers/reps_b200/csrc/kr_api.cu",
     "if (piped) { CU(cudaStreamWaitEvent(ctx->stream, ev_piece[c], 0)); ctx->ev_pool.push_back(ev_piece[c]); }",
     "if (piped) { ctx->ev_pool.push_back(ev_piece[c]); }", "warp"),
]


def main():
    failed = 0
    for what, rel, old, new, ident in MUTATIONS:
        with tempfile.TemporaryDirectory() as tmp:
            for d in ("kmers/reps_b200/csrc", "include", "tests/emul"):
                shutil.copytree(os.path.join(ROOT, d), os.path.join(tmp, d), ignore=shutil.ignore_patterns("_build"))
            path = os.path.join(tmp, rel)
            src = open(path).read()
            assert src.count(old) == 1, "mutation target not found exactly once: " + what
            open(path, "w").write(src.replace(old, new))
            extra = ["--param", "tsan-instrument-func-entry-exit=0"] if ident == "warp" else []
            binary = os.path.join(tmp, "race_check")
            subprocess.check_call(["g++"] + FLAGS + extra + ["-o", binary, "tests/emul/race_check.cpp"], cwd=tmp)
            env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66 history_size=4", KR_EMUL_SMS="2")
            if ident == "warp":
                env["KR_EMUL_RACE_ID"] = "warp"
            p = subprocess.run([binary, "smoke"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
            n = p.stdout.decode().count("WARNING: ThreadSanitizer")
            ok = p.returncode == 66 and n > 0
            print("%s: %s (%s identities, exit %d, %d report(s))" % ("reported" if ok else "NOT REPORTED", what, ident, p.returncode, n), flush=True)
            failed += not ok
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
