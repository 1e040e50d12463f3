#!/usr/bin/env python
"""Static evidence for the built libkmersreps_cuda.so (no GPU needed): cuobjdump -res-usage and the SASS mnemonic mix of
every kernel, written to profiles/r02_static.txt with the commit and a hash of the kernel sources in the header.

    python tools/static_evidence.py"""
import collections
import hashlib
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kmers", "reps_b200", "libkmersreps_cuda.so")
CSRC = os.path.join(ROOT, "kmers", "reps_b200", "csrc")
KEYS = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "RED", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "WARPSYNC", "VOTE", "IMAD", "LOP3", "ISETP"]


def demangle(name):
    d = subprocess.check_output(["c++filt", name]).decode().strip()
    return re.sub(r"\(.*\)$", "", d).replace("void ", "")


def main():
    rev = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT).decode().strip()
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        h.update(open(os.path.join(CSRC, f), "rb").read())
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT).stdout.decode()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT).stdout.decode()
    out = ["# Static evidence for kmers/reps_b200/libkmersreps_cuda.so (cuobjdump of the sm_100a cubin; no GPU involved)",
           "# git HEAD when generated: %s; sha256 of kmers/reps_b200/csrc/*: %s" % (rev, h.hexdigest()[:16]),
           "# cubin arch: " + ", ".join(sorted(set(re.findall(r"arch = (\S+)", sass)))), "",
           "## registers / stack / static shared memory / local memory per kernel (cuobjdump -res-usage)"]
    for m in re.finditer(r"Function (\S+):\n\s*(REG:\d+ STACK:\d+ SHARED:\d+ LOCAL:\d+)", res):
        out.append("%-44s %s" % (demangle(m.group(1)), m.group(2)))
    out += ["", "## SASS instruction count and mnemonic mix per kernel (cuobjdump -sass)",
            "# UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier ops, ATOMS = shared-memory atomics, ATOMG/RED = global",
            "# atomics, SHFL = warp shuffles, LDS/STS = shared loads/stores, LDG/STG = global loads/stores, LDL/STL = local memory,",
            "# BAR = CTA barriers, WARPSYNC = __syncwarp"]
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n", 1)[0].strip()
        ins = re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", f, flags=re.M)
        c = collections.Counter(i.split(".")[0] for i in ins)
        out.append("%-44s total %6d  " % (demangle(name), len(ins)) + " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k]))
    path = os.path.join(ROOT, "profiles", "r02_static.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
