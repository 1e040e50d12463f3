"""CPU tests of the C++ host layer's pure-host parts (kmers/reps_b200/host/selftest makes no GPU call):
java.util.HashMap order emulation, Java number formatting, FASTA I/O and the RepGenomeDb load->save round trip,
each against the oracle's independent implementation or the reference's own data file."""
import os
import subprocess
from decimal import Decimal

import numpy as np
import pytest

from oracle import kr_oracle as ko

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "kmers", "reps_b200", "host", "selftest")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(TOOL):
        import __graft_entry__ as g
        g.build()
    assert os.path.exists(TOOL)


def run(args, text="", check=True):
    p = subprocess.run([TOOL] + args, input=text.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    if check:
        assert p.returncode == 0, p.stderr.decode()
    return p


def test_hashmap_order_matches_oracle_and_rep50(rep50, oracle):
    ids = [ko.genome_id_of(f) for (f, _, _) in rep50["reps"]]
    got = [int(x) for x in run(["order"], "\n".join(ids) + "\n").stdout.split()]
    assert got == list(range(len(ids)))                    # file order is iteration order: a fixed point
    rng = np.random.default_rng(21)
    for trial in range(6):
        n = int(rng.integers(1, 5000))
        keys = ["%d.%d" % (rng.integers(1, 3000000), rng.integers(1, 400)) for _ in range(n)]
        keys += keys[: n // 10]                            # re-put of existing keys keeps their position
        want = list(oracle.java_hashmap_order(keys))
        got = [int(x) for x in run(["order"], "\n".join(keys) + "\n").stdout.split()]
        assert got == want
        assert got == ko.py_java_hashmap_order(keys)


def order_check(script, every=False):
    out = run(["order-remove"] + (["every"] if every else []), "\n".join(script) + "\n").stdout.decode().split("\n")
    tail = dict(kv.split("=") for kv in out[-2].split())
    return [int(x) for x in out[:-2]], int(tail["inv"]), int(tail["treebins"])


def colliding_ids(nbits, groups, per_group, seed):
    """genome-id-like keys whose spread hashes agree in their low `nbits` bits: they share a bin until the table
    outgrows 2**nbits slots"""
    rng = np.random.default_rng(seed)
    by_low, out = {}, []
    while len(out) < groups:
        k = "%d.%d" % (rng.integers(1, 4000000), rng.integers(1, 500))
        low = ko.JavaHashMap.spread(k) & ((1 << nbits) - 1)
        g = by_low.setdefault(low, [])
        if k not in g:
            g.append(k)
            if len(g) == per_group:
                out.append(g)
    return out


def test_hashmap_tree_bins_same_hash_keys(oracle):
    # 32 keys with ONE String.hashCode ("Aa" and "BB" hash alike): the 9th to arrive at a 64-slot table turns the
    # bin into a red-black tree ordered by String.compareTo; the C++ host emulation and the oracle's Python
    # restatement of OpenJDK's HashMap must agree on the iteration order at every size
    import itertools
    keys = ["".join(p) for p in itertools.product(("Aa", "BB"), repeat=5)]
    assert len({ko.java_string_hash(k) for k in keys}) == 1
    rng = np.random.default_rng(5)
    for trial in range(8):
        ks = [keys[i] for i in rng.permutation(len(keys))]
        fill = ["f%d" % i for i in range(int(rng.integers(0, 300)))]
        seq = [k for pair in itertools.zip_longest(ks, fill) for k in pair if k is not None]
        for n in (9, 11, 12, 13, 20, len(seq)):
            got, inv, bins = order_check(seq[:n], every=True)
            assert inv == 1
            assert got == ko.py_java_hashmap_order(seq[:n])
            assert got == list(oracle.java_hashmap_order(seq[:n]))
        assert bins == 1
    # the order is NOT insertion order inside the bin (the tree root moves to the front)
    got, _, _ = order_check(keys[:12])
    chain = [i for i in got if i < 12]
    assert sorted(chain) == list(range(12)) and chain != list(range(12))


def test_hashmap_tree_bins_split_and_untreeify():
    # genome-id keys sharing the low 10 bits of their spread hash: trees form at 64..1024 slots, then every further
    # doubling splits them (TreeNode.split), re-treeifying halves of > 6 nodes and flattening the smaller ones
    groups = colliding_ids(10, 12, 40, seed=77)
    rng = np.random.default_rng(78)
    filler = ["%d.%d" % (rng.integers(1, 4000000), rng.integers(1, 500)) for _ in range(7000)]
    seq, gi = [], [0] * len(groups)
    for i, f in enumerate(filler):
        seq.append(f)
        g = i % len(groups)
        if i % 3 == 0 and gi[g] < 40:
            seq.append(groups[g][gi[g]])
            gi[g] += 1
    seen_trees = 0
    for n in (100, 300, 700, 900, 1500, 1700, 3200, 6500, len(seq)):
        got, inv, bins = order_check(seq[:n])
        assert inv == 1
        assert got == ko.py_java_hashmap_order(seq[:n])
        seen_trees = max(seen_trees, bins)
    assert seen_trees >= 6


def test_hashmap_tree_bins_removal():
    # Iterator.remove() ('-', UpdateProcessor.java:96-103: no re-rooting, no untreeify) and Map.remove ('~') on
    # tree bins, interleaved with more puts
    groups = colliding_ids(12, 6, 30, seed=91)
    rng = np.random.default_rng(92)
    script, removals, live, keys = [], {}, [], []
    pool = [k for g in groups for k in g] + ["%d.%d" % (rng.integers(1, 4000000), rng.integers(1, 500)) for _ in range(400)]
    pool = [pool[i] for i in rng.permutation(len(pool))]
    for k in dict.fromkeys(pool):
        keys.append(k)
        script.append(k)
        live.append(k)
        if len(live) > 40 and rng.random() < 0.35:
            victim = live.pop(int(rng.integers(0, len(live))))
            movable = bool(rng.random() < 0.5)
            script.append(("~" if movable else "-") + victim)
            removals.setdefault(len(keys) - 1, []).append((victim, movable))
    got, inv, bins = order_check(script, every=True)
    assert inv == 1
    assert got == ko.py_java_hashmap_order(keys, removals=removals)
    assert sorted(keys[i] for i in got) == sorted(live)


def test_hashmap_remove_keeps_order():
    keys = ["k%d" % i for i in range(200)]
    script = keys + ["-k%d" % i for i in range(0, 200, 3)]
    got, inv, _ = order_check(script)
    assert inv == 1
    full = ko.py_java_hashmap_order(keys)
    assert got == [i for i in full if i % 3 != 0]


def py_double_to_string(x: float) -> str:
    if x == 0:
        return "0.0"
    d = Decimal(repr(x))
    sign, digits, exp = d.as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:
        digits.pop()
        exp += 1
    e10 = len(digits) - 1 + exp
    s = "-" if sign else ""
    if 1e-3 <= abs(x) < 1e7:
        t = format(d, "f")
        if "." not in t:
            t += ".0"
        elif t.endswith("0") and not t.endswith(".0"):
            t = t.rstrip("0")
            if t.endswith("."):
                t += "0"
        return t
    return s + str(digits[0]) + "." + ("".join(map(str, digits[1:])) or "0") + "E" + str(e10)


def test_java_number_formatting():
    rng = np.random.default_rng(5)
    vals = list(rng.random(3000)) + [0.125, 0.375, 0.0005, 0.9999995, 1.0, 0.0, 0.5110619469026549, 0.06470588235294117,
                                     2.5, 0.045, 0.005, 0.015, 0.025, 1e-9, 123456.789]
    text = "\n".join(repr(float(v)) for v in vals) + "\n"
    for (w, p) in ((8, 6), (8, 4), (12, 8), (4, 2), (1, 0)):
        got = run(["fixed", str(w), str(p)], text).stdout.decode().split("\n")[:-1]
        assert got == [ko.java_fixed(float(v), w, p) for v in vals], (w, p)
    dv = vals + [1e7, 9999999.0, 1.2345678e7, 1e-3, 9.99e-4, 2.5e-5, 1e21, 343.7, 100.0, 1234567.125]
    got = run(["dtos"], "\n".join(repr(float(v)) for v in dv) + "\n").stdout.decode().split("\n")[:-1]
    assert got == [py_double_to_string(float(v)) for v in dv]
    assert py_double_to_string(1.2345678e7) == "1.2345678E7" and py_double_to_string(2.5e-5) == "2.5E-5"


def test_repdb_load_save_is_byte_identical(tmp_path, rep50):
    src, out = tmp_path / "rep50.db", tmp_path / "resaved.db"
    src.write_bytes(rep50["raw"])
    p = run(["resave", str(src), str(out)])
    assert out.read_bytes() == rep50["raw"]
    assert p.stdout.decode().strip() == "1480 50 8 Phenylalanyl-tRNA synthetase alpha chain"
    # a database whose records arrive in another order is re-ordered exactly as java.util.HashMap would
    rng = np.random.default_rng(3)
    shuffled = [rep50["reps"][i] for i in rng.permutation(1480)]
    src.write_bytes(ko.save_repdb(50, 8, rep50["prot_name"], shuffled))
    run(["resave", str(src), str(out)])
    ids = [ko.genome_id_of(f) for (f, _, _) in shuffled]
    want = ko.save_repdb(50, 8, rep50["prot_name"], [shuffled[i] for i in ko.py_java_hashmap_order(ids)])
    assert out.read_bytes() == want
    bad = tmp_path / "bad.db"
    bad.write_bytes(b">NotARepDb\n>fig|1.1.peg.1 x\nACD\n")
    p = run(["resave", str(bad), str(out)], check=False)
    assert p.returncode == 1 and "invalid rep-genome header" in p.stderr.decode()


def test_fasta_reader_and_genome_ids(vapb, rep50):
    from tests.conftest import golden_bytes
    got = [l.split("\t") for l in run(["fasta"], golden_bytes("VapbProt.fa").decode()).stdout.decode().split("\n")[:-1]]
    assert [tuple(g) for g in got] == [tuple(r) for r in vapb]
    # CRLF input, blank lines, a header without sequence, trailing record without newline
    txt = ">a first one\r\nACD\r\n\r\nEFG\r\n>empty\r\n>b\r\nKLM"
    got = run(["fasta"], txt).stdout.decode().split("\n")[:-1]
    assert got == ["a\tfirst one\tACDEFG", "empty\t\t", "b\t\tKLM"]
    fids = [f for (f, _, _) in rep50["reps"][:50]] + ["fig|83333.1.peg.4", "notafig", "fig|12.peg.1"]
    got = run(["genomeof"], "\n".join(fids) + "\n").stdout.decode().split("\n")[:-1]
    assert got == [ko.genome_id_of(f) for f in fids]


def test_gto_reader_and_seed_lookup(gto_seeds):
    """genome.hpp: the JSON reader, the GTO fields and GenomeDescriptor.findSeed's rule on the reduced GTOs of tests/golden
    (escapes, a eukaryote, a genome without seed, a shorter decoy with EC number + comment) -- and, where the reference tree
    is present (this container, not the GPU box), on the ten bundled full-size GTOs against the golden seed proteins"""
    gdir = os.path.join(ROOT, "tests", "golden", "gto_mini")
    files = sorted(os.path.join(gdir, f) for f in os.listdir(gdir))
    rows = [ln.split("\t") for ln in run(["seed"] + files).stdout.decode().split("\n")[:-1]]
    by_id = {r[0]: r for r in rows}
    seeds = {ko.genome_id_of(f): (f, n, s) for (f, n, s) in gto_seeds}
    for g in ("202462.4", "1001994.6", "1408469.3"):
        assert by_id[g][2:] == [seeds[g][0], seeds[g][1], seeds[g][2]]
    assert by_id["1001994.6"][1] == "Archaea" and by_id["4932.99"][1] == "Eukaryota"
    assert by_id["4932.99"][3] == "Saccharomyces test genome \u00e9"            # \u escape -> UTF-8
    assert by_id["99999.1"][2] == "" and by_id["99999.1"][3] == 'Genome without "seed" protein'
    ref = "/root/reference/data"
    if os.path.isdir(ref):
        full = [os.path.join(ref, "202462.4.gto")] + sorted(os.path.join(ref, "P3Master", f) for f in os.listdir(os.path.join(ref, "P3Master")) if f.endswith(".gto"))
        rows = [ln.split("\t") for ln in run(["seed"] + full).stdout.decode().split("\n")[:-1]]
        assert [(r[2], r[3], r[4]) for r in rows] == [tuple(x) for x in gto_seeds]


def test_repgenome_id_order_policy(rep50):
    """RepGenome.compareTo lives in an un-vendored module: lexical vs natural order of genome ids is a run-time policy
    (KR_ID_ORDER / RepGenome::setIdOrder), checked against the oracle's independent restatement of both orders."""
    from oracle import ref_commands as rc
    ids = [ko.genome_id_of(f) for (f, _, _) in rep50["reps"][:120]] + ["83333.1", "511145.12", "9.10", "10.9", "10.10", "1.2.3", "1.2", "x.1", "1", "01.1", "1.1"]
    rng = np.random.default_rng(7)
    pairs = [(ids[a], ids[b]) for a, b in rng.integers(0, len(ids), (600, 2))] + [(i, i) for i in ids[-11:]]
    got = [tuple(int(v) for v in l.split("\t")) for l in run(["idcmp"], "".join("%s\t%s\n" % p for p in pairs)).stdout.decode().split("\n")[:-1]]
    want = [(rc.compare_ids(a, b, "lexical"), rc.compare_ids(a, b, "natural")) for a, b in pairs]
    assert got == want
    assert any(l != n for l, n in got), "the two orders never differed on the sample"
    # both are strict total orders on distinct ids: antisymmetric, and 0 only for equal strings
    for (a, b), (l, n) in zip(pairs, got):
        assert (l == 0) == (a == b) and (n == 0) == (a == b)
        assert (rc.compare_ids(b, a, "natural"), rc.compare_ids(b, a, "lexical")) == (-n, -l)
