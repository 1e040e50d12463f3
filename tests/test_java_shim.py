"""JDK-free structural check of java/ (tools/check_java_shim.py): every use the reference's sources make of the classes
the shim replaces resolves to a declared member of the shim, and every FFM handle matches include/kmersreps.h."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import check_java_shim as cj  # noqa: E402

REF = "/root/reference"


def uses():
    if os.path.isdir(os.path.join(REF, "src")):
        return cj.reference_uses(REF)
    return json.load(open(cj.GOLDEN))


def test_every_reference_use_resolves_in_the_shim():
    u = uses()
    assert len(u) > 100 and {x["cls"] for x in u} == set(cj.CLASSES)
    missing = cj.check(u)
    assert not missing, missing
    # the members VERDICT r01 listed as absent
    members = cj.shim_members()
    assert 3 in members["Representation"]["<init>"] and 2 in members["Representation"]["<init>"] and 0 in members["Representation"]["<init>"]
    assert 1 in members["RepGenomeDb"]["getSeedProtein"]
    for m in ("iterator", "retainAll", "removeAll", "equals", "hashCode", "size", "similarity", "getLen"):
        assert m in members["SequenceKmers"], m
    src = open(os.path.join(cj.JAVA, cj.CLASSES["SequenceKmers"])).read() + open(os.path.join(cj.JAVA, cj.CLASSES["ProteinKmers"])).read()
    assert "UnsupportedOperationException" not in src


def test_the_check_notices_a_missing_member():
    bogus = [{"file": "X.java", "line": 1, "cls": "RepGenomeDb", "member": "noSuchMember", "kind": "call", "nargs": 0},
             {"file": "X.java", "line": 2, "cls": "Representation", "member": "<init>", "kind": "new", "nargs": 5},
             {"file": "X.java", "line": 3, "cls": "RepGenome", "member": "size", "kind": "call", "nargs": 0}]       # inherited: resolves
    assert [m["line"] for m in cj.check(bogus)] == [1, 2]


def test_ffm_handles_match_the_header():
    assert cj.check_ffm_against_header() == []
    assert cj.check_brackets() == []


def test_committed_use_list_is_current():
    if not os.path.isdir(os.path.join(REF, "src")):
        return
    key = lambda x: (x["file"], x["line"], x["cls"], x["member"], x["kind"], x["nargs"] if x["nargs"] is not None else -1)
    assert sorted(cj.reference_uses(REF), key=key) == sorted(json.load(open(cj.GOLDEN)), key=key)
