#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference's data files (run in the build container,
where /root/reference exists; the GPU box only ever sees the committed outputs).

Inputs are the reference's own fixtures:
  data/rep50.db, data/VapbProt.fa, data/seq_test.fa, the 10 bundled GTOs
  (data/202462.4.gto, data/P3Master/*.gto) and data/rickettsia.roles.tbl.
They are stored gzip-compressed (data, not source).  Expected OUTPUTS in expected.json are
produced by the pure-Python set-of-substrings model in oracle/kr_oracle.py (not by the Java
code, which cannot run here) and are cross-checked against the C oracle by
tests/test_oracle_golden.py and against the figures measured in SURVEY.md section 8c.

usage: python tests/golden/make_golden.py [/root/reference]
"""
import gzip
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import kr_oracle as ko  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
DATA = os.path.join(REF, "data")
SEED_FUNCTION = "Phenylalanyl-tRNA synthetase alpha chain"


def gz_write(name, data: bytes):
    with open(os.path.join(HERE, name), "wb") as f:
        f.write(gzip.compress(data, 9, mtime=0))


def comment_free(fn: str) -> str:
    """RoleUtilities.commentFree: strip a trailing '# comment' / '! comment'."""
    return re.split(r"\s*[#!]", fn, 1)[0].strip()


def seed_protein(gto_path):
    """Longest feature whose comment-free function is the seed protein role, with or without
    the EC number (ScoreCouplingReporter.java:135,164-168; TestRolePatterns.java:39-41)."""
    g = json.load(open(gto_path))
    best = None
    for ft in g["features"]:
        fn = comment_free(ft.get("function") or "")
        if fn in (SEED_FUNCTION, SEED_FUNCTION + " (EC 6.1.1.20)"):
            prot = ft.get("protein_translation") or ""
            if best is None or len(prot) > len(best[2]):
                best = (ft["id"], g["scientific_name"], prot)
    return g["id"], best


def main():
    exp = {}
    # --- raw fixtures ------------------------------------------------------------------
    rep50 = open(os.path.join(DATA, "rep50.db"), "rb").read()
    vapb = open(os.path.join(DATA, "VapbProt.fa"), "rb").read()
    seqtest = open(os.path.join(DATA, "seq_test.fa"), "rb").read()
    gz_write("rep50.db.gz", rep50)
    gz_write("VapbProt.fa.gz", vapb)
    gz_write("seq_test.fa.gz", seqtest)
    exp["sha256"] = {"rep50.db": hashlib.sha256(rep50).hexdigest(),
                     "VapbProt.fa": hashlib.sha256(vapb).hexdigest()}

    # --- config 1 inputs: seed proteins of the 10 bundled GTOs ----------------------------
    gtos = [os.path.join(DATA, "202462.4.gto")] + sorted(
        os.path.join(DATA, "P3Master", f) for f in os.listdir(os.path.join(DATA, "P3Master")) if f.endswith(".gto"))
    seeds = []
    for p in gtos:
        gid, best = seed_protein(p)
        assert best is not None, p
        seeds.append((best[0], best[1], best[2]))
    with open(os.path.join(HERE, "gto_seeds.fa"), "wb") as f:
        for (fid, name, prot) in seeds:
            f.write(ko.write_fasta_record(fid, name, prot))

    # --- reduced GTOs for the gtoclass command (GtoClassProcessor): identity fields + a dozen features each ------------
    mini = os.path.join(HERE, "gto_mini")
    os.makedirs(mini, exist_ok=True)
    for p in (gtos[0], gtos[1], gtos[6]):                      # 202462.4 (Bacteria), 1001994.6 (Archaea), 1408469.3
        g = json.load(open(p))
        gid, best = seed_protein(p)
        keep = [ft for ft in g["features"] if ft["id"] == best[0]]
        keep += [ft for ft in g["features"] if ft["id"] != best[0]][:11]
        feats = [{k: ft.get(k) for k in ("id", "type", "function", "protein_translation")} for ft in keep]
        if gid == "202462.4":
            # a shorter feature with the same role written with its EC number and a comment: the longest one must still win
            feats.insert(0, {"id": "fig|202462.4.peg.9001", "type": "CDS", "function": SEED_FUNCTION + " (EC 6.1.1.20) # fragment",
                             "protein_translation": best[2][:120]})
        out = {"id": g["id"], "scientific_name": g["scientific_name"], "domain": g["domain"], "genetic_code": g.get("genetic_code", 11),
               "features": feats, "contigs": [], "note": "reduced copy of data/%s for tests (tests/golden/make_golden.py)" % os.path.basename(p)}
        with open(os.path.join(mini, gid + ".gto"), "w") as f:
            json.dump(out, f, indent=1)
    euk = {"id": "4932.99", "scientific_name": "Saccharomyces test genome \u00e9", "domain": "Eukaryota",
           "features": [{"id": "fig|4932.99.peg.1", "type": "CDS", "function": SEED_FUNCTION, "protein_translation": seeds[0][2]}]}
    json.dump(euk, open(os.path.join(mini, "4932.99.gto"), "w"), indent=1)
    noseed = {"id": "99999.1", "scientific_name": "Genome without \"seed\" protein", "domain": "Bacteria",
              "features": [{"id": "fig|99999.1.peg.1", "type": "CDS", "function": "Phenylalanyl-tRNA synthetase beta chain (EC 6.1.1.20)",
                            "protein_translation": "MKVLAAGIVGLAAA"}, {"id": "fig|99999.1.rna.1", "type": "rna", "function": "tRNA-Thr", "protein_translation": ""}]}
    json.dump(noseed, open(os.path.join(mini, "99999.1.gto"), "w"), indent=1)

    # --- duplicate-heavy real set -------------------------------------------------------
    lines = open(os.path.join(DATA, "rickettsia.roles.tbl")).read().split("\n")
    hdr = lines[0].split("\t")
    gi, si = hdr.index("genome"), hdr.index("seed_prot")
    rick = []
    for ln in lines[1:]:
        if ln.strip():
            c = ln.split("\t")
            rick.append((c[gi], c[si]))
    gz_write("rickettsia_seed.tsv.gz", "".join("%s\t%s\n" % r for r in rick).encode())

    # --- expected outputs (pure-Python model) -----------------------------------------
    db = ko.parse_repdb(rep50)
    K = db["K"]
    reps = db["reps"]
    rep_sets = [ko.py_kmers(s, K) for (_, _, s) in reps]
    exp["rep50"] = dict(threshold=db["threshold"], K=K, prot_name=db["prot_name"], n=len(reps),
                        total_kmers=sum(len(s) for s in rep_sets),
                        distinct_kmers=len(set().union(*rep_sets)))
    # inverted index for the all-pairs stats (exact, fast enough in Python)
    from collections import defaultdict
    post = defaultdict(list)
    for i, s in enumerate(rep_sets):
        for k in s:
            post[k].append(i)
    pair = defaultdict(int)
    for k, lst in post.items():
        for x in range(len(lst)):
            for y in range(x + 1, len(lst)):
                pair[(lst[x], lst[y])] += 1
    hist = defaultdict(int)
    for v in pair.values():
        hist[v] += 1
    exp["rep50"].update(max_sim=max(pair.values()), pairs_at_max=hist[max(pair.values())],
                        nonzero_pairs=len(pair), total_pairs=len(reps) * (len(reps) - 1) // 2,
                        sum_sim=sum(pair.values()), max_posting=max(len(v) for v in post.values()))
    # config 1: closest rep of each GTO seed in rep50.db (file order == HashMap order)
    c1 = []
    for (fid, name, prot) in seeds:
        q = ko.py_kmers(prot, K)
        best, bi, ties = 0, -1, 0
        for i, s in enumerate(rep_sets):
            sim = ko.INFINITY if prot == reps[i][2] else len(q & s)
            if sim > best:
                best, bi, ties = sim, i, 1
            elif sim == best and sim > 0:
                ties += 1
        dist = ko.py_distance(best, len(q), len(rep_sets[bi]))
        c1.append(dict(genome=ko.genome_id_of(fid), fid=fid, rep=ko.genome_id_of(reps[bi][0]), rep_index=bi,
                       sim=best, ties=ties, distance=dist, size=len(q)))
    exp["config1"] = c1
    # config 2: VapbProt all-vs-all raw counts (diag included, no INFINITY substitution)
    vrecs = ko.read_fasta(vapb)
    vsets = [ko.py_kmers(s, 8) for (_, _, s) in vrecs]
    n = len(vrecs)
    rows = []
    tot = nz = mx = ident = 0
    for i in range(n):
        row = [len(vsets[i] & vsets[j]) for j in range(n)]
        rows.append(",".join(map(str, row)))
        for j in range(i + 1, n):
            tot += row[j]
            nz += row[j] > 0
            mx = max(mx, row[j])
            ident += vrecs[i][2] == vrecs[j][2]
    exp["config2"] = dict(n=n, pairs=n * (n - 1) // 2, sum_sim=tot, nonzero_pairs=nz, max_raw_sim=mx,
                          identical_pairs=ident, first_id=vrecs[0][0],
                          sizes_first2=[len(vsets[0]), len(vsets[1])],
                          sha256_raw_rows=hashlib.sha256("\n".join(rows).encode()).hexdigest())
    # rickettsia duplicates
    exp["rickettsia"] = dict(n=len(rick), distinct=len({p for (_, p) in rick}))
    # upstream KAT recalled in SURVEY.md section 4 (K=10)
    p1 = "MGMLVPLISKISDLSEEAKACVAACSSVEELDEVRGRYIGRAGALTALLA"
    p2 = "MDINLFKEELEELAKKAKHMLNETASKNDLEQVKVSLLGKKGLLTLQSAA"
    p3 = "MDINLFKEELKHMLNETASKKGLLTLQSA"
    exp["kat_k10"] = dict(p1=p1, p2=p2, p3=p3, size_p1=len(ko.py_kmers(p1, 10)),
                          sim_p2_p3=ko.py_similarity(p2, p3, 10), sim_p3_p2=ko.py_similarity(p3, p2, 10),
                          dist_p1_p2=ko.py_distance(ko.py_similarity(p1, p2, 10), len(ko.py_kmers(p1, 10)), len(ko.py_kmers(p2, 10))),
                          dist_p2_p3=ko.py_distance(ko.py_similarity(p2, p3, 10), len(ko.py_kmers(p2, 10)), len(ko.py_kmers(p3, 10))))
    with open(os.path.join(HERE, "expected.json"), "w") as f:
        json.dump(exp, f, indent=1, sort_keys=True)
        f.write("\n")
    print(json.dumps({k: exp[k] for k in ("rep50", "config2", "rickettsia", "kat_k10")}, indent=1))
    for r in c1:
        print(r["genome"], "->", r["rep"], r["sim"], "%.6f" % r["distance"], "ties", r["ties"])


if __name__ == "__main__":
    main()
