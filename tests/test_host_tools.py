"""GPU tests of the host layer (C++ mirror of the reference's Java classes + commands): the text the tools
print and the database files they write are compared byte for byte with the restatement in
oracle/ref_commands.py and with the reference's own data/rep50.db."""
import os
import subprocess

import numpy as np
import pytest

from kmers.reps_b200 import synth
from oracle import kr_oracle as ko
from oracle import ref_commands as rc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.environ.get("KR_TOOL_PATH") or os.path.join(ROOT, "kmers", "reps_b200", "host", "kmers_reps")   # (tools/emul_gpu_suite.sh points it at a build linked against the emulator)


def run(args, stdin=b"", check=True, env=None):
    p = subprocess.run([TOOL] + args, input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600,
                       env=dict(os.environ, **env) if env else None)
    if check:
        assert p.returncode == 0, p.stderr.decode()
    return p


def fasta(records):
    return b"".join(ko.write_fasta_record(*r) for r in records)


def test_tool_is_built():
    assert os.path.exists(TOOL), "run __graft_entry__.build()"


def test_repdb_load_report_and_list(tmp_path, rep50, gto_seeds, expected):
    db = tmp_path / "rep50.db"
    db.write_bytes(rep50["raw"])
    lst = tmp_path / "list.tbl"
    p = run(["repdb", "--list", str(lst), str(db)], stdin=fasta(gto_seeds))
    assert p.stdout.decode() == rc.repdb_report(rep50["reps"], 50, gto_seeds)
    rows = p.stdout.decode().split("\n")[1:-1]
    for row, e in zip(rows, expected["config1"]):
        g, name, rep, score = row.split("\t")
        assert (g, rep, int(score)) == (e["genome"], e["rep"], e["sim"])
    want = sorted((ko.genome_id_of(f), n) for (f, n, _) in rep50["reps"])
    got = [tuple(l.split("\t")) for l in lst.read_text().split("\n")[1:-1]]
    assert got == want
    # a member of the database finds itself with score MAX
    p = run(["repdb", str(db)], stdin=fasta(rep50["reps"][:5]))
    assert [l.split("\t")[3] for l in p.stdout.decode().split("\n")[1:-1]] == ["MAX"] * 5


def test_repdb_create_reproduces_rep50_byte_for_byte(tmp_path, rep50):
    # every pair of rep50's reps is below the threshold, so a build at 50 keeps them all; fed in the file's own
    # order the saved file (java.util.HashMap order by genome id, bins in insertion order) is the reference's
    # file again, byte for byte
    src = tmp_path / "in.fa"
    src.write_bytes(fasta(rep50["reps"]))
    out = tmp_path / "out.db"
    run(["repdb", "--create", str(src), "-m", "50", "--null", "--batch", "400", str(out)])
    assert out.read_bytes() == rep50["raw"]
    # any other input order changes the order inside hash bins exactly as java.util.HashMap would
    rng = np.random.default_rng(17)
    shuffled = [rep50["reps"][i] for i in rng.permutation(len(rep50["reps"]))]
    src.write_bytes(fasta(shuffled))
    run(["repdb", "--create", str(src), "-m", "50", "--null", str(out)])
    want, reps = rc.repdb_create(shuffled, 50)
    assert len(reps) == 1480 and out.read_bytes() == want and want != rep50["raw"]


def test_repdb_create_synthetic_with_deferral(tmp_path):
    res, off, _ = synth.generate(0x5EED0004, 3000, 75, xrun_per_10k=200)
    seqs = synth.to_strings(res, off)
    recs = [("fig|%d.%d.peg.%d" % (1000 + 7 * i, 1 + i % 9, i), "Synthetic genome %d" % i, s) for i, s in enumerate(seqs)]
    assert sum(s.count("X") >= 20 for s in seqs) > 10
    src = tmp_path / "cands.fa"
    src.write_bytes(fasta(recs))
    for thr in (50, 100):
        out = tmp_path / ("syn%d.db" % thr)
        run(["repdb", "--create", str(src), "-m", str(thr), "--null", "--batch", "512", str(out)])
        want, reps = rc.repdb_create(recs, thr)
        assert out.read_bytes() == want
        assert 0 < len(reps) < len(recs)


def test_distances(tmp_path, rep50):
    reps = rep50["reps"][:300]
    db = tmp_path / "small.db"
    db.write_bytes(ko.save_repdb(50, 8, rep50["prot_name"], reps))
    p = run(["distances", str(db)])
    assert p.stdout.decode() == rc.distances(reps)
    # RepGenome.compareTo's order of genome ids is a policy (un-vendored class): the natural order visits the same pairs,
    # oriented and ordered differently
    pn = run(["distances", str(db)], env={"KR_ID_ORDER": "natural"})
    assert pn.stdout.decode() == rc.distances(reps, id_order="natural")
    assert pn.stdout != p.stdout and len(pn.stdout.splitlines()) == len(p.stdout.splitlines())


def test_group_vapb(vapb):
    p = run(["group"], stdin=fasta(vapb))
    want, n = rc.group(vapb)
    assert p.stdout.decode() == want
    assert p.stderr.decode().strip().endswith("%d sequences form the representation set." % n)
    p2 = run(["group", "-m", "0.9", "-K", "6"], stdin=fasta(vapb))
    want2, n2 = rc.group(vapb, max_dist=0.9, K=6)
    assert p2.stdout.decode() == want2 and n2 != n


def test_classify(tmp_path, vapb, rickettsia):
    ga = [r for r in vapb[:120]]
    gb = [("r%d" % i, "", p) for i, (_, p) in enumerate(rickettsia[:40])]
    (tmp_path / "a.fa").write_bytes(fasta(ga))
    (tmp_path / "b.fa").write_bytes(fasta(gb))
    ctl = tmp_path / "control.tbl"
    ctl.write_text("anti\t%s\nphes\t%s\n" % (tmp_path / "a.fa", tmp_path / "b.fa"))
    queries = [(l, "anti", s) for (l, _, s) in vapb[100:160]] + [("q%d" % i, "phes", p) for i, (_, p) in enumerate(rickettsia[30:60])]
    queries += [("lost", "nogroup", vapb[0][2])]
    p = run(["classify", str(ctl)], stdin=fasta(queries))
    want = rc.classify({"anti": [s for (_, _, s) in ga], "phes": [s for (_, _, s) in gb]}, queries)
    assert p.stdout.decode() == want and want.count("\n") > 20


def test_error_behaviour(tmp_path):
    p = run(["repdb", str(tmp_path / "missing.db")], check=False)
    assert p.returncode == 1 and "FileNotFoundException" in p.stderr.decode()
    p = run(["repdb", "--bogus", "x"], check=False)
    assert p.returncode == 1 and "not a valid option" in p.stderr.decode()
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b">fig|1.1.peg.1 x\nACDEFGHIKLMNPQRSTVWYacd\n")
    p = run(["repdb", "--create", str(bad), "--null", str(tmp_path / "o.db")], check=False)
    assert p.returncode == 1 and "outside" in p.stderr.decode()


def test_neighbors_top_n_with_ties(tmp_path, rep50, gto_seeds, rickettsia, oracle):
    """`neighbors` (NeighborProcessor.java:155-230): the N closest representatives of each genome with the
    reference's tie retention.  The database holds 100 Rickettsia seed proteins of which only 34 are distinct, so
    equal similarities (and equal-sequence INFINITY hits) straddle the cut for most queries."""
    recs = list(rep50["reps"][:300]) + [("fig|%s.peg.1" % g, "Rickettsia %s" % g, p) for (g, p) in rickettsia]
    ids = [ko.genome_id_of(r[0]) for r in recs]
    assert len(set(ids)) == len(ids)
    db = tmp_path / "mix.db"
    db.write_bytes(ko.save_repdb(50, 8, "Phenylalanyl-tRNA synthetase alpha chain", recs))
    it_order = [recs[i] for i in oracle.java_hashmap_order(ids)]          # load() re-inserts in file order
    seeds = [(f, n, 1000000 + 1234 * k, s) for k, (f, n, s) in enumerate(gto_seeds)]
    seeds += [("fig|%s.peg.7" % g, "query %s" % g, 1100000 + 77 * k, p) for k, (g, p) in enumerate(rickettsia[:12])]
    seeds.append(("fig|9.9.peg.1", "nothing close", 5, "MKV" * 30))
    src = tmp_path / "seeds.tbl"
    src.write_text("seed_fid\tgenome_name\tdna_size\tseed_protein\n" + "".join("%s\t%s\t%d\t%s\n" % s for s in seeds))
    for num, min_sim in ((3, 1), (5, 1), (1, 60), (2, 0)):
        p = run(["neighbors", "-n", str(num), "--min", str(min_sim), str(db), str(src)])
        want = rc.neighbors(it_order, seeds, num=num, min_sim=min_sim)
        assert p.stdout.decode() == want
    text = rc.neighbors(it_order, seeds, num=3, min_sim=1)
    per = {}
    for ln in text.split("\n")[1:-1]:
        per.setdefault(ln.split("\t")[0], []).append(ln.split("\t"))
    assert max(len(v) for v in per.values()) > 3                          # ties across the cut were kept
    assert any(r[5] == "9999" and r[6] == "0.0" for v in per.values() for r in v)
    assert per["9.9"][0][3:] == ["", "", "", "", "", ""]
    assert run(["neighbors", "-n", "0", str(db), str(src)], check=False).returncode != 0


def test_gtoclass_on_reduced_gtos(tmp_path, rep50, gto_seeds, oracle):
    """`gtoclass` (GtoClassProcessor.java:114-151, BASELINE configs[0]): GTO ingest, seed-protein lookup (longest feature
    with the role, EC number and comment tolerated), domain filter, one batched findClosest per database; blank column when
    the genome is not represented.  tests/golden/gto_mini holds reduced copies of three bundled GTOs plus a eukaryote and a
    genome without seed protein (tests/golden/make_golden.py)."""
    gdir = os.path.join(ROOT, "tests", "golden", "gto_mini")
    db50 = tmp_path / "rep50.db"
    db50.write_bytes(rep50["raw"])
    db300 = tmp_path / "rep300.db"
    db300.write_bytes(rep50["raw"].replace(b">Rep50,K=8", b">Rep300,K=8", 1))
    p = run(["gtoclass", "-v", gdir, str(db50), str(db300)])
    seeds = {ko.genome_id_of(f): (n, s) for (f, n, s) in gto_seeds}
    rb = oracle.batch([s for (_, _, s) in rep50["reps"]], K=8)
    want = ["genome_id\tgenome_name\tRepGen.50\tRepGen.300"]
    for g in sorted(["202462.4", "1001994.6", "1408469.3"]):                      # directory order: by genome id
        bs, bi = oracle.find_closest(rb, oracle.batch([seeds[g][1]], K=8))
        rep = ko.genome_id_of(rep50["reps"][int(bi[0])][0])
        want.append("%s\t%s\t%s\t%s" % (g, seeds[g][0], rep if bs[0] >= 50 else "", rep if bs[0] >= 300 else ""))
    assert p.stdout.decode() == "\n".join(want) + "\n"
    assert "5 genomes processed. 3 representatives found; 5 failures." in p.stderr.decode()   # 3 misses at 300 + the seedless genome twice
    assert run(["gtoclass", str(tmp_path / "nowhere"), str(db50)], check=False).returncode != 0


def write_gtos(gdir, genomes, role="Phenylalanyl-tRNA synthetase alpha chain"):
    """synthetic GTO files: each genome gets its seed protein, a shorter decoy with the same role and an unrelated feature"""
    import json
    os.makedirs(gdir, exist_ok=True)
    for (gid, name, prot, ssu) in genomes:
        feats = [{"id": "fig|%s.peg.2" % gid, "type": "CDS", "function": role + " (EC 6.1.1.20) ## truncated", "protein_translation": prot[: len(prot) // 2]},
                 {"id": "fig|%s.peg.7" % gid, "type": "CDS", "function": role, "protein_translation": prot},
                 {"id": "fig|%s.rna.1" % gid, "type": "rna", "function": "tRNA-Thr", "protein_translation": ""}]
        json.dump({"id": gid, "scientific_name": name, "domain": "Bacteria", "ssu_rrna": ssu, "features": feats}, open(os.path.join(gdir, gid + ".gto"), "w"))


def test_gtoreps_and_buildfromgtos(tmp_path, oracle):
    """`gtoreps` (GtoRepGenomeProcessor.java:198-314): report mode, --create (every outlier joins the database and is seen by
    the genomes after it), --update with a black list (deferred genomes go last) and --rna; `buildfromgtos`
    (GtoRepBuildProcessor.java:96-163) with its --verify report.  Byte for byte against the sequential restatements."""
    res, off, _ = synth.generate(0x5EED000E, 90, 9)
    seqs = synth.to_strings(res, off)
    genomes = [("%d.%d" % (3000 + 7 * i, 1 + i % 3), "Genome number %d" % i, s, "acgt" * (400 if i % 9 else 100)) for i, s in enumerate(seqs)]
    genomes[13] = (genomes[13][0], genomes[13][1], genomes[13][2], "")                      # no SSU rRNA at all
    gdir = str(tmp_path / "gtos")
    write_gtos(gdir, genomes)
    order = sorted(genomes, key=lambda g: g[0])                                              # directory order: by genome id
    recs = [("fig|%s.peg.7" % g[0], g[1], g[2]) for g in order]
    # --create: the database grows while the directory is scanned
    db = tmp_path / "new.db"
    p = run(["gtoreps", "--create", "60", "-v", str(db), gdir])
    text, final = rc.gtoreps([], 60, recs, update=True)
    assert p.stdout.decode() == text
    assert db.read_bytes() == ko.save_repdb(60, 8, "Phenylalanyl-tRNA synthetase alpha chain", final)
    n_out = text.count("\t*\n")
    assert 5 < n_out < 80 and ("%d outliers found, 0 skipped" % n_out) in p.stderr.decode()
    # report mode against that database: nothing is added, every genome finds itself or its representative
    p = run(["gtoreps", str(db), gdir])
    text2, _ = rc.gtoreps(final, 60, recs, update=False)
    assert p.stdout.decode() == text2 and "\t*\n" not in text2
    # --update from a smaller database, black list + --rna: deferred genomes are processed last, the seedless SSU is skipped
    small = tmp_path / "small.db"
    small.write_bytes(ko.save_repdb(60, 8, "Phenylalanyl-tRNA synthetase alpha chain", final[:3]))
    black = tmp_path / "black.tbl"
    black.write_text("genome_id\tnote\n" + "".join("%s\tx\n" % g[0] for g in order[5:9]))
    upd = tmp_path / "updated.db"
    p = run(["gtoreps", "--update", str(upd), "--blackList", str(black), "--rna", "-o", str(tmp_path / "rep.tbl"), str(small), gdir])
    main = [r for g, r in zip(order, recs) if g[3] != "" and len(g[3]) >= 1400 and g not in order[5:9]]
    deferred = [r for g, r in zip(order, recs) if g[3] != "" and (g in order[5:9] or len(g[3]) < 1400)]
    # deferral order: black-listed and short-RNA genomes in directory order
    deferred = [r for g, r in zip(order, recs) if r in deferred]
    text3, final3 = rc.gtoreps(final[:3], 60, main + deferred, update=True)
    assert (tmp_path / "rep.tbl").read_text() == text3
    assert upd.read_bytes() == ko.save_repdb(60, 8, "Phenylalanyl-tRNA synthetase alpha chain", final3)
    # buildfromgtos: the same greedy set from the directory, then the verification report
    roles = tmp_path / "roles.tbl"
    roles.write_text("PhenTrnaSyntAlph\t\tPhenylalanyl-tRNA synthetase alpha chain\n")
    ver = tmp_path / "verify.fa"
    res2, off2, _ = synth.generate(0x5EED000F, 60, 12)
    vq = [("fig|%d.1.peg.1" % (9000 + i), "query %d" % i, s) for i, s in enumerate(synth.to_strings(res2, off2))] + [recs[0]]
    ver.write_bytes(fasta(vq))
    out_db = tmp_path / "built.db"
    p = run(["buildfromgtos", "-m", "60", "--verify", str(ver), gdir, str(roles), str(out_db)])
    want_db, reps = rc.repdb_create(recs, 60, max_err=10 ** 9)
    assert out_db.read_bytes() == want_db
    assert p.stdout.decode() == rc.buildfromgtos_verify(reps, 60, vq)
    assert run(["buildfromgtos", gdir, str(roles), str(tmp_path / "x.db")]).stdout.decode() == "No verification requested.\n"


def test_kmerops_set_members(vapb):
    """SequenceKmers.retainAll / removeAll / size / iteration / equals (TargetProcessor.java:160-195) through the host
    classes: every number and every retained k-mer equals Python set arithmetic on substrings"""
    recs = vapb[:40] + [vapb[0], ("short", "", "MKV")]
    p = run(["kmerops"], stdin=fasta(recs))
    other = ko.py_kmers(recs[0][2], 8)
    rows = p.stdout.decode().split("\n")[:-1]
    assert len(rows) == len(recs) - 1
    for (label, _, seq), row in zip(recs[1:], rows):
        f = row.split("\t")
        mine = ko.py_kmers(seq, 8)
        kept = mine & other
        assert f[0] == label and int(f[1]) == len(mine) and int(f[2]) == len(kept) and int(f[3]) == len(mine - other)
        want_sim = 9999 if (seq == recs[0][2] and kept == other) else len(kept)      # kept is a subset of other
        assert int(f[4]) == want_sim
        assert (f[5].split(",") if f[5] else []) == sorted(kept)
        assert f[6] == ("1" if kept == mine else "0") + ("1" if not (mine & other) else "0")


def test_commands_over_a_multi_gpu_handle(tmp_path, rep50, gto_seeds, vapb, monkeypatch):
    """KR_DEVICES: the same commands with the batch calls sharded inside the library over two contexts (kr_multi) print
    the same bytes: repdb report, repdb --create (greedy build), distances, group"""
    import torch
    ndev = 0 if "emul" in TOOL else torch.cuda.device_count()
    monkeypatch.setenv("KR_DEVICES", ",".join(str(i) for i in range(ndev)) if ndev > 1 else "0,0")
    db = tmp_path / "rep50.db"
    db.write_bytes(rep50["raw"])
    p = run(["repdb", str(db)], stdin=fasta(gto_seeds))
    assert p.stdout.decode() == rc.repdb_report(rep50["reps"], 50, gto_seeds)
    src = tmp_path / "in.fa"
    src.write_bytes(fasta(rep50["reps"]))
    out = tmp_path / "out.db"
    run(["repdb", "--create", str(src), "-m", "50", "--null", "--batch", "400", str(out)])
    assert out.read_bytes() == rep50["raw"]
    small = tmp_path / "small.db"
    small.write_bytes(ko.save_repdb(50, 8, rep50["prot_name"], rep50["reps"][:300]))
    assert run(["distances", str(small)]).stdout.decode() == rc.distances(rep50["reps"][:300])
    want, _ = rc.group(vapb)
    assert run(["group"], stdin=fasta(vapb)).stdout.decode() == want
