"""TEST INFRASTRUCTURE ONLY -- expected text output of the reference's hot-path commands, restated in
Python on top of the CPU oracle, for byte-level comparison with kmers/reps_b200/host/kmers_reps.

Follows (all under src/main/java/org/theseed/proteins/kmers/reps/):
  repdb      RepGenomeDbProcessor.java:119-169 (create: file order, >= maxErr 'X' deferred), :224-252 (report)
  distances  DistanceMatrixProcessor.java:59,71-83
  group      RepMatrixProcessor.java:88-125 + SequenceInfo.java:62-148 (double accumulation in that exact order)
  classify   ClassifyProcessor.java:112-120
  neighbors  NeighborProcessor.java:155-230 (insertion-sorted top-N with tie retention; Double.toString columns)
  buildsets  BaseGenomeProcessor.java:151-214 (collateGenomes / copyReps / buildRepGenSet), :348-407 (list files)
"""
import math

import numpy as np

from . import kr_oracle as ko


def repdb_create(records, threshold, K=8, prot_name="Phenylalanyl-tRNA synthetase alpha chain", max_err=20, ident_inf=True):
    """records: [(fid, name, seq)] in file order -> bytes of the saved DB"""
    o = ko.get()
    main = [r for r in records if r[2].count("X") < max_err]
    deferred = [r for r in records if r[2].count("X") >= max_err]
    order = main + deferred
    b = o.batch([r[2] for r in order], K=K)
    n, added, _ = o.greedy_build(b, threshold, ident_inf=ident_inf)
    kept = [order[i] for i in np.nonzero(added)[0]]
    # HashMap<genomeId, RepGenome>: a later rep of the same genome replaces the earlier one in place
    ids, by_id = [], {}
    for r in kept:
        g = ko.genome_id_of(r[0])
        if g not in by_id:
            ids.append(g)
        by_id[g] = r
    it = o.java_hashmap_order(ids)
    reps = [by_id[ids[i]] for i in it]
    return ko.save_repdb(threshold, K, prot_name, reps), reps


def repdb_report(db_reps, threshold, queries, K=8, ident_inf=True):
    """db_reps in iteration (file) order; queries: [(label, comment, seq)] -> report text"""
    o = ko.get()
    rb = o.batch([r[2] for r in db_reps], K=K)
    qb = o.batch([q[2] for q in queries], K=K)
    bs, bi = o.find_closest(rb, qb, ident_inf=ident_inf)
    out = ["genome_id\tname\trep_id\tscore"]
    for q, s, i in zip(queries, bs, bi):
        rep_id = ko.genome_id_of(db_reps[i][0]) if s > 0 else ""
        out.append("%s\t%s\t%s\t%s" % (ko.genome_id_of(q[0]), q[1], rep_id, "MAX" if s == ko.INFINITY else str(s)))
    return "\n".join(out) + "\n"


def _natural_key(gid):
    """genome id as (taxon number, version number, ...): the other order RepGenome.compareTo may impose"""
    return [(0, int(f), f) if f.isdigit() else (1, 0, f) for f in gid.split(".")]


def compare_ids(a, b, id_order="lexical"):
    """sign of RepGenome.compareTo under the two candidate orders of genome-id strings"""
    if id_order == "natural":
        ka, kb = _natural_key(a), _natural_key(b)
        for x, y in zip(ka, kb):
            if x[0] == 0 and y[0] == 0:
                if x[1] != y[1]:
                    return -1 if x[1] < y[1] else 1
            elif x[2] != y[2]:
                return -1 if x[2] < y[2] else 1
        if len(ka) != len(kb):
            return -1 if len(ka) < len(kb) else 1
    return -1 if a < b else 1 if a > b else 0


def distances(db_reps, K=8, ident_inf=True, id_order="lexical"):
    o = ko.get()
    b = o.batch([r[2] for r in db_reps], K=K)
    m = o.sim_matrix(b, ident_inf=ident_inf, threads=4)
    sz = b.sizes()
    ids = [ko.genome_id_of(r[0]) for r in db_reps]
    out = ["genome_1\tgenome_2\tdistance\tsim"]
    for i in range(len(ids)):
        for j in range(len(ids)):
            if compare_ids(ids[i], ids[j], id_order) < 0:
                d = o.distance(m[i, j], sz[i], sz[j])
                out.append("%s\t%s\t%s\t%d" % (ids[i], ids[j], ko.java_fixed(d, 8, 6), m[i, j]))
    return "\n".join(out) + "\n"


def group(records, max_dist=0.995, K=8, ident_inf=True):
    """-> (fasta text of kept sequences, count)"""
    o = ko.get()
    b = o.batch([r[2] for r in records], K=K)
    m = o.sim_matrix(b, ident_inf=ident_inf, threads=4)
    sz = b.sizes()
    n = len(records)
    cnt = [0] * n
    s1 = [0.0] * n
    s2 = [0.0] * n
    mx = [0.0] * n
    mn = [1.7976931348623157e308] * n

    def add(k, d):
        cnt[k] += 1
        s1[k] += d
        s2[k] += d * d
        if d > mx[k]:
            mx[k] = d
        if d < mn[k]:
            mn[k] = d

    def dist(i, j):
        return o.distance(m[i, j], sz[i], sz[j])

    for i in range(n):
        for j in range(i):
            d = dist(i, j)
            add(i, d)
            add(j, d)

    def sdev(k):
        N = float(cnt[k])
        Nm1 = N - 1
        return math.sqrt(s2[k] / Nm1 - s1[k] * s1[k] / (N * Nm1)) if N > 1.0 else 0.0

    order = sorted(range(n), key=lambda k: (s1[k], sdev(k), records[k][0]))
    kept, out = [], []
    for x in order:
        if mn[x] > max_dist:
            continue
        if not any(dist(x, y) <= max_dist for y in kept):
            kept.append(x)
            out.append(ko.write_fasta_record(*records[x]).decode())
    return "".join(out), len(kept)


def classify(groups, queries, max_dist=0.995, K=8, ident_inf=True):
    """groups: {name: [seq,...]}; queries: [(label, comment(group), seq)]"""
    o = ko.get()
    out = ["seq_id\tgroup\tdistance"]
    for (label, comment, seq) in queries:
        best = 1.0
        members = groups.get(comment, [])
        if members:
            qb = o.batch([seq], K=K)
            mb = o.batch(members, K=K)
            m = o.sim_matrix(qb, mb, ident_inf=ident_inf)
            qs, ms = qb.sizes(), mb.sizes()
            for k in range(len(members)):
                best = min(best, o.distance(m[0, k], qs[0], ms[k]))
        if best <= max_dist:
            out.append("%s\t%s\t%s" % (label, comment, ko.java_fixed(best, 12, 8)))
    return "\n".join(out) + "\n"


def buildsets(records, thresholds, K=8, prot_name="Phenylalanyl-tRNA synthetase alpha chain", ident_inf=True):
    """records: [(fid, name, seq)] in master order -> {thr: (db bytes, list-file text)}"""
    o = ko.get()
    out = {}
    prev = []                                     # reps of the previous (smaller) set, in its iteration order
    qb = o.batch([r[2] for r in records], K=K)
    qs = qb.sizes()
    for thr in thresholds:
        order = prev + list(records)              # copyReps first, then every genome in master order
        b = o.batch([r[2] for r in order], K=K)
        n, added, _ = o.greedy_build(b, thr, ident_inf=ident_inf)
        kept = [order[i] for i in np.nonzero(added)[0]]
        ids, by_id = [], {}
        for r in kept:
            g = ko.genome_id_of(r[0])
            if g not in by_id:
                ids.append(g)
            by_id[g] = r
        reps = [by_id[ids[i]] for i in o.java_hashmap_order(ids)]
        db_bytes = ko.save_repdb(thr, K, prot_name, reps)
        rb = o.batch([r[2] for r in reps], K=K)
        rs = rb.sizes()
        bs, bi = o.find_closest(rb, qb, ident_inf=ident_inf, threads=4)
        lines = ["genome_id\tgenome_name\trep_id\tscore\tdistance"]
        for q, (rec, s, i) in enumerate(zip(records, bs, bi)):
            rep_id = ko.genome_id_of(reps[i][0]) if s > 0 else "null"
            d = o.distance(s, qs[q], rs[i]) if s > 0 else 1.0
            lines.append("%s\t%s\t%s\t%d\t%s" % (ko.genome_id_of(rec[0]), rec[1], rep_id, s, ko.java_fixed(d, 4, 2)))
        out[thr] = (db_bytes, "\n".join(lines) + "\n")
        prev = reps
    return out


def java_double_to_string(x: float) -> str:
    """Double.toString (JDK 19+: shortest round-trip digits; plain notation for 1e-3 <= |x| < 1e7)"""
    from decimal import Decimal
    if x == 0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign, digits, exp = Decimal(repr(x)).as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:
        digits.pop()
        exp += 1
    point = len(digits) + exp                      # value = 0.d1d2... * 10**point
    ds = "".join(map(str, digits))
    if 1e-3 <= abs(x) < 1e7:
        if point <= 0:
            body = "0." + "0" * (-point) + ds
        elif point >= len(ds):
            body = ds + "0" * (point - len(ds)) + ".0"
        else:
            body = ds[:point] + "." + ds[point:]
    else:
        body = ds[0] + "." + (ds[1:] or "0") + "E" + str(point - 1)
    return ("-" if sign else "") + body


def neighbors(db_reps, seeds, num=5, min_sim=1, K=8, ident_inf=True):
    """db_reps: [(fid, name, seq)] in iteration (file) order; seeds: [(fid, name, dna_size, seq)] -> report text"""
    o = ko.get()
    rb = o.batch([r[2] for r in db_reps], K=K)
    qb = o.batch([q[3] for q in seeds], K=K)
    m = o.sim_matrix(qb, rb, ident_inf=ident_inf, threads=4)
    rs, qs = rb.sizes(), qb.sizes()
    out = ["genome_id\tgenome_name\tsize\trep_id\trep_name\tsim\tdistance\tclosest\tsignal"]
    for g, (fid, name, dna_size, _) in enumerate(seeds):
        found = []                                  # (sim, rep index), as the ArrayList of :161
        for r in range(len(db_reps)):
            sim = int(m[g, r])
            if sim >= min_sim:
                idx, n = 0, len(found)
                while idx < n and found[idx][0] > sim:
                    idx += 1
                if idx < n or n < num:
                    found.insert(idx, (sim, r))
                    if len(found) > num:
                        if found[num - 1][0] > found[num][0]:
                            while len(found) > num:
                                del found[num]
        prefix = "%s\t%s\t%d" % (ko.genome_id_of(fid), name, dna_size)
        if not found:
            out.append(prefix + "\t\t\t\t\t\t")
            continue
        flag = "Y"
        for sim, r in found:
            dist = o.distance(sim, qs[g], rs[r])
            signal = dna_size * (1.0 - dist)
            out.append("%s\t%s\t%s\t%d\t%s\t%s\t%s" % (prefix, ko.genome_id_of(db_reps[r][0]), db_reps[r][1], sim,
                                                     java_double_to_string(dist), flag, java_double_to_string(signal)))
            flag = " "
    return "\n".join(out) + "\n"


def gtoreps(db_reps, threshold, genomes, update, K=8):
    """GtoRepGenomeProcessor.java:198,270-314.  db_reps: [(fid, name, seq)] of the loaded database in file order (empty for
    --create); genomes: [(fid, genome_name, seq)] in processing order (deferrals last); update: outliers are added.
    -> (report text, final reps in HashMap iteration order)"""
    o = ko.get()
    ids, by_id = [], {}
    for r in db_reps:
        g = ko.genome_id_of(r[0])
        if g not in by_id:
            ids.append(g)
        by_id[g] = r
    out = ["genome_id\tgenome_name\trep_id\trep_name\tsimilarity\tdistance\toutlier"]
    for (fid, name, seq) in genomes:
        gid = ko.genome_id_of(fid)
        reps = [by_id[ids[i]] for i in o.java_hashmap_order(ids)] if ids else []
        rep_id, rep_name, sim, dist, flag = "", "", 0, 1.0, ""
        if reps:
            rb, qb = o.batch([r[2] for r in reps], K=K), o.batch([seq], K=K)
            bs, bi = o.find_closest(rb, qb)
            if bs[0] > 0:
                rep_id, rep_name, sim = ko.genome_id_of(reps[bi[0]][0]), reps[bi[0]][1], int(bs[0])
                dist = o.distance(sim, qb.sizes()[0], rb.sizes()[bi[0]])
        if sim == 0 or sim < threshold:
            flag = "*"
            if update:
                if gid not in by_id:
                    ids.append(gid)
                by_id[gid] = (fid, name, seq)
                rep_id, sim, dist = gid, 9999, 0.0
        out.append("%s\t%s\t%s\t%s\t%d\t%s\t%s" % (gid, name, rep_id, rep_name, sim, ko.java_fixed(dist, 8, 4), flag))
    final = [by_id[ids[i]] for i in o.java_hashmap_order(ids)] if ids else []
    return "\n".join(out) + "\n", final


def buildfromgtos_verify(db_reps, threshold, queries, K=8):
    """GtoRepBuildProcessor.java:147-160: the sequences of the verification FASTA that the database does not represent"""
    o = ko.get()
    out = ["seq_id\tname\tsimilarity"]
    if queries:
        bs, _ = o.find_closest(o.batch([r[2] for r in db_reps], K=K), o.batch([q[2] for q in queries], K=K))
        for q, s in zip(queries, bs):
            if s < threshold:
                out.append("%s\t%s\t%d" % (q[0], q[1], s))
    return "\n".join(out) + "\n"
