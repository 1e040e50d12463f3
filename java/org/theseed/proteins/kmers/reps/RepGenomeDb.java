package org.theseed.proteins.kmers.reps;

import java.io.File;
import java.io.IOException;
import java.io.PrintWriter;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.nio.charset.StandardCharsets;
import java.util.ArrayList;
import java.util.Collection;
import java.util.HashMap;
import java.util.Iterator;
import java.util.List;
import java.util.SortedSet;
import java.util.TreeSet;

import org.theseed.genome.Feature;
import org.theseed.genome.Genome;
import org.theseed.sequence.FastaInputStream;
import org.theseed.sequence.KmersRepsNative;
import org.theseed.sequence.ProteinKmers;
import org.theseed.sequence.Sequence;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Drop-in for org.theseed.proteins.kmers.reps.RepGenomeDb (org.theseed:shared): same public members as used by
 * kmers.reps (SURVEY.md appendix C), with the scan and build loops batched on the GPU.  The representatives stay
 * in a java.util.HashMap keyed by genome id, exactly as upstream, so save() order and findClosest tie-breaking
 * are the JVM's own; the device copy is told each representative's position in that iteration order
 * (kr_db_set_ranks).  New entry points: findClosest(Collection), checkGenomes(List).
 * COMPILE-UNVERIFIED (no JDK in the build container); the same logic is implemented and tested in
 * kmers/reps_b200/host/theseed.hpp (class RepGenomeDb).
 */
public class RepGenomeDb implements Iterable<RepGenome> {

    public static final String DEFAULT_PROTEIN = "Phenylalanyl-tRNA synthetase alpha chain";

    private final HashMap<String, RepGenome> genomeMap = new HashMap<>();
    private final int threshold;
    private final int kmerSize;
    private final String protName;
    private final String[] protNames;
    // device mirror
    // With -Dkmersreps.devices=0,1,... the mirror is a kr_multi_db (one replica per GPU, calls sharded inside the library);
    // otherwise a kr_db on the single context.  `db` holds whichever handle is in use.
    private final MemorySegment multi = KmersRepsNative.multi(ProteinKmers.kmerSize());
    private MemorySegment db = MemorySegment.NULL;
    private final List<RepGenome> onDevice = new ArrayList<>();   // device index -> rep
    private boolean rebuild = false, ranksDirty = true;

    public class Representation {
        private final int similarity;
        private final RepGenome representative;
        private final String genomeId;
        private final double distance;

        public Representation() { this((RepGenome) null, 0, 1.0); }
        public Representation(RepGenome rep, ProteinKmers query) {   // NeighborProcessor.java:168
            this(rep, rep.similarity(query), rep.distance(query));
        }
        /** a stored result: representative id, score and distance read back from a file (ProteinData.java:242) */
        public Representation(String repId, int score, double distance) {
            this.representative = repId == null ? null : genomeMap.get(repId);
            this.genomeId = repId; this.similarity = score; this.distance = distance;
        }
        Representation(RepGenome rep, int sim, double dist) {
            this.representative = rep; this.genomeId = rep == null ? null : rep.getGenomeId(); this.similarity = sim; this.distance = dist;
        }
        public int getSimilarity() { return similarity; }
        public double getDistance() { return distance; }
        public RepGenome getRepresentative() { return representative; }
        public String getGenomeId() { return genomeId; }
        public boolean isRepresented() { return similarity >= threshold; }
    }

    public RepGenomeDb(int threshold, String... protNames) {
        this.threshold = threshold;
        this.kmerSize = ProteinKmers.kmerSize();
        this.protName = protNames.length > 0 ? protNames[0] : DEFAULT_PROTEIN;
        this.protNames = protNames.length > 0 ? protNames.clone() : new String[] { DEFAULT_PROTEIN };
    }

    /**
     * The genome's seed protein as a RepGenome, or null when it has none (GtoRepGenomeProcessor.java:213,
     * NeighborProcessor.java:96, BuildRepDbProcessor.java:134): the LONGEST protein feature whose comment-free function
     * is one of this database's protein names, with or without a trailing EC number -- the rule of
     * GtoRepBuildProcessor.computeRepGenome (:176-192) and ScoreCouplingReporter.java:135,164-168.
     */
    public RepGenome getSeedProtein(Genome genome) {
        String bestProtein = "", fid = null;
        for (Feature feat : genome.getFeatures()) {
            String prot = feat.getProteinTranslation();
            if (prot == null || prot.isEmpty() || !isSeedFunction(feat.getFunction())) continue;
            if (prot.length() > bestProtein.length()) { bestProtein = prot; fid = feat.getId(); }
        }
        return fid == null ? null : new RepGenome(fid, genome.getName(), bestProtein);
    }

    private boolean isSeedFunction(String function) {
        if (function == null) return false;
        int cut = function.length();
        for (int i = 0; i < function.length(); i++) { char c = function.charAt(i); if (c == '#' || c == '!') { cut = i; break; } }
        String fn = function.substring(0, cut).strip();
        int ec = fn.lastIndexOf(" (EC ");
        String bare = (ec > 0 && fn.endsWith(")")) ? fn.substring(0, ec) : fn;
        for (String name : protNames) if (fn.equals(name) || bare.equals(name)) return true;
        return false;
    }

    public int getThreshold() { return threshold; }
    public int getKmerSize() { return kmerSize; }
    public String getProtName() { return protName; }
    public int size() { return genomeMap.size(); }
    public RepGenome get(String genomeId) { return genomeMap.get(genomeId); }
    public boolean contains(String genomeId) { return genomeMap.containsKey(genomeId); }
    /** the representatives in iteration order, as an array (NeighborProcessor.java:140, ListProcessor.java:62) */
    public RepGenome[] all() { return genomeMap.values().toArray(new RepGenome[0]); }
    public SortedSet<RepGenome> sorted() { return new TreeSet<>(genomeMap.values()); }
    public String getListFileName() { return String.format("rep%d.list.tbl", threshold); }
    @Override public String toString() { return "RepGen" + threshold; }

    @Override
    public Iterator<RepGenome> iterator() {
        Iterator<RepGenome> it = genomeMap.values().iterator();
        return new Iterator<>() {   // remove() must also invalidate the device copy (UpdateProcessor.java:96-103)
            public boolean hasNext() { return it.hasNext(); }
            public RepGenome next() { return it.next(); }
            public void remove() { it.remove(); rebuild = true; ranksDirty = true; }
        };
    }

    public void addRep(RepGenome rep) {
        if (genomeMap.put(rep.getGenomeId(), rep) != null) rebuild = true;   // replaced in place
        ranksDirty = true;
    }

    public Representation findClosest(ProteinKmers query) { return findClosest(List.of(query)).get(0); }
    public Representation findClosest(Sequence seq) { return findClosest(new ProteinKmers(seq.getSequence())); }
    public Representation findClosest(RepGenome rep) { return findClosest((ProteinKmers) rep); }   // GtoRepGenomeProcessor.java:271

    /** The batched scan: one launch for the whole list (RepGenomeDbProcessor.java:235-252 rewritten). */
    public List<Representation> findClosest(Collection<? extends ProteinKmers> queries) {
        List<Representation> out = new ArrayList<>(queries.size());
        if (queries.isEmpty()) return out;
        sync();
        List<ProteinKmers> q = new ArrayList<>(queries);
        List<String> prots = new ArrayList<>();
        for (ProteinKmers p : q) prots.add(p.getProtein());
        if (multi != null) return findClosestMulti(q, prots);
        try (Arena a = Arena.ofConfined(); KmersRepsNative.SeqSet set = KmersRepsNative.encode(kmerSize, prots)) {
            MemorySegment sim = a.allocate(JAVA_INT, q.size()), first = a.allocate(JAVA_INT, q.size());
            MemorySegment ctx = KmersRepsNative.context(kmerSize);
            KmersRepsNative.check(ctx, (int) KmersRepsNative.kr_db_find_closest.invokeExact(ctx, db, set.handle(), sim, first, MemorySegment.NULL));
            for (int i = 0; i < q.size(); i++) {
                int s = sim.getAtIndex(JAVA_INT, i), f = first.getAtIndex(JAVA_INT, i);
                if (s == 0 || f == KmersRepsNative.NO_REP) { out.add(new Representation()); continue; }
                RepGenome rep = onDevice.get(f);
                out.add(new Representation(rep, s, KmersRepsNative.distance(s, set.sizeOf(i), rep.size())));
            }
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
        return out;
    }

    /** findClosest over every GPU of the kr_multi handle: the library splits the queries; sizes come from the objects */
    private List<Representation> findClosestMulti(List<ProteinKmers> q, List<String> prots) {
        List<Representation> out = new ArrayList<>(q.size());
        ProteinKmers.encodeBatch(q);                                   // sizes for the distances (one launch, first context)
        try (Arena a = Arena.ofConfined()) {
            KmersRepsNative.Flat f = KmersRepsNative.flat(a, prots);
            MemorySegment sim = a.allocate(JAVA_INT, q.size()), first = a.allocate(JAVA_INT, q.size());
            KmersRepsNative.checkMulti(multi, (int) KmersRepsNative.kr_multi_db_find_closest.invokeExact(multi, db, f.residues, f.offsets, q.size(), sim, first, MemorySegment.NULL));
            for (int i = 0; i < q.size(); i++) {
                int s = sim.getAtIndex(JAVA_INT, i), fi = first.getAtIndex(JAVA_INT, i);
                if (s == 0 || fi == KmersRepsNative.NO_REP) { out.add(new Representation()); continue; }
                RepGenome rep = onDevice.get(fi);
                out.add(new Representation(rep, s, KmersRepsNative.distance(s, q.get(i).size(), rep.size())));
            }
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
        return out;
    }

    /**
     * Similarity of every query to EVERY representative, representatives in iteration order (this.all()): the
     * rewritten inner loop of NeighborProcessor.processGenome (:165-168), one rectangular GPU matrix for all queries.
     * Row q, column r = new Representation(all()[r], queries[q]).getSimilarity().
     */
    public int[][] similarities(List<? extends ProteinKmers> queries) {
        RepGenome[] reps = all();
        int[][] out = new int[queries.size()][reps.length];
        if (queries.isEmpty() || reps.length == 0) return out;
        List<String> q = new ArrayList<>(), r = new ArrayList<>();
        for (ProteinKmers p : queries) q.add(p.getProtein());
        for (RepGenome g : reps) r.add(g.getProtein());
        try (KmersRepsNative.SeqSet qs = KmersRepsNative.encode(kmerSize, q); KmersRepsNative.SeqSet rs = KmersRepsNative.encode(kmerSize, r)) {
            int[] cells = KmersRepsNative.simMatrix(qs, rs);
            for (int i = 0; i < q.size(); i++) System.arraycopy(cells, i * reps.length, out[i], 0, reps.length);
        }
        return out;
    }

    /**
     * The N closest representatives of one query with the reference's tie retention (NeighborProcessor.java:161-193):
     * positions into `row` (similarities in iteration order), best first; a new result goes before the entries it
     * ties with, and the list is cut back to `num` only when the entry at the break beats the one after it.
     */
    public static List<Integer> closestN(int[] row, int num, int minSim) {
        List<Integer> found = new ArrayList<>(num + 1);
        for (int r = 0; r < row.length; r++) {
            int sim = row[r];
            if (sim < minSim) continue;
            int idx = 0, n = found.size();
            while (idx < n && row[found.get(idx)] > sim) idx++;
            if (idx < n || n < num) {
                found.add(idx, r);
                if (found.size() > num && row[found.get(num - 1)] > row[found.get(num)])
                    while (found.size() > num) found.remove(num);
            }
        }
        return found;
    }

    public boolean checkSimilarity(ProteinKmers query, int sim) {
        return size() > 0 && findClosest(query).getSimilarity() >= sim;
    }

    public boolean checkGenome(RepGenome rep) { return checkGenomes(List.of(rep))[0]; }

    /** The greedy build for a list of candidates in list order (RepGenomeDbProcessor.java:134-147 rewritten). */
    public boolean[] checkGenomes(List<RepGenome> cands) {
        boolean[] added = new boolean[cands.size()];
        if (cands.isEmpty()) return added;
        sync();
        List<String> prots = new ArrayList<>();
        for (RepGenome c : cands) prots.add(c.getProtein());
        try (Arena a = Arena.ofConfined()) {
            MemorySegment flags = a.allocate(JAVA_BYTE, cands.size());
            if (multi != null) {
                KmersRepsNative.Flat f = KmersRepsNative.flat(a, prots);
                KmersRepsNative.checkMulti(multi, (int) KmersRepsNative.kr_multi_greedy_build.invokeExact(multi, db, f.residues, f.offsets, cands.size(), threshold, 0, flags));
            } else {
                try (KmersRepsNative.SeqSet set = KmersRepsNative.encode(kmerSize, prots)) {
                    MemorySegment ctx = KmersRepsNative.context(kmerSize);
                    KmersRepsNative.check(ctx, (int) KmersRepsNative.kr_greedy_build.invokeExact(ctx, db, set.handle(), threshold, 0, flags));
                }
            }
            for (int i = 0; i < added.length; i++)
                if (flags.get(JAVA_BYTE, i) != 0) {
                    added[i] = true;
                    RepGenome c = cands.get(i);
                    if (genomeMap.put(c.getGenomeId(), c) != null) rebuild = true; else onDevice.add(c);
                }
            ranksDirty = true;
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
        return added;
    }

    /** FASTA: ">Rep<thr>,K=<K> <protName>" with no sequence, then one record per representative (data/rep50.db). */
    public void save(File saveFile) throws IOException {
        try (PrintWriter out = new PrintWriter(saveFile, StandardCharsets.US_ASCII)) {
            out.print(String.format(">Rep%d,K=%d %s\n", threshold, kmerSize, protName));
            for (RepGenome rep : this) {
                out.print(">" + rep.getFid() + " " + rep.getName() + "\n");
                String p = rep.getProtein();
                for (int i = 0; i < p.length(); i += 60) out.print(p.substring(i, Math.min(p.length(), i + 60)) + "\n");
            }
        }
    }

    public static RepGenomeDb load(File loadFile) throws IOException {
        try (FastaInputStream in = new FastaInputStream(loadFile)) {
            Iterator<Sequence> it = in.iterator();
            Sequence header = it.next();
            String[] parts = header.getLabel().substring(3).split(",K=");
            ProteinKmers.setKmerSize(Integer.parseInt(parts[1]));
            RepGenomeDb db = new RepGenomeDb(Integer.parseInt(parts[0]), header.getComment());
            while (it.hasNext()) db.addRep(new RepGenome(it.next()));
            return db;
        }
    }

    /** Bring the device copy up to date: append new representatives, refresh iteration ranks. */
    private void sync() {
        MemorySegment ctx = KmersRepsNative.context(kmerSize);
        try (Arena a = Arena.ofConfined()) {
            if (rebuild && !db.equals(MemorySegment.NULL)) {
                if (multi != null) KmersRepsNative.kr_multi_db_free.invokeExact(db); else KmersRepsNative.kr_db_free.invokeExact(db);
                db = MemorySegment.NULL; onDevice.clear();
            }
            if (db.equals(MemorySegment.NULL)) {
                MemorySegment out = a.allocate(ADDRESS);
                if (multi != null) KmersRepsNative.checkMulti(multi, (int) KmersRepsNative.kr_multi_db_create.invokeExact(multi, out));
                else KmersRepsNative.check(ctx, (int) KmersRepsNative.kr_db_create.invokeExact(ctx, out));
                db = out.get(ADDRESS, 0);
            }
            rebuild = false;
            java.util.Set<RepGenome> have = java.util.Collections.newSetFromMap(new java.util.IdentityHashMap<>());
            have.addAll(onDevice);
            List<RepGenome> fresh = new ArrayList<>();
            for (RepGenome r : genomeMap.values()) if (!have.contains(r)) fresh.add(r);
            if (!fresh.isEmpty()) {
                long total = 0;
                for (RepGenome r : fresh) total += r.getProtein().length();
                MemorySegment res = a.allocate(Math.max(total, 1)), off = a.allocate(JAVA_LONG, fresh.size() + 1L);
                long pos = 0;
                for (int i = 0; i < fresh.size(); i++) {
                    byte[] b = fresh.get(i).getProtein().getBytes(StandardCharsets.ISO_8859_1);
                    MemorySegment.copy(b, 0, res, JAVA_BYTE, pos, b.length);
                    off.setAtIndex(JAVA_LONG, i, pos);
                    pos += b.length;
                }
                off.setAtIndex(JAVA_LONG, fresh.size(), pos);
                if (multi != null) KmersRepsNative.checkMulti(multi, (int) KmersRepsNative.kr_multi_db_append.invokeExact(multi, db, res, off, fresh.size()));
                else KmersRepsNative.check(ctx, (int) KmersRepsNative.kr_db_append.invokeExact(ctx, db, res, off, fresh.size()));
                onDevice.addAll(fresh);
                ranksDirty = true;
            }
            if (ranksDirty && !onDevice.isEmpty()) {
                java.util.IdentityHashMap<RepGenome, Integer> rankOf = new java.util.IdentityHashMap<>();
                int k = 0;
                for (RepGenome r : genomeMap.values()) rankOf.put(r, k++);   // the JVM's own HashMap order
                MemorySegment rank = a.allocate(JAVA_INT, onDevice.size());
                for (int d = 0; d < onDevice.size(); d++) rank.setAtIndex(JAVA_INT, d, rankOf.get(onDevice.get(d)));
                if (multi != null) KmersRepsNative.checkMulti(multi, (int) KmersRepsNative.kr_multi_db_set_ranks.invokeExact(multi, db, rank, onDevice.size()));
                else KmersRepsNative.check(ctx, (int) KmersRepsNative.kr_db_set_ranks.invokeExact(ctx, db, rank, onDevice.size()));
                ranksDirty = false;
            }
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }
}
