package org.theseed.sequence;

import java.util.Arrays;
import java.util.Iterator;
import java.util.NoSuchElementException;

/**
 * Drop-in for org.theseed.sequence.SequenceKmers (org.theseed:shared): the public members kmers.reps and its siblings
 * call (SURVEY.md appendix C) -- INFINITY, size, similarity, distance (both forms), iterator, retainAll, removeAll,
 * getLen, equals, hashCode.  The HashSet&lt;String&gt; of upstream is a sorted array of packed k-mers that lives on the
 * GPU (one kr_seqset per batch of objects); the k-mers come back to the JVM only when a caller iterates or edits the
 * set.  GenomeKmers and DnaKmers keep their upstream implementations (other alphabets; SURVEY.md 8f-4).
 * COMPILE-UNVERIFIED (no JDK in the build container): tools/check_java_shim.py checks every member use in the reference
 * tree against these declarations, and the same logic runs tested in kmers/reps_b200/host/theseed.hpp.
 */
public abstract class SequenceKmers implements Iterable<String> {

    /** similarity of identical sequences (RepGenomeDbProcessor.java:249, GtoRepGenomeProcessor.java:300) */
    public static final int INFINITY = KmersRepsNative.INFINITY;

    protected String sequence;

    /** engine key of the context this object's k-mers live in (the k-mer size for proteins) */
    protected abstract int engineKey();

    // The object's place on the device: element `index` of batch `set`.  Filled lazily (constructors never touch the
    // device: they run inside parallel streams, GtoRepBuildProcessor.java:127) by ensure() or by an encodeBatch call.
    KmersRepsNative.SeqSet set;
    int index;
    /** non-null once retainAll / removeAll changed the set: the sorted packed k-mers it now holds */
    private long[] edited;

    protected abstract void encodeSelf();

    boolean isEdited() { return this.edited != null; }

    synchronized void ensure() {
        if (this.set == null) {
            if (this.edited != null) {
                this.set = KmersRepsNative.fromKeys(engineKey(), new long[][] { this.edited }, java.util.List.of(this.sequence));
                this.index = 0;
            } else encodeSelf();
        }
    }

    /** number of distinct k-mers (ScoreCouplingReporter.java:73, TargetProcessor.java:168) */
    public int size() {
        if (this.edited != null) return this.edited.length;
        ensure();
        return this.set.sizes[this.index];
    }

    /** number of k-mers in common; INFINITY for identical sequences (ScoreCouplingReporter.java:66-70) */
    public int similarity(SequenceKmers other) {
        ensure(); other.ensure();
        return KmersRepsNative.similarity(this.set, this.index, other.set, other.index);
    }

    /** Jaccard distance (SequenceInfo.java:75, DistanceMatrixProcessor.java:75) */
    public double distance(SequenceKmers other) {
        return distance(this.similarity(other), this, other);
    }

    /** ScoreCouplingReporter.java:72-73, SsuRepresentation.java:46-47 */
    public static double distance(int similarity, SequenceKmers a, SequenceKmers b) {
        return KmersRepsNative.distance(similarity, a.size(), b.size());
    }

    public int getLen() { return this.sequence.length(); }   // ProteinDataFactory.java:308

    /** the sorted packed k-mers of this object (kr_seqset_keys) */
    long[] keys() {
        if (this.edited != null) return this.edited;
        ensure();
        return KmersRepsNative.keys(this.set, this.index);
    }

    /** the k-mers as strings, in packed-key order (TargetProcessor.java:193: for (String kmer : kmerSet)) */
    @Override
    public Iterator<String> iterator() {
        final long[] k = keys();
        final int key = engineKey();
        return new Iterator<>() {
            private int next = 0;
            @Override public boolean hasNext() { return next < k.length; }
            @Override public String next() {
                if (next >= k.length) throw new NoSuchElementException();
                return KmersRepsNative.decode(key, k[next++]);
            }
        };
    }

    /** keep only the k-mers also in the other object (TargetProcessor.java:166) */
    public void retainAll(SequenceKmers other) { edit(other, true); }

    /** drop the k-mers also in the other object (TargetProcessor.java:184) */
    public void removeAll(SequenceKmers other) { edit(other, false); }

    private synchronized void edit(SequenceKmers other, boolean keepCommon) {
        long[] a = keys(), b = other.keys(), out = new long[a.length];
        int n = 0, j = 0;
        for (long x : a) {                                       // merge of two ascending lists
            while (j < b.length && b[j] < x) j++;
            boolean common = j < b.length && b[j] == x;
            if (common == keepCommon) out[n++] = x;
        }
        this.edited = Arrays.copyOf(out, n);
        this.set = null;                                         // re-created from the key list on next use
    }

    /** same class, same sequence, same k-mers (SsuGenome.java:77-104 calls these through super) */
    @Override
    public boolean equals(Object obj) {
        if (this == obj) return true;
        if (obj == null || getClass() != obj.getClass()) return false;
        SequenceKmers o = (SequenceKmers) obj;
        if (!this.sequence.equals(o.sequence)) return false;
        return (this.edited == null && o.edited == null) || Arrays.equals(keys(), o.keys());
    }

    @Override
    public int hashCode() {
        return 31 * this.sequence.hashCode() + (this.edited == null ? 0 : Arrays.hashCode(this.edited));
    }
}
