package org.theseed.sequence;

import java.util.ArrayList;
import java.util.Collection;
import java.util.List;

/**
 * NEW class (not in upstream shared): all-vs-all similarities of a list of ProteinKmers in one GPU pass
 * (kr_sim_matrix).  It replaces the per-pair calls inside the double loops of
 * DistanceMatrixProcessor.java:71-76 and RepMatrixProcessor.java:88-93; the loops themselves stay, in the same
 * order, and only read this matrix, so every floating-point accumulation (SequenceInfo.java:83-89) happens in the
 * reference's order on the JVM.  COMPILE-UNVERIFIED (no JDK in the build container); the C++ twin is
 * theseed::SimilarityMatrix in kmers/reps_b200/host/theseed.hpp.
 */
public final class SimilarityMatrix {

    private final int n;
    private final int[] sim;     // row-major n x n
    private final int[] sizes;

    private SimilarityMatrix(int n, int[] sim, int[] sizes) { this.n = n; this.sim = sim; this.sizes = sizes; }

    /** one batch on the device for a list of objects: their sequences re-encoded together, or -- when some object's set was
     *  edited by retainAll / removeAll -- their key lists (kr_seqset_from_keys) */
    private static KmersRepsNative.SeqSet batch(Collection<? extends ProteinKmers> kmers) {
        List<String> prots = new ArrayList<>(kmers.size());
        int k = ProteinKmers.kmerSize();
        boolean edited = false;
        for (ProteinKmers p : kmers) { prots.add(p.getProtein()); k = p.k; edited |= p.isEdited(); }
        if (!edited) return KmersRepsNative.encode(k, prots);
        long[][] keys = new long[kmers.size()][];
        int i = 0;
        for (ProteinKmers p : kmers) keys[i++] = p.keys();
        return KmersRepsNative.fromKeys(k, keys, prots);
    }

    public static SimilarityMatrix compute(Collection<? extends ProteinKmers> kmers) {
        try (KmersRepsNative.SeqSet set = batch(kmers)) {
            int[] sizes = new int[kmers.size()];
            for (int i = 0; i < sizes.length; i++) sizes[i] = set.sizeOf(i);
            return new SimilarityMatrix(sizes.length, sizes.length == 0 ? new int[0] : KmersRepsNative.simMatrix(set, null), sizes);
        }
    }

    /**
     * The same matrix built from the cells with similarity >= 1 only (kr_sim_matrix_sparse): a few cells per sequence cross
     * PCIe instead of N * N values; every cell not returned is 0 by definition, the diagonal is the size (INFINITY under the
     * equal-sequence rule).  This is what `distances` / `group` use on large inputs: the callers' loops still visit every
     * pair in the reference's order (DistanceMatrixProcessor.java:71-83, SequenceInfo.java:83-89) and read zeros here.
     */
    public static SimilarityMatrix computeSparse(Collection<? extends ProteinKmers> kmers) {
        try (KmersRepsNative.SeqSet set = batch(kmers)) {
            int n = kmers.size();
            int[] sizes = new int[n];
            for (int i = 0; i < n; i++) sizes[i] = set.sizeOf(i);
            int[] sim = new int[n * n];
            for (int i = 0; i < n; i++) sim[i * n + i] = KmersRepsNative.INFINITY;      // an object and itself: identical sequences
            if (n > 0) {
                int[] cells = KmersRepsNative.simCells(set, null, 1);
                for (int c = 0; c < cells.length; c += 3) {
                    sim[cells[c] * n + cells[c + 1]] = cells[c + 2];
                    sim[cells[c + 1] * n + cells[c]] = cells[c + 2];
                }
            }
            return new SimilarityMatrix(n, sim, sizes);
        }
    }

    public int size() { return n; }
    /** SequenceKmers.similarity of elements i and j (INFINITY for equal sequences) */
    public int similarity(int i, int j) { return sim[i * n + j]; }
    /** SequenceKmers.distance of elements i and j */
    public double distance(int i, int j) { return KmersRepsNative.distance(sim[i * n + j], sizes[i], sizes[j]); }
}
