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

    public static SimilarityMatrix compute(Collection<? extends ProteinKmers> kmers) {
        List<String> prots = new ArrayList<>(kmers.size());
        int k = ProteinKmers.kmerSize();
        for (ProteinKmers p : kmers) { prots.add(p.getProtein()); k = p.k; }
        try (KmersRepsNative.SeqSet set = KmersRepsNative.encode(k, prots)) {
            int[] sizes = new int[prots.size()];
            for (int i = 0; i < sizes.length; i++) sizes[i] = set.sizeOf(i);
            return new SimilarityMatrix(prots.size(), prots.isEmpty() ? new int[0] : KmersRepsNative.simMatrix(set, null), sizes);
        }
    }

    public int size() { return n; }
    /** SequenceKmers.similarity of elements i and j (INFINITY for equal sequences) */
    public int similarity(int i, int j) { return sim[i * n + j]; }
    /** SequenceKmers.distance of elements i and j */
    public double distance(int i, int j) { return KmersRepsNative.distance(sim[i * n + j], sizes[i], sizes[j]); }
}
