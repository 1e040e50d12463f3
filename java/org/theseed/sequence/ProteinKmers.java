package org.theseed.sequence;

import java.util.ArrayList;
import java.util.Collection;
import java.util.List;

/**
 * Drop-in for org.theseed.sequence.ProteinKmers of org.theseed:shared (same FQCN, same public members as used
 * by kmers.reps: SequenceInfo.java:47,75, RepGenomeDbProcessor.java:77-80,121, GtoClassProcessor.java:129,
 * ScoreCouplingReporter.java:66-73, SeqTestProcessor.java:108,121).  The HashSet&lt;String&gt; of k-mers is replaced
 * by a sorted packed-key array that lives on the GPU; the constructor does not touch the device (it is called from
 * parallel streams, GtoRepBuildProcessor.java:127) -- encoding happens lazily, alone or in a batch via
 * {@link #encodeBatch}.
 * COMPILE-UNVERIFIED (no JDK in the build container); the same logic is implemented and tested in
 * kmers/reps_b200/host/theseed.hpp.
 */
public class ProteinKmers extends SequenceKmers {

    private static int K = 8;

    public static void setKmerSize(int kSize) { K = kSize; }
    public static int kmerSize() { return K; }

    private final String protein;
    final int k;

    public ProteinKmers(String protein) {
        this.protein = protein;
        this.sequence = protein;
        this.k = K;
    }

    public String getProtein() { return this.protein; }

    @Override protected int engineKey() { return this.k; }

    /** Encode many objects with one kernel launch; already-encoded objects are left alone. */
    public static void encodeBatch(Collection<? extends ProteinKmers> list) {
        List<ProteinKmers> todo = new ArrayList<>();
        List<String> seqs = new ArrayList<>();
        for (ProteinKmers p : list) {
            if (p.set != null) continue;
            if (p.isEdited()) p.ensure();                        // its k-mers are no longer those of the sequence: kr_seqset_from_keys
            else { todo.add(p); seqs.add(p.protein); }
        }
        if (todo.isEmpty()) return;
        KmersRepsNative.SeqSet s = KmersRepsNative.encode(todo.get(0).k, seqs);
        for (int i = 0; i < todo.size(); i++) { todo.get(i).set = s; todo.get(i).index = i; }
    }

    @Override protected void encodeSelf() { encodeBatch(List.of(this)); }
}
