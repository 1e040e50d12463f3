package org.theseed.sequence;

import java.util.ArrayList;
import java.util.Collection;
import java.util.Iterator;
import java.util.List;

/**
 * Drop-in for org.theseed.sequence.DnaKmers of org.theseed:shared (call sites SsuRepGenomeProcessor.java:88,
 * SeqCompProcessor.java:102, ProteinDataFactory.java:277-300, TargetProcessor.java:162-193; SsuGenome.java:22 extends it):
 * the K-mers of the lower-cased sequence and of its reverse complement.  Same structure as {@link ProteinKmers}; the
 * context is created with the DNA alphabet flag, which only library builds with the (still experimental) DNA encoder
 * accept -- with the default library the first use throws the library's message, and the upstream class should stay in
 * place.  The mutating set operations of TargetProcessor (retainAll / removeAll / iteration) are not covered.
 * COMPILE-UNVERIFIED (no JDK in the build container); the C++ twin is kmers/reps_b200/host/theseed.hpp (DnaKmers).
 */
public class DnaKmers extends SequenceKmers {

    private static int K = 15;

    public static void setKmerSize(int kSize) { K = kSize; }
    public static int kmerSize() { return K; }

    final int k;
    KmersRepsNative.SeqSet set;
    int index;

    public DnaKmers(String dna) {
        this.sequence = dna;
        this.k = K;
    }

    public String getDna() { return this.sequence; }

    /** Encode many objects with one kernel launch; already-encoded objects are left alone. */
    public static void encodeBatch(Collection<? extends DnaKmers> list) {
        List<DnaKmers> todo = new ArrayList<>();
        List<String> seqs = new ArrayList<>();
        for (DnaKmers p : list) if (p.set == null) { todo.add(p); seqs.add(p.sequence); }
        if (todo.isEmpty()) return;
        KmersRepsNative.SeqSet s = KmersRepsNative.encode(KmersRepsNative.DNA_KEY + todo.get(0).k, seqs);
        for (int i = 0; i < todo.size(); i++) { todo.get(i).set = s; todo.get(i).index = i; }
    }

    synchronized void ensure() { if (this.set == null) encodeBatch(List.of(this)); }

    @Override
    public int size() { ensure(); return this.set.sizes[this.index]; }

    @Override
    public int similarity(SequenceKmers other) {
        DnaKmers o = (DnaKmers) other;
        ensure(); o.ensure();
        return KmersRepsNative.similarity(this.set, this.index, o.set, o.index);
    }

    @Override
    public Iterator<String> iterator() {
        throw new UnsupportedOperationException("decode with kr_seqset_keys + kr_key_decode; see INTEGRATION.md");
    }
}
