package org.theseed.proteins.kmers;

import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

import org.theseed.sequence.KmersRepsNative;
import org.theseed.sequence.ProteinKmers;
import org.theseed.sequence.Sequence;

/**
 * Drop-in for org.theseed.proteins.kmers.KmerCollectionGroup (org.theseed:shared) as used by
 * ClassifyProcessor.java:95,108,116: named groups of ProteinKmers; getDistance(seq, group) is the smallest
 * distance from seq to any member of the group (1.0 for an unknown or empty group).  getDistances is the batched
 * form the rewritten classify loop calls: one rectangular kr_sim_matrix per group.
 * COMPILE-UNVERIFIED (no JDK in the build container); C++ twin: theseed::KmerCollectionGroup.
 */
public class KmerCollectionGroup {

    private final Map<String, List<ProteinKmers>> groups = new HashMap<>();

    public KmerCollectionGroup() { }                               // ClassifyProcessor.java:71

    public void addSequence(Sequence seq, String group) {
        groups.computeIfAbsent(group, g -> new ArrayList<>()).add(new ProteinKmers(seq.getSequence()));
    }

    public int size() { return groups.size(); }

    public double getDistance(Sequence seq, String group) {
        return getDistances(List.of(seq), List.of(group))[0];
    }

    /** result of getBest: the closest group and the distance to it */
    public static final class Result {
        private final String group; private final double distance;
        Result(String group, double distance) { this.group = group; this.distance = distance; }
        public String getGroup() { return group; }
        public double getDistance() { return distance; }
    }

    /** the group closest to the sequence [recalled member; groups in map order, first strict minimum wins; null group at distance 1.0 when nothing is closer] */
    public Result getBest(Sequence seq) {
        String best = null; double bestDist = 1.0;
        for (String g : groups.keySet()) {
            double d = getDistance(seq, g);
            if (d < bestDist) { bestDist = d; best = g; }
        }
        return new Result(best, bestDist);
    }

    public double[] getDistances(List<Sequence> seqs, List<String> groupOf) {
        double[] out = new double[seqs.size()];
        java.util.Arrays.fill(out, 1.0);
        Map<String, List<Integer>> byGroup = new HashMap<>();
        for (int i = 0; i < seqs.size(); i++) byGroup.computeIfAbsent(groupOf.get(i), g -> new ArrayList<>()).add(i);
        for (Map.Entry<String, List<Integer>> e : byGroup.entrySet()) {
            List<ProteinKmers> members = groups.get(e.getKey());
            if (members == null || members.isEmpty()) continue;
            List<String> q = new ArrayList<>(), m = new ArrayList<>();
            for (int i : e.getValue()) q.add(seqs.get(i).getSequence());
            for (ProteinKmers p : members) m.add(p.getProtein());
            int k = ProteinKmers.kmerSize();
            try (KmersRepsNative.SeqSet qs = KmersRepsNative.encode(k, q); KmersRepsNative.SeqSet ms = KmersRepsNative.encode(k, m)) {
                int[] sim = KmersRepsNative.simMatrix(qs, ms);
                for (int a = 0; a < q.size(); a++) {
                    double best = 1.0;
                    for (int b = 0; b < m.size(); b++)
                        best = Math.min(best, KmersRepsNative.distance(sim[a * m.size() + b], qs.sizeOf(a), ms.sizeOf(b)));
                    out[e.getValue().get(a)] = best;
                }
            }
        }
        return out;
    }
}
