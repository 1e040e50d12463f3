package org.theseed.proteins.kmers.reps;

import java.util.regex.Matcher;
import java.util.regex.Pattern;

import org.theseed.sequence.ProteinKmers;
import org.theseed.sequence.Sequence;

/**
 * Drop-in for org.theseed.proteins.kmers.reps.RepGenome (org.theseed:shared): a representative genome's seed
 * protein.  Members as used by kmers.reps: RepGenomeDbProcessor.java:161,218, BaseGenomeProcessor.java:193-202,314,
 * DistanceMatrixProcessor.java:74, UpdateProcessor.java:101.  COMPILE-UNVERIFIED (no JDK in the build container).
 */
public class RepGenome extends ProteinKmers implements Comparable<RepGenome> {

    private static final Pattern GENOME_ID = Pattern.compile("fig\\|(\\d+\\.\\d+)\\.");
    private final String fid, name, genomeId;

    public RepGenome(Sequence seq) { this(seq.getLabel(), seq.getComment(), seq.getSequence()); }

    public RepGenome(String fid, String name, String protein) {
        super(protein);
        this.fid = fid;
        this.name = name;
        Matcher m = GENOME_ID.matcher(fid);
        this.genomeId = m.lookingAt() ? m.group(1) : fid;
    }

    public String getFid() { return fid; }
    public String getName() { return name; }
    public String getGenomeId() { return genomeId; }
    /**
     * Order of genome ids.  Upstream's compareTo cannot be read here (un-vendored module), so the order is a policy:
     * String.compareTo of the ids (default) or, with -Dkmersreps.idOrder=natural, taxon number then version number.
     * Same rule as RepGenome::compareIds in kmers/reps_b200/host/theseed.hpp.  Only the orientation / order of the
     * rows of `distances` depends on it (DistanceMatrixProcessor.java:74), never a value.
     */
    private static final boolean NATURAL_ORDER = "natural".equals(System.getProperty("kmersreps.idOrder"));

    public static int compareIds(String a, String b) {
        if (NATURAL_ORDER) {
            String[] fa = a.split("\\.", -1), fb = b.split("\\.", -1);
            for (int i = 0; i < Math.min(fa.length, fb.length); i++) {
                boolean na = !fa[i].isEmpty() && fa[i].chars().allMatch(c -> c >= '0' && c <= '9');
                boolean nb = !fb[i].isEmpty() && fb[i].chars().allMatch(c -> c >= '0' && c <= '9');
                int c = (na && nb) ? new java.math.BigInteger(fa[i]).compareTo(new java.math.BigInteger(fb[i])) : fa[i].compareTo(fb[i]);
                if (c != 0) return c < 0 ? -1 : 1;
            }
            if (fa.length != fb.length) return fa.length < fb.length ? -1 : 1;
        }
        return Integer.signum(a.compareTo(b));
    }

    @Override public int compareTo(RepGenome o) { return compareIds(genomeId, o.genomeId); }
    @Override public boolean equals(Object o) { return o instanceof RepGenome && ((RepGenome) o).genomeId.equals(genomeId); }
    @Override public int hashCode() { return genomeId.hashCode(); }
    @Override public String toString() { return genomeId + " (" + name + ")"; }
}
