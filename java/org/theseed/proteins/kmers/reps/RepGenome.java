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
    @Override public int compareTo(RepGenome o) { return genomeId.compareTo(o.genomeId); }
    @Override public boolean equals(Object o) { return o instanceof RepGenome && ((RepGenome) o).genomeId.equals(genomeId); }
    @Override public int hashCode() { return genomeId.hashCode(); }
    @Override public String toString() { return genomeId + " (" + name + ")"; }
}
