package org.theseed.sequence;

/**
 * Drop-in for org.theseed.sequence.SequenceKmers (org.theseed:shared).  Only what the similarity path needs;
 * GenomeKmers keeps the upstream implementation, and so does DnaKmers until the DNA encoder leaves the experimental
 * library (java/org/theseed/sequence/DnaKmers.java is its replacement; SURVEY.md 8f-4).
 */
public abstract class SequenceKmers implements Iterable<String> {

    /** similarity of identical sequences (RepGenomeDbProcessor.java:249, GtoRepGenomeProcessor.java:300) */
    public static final int INFINITY = KmersRepsNative.INFINITY;

    protected String sequence;

    public abstract int size();
    public abstract int similarity(SequenceKmers other);

    /** Jaccard distance (SequenceInfo.java:75, DistanceMatrixProcessor.java:75) */
    public double distance(SequenceKmers other) {
        return distance(this.similarity(other), this, other);
    }

    /** ScoreCouplingReporter.java:72-73, SsuRepresentation.java:46-47 */
    public static double distance(int similarity, SequenceKmers a, SequenceKmers b) {
        return KmersRepsNative.distance(similarity, a.size(), b.size());
    }

    public int getLen() { return this.sequence.length(); }
}
