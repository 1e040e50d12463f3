/*
 * kr_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the protein k-mer
 * similarity path of SEEDtk/kmers.reps.  Nothing under kmers/ (the product) may
 * link, load or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py do, and only as the checker
 * or as the timed CPU baseline.
 *
 * PARITY UNPINNED.  The arithmetic of this path (SequenceKmers / ProteinKmers /
 * RepGenome / RepGenomeDb) lives in the un-vendored Maven modules
 * org.theseed:shared:1.0.0 and org.theseed:sequence:1.0.0 (reference pom.xml:80-94),
 * which are absent from /root/reference, and there is no JVM in this image.  The
 * reference's own tests assert no similarity or distance value
 * (src/test/java/org/theseed/proteins/kmers/reps/ProtTest.java:175-219 checks
 * self-consistency only).  The restatement therefore follows the published
 * algorithm of those classes as used at the reference's call sites, and is
 * anchored on data the reference ships (tests/test_oracle_golden.py):
 *   - data/rep50.db record order == java.util.HashMap<String genomeId,..> order,
 *   - every pair of the 1480 reps of data/rep50.db has similarity <= 49 at K=8
 *     (the stingy-add invariant of a threshold-50 database),
 *   - save(load(data/rep50.db)) is byte-identical.
 *
 * What each function follows:
 *   kro_set_*      ProteinKmers(String): HashSet<String> of all length-K substrings,
 *                  no alphabet filtering; call sites SequenceInfo.java:47,
 *                  RepGenomeDbProcessor.java:161, GtoClassProcessor.java:129.
 *   kro_similarity SequenceKmers.similarity(other): number of other's distinct
 *                  k-mers contained in this; INFINITY (9999) for equal sequences.
 *                  Uses: DistanceMatrixProcessor.java:76, RepGenomeDbProcessor.java:249,
 *                  ScoreCouplingReporter.java:66-70.
 *   kro_distance   1 - sim/(|A|+|B|-sim): ScoreCouplingReporter.java:72-73,
 *                  SsuRepresentation.java:46-47, DistanceMatrixProcessor.java:75.
 *   kro_find_closest / kro_greedy_build
 *                  RepGenomeDb.findClosest / checkGenome+addRep as driven by
 *                  RepGenomeDbProcessor.java:119-169,235-252.
 *   kro_java_hashmap_order
 *                  java.util.HashMap<String,V> iteration order (OpenJDK 8+),
 *                  which RepGenomeDb.save() and findClosest tie-breaking inherit
 *                  (data/rep50.db order).
 *
 * The data structure mirrors the reference's choice on purpose (a hash set of
 * substrings probed key by key) so that timing it is a fair "port" baseline.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define KRO_INFINITY 9999

typedef struct kro_set {
    const char *seq;     /* not owned */
    int len, K;
    int nk;              /* number of distinct k-mers */
    uint32_t cap;        /* power of two */
    int32_t *slot;       /* offset of the k-mer stored in the slot, or -1 */
    uint32_t *slot_hash; /* cached String.hashCode of that k-mer */
    int32_t *kmers;      /* offsets of the distinct k-mers, in first-seen order */
    uint32_t *kmer_hash;
} kro_set;

/* java.lang.String.hashCode over K chars */
static inline uint32_t jhash(const char *s, int n) {
    uint32_t h = 0;
    for (int i = 0; i < n; i++) h = 31u * h + (uint8_t)s[i];
    return h;
}
/* java.util.HashMap.hash spread */
static inline uint32_t jspread(uint32_t h) { return h ^ (h >> 16); }

kro_set *kro_set_new(const char *seq, int len, int K) {
    kro_set *s = (kro_set *)calloc(1, sizeof(kro_set));
    s->seq = seq; s->len = len; s->K = K;
    int n = len - K + 1; if (n < 0) n = 0;
    uint32_t cap = 16; while (cap < (uint32_t)(2 * n + 1)) cap <<= 1;
    s->cap = cap;
    s->slot = (int32_t *)malloc(sizeof(int32_t) * cap);
    s->slot_hash = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    s->kmers = (int32_t *)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
    s->kmer_hash = (uint32_t *)malloc(sizeof(uint32_t) * (n > 0 ? n : 1));
    for (uint32_t i = 0; i < cap; i++) s->slot[i] = -1;
    for (int p = 0; p < n; p++) {
        uint32_t h = jhash(seq + p, K);
        uint32_t i = jspread(h) & (cap - 1);
        int found = 0;
        while (s->slot[i] >= 0) {
            if (s->slot_hash[i] == h && memcmp(seq + s->slot[i], seq + p, K) == 0) { found = 1; break; }
            i = (i + 1) & (cap - 1);
        }
        if (!found) {
            s->slot[i] = p; s->slot_hash[i] = h;
            s->kmers[s->nk] = p; s->kmer_hash[s->nk] = h; s->nk++;
        }
    }
    return s;
}

void kro_set_free(kro_set *s) {
    if (!s) return;
    free(s->slot); free(s->slot_hash); free(s->kmers); free(s->kmer_hash); free(s);
}

int kro_set_size(const kro_set *s) { return s->nk; }

static inline int set_contains(const kro_set *s, const char *kmer, uint32_t h) {
    uint32_t i = jspread(h) & (s->cap - 1);
    while (s->slot[i] >= 0) {
        if (s->slot_hash[i] == h && memcmp(s->seq + s->slot[i], kmer, s->K) == 0) return 1;
        i = (i + 1) & (s->cap - 1);
    }
    return 0;
}

/* raw shared distinct k-mer count: probe `a` with every distinct k-mer of `b` */
int kro_set_raw_similarity(const kro_set *a, const kro_set *b) {
    int c = 0;
    for (int i = 0; i < b->nk; i++)
        c += set_contains(a, b->seq + b->kmers[i], b->kmer_hash[i]);
    return c;
}

int kro_set_identical(const kro_set *a, const kro_set *b) {
    return a->len == b->len && memcmp(a->seq, b->seq, (size_t)a->len) == 0;
}

/* SequenceKmers.similarity: ident_inf != 0 selects the "equal sequences => INFINITY" rule */
int kro_set_similarity(const kro_set *a, const kro_set *b, int ident_inf) {
    if (ident_inf && kro_set_identical(a, b)) return KRO_INFINITY;
    return kro_set_raw_similarity(a, b);
}

/* SequenceKmers.distance(sim, a, b) */
double kro_distance(int sim, int size_a, int size_b) {
    if (sim == KRO_INFINITY) return 0.0;
    if (sim == 0) return 1.0;
    return 1.0 - (sim / (double)(size_a + size_b - sim));
}

/* convenience for one-off pairs */
int kro_similarity(const char *a, int la, const char *b, int lb, int K, int ident_inf) {
    kro_set *sa = kro_set_new(a, la, K), *sb = kro_set_new(b, lb, K);
    int r = kro_set_similarity(sa, sb, ident_inf);
    kro_set_free(sa); kro_set_free(sb);
    return r;
}

/* ---- batches over a flat residue buffer: seq i = res[off[i] .. off[i+1]) ---------- */

typedef struct kro_batch { int n; kro_set **sets; } kro_batch;

kro_batch *kro_batch_new(const char *res, const uint64_t *off, int n, int K) {
    kro_batch *b = (kro_batch *)malloc(sizeof(kro_batch));
    b->n = n; b->sets = (kro_set **)malloc(sizeof(kro_set *) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) b->sets[i] = kro_set_new(res + off[i], (int)(off[i + 1] - off[i]), K);
    return b;
}
void kro_batch_free(kro_batch *b) {
    if (!b) return;
    for (int i = 0; i < b->n; i++) kro_set_free(b->sets[i]);
    free(b->sets); free(b);
}
void kro_batch_sizes(const kro_batch *b, int32_t *sizes) {
    for (int i = 0; i < b->n; i++) sizes[i] = b->sets[i]->nk;
}

/* rows [r0,r1) of the |A| x |B| similarity matrix, row-major into out[(i-r0)*nB + j].
 * Follows the double loop of DistanceMatrixProcessor.java:71-76 / RepMatrixProcessor.java:88-93. */
void kro_sim_rows(const kro_batch *A, const kro_batch *B, int r0, int r1, int ident_inf, int32_t *out) {
    for (int i = r0; i < r1; i++)
        for (int j = 0; j < B->n; j++)
            out[(size_t)(i - r0) * B->n + j] = kro_set_similarity(A->sets[i], B->sets[j], ident_inf);
}

/* strictly-upper triangle (i<j) of the self matrix, packed row by row; rows [r0,r1) */
void kro_sim_triangle_rows(const kro_batch *A, int r0, int r1, int ident_inf, int32_t *out) {
    size_t k = 0;
    for (int i = r0; i < r1; i++)
        for (int j = i + 1; j < A->n; j++)
            out[k++] = kro_set_similarity(A->sets[i], A->sets[j], ident_inf);
}

typedef struct { const kro_batch *A, *B; int r0, r1, ident_inf; int32_t *out; int tri; uint64_t count; } mt_job;
static void *mt_run(void *p) {
    mt_job *j = (mt_job *)p;
    if (j->tri) {
        uint64_t c = 0; int64_t sum = 0;
        for (int i = j->r0; i < j->r1; i++)
            for (int k = i + 1; k < j->A->n; k++) { sum += kro_set_similarity(j->A->sets[i], j->A->sets[k], j->ident_inf); c++; }
        j->count = c; j->out[0] = (int32_t)(sum & 0x7fffffff);
    } else kro_sim_rows(j->A, j->B, j->r0, j->r1, j->ident_inf, j->out);
    return 0;
}

/* multi-threaded full matrix (rows interleaved in chunks over nthreads) */
void kro_sim_matrix_mt(const kro_batch *A, const kro_batch *B, int ident_inf, int nthreads, int32_t *out) {
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    mt_job *jobs = (mt_job *)malloc(sizeof(mt_job) * nthreads);
    int per = (A->n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int r0 = t * per, r1 = r0 + per; if (r0 > A->n) r0 = A->n; if (r1 > A->n) r1 = A->n;
        jobs[t] = (mt_job){A, B, r0, r1, ident_inf, out + (size_t)r0 * B->n, 0, 0};
        pthread_create(&th[t], 0, mt_run, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
    free(th); free(jobs);
}

/* timing helper for the CPU baseline: all i<j pairs with i in rows dealt round-robin to
 * threads; returns the number of pairs evaluated and a checksum (sum of sims). */
uint64_t kro_all_pairs_count_mt(const kro_batch *A, int ident_inf, int nthreads, int64_t *checksum) {
    if (nthreads < 1) nthreads = 1;
    /* balance: thread t takes rows i with (i % nthreads == t) -- done by splitting into stripes */
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    mt_job *jobs = (mt_job *)malloc(sizeof(mt_job) * nthreads);
    int32_t *sums = (int32_t *)calloc(nthreads, sizeof(int32_t));
    /* stripes of equal pair count: row boundary b_t with cumulative pairs ~ t/nthreads */
    int n = A->n; double total = (double)n * (n - 1) / 2.0; int r = 0; double acc = 0;
    for (int t = 0; t < nthreads; t++) {
        int r0 = r; double want = total * (t + 1) / nthreads;
        while (r < n && acc < want) { acc += (double)(n - 1 - r); r++; }
        if (t == nthreads - 1) r = n;
        jobs[t] = (mt_job){A, A, r0, r, ident_inf, &sums[t], 1, 0};
        pthread_create(&th[t], 0, mt_run, &jobs[t]);
    }
    uint64_t c = 0; int64_t cs = 0;
    for (int t = 0; t < nthreads; t++) { pthread_join(th[t], 0); c += jobs[t].count; cs += sums[t]; }
    if (checksum) *checksum = cs;
    free(th); free(jobs); free(sums);
    return c;
}

/* RepGenomeDb.findClosest: scan the reps in iteration order `order` (a permutation of
 * 0..R-1: order[k] = index of the k-th rep visited); keep the FIRST strict maximum
 * (tie_last == 0) or the LAST maximum (tie_last != 0).  best_idx = -1 when every sim is 0. */
void kro_find_closest(const kro_batch *reps, const int32_t *order, const kro_batch *queries,
                      int q0, int q1, int ident_inf, int tie_last, int32_t *best_sim, int32_t *best_idx) {
    for (int q = q0; q < q1; q++) {
        int bs = 0, bi = -1;
        for (int k = 0; k < reps->n; k++) {
            int r = order ? order[k] : k;
            int s = kro_set_similarity(reps->sets[r], queries->sets[q], ident_inf);
            if (tie_last ? (s >= bs && s > 0) : (s > bs)) { bs = s; bi = r; }
        }
        best_sim[q] = bs; best_idx[q] = bi;
    }
}

typedef struct { const kro_batch *reps; const int32_t *order; const kro_batch *q; int q0, q1, ident_inf, tie_last; int32_t *bs, *bi; } fc_job;
static void *fc_run(void *p) { fc_job *j = (fc_job *)p; kro_find_closest(j->reps, j->order, j->q, j->q0, j->q1, j->ident_inf, j->tie_last, j->bs, j->bi); return 0; }
void kro_find_closest_mt(const kro_batch *reps, const int32_t *order, const kro_batch *queries, int ident_inf, int tie_last,
                         int nthreads, int32_t *best_sim, int32_t *best_idx) {
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    fc_job *jobs = (fc_job *)malloc(sizeof(fc_job) * nthreads);
    int per = (queries->n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int q0 = t * per, q1 = q0 + per; if (q0 > queries->n) q0 = queries->n; if (q1 > queries->n) q1 = queries->n;
        jobs[t] = (fc_job){reps, order, queries, q0, q1, ident_inf, tie_last, best_sim, best_idx};
        pthread_create(&th[t], 0, fc_run, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
    free(th); free(jobs);
}

/* Greedy representative-set build: RepGenomeDb.checkGenome for each candidate in the given
 * order (RepGenomeDbProcessor.java:160-163): a candidate is added iff no current rep has
 * similarity >= threshold (early exit on the first hit).  `initial` reps (indices into
 * cands marked added_in[i]!=0 on entry) are taken as already present.  added[i] = 1 when
 * candidate i became a rep.  Returns the number of reps; *pairs = similarity evaluations. */
int kro_greedy_build(const kro_batch *cands, int threshold, int ident_inf, uint8_t *added, uint64_t *pairs) {
    int n = cands->n, nreps = 0; uint64_t np = 0;
    int32_t *reps = (int32_t *)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        int represented = 0;
        for (int k = 0; k < nreps; k++) {
            np++;
            if (kro_set_similarity(cands->sets[reps[k]], cands->sets[i], ident_inf) >= threshold) { represented = 1; break; }
        }
        added[i] = (uint8_t)!represented;
        if (!represented) reps[nreps++] = i;
    }
    free(reps);
    if (pairs) *pairs = np;
    return nreps;
}

/* ---- java.util.HashMap<String,V> iteration order -------------------------------------
 * Emulates OpenJDK HashMap for String keys inserted one by one with put(): table created at
 * capacity 16 on first put (or `initial_cap` if > 0, rounded up to a power of two),
 * load factor 0.75, resize doubles and splits each bin preserving relative order, new keys
 * append at the bin tail, re-put of an existing key keeps its position.  Tree bins are NOT
 * emulated: returns -1 if any bin would be treeified (9th node in a bin at capacity >= 64).
 * keys: n NUL-terminated strings.  order_out[k] = index (first insertion) of the k-th key in
 * iteration order; returns the number of distinct keys. */
typedef struct hm_node { uint32_t hash; int32_t idx; struct hm_node *next; } hm_node;

/* HashMap.resize(): double the table, split every bin into lo/hi keeping relative order */
static hm_node **hm_resize(hm_node **tab, uint32_t *cap, uint32_t *thr) {
    uint32_t ocap = *cap, ncap = ocap << 1;
    hm_node **nt = (hm_node **)calloc(ncap, sizeof(hm_node *));
    for (uint32_t j = 0; j < ocap; j++) {
        hm_node *lo_h = 0, *lo_t = 0, *hi_h = 0, *hi_t = 0, *e = tab[j];
        while (e) {
            hm_node *nx = e->next; e->next = 0;
            if ((e->hash & ocap) == 0) { if (lo_t) lo_t->next = e; else lo_h = e; lo_t = e; }
            else { if (hi_t) hi_t->next = e; else hi_h = e; hi_t = e; }
            e = nx;
        }
        nt[j] = lo_h; nt[j + ocap] = hi_h;
    }
    free(tab);
    *thr = (ocap >= 16) ? (*thr << 1) : (uint32_t)((float)ncap * 0.75f);
    *cap = ncap;
    return nt;
}

int kro_java_hashmap_order(const char **keys, int n, int initial_cap, int32_t *order_out) {
    uint32_t cap = 0, thr = 0; int size = 0;
    hm_node **tab = 0; hm_node *pool = (hm_node *)malloc(sizeof(hm_node) * (n > 0 ? n : 1)); int npool = 0;
    uint32_t init = 16;
    if (initial_cap > 0) { init = 1; while (init < (uint32_t)initial_cap) init <<= 1; }
    for (int i = 0; i < n; i++) {
        const char *k = keys[i];
        uint32_t h = jspread(jhash(k, (int)strlen(k)));
        if (!tab) { cap = init; thr = (uint32_t)((float)cap * 0.75f); tab = (hm_node **)calloc(cap, sizeof(hm_node *)); }
        uint32_t b = h & (cap - 1);
        hm_node *p = tab[b], *last = 0; int depth = 0, dup = 0;
        while (p) { if (p->hash == h && strcmp(keys[p->idx], k) == 0) { dup = 1; break; } last = p; p = p->next; depth++; }
        if (dup) continue;                  /* put() of an existing key keeps its position */
        hm_node *nn = &pool[npool++]; nn->hash = h; nn->idx = i; nn->next = 0;
        if (last) last->next = nn; else tab[b] = nn;
        if (depth >= 8) {                   /* binCount >= TREEIFY_THRESHOLD-1  ->  treeifyBin() */
            if (cap < 64) tab = hm_resize(tab, &cap, &thr);   /* small table: resize instead */
            else { free(tab); free(pool); return -1; }        /* tree bins not emulated */
        }
        if ((uint32_t)(++size) > thr) tab = hm_resize(tab, &cap, &thr);
    }
    int k = 0;
    for (uint32_t j = 0; j < cap; j++) for (hm_node *e = tab ? tab[j] : 0; e; e = e->next) order_out[k++] = e->idx;
    free(tab); free(pool);
    return k;
}
