/*
 * kmersreps.h -- C ABI of libkmersreps_cuda.so: the B200 (sm_100a) implementation of the
 * protein k-mer similarity hot path of SEEDtk/kmers.reps.
 *
 * The reference has no FFI of its own: the boundary it exposes for this path is the public Java
 * API of org.theseed.sequence.{SequenceKmers,ProteinKmers} and
 * org.theseed.proteins.kmers.reps.{RepGenome,RepGenomeDb} (un-vendored module
 * org.theseed:shared:1.0.0, reference pom.xml:80-94).  Each entry point below names the Java
 * member(s) it replaces and the reference call site (file:line under
 * src/main/java/org/theseed/proteins/kmers/reps/ unless a longer path is given) that drives it.
 * INTEGRATION.md shows the Panama-FFM binding a maintainer adds on the Java side.
 *
 * Conventions
 *   - plain C types only; every function returns a kr_status (0 = OK, < 0 = error) unless noted;
 *     the message of the last error is kr_last_error(ctx) (ctx may be NULL for creation errors).
 *   - host buffers are caller-owned and only touched during the call; kr_* handles are opaque,
 *     library-owned, and freed by their _free / _destroy; "_device" variants take device pointers
 *     of the context's GPU instead of host pointers and do not synchronise the host.
 *   - there is NO CPU fallback: kr_create fails when no CUDA device is usable (reference
 *     convention for fatal errors: RuntimeException, App.java:129).
 *   - a kr_ctx is bound to one GPU and is internally serialised: it may be called from any
 *     thread (NeighborProcessor.java:143-145 and GtoRepBuildProcessor.java:127 use parallel
 *     streams).  Several GPUs: either one context per GPU driven by the caller (one process or
 *     thread each, with the stripe / shard arguments below -- what bench.py does under torchrun),
 *     or ONE kr_multi handle (last section) that owns a context and a worker thread per GPU and
 *     shards every call internally.  The path has no data-path exchange between GPUs.
 *   - sequences are passed as one flat residue buffer plus offsets: sequence i is
 *     residues[offsets[i] .. offsets[i+1]).  Residues are the characters of the Java String;
 *     'A'..'Z', '*' and '-' are accepted (5-bit packing must stay injective);
 *     any other byte fails the call with KR_ERR_ALPHABET and names the byte.
 */
#ifndef KMERSREPS_H
#define KMERSREPS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kr_ctx kr_ctx;
typedef struct kr_seqset kr_seqset; /* a batch of ProteinKmers objects, encoded and resident in HBM */
typedef struct kr_db kr_db;         /* the representatives of a RepGenomeDb, resident in HBM */

typedef int32_t kr_status;
#define KR_OK 0
#define KR_ERR_ARG (-1)      /* bad argument (null pointer, K out of range, index out of range ...) */
#define KR_ERR_CUDA (-2)     /* a CUDA call failed; message has the CUDA error string */
#define KR_ERR_NODEVICE (-3) /* no usable CUDA device: there is no CPU fallback */
#define KR_ERR_ALPHABET (-4) /* a residue outside 'A'..'Z', '*', '-' */
#define KR_ERR_STATE (-5)    /* handle belongs to another context / K changed / destroyed */

/* SequenceKmers.INFINITY (RepGenomeDbProcessor.java:249, GtoRepGenomeProcessor.java:300) */
#define KR_INFINITY 9999
#define KR_NO_REP 0xFFFFFFFFu

/* context flags (policy switches for the semantics SURVEY.md 8c could not verify against a JVM) */
#define KR_IDENT_IS_INFINITY 1u /* similarity of two equal sequences is KR_INFINITY, not the raw count */
#define KR_TIE_LAST 2u          /* findClosest keeps the last maximum in iteration order, not the first */
/* Similarity matrices, findClosest and the greedy build each have two exact device paths: the tile join (work
 * proportional to the number of pairs; any set) and the sorted inverted index with its warp-per-query scan (work
 * proportional to the k-mers the pairs share; sets that keep key lists, 5K + id bits <= 64).  The library picks per
 * call from the index's posting-list statistics; these flags pin one path (the parity tests run both). */
#define KR_FORCE_JOIN 4u
#define KR_FORCE_INDEX 8u

/* K = 1..11.  The reference's validators accept any K >= 1 (SeqTestProcessor.java:118, BuildRepDbProcessor.java:90,
 * GtoRepBuildProcessor.java:98); 12 and above are refused by kr_create with a message: 5 bits per residue leave no room
 * beside a 60-bit key for the sequence id of an index entry or the block-local id of a block-list entry in 64 bits. */
#define KR_MIN_K 1
#define KR_MAX_K 11

/* ---- context -------------------------------------------------------------------------------
 * Replaces: static state of ProteinKmers -- ProteinKmers.setKmerSize(int) / kmerSize()
 * (RepGenomeDbProcessor.java:77-80,121; RepMatrixProcessor.java:51-54). */
kr_status kr_create(int device, int K, uint32_t flags, kr_ctx **out);
void kr_destroy(kr_ctx *ctx);
/* message of the context's last error (ctx == NULL: of this thread's last failed kr_create), copied into a buffer of the
 * calling thread: valid until this thread calls kr_last_error again, whatever other threads do with the context */
const char *kr_last_error(const kr_ctx *ctx);
int kr_kmer_size(const kr_ctx *ctx);
/* build tag of the library ("default": there is one build) */
const char *kr_build_info(void);
/* block until every kernel and copy the context issued has finished */
kr_status kr_synchronize(kr_ctx *ctx);

/* ---- ProteinKmers(String) in batches -------------------------------------------------------
 * Replaces: new ProteinKmers(String) / new RepGenome(Sequence) (SequenceInfo.java:47,
 * RepGenomeDbProcessor.java:161, GtoClassProcessor.java:129, ProteinData.java:211): the set of
 * distinct length-K substrings, stored as a sorted array of packed 64-bit keys per sequence. */
kr_status kr_encode(kr_ctx *ctx, const uint8_t *residues, const uint64_t *offsets, uint32_t nseq, kr_seqset **out);
/* same, residues already on the context's device (offsets stay a host array) */
kr_status kr_encode_device(kr_ctx *ctx, const uint8_t *d_residues, const uint64_t *offsets, uint32_t nseq, kr_seqset **out);
void kr_seqset_free(kr_seqset *set);
uint32_t kr_seqset_count(const kr_seqset *set);
/* SequenceKmers.size() for every sequence (ScoreCouplingReporter.java:73) */
kr_status kr_seqset_sizes(const kr_seqset *set, uint32_t *sizes);
/* SequenceKmers.iterator(): the sorted packed keys of sequence i (TargetProcessor.java:193);
 * kr_key_decode turns one key back into its K characters */
kr_status kr_seqset_keys(const kr_seqset *set, uint32_t i, uint64_t *keys, uint32_t cap, uint32_t *n);
kr_status kr_key_decode(const kr_ctx *ctx, uint64_t key, char *out /* K bytes, not NUL-terminated */);
/* new set holding sequences idx[0..n) of `set` (used to batch sub-lists of ProteinKmers objects) */
kr_status kr_seqset_gather(kr_ctx *ctx, const kr_seqset *set, const uint32_t *idx, uint32_t n, kr_seqset **out);
/* SequenceKmers.retainAll(other) / removeAll(other) (TargetProcessor.java:166,184) leave objects whose k-mer set is no
 * longer "all substrings of the sequence".  The shim merges the two sorted key lists it read with kr_seqset_keys and
 * makes the result a set again with this call, so everything downstream still runs on the GPU.  keys: strictly
 * ascending packed k-mers per sequence (key_offsets[nseq + 1]); residues / res_offsets: the objects' unchanged
 * sequences (what upstream's equal-sequence test compares), or NULL / NULL: then an object equals only itself.
 * KR_INFINITY is reported for two objects whose sequences AND k-mer sets are equal (the device examines only pairs
 * whose count equals both sizes); an object whose set really changed is no longer "identical" to a plain object of
 * the same sequence -- the one corner where this differs from a string comparison, and no reference caller takes a
 * similarity after retainAll / removeAll. */
kr_status kr_seqset_from_keys(kr_ctx *ctx, const uint64_t *keys, const uint64_t *key_offsets, const uint8_t *residues,
                              const uint64_t *res_offsets, uint32_t nseq, kr_seqset **out);

/* ---- multi-GPU: exchanging encoded block lists ------------------------------------------------
 * The tile join reads only a set's block lists.  With several GPUs each one encodes a contiguous slice of the
 * batch (kr_encode_device on offsets + begin; slices must start at multiples of KR_TILE), exports its block
 * lists as flat device arrays, the caller all-gathers those arrays (NCCL), and every GPU imports the
 * concatenation as one set.  An imported set supports kr_sim_matrix* and kr_db_find_closest* as queries via the
 * join; it has no per-sequence key lists (kr_sim_pairs, kr_seqset_keys, kr_seqset_gather refuse it). */
typedef struct kr_blocklists {
    uint32_t nseq;      /* sequences in the shard */
    uint32_t nblk;      /* blocks of KR_TILE sequences */
    uint64_t entries;   /* length of hi / lo, including padding */
    uint32_t lo_size;   /* bytes per lo element (2 or 4) */
    uint32_t nchunk;    /* key-range chunks per block */
    const void *hi;        /* uint32[entries] */
    const void *lo;        /* lo_size * entries bytes */
    const void *chunk_off; /* uint32[nblk * (nchunk + 1)] */
    const void *chunk_cnt; /* uint32[nblk * nchunk] */
    const void *blk_base;  /* uint64[nblk + 1], relative to hi / lo */
    const void *meta;      /* 24 bytes per sequence: size, length, hash, (residue pointer: rewritten on import) */
} kr_blocklists;
/* allocated length (elements) of the hi / lo arrays of a set encoded from these offsets: an upper bound of
 * `entries` every GPU can compute without communication (pure host arithmetic) */
uint64_t kr_blocklist_capacity(const kr_ctx *ctx, const uint64_t *offsets, uint32_t nseq);
/* device pointers into the set's own arrays; valid until the set is freed; synchronises the host */
kr_status kr_seqset_export_blocklists(kr_ctx *ctx, kr_seqset *set, kr_blocklists *out);
/* shards: nshards descriptors whose pointers are on this context's device, in batch order; residues / offsets
 * describe the WHOLE batch (d_residues on the device, offsets on the host) */
kr_status kr_seqset_import_blocklists(kr_ctx *ctx, const kr_blocklists *shards, uint32_t nshards, const uint8_t *d_residues,
                                      const uint64_t *offsets, uint32_t nseq, kr_seqset **out);

/* ---- SequenceKmers.similarity / distance ---------------------------------------------------
 * Replaces: a.similarity(b) for a list of pairs (DistanceMatrixProcessor.java:76,
 * ScoreCouplingReporter.java:66-70, NeighborProcessor.java:168, KmerCollectionGroup.getDistance
 * via ClassifyProcessor.java:116).  sim[p] is the raw shared-k-mer count, or KR_INFINITY for equal
 * sequences under KR_IDENT_IS_INFINITY; identical[p] (may be NULL) is 1 for equal sequences under
 * either policy. */
kr_status kr_sim_pairs(kr_ctx *ctx, const kr_seqset *A, const uint32_t *ia, const kr_seqset *B, const uint32_t *ib,
                       uint64_t npairs, int32_t *sim, uint8_t *identical);
/* SequenceKmers.distance(int sim, a, b): 1 - sim/(|a|+|b|-sim); 1.0 for sim 0; 0.0 for INFINITY
 * (ScoreCouplingReporter.java:72-73, SsuRepresentation.java:46-47).  Pure host arithmetic. */
double kr_distance(int32_t sim, uint32_t size_a, uint32_t size_b);

/* ---- all-vs-all matrices -------------------------------------------------------------------
 * Replaces: the double loops of DistanceMatrixProcessor.java:71-76 and RepMatrixProcessor.java:88-93
 * (+ SequenceInfo.java:62-76).  Rows are sequences of A, columns sequences of B (B == NULL: A
 * against itself). */
#define KR_ELEM_I32 0u
#define KR_ELEM_U16 1u /* every similarity fits: counts are < 65536 and KR_INFINITY is 9999 */
#define KR_LAYOUT_FULL 0u     /* row-major, leading dimension = number of columns */
#define KR_LAYOUT_TRIANGLE 1u /* self only: cells j > i, packed row after row (N(N-1)/2 cells) */
#define KR_LAYOUT_BLOCKTRI 2u /* self only: row i holds columns j >= 256*(i/256), row-major per
                                 256-row stripe with leading dimension N - 256*(i/256) */
#define KR_TILE 256u          /* rows per stripe unit; row_begin must be a multiple of it */

typedef struct kr_matrix_desc {
    uint32_t elem;      /* KR_ELEM_* */
    uint32_t layout;    /* KR_LAYOUT_* */
    uint32_t row_begin; /* first row computed by this call (multiple of KR_TILE) */
    uint32_t row_end;   /* one past the last row (multiple of KR_TILE, or the row count) */
} kr_matrix_desc;

/* number of elements rows [row_begin,row_end) occupy in the given layout (n_rows x n_cols matrix) */
uint64_t kr_matrix_elems(uint32_t n_rows, uint32_t n_cols, uint32_t layout, uint32_t row_begin, uint32_t row_end);
/* split rows into nshards stripes of (nearly) equal cell count, boundaries multiples of KR_TILE:
 * row_bounds[0..nshards] */
kr_status kr_matrix_stripes(uint32_t n_rows, uint32_t n_cols, uint32_t layout, uint32_t nshards, uint32_t *row_bounds);
/* out receives the cells of rows [row_begin,row_end), starting at out[0] */
kr_status kr_sim_matrix(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, void *out_host);
kr_status kr_sim_matrix_device(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, void *out_dev);

/* The same matrix as a list of its cells with sim >= min_sim (min_sim >= 1): four uint32 per cell -- row, column,
 * similarity, 0 -- in no particular order.  Layout FULL: every column; TRIANGLE (self only): columns j > i.  All other
 * cells are 0 by definition, so the callers that walk the whole matrix (DistanceMatrixProcessor.java:71-83 prints every
 * pair, SequenceInfo.java:62-89 sums every distance) regenerate the zero cells on the host in the reference's order
 * instead of moving N^2 values over PCIe.  *count receives the number of qualifying cells even when it exceeds
 * `cap`; then only `cap` cells are returned and the status is KR_ERR_ARG (call again with room). */
kr_status kr_sim_matrix_sparse(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, int32_t min_sim,
                               uint32_t *cells_host, uint64_t cap, uint64_t *count);
/* same, cells left in device memory (16 bytes each; used by the multi-GPU result gather); synchronises the host to
 * return the count */
kr_status kr_sim_matrix_sparse_device(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, int32_t min_sim,
                                      void *cells_dev, uint64_t cap, uint64_t *count);
/* The same with PACKED cells, 8 bytes each: cell = row << 40 | col << 16 | sim (KR_CELL_ROW / _COL / _SIM).  Half the bytes
 * over PCIe for callers that fetch millions of cells; both sets must hold at most 2^24 sequences (KR_ERR_ARG otherwise). */
#define KR_CELL_ROW(c) ((uint32_t)((c) >> 40))
#define KR_CELL_COL(c) ((uint32_t)(((c) >> 16) & 0xFFFFFFu))
#define KR_CELL_SIM(c) ((int32_t)((c) & 0xFFFFu))
kr_status kr_sim_matrix_sparse64(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, int32_t min_sim,
                                 uint64_t *cells_host, uint64_t cap, uint64_t *count);
kr_status kr_sim_matrix_sparse64_device(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc,
                                        int32_t min_sim, void *cells_dev, uint64_t cap, uint64_t *count);

/* ---- RepGenomeDb ---------------------------------------------------------------------------
 * Replaces the representative store of RepGenomeDb: new RepGenomeDb(threshold, protName)
 * (RepGenomeDbProcessor.java:122), addRep (BaseGenomeProcessor.java:172-176,204),
 * load (RepGenomeDbProcessor.java:209).  Representatives are numbered 0,1,2... in the order
 * they are appended.  `rank` gives each one's position in the database's iteration order
 * (java.util.HashMap order by genome id, emulated by the host layer); ties in findClosest are
 * broken by it.  Without kr_db_set_ranks the rank is the append order. */
kr_status kr_db_create(kr_ctx *ctx, kr_db **out);
void kr_db_free(kr_db *db);
uint32_t kr_db_size(const kr_db *db);
kr_status kr_db_append(kr_ctx *ctx, kr_db *db, const uint8_t *residues, const uint64_t *offsets, uint32_t nseq);
kr_status kr_db_append_set(kr_ctx *ctx, kr_db *db, const kr_seqset *set, const uint32_t *idx, uint32_t n);
kr_status kr_db_set_ranks(kr_ctx *ctx, kr_db *db, const uint32_t *rank, uint32_t n);

/* RepGenomeDb.findClosest(ProteinKmers|Sequence|RepGenome) for a batch of queries
 * (RepGenomeDbProcessor.java:237, GtoRepGenomeProcessor.java:271, GtoClassProcessor.java:132,
 * ProteinData.java:214).  best_sim[q] = max over representatives of similarity(rep, q);
 * best_first[q] / best_last[q] = the representative with that similarity that comes first / last
 * in iteration order (KR_NO_REP when best_sim is 0: "nothing close", GtoRepGenomeProcessor.java:277).
 * Either index array may be NULL. */
kr_status kr_db_find_closest(kr_ctx *ctx, const kr_db *db, const kr_seqset *queries, int32_t *best_sim,
                             uint32_t *best_first, uint32_t *best_last);
kr_status kr_db_find_closest_device(kr_ctx *ctx, const kr_db *db, const kr_seqset *queries, int32_t *d_best_sim,
                                    uint32_t *d_best_first, uint32_t *d_best_last);
/* RepGenomeDb.checkSimilarity(ProteinKmers, int sim) for a batch (BaseGenomeProcessor.java:203):
 * represented[c] = 1 iff some representative has similarity >= threshold.  (The reference's loop stops at the first
 * such representative; the index scan scores all representatives that share a k-mer with the candidate in one pass
 * over the candidate's ~300 keys, so there is nothing left to cut short.) */
kr_status kr_db_any_ge(kr_ctx *ctx, const kr_db *db, const kr_seqset *cands, int32_t threshold, uint8_t *represented);
/* The greedy build loop: for each candidate in order, RepGenomeDb.checkGenome(candidate)
 * (RepGenomeDbProcessor.java:160-163; GtoRepBuildProcessor.java:130-137;
 * BaseGenomeProcessor.java:203-204): it becomes a representative iff no current representative
 * (including candidates accepted earlier in this call) has similarity >= threshold.
 * added[c] = 1 for the candidates appended to db, in candidate order.  `batch` candidates are
 * scored per launch (0 = library default); the result does not depend on it. */
kr_status kr_greedy_build(kr_ctx *ctx, kr_db *db, const kr_seqset *cands, int32_t threshold, uint32_t batch,
                          uint8_t *added);

/* ---- instrumentation -----------------------------------------------------------------------
 * Replaces the rate logging of the reference (RepGenomeDbProcessor.java:172-179,253-261,
 * DistanceMatrixProcessor.java:84-93). */
typedef struct kr_stats {
    uint64_t kernel_launches; /* kernels of this library launched since the last reset */
    uint64_t pairs;           /* sequence pairs whose similarity was produced */
    double encode_ms;         /* encode + per-sequence sort/dedupe kernels (device time) */
    double blocklist_ms;      /* block key-list build kernels */
    double join_ms;           /* tile join kernels (matrix / argmax / sparse) */
    double other_ms;          /* pair, merge, zero-fill and bookkeeping kernels */
    uint64_t join_launches;
    uint64_t h2d_bytes, d2h_bytes;
    double index_ms;          /* sorted-index build kernels (partition, sort, directory) */
    double scan_ms;           /* index scan kernels (argmax / cells / dense cells) */
    uint64_t scan_launches;   /* index scans launched: tells which path served a call */
    uint64_t index_builds;
    /* posting-list statistics of the index most recently consulted for a path choice: entries (sum of list lengths m),
     * sum of m^2 (postings a self scan walks = sum over ordered pairs of their similarity, diagonal included) */
    uint64_t index_entries, index_walk;
    /* zero-fill of dense results that ran on a second stream beside an index build (device time of the fill kernels; it
     * overlaps index_ms, so the stage times no longer add up to the step time) */
    double fill_ms;
    uint64_t fill_launches;
} kr_stats;
kr_status kr_stats_get(kr_ctx *ctx, kr_stats *out); /* synchronises the context's stream */
kr_status kr_stats_reset(kr_ctx *ctx);
/* device-time bracket on the context's own stream (CUDA events): begin records, end records,
 * waits for the stream and returns the elapsed milliseconds in between */
kr_status kr_timer_begin(kr_ctx *ctx);
kr_status kr_timer_end(kr_ctx *ctx, double *ms);

/* pinned host memory for callers that want full-speed copies of large matrices */
kr_status kr_host_alloc(kr_ctx *ctx, size_t bytes, void **out);
void kr_host_free(kr_ctx *ctx, void *p);

/* ---- several GPUs behind one handle ---------------------------------------------------------
 * "kr_create(const int* devices, int ndev, ...)", one worker thread + stream per
 * GPU inside, so that one JVM reaches every GPU of the box.  A kr_multi owns one kr_ctx per listed
 * device (a device may be listed more than once).  Every call below takes HOST buffers, splits the
 * work (row stripes of a matrix, query ranges of a findClosest, candidate ranges of a build batch),
 * runs the single-GPU entry point on each context from that context's worker thread and lets each
 * GPU copy its share of the result into its own slice of the caller's buffer -- the "allgather of
 * per-shard results" of a caller whose consumer is host code.  Results are identical to the
 * single-GPU calls.  A kr_multi is internally serialised (one multi call at a time, any thread). */
typedef struct kr_multi kr_multi;
typedef struct kr_multi_db kr_multi_db; /* a RepGenomeDb replicated on every GPU of a kr_multi */
kr_status kr_multi_create(const int *devices, int ndev, int K, uint32_t flags, kr_multi **out);
void kr_multi_destroy(kr_multi *m);
int kr_multi_ndev(const kr_multi *m);
kr_ctx *kr_multi_ctx(kr_multi *m, int i);          /* the i-th context, for single-GPU calls */
const char *kr_multi_last_error(const kr_multi *m); /* m may be NULL for creation errors */
/* kr_encode + kr_sim_matrix over all GPUs.  off_b == NULL: all-vs-all of batch A in `layout`
 * (DistanceMatrixProcessor.java:71-76, RepMatrixProcessor.java:88-93); otherwise rows of A against B,
 * KR_LAYOUT_FULL (NeighborProcessor.java:165-193).  out_host holds kr_matrix_elems(n_a, n_cols, layout, 0, n_a)
 * elements of `elem` (KR_ELEM_*). */
kr_status kr_multi_sim_matrix(kr_multi *m, const uint8_t *res_a, const uint64_t *off_a, uint32_t n_a, const uint8_t *res_b,
                              const uint64_t *off_b, uint32_t n_b, uint32_t elem, uint32_t layout, void *out_host);
/* the same as sparse cells (kr_sim_matrix_sparse): (row, col, sim, 0) with sim >= min_sim, unordered;
 * layout KR_LAYOUT_FULL, or KR_LAYOUT_TRIANGLE for an all-vs-all */
kr_status kr_multi_sim_matrix_sparse(kr_multi *m, const uint8_t *res_a, const uint64_t *off_a, uint32_t n_a, const uint8_t *res_b,
                                     const uint64_t *off_b, uint32_t n_b, int32_t min_sim, uint32_t layout, uint32_t *cells_host,
                                     uint64_t cap, uint64_t *count);
kr_status kr_multi_db_create(kr_multi *m, kr_multi_db **out);
void kr_multi_db_free(kr_multi_db *db);
uint32_t kr_multi_db_size(const kr_multi_db *db);
kr_status kr_multi_db_append(kr_multi *m, kr_multi_db *db, const uint8_t *residues, const uint64_t *offsets, uint32_t nseq);
kr_status kr_multi_db_set_ranks(kr_multi *m, kr_multi_db *db, const uint32_t *rank, uint32_t n);
/* kr_encode + kr_db_find_closest, queries split into one contiguous range per GPU
 * (RepGenomeDbProcessor.java:235-252, GtoRepGenomeProcessor.java:271) */
kr_status kr_multi_db_find_closest(kr_multi *m, const kr_multi_db *db, const uint8_t *residues, const uint64_t *offsets, uint32_t nq,
                                   int32_t *best_sim, uint32_t *best_first, uint32_t *best_last);
kr_status kr_multi_db_any_ge(kr_multi *m, const kr_multi_db *db, const uint8_t *residues, const uint64_t *offsets, uint32_t nc,
                             int32_t threshold, uint8_t *represented);
/* kr_encode + kr_greedy_build: per batch every GPU scores its share of the candidates against its replica, the
 * survivors are resolved once (first GPU), the accepted ones are appended to every replica
 * (RepGenomeDbProcessor.java:134-147, 160-163) */
kr_status kr_multi_greedy_build(kr_multi *m, kr_multi_db *db, const uint8_t *residues, const uint64_t *offsets, uint32_t ncand,
                                int32_t threshold, uint32_t batch, uint8_t *added);
kr_status kr_multi_stats_get(kr_multi *m, kr_stats *out); /* sums over the GPUs (times are device-ms summed) */

#ifdef __cplusplus
}
#endif
#endif /* KMERSREPS_H */
