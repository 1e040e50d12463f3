// kr_index.cuh -- sm_100a kernels of the SORTED INVERTED INDEX and of the warp-per-query scan that serves
// findClosest, checkSimilarity / the greedy build and the sparse all-vs-all.
//
// Index of a collection of sequences (a RepGenomeDb, or the set an all-vs-all runs over):
//   entry   (H(key) << idbits) | id          one uint64 per (distinct k-mer, sequence); H is a bijection of the
//                                            5K-bit key space, so equal hashed keys are equal keys (exact, no
//                                            false sharing) and hashed keys are uniform whatever the residue
//                                            composition is
//   ent[]   all entries, sorted by hashed key: the posting list of a k-mer is a contiguous run
//   dir[]   2^D + 1 offsets: entries whose hashed key starts with the D-bit prefix b are ent[dir[b] .. dir[b+1])
//           (D is chosen so that a prefix holds 2..4 entries: one or two 32-byte sectors per probe)
// Build (K4): count entries per B-bit prefix -> exclusive scan -> scatter into the prefix's slot (one streaming
//   pass each, global atomics only on the 2^B counters) -> one CTA per prefix sorts its ~2,000 entries in shared
//   memory -> directory.  About 50 B of HBM traffic per entry; nothing is hashed into a table, nothing is probed.
// Query (K5): one WARP per query.  Each lane takes keys of the query, reads the two directory words and walks the
//   few entries of the prefix; every matching entry adds 1 to the score of its sequence id in the warp's own
//   shared-memory score table (open addressing on the id, 512 slots, touched slots remembered in a list).  A
//   query that touches more ids than fit is redone in 4, 16, 64 ... passes over disjoint id classes, so the
//   result is exact for any input.  The epilogue turns the touched slots into
//     ARGMAX  (max similarity, first rank, last rank)          -- RepGenomeDb.findClosest / checkSimilarity
//     CELLS   (row, col, sim) for every sim >= min_sim         -- similarity matrices as sparse cells
//   with the equal-sequence rule applied (score == both sizes -> hash -> byte compare -> INFINITY).
// Every similarity is an exact integer count.
#pragma once
#include "kr_kernels.cuh"

namespace kr {

// ---- bijective hash of a `bits`-wide key (xorshift and odd multiplications are invertible modulo 2^bits) ----------
__host__ __device__ __forceinline__ u64 ix_hash(u64 k, int bits, u64 mask) {
    const int s = bits >> 1;
    k ^= k >> s;
    k = (k * 0x9E3779B97F4A7C15ULL) & mask;
    k ^= k >> s;
    k = (k * 0xD6E8FEB86659FD93ULL) & mask;
    k ^= k >> s;
    return k;
}

struct IxGeom {      // bit layout of an index: plain values, identical on host and device
    int keybits;     // 5K
    int idbits;      // bits of a sequence id (>= 1)
    int B;           // prefix bits of the build partition
    int D;           // prefix bits of the directory
    u64 keymask;
};

struct KIndex {
    const u64 *ent;  // entries, sorted by hashed key
    const u32 *dir;  // 2^D + 1 offsets
    IxGeom g;
};

constexpr u32 IX_SORT_CAP = 4096;      // entries a prefix may hold to be sorted in shared memory (32 KB)
constexpr u32 IX_SORT_THREADS = 256;

// one segment of the collection being indexed: sequences [0, nseq) of a set become ids id_base + s
struct IxSegment { const u64 *keys; const u64 *koff; const SeqMeta *meta; u32 nseq; u32 id_base; };

// K4.1 / K4.3: count, then scatter, the entries of a segment by B-bit prefix.  One warp per sequence (its key list is
// contiguous); part / nparts select the hashed keys of one partition (multi-GPU build: partition p owns the prefixes
// [p * 2^B / nparts, (p + 1) * 2^B / nparts)); nparts == 1 takes everything.
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_ix_partition(IxSegment sg, IxGeom g, u32 pfx_lo, u32 pfx_hi, u32 *cnt_or_cursor, u64 *ent) {
    const u32 lane = threadIdx.x & 31;
    const u32 s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= sg.nseq) return;
    const u32 n = sg.meta[s].size;
    const u64 *kp = sg.keys + sg.koff[s];
    const u64 id = sg.id_base + s;
    const int pshift = g.keybits - g.B;
    for (u32 i = lane; i < n; i += 32) {
        const u64 hk = ix_hash(kp[i], g.keybits, g.keymask);
        const u32 pfx = (u32)(hk >> pshift);
        if (pfx < pfx_lo || pfx >= pfx_hi) continue;
        if (SCATTER) { const u32 pos = atomicAdd(&cnt_or_cursor[pfx], 1u); ent[pos] = (hk << g.idbits) | id; }
        else atomicAdd(&cnt_or_cursor[pfx], 1u);
    }
}

// K4.2: exclusive scan of the 2^B prefix counts (one CTA): start[p] = first slot of prefix p, start[2^B] = total;
// cursor[p] = start[p] for the scatter.  `base` shifts every start (entries of earlier partitions precede).
__global__ void __launch_bounds__(1024) k_ix_scan(const u32 *cnt, u32 nbins, u32 *start, u32 *cursor) {
    __shared__ u32 part[1024];
    const u32 tid = threadIdx.x, per = (nbins + 1023) / 1024;
    const u32 b0 = min(nbins, tid * per), b1 = min(nbins, b0 + per);
    u32 acc = 0;
    for (u32 i = b0; i < b1; ++i) acc += cnt[i];
    part[tid] = acc;
    __syncthreads();
    if (tid == 0) { u32 run = 0; for (int i = 0; i < 1024; ++i) { const u32 v = part[i]; part[i] = run; run += v; } start[nbins] = run; }
    __syncthreads();
    u32 run = part[tid];
    for (u32 i = b0; i < b1; ++i) { const u32 v = cnt[i]; start[i] = run; cursor[i] = run; run += v; }
}

// K4.4: one CTA per prefix sorts the prefix's entries.  The entries of a prefix share their top B bits and the bits
// below are a hash, i.e. uniform: a counting sort on the next 11 bits leaves about one entry per bucket, so the sort is
// three passes over shared memory (histogram with the rank kept in a register, scan, scatter) plus an insertion sort
// of the few buckets that hold more than one entry -- an order of magnitude fewer instructions than a sorting network
// over 2,000 entries.  Inputs that are not uniform (one k-mer shared by thousands of sequences puts thousands of equal
// hashed keys into one bucket) take the network instead; a prefix beyond IX_SORT_CAP entries is sorted in place in
// global memory.  `low_bits` = number of entry bits below the prefix (key bits below the prefix + id bits).
constexpr u32 IX_SORT_BUCKET_BITS = 11, IX_SORT_NB = 1u << IX_SORT_BUCKET_BITS;
constexpr u32 IX_SORT_PER_THREAD = IX_SORT_CAP / IX_SORT_THREADS;      // entries a thread keeps in registers
constexpr u32 IX_SORT_INSERTION_MAX = 24;                              // a bucket above this sends the prefix to the network
static_assert(IX_SORT_NB == 8 * IX_SORT_THREADS, "the scan gives every thread 8 consecutive buckets");
// bucket b lives in word b + (b >> 5): thread t owns buckets 8 t .. 8 t + 7, and the skew spreads the 32 threads of a warp over
// all 32 shared-memory banks in every step of its loops (unskewed, they would share four)
__device__ __forceinline__ u32 ix_sort_slot(u32 b) { return b + (b >> 5); }
__global__ void __launch_bounds__(IX_SORT_THREADS) k_ix_sort(const u32 *start, u64 *ent, int low_bits) {
    __shared__ __align__(16) u64 sk[IX_SORT_CAP];
    __shared__ u32 cnt[IX_SORT_NB + IX_SORT_NB / 32 + 2];
    __shared__ u32 wsum[IX_SORT_THREADS / 32];
    __shared__ u32 s_big;
    const u32 p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 lo = start[p], n = start[p + 1] - lo;
    if (n < 2) return;
    u64 *g = ent + lo;
    if (n > IX_SORT_CAP) { bitonic_sort(g, n, tid, IX_SORT_THREADS); return; }
    const int bshift = low_bits > (int)IX_SORT_BUCKET_BITS ? low_bits - (int)IX_SORT_BUCKET_BITS : 0;
    u64 e[IX_SORT_PER_THREAD];
    u32 rank[IX_SORT_PER_THREAD];
    for (u32 i = tid; i < IX_SORT_NB + IX_SORT_NB / 32 + 2; i += IX_SORT_THREADS) cnt[i] = 0;
    if (tid == 0) s_big = 0;
    __syncthreads();
    // 1. histogram; the value the atomic returns is the entry's place inside its bucket
#pragma unroll
    for (u32 j = 0; j < IX_SORT_PER_THREAD; ++j) {
        const u32 i = tid + j * IX_SORT_THREADS;
        if (i < n) {
            e[j] = g[i];
            rank[j] = atomicAdd(&cnt[ix_sort_slot((u32)(e[j] >> bshift) & (IX_SORT_NB - 1))], 1u);
        }
    }
    __syncthreads();
    // 2. exclusive scan of the bucket counts, in place: thread t owns buckets 8t .. 8t+7
    u32 c[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k] = cnt[ix_sort_slot(tid * 8 + k)]; sum += c[k]; if (c[k] > IX_SORT_INSERTION_MAX) atomicExch(&s_big, 1u); }   // (several threads may: an atomic, not a plain store)
    u32 incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += up; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    u32 base = incl - sum;
    for (u32 w = 0; w < warp; ++w) base += wsum[w];
#pragma unroll
    for (int k = 0; k < 8; ++k) { cnt[ix_sort_slot(tid * 8 + k)] = base; base += c[k]; }
    __syncthreads();
    // 3. scatter into bucket order
#pragma unroll
    for (u32 j = 0; j < IX_SORT_PER_THREAD; ++j) {
        const u32 i = tid + j * IX_SORT_THREADS;
        if (i < n) sk[cnt[ix_sort_slot((u32)(e[j] >> bshift) & (IX_SORT_NB - 1))] + rank[j]] = e[j];
    }
    __syncthreads();
    if (s_big) {
        bitonic_sort(sk, n, tid, IX_SORT_THREADS);          // (uniform decision: s_big was read after the barrier)
    } else {
        // 4. buckets with more than one entry: insertion sort by the thread that owns the bucket
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (c[k] > 1) {
                u64 *q = sk + cnt[ix_sort_slot(tid * 8 + k)];
                for (u32 x = 1; x < c[k]; ++x) {
                    const u64 v = q[x];
                    u32 y = x;
                    while (y > 0 && q[y - 1] > v) { q[y] = q[y - 1]; --y; }
                    q[y] = v;
                }
            }
        }
        __syncthreads();
    }
    for (u32 i = tid; i < n; i += IX_SORT_THREADS) g[i] = sk[i];
}

// K4.5: the directory and the posting-list statistics, one thread per entry (one pass over the sorted entries).
// Directory: entry i opens every D-bit prefix after its predecessor's up to its own; the last entry also closes the
// prefixes above it.  An empty index is all zeros (memset by the host).
// Statistics (run starts only): out[0] += sum over the posting lists of m^2 (the postings a self scan walks: every entry
// walks its own list), out[1] = longest list, out[2] = number of entries, out[3] += number of distinct keys.
__global__ void __launch_bounds__(256) k_ix_dir(const u64 *ent, const u32 *total, int dshift, u32 ndir, u32 *dir, int idbits, unsigned long long *out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n = *total;      // the number of entries is only known on the device (sum of the distinct-k-mer counts)
    if (i == 0) out[2] = n;
    unsigned long long w = 0, mx = 0, runs = 0;
    if (i < n) {
        const u64 e = ent[i], prev = i ? ent[i - 1] : 0;
        const u32 d = (u32)(e >> dshift);
        const u32 first = i ? (u32)(prev >> dshift) + 1 : 0;
        for (u32 b = first; b <= d; ++b) dir[b] = i;
        if (i == n - 1) for (u32 b = d + 1; b <= ndir; ++b) dir[b] = n;
        const u64 hk = e >> idbits;
        if (i == 0 || (prev >> idbits) != hk) {
            u32 m = 1;
            while (i + m < n && (ent[i + m] >> idbits) == hk) ++m;
            w = (unsigned long long)m * m; mx = m; runs = 1;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        w += __shfl_xor_sync(0xffffffffu, w, o);
        runs += __shfl_xor_sync(0xffffffffu, runs, o);
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = t > mx ? t : mx;
    }
    if ((threadIdx.x & 31) == 0 && w) { atomicAdd(&out[0], w); atomicMax(&out[1], mx); atomicAdd(&out[3], runs); }
}

// ---------------------------------------------------------------------------------------------------------------
// K5: the scan.
enum { IXQ_ARGMAX = 0, IXQ_CELLS = 1, IXQ_DENSE = 2 };
constexpr u32 IXQ_WARPS = 8;            // queries in flight per CTA
#ifndef KR_IXQ_H
#define KR_IXQ_H 512
#endif
constexpr u32 IXQ_H = KR_IXQ_H;         // slots of a warp's score table (a power of two; the tests also build a 64-slot variant)
constexpr u32 IXQ_FILL = IXQ_H >= 256 ? IXQ_H * 3 / 4 - 32 : IXQ_H / 4;   // ids a pass may touch before the query is redone in more passes
constexpr int ixq_log2(u32 v) { return v <= 1 ? 0 : 1 + ixq_log2(v >> 1); }
constexpr int IXQ_SHIFT = 32 - ixq_log2(IXQ_H);
static_assert((IXQ_H & (IXQ_H - 1)) == 0 && IXQ_H >= 64, "the score table is indexed by a multiplicative hash");
static_assert(IXQ_FILL + 33 <= IXQ_H, "after the fill stops, each lane may still add one id");
constexpr u32 IXQ_EMPTY = 0xFFFFFFFFu;
// per warp: ids (u32), scores (u16, two per word), touched-slot list (u16), two counters
constexpr u32 IXQ_SMEM_PER_WARP = IXQ_H * 4 + IXQ_H * 2 + IXQ_H * 2 + 16;
constexpr u32 IXQ_SMEM = IXQ_WARPS * IXQ_SMEM_PER_WARP;     // 32,896 B at 512 slots: six CTAs (48 queries) per SM
constexpr u32 IXQ_CTAS_PER_SM = (232448u / (IXQ_SMEM + 1024u)) < 8u ? (232448u / (IXQ_SMEM + 1024u)) : 8u;

struct IxQueryParams {
    // queries: sequences [q_begin, q_end) of a set
    const u64 *qkeys; const u64 *qkoff; const SeqMeta *qmeta; u32 q_begin, q_end;
    KIndex ix;
    const SeqMeta *rmeta;          // records of the indexed sequences, by id
    u32 flags;                     // KR_IDENT_IS_INFINITY
    u32 *next;                     // work counter (starts at 0)
    // ARGMAX
    const u32 *rank; int4 *partial;   // partial[q - q_begin] = (sim, first rank, last rank)
    // CELLS / DENSE: the cells (row = q, col = id) with sim >= min_sim and col >= col_min(q)
    u32 col_mode;                  // 0: every column, 1: columns > q (strict triangle), 2: columns >= 256 * (q / 256)
    u32 min_sim;
    uint4 *cells; unsigned long long cells_cap; unsigned long long *cells_count;      // CELLS: (row, col, sim, 0), any order
    u32 cells_packed;              // CELLS: 8-byte cells (put_cell)
    void *out; u32 elem, layout; u64 origin; u32 nrows, ncols;                        // DENSE: the result
    u32 prezeroed;                 // DENSE: the rows were already zero-filled (k_zero_fill overlapped with the index build)
};

template <int MODE>
__global__ void __launch_bounds__(IXQ_WARPS * 32) k_ix_query(IxQueryParams p) {
    KR_DYN_SMEM(u32, smem_words, 16);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 *tk = smem_words + (size_t)warp * (IXQ_SMEM_PER_WARP / 4);   // ids
    u32 *ts = tk + IXQ_H;                                            // scores: slot h is the (h & 1) half of word h >> 1
    u16 *tl = reinterpret_cast<u16 *>(ts + IXQ_H / 2);               // touched slots, in insertion order
    volatile u32 *tc = reinterpret_cast<volatile u32 *>(tl + IXQ_H); // [0] = touched count, [1] = "table nearly full"
    for (u32 i = lane; i < IXQ_H; i += 32) tk[i] = IXQ_EMPTY;
    for (u32 i = lane; i < IXQ_H / 2; i += 32) ts[i] = 0;
    if (lane == 0) { tc[0] = 0; tc[1] = 0; }
    __syncwarp();
    const bool ident_inf = (p.flags & 1u) != 0;
    const IxGeom g = p.ix.g;
    const int dshift = g.keybits - g.D;
    const u64 idmask = (1ull << g.idbits) - 1;
    const u32 nq = p.q_end - p.q_begin;
    for (;;) {
        u32 qi = 0;
        if (lane == 0) qi = atomicAdd(p.next, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= nq) break;
        const u32 q = p.q_begin + qi;
        const SeqMeta qm = p.qmeta[q];
        const u32 n = qm.size;
        const u64 *kp = p.qkeys + p.qkoff[q];
        unsigned long long bf = 0, bl = 0;   // ARGMAX: (score << 32) | ~rank  and  (score << 32) | rank
        u32 col_min = 0;
        if (MODE != IXQ_ARGMAX) col_min = p.col_mode == 1 ? q + 1 : p.col_mode == 2 ? (q & ~255u) : 0;
        if (MODE == IXQ_DENSE && !p.prezeroed && p.ncols > col_min) {
            // The warp owns row q of the result: it writes the whole row as zeros (streaming 16-byte stores; the head and
            // tail up to the next 16-byte boundary element-wise) and afterwards the few cells that are not zero.  No
            // separate zero-fill pass over the N^2 result, and the stores overlap the latency-bound walk of other warps.
            const u32 esz = p.elem == 1 ? 2u : 4u;
            u8 *bp = reinterpret_cast<u8 *>(p.out) + (cell_index(p.layout, q, col_min, p.nrows, p.ncols) - p.origin) * esz;
            const u64 bytes = (u64)(p.ncols - col_min) * esz;
            u64 head = (16 - (reinterpret_cast<uintptr_t>(bp) & 15)) & 15;
            if (head > bytes) head = bytes;
            const u64 body = (bytes - head) & ~15ull;
            for (u64 k = lane * 2; k < head; k += 64) *reinterpret_cast<u16 *>(bp + k) = 0;
            uint4 *vp = reinterpret_cast<uint4 *>(bp + head);
            for (u64 k = lane; k < body / 16; k += 32) __stcs(vp + k, make_uint4(0u, 0u, 0u, 0u));   // streaming: never re-read here, L2 is for the index
            for (u64 k = head + body + lane * 2; k < bytes; k += 64) *reinterpret_cast<u16 *>(bp + k) = 0;
            __syncwarp();      // orders the zeros before the cells this warp writes below
        }
        // A pass scores the ids of one class (id % npass == pass).  One pass is the rule; when the ids a query touches
        // do not fit the table, NOTHING of the overflowing attempt is kept (no cell emitted, no maximum taken) and the
        // query is redone with four times as many classes, until every class fits.
        for (u32 npass = 1;; npass *= 4) {
            bool redo = false;
            bf = 0; bl = 0;
            // emitting modes need every pass of this attempt to fit before the first cell leaves: try the fills first
            const u32 rounds = (MODE != IXQ_ARGMAX && npass > 1) ? 2 : 1;
            for (u32 round = 0; round < rounds && !redo; ++round) {
                const bool dry = (rounds == 2 && round == 0);       // fill and clear only: does every class fit?
                for (u32 pass = 0; pass < npass && !redo; ++pass) {
                    // ---- walk: every entry of the index that carries one of the query's keys scores its id -----------
                    // (two keys per lane in flight: both directory reads, then both first entries, before either list is walked)
                    auto score = [&](u32 id) {
                        if (npass > 1 && (id % npass) != pass) return;
                        if (MODE != IXQ_ARGMAX && id < col_min) return;             // a cell that is never emitted needs no score
                        u32 h = (id * 0x9E3779B1u) >> IXQ_SHIFT;                    // multiplicative hash, top log2(IXQ_H) bits
                        for (;;) {
                            const u32 cur = *reinterpret_cast<volatile u32 *>(&tk[h]);
                            if (cur == id) break;
                            if (cur == IXQ_EMPTY) {
                                const u32 old = atomicCAS(&tk[h], IXQ_EMPTY, id);
                                if (old == IXQ_EMPTY) {
                                    const u32 pos = atomicAdd(const_cast<u32 *>(&tc[0]), 1u);
                                    if (pos < IXQ_H) tl[pos] = (u16)h;
                                    if (pos >= IXQ_FILL) tc[1] = 1;
                                    break;
                                }
                                if (old == id) break;
                            }
                            h = (h + 1) & (IXQ_H - 1);
                        }
                        atomicAdd(&ts[h >> 1], (h & 1) ? 0x10000u : 1u);
                    };
                    auto walk = [&](u64 hk, u32 lo, u32 hi, u64 e) {                 // e = ent[lo], already loaded when lo < hi
                        for (u32 t = lo; t < hi; ++t) {
                            if (t != lo) e = __ldg(p.ix.ent + t);
                            const u64 ek = e >> g.idbits;
                            if (ek < hk) continue;
                            if (ek > hk) break;
                            score((u32)(e & idmask));
                            if (tc[1]) break;      // table nearly full: stop filling (at most one more id per lane gets in)
                        }
                    };
                    for (u32 i = lane; i < n; i += 64) {
                        const bool two = i + 32 < n;
                        const u64 hk0 = ix_hash(kp[i], g.keybits, g.keymask);
                        const u64 hk1 = two ? ix_hash(kp[i + 32], g.keybits, g.keymask) : 0;
                        const u32 b0 = (u32)(hk0 >> dshift), b1 = (u32)(hk1 >> dshift);
                        const u32 lo0 = __ldg(p.ix.dir + b0), hi0 = __ldg(p.ix.dir + b0 + 1);
                        const u32 lo1 = two ? __ldg(p.ix.dir + b1) : 0, hi1 = two ? __ldg(p.ix.dir + b1 + 1) : 0;
                        const u64 e0 = lo0 < hi0 ? __ldg(p.ix.ent + lo0) : 0, e1 = lo1 < hi1 ? __ldg(p.ix.ent + lo1) : 0;
                        walk(hk0, lo0, hi0, e0);
                        if (tc[1]) break;
                        walk(hk1, lo1, hi1, e1);
                        if (tc[1]) break;
                    }
                    __syncwarp();
                    const u32 T = min(tc[0], IXQ_H);
                    redo = tc[1] != 0;
                    // ---- epilogue over the touched slots (also clears them) ----------------------------------------
                    for (u32 t0 = 0; t0 < T; t0 += 32) {
                        const u32 t = t0 + lane;
                        u32 id = 0, sc = 0;
                        if (t < T) {
                            const u32 h = tl[t];
                            id = tk[h]; sc = (ts[h >> 1] >> ((h & 1) * 16)) & 0xFFFFu;
                            tk[h] = IXQ_EMPTY;
                        }
                        if (redo || dry) continue;
                        // the equal-sequence rule: only a sequence holding every key of the query can be identical
                        if (ident_inf && t < T && sc == n) { const SeqMeta rm = p.rmeta[id]; if (rm.size == n && seq_equal(qm, rm)) sc = KR_INF; }
                        if (MODE == IXQ_ARGMAX) {
                            if (t < T && sc >= (u32)(bf >> 32)) {
                                const u32 rk = p.rank[id];
                                const unsigned long long cf = ((unsigned long long)sc << 32) | (0xFFFFFFFFu - rk);
                                const unsigned long long cl = ((unsigned long long)sc << 32) | rk;
                                bf = cf > bf ? cf : bf; bl = cl > bl ? cl : bl;
                            }
                        } else if (MODE == IXQ_CELLS) {
                            const bool emit = t < T && sc >= p.min_sim && id >= col_min;
                            const u32 bal = __ballot_sync(0xffffffffu, emit);
                            if (bal) {
                                unsigned long long base = 0;
                                if (lane == 0) base = atomicAdd(p.cells_count, (unsigned long long)__popc(bal));
                                base = ((unsigned long long)__shfl_sync(0xffffffffu, (u32)(base >> 32), 0) << 32) | __shfl_sync(0xffffffffu, (u32)base, 0);
                                const unsigned long long pos = base + __popc(bal & ((1u << lane) - 1));
                                if (emit && pos < p.cells_cap) put_cell(p.cells, p.cells_packed, pos, q, id, sc);
                            }
                        } else {
                            if (t < T && sc >= p.min_sim && id >= col_min && id < p.ncols) {
                                const u64 idx = cell_index(p.layout, q, id, p.nrows, p.ncols) - p.origin;
                                if (p.elem == 1) reinterpret_cast<u16 *>(p.out)[idx] = (u16)sc; else reinterpret_cast<int *>(p.out)[idx] = (int)sc;
                            }
                        }
                    }
                    __syncwarp();
                    // two slots share a score word: clear whole words (a word next to a touched slot holds no other live score)
                    // (two lanes may clear the same word: volatile stores, i.e. plain st.volatile.shared, not a data race on one word)
                    for (u32 t = lane; t < T; t += 32) *reinterpret_cast<volatile u32 *>(&ts[tl[t] >> 1]) = 0;
                    __syncwarp();
                    if (lane == 0) { tc[0] = 0; tc[1] = 0; }
                    __syncwarp();
                }
            }
            if (!redo) break;
            if (npass >= (1u << 24)) break;   // cannot happen: a class then holds at most one id
        }
        if (MODE == IXQ_ARGMAX) {
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long t1 = ((unsigned long long)__shfl_xor_sync(0xffffffffu, (u32)(bf >> 32), o) << 32) | __shfl_xor_sync(0xffffffffu, (u32)bf, o);
                const unsigned long long t2 = ((unsigned long long)__shfl_xor_sync(0xffffffffu, (u32)(bl >> 32), o) << 32) | __shfl_xor_sync(0xffffffffu, (u32)bl, o);
                bf = t1 > bf ? t1 : bf; bl = t2 > bl ? t2 : bl;
            }
            if (lane == 0) {
                const int sim = (int)(bf >> 32);
                p.partial[qi] = make_int4(sim, sim ? (int)(0xFFFFFFFFu - (u32)bf) : -1, sim ? (int)(u32)bl : 0, 0);
            }
        }
    }
}

// Zero-fill of a dense result (16-byte streaming stores, grid-stride).  The host launches it on a second stream when a
// dense matrix call has to build its index first: the fill is pure HBM write traffic, the index build is atomics and
// shared-memory sorting, so the two overlap almost perfectly; the scan then only writes the non-zero cells.
__global__ void __launch_bounds__(256) k_zero_fill(u8 *p, u64 bytes) {
    u64 head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > bytes) head = bytes;
    const u64 body = (bytes - head) >> 4;
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (u64)gridDim.x * blockDim.x;
    uint4 *v = reinterpret_cast<uint4 *>(p + head);
    for (u64 k = tid; k < body; k += nthr) __stcs(v + k, make_uint4(0u, 0u, 0u, 0u));
    if (tid < head) p[tid] = 0;
    const u64 tail0 = head + (body << 4);
    if (tail0 + tid < bytes && tid < 16) p[tail0 + tid] = 0;
}

// Sequences without any k-mer (shorter than K) never meet the index, but two EQUAL ones are still "identical": under
// KR_IDENT_IS_INFINITY their cell is INFINITY.  rows_e / cols_e list the k-mer-less sequences of the two sides (device
// counts); one thread per (row, col) candidate pair.  Writes a dense cell (out != null) or appends a sparse cell.
struct EmptyPairsParams {
    const u32 *rows_e; const u32 *nrows_e; const u32 *cols_e; const u32 *ncols_e;
    const SeqMeta *metaA, *metaB;
    u32 row_begin, row_end, col_mode, min_sim;
    void *out; u32 elem, layout; u64 origin; u32 nrows, ncols;
    uint4 *cells; unsigned long long cells_cap; unsigned long long *cells_count; u32 cells_packed;
};
__global__ void __launch_bounds__(256) k_ix_empty_pairs(EmptyPairsParams p) {
    const u64 na = *p.nrows_e, nb = *p.ncols_e, total = na * nb;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x) {
        const u32 i = p.rows_e[t / nb], j = p.cols_e[t % nb];
        if (i < p.row_begin || i >= p.row_end) continue;
        const u32 col_min = p.col_mode == 1 ? i + 1 : p.col_mode == 2 ? (i & ~255u) : 0;
        if (j < col_min || (u32)KR_INF < p.min_sim) continue;
        if (!seq_equal(p.metaA[i], p.metaB[j])) continue;
        if (p.out) {
            const u64 idx = cell_index(p.layout, i, j, p.nrows, p.ncols) - p.origin;
            if (p.elem == 1) reinterpret_cast<u16 *>(p.out)[idx] = (u16)KR_INF; else reinterpret_cast<int *>(p.out)[idx] = KR_INF;
        } else {
            const unsigned long long pos = atomicAdd(p.cells_count, 1ull);
            if (pos < p.cells_cap) put_cell(p.cells, p.cells_packed, pos, i, j, (u32)KR_INF);
        }
    }
}

}  // namespace kr
