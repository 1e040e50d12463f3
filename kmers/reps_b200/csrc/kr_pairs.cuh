// kr_pairs.cuh -- EXPERIMENTAL (compiled only with -DKR_EXPERIMENTAL; unverified on hardware when written):
// all-vs-all similarity from posting lists.
//
// The tile join (kr_kernels.cuh) does work proportional to (number of tiles) x (block-list length) whatever the
// data.  When few pairs share k-mers (families of related proteins among unrelated ones, representative sets whose
// members are pairwise below a threshold) almost all of that work finds nothing.  This path turns it around:
//   1. the inverted index kernels (k_idx_insert / k_scan_* / k_idx_fill) group the (key, sequence) entries of the
//      whole set by key: one posting list of sequence numbers per distinct key;
//   2. k_pairs_emit visits every posting list of m >= 2 sequences and adds 1 to the result cell of each of its
//      m(m-1)/2 unordered pairs -- the similarity of two sequences IS the number of lists that hold both;
//   3. equal sequences (found through a hash chain over the per-sequence hashes) and the diagonal are written last.
// Work = sum over keys of m(m-1)/2 = sum over pairs of their similarity, plus zero-filling the result.  The host
// picks this path or the tile join from that sum (a dense, highly redundant collection is cheaper in the join).
//
// Every kernel here is written WITHOUT __syncthreads and warp intrinsics: threads interact only through global
// atomics.  That keeps them executable, thread after thread, by the sequential CPU shim the tests use to check the
// indexing logic (tests/emul/), which matters because this file was written without access to a GPU.
#pragma once
#include "kr_types.cuh"

#if defined(KR_EMUL)
#include "cuda_emul.h"
#define KRP_DEVICE __device__ __forceinline__
#define KRP_GLOBAL __global__
#elif !defined(KR_PAIRS_EMUL)
#include <cuda_runtime.h>
#define KRP_DEVICE __device__ __forceinline__
#define KRP_GLOBAL __global__
#else
#define KRP_DEVICE inline
#define KRP_GLOBAL
#endif

namespace kr {

constexpr u32 PAIRS_NIL = 0xFFFFFFFFu;

struct PairsParams {
    // posting lists: slot s of the index holds postings[tab_off[s] .. tab_off[s] + tab_cnt[s])
    const u32 *tab_cnt, *tab_off, *postings;
    u64 nslots;
    // result: the cells of rows [row_begin, row_end) of the N x N self matrix, cell 0 = first cell of row_begin
    void *out; u32 elem, layout; u64 origin, total;   // total = number of cells in `out`
    u32 N, row_begin, row_end;
    u32 flags;
    // equal-sequence rule
    const SeqMeta *meta; u32 *head, *next; u32 head_mask;
};

// same cell numbering as cell_index() of the tile join
KRP_DEVICE u64 pairs_cell_index(u32 layout, u64 i, u64 j, u64 n) {
    if (layout == 0) return i * n + j;
    if (layout == 1) return i * n - i * (i + 1) / 2 + (j - i - 1);
    const u64 I = i >> 8, c0 = I << 8;
    return 256ull * (I * n - 128ull * I * (I - 1)) + (i - c0) * (n - c0) + (j - c0);
}

// largest b with b(b-1)/2 <= t: pair number t of the enumeration (0,1),(0,2),(1,2),(0,3),... is (t - b(b-1)/2, b)
KRP_DEVICE u32 pairs_tri_row(u64 t) {
    u64 b = (u64)((1.0 + sqrt(1.0 + 8.0 * (double)t)) * 0.5);
    while (b * (b - 1) / 2 > t) --b;
    while ((b + 1) * b / 2 <= t) ++b;
    return (u32)b;
}

KRP_DEVICE void pairs_cell_add(const PairsParams &p, u64 idx, u32 v = 1) {
    if (p.elem == 1) {
        u16 *c = reinterpret_cast<u16 *>(p.out) + idx;
        if ((idx & 1) == 0 && idx + 1 >= p.total) {
            // the last cell of an odd-sized result has no upper half word inside the buffer: 16-bit CAS loop
            unsigned short old = *c, assumed;
            do { assumed = old; old = atomicCAS(reinterpret_cast<unsigned short *>(c), assumed, (unsigned short)(assumed + v)); } while (old != assumed);
        } else {
            atomicAdd(reinterpret_cast<u32 *>(p.out) + (idx >> 1), (idx & 1) ? (v << 16) : v);
        }
    } else {
        atomicAdd(reinterpret_cast<int *>(p.out) + idx, (int)v);
    }
}
KRP_DEVICE void pairs_cell_set(const PairsParams &p, u64 idx, u32 v) {
    if (p.elem == 1) reinterpret_cast<u16 *>(p.out)[idx] = (u16)v;
    else reinterpret_cast<int *>(p.out)[idx] = (int)v;
}

// the cells of the unordered pair {i < j} that lie in rows [row_begin, row_end): (i, j) in every layout, and the
// mirrored (j, i) where the layout stores it (FULL: always; BLOCKTRI: inside a diagonal 256 x 256 block)
template <bool SET>
KRP_DEVICE void pairs_touch(const PairsParams &p, u32 i, u32 j, u32 v) {
    if (i >= p.row_begin && i < p.row_end) {
        const u64 idx = pairs_cell_index(p.layout, i, j, p.N) - p.origin;
        if (SET) pairs_cell_set(p, idx, v); else pairs_cell_add(p, idx, v);
    }
    if (p.layout != 1 && j >= p.row_begin && j < p.row_end && (p.layout == 0 || (i >> 8) == (j >> 8))) {
        const u64 idx = pairs_cell_index(p.layout, j, i, p.N) - p.origin;
        if (SET) pairs_cell_set(p, idx, v); else pairs_cell_add(p, idx, v);
    }
}

// Step 2.  Warp w takes the groups of 8 consecutive slots w, w + W, w + 2W ... (hashing has already spread the
// long lists evenly); its 32 lanes share the pairs of each list.  All lanes of a warp run the same loops.
KRP_GLOBAL void k_pairs_emit(PairsParams p) {
    const u32 lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 g = warp; g * 8 < p.nslots; g += nwarp) {
        for (u32 k = 0; k < 8; ++k) {
            const u64 slot = g * 8 + k;
            if (slot >= p.nslots) break;
            const u32 m = p.tab_cnt[slot];
            if (m < 2) continue;
            const u32 off = p.tab_off[slot];
            const u64 P = (u64)m * (m - 1) / 2;
            for (u64 t = lane; t < P; t += 32) {
                const u32 b = pairs_tri_row(t), a = (u32)(t - (u64)b * (b - 1) / 2);
                u32 i = p.postings[off + a], j = p.postings[off + b];
                if (i > j) { const u32 x = i; i = j; j = x; }
                pairs_touch<false>(p, i, j, 1);
            }
        }
    }
}

// sum over the lists of m(m-1)/2: the work of k_pairs_emit, for the host's choice between this path and the join
KRP_GLOBAL void k_pairs_work(const u32 *tab_cnt, u64 nslots, unsigned long long *work) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (u64)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (u64 s = tid; s < nslots; s += nthr) { const u64 m = tab_cnt[s]; acc += m * (m - (m ? 1 : 0)) / 2; }
    if (acc) atomicAdd(work, acc);
}

KRP_DEVICE u32 pairs_head_slot(u64 hash, u32 mask) { return (u32)(hash >> 17) & mask; }

KRP_DEVICE bool pairs_seq_equal(const SeqMeta &a, const SeqMeta &b) {
    if (a.len != b.len || a.hash != b.hash || a.size != b.size) return false;
    if (a.res == b.res) return true;
    for (u32 i = 0; i < a.len; ++i) if (a.res[i] != b.res[i]) return false;
    return true;
}

// Step 3a: chain the sequences by hash (head[] preset to PAIRS_NIL)
KRP_GLOBAL void k_pairs_dup_insert(PairsParams p) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    p.next[i] = atomicExch(&p.head[pairs_head_slot(p.meta[i].hash, p.head_mask)], i);
}
// Step 3b: the thread of sequence i writes INFINITY for every equal sequence j > i; then the diagonal cell
KRP_GLOBAL void k_pairs_dup_fix(PairsParams p) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    const SeqMeta mi = p.meta[i];
    if (p.flags & 1u) {
        for (u32 j = p.head[pairs_head_slot(mi.hash, p.head_mask)]; j != PAIRS_NIL; j = p.next[j])
            if (j > i && pairs_seq_equal(mi, p.meta[j])) pairs_touch<true>(p, i, j, (u32)KR_INF);
    }
    if (p.layout != 1 && i >= p.row_begin && i < p.row_end)
        pairs_cell_set(p, pairs_cell_index(p.layout, i, i, p.N) - p.origin, (p.flags & 1u) ? (u32)KR_INF : mi.size);
}

// ---- sparse result: the pairs with similarity >= min_sim as (row, col, sim, 0) cells ------------------------------
// Same posting lists, but the increments go to a hash table keyed by the pair (i < j, row i inside the stripe)
// instead of a dense matrix, so nothing proportional to N^2 is ever written: the table holds one slot per pair that
// shares a k-mer, and a last kernel copies the slots that reach min_sim into the cell list.
constexpr unsigned long long PAIRTAB_EMPTY = ~0ull;

struct PairTabParams {
    const u32 *tab_cnt, *tab_off, *postings;
    u64 nslots;
    unsigned long long *pt_key; u32 *pt_cnt; u64 pt_mask;   // pair table: capacity pt_mask + 1 (a power of two)
    u32 N, row_begin, row_end, flags, min_sim;
    const SeqMeta *meta; u32 *head, *next; u32 head_mask;
    uint4 *cells; unsigned long long cells_cap; unsigned long long *cells_count;
    u32 *overflow;   // set when the pair table turns out to be full (the host sized it from the pair increments: a bug if ever set)
};

KRP_DEVICE u64 pairtab_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
// slot of pair (i, j), claiming an empty one if the pair is new
KRP_DEVICE u64 pairtab_slot(const PairTabParams &p, u32 i, u32 j) {
    const unsigned long long k = ((unsigned long long)i << 32) | j;
    u64 slot = pairtab_hash(k) & p.pt_mask;
    for (u64 probes = 0;; ++probes) {
        const unsigned long long old = atomicCAS(&p.pt_key[slot], PAIRTAB_EMPTY, k);
        if (old == PAIRTAB_EMPTY || old == k) return slot;
        slot = (slot + 1) & p.pt_mask;
        if (probes > p.pt_mask) { if (p.overflow) atomicExch(p.overflow, 1u); return slot; }   // full table: reported, never a hang
    }
}

KRP_GLOBAL void k_pairs_emit_tab(PairTabParams p) {
    const u32 lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 g = warp; g * 8 < p.nslots; g += nwarp) {
        for (u32 k = 0; k < 8; ++k) {
            const u64 slot = g * 8 + k;
            if (slot >= p.nslots) break;
            const u32 m = p.tab_cnt[slot];
            if (m < 2) continue;
            const u32 off = p.tab_off[slot];
            const u64 P = (u64)m * (m - 1) / 2;
            for (u64 t = lane; t < P; t += 32) {
                const u32 b = pairs_tri_row(t), a = (u32)(t - (u64)b * (b - 1) / 2);
                u32 i = p.postings[off + a], j = p.postings[off + b];
                if (i > j) { const u32 x = i; i = j; j = x; }
                if (i >= p.row_begin && i < p.row_end) atomicAdd(&p.pt_cnt[pairtab_slot(p, i, j)], 1u);
            }
        }
    }
}
// equal sequences (after every increment has landed: a separate launch): their cell is INFINITY
KRP_GLOBAL void k_pairs_dup_tab(PairTabParams p) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N || i < p.row_begin || i >= p.row_end || !(p.flags & 1u)) return;
    const SeqMeta mi = p.meta[i];
    for (u32 j = p.head[pairs_head_slot(mi.hash, p.head_mask)]; j != PAIRS_NIL; j = p.next[j])
        if (j > i && pairs_seq_equal(mi, p.meta[j])) p.pt_cnt[pairtab_slot(p, i, j)] = (u32)KR_INF;
}
KRP_GLOBAL void k_pairs_dup_insert_tab(PairTabParams p) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    p.next[i] = atomicExch(&p.head[pairs_head_slot(p.meta[i].hash, p.head_mask)], i);
}
KRP_GLOBAL void k_pairs_compact(PairTabParams p) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (u64)gridDim.x * blockDim.x;
    for (u64 s = tid; s <= p.pt_mask; s += nthr) {
        const unsigned long long k = p.pt_key[s];
        if (k == PAIRTAB_EMPTY) continue;
        const u32 c = p.pt_cnt[s];
        if (c < p.min_sim) continue;
        const unsigned long long pos = atomicAdd(p.cells_count, 1ull);
        if (pos < p.cells_cap) p.cells[pos] = make_uint4((u32)(k >> 32), (u32)k, c, 0u);
    }
}

}  // namespace kr

// =====================================================================================================================
// Bucketed grouping (second generation of the posting-list path).  Building one big hash table in HBM costs about four
// random DRAM accesses per (key, sequence) entry.  Here the entries are first PARTITIONED by a hash of the key into
// buckets of ~1,000 entries with two streaming passes (count, scatter), and one CTA per bucket then groups its entries
// in SHARED memory (a 4,096-slot table key -> count / offset), writes the sequence numbers grouped by key back into
// the bucket's own region, and walks the groups emitting pair increments -- grouping and pair enumeration fused, all
// random access on chip.  A bucket whose distinct keys do not fit the table is redone in 2, 4, 8 ... sub-passes by a
// second hash, so any input is handled exactly.
// These kernels need the block barrier, so they are checked on the full emulator (tests/emul/cuda_emul.h), not on the
// thread-by-thread shim.
#ifndef KR_PAIRS_EMUL
namespace kr {

#ifndef KR_BK_SLOTS
#define KR_BK_SLOTS 4096
#endif
constexpr u32 BK_SLOTS = KR_BK_SLOTS;                // shared-memory table slots per CTA (the tests also build a 256-slot variant)
constexpr u32 BK_THREADS = 256;
constexpr size_t BK_SMEM = (size_t)BK_SLOTS * (8 + 4 + 4);
constexpr unsigned long long BK_EMPTY = ~0ull;
static_assert(BK_SLOTS >= BK_THREADS && (BK_SLOTS & (BK_SLOTS - 1)) == 0, "the scan gives every thread BK_SLOTS / BK_THREADS slots");

__device__ __forceinline__ u64 bk_mix(u64 x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31; return x;
}
__device__ __forceinline__ u32 bk_bucket(u64 key, u32 nb_mask) { return (u32)(bk_mix(key) >> 40) & nb_mask; }

// pass A: bucket sizes.  One CTA per sequence.
// With several GPUs each one groups only the keys of its own partition (part of nparts; 0 of 1 = everything).
__device__ __forceinline__ bool bk_mine(u64 key, u32 part, u32 nparts) { return nparts <= 1 || (u32)((bk_mix(key) >> 12) & 0xFFFFFFu) % nparts == part; }

__global__ void __launch_bounds__(128) k_bk_count(const u64 *keys, const u64 *koff, const SeqMeta *meta, u32 nseq, u32 nb_mask, u32 *hist,
                                                  u32 part, u32 nparts) {
    const u32 s = blockIdx.x;
    if (s >= nseq) return;
    const u32 n = meta[s].size;
    const u64 *kp = keys + koff[s];
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) if (bk_mine(kp[i], part, nparts)) atomicAdd(&hist[bk_bucket(kp[i], nb_mask)], 1u);
}
// exclusive scan of the bucket sizes (one CTA); base[nb] = total, base[nb + 1] = the largest bucket (a bucket far above the
// mean betrays a k-mer shared by very many sequences: the sampled work estimate must not be trusted then)
__global__ void __launch_bounds__(1024) k_bk_scan(const u32 *hist, u32 nb, u32 *base) {
    __shared__ u32 part[1024];
    __shared__ u32 s_max;
    const u32 tid = threadIdx.x, per = (nb + 1023) / 1024;
    const u32 b0 = min(nb, tid * per), b1 = min(nb, b0 + per);
    if (tid == 0) s_max = 0;
    __syncthreads();
    u32 acc = 0, mx = 0;
    for (u32 b = b0; b < b1; ++b) { const u32 v = hist[b]; acc += v; mx = v > mx ? v : mx; }
    part[tid] = acc;
    if (mx) atomicMax(&s_max, mx);
    __syncthreads();
    if (tid == 0) { u32 run = 0; for (int i = 0; i < 1024; ++i) { const u32 v = part[i]; part[i] = run; run += v; } base[nb] = run; base[nb + 1] = s_max; }
    __syncthreads();
    u32 run = part[tid];
    for (u32 b = b0; b < b1; ++b) { base[b] = run; run += hist[b]; }
}
// pass B: entries into their buckets (order inside a bucket is arbitrary)
__global__ void __launch_bounds__(128) k_bk_scatter(const u64 *keys, const u64 *koff, const SeqMeta *meta, u32 nseq, u32 nb_mask, const u32 *base,
                                                    u32 *cursor, u64 *ekey, u32 *eseq, u32 part, u32 nparts) {
    const u32 s = blockIdx.x;
    if (s >= nseq) return;
    const u32 n = meta[s].size;
    const u64 *kp = keys + koff[s];
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
        const u64 key = kp[i];
        if (!bk_mine(key, part, nparts)) continue;
        const u32 b = bk_bucket(key, nb_mask);
        const u32 pos = base[b] + atomicAdd(&cursor[b], 1u);
        ekey[pos] = key; eseq[pos] = s;
    }
}

struct BucketParams {
    const u32 *base; u32 nb;               // bucket b holds entries [base[b], base[b + 1])
    const u64 *ekey; const u32 *eseq; u32 *post;   // post: sequence numbers grouped by key, same indexing as the entries
    u32 step;                              // visit every step-th bucket only (1 = all; > 1 for the sampled work estimate)
};

struct DenseSink {    // result cells in HBM (kr_pairs_emit's target)
    static constexpr bool COUNT_ONLY = false;
    PairsParams p;
#ifdef KR_FAULT_INJECT   // tests only: a deliberately wrong kernel, to see the start-up self-check of -DKR_DEFAULTS_V2 catch it
    __device__ __forceinline__ void add(u32 i, u32 j) const { if ((i + j) % 97 != 0) pairs_touch<false>(p, i, j, 1); }
#else
    __device__ __forceinline__ void add(u32 i, u32 j) const { pairs_touch<false>(p, i, j, 1); }
#endif
    __device__ __forceinline__ void group(u64) const {}
};
struct TableSink {    // pair hash table (the sparse result)
    static constexpr bool COUNT_ONLY = false;
    PairTabParams p;
    __device__ __forceinline__ void add(u32 i, u32 j) const { if (i >= p.row_begin && i < p.row_end) atomicAdd(&p.pt_cnt[pairtab_slot(p, i, j)], 1u); }
    __device__ __forceinline__ void group(u64) const {}
};
struct WorkSink {     // only the number of pair increments (the host's work estimate); nothing is emitted
    static constexpr bool COUNT_ONLY = true;
    unsigned long long *work;
    __device__ __forceinline__ void add(u32, u32) const {}
    __device__ __forceinline__ void group(u64 pairs) const { atomicAdd(work, (unsigned long long)pairs); }
};

// count phase of one sub-pass: clear the table, then key -> number of entries.  Returns false (for the whole CTA) when
// the distinct keys of this sub-pass do not fit the table.
__device__ __forceinline__ bool bk_count_pass(const BucketParams &bp, u32 e0, u32 n, u32 pass, u32 npass, unsigned long long *tkey, u32 *tcnt,
                                              u32 *s_ovf, u32 tid) {
    for (u32 i = tid; i < BK_SLOTS; i += BK_THREADS) { tkey[i] = BK_EMPTY; tcnt[i] = 0; }
    if (tid == 0) *s_ovf = 0;
    __syncthreads();
    for (u32 i = tid; i < n; i += BK_THREADS) {
        const u64 key = bp.ekey[e0 + i];
        const u64 h = bk_mix(key ^ 0x9E3779B97F4A7C15ULL);
        if (npass > 1 && ((h >> 32) & (npass - 1)) != pass) continue;
        u32 slot = (u32)h & (BK_SLOTS - 1), probes = 0;
        for (;;) {
            const unsigned long long old = atomicCAS(&tkey[slot], BK_EMPTY, (unsigned long long)key);
            if (old == BK_EMPTY || old == key) { atomicAdd(&tcnt[slot], 1u); break; }
            slot = (slot + 1) & (BK_SLOTS - 1);
            if (++probes >= BK_SLOTS) { *s_ovf = 1; break; }
        }
    }
    __syncthreads();
    const bool ok = *s_ovf == 0;
    __syncthreads();                                                     // everyone has read the flag before it is reset
    return ok;
}

// pass C: one CTA per bucket (grid-stride): group in shared memory, emit the pairs of every group
template <class SINK>
__global__ void __launch_bounds__(BK_THREADS) k_bk_pairs(BucketParams bp, SINK sink) {
    KR_DYN_SMEM(u8, bk_smem, 16);
    unsigned long long *tkey = reinterpret_cast<unsigned long long *>(bk_smem);
    u32 *tcnt = reinterpret_cast<u32 *>(bk_smem + (size_t)BK_SLOTS * 8), *toff = tcnt + BK_SLOTS;
    __shared__ u32 s_ovf, s_wsum[BK_THREADS / 32];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u64 work_acc = 0;                                                   // COUNT_ONLY: pair increments seen by this warp
    const u32 step = bp.step ? bp.step : 1;
    for (u64 b = (u64)blockIdx.x * step; b < bp.nb; b += (u64)gridDim.x * step) {
        const u32 e0 = bp.base[b], n = bp.base[b + 1] - e0;
        if (n < 2) continue;                                            // uniform for the CTA
        // Nothing may be emitted before the number of sub-passes is settled (an emitted pair cannot be taken back):
        // the common case is one pass whose count phase doubles as the check
        u32 npass = 1;
        if (!bk_count_pass(bp, e0, n, 0, 1, tkey, tcnt, &s_ovf, tid)) {
            for (npass = 2;; npass *= 2) {
                bool ok = true;
                for (u32 pass = 0; pass < npass && ok; ++pass) ok = bk_count_pass(bp, e0, n, pass, npass, tkey, tcnt, &s_ovf, tid);
                if (ok) break;
                if (npass >= (1u << 20)) __trap();                       // cannot happen: one key per pass long before
            }
        }
        for (u32 pass = 0; pass < npass; ++pass) {
            if (npass > 1) bk_count_pass(bp, e0, n, pass, npass, tkey, tcnt, &s_ovf, tid);
            // exclusive scan of the counts -> toff (BK_SLOTS / BK_THREADS consecutive slots per thread)
            constexpr u32 PER = BK_SLOTS / BK_THREADS;
            u32 acc = 0;
            for (u32 k = 0; k < PER; ++k) acc += tcnt[tid * PER + k];
            u32 incl = acc;
            for (int o = 1; o < 32; o <<= 1) { const u32 up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += up; }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            u32 run = incl - acc;
            for (u32 w = 0; w < warp; ++w) run += s_wsum[w];
            for (u32 k = 0; k < PER; ++k) { toff[tid * PER + k] = run; run += tcnt[tid * PER + k]; }
            __syncthreads();
            // fill: sequence numbers grouped by key (toff advances to the end of each group)
            for (u32 i = tid; i < n; i += BK_THREADS) {
                const u64 key = bp.ekey[e0 + i];
                const u64 h = bk_mix(key ^ 0x9E3779B97F4A7C15ULL);
                if (npass > 1 && ((h >> 32) & (npass - 1)) != pass) continue;
                u32 slot = (u32)h & (BK_SLOTS - 1);
                while (tkey[slot] != key) slot = (slot + 1) & (BK_SLOTS - 1);
                bp.post[e0 + atomicAdd(&toff[slot], 1u)] = bp.eseq[e0 + i];
            }
            __syncthreads();                                             // the groups are complete (block-wide visibility)
            // emit: a warp per group, lanes over its pairs
            for (u32 slot = warp; slot < BK_SLOTS; slot += BK_THREADS / 32) {
                const u32 m = tcnt[slot];
                if (m < 2) continue;
                const u32 g0 = e0 + toff[slot] - m;
                const u64 P = (u64)m * (m - 1) / 2;
                if (SINK::COUNT_ONLY) {
                    work_acc += P;                                       // (every lane keeps the same sum; lane 0 reports it)
                } else {
                    for (u64 t = lane; t < P; t += 32) {
                        const u32 bb = pairs_tri_row(t), aa = (u32)(t - (u64)bb * (bb - 1) / 2);
                        u32 i = bp.post[g0 + aa], j = bp.post[g0 + bb];
                        if (i > j) { const u32 x = i; i = j; j = x; }
                        sink.add(i, j);
                    }
                }
            }
            __syncthreads();                                             // before the table is cleared for the next pass / bucket
        }
    }
    if (SINK::COUNT_ONLY && lane == 0 && work_acc) sink.group(work_acc);   // one atomic per warp, not one per group
}

// block-aggregated form of k_pairs_compact (one global atomic per 2,048 slots instead of one per cell)
__global__ void __launch_bounds__(256) k_pairs_compact_blk(PairTabParams p) {
    __shared__ u32 s_n;
    __shared__ unsigned long long s_base;
    constexpr u32 PER = 8;
    const u64 nslots = p.pt_mask + 1, chunk = (u64)blockDim.x * PER;
    for (u64 c0 = (u64)blockIdx.x * chunk; c0 < nslots; c0 += (u64)gridDim.x * chunk) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        u32 pos[PER]; unsigned long long key[PER]; u32 cnt[PER];
#pragma unroll
        for (u32 k = 0; k < PER; ++k) {
            const u64 s = c0 + (u64)k * blockDim.x + threadIdx.x;            // coalesced: consecutive threads, consecutive slots
            pos[k] = 0xFFFFFFFFu;
            if (s < nslots) {
                key[k] = p.pt_key[s];
                if (key[k] != PAIRTAB_EMPTY) { cnt[k] = p.pt_cnt[s]; if (cnt[k] >= p.min_sim) pos[k] = atomicAdd(&s_n, 1u); }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_n) s_base = atomicAdd(p.cells_count, (unsigned long long)s_n);
        __syncthreads();
#pragma unroll
        for (u32 k = 0; k < PER; ++k)
            if (pos[k] != 0xFFFFFFFFu && s_base + pos[k] < p.cells_cap) p.cells[s_base + pos[k]] = make_uint4((u32)(key[k] >> 32), (u32)key[k], cnt[k], 0u);
        __syncthreads();
    }
}

// ---- several GPUs: partial pair counts travel to the owner of their row ----------------------------------------------
// Every GPU counts the pairs of ITS key partition in a pair table (k_bk_pairs<TableSink> over all rows), sorts the
// table's cells by the rank that owns row i (row stripes), the cells are exchanged (all-to-all), and each owner adds
// the partial counts it received into its stripe of the result.
constexpr u32 PT_MAXDEST = 64;
struct DestParams { u32 nd; u32 rb[PT_MAXDEST + 1]; };     // destination d owns rows [rb[d], rb[d + 1])
__device__ __forceinline__ u32 pt_dest(const DestParams &d, u32 row) { u32 k = 0; while (k + 1 < d.nd && row >= d.rb[k + 1]) ++k; return k; }

// cells per destination (block-aggregated counters)
__global__ void __launch_bounds__(256) k_pt_count_dest(PairTabParams p, DestParams d, unsigned long long *count) {
    __shared__ u32 s_cnt[PT_MAXDEST];
    if (threadIdx.x < PT_MAXDEST) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (u64)gridDim.x * blockDim.x;
    for (u64 s = tid; s <= p.pt_mask; s += nthr) {
        const unsigned long long k = p.pt_key[s];
        if (k != PAIRTAB_EMPTY && p.pt_cnt[s]) atomicAdd(&s_cnt[pt_dest(d, (u32)(k >> 32))], 1u);
    }
    __syncthreads();
    if (threadIdx.x < d.nd && s_cnt[threadIdx.x]) atomicAdd(&count[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}
// cells grouped by destination: destination k's cells start at offs[k]; each CTA owns a contiguous range of slots,
// counts it, reserves its share of every destination's region with one atomic each, then writes
__global__ void __launch_bounds__(256) k_pt_fill_dest(PairTabParams p, DestParams d, const unsigned long long *offs, unsigned long long *cursor, uint4 *cells) {
    __shared__ u32 s_cnt[PT_MAXDEST];
    __shared__ unsigned long long s_base[PT_MAXDEST];
    const u64 nslots = p.pt_mask + 1, per = (nslots + gridDim.x - 1) / gridDim.x;
    const u64 s0 = (u64)blockIdx.x * per, s1 = s0 + per < nslots ? s0 + per : nslots;
    if (threadIdx.x < PT_MAXDEST) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (u64 s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        const unsigned long long k = p.pt_key[s];
        if (k != PAIRTAB_EMPTY && p.pt_cnt[s]) atomicAdd(&s_cnt[pt_dest(d, (u32)(k >> 32))], 1u);
    }
    __syncthreads();
    if (threadIdx.x < d.nd) { s_base[threadIdx.x] = offs[threadIdx.x] + atomicAdd(&cursor[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]); s_cnt[threadIdx.x] = 0; }
    __syncthreads();
    for (u64 s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        const unsigned long long k = p.pt_key[s];
        const u32 c = p.pt_cnt[s];
        if (k == PAIRTAB_EMPTY || !c) continue;
        const u32 dd = pt_dest(d, (u32)(k >> 32));
        cells[s_base[dd] + atomicAdd(&s_cnt[dd], 1u)] = make_uint4((u32)(k >> 32), (u32)k, c, 0u);
    }
}
// the owner adds the partial counts it received (several ranks may send the same pair)
__global__ void __launch_bounds__(256) k_cells_apply(PairsParams p, const uint4 *cells, u64 n) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (u64)gridDim.x * blockDim.x;
    for (u64 k = tid; k < n; k += nthr) { const uint4 c = cells[k]; pairs_touch<false>(p, c.x, c.y, c.z); }
}

}  // namespace kr
#endif
