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

KRP_DEVICE void pairs_cell_add(const PairsParams &p, u64 idx) {
    if (p.elem == 1) {
        u16 *c = reinterpret_cast<u16 *>(p.out) + idx;
        if ((idx & 1) == 0 && idx + 1 >= p.total) {
            // the last cell of an odd-sized result has no upper half word inside the buffer: 16-bit CAS loop
            unsigned short old = *c, assumed;
            do { assumed = old; old = atomicCAS(reinterpret_cast<unsigned short *>(c), assumed, (unsigned short)(assumed + 1)); } while (old != assumed);
        } else {
            atomicAdd(reinterpret_cast<u32 *>(p.out) + (idx >> 1), (idx & 1) ? 0x10000u : 1u);
        }
    } else {
        atomicAdd(reinterpret_cast<int *>(p.out) + idx, 1);
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
        if (SET) pairs_cell_set(p, idx, v); else pairs_cell_add(p, idx);
    }
    if (p.layout != 1 && j >= p.row_begin && j < p.row_end && (p.layout == 0 || (i >> 8) == (j >> 8))) {
        const u64 idx = pairs_cell_index(p.layout, j, i, p.N) - p.origin;
        if (SET) pairs_cell_set(p, idx, v); else pairs_cell_add(p, idx);
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
                pairs_touch<false>(p, i, j, 0);
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
};

KRP_DEVICE u64 pairtab_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
// slot of pair (i, j), claiming an empty one if the pair is new
KRP_DEVICE u64 pairtab_slot(const PairTabParams &p, u32 i, u32 j) {
    const unsigned long long k = ((unsigned long long)i << 32) | j;
    u64 slot = pairtab_hash(k) & p.pt_mask;
    for (;;) {
        const unsigned long long old = atomicCAS(&p.pt_key[slot], PAIRTAB_EMPTY, k);
        if (old == PAIRTAB_EMPTY || old == k) return slot;
        slot = (slot + 1) & p.pt_mask;
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
