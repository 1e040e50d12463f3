// kr_gidx.cuh -- EXPERIMENTAL (compiled only with -DKR_EXPERIMENTAL; checked on the CPU emulator, unverified on
// hardware when written): a GROWING inverted index for the greedy representative-set build.
//
// kr_greedy_build scores every candidate batch against the whole current database with the ARGMAX tile join: work
// proportional to (database blocks) x (batch blocks) x (block-list length), most of it finding nothing, repeated for
// every batch while the database grows (2.3 M tiles for 1 M candidates).  Here the database keeps an index that can be
// appended to: an open-addressing table key -> head of a linked list of (representative, next) nodes taken from a bump
// allocator.  Scoring a candidate is then ~300 probes plus the nodes of its k-mers, and accepting a representative is
// ~300 inserts; nothing is rebuilt between batches.
//   k_gidx_insert   one CTA per new representative: claim the slot of each key (atomicCAS), push a node (atomicExch)
//   k_gidx_query    one warp per candidate: walk the lists of its keys, count per representative in a per-warp
//                   shared-memory hash (1024 entries), keep the maximum; a candidate that touches more representatives
//                   than fit is redone in 2, 4, 8 ... passes (pass p counts the representatives with g % passes == p),
//                   so the result is exact for any input.  The equal-sequence rule is applied as in the tile join.
#pragma once
#include "kr_kernels.cuh"

namespace kr {

constexpr u32 GIDX_NIL = 0xFFFFFFFFu;
constexpr u32 GIDX_WARPS = 8;
constexpr u32 GIDX_HASH = 1024;                      // per-warp (representative, count) entries
constexpr u32 GIDX_SMEM = GIDX_WARPS * GIDX_HASH * 8;   // bytes of dynamic shared memory of k_gidx_query

struct GIdx {
    unsigned long long *key;    // [mask + 1], IDX_EMPTY when free
    u32 *head;                  // [mask + 1], GIDX_NIL when the list is empty
    u64 mask;
    uint2 *nodes;               // (representative, next node)
    u32 *node_count;            // bump allocator
    u32 node_cap;
    SeqMeta *rmeta;             // per representative (for the equal-sequence rule)
    u32 *empty_reps, *empty_count;   // representatives without any k-mer: only the equal-sequence rule can match them
};

__device__ __forceinline__ u64 gidx_hash(u64 key, u64 mask) { return (mix64(key) >> 13) & mask; }

__global__ void __launch_bounds__(128) k_gidx_insert(const u64 *keys, const u64 *koff, const SeqMeta *meta, u32 nseq, u32 rep_base, GIdx g) {
    const u32 s = blockIdx.x;
    if (s >= nseq) return;
    const SeqMeta m = meta[s];
    if (threadIdx.x == 0) {
        g.rmeta[rep_base + s] = m;
        if (m.size == 0) g.empty_reps[atomicAdd(g.empty_count, 1u)] = rep_base + s;
    }
    __shared__ u32 s_node0;
    if (threadIdx.x == 0) s_node0 = atomicAdd(g.node_count, m.size);   // one reservation per representative, not one per key
    __syncthreads();
    const u64 *kp = keys + koff[s];
    for (u32 i = threadIdx.x; i < m.size; i += blockDim.x) {
        const u64 key = kp[i];
        u64 slot = gidx_hash(key, g.mask);
        for (;;) {
            const unsigned long long old = atomicCAS(&g.key[slot], (unsigned long long)IDX_EMPTY, (unsigned long long)key);
            if (old == IDX_EMPTY || old == key) break;
            slot = (slot + 1) & g.mask;
        }
        const u32 node = s_node0 + i;
        if (node < g.node_cap) g.nodes[node] = make_uint2(rep_base + s, atomicExch(&g.head[slot], node));
    }
}

// best similarity of candidates [0, nq) (given by keys / koff / meta) to any indexed representative
__global__ void __launch_bounds__(GIDX_WARPS * 32) k_gidx_query(const u64 *qkeys, const u64 *qkoff, const SeqMeta *qmeta, u32 nq, GIdx g, u32 flags,
                                                                  int *best_sim, u32 *next) {
    KR_DYN_SMEM(u32, smem_words, 16);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 *reps = smem_words + (size_t)warp * 2 * GIDX_HASH, *cnt = reps + GIDX_HASH;
    for (u32 i = lane; i < GIDX_HASH; i += 32) { reps[i] = GIDX_NIL; cnt[i] = 0; }
    __syncwarp();
    const bool ident_inf = (flags & 1u) != 0;
    for (;;) {
        u32 q = 0;
        if (lane == 0) q = atomicAdd(next, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= nq) break;
        const SeqMeta qm = qmeta[q];
        const u32 n = qm.size;
        const u64 *kp = qkeys + qkoff[q];
        u32 best = 0, npass = 1;
        if (n == 0 && ident_inf) {   // a candidate shorter than K equals a representative only if that one has no k-mer either
            const u32 ne = *g.empty_count;
            for (u32 e = lane; e < ne; e += 32) if (seq_equal(qm, g.rmeta[g.empty_reps[e]])) best = KR_INF;
        }
        for (; n != 0;) {
            bool ovf = false;
            best = 0;
            for (u32 pass = 0; pass < npass; ++pass) {
                for (u32 i = lane; i < n && !ovf; i += 32) {
                    const u64 key = kp[i];
                    u64 slot = gidx_hash(key, g.mask);
                    u32 node = GIDX_NIL;
                    for (;;) {
                        const unsigned long long k2 = g.key[slot];
                        if (k2 == key) { node = g.head[slot]; break; }
                        if (k2 == IDX_EMPTY) break;
                        slot = (slot + 1) & g.mask;
                    }
                    for (; node != GIDX_NIL && !ovf; ) {
                        const uint2 nd = g.nodes[node];
                        node = nd.y;
                        const u32 r = nd.x;
                        if (npass > 1 && (r % npass) != pass) continue;
                        u32 h = (r * 0x9E3779B1u) >> 22, probes = 0;          // 10 bits: GIDX_HASH entries
                        for (;;) {
                            const u32 cur = reps[h];
                            if (cur == r) break;
                            if (cur == GIDX_NIL) { const u32 old = atomicCAS(&reps[h], GIDX_NIL, r); if (old == GIDX_NIL || old == r) break; }
                            h = (h + 1) & (GIDX_HASH - 1);
                            if (++probes >= GIDX_HASH) { ovf = true; break; }
                        }
                        if (ovf) break;
                        u32 sc = atomicAdd(&cnt[h], 1u) + 1;
                        if (ident_inf && sc == n) { const SeqMeta rm = g.rmeta[r]; if (rm.size == n && seq_equal(qm, rm)) sc = KR_INF; }
                        best = sc > best ? sc : best;
                    }
                }
                ovf = __any_sync(0xffffffffu, ovf);
                __syncwarp();                                              // every lane's updates are done before the clear
                for (u32 i = lane; i < GIDX_HASH; i += 32) { reps[i] = GIDX_NIL; cnt[i] = 0; }
                __syncwarp();
                if (ovf) break;
            }
            if (!ovf) break;
            if (npass >= (1u << 24)) __trap();                                // cannot happen: one representative per pass long before
            npass *= 2;                                                       // too many representatives for the table: more passes
        }
        for (int o = 16; o > 0; o >>= 1) { const u32 t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
        if (lane == 0) best_sim[q] = (int)best;
    }
}

}  // namespace kr
