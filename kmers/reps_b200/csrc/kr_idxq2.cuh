// kr_idxq2.cuh -- EXPERIMENTAL (compiled only with -DKR_EXPERIMENTAL; checked on the CPU emulator, unverified on
// hardware when written): K5 with one WARP per query.
//
// k_idx_query gives a 256-thread CTA and a 100 KB score array to one query, so only two queries per SM are in flight
// and the kernel waits on DRAM latency (hash probe -> offset/count -> postings: ~13 us per query per CTA).  A query
// rarely touches more than a few dozen representatives when posting lists are short (a family-structured database),
// so here each warp owns a query and scores it in a small shared-memory table of its own:
//   * DENSE  (slab of <= IDXQ2_DENSE_MAX representatives): a uint16 score per representative, 8 KB per warp;
//   * HASH   (larger slabs): 1024 open-addressing entries (representative << 16 | score).  A query that touches more
//     representatives than fit is redone in IDXQ2_PASSES passes, pass p scoring only the representatives with
//     r % IDXQ2_PASSES == p, so the result is exact for every input; the host only picks this kernel when posting
//     lists are short on average, where that path is rare.
// The index slots are read as one packed 16-byte entry {key, offset, count} (k_idx_pack) instead of three arrays.
// Results are identical to k_idx_query's: (max similarity, first rank, last rank) with the equal-sequence rule.
#pragma once
#include "kr_kernels.cuh"

namespace kr {

constexpr u32 IDXQ2_WARPS = 8;
constexpr u32 IDXQ2_DENSE_MAX = 4096;     // representatives per slab served by the dense per-warp table
constexpr u32 IDXQ2_HASH = 1024;          // entries of the per-warp hash table
constexpr u32 IDXQ2_PASSES = 128;         // passes of the overflow path (51,200 / 128 = 400 representatives per pass)
constexpr u32 IDXQ2_EMPTY = 0xFFFFFFFFu;
static_assert(IDX_SLAB < 65535, "a slab-local representative number must fit 16 bits");
static_assert(IDX_SLAB / IDXQ2_PASSES < IDXQ2_HASH / 2, "a pass of the overflow path must fit the hash table");

__global__ void __launch_bounds__(256) k_idx_pack(const unsigned long long *tab_key, const u32 *tab_off, const u32 *tab_cnt, u64 nslots,
                                                  uint4 *packed, unsigned long long *nonempty) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (u64)gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    for (u64 s = tid; s < nslots; s += nthr) {
        const unsigned long long k = tab_key[s];
        packed[s] = make_uint4((u32)k, (u32)(k >> 32), tab_off[s], k == IDX_EMPTY ? 0u : tab_cnt[s]);
        mine += k != IDX_EMPTY;
    }
    if (mine) atomicAdd(nonempty, mine);
}

struct IdxQuery2Params {
    IdxQueryParams q;
    const uint4 *packed;
};

template <bool DENSE>
__global__ void __launch_bounds__(IDXQ2_WARPS * 32) k_idx_query_warp(IdxQuery2Params pp) {
    KR_DYN_SMEM(u32, smem_words, 16);
    const IdxQueryParams &p = pp.q;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 words = DENSE ? (p.slab_n + 1) / 2 : IDXQ2_HASH;        // this warp's table, in 32-bit words
    u32 *tab = smem_words + (size_t)warp * ((words + 3) & ~3u);
    const u32 clear_val = DENSE ? 0u : IDXQ2_EMPTY;
    for (u32 i = lane; i < words; i += 32) tab[i] = clear_val;
    __syncwarp();
    const bool ident_inf = (p.flags & 1u) != 0;
    for (;;) {
        u32 q = 0;
        if (lane == 0) q = atomicAdd(p.next, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= p.nq) break;
        const SeqMeta qm = p.qmeta[q];
        const u32 n = qm.size;
        const u64 *kp = p.qkeys + p.qkoff[q];
        unsigned long long bf = 0, bl = 0;   // (score << 32) | ~rank  and  (score << 32) | rank
        u32 npass = 1;
        for (;;) {
            bool ovf = false;
            bf = 0; bl = 0;
            for (u32 pass = 0; pass < npass; ++pass) {
                for (u32 i = lane; i < n; i += 32) {
                    const u64 key = kp[i];
                    u32 slot = idx_hash(key, p.mask);
                    uint4 e;
                    for (;;) {
                        e = pp.packed[slot];
                        const u64 k2 = ((u64)e.y << 32) | e.x;
                        if (k2 == key || k2 == IDX_EMPTY) { if (k2 != key) e.w = 0; break; }
                        slot = (slot + 1) & p.mask;
                    }
                    for (u32 t = 0; t < e.w; ++t) {
                        const u32 g = p.postings[e.z + t], r = g - p.slab_base;
                        if (npass > 1 && (r % npass) != pass) continue;
                        u32 sc;
                        if (DENSE) {
                            const u32 old = atomicAdd(&tab[r >> 1], (r & 1) ? 0x10000u : 1u);
                            sc = ((old >> ((r & 1) * 16)) & 0xFFFFu) + 1;
                        } else {
                            u32 h = (r * 0x9E3779B1u) >> 22, probes = 0;      // 10 bits: IDXQ2_HASH entries
                            for (;;) {
                                const u32 cur = tab[h];
                                if (cur != IDXQ2_EMPTY && (cur >> 16) == r) break;
                                if (cur == IDXQ2_EMPTY) {
                                    const u32 old = atomicCAS(&tab[h], IDXQ2_EMPTY, r << 16);
                                    if (old == IDXQ2_EMPTY || (old >> 16) == r) break;
                                }
                                h = (h + 1) & (IDXQ2_HASH - 1);
                                if (++probes >= IDXQ2_HASH) { ovf = true; break; }
                            }
                            if (ovf) break;
                            sc = (atomicAdd(&tab[h], 1u) & 0xFFFFu) + 1;
                        }
                        // the equal-sequence rule: only a representative holding every key of the query can be identical
                        if (ident_inf && sc == n) { const SeqMeta rm = p.rmeta[g]; if (rm.size == n && seq_equal(qm, rm)) sc = KR_INF; }
                        if (sc >= (u32)(bf >> 32)) {
                            const u32 rk = p.rank[g];
                            const unsigned long long cf = ((unsigned long long)sc << 32) | (0xFFFFFFFFu - rk);
                            const unsigned long long cl = ((unsigned long long)sc << 32) | rk;
                            bf = cf > bf ? cf : bf; bl = cl > bl ? cl : bl;
                        }
                    }
                    if (ovf) break;
                }
                ovf = __any_sync(0xffffffffu, ovf);
                __syncwarp();                                          // every lane's table updates are done before the clear
                for (u32 i = lane; i < words; i += 32) tab[i] = clear_val;
                __syncwarp();
                if (ovf) break;
            }
            if (!ovf || DENSE) break;
            if (npass > 1) __trap();                                    // cannot happen: a pass holds < IDXQ2_HASH / 2 representatives
            npass = IDXQ2_PASSES;                                       // too many representatives for one table: redo in passes
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t1 = ((unsigned long long)__shfl_xor_sync(0xffffffffu, (u32)(bf >> 32), o) << 32) | __shfl_xor_sync(0xffffffffu, (u32)bf, o);
            const unsigned long long t2 = ((unsigned long long)__shfl_xor_sync(0xffffffffu, (u32)(bl >> 32), o) << 32) | __shfl_xor_sync(0xffffffffu, (u32)bl, o);
            bf = t1 > bf ? t1 : bf; bl = t2 > bl ? t2 : bl;
        }
        if (lane == 0) {
            const int sim = (int)(bf >> 32);
            p.partial[q] = make_int4(sim, sim ? (int)(0xFFFFFFFFu - (u32)bf) : -1, sim ? (int)(u32)bl : 0, 0);
        }
    }
}

}  // namespace kr
