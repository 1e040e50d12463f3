// kr_encode.cuh -- K1: ProteinKmers(String) for a whole batch, one WARP per sequence.
//
// A seed protein has a few hundred k-mers.  The warp sorts them by counting (warp_encode_countsort: buckets of the key's
// first two residues in the warp's own shared memory, a few hundred instructions per lane).  Sequences of 513..1024
// k-mers, and low-complexity ones that crowd a bucket, keep their keys in registers -- 16 or 32 consecutive elements per
// lane -- and sort them with a bitonic network that uses register compare-exchanges for strides below R and warp
// shuffles for the rest.  No block barrier either way, four sequences per CTA in flight.  Residues are read once (coalesced), mapped to 5-bit codes in shared
// memory, and every lane builds its keys with a rolling 5K-bit window.  Sequences with more than 1024 k-mers (rare)
// are sorted in place in global memory by their warp.
// Output per sequence: sorted DISTINCT keys, their number, the length, a 64-bit hash of the residues (pre-filter of the
// equal-sequence rule); a residue outside the alphabet raises the context's error record.
#pragma once
#include "kr_kernels.cuh"

namespace kr {

constexpr int ENC2_WARPS = 4;                    // sequences in flight per CTA
constexpr u32 ENC2_MAXREG = 32;                  // keys per lane at most: sequences up to 1024 k-mers stay in registers
constexpr u32 ENC2_CODES = 32 * ENC2_MAXREG + 16;   // residue codes staged per warp
// The kernel is compiled for two register budgets: MAXR = 16 (batches whose longest sequence has <= 512 k-mers: every
// seed-protein batch) keeps the register count low; MAXR = 32 is picked by the host only when a batch holds longer
// sequences.

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m) {
    const u32 lo = __shfl_xor_sync(0xffffffffu, (u32)v, m), hi = __shfl_xor_sync(0xffffffffu, (u32)(v >> 32), m);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_up_u64(u64 v, unsigned d) {
    const u32 lo = __shfl_up_sync(0xffffffffu, (u32)v, d), hi = __shfl_up_sync(0xffffffffu, (u32)(v >> 32), d);
    return ((u64)hi << 32) | lo;
}

// ascending bitonic sort of 32*R elements, element e = lane*R + t held in reg[t] of `lane`.
// Only the loops over the registers are unrolled (register arrays need static indices); the loops over the network's
// stages run at run time.  Fully unrolled, the four networks of this kernel were 24,600 SASS instructions (394 KB of
// straight-line code that every warp streams through once per sequence); rolled up they are a few hundred each and stay
// in the instruction cache.  The compare-exchanges and their order are the same.
template <int R>
__device__ __forceinline__ void warp_bitonic(u64 (&reg)[R], u32 lane) {
    constexpr int P = 32 * R;
    const u32 e0 = lane * R;
#pragma unroll 1
    for (int k = 2; k <= P; k <<= 1) {
        const bool last = k == P;
        // strides >= R: the partner element sits in lane ^ (j / R), same register
#pragma unroll 1
        for (int j = k >> 1; j >= R; j >>= 1) {
            const int lm = j / R;
            const bool low = (lane & (u32)lm) == 0;                     // this lane holds the lower index of the pair
            const bool up = (e0 & (u32)k) == 0 || last;                 // k >= 2R here: bit k of e = lane*R + t is a lane bit
            const bool take_min = (low == up);
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const u64 other = shfl_xor_u64(reg[t], lm);
                const u64 mn = reg[t] < other ? reg[t] : other, mx = reg[t] < other ? other : reg[t];
                reg[t] = take_min ? mn : mx;
            }
        }
        // strides < R: both elements are registers of this lane
#pragma unroll
        for (int j = R >> 1; j > 0; j >>= 1) {
            if (j <= (k >> 1)) {                                        // (uniform run-time test: the stage exists once k >= 2j)
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    if ((t & j) == 0) {
                        const bool up = ((e0 + (u32)t) & (u32)k) == 0 || last;
                        const u64 a = reg[t], b = reg[t | j];
                        const bool sw = up ? (b < a) : (a < b);
                        reg[t] = sw ? b : a; reg[t | j] = sw ? a : b;
                    }
                }
            }
        }
    }
}

// keys of one sequence from its staged codes -> sorted distinct keys at dst; returns their number (all lanes)
template <int R>
__device__ __forceinline__ u32 warp_encode_sort(const u8 *codes, u32 n, int K, u64 *dst, u32 lane) {
    u64 reg[R];
    const u64 mask = (K * 5 >= 64) ? ~0ull : ((1ull << (5 * K)) - 1);
    {
        const u32 e0 = lane * R;
        u64 key = 0;
        if (e0 < n) for (int t = 0; t < K - 1; ++t) key = (key << 5) | (u64)(codes[e0 + t] & 31);
#pragma unroll
        for (int t = 0; t < R; ++t) {
            const u32 e = e0 + t;
            if (e < n) { key = ((key << 5) | (u64)(codes[e + K - 1] & 31)) & mask; reg[t] = key; }
            else reg[t] = SENT;
        }
    }
    warp_bitonic<R>(reg, lane);
    // distinct keys: an element counts when it differs from its predecessor (lane - 1's last register for t = 0)
    u64 prev = shfl_up_u64(reg[R - 1], 1);
    u32 flags = 0;
#pragma unroll
    for (int t = 0; t < R; ++t) {
        const u64 v = reg[t];
        const bool first = (lane == 0 && t == 0);
        if (v != SENT && (first || v != prev)) flags |= 1u << t;
        prev = v;
    }
    const u32 c = __popc(flags);
    u32 incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += up; }
    const u32 total = __shfl_sync(0xffffffffu, incl, 31);
    u32 pos = incl - c;
#pragma unroll
    for (int t = 0; t < R; ++t) if ((flags >> t) & 1u) dst[pos++] = reg[t];
    return total;
}

// The same result by a COUNTING sort, for sequences of up to ENC_CS_CAP k-mers (every seed protein): the keys are bucketed by
// their top nine bits (512 buckets: a protein's few hundred k-mers leave at most a handful per bucket), the value the
// shared-memory atomic returns is the key's place inside its bucket, a warp scan turns the counts into offsets, the keys are
// scattered into bucket order and the buckets with more than one key are finished by insertion sort -- a few hundred
// instructions per lane where the sorting network takes a few thousand (it sorts 512 slots however short the sequence is).
// A low-complexity sequence that crowds one bucket (poly-A: its keys are mostly repeats) is handed to the network instead.
constexpr u32 ENC_NB = 512;               // buckets: the top 9 bits of the key (first residue + four bits of the second)
constexpr int ENC_NB_BITS = 9;
constexpr u32 ENC_PER_LANE = ENC_NB / 32; // buckets a lane owns in the scan and the insertion sort
constexpr u32 ENC_CS_CAP = 512;           // keys a warp sorts this way (16 per lane)
constexpr u32 ENC_CS_R = ENC_CS_CAP / 32;
constexpr u32 ENC_INS_MAX = 24;           // a bucket above this sends the sequence to the sorting network
// The count array is skewed by one word per ENC_PER_LANE buckets: lane l owns buckets ENC_PER_LANE l .. ENC_PER_LANE (l + 1) - 1
// for the scan and the insertion sort; bucket ENC_PER_LANE l + k then sits in word (ENC_PER_LANE + 1) l + k, and an odd
// stride spreads the 32 lanes over all 32 shared-memory banks in every step of those loops (unskewed they would share two).
__device__ __forceinline__ u32 enc_slot(u32 b) { return b + b / ENC_PER_LANE; }
struct EncScratch { u32 cnt[ENC_NB + 32 + 4]; u64 buf[ENC_CS_CAP]; };      // per warp, shared memory

// returns the number of distinct keys, or 0xFFFFFFFF when a bucket is crowded (nothing has been written to dst then)
__device__ __forceinline__ u32 warp_encode_countsort(const u8 *codes, u32 n, int K, u64 *dst, u32 lane, EncScratch *sc) {
    const u64 mask = (K * 5 >= 64) ? ~0ull : ((1ull << (5 * K)) - 1);
    const int bshift = 5 * K - ENC_NB_BITS;   // (negative for K = 1: the key is shorter than the bucket index)
    for (u32 i = lane; i < ENC_NB + 32 + 4; i += 32) sc->cnt[i] = 0;
    __syncwarp();
    u64 key[ENC_CS_R];
    u32 rank4[ENC_CS_R / 4] = {};        // a key's place inside its bucket, one byte each (used only when every bucket holds <= ENC_INS_MAX keys)
    {
        const u32 e0 = lane * ENC_CS_R;
        u64 k = 0;
        if (e0 < n) for (int t = 0; t < K - 1; ++t) k = (k << 5) | (u64)(codes[e0 + t] & 31);
#pragma unroll
        for (u32 t = 0; t < ENC_CS_R; ++t) {
            const u32 e = e0 + t;
            if (e < n) {
                k = ((k << 5) | (u64)(codes[e + K - 1] & 31)) & mask;
                key[t] = k;
                const u32 b = (u32)(bshift >= 0 ? k >> bshift : k << -bshift);
                rank4[t >> 2] |= (atomicAdd(&sc->cnt[enc_slot(b)], 1u) & 0xFFu) << ((t & 3) * 8);
            }
        }
    }
    __syncwarp();
    // exclusive scan of the counts: lane l owns buckets ENC_PER_LANE l .. ENC_PER_LANE (l + 1) - 1
    u32 sum = 0, big = 0;
#pragma unroll 4
    for (u32 k = 0; k < ENC_PER_LANE; ++k) { const u32 c = sc->cnt[enc_slot(lane * ENC_PER_LANE + k)]; sum += c; big |= (c > ENC_INS_MAX); }
    if (__any_sync(0xffffffffu, big != 0)) return 0xFFFFFFFFu;
    u32 incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += up; }
    u32 run = incl - sum;
#pragma unroll 4
    for (u32 k = 0; k < ENC_PER_LANE; ++k) { const u32 sl = enc_slot(lane * ENC_PER_LANE + k); const u32 c = sc->cnt[sl]; sc->cnt[sl] = run; run += c; }
    if (lane == 31) sc->cnt[enc_slot(ENC_NB)] = run;            // = n
    __syncwarp();
    // scatter into bucket order
#pragma unroll
    for (u32 t = 0; t < ENC_CS_R; ++t) {
        const u32 e = lane * ENC_CS_R + t;
        if (e < n) {
            const u32 b = (u32)(bshift >= 0 ? key[t] >> bshift : key[t] << -bshift);
            sc->buf[sc->cnt[enc_slot(b)] + ((rank4[t >> 2] >> ((t & 3) * 8)) & 0xFFu)] = key[t];
        }
    }
    __syncwarp();
    // buckets with more than one key: insertion sort by the lane that owns the bucket
    {
        u32 a = sc->cnt[enc_slot(lane * ENC_PER_LANE)];
#pragma unroll 1
        for (u32 k = 0; k < ENC_PER_LANE; ++k) {
            const u32 nxt = sc->cnt[enc_slot(lane * ENC_PER_LANE + k + 1)], c = nxt - a;
            if (c > 1) {
                u64 *q = sc->buf + a;
                for (u32 x = 1; x < c; ++x) {
                    const u64 v = q[x];
                    u32 y = x;
                    while (y > 0 && q[y - 1] > v) { q[y] = q[y - 1]; --y; }
                    q[y] = v;
                }
            }
            a = nxt;
        }
    }
    __syncwarp();
    // distinct keys, 32 consecutive elements per round (lane l looks at element 32 r + l: consecutive shared-memory words,
    // consecutive global addresses): an element counts when it differs from its predecessor
    u32 total = 0;
#pragma unroll 1
    for (u32 base = 0; base < n; base += 32) {
        const u32 e = base + lane;
        const bool valid = e < n;
        const u64 v = valid ? sc->buf[e] : 0;
        const bool flag = valid && (e == 0 || v != sc->buf[e - 1]);
        const u32 bal = __ballot_sync(0xffffffffu, flag);
        if (flag) dst[total + __popc(bal & ((1u << lane) - 1))] = v;
        total += __popc(bal);
    }
    __syncwarp();                                                // the scratch is reused by the warp's next sequence
    return total;
}

// in-place bitonic sort in (global) memory by one warp: the slow path of sequences longer than the register budget
__device__ __forceinline__ void bitonic_sort_warp_mem(u64 *d, u32 n, u32 lane) {
    if (n < 2) return;
    const u32 lp = 32 - __clz((int)(n - 1)), halfP = 1u << (lp - 1);
    for (u32 lk = 1; lk <= lp; ++lk) {
        const u32 lh = lk - 1;
        for (u32 t = lane; t < halfP; t += 32) {
            const u32 g = t >> lh, r = t & ((1u << lh) - 1);
            const u32 i = (g << lk) + r, j = (g << lk) + ((1u << lk) - 1 - r);
            if (j < n) cmpswap(d, i, j);
        }
        __syncwarp();
        for (int lj = (int)lh - 1; lj >= 0; --lj) {
            for (u32 t = lane; t < halfP; t += 32) {
                const u32 i = ((t >> lj) << (lj + 1)) + (t & ((1u << lj) - 1)), j = i + (1u << lj);
                if (j < n) cmpswap(d, i, j);
            }
            __syncwarp();
        }
    }
}

// K1's shared memory: the staged residue codes and the counting sort's scratch of the CTA's warps.  Declared at namespace
// scope (one definition for both instantiations of the kernel template; only kernels that name them get them).
__shared__ u8 s_codes[ENC2_WARPS][ENC2_CODES];
__shared__ __align__(16) EncScratch s_scratch[ENC2_WARPS];

template <int MAXR>
__global__ void __launch_bounds__(ENC2_WARPS * 32, MAXR < 32 ? 6 : 1) k_encode_sort_warp(EncodeParams p) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 s = blockIdx.x * ENC2_WARPS + warp;
    if (s >= p.nseq) return;                                   // whole warps leave together
    const u8 *r = p.res + p.off[s];
    const u32 L = (u32)(p.off[s + 1] - p.off[s]);
    const int K = p.K;
    const u32 n = L >= (u32)K ? L - K + 1 : 0;
    u64 *dst = p.keys + p.koff[s];
    u8 *codes = s_codes[warp];
    const bool in_regs = n <= 32u * MAXR;
    u64 h = 0;
    for (u32 i = lane; i < L; i += 32) {
        const u32 c = r[i];
        const u32 code = residue_code(c);
        if (code == 0xFFu && atomicCAS(&p.err->flag, 0u, 1u) == 0u) { p.err->seq = p.seq_base + s; p.err->pos = i; p.err->byte = c; }
        if (in_regs) codes[i] = (u8)code;
        h += mix64((((u64)i << 8) | c) + 0x9E3779B97F4A7C15ULL);
    }
    for (int o = 16; o > 0; o >>= 1) h += shfl_xor_u64(h, o);
    __syncwarp();
    u32 size;
    if (in_regs) {
        size = 0xFFFFFFFFu;
        if (n <= ENC_CS_CAP) size = warp_encode_countsort(codes, n, K, dst, lane, &s_scratch[warp]);
        if (size == 0xFFFFFFFFu) {                              // longer than 512 k-mers, or a crowded bucket: the sorting network
            if (n <= 512 || MAXR < 32) size = warp_encode_sort<16>(codes, n, K, dst, lane);
            else size = warp_encode_sort<(MAXR < 32 ? 16 : 32)>(codes, n, K, dst, lane);
        }
    } else {
        for (u32 i = lane; i < n; i += 32) {
            u64 key = 0;
            for (int t = 0; t < K; ++t) key = (key << 5) | (u64)(residue_code(r[i + t]) & 31);
            dst[i] = key;
        }
        __syncwarp();
        bitonic_sort_warp_mem(dst, n, lane);
        // compact the distinct keys in place, 32 at a time (an output index never exceeds its input index)
        u32 out = 0; u64 last = SENT;
        for (u32 base = 0; base < n; base += 32) {
            const u32 i = base + lane;
            const bool valid = i < n;
            const u64 v = valid ? dst[i] : SENT;
            u64 pv = shfl_up_u64(v, 1);
            if (lane == 0) pv = last;
            const bool flag = valid && (i == 0 || v != pv);
            const u32 bal = __ballot_sync(0xffffffffu, flag);
            last = __shfl_sync(0xffffffffu, (u32)v, 31) | ((u64)__shfl_sync(0xffffffffu, (u32)(v >> 32), 31) << 32);   // lane 31's value
            __syncwarp();                                       // every read of this group is done before the writes
            if (flag) dst[out + __popc(bal & ((1u << lane) - 1))] = v;
            out += __popc(bal);
            __syncwarp();
        }
        size = out;
    }
    if (lane == 0) { SeqMeta m; m.size = size; m.len = L; m.hash = h; m.res = r; p.meta[s] = m; }
}

}  // namespace kr
