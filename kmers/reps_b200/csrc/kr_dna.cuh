// kr_dna.cuh -- EXPERIMENTAL (compiled only with -DKR_EXPERIMENTAL; checked on the CPU emulator, unverified on
// hardware when written): the DNA k-mer variant.
//
// org.theseed.sequence.DnaKmers (un-vendored module; driven from SsuRepGenomeProcessor.java:88, SeqCompProcessor.java:102,
// ProteinDataFactory.java:277-300, TargetProcessor.java:162-193) is the same set arithmetic on DNA: the set of the
// length-K substrings of the lower-cased sequence AND of its reverse complement; similarity / distance / INFINITY as in
// SequenceKmers.  Only the encoder differs, so this file holds just that: residues -> 3-bit codes (a c g t n u; any
// other character fails the call, so the packing stays injective) -> the keys of both strands -> sorted distinct keys.
// A DNA K-mer needs 3K bits; the context stores keys as ceil(3K / 5) five-bit units, so block lists, the tile join, the
// index and the posting-list path take DNA sets unchanged (K <= 18: 54 bits + the 8-bit block-local sequence number).
// The rules that could not be checked against a JVM are stated where they are applied ([recalled]).
#pragma once
#include "kr_kernels.cuh"

namespace kr {

// [recalled] DnaKmers lower-cases its input; the library's own copy of the residues is normalised in place so that the
// equal-sequence rule (hash + byte compare) sees what the Java code compares
__global__ void k_dna_lowercase(u8 *res, u64 n) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (u64)gridDim.x * blockDim.x;
    for (u64 i = tid; i < n; i += nthr) { const u8 c = res[i]; if (c >= 'A' && c <= 'Z') res[i] = c + 32; }
}

__device__ __forceinline__ u32 dna_code(u32 c) {   // lower-case input
    return c == 'a' ? 0u : c == 'c' ? 1u : c == 'g' ? 2u : c == 't' ? 3u : c == 'n' ? 4u : c == 'u' ? 5u : 0xFFu;
}
// [recalled] Contig.reverse: a <-> t, c <-> g, u -> a, anything else -> n
__device__ __forceinline__ u32 dna_comp(u32 code) { return code < 4 ? 3 - code : (code == 5 ? 0u : 4u); }

// One CTA per sequence, same structure as k_encode_sort: 2 (L - K + 1) keys, sorted, distinct.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_encode_dna(EncodeParams p) {
    KR_DYN_SMEM(u64, sk, 16);                          // smem_cap keys, then smem_cap + 16 code bytes
    __shared__ u64 s_hash, s_prev;
    __shared__ u32 s_wtot[THREADS / 32];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 s = blockIdx.x;
    const u8 *r = p.res + p.off[s];
    const u32 L = (u32)(p.off[s + 1] - p.off[s]);
    const int K = p.K;
    const u32 n1 = L >= (u32)K ? L - K + 1 : 0, n = 2 * n1;
    u64 *dst = p.keys + p.koff[s];
    const bool small = n <= p.smem_cap && L <= p.smem_cap + 16;
    u64 *w = small ? sk : dst;
    u8 *codes = reinterpret_cast<u8 *>(sk + p.smem_cap);
    if (tid == 0) s_hash = 0;
    __syncthreads();
    u64 h = 0;
    for (u32 i = tid; i < L; i += THREADS) {
        const u32 c = r[i];
        const u32 code = dna_code(c);
        if (code == 0xFFu && atomicCAS(&p.err->flag, 0u, 1u) == 0u) { p.err->seq = s; p.err->pos = i; p.err->byte = c; }
        if (small) codes[i] = (u8)code;
        h += mix64((((u64)i << 8) | c) + 0x9E3779B97F4A7C15ULL);
    }
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&s_hash), (unsigned long long)h);
    __syncthreads();
    for (u32 i = tid; i < n1; i += THREADS) {
        u64 fwd = 0, rev = 0;                           // window i of the forward strand and its reverse complement
        for (int t = 0; t < K; ++t) {
            const u32 cf = (small ? codes[i + t] : dna_code(r[i + t])) & 7;
            const u32 cr = (small ? codes[i + K - 1 - t] : dna_code(r[i + K - 1 - t])) & 7;
            fwd = (fwd << 3) | cf;
            rev = (rev << 3) | dna_comp(cr);
        }
        w[i] = fwd; w[n1 + i] = rev;
    }
    __syncthreads();
    bitonic_sort(w, n, tid, THREADS);
    u32 out_base = 0;
    for (u32 tile = 0; tile < n; tile += THREADS) {
        const u32 i = tile + tid;
        const bool valid = i < n;
        const u64 v = valid ? w[i] : 0;
        const bool flag = valid && (i == 0 || v != (tid == 0 ? s_prev : w[i - 1]));
        const u32 bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_wtot[warp] = __popc(bal);
        __syncthreads();
        u32 before = 0, total = 0;
        for (u32 k = 0; k < THREADS / 32; ++k) { const u32 c = s_wtot[k]; if (k < warp) before += c; total += c; }
        if (flag) dst[out_base + before + __popc(bal & ((1u << lane) - 1))] = v;
        if (valid && (i == n - 1 || tid == THREADS - 1)) s_prev = v;
        out_base += total;
        __syncthreads();
    }
    if (tid == 0) { SeqMeta m; m.size = out_base; m.len = L; m.hash = s_hash; m.res = r; p.meta[s] = m; }
}

}  // namespace kr
