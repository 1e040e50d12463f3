// kr_kernels.cuh -- sm_100a kernels of the protein k-mer similarity path.
//
// Data model (replaces the HashSet<String> of ProteinKmers):
//   residue  -> 5-bit code; k-mer -> 5K-bit key in a uint64 (K = 1..11)
//   sequence -> sorted array of its DISTINCT keys            ("key list",   K1)
//   block    -> 256 consecutive sequences; all their (key,lid) entries, key<<8|lid, sorted and
//               cut into NCHUNK key-range chunks by a fixed prefix table ("block list", K2);
//               stored split: hi = top 32 bits of the entry (searched), lo = the remaining low
//               bits (rest of the key + lid; 16 bits for K <= 8, 32 bits for K = 9..11)
//   tile     -> block I x block J: one merge-style pass over the two block lists, staged chunk by
//               chunk into shared memory with 1-D TMA bulk copies, counts the shared keys of all
//               256x256 sequence pairs at once in a shared-memory count tile                (K3)
// Every similarity is an exact integer; nothing here is floating point.
#pragma once
#ifdef KR_EMUL
#include "cuda_emul.h"   // tests/emul: the same sources compiled with g++ and run on CPU threads (test infrastructure)
#define KR_DYN_SMEM(type, name, align) type *name = reinterpret_cast<type *>(emu::dyn_smem())
#else
#include <cuda_runtime.h>
#define KR_DYN_SMEM(type, name, align) extern __shared__ __align__(align) type name[]
#endif
#include <stdint.h>

#include "kr_types.cuh"

namespace kr {


constexpr int TB = 256;          // sequences per block (KR_TILE); lid fits LID_BITS
constexpr int LID_BITS = 8;
constexpr u64 LID_MASK = 0xFFull;
#ifndef KR_NCHUNK
#define KR_NCHUNK 64
#endif
constexpr int NCHUNK = KR_NCHUNK;  // key-range chunks per block list (32 or 64)
constexpr int CHUNK_ALIGN = 16;  // chunk starts are multiples of 16 entries (128 B)
#ifndef KR_STAGE_CAP
#define KR_STAGE_CAP 2040
#endif
#ifndef KR_NSTAGE
#define KR_NSTAGE 4
#endif
#ifndef KR_NCWARP
#define KR_NCWARP 24
#endif
#ifndef KR_ILP
#define KR_ILP 4
#endif
constexpr int STAGE_CAP = KR_STAGE_CAP;  // entries per side per job (a multiple of 8 so sub-chunk starts stay 16-byte aligned)
constexpr int STAGE_BUF = (STAGE_CAP + 7) / 8 * 8;   // entries reserved per side in a stage
static_assert(STAGE_CAP % 8 == 0, "sub-chunk starts must stay 16-byte aligned for the TMA copies");
constexpr int NSTAGE = KR_NSTAGE;
#ifndef KR_STEP_TOP
#define KR_STEP_TOP 1024
#endif
constexpr uint32_t STEP_TOP = KR_STEP_TOP;     // first probe step of the lower bound
constexpr uint32_t WINDOW = 2 * STEP_TOP;      // hi entries staged per B side: probes reach index WINDOW-2
constexpr uint32_t TAILPAD = WINDOW;           // sentinel hi entries after every block list
constexpr uint32_t HI_SENT = 0xFFFFFFFFu;      // above every real hi (top residue code <= 27)
static_assert(STAGE_CAP <= WINDOW - 1, "a staged B sub-chunk must fit the probe window");
constexpr int NCWARP = KR_NCWARP;  // consumer warps of the join kernel (+1 producer warp)
constexpr int ILP = KR_ILP;        // A entries searched together per lane
constexpr int JOIN_THREADS = (NCWARP + 1) * 32;
#ifndef KR_NGROUP
#define KR_NGROUP 2
#endif
constexpr int NGROUP = KR_NGROUP;  // consumer warp groups; the jobs of a tile are dealt to the groups round-robin
constexpr int GW = NCWARP / NGROUP;
static_assert(NCWARP % NGROUP == 0, "consumer warps must split evenly into groups");
constexpr u64 SENT = ~0ull;
constexpr int LUT_BITS = 15;     // chunk table is indexed by the first three residues of the key

constexpr size_t JOIN_SMEM_CNT = (size_t)TB * TB * 2;                         // u16 count tile
// one pipeline stage: A hi | B hi window | A lo | B lo
__host__ __device__ constexpr size_t stage_bytes(bool lo32) { return (size_t)STAGE_BUF * 4 + WINDOW * 4 + 2 * (size_t)STAGE_BUF * (lo32 ? 4 : 2); }
__host__ __device__ constexpr int nstage(bool lo32) { return lo32 ? NSTAGE - 1 : NSTAGE; }
__host__ __device__ constexpr size_t join_smem_bytes(bool lo32) { return JOIN_SMEM_CNT + nstage(lo32) * stage_bytes(lo32) + 256; }
constexpr size_t JOIN_SMEM_BYTES = join_smem_bytes(false) > join_smem_bytes(true) ? join_smem_bytes(false) : join_smem_bytes(true);
static_assert(JOIN_SMEM_BYTES <= 232448, "join kernel shared memory exceeds 227 KB");
// low bits of an entry kept in `lo`: the lid plus whatever of the key does not fit the 32-bit hi
__host__ __device__ constexpr int lo_bits(int K) { return 5 * K + LID_BITS - 32 > LID_BITS ? 5 * K + LID_BITS - 32 : LID_BITS; }


struct BlockRef {    // one per block of 256 sequences
    const u32 *hi;   // block list, high words (sorted; sentinel-padded between chunks and after the block)
    const void *lo;  // block list, low bits (u16 or u32 per entry, same indexing as hi)
    const u32 *off;  // NCHUNK+1 chunk starts (entries, relative to hi/lo, multiples of CHUNK_ALIGN)
    const u32 *cnt;  // NCHUNK chunk lengths
    u32 seq0;        // index of the block's first sequence in its set / database
    u32 nseq;        // sequences in the block (<= 256)
};

struct ErrInfo {       // device-side error record of a context, read back at the next host synchronisation
    u32 flag, seq, pos, byte;   // K1: first residue outside the alphabet
    u32 join_fault;             // K3: JOIN_FAULT_* (the launch finished early or skipped work; results are invalid)
    u32 join_info[3];
};
constexpr u32 JOIN_FAULT_WATCHDOG = 1;   // an mbarrier wait did not complete (protocol bug / lost copy): CTA drained
constexpr u32 JOIN_FAULT_ALIGN = 2;      // a block list handed to the kernel is not 16-byte aligned: job skipped, no copy issued


// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31; return x;
}

__device__ __forceinline__ void cmpswap(u64 *d, u32 i, u32 j) {
    u64 a = d[i], b = d[j];
    if (b < a) { d[i] = b; d[j] = a; }
}

// Bitonic sorting network in its "always ascending" form (flip step + disperse steps), so indices
// >= n behave as +infinity without being stored: any n, no padding.  Whole CTA must call.
__device__ void bitonic_sort(u64 *d, u32 n, u32 tid, u32 nthr) {
    if (n < 2) return;
    u32 lp = 32 - __clz(n - 1);           // P = 1 << lp >= n
    u32 halfP = 1u << (lp - 1);
    for (u32 lk = 1; lk <= lp; ++lk) {    // k = 1 << lk
        u32 lh = lk - 1;                  // half = k/2
        for (u32 t = tid; t < halfP; t += nthr) {
            u32 g = t >> lh, r = t & ((1u << lh) - 1);
            u32 i = (g << lk) + r, j = (g << lk) + ((1u << lk) - 1 - r);
            if (j < n) cmpswap(d, i, j);
        }
        __syncthreads();
        for (int lj = (int)lh - 1; lj >= 0; --lj) {
            for (u32 t = tid; t < halfP; t += nthr) {
                u32 i = ((t >> lj) << (lj + 1)) + (t & ((1u << lj) - 1));
                u32 j = i + (1u << lj);
                if (j < n) cmpswap(d, i, j);
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1 (kr_encode.cuh): one warp per sequence: residues -> codes -> keys -> sort -> distinct keys + size + hash.
struct EncodeParams {
    const u8 *res; const u64 *off; const u64 *koff;
    u64 *keys; SeqMeta *meta; ErrInfo *err;
    u32 nseq; int K; u32 seq_base;     // sequences [0, nseq) of this launch are sequences seq_base + s of the batch (error reports)
};

// residue byte -> 5-bit code without a table: 'A'..'Z' -> 0..25, '*' -> 26, '-' -> 27, anything else 0xFF
__device__ __forceinline__ u32 residue_code(u32 c) {
    const u32 letter = c - 'A';
    return letter < 26u ? letter : (c == '*' ? 26u : (c == '-' ? 27u : 0xFFu));
}

// copy the key lists of selected sequences into a new compact arena (kr_seqset_gather)
__global__ void k_gather_keys(const u64 *src_keys, const u64 *src_koff, const SeqMeta *src_meta, const u32 *idx,
                              u32 n, const u64 *dst_koff, u64 *dst_keys, SeqMeta *dst_meta) {
    u32 i = blockIdx.x;
    if (i >= n) return;
    u32 s = idx[i];
    SeqMeta m = src_meta[s];
    const u64 *a = src_keys + src_koff[s];
    u64 *d = dst_keys + dst_koff[i];
    for (u32 t = threadIdx.x; t < m.size; t += blockDim.x) d[t] = a[t];
    if (threadIdx.x == 0) dst_meta[i] = m;
}

// ---------------------------------------------------------------------------------------------
// K2: block lists.  count -> offsets -> scatter -> sort each chunk.
// chunk of a key: the table is indexed by the key's first three residues (15 bits); keys of K < 3 are shorter than
// that and are left-aligned into the index instead (shift < 0)
__device__ __forceinline__ u32 chunk_prefix(u64 key, int shift) { return (u32)(shift >= 0 ? key >> shift : key << -shift); }
__global__ void __launch_bounds__(256) k_bl_count(const u64 *keys, const u64 *koff, const SeqMeta *meta, u32 nseq,
                                                  const u8 *lut, int shift, u32 *chcnt) {
    __shared__ u32 hist[NCHUNK];
    const u32 b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < NCHUNK) hist[tid] = 0;
    __syncthreads();
    u32 s0 = b * TB, s1 = min(nseq, s0 + TB);
    for (u32 s = s0 + warp; s < s1; s += 8) {
        u32 sz = meta[s].size; const u64 *kp = keys + koff[s];
        for (u32 i = lane; i < sz; i += 32) atomicAdd(&hist[lut[chunk_prefix(kp[i], shift)]], 1u);
    }
    __syncthreads();
    if (tid < NCHUNK) chcnt[b * NCHUNK + tid] = hist[tid];
}

// one CTA: per-block padded chunk starts, then an exclusive scan of the block totals
__global__ void __launch_bounds__(1024) k_bl_offsets(const u32 *chcnt, u32 nblk, u32 *choff, u64 *blkbase) {
    __shared__ u64 part[1024];
    const u32 tid = threadIdx.x;
    u32 per = (nblk + 1023) / 1024;
    u32 b0 = min(nblk, tid * per), b1 = min(nblk, b0 + per);
    u64 sum = 0;
    for (u32 b = b0; b < b1; ++b) {
        u32 o = 0;
        for (int c = 0; c < NCHUNK; ++c) {
            choff[b * (NCHUNK + 1) + c] = o;
            o += (chcnt[b * NCHUNK + c] + CHUNK_ALIGN - 1) / CHUNK_ALIGN * CHUNK_ALIGN;
        }
        choff[b * (NCHUNK + 1) + NCHUNK] = o;
        sum += o + TAILPAD;
    }
    part[tid] = sum;
    __syncthreads();
    if (tid == 0) { u64 run = 0; for (int i = 0; i < 1024; ++i) { u64 v = part[i]; part[i] = run; run += v; } blkbase[nblk] = run; }
    __syncthreads();
    u64 run = part[tid];
    for (u32 b = b0; b < b1; ++b) { blkbase[b] = run; run += choff[b * (NCHUNK + 1) + NCHUNK] + TAILPAD; }
}

__global__ void __launch_bounds__(256) k_bl_scatter(const u64 *keys, const u64 *koff, const SeqMeta *meta, u32 nseq,
                                                    const u8 *lut, int shift, const u32 *chcnt, const u32 *choff,
                                                    const u64 *blkbase, u64 *ent, u32 *hi, void *lo, int lo_size, BlockRef *refs) {
    __shared__ u32 cur[NCHUNK];
    __shared__ u32 so[NCHUNK + 1];
    const u32 b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < NCHUNK) cur[tid] = 0;
    if (tid <= NCHUNK) so[tid] = choff[b * (NCHUNK + 1) + tid];
    __syncthreads();
    u64 *e = ent + blkbase[b];
    u32 s0 = b * TB, s1 = min(nseq, s0 + TB);
    for (u32 s = s0 + warp; s < s1; s += 8) {
        u32 sz = meta[s].size; const u64 *kp = keys + koff[s];
        u64 lid = s - s0;
        for (u32 i = lane; i < sz; i += 32) {
            u64 key = kp[i];
            u32 c = lut[chunk_prefix(key, shift)];
            u32 pos = atomicAdd(&cur[c], 1u);
            e[so[c] + pos] = (key << LID_BITS) | lid;
        }
    }
    __syncthreads();
    // sentinel hi words between the end of each chunk and the start of the next, and after the block
    u32 *h = hi + blkbase[b];
    for (u32 c = warp; c < NCHUNK; c += 8)
        for (u32 i = so[c] + chcnt[b * NCHUNK + c] + lane; i < so[c + 1]; i += 32) h[i] = HI_SENT;
    for (u32 i = so[NCHUNK] + tid; i < so[NCHUNK] + TAILPAD; i += 256) h[i] = HI_SENT;
    if (tid == 0) {
        BlockRef r; r.hi = h; r.lo = (const u8 *)lo + blkbase[b] * lo_size;
        r.off = choff + b * (NCHUNK + 1); r.cnt = chcnt + b * NCHUNK; r.seq0 = s0; r.nseq = s1 - s0;
        refs[b] = r;
    }
}

// sort every chunk by the full entry, then store it split into hi / lo
constexpr int SORT_SMEM_CAP = 4096;
__global__ void __launch_bounds__(256) k_bl_sort(const BlockRef *refs, const u64 *blkbase, u64 *ent, int lobits) {
    __shared__ __align__(16) u64 sk[SORT_SMEM_CAP];
    const u32 b = blockIdx.x / NCHUNK, c = blockIdx.x % NCHUNK, tid = threadIdx.x;
    BlockRef r = refs[b];
    u32 n = r.cnt[c];
    if (n == 0) return;
    u64 *g = ent + blkbase[b] + r.off[c];
    u64 *src = g;
    if (n <= SORT_SMEM_CAP) {
        for (u32 i = tid; i < n; i += 256) sk[i] = g[i];
        __syncthreads();
        bitonic_sort(sk, n, tid, 256);
        src = sk;
    } else {
        bitonic_sort(g, n, tid, 256);   // rare: very long sequences; in place in global memory
    }
    u32 *h = const_cast<u32 *>(r.hi) + r.off[c];
    const u64 lomask = (1ull << lobits) - 1;
    if (lobits <= 16) {
        u16 *l = reinterpret_cast<u16 *>(const_cast<void *>(r.lo)) + r.off[c];
        for (u32 i = tid; i < n; i += 256) { u64 e = src[i]; h[i] = (u32)(e >> lobits); l[i] = (u16)(e & lomask); }
    } else {
        u32 *l = reinterpret_cast<u32 *>(const_cast<void *>(r.lo)) + r.off[c];
        for (u32 i = tid; i < n; i += 256) { u64 e = src[i]; h[i] = (u32)(e >> lobits); l[i] = (u32)(e & lomask); }
    }
}

// import of exchanged block lists: block refs of one shard inside the concatenated arrays, residue pointers
__global__ void k_import_refs(u32 nblk, u32 blk0, u32 seq0, u32 nseq_shard, const u32 *hi, const u8 *lo, u32 lo_size,
                              const u32 *choff, const u32 *chcnt, const u64 *blkbase_local, u64 ent0, BlockRef *refs) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    BlockRef r;
    const u64 base = ent0 + blkbase_local[b];
    r.hi = hi + base; r.lo = lo + base * lo_size;
    r.off = choff + (size_t)(blk0 + b) * (NCHUNK + 1); r.cnt = chcnt + (size_t)(blk0 + b) * NCHUNK;
    r.seq0 = seq0 + b * TB; r.nseq = min((u32)TB, nseq_shard - b * TB);
    refs[blk0 + b] = r;
}
// layout rules of an exchanged shard, checked before anything is built from it (a violation would hand the tile join a
// misaligned or out-of-range TMA source): block starts and chunk starts are multiples of CHUNK_ALIGN entries, chunk
// starts ascend, every chunk fits its slot, and every block leaves TAILPAD entries inside the shard's arrays.
// bad[0] = number of violating blocks, bad[1] = first one seen
__global__ void k_validate_shard(u32 nblk, u64 entries, const u32 *choff, const u32 *chcnt, const u64 *blkbase, u32 *bad) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const u64 base = blkbase[b];
    const u32 *o = choff + (size_t)b * (NCHUNK + 1), *c = chcnt + (size_t)b * NCHUNK;
    bool ok = (base % CHUNK_ALIGN) == 0 && o[0] == 0;
    for (int k = 0; k < NCHUNK && ok; ++k) ok = (o[k] % CHUNK_ALIGN) == 0 && o[k + 1] >= o[k] && c[k] <= o[k + 1] - o[k];
    ok = ok && (o[NCHUNK] % CHUNK_ALIGN) == 0 && base + o[NCHUNK] + TAILPAD <= entries && blkbase[b + 1] >= base + o[NCHUNK] + TAILPAD;
    if (!ok && atomicAdd(&bad[0], 1u) == 0u) bad[1] = b;
}
__global__ void k_import_meta(SeqMeta *meta, u32 n, const u8 *res, const u64 *off) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) meta[i].res = res + off[i];
}

// re-base the block refs of a set into a database-wide array (seq0 += base)
__global__ void k_copy_refs(const BlockRef *src, u32 n, u32 seq_base, BlockRef *dst) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { BlockRef r = src[i]; r.seq0 += seq_base; dst[i] = r; }
}

// ---------------------------------------------------------------------------------------------
#ifdef KR_EMUL
// the same primitives on the CPU emulator (tests/emul/cuda_emul.h)
__device__ __forceinline__ u32 smem_u32(const void *p) { return emu::smem_addr(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) { emu::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) { emu::mbar_expect_tx(bar, bytes); }
__device__ __forceinline__ void mbar_arrive_n(u64 *bar, u32 count) { emu::mbar_arrive(bar, count); }
__device__ __forceinline__ void mbar_arrive(u64 *bar) { emu::mbar_arrive(bar, 1); }
__device__ __forceinline__ bool mbar_wait(u64 *bar, u32 parity, volatile u32 *abort_flag) {
    for (u32 spins = 0; !emu::mbar_try_wait(bar, parity); ++spins)
        if (*abort_flag || spins > (1u << 22)) { *abort_flag = 1; return false; }
    return true;
}
__device__ __forceinline__ bool consumer_bar_or(bool pred) { return emu::named_barrier_or(1, NCWARP * 32, pred); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) { emu::bulk_g2s(dst, src, bytes, bar); }
__device__ __forceinline__ void consumer_bar();
#else
// mbarrier / TMA bulk-copy primitives (sm_90+; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(u64 *bar, u32 count) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait for the phase of the given parity.  try_wait suspends in hardware, so the loop turns slowly.  A wait that does
// not end within WATCHDOG_NS is a protocol bug or a lost copy: it never traps (a trap is a device fault for the whole
// process) -- it raises the CTA's abort flag and returns false; every other wait of the CTA then returns false at
// once, the consumers leave together through consumer_bar_or() and the host reports JOIN_FAULT_WATCHDOG.
constexpr unsigned long long WATCHDOG_NS = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ bool mbar_try(u32 addr, u32 parity) {
    u32 done;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    return done != 0;
}
__device__ __noinline__ bool mbar_wait_slow(u32 addr, u32 parity, volatile u32 *abort_flag) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (u32 spins = 0;; ++spins) {
        if (mbar_try(addr, parity)) return true;
        if ((spins & 63) == 63) {
            if (*abort_flag) return false;
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > WATCHDOG_NS) { *abort_flag = 1; return false; }
        }
    }
}
__device__ __forceinline__ bool mbar_wait(u64 *bar, u32 parity, volatile u32 *abort_flag) {
    const u32 addr = smem_u32(bar);
    if (mbar_try(addr, parity)) return true;
    return mbar_wait_slow(addr, parity, abort_flag);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NCWARP * 32) : "memory"); }
// the same barrier, returning the OR of `pred` over all consumer threads (uniform: every thread gets the same answer)
__device__ __forceinline__ bool consumer_bar_or(bool pred) {
    u32 r;
    asm volatile(
        "{\n"
        ".reg .pred P1, P2;\n"
        "setp.ne.u32 P1, %1, 0;\n"
        "barrier.cta.red.or.pred P2, 1, %2, P1;\n"
        "selp.u32 %0, 1, 0, P2;\n"
        "}\n" : "=r"(r) : "r"((u32)pred), "n"(NCWARP * 32) : "memory");
    return r != 0;
}
#endif
#ifdef KR_EMUL
__device__ __forceinline__ void consumer_bar() { emu::named_barrier(1, NCWARP * 32); }
#endif

// ---------------------------------------------------------------------------------------------
// K3: the tile join.
enum { MODE_MATRIX = 0, MODE_ARGMAX = 1, MODE_SPARSE = 2 };

struct JoinParams {
    const BlockRef *A, *B;
    const SeqMeta *metaA, *metaB;
    u32 nA, nB;              // sequences on each side
    u32 nBblk;
    const uint2 *tasks; u32 ntasks;   // MATRIX: (I, J).  ARGMAX: (I, split)
    u32 flags;               // KR_IDENT_IS_INFINITY
    ErrInfo *err;            // the context's device-side error record (join_fault)
    // MATRIX
    void *out; u32 elem, layout; u64 origin;
    // ARGMAX
    u32 nsplit; const u32 *rankB; int4 *partial;
    // SPARSE: cells (row, col, sim, 0) with sim >= min_sim, appended in no particular order
    uint4 *cells; unsigned long long cells_cap; unsigned long long *cells_count; u32 min_sim; u32 cells_packed;
};

// one sparse result cell: 16 bytes (row, col, sim, 0) or, packed, one uint64 = row << 40 | col << 16 | sim (rows and columns
// below 2^24, which the host checks; INFINITY = 9999 and every count <= 65535 fit the low 16 bits)
__device__ __forceinline__ void put_cell(uint4 *cells, u32 packed, unsigned long long pos, u32 row, u32 col, u32 sim) {
    if (packed) reinterpret_cast<u64 *>(cells)[pos] = ((u64)row << 40) | ((u64)col << 16) | (u64)sim;
    else cells[pos] = make_uint4(row, col, sim, 0u);
}

__device__ __forceinline__ u64 cell_index(u32 layout, u64 i, u64 j, u64 nrows, u64 ncols) {
    if (layout == 0) return i * ncols + j;
    if (layout == 1) return i * nrows - i * (i + 1) / 2 + (j - i - 1);
    u64 I = i >> 8, c0 = I << 8;
    return 256ull * (I * nrows - 128ull * I * (I - 1)) + (i - c0) * (nrows - c0) + (j - c0);
}

// byte compare of two stored sequences of the same length (the rare exact step of the equal-sequence rule)
__device__ __noinline__ bool res_equal(const u8 *a, const u8 *b, u32 len) {
    for (u32 i = 0; i < len; ++i) if (a[i] != b[i]) return false;
    return true;
}
// arguments by value: taking SeqMeta by reference put the records on the local stack of every caller
__device__ __forceinline__ bool seq_equal(const SeqMeta a, const SeqMeta b) {
    if (a.len != b.len || a.hash != b.hash) return false;
    if (a.res == b.res) return true;   // the same stored sequence (diagonal cells)
    return res_equal(a.res, b.res, a.len);
}

#ifdef KR_EMUL
__device__ __forceinline__ u32 lds32(u32 addr) { return *reinterpret_cast<const u32 *>(emu::smem_ptr(addr)); }
__device__ __forceinline__ u32 lds16(u32 addr) { return *reinterpret_cast<const u16 *>(emu::smem_ptr(addr)); }
#else
__device__ __forceinline__ u32 lds32(u32 addr) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ u32 lds16(u32 addr) {
    u32 v;
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(addr));
    return v;
}
#endif
template <bool LO32> __device__ __forceinline__ u32 ldlo(u32 base, u32 i) { return LO32 ? lds32(base + i * 4) : lds16(base + i * 2); }

// All matches of the staged A entries against the staged B sub-chunk -> count tile.
//   * a pass is 32*ILP consecutive A entries (ILP per lane, searched together); the passes of a job are dealt
//     round-robin to the GW warps of the group that owns the job, continuing across the group's jobs
//     (pass counter `pc`);
//   * the lower bound runs on the 32-bit hi words only.  The B window always holds WINDOW staged words and
//     everything past the sub-chunk is >= every target of this chunk (later chunks, or sentinels), so the
//     binary lifting needs neither a bounds check nor a clamp; the position is kept as an absolute shared
//     address, so one probe is LDS [p+imm]; ISETP; @P IADD;
//   * entries whose hi equals the target are then checked on the low key bits and counted.
template <bool LO32>
__device__ __forceinline__ void join_job(u32 sAhi, u32 sAlo, u32 na, u32 sBhi, u32 sBlo, u32 nb, u32 *cnt, u32 warp,
                                         u32 lane, u32 pc) {
    const u32 w = (warp + GW - pc % GW) % GW;   // first pass of this job that falls to this warp (warp = index in group)
    const u32 sBm = sBhi - 4;
    for (u32 base = w * (32 * ILP); base < na; base += GW * 32 * ILP) {
        u32 th[ILP], p[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const u32 i = base + lane + 32 * u;
            // lanes past the end search for a value no entry has (real hi words are < 0xE0000000)
            th[u] = i < na ? lds32(sAhi + i * 4) : HI_SENT - 1;
            p[u] = sBm;
        }
#pragma unroll
        for (u32 step = STEP_TOP; step; step >>= 1) {
            u32 v[ILP];
#pragma unroll
            for (int u = 0; u < ILP; ++u) v[u] = lds32(p[u] + step * 4);
#pragma unroll
            for (int u = 0; u < ILP; ++u) if (v[u] < th[u]) p[u] += step * 4;
        }
        // p[u] + 4 is the address of the lower-bound entry: a hit needs equal hi words inside the sub-chunk
        bool hit = false;
        u32 hv[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) hv[u] = lds32(p[u] + 4);
#pragma unroll
        for (int u = 0; u < ILP; ++u) hit |= (hv[u] == th[u]) & (p[u] - sBm < nb * 4);
        if (__any_sync(0xffffffffu, hit)) {   // rare for unrelated sequences: whole warp skips otherwise
#pragma unroll
            for (int u = 0; u < ILP; ++u) {
                u32 idx = (p[u] - sBm) >> 2;
                if (hv[u] == th[u] && idx < nb) {
                    const u32 al = ldlo<LO32>(sAlo, base + lane + 32 * u);
                    const u32 a = al & (u32)LID_MASK;
                    do {
                        const u32 bl = ldlo<LO32>(sBlo, idx);
                        if ((bl >> LID_BITS) == (al >> LID_BITS)) {
                            const u32 b = bl & (u32)LID_MASK;
                            atomicAdd(&cnt[(a * TB + b) >> 1], (b & 1) ? 0x10000u : 1u);
                        }
                        ++idx;
                    } while (idx < nb && lds32(sBhi + idx * 4) == th[u]);
                }
            }
        }
    }
}

// Job flags published by the producer with every pipeline slot
constexpr u32 JF_TASK_END = 2;   // head slot: this tile is the last J of its task
constexpr u32 JF_DONE = 4;       // head slot: no more work for this CTA

// The kernel is warp-specialised.  Warp NCWARP is the producer: it alone walks tasks -> tiles -> chunk pairs ->
// sub-chunks.  Per tile it first publishes a HEAD slot (number of jobs of the tile, J, flags) that every consumer
// warp reads, then one slot per job: wait for a free stage, write a 16-byte descriptor (na, nb) and issue the four
// TMA bulk copies that complete the stage's full barrier.  The NCWARP consumer warps form NGROUP groups; job r of
// a tile belongs to group r % NGROUP, so only GW warps pay the wait/arrive cost of a job and the groups work on
// different jobs at the same time.  All consumers meet at a named barrier twice per tile (epilogue).
template <int MODE, bool LO32>
__global__ void __launch_bounds__(JOIN_THREADS, 1) k_join(JoinParams p) {
    KR_DYN_SMEM(u8, smem, 128);
    constexpr int NST = nstage(LO32);
    constexpr u32 LOSZ = LO32 ? 4 : 2;
    constexpr u32 STAGE_B = (u32)stage_bytes(LO32);
    constexpr u32 OFF_BHI = STAGE_BUF * 4, OFF_ALO = OFF_BHI + WINDOW * 4, OFF_BLO = OFF_ALO + STAGE_BUF * LOSZ;
    u32 *cnt = reinterpret_cast<u32 *>(smem);
    u8 *stage = smem + JOIN_SMEM_CNT;
    u64 *full = reinterpret_cast<u64 *>(smem + JOIN_SMEM_CNT + (size_t)NST * STAGE_B);
    u64 *empty = full + NST;
    uint4 *desc = reinterpret_cast<uint4 *>(empty + NST);   // 2*NST barriers of 8 bytes: 16-byte aligned
    volatile u32 *s_abort = reinterpret_cast<volatile u32 *>(desc + NST);   // raised by a wait that timed out (see mbar_wait)
    static_assert(2 * NST * 8 + NST * 16 + 4 <= 256, "barriers, descriptors and the abort flag live in the last 256 bytes");
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        *s_abort = 0;
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCWARP); }
#ifndef KR_EMUL
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
    }
    for (u32 i = tid; i < TB * TB / 2; i += JOIN_THREADS) cnt[i] = 0;
    __syncthreads();

    if (warp == NCWARP) {
        // ================================ producer ================================================
        u32 k = 0;
        bool alive = true;       // false once a wait of this CTA timed out: nothing more is published
        auto emit = [&](u32 na, u32 nb, u32 flags, u32 J, const BlockRef &a, const BlockRef &b, u32 posA, u32 posB) {
            const u32 s = k % NST, ph = (k / NST) & 1;
            if (lane == 0 && alive) {
                u8 *st = stage + (size_t)s * STAGE_B;
                if (!mbar_wait(&empty[s], ph ^ 1, s_abort)) {
                    alive = false;
                    if (atomicCAS(&p.err->join_fault, 0u, JOIN_FAULT_WATCHDOG) == 0u) { p.err->join_info[0] = blockIdx.x; p.err->join_info[1] = k; }
                } else {
                    const u8 *gAhi = reinterpret_cast<const u8 *>(a.hi + posA), *gBhi = reinterpret_cast<const u8 *>(b.hi + posB);
                    const u8 *gAlo = (const u8 *)a.lo + (size_t)posA * LOSZ, *gBlo = (const u8 *)b.lo + (size_t)posB * LOSZ;
                    // cp.async.bulk needs 16-byte aligned sources; a misaligned source is a device fault, so a block list that
                    // breaks the layout rule (kr_seqset_import_blocklists checks it too) is skipped and reported instead
                    const bool copy = na && !(flags & 1u);
                    const bool aligned = ((reinterpret_cast<uintptr_t>(gAhi) | reinterpret_cast<uintptr_t>(gBhi) |
                                           reinterpret_cast<uintptr_t>(gAlo) | reinterpret_cast<uintptr_t>(gBlo)) & 15) == 0;
                    if (copy && !aligned) {
                        if (atomicCAS(&p.err->join_fault, 0u, JOIN_FAULT_ALIGN) == 0u) { p.err->join_info[0] = blockIdx.x; p.err->join_info[1] = posA; p.err->join_info[2] = posB; }
                        na = 0; nb = 0;
                    }
                    desc[s] = make_uint4(na, nb, flags, J);
#if defined(KR_DEBUG_KNOBS) && (KR_DEBUG_KNOBS & 4)
                    if (p.nA != 0xFFFFFFFFu) { mbar_arrive(&full[s]); } else
#endif
                    if (copy && aligned) {
                        const u32 bAhi = ((na + 3) & ~3u) * 4, bAlo = LO32 ? ((na + 3) & ~3u) * 4 : ((na + 7) & ~7u) * 2;
                        const u32 bBlo = LO32 ? ((nb + 3) & ~3u) * 4 : ((nb + 7) & ~7u) * 2;
                        mbar_expect_tx(&full[s], bAhi + WINDOW * 4 + bAlo + bBlo);
                        bulk_g2s(st, gAhi, bAhi, &full[s]);
                        bulk_g2s(st + OFF_BHI, gBhi, WINDOW * 4, &full[s]);
                        bulk_g2s(st + OFF_ALO, gAlo, bAlo, &full[s]);
                        bulk_g2s(st + OFF_BLO, gBlo, bBlo, &full[s]);
                    } else {
                        mbar_arrive(&full[s]);
                    }
                }
            }
            ++k;
        };
        constexpr u32 HEAD = 1u;   // descriptor flag bit 0: head slot (x = jobs of the tile, w = J)
        static_assert(NCHUNK == 64 || NCHUNK == 32, "chunk tables are held as one or two registers per lane");
        constexpr int HI = NCHUNK > 32 ? 32 : 0;   // offset of the second table register (unused for 32 chunks)
        BlockRef a{}, b{};
        for (u32 task = blockIdx.x; task < p.ntasks; task += gridDim.x) {
            if (!__shfl_sync(0xffffffffu, (u32)alive, 0)) return;   // the CTA is draining after a watchdog timeout
            const uint2 tk = p.tasks[task];
            u32 j0, j1;
            if (MODE != MODE_ARGMAX) { j0 = tk.y; j1 = tk.y + 1; }
            else { j0 = (u32)((u64)tk.y * p.nBblk / p.nsplit); j1 = (u32)((u64)(tk.y + 1) * p.nBblk / p.nsplit); }
            a = p.A[tk.x];
            // chunk tables live in registers (lane c%32 holds chunk c) and are broadcast by shuffle
            const u32 tAc0 = __ldg(a.cnt + lane), tAc1 = __ldg(a.cnt + HI + lane), tAo0 = __ldg(a.off + lane), tAo1 = __ldg(a.off + HI + lane);
            for (u32 J = j0; J < j1; ++J) {
                b = p.B[J];
                const u32 tBc0 = __ldg(b.cnt + lane), tBc1 = __ldg(b.cnt + HI + lane), tBo0 = __ldg(b.off + lane), tBo1 = __ldg(b.off + HI + lane);
                // jobs of this tile: each lane counts the sub-chunk pairs of its one or two chunks
                u32 mine = 0;
                {
                    const u32 x0 = (tAc0 + STAGE_CAP - 1) / STAGE_CAP * ((tBc0 + STAGE_CAP - 1) / STAGE_CAP);
                    const u32 x1 = (tAc1 + STAGE_CAP - 1) / STAGE_CAP * ((tBc1 + STAGE_CAP - 1) / STAGE_CAP);
                    mine = x0 + (NCHUNK > 32 ? x1 : 0);
                }
                for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
                emit(mine, 0, HEAD | (J + 1 == j1 ? JF_TASK_END : 0), J, a, b, 0, 0);
                for (int c = 0; c < NCHUNK; ++c) {
                    const u32 nAc = __shfl_sync(0xffffffffu, c < 32 ? tAc0 : tAc1, c & 31);
                    const u32 nBc = __shfl_sync(0xffffffffu, c < 32 ? tBc0 : tBc1, c & 31);
                    if (nAc == 0 || nBc == 0) continue;
                    const u32 oA = __shfl_sync(0xffffffffu, c < 32 ? tAo0 : tAo1, c & 31);
                    const u32 oB = __shfl_sync(0xffffffffu, c < 32 ? tBo0 : tBo1, c & 31);
                    for (u32 as = 0; as < nAc; as += STAGE_CAP)
                        for (u32 bs = 0; bs < nBc; bs += STAGE_CAP)
                            emit(min((u32)STAGE_CAP, nAc - as), min((u32)STAGE_CAP, nBc - bs), 0, J, a, b, oA + as, oB + bs);
                }
            }
        }
        emit(0, 0, HEAD | JF_DONE, 0, a, b, 0, 0);
        return;
    }

    // ==================================== consumers ===============================================
    u32 k = 0, pc = 0;
    u32 task = blockIdx.x;
    const u32 grp = warp / GW, gw = warp % GW;
    // ARGMAX: row r of the tile is owned by warp r%NCWARP, its running best lives in lane r/NCWARP
    int best_sim = 0; u32 best_first = 0xFFFFFFFFu, best_last = 0;
    bool dead = false;           // a wait of this warp (or of another warp of the CTA) timed out
    for (;;) {
        // head slot of the next tile: read by every consumer warp
        uint4 d = make_uint4(0, 0, 0, 0);
        {
            const u32 s = k % NST, ph = (k / NST) & 1;
            if (mbar_wait(&full[s], ph, s_abort)) {
                d = desc[s];
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            } else dead = true;
            ++k;
        }
        // this group's jobs of the tile
        if (!dead && !(d.z & JF_DONE))
        for (u32 r = grp; r < d.x; r += NGROUP) {
            const u32 kk = k + r, s = kk % NST, ph = (kk / NST) & 1;
            if (!mbar_wait(&full[s], ph, s_abort)) { dead = true; break; }
            const uint4 dj = desc[s];
            const u32 sb = smem_u32(stage + (size_t)s * STAGE_B);
#if defined(KR_DEBUG_KNOBS) && (KR_DEBUG_KNOBS & 1)
            if (p.nA == 0xFFFFFFFFu)
#endif
            join_job<LO32>(sb, sb + OFF_ALO, dj.x, sb + OFF_BHI, sb + OFF_BLO, dj.y, cnt, gw, lane, pc);
            pc += (dj.x + 32 * ILP - 1) / (32 * ILP);
            __syncwarp();
            if (lane == 0) mbar_arrive_n(&empty[s], NGROUP);   // the GW owners stand in for all NCWARP warps
        }
        k += d.x;
        // every consumer warp meets here once per head slot, whatever happened to it above, and all of them take the same
        // decision: leave (end of work, or a timed-out wait anywhere in the CTA) or run the epilogue
        if (consumer_bar_or(dead || (d.z & JF_DONE))) {
            if (dead && lane == 0 && atomicCAS(&p.err->join_fault, 0u, JOIN_FAULT_WATCHDOG) == 0u) { p.err->join_info[0] = blockIdx.x; p.err->join_info[1] = k; }
            break;
        }

        // ---- epilogue of tile (I,J) ----------------------------------------------------------
        const BlockRef a = p.A[p.tasks[task].x];
        const BlockRef b = p.B[d.w];
        const bool ident_inf = (p.flags & 1u) != 0;
        // sizes of this warp's rows, one per lane (row warp + NCWARP*lane), fetched together
        u32 my_szA = 0;
        if (ident_inf && warp + NCWARP * lane < a.nseq) my_szA = p.metaA[a.seq0 + warp + NCWARP * lane].size;
        for (u32 row = warp; row < a.nseq; row += NCWARP) {
            uint4 *rp = reinterpret_cast<uint4 *>(cnt + row * (TB / 2)) + lane;
            uint4 v = *rp;
            *rp = make_uint4(0, 0, 0, 0);
            int cv[8] = {(int)(v.x & 0xFFFF), (int)(v.x >> 16), (int)(v.y & 0xFFFF), (int)(v.y >> 16),
                         (int)(v.z & 0xFFFF), (int)(v.z >> 16), (int)(v.w & 0xFFFF), (int)(v.w >> 16)};
            const u32 i = a.seq0 + row;
            const u32 jb = b.seq0 + lane * 8;      // first column of this lane
            if (ident_inf) {
                const u32 szA = __shfl_sync(0xffffffffu, my_szA, row / NCWARP);
                bool cand = false;
#pragma unroll
                for (int t = 0; t < 8; ++t) cand |= (cv[t] == (int)szA) && (lane * 8 + t < b.nseq);
                if (cand) {
                    const SeqMeta ma = p.metaA[i];
                    for (int t = 0; t < 8; ++t)
                        if (cv[t] == (int)szA && lane * 8 + t < b.nseq) {
                            const SeqMeta mb = p.metaB[jb + t];
                            if (mb.size == szA && seq_equal(ma, mb)) cv[t] = KR_INF;
                        }
                }
            }
#if defined(KR_DEBUG_KNOBS) && (KR_DEBUG_KNOBS & 2)
            if (p.nA == 0xFFFFFFFFu)
#endif
            if (MODE == MODE_MATRIX) {
                const u64 N = p.nA, M = p.nB;
                // valid columns of this row inside the tile
                u32 jlo = b.seq0, jhi = b.seq0 + b.nseq;
                if (p.layout == 1 && jlo < i + 1) jlo = i + 1;
                if (jlo < jhi && jb + 8 > jlo && jb < jhi) {
                    const u32 ja = max(jb, jlo);
                    const u64 idx = cell_index(p.layout, i, ja, N, M) - p.origin - (ja - jb);  // index of column jb
                    const bool fullv = jb >= jlo && jb + 8 <= jhi;
                    if (p.elem == 1) {
                        u16 *o = reinterpret_cast<u16 *>(p.out) + idx;
                        if (fullv && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                            uint4 w;
                            w.x = (u32)cv[0] | ((u32)cv[1] << 16); w.y = (u32)cv[2] | ((u32)cv[3] << 16);
                            w.z = (u32)cv[4] | ((u32)cv[5] << 16); w.w = (u32)cv[6] | ((u32)cv[7] << 16);
                            *reinterpret_cast<uint4 *>(o) = w;
                        } else {
#pragma unroll
                            for (int t = 0; t < 8; ++t) if (jb + t >= jlo && jb + t < jhi) o[t] = (u16)cv[t];
                        }
                    } else {
                        int *o = reinterpret_cast<int *>(p.out) + idx;
                        if (fullv && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                            reinterpret_cast<int4 *>(o)[0] = make_int4(cv[0], cv[1], cv[2], cv[3]);
                            reinterpret_cast<int4 *>(o)[1] = make_int4(cv[4], cv[5], cv[6], cv[7]);
                        } else {
#pragma unroll
                            for (int t = 0; t < 8; ++t) if (jb + t >= jlo && jb + t < jhi) o[t] = cv[t];
                        }
                    }
                }
            }
            else if (MODE == MODE_SPARSE) {
                // only the cells that reach min_sim leave the SM: one atomicAdd per (row, tile) that has any
                u32 jlo = b.seq0;
                const u32 jhi = b.seq0 + b.nseq;
                if (p.layout == 1 && jlo < i + 1) jlo = i + 1;
                u32 m = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) if (jb + t >= jlo && jb + t < jhi && cv[t] >= (int)p.min_sim) m |= 1u << t;
                if (__any_sync(0xffffffffu, m != 0)) {
                    const u32 c = __popc(m);
                    u32 incl = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const u32 up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += up; }
                    const u32 total = __shfl_sync(0xffffffffu, incl, 31);
                    unsigned long long base = 0;
                    if (lane == 31) base = atomicAdd(p.cells_count, (unsigned long long)total);
                    base = __shfl_sync(0xffffffffu, base, 31);
                    unsigned long long pos = base + incl - c;
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        if ((m >> t) & 1u) { if (pos < p.cells_cap) put_cell(p.cells, p.cells_packed, pos, i, jb + t, (u32)cv[t]); ++pos; }
                }
            }
            else {
                // row maximum with first/last rank among the maxima
                int m = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) if (lane * 8 + t < b.nseq) m = max(m, cv[t]);
                int wm = m;
                for (int o = 16; o > 0; o >>= 1) wm = max(wm, __shfl_xor_sync(0xffffffffu, wm, o));
                if (wm > 0) {
                    u32 rf = 0xFFFFFFFFu, rl = 0;
                    if (m == wm) {
#pragma unroll
                        for (int t = 0; t < 8; ++t)
                            if (lane * 8 + t < b.nseq && cv[t] == wm) { u32 rk = __ldg(p.rankB + jb + t); rf = min(rf, rk); rl = max(rl, rk); }
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        rf = min(rf, __shfl_xor_sync(0xffffffffu, rf, o));
                        rl = max(rl, __shfl_xor_sync(0xffffffffu, rl, o));
                    }
                    if (lane == row / NCWARP) {
                        if (wm > best_sim) { best_sim = wm; best_first = rf; best_last = rl; }
                        else if (wm == best_sim) { best_first = min(best_first, rf); best_last = max(best_last, rl); }
                    }
                }
            }
        }
        consumer_bar();
        if (d.z & JF_TASK_END) {
            if (MODE == MODE_ARGMAX) {
                const u32 row = warp + NCWARP * lane;
                if (row < a.nseq) p.partial[(size_t)task * TB + row] = make_int4(best_sim, (int)best_first, (int)best_last, 0);
                best_sim = 0; best_first = 0xFFFFFFFFu; best_last = 0;
            }
            task += gridDim.x;
        }
    }
}

// merge the per-split partial maxima of each query; rank -> representative index via `order`
__global__ void k_argmax_merge(const int4 *partial, u32 nq, u32 nsplit, const u32 *order, int *best_sim, u32 *best_first,
                               u32 *best_last) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    u32 I = q / TB, row = q % TB;
    int bs = 0; u32 bf = 0xFFFFFFFFu, bl = 0;
    for (u32 sp = 0; sp < nsplit; ++sp) {
        int4 v = partial[((size_t)I * nsplit + sp) * TB + row];
        if (v.x > bs) { bs = v.x; bf = (u32)v.y; bl = (u32)v.z; }
        else if (v.x == bs && v.x > 0) { bf = min(bf, (u32)v.y); bl = max(bl, (u32)v.z); }
    }
    best_sim[q] = bs;
    if (best_first) best_first[q] = bs > 0 ? order[bf] : 0xFFFFFFFFu;
    if (best_last) best_last[q] = bs > 0 ? order[bl] : 0xFFFFFFFFu;
}

// ---------------------------------------------------------------------------------------------
// SequenceKmers.similarity for an explicit list of pairs: one warp per pair, the shorter key
// list probes the longer one by binary search (both lists are L2-resident, sorted, distinct).
__global__ void __launch_bounds__(256) k_sim_pairs(const u64 *keysA, const u64 *koffA, const SeqMeta *metaA, const u32 *ia,
                                                   const u64 *keysB, const u64 *koffB, const SeqMeta *metaB, const u32 *ib,
                                                   u64 npairs, u32 flags, int *sim, u8 *ident) {
    const u32 lane = threadIdx.x & 31;
    u64 wid = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 p = wid; p < npairs; p += nw) {
        u32 a = ia[p], b = ib[p];
        SeqMeta ma = metaA[a], mb = metaB[b];
        const u64 *la = keysA + koffA[a], *lb = keysB + koffB[b];
        u32 na = ma.size, nb = mb.size;
        if (na > nb) { const u64 *t = la; la = lb; lb = t; u32 tn = na; na = nb; nb = tn; }
        int c = 0;
        for (u32 i = lane; i < na; i += 32) {
            u64 key = la[i];
            u32 lo = 0, len = nb;
            while (len) { u32 half = len >> 1; if (lb[lo + half] < key) { lo += half + 1; len -= half + 1; } else len = half; }
            c += (lo < nb && lb[lo] == key);
        }
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) {
            bool same = (c == (int)ma.size && c == (int)mb.size) && seq_equal(ma, mb);
            if (ident) ident[p] = same;
            sim[p] = (same && (flags & 1u)) ? KR_INF : c;
        }
    }
}

// records of a set built from explicit key lists (kr_seqset_from_keys): one warp per sequence.  With the sequences at
// hand the record is K1's (length, the same 64-bit hash, residue pointer), so such an object is "equal" to whatever its
// sequence equals; without them every object gets a hash of its own and equals only itself.
__global__ void __launch_bounds__(256) k_meta_from_keys(const u64 *koff, const u8 *res, const u64 *off, u32 nseq, u64 nonce, SeqMeta *meta) {
    const u32 lane = threadIdx.x & 31;
    const u32 s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nseq) return;                                    // whole warps leave together
    SeqMeta m; m.size = (u32)(koff[s + 1] - koff[s]); m.len = 0; m.res = nullptr; m.hash = 0;
    if (res) {
        const u8 *r = res + off[s];
        const u32 L = (u32)(off[s + 1] - off[s]);
        u64 h = 0;
        for (u32 i = lane; i < L; i += 32) h += mix64((((u64)i << 8) | r[i]) + 0x9E3779B97F4A7C15ULL);
        for (int o = 16; o > 0; o >>= 1) {
            const u32 lo = __shfl_xor_sync(0xffffffffu, (u32)h, o), hi = __shfl_xor_sync(0xffffffffu, (u32)(h >> 32), o);
            h += ((u64)hi << 32) | lo;
        }
        m.len = L; m.hash = h; m.res = r;
    } else {
        m.hash = mix64(nonce + s) | 1ull;
    }
    if (lane == 0) meta[s] = m;
}

__global__ void k_fill_u32_iota(u32 *p, u32 n, u32 base) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = base + i;
}

}  // namespace kr

// =============================================================================================
// Helpers of the index scan (kr_index.cuh) that work on per-query partial results.
namespace kr {

// queries with no k-mer at all (shorter than K): under the equal-sequence rule they still match an identical
// representative.  One warp per such query scans the (few) representatives that are also empty.
__global__ void k_idx_empty_queries(const SeqMeta *qmeta, u32 nq, const SeqMeta *rmeta, const u32 *rank, const u32 *empty_reps,
                                    const u32 *nempty_ptr, int4 *partial) {
    const u32 q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    const SeqMeta qm = qmeta[q];
    if (qm.size != 0) return;
    const u32 nempty = *nempty_ptr;
    u32 rf = 0xFFFFFFFFu, rl = 0; bool any = false;
    for (u32 e = lane; e < nempty; e += 32) {
        const u32 g = empty_reps[e];
        if (seq_equal(qm, rmeta[g])) { any = true; rf = min(rf, rank[g]); rl = max(rl, rank[g]); }
    }
    any = __any_sync(0xffffffffu, any);
    for (int o = 16; o > 0; o >>= 1) { rf = min(rf, __shfl_xor_sync(0xffffffffu, rf, o)); rl = max(rl, __shfl_xor_sync(0xffffffffu, rl, o)); }
    if (lane == 0 && any) partial[q] = make_int4(KR_INF, (int)rf, (int)rl, 0);
}

// merge per-level partials (partial[level * nq + q]) and turn ranks into representative indices
__global__ void k_idx_merge(const int4 *partial, u32 nq, u32 nslab, const u32 *order, int *best_sim, u32 *best_first, u32 *best_last) {
    const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int bs = 0; u32 bf = 0xFFFFFFFFu, bl = 0;
    for (u32 s = 0; s < nslab; ++s) {
        const int4 v = partial[(size_t)s * nq + q];
        if (v.x > bs) { bs = v.x; bf = (u32)v.y; bl = (u32)v.z; }
        else if (v.x == bs && v.x > 0) { bf = min(bf, (u32)v.y); bl = max(bl, (u32)v.z); }
    }
    best_sim[q] = bs;
    if (best_first) best_first[q] = bs > 0 ? order[bf] : 0xFFFFFFFFu;
    if (best_last) best_last[q] = bs > 0 ? order[bl] : 0xFFFFFFFFu;
}

__global__ void k_collect_empty(const SeqMeta *meta, u32 n, u32 *out, u32 *count) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && meta[i].size == 0) out[atomicAdd(count, 1u)] = i;
}

}  // namespace kr
