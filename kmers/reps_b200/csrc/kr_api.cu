// kr_api.cu -- C ABI of libkmersreps_cuda.so (see include/kmersreps.h) and the host-side
// orchestration of the kernels in kr_kernels.cuh.  No CPU fallback anywhere: every entry point
// that produces a similarity launches a kernel on the context's GPU.
#include "../../../include/kmersreps.h"
#include "kr_kernels.cuh"
#include "kr_encode.cuh"
#include "kr_index.cuh"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

using namespace kr;

// Kernel launches go through one macro so that the test emulator (tests/emul/, -DKR_EMUL, g++) can run the same host
// code; for nvcc it expands to exactly the <<< >>> launch it replaces.
#define KR_T(...) __VA_ARGS__
#ifdef KR_EMUL
#define KR_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), stream, [kr_args_ = std::make_tuple(__VA_ARGS__)] { std::apply(kernel, kr_args_); })   // (arguments by value, as a launch takes them)
#else
#define KR_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_create_error;

struct EvPair { cudaEvent_t a, b; int stage; };
enum { ST_ENCODE = 0, ST_BLOCKLIST = 1, ST_JOIN = 2, ST_OTHER = 3, ST_INDEX = 4, ST_SCAN = 5, ST_FILL = 6 };

struct DevMem {   // refcounted device allocation (residue buffers are shared by gathered sets)
    void *p = nullptr;
    cudaStream_t stream = nullptr;
    ~DevMem() { if (p) cudaFreeAsync(p, stream); }   // stream-ordered: safe after every kernel already queued
};
static cudaError_t pool_alloc(struct kr_ctx *ctx, void **p, size_t bytes);

struct kr_ctx {
    int dev = 0, K = 8;
    u32 flags = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, fill_stream = nullptr;   // fill_stream: zero-fill of a dense result beside the index build
    cudaEvent_t ev_fill[2] = {nullptr, nullptr};
    cudaMemPool_t pool = nullptr;   // this context's own memory pool: the device's default pool (shared with the JVM / torch) is never touched
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    std::mutex mu;              // serialises the calls on this context
    mutable std::mutex err_mu;  // guards `err` alone (fail() also runs before `mu` is taken)
    std::string err;
    u8 *d_lut = nullptr;
    ErrInfo *d_err = nullptr;
    kr_stats st{};
    std::vector<EvPair> pending;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool err_pending = false;   // an encode launched by a _device call has not had its error flag read yet
    u64 keyset_nonce = 1;       // kr_seqset_from_keys without sequences: every such object gets a hash of its own
    // cached tile list of the last matrix launch (the same stripes are usually computed again and again)
    uint2 *d_tasks_cache = nullptr; u32 tasks_cache_n = 0; u32 tasks_key[5] = {0, 0, 0, 0, 0xFFFFFFFFu};
};

struct kr_seqset {
    kr_ctx *ctx = nullptr;
    u32 n = 0;
    std::shared_ptr<DevMem> res;   // residues (kept alive for the equal-sequence check)
    u64 *d_off = nullptr;          // residue offsets (null for gathered sets)
    u64 *d_koff = nullptr;         // key-list offsets [n+1]
    u64 *d_keys = nullptr;
    SeqMeta *d_meta = nullptr;
    u64 key_cap = 0;
    std::vector<u32> h_sizes;      // lazily fetched
    // block lists (built on first use)
    bool bl_ready = false;
    u32 nblk = 0;
    u32 *d_hi = nullptr; void *d_lo = nullptr; u32 *d_choff = nullptr, *d_chcnt = nullptr; u64 *d_blkbase = nullptr;
    BlockRef *d_refs = nullptr;
    std::vector<std::shared_ptr<DevMem>> keep;   // parents' residue buffers
    // sorted inverted index over the set's own sequences (built on first use by a matrix scan), and its k-mer-less sequences
    struct IndexOwned *ix = nullptr;
    u32 *d_empty = nullptr, *d_nempty = nullptr;
};

struct DbSegment { kr_seqset *set; u32 base; };
struct kr_db {
    kr_ctx *ctx = nullptr;
    std::vector<DbSegment> segs;
    u32 n = 0;
    // database-wide arrays, rebuilt when the database changes
    bool dirty = true;
    SeqMeta *d_meta = nullptr; BlockRef *d_refs = nullptr; u32 *d_rank = nullptr, *d_order = nullptr;
    u32 nblk = 0, cap = 0, cap_blk = 0;
    std::vector<u32> h_rank;   // empty = append order
    bool refs_ready = false;   // block lists / block refs (tile-join path only) are current
    // Sorted inverted index in at most two levels: `main` over representatives [0, main_n) and `tail` over
    // [main_n, tail_n).  Representatives are only ever appended, so a level stays valid; the tail is rebuilt when the
    // database has grown and folded into the main level when it reaches a quarter of it (greedy build).
    struct IndexOwned *ix_main = nullptr, *ix_tail = nullptr;
    u32 main_n = 0, tail_n = 0;
    u32 *d_empty_reps = nullptr, *d_nempty_reps = nullptr; u32 empty_n = 0;   // k-mer-less representatives among [0, empty_n)
};

// ------------------------------------------------------------------------------------------------
static kr_status fail(kr_ctx *ctx, kr_status code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (ctx) { std::lock_guard<std::mutex> lk(ctx->err_mu); ctx->err = buf; } else g_create_error = buf;
    return code;
}
// key slots a sequence of L residues can need (one per k-mer position)
static inline u64 kmer_slots(const kr_ctx *ctx, u64 L) {
    return L >= (u64)ctx->K ? L - ctx->K + 1 : 0;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(ctx, KR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T> static cudaError_t dalloc(kr_ctx *ctx, T **p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    return pool_alloc(ctx, (void **)p, count * sizeof(T));
}
static cudaError_t pool_alloc(kr_ctx *ctx, void **p, size_t bytes) { return cudaMallocFromPoolAsync(p, bytes ? bytes : 1, ctx->pool, ctx->stream); }
template <typename T> static void dfree(kr_ctx *ctx, T *p) { if (p) cudaFreeAsync((void *)p, ctx->stream); }

static cudaEvent_t get_event(kr_ctx *ctx) {
    if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct StageTimer {   // device time of a group of launches on one stream, resolved lazily in kr_stats_get
    kr_ctx *ctx; EvPair ev; cudaStream_t stream;
    StageTimer(kr_ctx *c, int stage, cudaStream_t on = nullptr) : ctx(c), stream(on ? on : c->stream) {
        ev.a = get_event(c); ev.b = get_event(c); ev.stage = stage; cudaEventRecord(ev.a, stream);
    }
    ~StageTimer() { cudaEventRecord(ev.b, stream); ctx->pending.push_back(ev); if (ctx->pending.size() > 4096) drain(ctx); }
    static void drain(kr_ctx *c) {
        for (auto &e : c->pending) {
            cudaEventSynchronize(e.b);
            float ms = 0; cudaEventElapsedTime(&ms, e.a, e.b);
            double *dst = e.stage == ST_ENCODE ? &c->st.encode_ms : e.stage == ST_BLOCKLIST ? &c->st.blocklist_ms
                        : e.stage == ST_JOIN ? &c->st.join_ms : e.stage == ST_INDEX ? &c->st.index_ms
                        : e.stage == ST_SCAN ? &c->st.scan_ms : e.stage == ST_FILL ? &c->st.fill_ms : &c->st.other_ms;
            *dst += ms;
            c->ev_pool.push_back(e.a); c->ev_pool.push_back(e.b);
        }
        c->pending.clear();
    }
};
#define LAUNCHED(ctx) ((ctx)->st.kernel_launches++)

// Chunk table: the 32^3 three-residue key prefixes, in key order, are cut into NCHUNK contiguous
// groups of about equal expected mass under the background residue frequencies of seed proteins
// (counts measured on data/rep50.db).  It only balances work; any monotone table is
// correct.
static void make_chunk_lut(u8 *lut) {
    double w[32];
    for (int i = 0; i < 32; ++i) w[i] = 40.0;
    const char *aa = "LEGARIVDKFTPSNQMHYCW";
    const double cnt[20] = {48827, 42607, 37306, 35185, 33689, 32513, 31252, 30691, 30145, 29359,
                            25571, 25444, 25220, 19059, 18890, 15768, 13392, 13332, 6459, 4267};
    for (int i = 0; i < 20; ++i) w[aa[i] - 'A'] = cnt[i];
    double tot = 0;
    for (int i = 0; i < 32; ++i) tot += w[i];
    double total3 = tot * tot * tot, cum = 0;
    for (int p = 0; p < (1 << LUT_BITS); ++p) {
        double m = w[p >> 10] * w[(p >> 5) & 31] * w[p & 31];
        int c = (int)((cum + 0.5 * m) * NCHUNK / total3);
        lut[p] = (u8)std::min(c, NCHUNK - 1);
        cum += m;
    }
    for (int p = 1; p < (1 << LUT_BITS); ++p) if (lut[p] < lut[p - 1]) lut[p] = lut[p - 1];
}

// ------------------------------------------------------------------------------------------------
extern "C" {

// The message is copied under its own lock into a buffer of the CALLING thread: argument checks fail before the context's
// call lock is taken, and another thread may be failing on the same context at the same moment.  The pointer stays valid
// until this thread asks again.
const char *kr_last_error(const kr_ctx *ctx) {
    static thread_local std::string tl_last_error;
    if (!ctx) return g_create_error.c_str();
    std::lock_guard<std::mutex> lk(ctx->err_mu);
    tl_last_error = ctx->err;
    return tl_last_error.c_str();
}
const char *kr_build_info(void) { return "default"; }
int kr_kmer_size(const kr_ctx *ctx) { return ctx ? ctx->K : 0; }


kr_status kr_create(int device, int K, uint32_t flags, kr_ctx **out) {
    kr_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, KR_ERR_ARG, "kr_create: out is null");
    *out = nullptr;
    if (flags & ~(KR_IDENT_IS_INFINITY | KR_TIE_LAST | KR_FORCE_JOIN | KR_FORCE_INDEX)) return fail(nullptr, KR_ERR_ARG, "kr_create: unknown flags 0x%x", flags);
    if (K < KR_MIN_K || K > KR_MAX_K)
        return fail(nullptr, KR_ERR_ARG, "kr_create: K=%d outside [%d,%d] (5 bits per residue: a 64-bit entry has no room for a sequence id beside a key of 12 or more residues)",
                    K, KR_MIN_K, KR_MAX_K);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, KR_ERR_NODEVICE, "kr_create: no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail(nullptr, KR_ERR_ARG, "kr_create: device %d of %d", device, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, KR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, KR_ERR_NODEVICE, "kr_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    if ((size_t)prop.sharedMemPerBlockOptin < JOIN_SMEM_BYTES)
        return fail(nullptr, KR_ERR_NODEVICE, "kr_create: device offers %zu B shared memory per block, need %zu", (size_t)prop.sharedMemPerBlockOptin, JOIN_SMEM_BYTES);
    ctx = new kr_ctx();
    ctx->dev = device; ctx->K = K; ctx->flags = flags; ctx->num_sms = prop.multiProcessorCount;
#define CUC(call) do { cudaError_t e2 = (call); if (e2 != cudaSuccess) { fail(nullptr, KR_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e2)); delete ctx; return KR_ERR_CUDA; } } while (0)
    CUC(cudaSetDevice(device));
    CUC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&ctx->fill_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) CUC(cudaEventCreateWithFlags(&ctx->ev_fill[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        CUC(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
    }
    {   // a private pool that keeps its memory between calls (steps re-use the same sizes); freed with the context
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        CUC(cudaMemPoolCreate(&ctx->pool, &props));
        unsigned long long keep = ~0ull;
        CUC(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    std::vector<u8> lut(1 << LUT_BITS); make_chunk_lut(lut.data());
    CUC(cudaMalloc(&ctx->d_lut, lut.size()));
    CUC(cudaMemcpy(ctx->d_lut, lut.data(), lut.size(), cudaMemcpyHostToDevice));
    CUC(cudaMalloc(&ctx->d_err, sizeof(ErrInfo)));
    CUC(cudaMemset(ctx->d_err, 0, sizeof(ErrInfo)));
    CUC(cudaFuncSetAttribute(k_ix_query<IXQ_ARGMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IXQ_SMEM));
    CUC(cudaFuncSetAttribute(k_ix_query<IXQ_CELLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IXQ_SMEM));
    CUC(cudaFuncSetAttribute(k_ix_query<IXQ_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IXQ_SMEM));
    CUC(cudaFuncSetAttribute(k_join<MODE_SPARSE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)join_smem_bytes(false)));
    CUC(cudaFuncSetAttribute(k_join<MODE_SPARSE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)join_smem_bytes(true)));
    CUC(cudaFuncSetAttribute(k_join<MODE_MATRIX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)join_smem_bytes(false)));
    CUC(cudaFuncSetAttribute(k_join<MODE_ARGMAX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)join_smem_bytes(false)));
    CUC(cudaFuncSetAttribute(k_join<MODE_MATRIX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)join_smem_bytes(true)));
    CUC(cudaFuncSetAttribute(k_join<MODE_ARGMAX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)join_smem_bytes(true)));
#undef CUC
    *out = ctx;
    return KR_OK;
}

void kr_destroy(kr_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->dev);
    cudaStreamSynchronize(ctx->stream);
    StageTimer::drain(ctx);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->t0) { cudaEventDestroy(ctx->t0); cudaEventDestroy(ctx->t1); }
    cudaFree(ctx->d_lut); cudaFree(ctx->d_err);
    if (ctx->d_tasks_cache) cudaFree(ctx->d_tasks_cache);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(ctx->ev_done[i]); cudaEventDestroy(ctx->ev_copied[i]); }
    cudaStreamSynchronize(ctx->fill_stream);
    for (int i = 0; i < 2; ++i) cudaEventDestroy(ctx->ev_fill[i]);
    cudaStreamDestroy(ctx->fill_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
}

static kr_status check_alphabet(kr_ctx *ctx);

kr_status kr_synchronize(kr_ctx *ctx) {
    if (!ctx) return KR_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->err_pending) return check_alphabet(ctx);
    return KR_OK;
}

kr_status kr_stats_get(kr_ctx *ctx, kr_stats *out) {
    if (!ctx || !out) return KR_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    CU(cudaStreamSynchronize(ctx->stream));
    StageTimer::drain(ctx);
    *out = ctx->st;
    return KR_OK;
}
kr_status kr_stats_reset(kr_ctx *ctx) {
    if (!ctx) return KR_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->dev);
    cudaStreamSynchronize(ctx->stream);
    StageTimer::drain(ctx);
    ctx->st = kr_stats{};
    return KR_OK;
}

kr_status kr_timer_begin(kr_ctx *ctx) {
    if (!ctx) return KR_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    if (!ctx->t0) { CU(cudaEventCreate(&ctx->t0)); CU(cudaEventCreate(&ctx->t1)); }
    CU(cudaEventRecord(ctx->t0, ctx->stream));
    return KR_OK;
}
kr_status kr_timer_end(kr_ctx *ctx, double *ms) {
    if (!ctx || !ms) return KR_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    if (!ctx->t0) return fail(ctx, KR_ERR_STATE, "kr_timer_end without kr_timer_begin");
    CU(cudaEventRecord(ctx->t1, ctx->stream));
    CU(cudaEventSynchronize(ctx->t1));
    float f = 0;
    CU(cudaEventElapsedTime(&f, ctx->t0, ctx->t1));
    *ms = f;
    if (ctx->err_pending) return check_alphabet(ctx);
    return KR_OK;
}

kr_status kr_host_alloc(kr_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return KR_ERR_ARG;
    CU(cudaSetDevice(ctx->dev));
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return KR_OK;
}
void kr_host_free(kr_ctx *ctx, void *p) { if (ctx) cudaSetDevice(ctx->dev); if (p) cudaFreeHost(p); }

double kr_distance(int32_t sim, uint32_t size_a, uint32_t size_b) {
    if (sim == KR_INFINITY) return 0.0;
    if (sim == 0) return 1.0;
    return 1.0 - (sim / (double)((int64_t)size_a + (int64_t)size_b - sim));
}

kr_status kr_key_decode(const kr_ctx *ctx, uint64_t key, char *out) {
    if (!ctx || !out) return KR_ERR_ARG;
    for (int i = ctx->K - 1; i >= 0; --i) {
        int c = (int)(key & 31); key >>= 5;
        out[i] = c < 26 ? (char)('A' + c) : c == 26 ? '*' : c == 27 ? '-' : '?';
    }
    return KR_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
static kr_status fetch_sizes(kr_ctx *ctx, kr_seqset *s);
#include "kr_api_index.inc"

static void seqset_destroy(kr_seqset *s) {
    if (!s) return;
    kr_ctx *ctx = s->ctx;
    cudaSetDevice(ctx->dev);
    if (s->ix) { index_free(ctx, s->ix); delete s->ix; s->ix = nullptr; }
    dfree(ctx, s->d_empty); dfree(ctx, s->d_nempty);
    dfree(ctx, s->d_off); dfree(ctx, s->d_koff); dfree(ctx, s->d_keys); dfree(ctx, s->d_meta);
    dfree(ctx, s->d_hi); dfree(ctx, (u8 *)s->d_lo); dfree(ctx, s->d_choff); dfree(ctx, s->d_chcnt); dfree(ctx, s->d_blkbase); dfree(ctx, s->d_refs);
    delete s;   // residue buffers are released stream-ordered by ~DevMem
}

// read (and clear) the context's device-side error record: a residue outside the alphabet (K1) or a tile join that
// had to skip work (K3: misaligned block list, watchdog); synchronises the stream
static kr_status check_alphabet(kr_ctx *ctx) {
    ErrInfo ei;
    CU(cudaMemcpyAsync(&ei, ctx->d_err, sizeof ei, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->err_pending = false;
    if (ei.flag || ei.join_fault) cudaMemsetAsync(ctx->d_err, 0, sizeof(ErrInfo), ctx->stream);
    if (ei.join_fault == JOIN_FAULT_ALIGN)
        return fail(ctx, KR_ERR_STATE, "tile join: a block list is not 16-byte aligned (CTA %u, entry offsets %u / %u); the tile was skipped and "
                    "the result of this call is invalid", ei.join_info[0], ei.join_info[1], ei.join_info[2]);
    if (ei.join_fault)
        return fail(ctx, KR_ERR_STATE, "tile join: a pipeline wait timed out in CTA %u (slot %u); the launch was drained and the result of this "
                    "call is invalid", ei.join_info[0], ei.join_info[1]);
    if (ei.flag) {
        return fail(ctx, KR_ERR_ALPHABET, "kr_encode: sequence %u position %u: residue 0x%02x ('%c') is outside %s",
                    ei.seq, ei.pos, ei.byte, (ei.byte >= 32 && ei.byte < 127) ? (char)ei.byte : '?',
                    "'A'..'Z', '*', '-'");
    }
    return KR_OK;
}

static kr_status encode_impl(kr_ctx *ctx, const u8 *residues, bool on_device, const u64 *offsets, u32 nseq, kr_seqset **out) {
    if (!ctx || !out || (!offsets && nseq)) return fail(ctx, KR_ERR_ARG, "kr_encode: null argument");
    *out = nullptr;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    const int K = ctx->K;
    u64 total = nseq ? offsets[nseq] - offsets[0] : 0;
    if (nseq && offsets[nseq] < offsets[0]) return fail(ctx, KR_ERR_ARG, "kr_encode: offsets not monotone");
    if (total && !residues) return fail(ctx, KR_ERR_ARG, "kr_encode: null residues");
    std::unique_ptr<kr_seqset, void (*)(kr_seqset *)> s(new kr_seqset(), seqset_destroy);
    s->ctx = ctx; s->n = nseq;
    s->res = std::make_shared<DevMem>();
    s->res->stream = ctx->stream;
    CU(pool_alloc(ctx, &s->res->p, total + 16));
    // Host batches of a few MB and more are encoded as a pipeline: the residues travel in ENC_PIPE pieces (cut at sequence
    // boundaries) on the copy stream, and the encoder of piece c starts as soon as piece c has arrived, while the later
    // pieces are still on the bus -- and while the host is still walking the offsets below.
    constexpr u32 ENC_PIPE = 4;
#ifndef KR_ENC_PIPE_MIN
#define KR_ENC_PIPE_MIN (4u << 20)      // bytes of residues from which the pipeline pays (the tests build a variant that always pipelines)
#endif
    const bool piped = !on_device && total >= KR_ENC_PIPE_MIN && nseq >= ENC_PIPE;
    u32 cut[ENC_PIPE + 1];
    cudaEvent_t ev_piece[ENC_PIPE];
    for (u32 c = 0; c <= ENC_PIPE; ++c) cut[c] = (u32)((u64)nseq * c / ENC_PIPE);
    if (on_device) CU(cudaMemcpyAsync(s->res->p, residues + offsets[0], total, cudaMemcpyDeviceToDevice, ctx->stream));
    else if (!piped) { CU(cudaMemcpyAsync(s->res->p, residues + offsets[0], total, cudaMemcpyHostToDevice, ctx->stream)); ctx->st.h2d_bytes += total; }
    else {
        cudaEvent_t ev_alloc = get_event(ctx);
        CU(cudaEventRecord(ev_alloc, ctx->stream));                     // the residue buffer exists (stream-ordered allocation)
        CU(cudaStreamWaitEvent(ctx->copy_stream, ev_alloc, 0));
        ctx->ev_pool.push_back(ev_alloc);
        for (u32 c = 0; c < ENC_PIPE; ++c) {
            const u64 b0 = offsets[cut[c]] - offsets[0], b1 = offsets[cut[c + 1]] - offsets[0];
            // (pieces of an offset table that is not monotone are not copied; the walk below reports it)
            if (b1 > b0 && b1 <= total) CU(cudaMemcpyAsync((u8 *)s->res->p + b0, residues + offsets[0] + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->copy_stream));
            ev_piece[c] = get_event(ctx);
            CU(cudaEventRecord(ev_piece[c], ctx->copy_stream));
        }
        ctx->st.h2d_bytes += total;
    }
    std::vector<u64> off(nseq + 1), koff(nseq + 1);
    u64 kc = 0; u32 maxn = 0;
    kr_status bad = KR_OK;
    for (u32 i = 0; i < nseq && bad == KR_OK; ++i) {
        if (offsets[i + 1] < offsets[i]) { bad = fail(ctx, KR_ERR_ARG, "kr_encode: offsets not monotone at %u", i); break; }
        u64 L = offsets[i + 1] - offsets[i];
        if (kmer_slots(ctx, L) > 65535u) { bad = fail(ctx, KR_ERR_ARG, "kr_encode: sequence %u has %llu residues; at most 65535 k-mers per sequence are supported", i, (unsigned long long)L); break; }
        off[i] = offsets[i] - offsets[0]; koff[i] = kc;
        u32 n = (u32)kmer_slots(ctx, L);
        kc += n; maxn = std::max(maxn, n);
    }
    if (bad != KR_OK) {     // copies already queued on the copy stream must have landed before the buffer is released
        if (piped) { cudaStreamSynchronize(ctx->copy_stream); for (u32 c = 0; c < ENC_PIPE; ++c) ctx->ev_pool.push_back(ev_piece[c]); }
        return bad;
    }
    off[nseq] = total; koff[nseq] = kc;
    s->key_cap = kc;
    CU(dalloc(ctx, &s->d_off, nseq + 1)); CU(dalloc(ctx, &s->d_koff, nseq + 1));
    CU(dalloc(ctx, &s->d_keys, kc + 16)); CU(dalloc(ctx, &s->d_meta, nseq));
    CU(cudaMemcpyAsync(s->d_off, off.data(), (nseq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(s->d_koff, koff.data(), (nseq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->st.h2d_bytes += 2ull * (nseq + 1) * 8;
    if (nseq) {
        StageTimer t(ctx, ST_ENCODE);
        // one warp per sequence, keys sorted in registers; the 16-keys-per-lane build serves every batch whose longest
        // sequence has at most 512 k-mers (seed proteins), the 32-keys-per-lane build the others
        for (u32 c = 0; c < (piped ? ENC_PIPE : 1u); ++c) {
            const u32 s0 = piped ? cut[c] : 0, s1 = piped ? cut[c + 1] : nseq;
            if (piped) { CU(cudaStreamWaitEvent(ctx->stream, ev_piece[c], 0)); ctx->ev_pool.push_back(ev_piece[c]); }   // piece c has arrived
            if (s1 == s0) continue;
            EncodeParams p;
            p.res = (const u8 *)s->res->p; p.off = s->d_off + s0; p.koff = s->d_koff + s0; p.keys = s->d_keys; p.meta = s->d_meta + s0;
            p.err = ctx->d_err; p.nseq = s1 - s0; p.K = K; p.seq_base = s0;
            const u32 grid = (s1 - s0 + ENC2_WARPS - 1) / ENC2_WARPS;
            if (maxn <= 512) KR_LAUNCH(KR_T(k_encode_sort_warp<16>), grid, ENC2_WARPS * 32, 0, ctx->stream, p);
            else KR_LAUNCH(KR_T(k_encode_sort_warp<32>), grid, ENC2_WARPS * 32, 0, ctx->stream, p);
            LAUNCHED(ctx);
        }
        CU(cudaGetLastError());
    }
    if (on_device) {   // "_device" calls do not synchronise the host: the flag is read at the next call that does
        ctx->err_pending = true;
        *out = s.release();
        return KR_OK;
    }
    kr_status st = check_alphabet(ctx);
    if (st) return st;
    *out = s.release();
    return KR_OK;
}

// entries + per-chunk alignment padding + the sentinel tail after every block
static u64 blocklist_capacity(u64 key_cap, u32 nblk) { return key_cap + (u64)nblk * NCHUNK * CHUNK_ALIGN + (u64)nblk * TAILPAD + 64; }

// block lists of a set (K2), built on first use
static kr_status ensure_blocklists(kr_ctx *ctx, kr_seqset *s) {
    if (s->bl_ready) return KR_OK;
    u32 nblk = (s->n + TB - 1) / TB;
    s->nblk = nblk;
    u64 cap = blocklist_capacity(s->key_cap, nblk);
    const int lobits = lo_bits(ctx->K), lo_size = lobits <= 16 ? 2 : 4;
    u64 *d_ent = nullptr; u8 *d_lo = nullptr;
    CU(dalloc(ctx, &d_ent, cap)); CU(dalloc(ctx, &s->d_hi, cap)); CU(dalloc(ctx, &d_lo, cap * lo_size));
    s->d_lo = d_lo;
    CU(dalloc(ctx, &s->d_choff, (size_t)nblk * (NCHUNK + 1)));
    CU(dalloc(ctx, &s->d_chcnt, (size_t)nblk * NCHUNK)); CU(dalloc(ctx, &s->d_blkbase, nblk + 1)); CU(dalloc(ctx, &s->d_refs, nblk));
    if (nblk) {
        StageTimer t(ctx, ST_BLOCKLIST);
        int shift = 5 * ctx->K - LUT_BITS;
        KR_LAUNCH(KR_T(k_bl_count), nblk, 256, 0, ctx->stream, s->d_keys, s->d_koff, s->d_meta, s->n, ctx->d_lut, shift, s->d_chcnt);
        KR_LAUNCH(KR_T(k_bl_offsets), 1, 1024, 0, ctx->stream, s->d_chcnt, nblk, s->d_choff, s->d_blkbase);
        KR_LAUNCH(KR_T(k_bl_scatter), nblk, 256, 0, ctx->stream, s->d_keys, s->d_koff, s->d_meta, s->n, ctx->d_lut, shift, s->d_chcnt, s->d_choff,
                                                    s->d_blkbase, d_ent, s->d_hi, s->d_lo, lo_size, s->d_refs);
        KR_LAUNCH(KR_T(k_bl_sort), nblk * NCHUNK, 256, 0, ctx->stream, s->d_refs, s->d_blkbase, d_ent, lobits);
        ctx->st.kernel_launches += 4;
        CU(cudaGetLastError());
    }
    dfree(ctx, d_ent);   // the unsplit entries were only scratch for the sort
    s->bl_ready = true;
    return KR_OK;
}

static kr_status fetch_sizes(kr_ctx *ctx, kr_seqset *s) {
    if (s->h_sizes.size() == s->n) return KR_OK;
    std::vector<SeqMeta> m(s->n);
    CU(cudaMemcpyAsync(m.data(), s->d_meta, (size_t)s->n * sizeof(SeqMeta), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->err_pending) { kr_status st = check_alphabet(ctx); if (st) return st; }
    ctx->st.d2h_bytes += (u64)s->n * sizeof(SeqMeta);
    s->h_sizes.resize(s->n);
    for (u32 i = 0; i < s->n; ++i) s->h_sizes[i] = m[i].size;
    return KR_OK;
}

static kr_status gather_impl(kr_ctx *ctx, const kr_seqset *src_c, const u32 *idx, u32 n, kr_seqset **out) {
    kr_seqset *src = const_cast<kr_seqset *>(src_c);
    if (!src->d_keys && src->n) return fail(ctx, KR_ERR_STATE, "this set was imported from block lists and has no key lists");
    kr_status st = fetch_sizes(ctx, src);
    if (st) return st;
    std::vector<u64> koff(n + 1);
    u64 kc = 0;
    for (u32 i = 0; i < n; ++i) {
        if (idx[i] >= src->n) return fail(ctx, KR_ERR_ARG, "gather: index %u out of range (%u)", idx[i], src->n);
        koff[i] = kc; kc += src->h_sizes[idx[i]];
    }
    koff[n] = kc;
    std::unique_ptr<kr_seqset, void (*)(kr_seqset *)> s(new kr_seqset(), seqset_destroy);
    s->ctx = ctx; s->n = n; s->key_cap = kc;
    if (src->res) s->keep.push_back(src->res);
    for (auto &k : src->keep) s->keep.push_back(k);
    u32 *d_idx = nullptr;
    CU(dalloc(ctx, &d_idx, n)); CU(dalloc(ctx, &s->d_koff, n + 1)); CU(dalloc(ctx, &s->d_keys, kc + 16)); CU(dalloc(ctx, &s->d_meta, n));
    CU(cudaMemcpyAsync(d_idx, idx, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(s->d_koff, koff.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
        StageTimer t(ctx, ST_OTHER);
        KR_LAUNCH(KR_T(k_gather_keys), n, 128, 0, ctx->stream, src->d_keys, src->d_koff, src->d_meta, d_idx, n, s->d_koff, s->d_keys, s->d_meta);
        LAUNCHED(ctx);
        CU(cudaGetLastError());
    }
    dfree(ctx, d_idx);   // stream-ordered; idx / koff were staged before their copy calls returned
    s->h_sizes.resize(n);
    for (u32 i = 0; i < n; ++i) s->h_sizes[i] = src->h_sizes[idx[i]];
    *out = s.release();
    return KR_OK;
}

// ------------------------------------------------------------------------------------------------
static u64 host_cell_index(u32 layout, u64 i, u64 j, u64 nrows, u64 ncols) {
    if (layout == KR_LAYOUT_FULL) return i * ncols + j;
    if (layout == KR_LAYOUT_TRIANGLE) return i * nrows - i * (i + 1) / 2 + (j - i - 1);
    u64 I = i >> 8, c0 = I << 8;
    return 256ull * (I * nrows - 128ull * I * (I - 1)) + (i - c0) * (nrows - c0) + (j - c0);
}
// element index of the first cell of row r (== number of cells in rows < r)
static u64 row_start(u32 layout, u64 r, u64 nrows, u64 ncols) {
    if (layout == KR_LAYOUT_FULL) return r * ncols;
    if (layout == KR_LAYOUT_TRIANGLE) return r * nrows - r * (r + 1) / 2;
    u64 I = r >> 8, c0 = I << 8;
    return 256ull * (I * nrows - 128ull * I * (I - 1)) + (r - c0) * (nrows - c0);
}

struct SparseSink { uint4 *cells; unsigned long long cap; unsigned long long *count; u32 min_sim; bool packed; };


static kr_status matrix_launch(kr_ctx *ctx, kr_seqset *A, kr_seqset *B, bool self, const kr_matrix_desc *d, void *out_dev,
                               const SparseSink *sparse = nullptr) {
    kr_status st;
    {
        // sparse collections go through the index scan: zero-fill the result (dense) and write only the cells of pairs
        // that share a k-mer; dense ones through the tile join below
        bool use_index = false;
        kr_seqset *Bx = self ? A : B;
        // A dense result of the index path is zeros plus a few cells.  When this call still has to BUILD the index (every
        // step of a pipeline that re-encodes its batch), the zero-fill -- pure HBM write traffic -- is started on a second
        // stream first and runs beside the build (atomics, shared-memory sorting).  It is started before the path is known
        // (the choice needs the built index's statistics); should the join be chosen after all it simply overwrites every
        // cell, the fill having cost a few percent of a join.  Small collections never take the index and are not filled.
        bool filled = false;
        const u64 rows = d->row_end > d->row_begin ? d->row_end - d->row_begin : 0;
        if (!sparse && rows && !(ctx->flags & KR_FORCE_JOIN) && A->d_keys && Bx->n && set_indexable(ctx, Bx) && !(Bx->ix && Bx->ix->ok) &&
            ((ctx->flags & KR_FORCE_INDEX) || (double)A->key_cap * 3.0 <= (double)A->n * (double)Bx->n)) {
            const u64 bytes = (row_start(d->layout, d->row_end, A->n, Bx->n) - row_start(d->layout, d->row_begin, A->n, Bx->n)) * (d->elem == KR_ELEM_U16 ? 2 : 4);
            CU(cudaEventRecord(ctx->ev_fill[0], ctx->stream));                  // after everything queued so far (earlier users of `out`)
            CU(cudaStreamWaitEvent(ctx->fill_stream, ctx->ev_fill[0], 0));
            {
                StageTimer t(ctx, ST_FILL, ctx->fill_stream);
                KR_LAUNCH(KR_T(k_zero_fill), std::max(1, ctx->num_sms * 2), 256, 0, ctx->fill_stream, (u8 *)out_dev, bytes);
                LAUNCHED(ctx); ctx->st.fill_launches++;
            }
            CU(cudaGetLastError());
            CU(cudaEventRecord(ctx->ev_fill[1], ctx->fill_stream));
            filled = true;
        }
        st = choose_index_path(ctx, A, Bx, self, &use_index);
        // whatever happens next on the main stream (scan, join, the caller's frees) comes after the fill
        if (filled) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_fill[1], 0));
        if (st) return st;
        if (use_index && rows) {
            MatrixTarget tg;
            if (sparse) { tg.cells = sparse->cells; tg.cells_cap = sparse->cap; tg.cells_count = sparse->count; tg.min_sim = sparse->min_sim; tg.packed = sparse->packed; }
            else {
                tg.out = out_dev; tg.elem = d->elem; tg.layout = d->layout; tg.origin = row_start(d->layout, d->row_begin, A->n, Bx->n); tg.min_sim = 1;
                // (the scan leaves every cell of its rows written: zeros -- by the fill above, or else row by row by the warp
                // that owns the row -- and then the cells of pairs that share a k-mer)
                tg.prezeroed = filled;
            }
            if ((st = index_matrix(ctx, A, Bx, self, d->layout, d->row_begin, d->row_end, tg))) return st;
            if (d->layout == KR_LAYOUT_FULL) ctx->st.pairs += rows * (u64)Bx->n;
            else ctx->st.pairs += kr_matrix_elems(A->n, A->n, KR_LAYOUT_TRIANGLE, d->row_begin, d->row_end);
            return KR_OK;
        }
        if (use_index) return KR_OK;
    }
    if ((st = ensure_blocklists(ctx, A))) return st;
    if (!self && (st = ensure_blocklists(ctx, B))) return st;
    kr_seqset *Bs = self ? A : B;
    u32 rb0 = d->row_begin / TB, rb1 = (d->row_end + TB - 1) / TB;
    // Tile order: super-tiles of ST x ST blocks, so that the ~150 tiles in flight at any time share
    // 2*ST block lists (a few MB) and re-reads of a block list hit L2 instead of HBM.
    const u32 key[5] = {rb0, rb1, A->nblk, Bs->nblk, d->layout == KR_LAYOUT_FULL ? 0u : 1u};
    if (memcmp(key, ctx->tasks_key, sizeof key) != 0 || !ctx->d_tasks_cache) {
        std::vector<uint2> tasks;
        const u32 ST = 16;
        for (u32 SI = rb0; SI < rb1; SI += ST)
            for (u32 SJ = (d->layout == KR_LAYOUT_FULL ? 0 : SI / ST * ST); SJ < Bs->nblk; SJ += ST)
                for (u32 I = SI; I < std::min(SI + ST, rb1); ++I)
                    for (u32 J = std::max(SJ, d->layout == KR_LAYOUT_FULL ? 0u : I); J < std::min(SJ + ST, Bs->nblk); ++J)
                        tasks.push_back(make_uint2(I, J));
        dfree(ctx, ctx->d_tasks_cache);
        ctx->d_tasks_cache = nullptr; ctx->tasks_cache_n = (u32)tasks.size();
        CU(dalloc(ctx, &ctx->d_tasks_cache, tasks.size()));
        CU(cudaMemcpyAsync(ctx->d_tasks_cache, tasks.data(), tasks.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
        memcpy(ctx->tasks_key, key, sizeof key);
    }
    if (ctx->tasks_cache_n == 0) return KR_OK;
    uint2 *d_tasks = ctx->d_tasks_cache;
    JoinParams p{};
    p.A = A->d_refs; p.B = Bs->d_refs; p.metaA = A->d_meta; p.metaB = Bs->d_meta;
    p.nA = A->n; p.nB = Bs->n; p.nBblk = Bs->nblk;
    p.tasks = d_tasks; p.ntasks = ctx->tasks_cache_n; p.flags = ctx->flags; p.err = ctx->d_err; ctx->err_pending = true;
    p.out = out_dev; p.elem = d->elem; p.layout = d->layout;
    p.origin = row_start(d->layout, d->row_begin, A->n, Bs->n);
    {
        StageTimer t(ctx, ST_JOIN);
        u32 grid = std::min<u32>(p.ntasks, (u32)ctx->num_sms);
        if (sparse) {   // the join's epilogue keeps only the cells that reach min_sim
            p.cells = sparse->cells; p.cells_cap = sparse->cap; p.cells_count = sparse->count; p.min_sim = sparse->min_sim; p.cells_packed = sparse->packed ? 1u : 0u;
            if (lo_bits(ctx->K) <= 16) KR_LAUNCH(KR_T(k_join<MODE_SPARSE, false>), grid, JOIN_THREADS, join_smem_bytes(false), ctx->stream, p);
            else KR_LAUNCH(KR_T(k_join<MODE_SPARSE, true>), grid, JOIN_THREADS, join_smem_bytes(true), ctx->stream, p);
        }
        else if (lo_bits(ctx->K) <= 16) KR_LAUNCH(KR_T(k_join<MODE_MATRIX, false>), grid, JOIN_THREADS, join_smem_bytes(false), ctx->stream, p);
        else KR_LAUNCH(KR_T(k_join<MODE_MATRIX, true>), grid, JOIN_THREADS, join_smem_bytes(true), ctx->stream, p);
        LAUNCHED(ctx); ctx->st.join_launches++;
        CU(cudaGetLastError());
    }
    u64 rows = d->row_end - d->row_begin;
    if (d->layout == KR_LAYOUT_FULL) ctx->st.pairs += rows * (u64)Bs->n;
    else ctx->st.pairs += kr_matrix_elems(A->n, A->n, KR_LAYOUT_TRIANGLE, d->row_begin, d->row_end);
    return KR_OK;
}

static kr_status check_desc(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *d) {
    if (!A || !d) return fail(ctx, KR_ERR_ARG, "kr_sim_matrix: null argument");
    if (A->ctx != ctx || (B && B->ctx != ctx)) return fail(ctx, KR_ERR_STATE, "kr_sim_matrix: set belongs to another context");
    if (d->elem > KR_ELEM_U16 || d->layout > KR_LAYOUT_BLOCKTRI) return fail(ctx, KR_ERR_ARG, "kr_sim_matrix: bad elem/layout");
    if (d->layout != KR_LAYOUT_FULL && B && B != A) return fail(ctx, KR_ERR_ARG, "kr_sim_matrix: triangle layouts need B == NULL");
    if (d->row_begin % TB || d->row_begin > d->row_end || d->row_end > A->n || (d->row_end % TB && d->row_end != A->n))
        return fail(ctx, KR_ERR_ARG, "kr_sim_matrix: rows [%u,%u) must be multiples of %d inside [0,%u]", d->row_begin, d->row_end, TB, A->n);
    return KR_OK;
}

// ------------------------------------------------------------------------------------------------
static void db_free_arrays(kr_db *db) {
    kr_ctx *ctx = db->ctx;
    dfree(ctx, db->d_meta); dfree(ctx, db->d_refs); dfree(ctx, db->d_rank); dfree(ctx, db->d_order);
    db->d_meta = nullptr; db->d_refs = nullptr; db->d_rank = nullptr; db->d_order = nullptr;
}

static void db_free_index(kr_db *db) {
    kr_ctx *ctx = db->ctx;
    if (db->ix_main) { index_free(ctx, db->ix_main); delete db->ix_main; db->ix_main = nullptr; }
    if (db->ix_tail) { index_free(ctx, db->ix_tail); delete db->ix_tail; db->ix_tail = nullptr; }
    db->main_n = db->tail_n = 0;
    dfree(ctx, db->d_empty_reps); dfree(ctx, db->d_nempty_reps); db->d_empty_reps = nullptr; db->d_nempty_reps = nullptr; db->empty_n = 0;
}

// (re)build the database-wide record / rank arrays after a change (block refs only on demand: db_refresh_refs)
static kr_status db_refresh(kr_ctx *ctx, kr_db *db) {
    if (!db->dirty) return KR_OK;
    db_free_arrays(db);
    db->refs_ready = false;
    CU(dalloc(ctx, &db->d_meta, db->n)); CU(dalloc(ctx, &db->d_rank, db->n)); CU(dalloc(ctx, &db->d_order, db->n));
    for (auto &sg : db->segs) {
        kr_seqset *s = sg.set;
        if (!s->n) continue;
        CU(cudaMemcpyAsync(db->d_meta + sg.base, s->d_meta, (size_t)s->n * sizeof(SeqMeta), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (db->n) {
        if (db->h_rank.size() == db->n) {
            std::vector<u32> order(db->n);
            for (u32 i = 0; i < db->n; ++i) order[db->h_rank[i]] = i;
            CU(cudaMemcpyAsync(db->d_rank, db->h_rank.data(), (size_t)db->n * 4, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(db->d_order, order.data(), (size_t)db->n * 4, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        } else {
            KR_LAUNCH(KR_T(k_fill_u32_iota), (db->n + 255) / 256, 256, 0, ctx->stream, db->d_rank, db->n, 0);
            KR_LAUNCH(KR_T(k_fill_u32_iota), (db->n + 255) / 256, 256, 0, ctx->stream, db->d_order, db->n, 0);
            ctx->st.kernel_launches += 2;
        }
    }
    CU(cudaGetLastError());
    db->dirty = false;
    return KR_OK;
}

// block lists of every segment and the database-wide block-ref array (tile-join path)
static kr_status db_refresh_refs(kr_ctx *ctx, kr_db *db) {
    if (db->refs_ready) return KR_OK;
    kr_status st;
    u32 nblk = 0;
    for (auto &sg : db->segs) { if ((st = ensure_blocklists(ctx, sg.set))) return st; nblk += sg.set->nblk; }
    dfree(ctx, db->d_refs); db->d_refs = nullptr;
    CU(dalloc(ctx, &db->d_refs, nblk));
    u32 bo = 0;
    for (auto &sg : db->segs) {
        kr_seqset *s = sg.set;
        if (!s->n) continue;
        KR_LAUNCH(KR_T(k_copy_refs), (s->nblk + 127) / 128, 128, 0, ctx->stream, s->d_refs, s->nblk, sg.base, db->d_refs + bo);
        LAUNCHED(ctx);
        bo += s->nblk;
    }
    CU(cudaGetLastError());
    db->nblk = nblk;
    db->refs_ready = true;
    return KR_OK;
}

// Segments other than the first are merged once there are too many of them (greedy build appends one
// small segment per batch); the first is folded in when the tail has grown to half its size.
static kr_status db_compact(kr_ctx *ctx, kr_db *db) {
    if (db->segs.size() <= 6) return KR_OK;
    size_t first = 1;
    u32 tail = 0;
    for (size_t i = 1; i < db->segs.size(); ++i) tail += db->segs[i].set->n;
    if (tail * 2 >= db->segs[0].set->n) first = 0;
    // concatenate key lists of segments [first, end) into one new set
    std::unique_ptr<kr_seqset, void (*)(kr_seqset *)> s(new kr_seqset(), seqset_destroy);
    u32 n = 0; u64 kc = 0;
    for (size_t i = first; i < db->segs.size(); ++i) { n += db->segs[i].set->n; }
    s->ctx = ctx; s->n = n;
    std::vector<u64> koff(n + 1);
    std::vector<u32> sizes; sizes.reserve(n);
    kr_status st;
    for (size_t i = first; i < db->segs.size(); ++i) {
        kr_seqset *g = db->segs[i].set;
        if ((st = fetch_sizes(ctx, g))) return st;
        for (u32 k = 0; k < g->n; ++k) { koff[sizes.size()] = kc; kc += g->h_sizes[k]; sizes.push_back(g->h_sizes[k]); }
        if (g->res) s->keep.push_back(g->res);
        for (auto &kp : g->keep) s->keep.push_back(kp);
    }
    koff[n] = kc; s->key_cap = kc; s->h_sizes = sizes;
    CU(dalloc(ctx, &s->d_koff, n + 1)); CU(dalloc(ctx, &s->d_keys, kc + 16)); CU(dalloc(ctx, &s->d_meta, n));
    CU(cudaMemcpyAsync(s->d_koff, koff.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    u32 at = 0;
    for (size_t i = first; i < db->segs.size(); ++i) {
        kr_seqset *g = db->segs[i].set;
        if (!g->n) continue;
        // gathered / compacted sets have dense key arenas only if built by gather; use the gather kernel for generality
        u32 *d_idx = nullptr;
        CU(dalloc(ctx, &d_idx, g->n));
        KR_LAUNCH(KR_T(k_fill_u32_iota), (g->n + 255) / 256, 256, 0, ctx->stream, d_idx, g->n, 0);
        KR_LAUNCH(KR_T(k_gather_keys), g->n, 128, 0, ctx->stream, g->d_keys, g->d_koff, g->d_meta, d_idx, g->n, s->d_koff + at, s->d_keys, s->d_meta + at);
        ctx->st.kernel_launches += 2;
        dfree(ctx, d_idx);
        at += g->n;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    u32 base = db->segs[first].base;
    for (size_t i = first; i < db->segs.size(); ++i) seqset_destroy(db->segs[i].set);
    db->segs.resize(first);
    db->segs.push_back(DbSegment{s.release(), base});
    db->dirty = true;
    return KR_OK;
}

static kr_status db_append_owned(kr_ctx *ctx, kr_db *db, kr_seqset *s) {
    if (s->n == 0) { seqset_destroy(s); return KR_OK; }
    db->segs.push_back(DbSegment{s, db->n});
    db->n += s->n;
    if (!db->h_rank.empty()) db->h_rank.clear();   // ranks must be supplied again after growth
    db->dirty = true;
    return db_compact(ctx, db);
}

// can the representatives of this database be indexed (every segment keeps key lists; K and size fit the entry layout)?
static bool db_indexable(const kr_ctx *ctx, const kr_db *db) {
    u64 cap = 0;
    for (auto &sg : db->segs) { if (sg.set->n && !sg.set->d_keys) return false; cap += sg.set->key_cap; }
    IxGeom g;
    return index_geometry(ctx, std::max<u32>(2 * db->n, 1024), cap, &g);
}

// K4 for a database: one index level over representatives [lo, hi)
static kr_status db_build_level(kr_ctx *ctx, kr_db *db, u32 lo, u32 hi, IndexOwned **level) {
    std::vector<IxSegment> segs;
    u64 cap = 0;
    for (auto &sg : db->segs) {
        const u32 a = std::max(lo, sg.base), b = std::min(hi, sg.base + sg.set->n);
        if (a >= b) continue;
        kr_seqset *s = sg.set; const u32 l = a - sg.base;
        segs.push_back(IxSegment{s->d_keys, s->d_koff + l, s->d_meta + l, b - a, a});
        if (l == 0 && b - a == s->n) cap += s->key_cap;
        else { kr_status st = fetch_sizes(ctx, s); if (st) return st; for (u32 k = l; k < l + (b - a); ++k) cap += s->h_sizes[k]; }
    }
    if (!*level) *level = new IndexOwned();
    // ids of every level share one bit width (that of the largest id the database may reach before the next rebuild)
    return index_build(ctx, segs, cap, lo, hi, std::max<u32>(2 * db->n, 1024), *level);
}

// Bring the index up to date with the representatives appended since it was last built: the tail level is rebuilt over
// everything after the main level, and folded into the main level once it holds a quarter of it.
static kr_status db_ensure_index(kr_ctx *ctx, kr_db *db) {
    kr_status st;
    // an id must still fit the bit width the levels were built with
    if (db->ix_main && db->ix_main->ok && (db->n > (1ull << db->ix_main->g.idbits))) db_free_index(db);
    const u32 tail = db->n - db->main_n;
    if (!db->ix_main || !db->ix_main->ok || db->main_n == 0 || (u64)tail * 4 >= db->main_n) {
        if (db->ix_tail) { index_free(ctx, db->ix_tail); delete db->ix_tail; db->ix_tail = nullptr; }
        if ((st = db_build_level(ctx, db, 0, db->n, &db->ix_main))) return st;
        db->main_n = db->tail_n = db->n;
    } else if (db->tail_n != db->n) {
        if ((st = db_build_level(ctx, db, db->main_n, db->n, &db->ix_tail))) return st;
        db->tail_n = db->n;
    }
    // the k-mer-less representatives (equal-sequence rule for k-mer-less queries)
    if ((ctx->flags & KR_IDENT_IS_INFINITY) && db->empty_n != db->n) {
        dfree(ctx, db->d_empty_reps); dfree(ctx, db->d_nempty_reps);
        CU(dalloc(ctx, &db->d_empty_reps, (size_t)db->n + 1)); CU(dalloc(ctx, &db->d_nempty_reps, 1));
        CU(cudaMemsetAsync(db->d_nempty_reps, 0, 4, ctx->stream));
        KR_LAUNCH(KR_T(k_collect_empty), (db->n + 255) / 256, 256, 0, ctx->stream, db->d_meta, db->n, db->d_empty_reps, db->d_nempty_reps);
        LAUNCHED(ctx);
        CU(cudaGetLastError());
        db->empty_n = db->n;
    }
    return KR_OK;
}

// Which path scans this database?  The index, unless the context pins the join, the sets cannot be indexed, or the
// database's posting lists are long (dense similarity: most representatives share k-mers with most queries), where the
// per-warp score table would overflow into passes and the tile join's similarity-independent work wins.
static kr_status db_use_index(kr_ctx *ctx, kr_db *db, const kr_seqset *Q, bool *use) {
    *use = false;
    if (!Q->d_keys || (ctx->flags & KR_FORCE_JOIN) || !db_indexable(ctx, db)) return KR_OK;
    kr_status st;
    if ((st = db_refresh(ctx, db)) || (st = db_ensure_index(ctx, db)) || (st = index_stats(ctx, db->ix_main))) return st;
    *use = (ctx->flags & KR_FORCE_INDEX) || index_list_length(*db->ix_main) <= IX_DENSE_LISTS;
    return KR_OK;
}

// K5: queries [q0, q1) of Q against the index levels of the database; outputs are indexed by q - q0
static kr_status find_closest_index(kr_ctx *ctx, kr_db *db, kr_seqset *Q, u32 q0, u32 q1, int *d_sim, u32 *d_first, u32 *d_last) {
    kr_status st = db_ensure_index(ctx, db);
    if (st) return st;
    const u32 nq = q1 - q0;
    const bool two = db->ix_tail && db->tail_n > db->main_n;
    const u32 nlev = two ? 2u : 1u;
    int4 *d_partial = nullptr; u32 *d_next = nullptr;
    CU(dalloc(ctx, &d_partial, (size_t)nlev * nq)); CU(dalloc(ctx, &d_next, 2));
    CU(cudaMemsetAsync(d_next, 0, 8, ctx->stream));
    {
        StageTimer t(ctx, ST_SCAN);
        if ((st = index_argmax(ctx, *db->ix_main, Q, q0, q1, db->d_meta, db->d_rank, d_partial, d_next))) return st;
        if (two && (st = index_argmax(ctx, *db->ix_tail, Q, q0, q1, db->d_meta, db->d_rank, d_partial + nq, d_next + 1))) return st;
        if (ctx->flags & KR_IDENT_IS_INFINITY) {
            KR_LAUNCH(KR_T(k_idx_empty_queries), (nq * 32 + 255) / 256, 256, 0, ctx->stream, Q->d_meta + q0, nq, db->d_meta, db->d_rank, db->d_empty_reps,
                                                                                  db->d_nempty_reps, d_partial);
            LAUNCHED(ctx);
        }
    }
    {
        StageTimer t(ctx, ST_OTHER);
        KR_LAUNCH(KR_T(k_idx_merge), (nq + 255) / 256, 256, 0, ctx->stream, d_partial, nq, nlev, db->d_order, d_sim, d_first, d_last);
        LAUNCHED(ctx);
    }
    CU(cudaGetLastError());
    dfree(ctx, d_partial); dfree(ctx, d_next);
    ctx->st.pairs += (u64)nq * db->n;
    return KR_OK;
}

// queries x database: per-query best similarity with first/last rank (device outputs)
static kr_status find_closest_launch(kr_ctx *ctx, kr_db *db, kr_seqset *Q, int *d_sim, u32 *d_first, u32 *d_last) {
    kr_status st;
    if ((st = db_refresh(ctx, db))) return st;
    if (Q->n == 0) return KR_OK;
    if (db->n == 0) {
        CU(cudaMemsetAsync(d_sim, 0, (size_t)Q->n * 4, ctx->stream));
        if (d_first) CU(cudaMemsetAsync(d_first, 0xFF, (size_t)Q->n * 4, ctx->stream));
        if (d_last) CU(cudaMemsetAsync(d_last, 0xFF, (size_t)Q->n * 4, ctx->stream));
        return KR_OK;
    }
    // Two exact paths.  The index scan (work ~ postings the queries touch) serves every batch whose sets keep key lists
    // and whose K fits the 64-bit entry; the tile join (work ~ queries x representatives) serves the rest: sets
    // imported from block lists, K > 9 with large databases, and contexts that pin it (KR_FORCE_JOIN).
    bool use_index = false;
    if ((st = db_use_index(ctx, db, Q, &use_index))) return st;
    if (use_index) return find_closest_index(ctx, db, Q, 0, Q->n, d_sim, d_first, d_last);
    if ((st = db_refresh_refs(ctx, db))) return st;
    if ((st = ensure_blocklists(ctx, Q))) return st;
    u32 nI = Q->nblk;
    u32 nsplit = std::max<u32>(1, std::min<u32>(db->nblk, (2 * (u32)ctx->num_sms + nI - 1) / nI));
    std::vector<uint2> tasks;
    for (u32 I = 0; I < nI; ++I) for (u32 sp = 0; sp < nsplit; ++sp) tasks.push_back(make_uint2(I, sp));
    uint2 *d_tasks = nullptr; int4 *d_partial = nullptr;
    CU(dalloc(ctx, &d_tasks, tasks.size())); CU(dalloc(ctx, &d_partial, tasks.size() * TB));
    CU(cudaMemcpyAsync(d_tasks, tasks.data(), tasks.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
    JoinParams p{};
    p.A = Q->d_refs; p.B = db->d_refs; p.metaA = Q->d_meta; p.metaB = db->d_meta;
    p.nA = Q->n; p.nB = db->n; p.nBblk = db->nblk;
    p.tasks = d_tasks; p.ntasks = (u32)tasks.size(); p.flags = ctx->flags; p.err = ctx->d_err; ctx->err_pending = true;
    p.nsplit = nsplit; p.rankB = db->d_rank; p.partial = d_partial;
    {
        StageTimer t(ctx, ST_JOIN);
        u32 grid = std::min<u32>(p.ntasks, (u32)ctx->num_sms);
        if (lo_bits(ctx->K) <= 16) KR_LAUNCH(KR_T(k_join<MODE_ARGMAX, false>), grid, JOIN_THREADS, join_smem_bytes(false), ctx->stream, p);
        else KR_LAUNCH(KR_T(k_join<MODE_ARGMAX, true>), grid, JOIN_THREADS, join_smem_bytes(true), ctx->stream, p);
        LAUNCHED(ctx); ctx->st.join_launches++;
    }
    {
        StageTimer t(ctx, ST_OTHER);
        KR_LAUNCH(KR_T(k_argmax_merge), (Q->n + 255) / 256, 256, 0, ctx->stream, d_partial, Q->n, nsplit, db->d_order, d_sim, d_first, d_last);
        LAUNCHED(ctx);
    }
    CU(cudaGetLastError());
    dfree(ctx, d_tasks); dfree(ctx, d_partial);
    ctx->st.pairs += (u64)Q->n * db->n;
    return KR_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" {

kr_status kr_encode(kr_ctx *ctx, const uint8_t *residues, const uint64_t *offsets, uint32_t nseq, kr_seqset **out) {
    return encode_impl(ctx, residues, false, offsets, nseq, out);
}
kr_status kr_encode_device(kr_ctx *ctx, const uint8_t *d_residues, const uint64_t *offsets, uint32_t nseq, kr_seqset **out) {
    return encode_impl(ctx, d_residues, true, offsets, nseq, out);
}
void kr_seqset_free(kr_seqset *set) {
    if (!set) return;
    std::lock_guard<std::mutex> lk(set->ctx->mu);
    seqset_destroy(set);
}
uint32_t kr_seqset_count(const kr_seqset *set) { return set ? set->n : 0; }

kr_status kr_seqset_sizes(const kr_seqset *set, uint32_t *sizes) {
    if (!set || !sizes) return KR_ERR_ARG;
    kr_ctx *ctx = set->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    kr_status st = fetch_sizes(ctx, const_cast<kr_seqset *>(set));
    if (st) return st;
    memcpy(sizes, set->h_sizes.data(), (size_t)set->n * 4);
    return KR_OK;
}

kr_status kr_seqset_keys(const kr_seqset *set, uint32_t i, uint64_t *keys, uint32_t cap, uint32_t *n) {
    if (!set || !n) return KR_ERR_ARG;
    kr_ctx *ctx = set->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    if (i >= set->n) return fail(ctx, KR_ERR_ARG, "kr_seqset_keys: index %u of %u", i, set->n);
    if (!set->d_keys) return fail(ctx, KR_ERR_STATE, "kr_seqset_keys: a set imported from block lists has no key lists");
    kr_status st = fetch_sizes(ctx, const_cast<kr_seqset *>(set));
    if (st) return st;
    u32 sz = set->h_sizes[i];
    *n = sz;
    if (!keys || cap < sz) return sz ? fail(ctx, KR_ERR_ARG, "kr_seqset_keys: need room for %u keys", sz) : KR_OK;
    u64 ko;
    CU(cudaMemcpyAsync(&ko, set->d_koff + i, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpyAsync(keys, set->d_keys + ko, (size_t)sz * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KR_OK;
}

kr_status kr_seqset_gather(kr_ctx *ctx, const kr_seqset *set, const uint32_t *idx, uint32_t n, kr_seqset **out) {
    if (!ctx || !set || !out || (!idx && n)) return fail(ctx, KR_ERR_ARG, "kr_seqset_gather: null argument");
    if (set->ctx != ctx) return fail(ctx, KR_ERR_STATE, "kr_seqset_gather: set belongs to another context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    return gather_impl(ctx, set, idx, n, out);
}

// SequenceKmers.retainAll / removeAll (TargetProcessor.java:166,184) leave an object whose k-mer set is no longer the set of
// its sequence's substrings: the shim computes the new key list on the host (a merge of two sorted lists it fetched with
// kr_seqset_keys) and hands it back here, so that size / similarity / distance / findClosest of such objects still run on
// the GPU like any other set.  The sequences themselves (optional) keep the equal-sequence rule working as upstream, which
// compares the unchanged `sequence` strings.
kr_status kr_seqset_from_keys(kr_ctx *ctx, const uint64_t *keys, const uint64_t *key_offsets, const uint8_t *residues, const uint64_t *res_offsets,
                              uint32_t nseq, kr_seqset **out) {
    if (!ctx || !out || (!key_offsets && nseq)) return fail(ctx, KR_ERR_ARG, "kr_seqset_from_keys: null argument");
    *out = nullptr;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    const u64 kc = nseq ? key_offsets[nseq] - key_offsets[0] : 0;
    if (kc && !keys) return fail(ctx, KR_ERR_ARG, "kr_seqset_from_keys: null keys");
    const u64 limit = 5 * ctx->K >= 64 ? ~0ull : (1ull << (5 * ctx->K));
    std::vector<u64> koff(nseq + 1), off(nseq + 1, 0);
    std::vector<u32> sizes(nseq);
    for (u32 i = 0; i < nseq; ++i) {
        if (key_offsets[i + 1] < key_offsets[i]) return fail(ctx, KR_ERR_ARG, "kr_seqset_from_keys: key offsets not monotone at %u", i);
        const u64 n = key_offsets[i + 1] - key_offsets[i];
        if (n > 65535u) return fail(ctx, KR_ERR_ARG, "kr_seqset_from_keys: sequence %u has %llu keys; at most 65535 per sequence are supported", i, (unsigned long long)n);
        const u64 *kp = keys + key_offsets[i];
        for (u64 t = 0; t < n; ++t)
            if ((limit != ~0ull && kp[t] >= limit) || (t && kp[t] <= kp[t - 1]))
                return fail(ctx, KR_ERR_ARG, "kr_seqset_from_keys: keys of sequence %u are not strictly ascending %d-mers (position %llu)", i, ctx->K, (unsigned long long)t);
        koff[i] = key_offsets[i] - key_offsets[0]; sizes[i] = (u32)n;
    }
    koff[nseq] = kc;
    u64 total = 0;
    if (res_offsets) {
        for (u32 i = 0; i < nseq; ++i) {
            if (res_offsets[i + 1] < res_offsets[i]) return fail(ctx, KR_ERR_ARG, "kr_seqset_from_keys: residue offsets not monotone at %u", i);
            off[i] = res_offsets[i] - res_offsets[0];
        }
        total = nseq ? res_offsets[nseq] - res_offsets[0] : 0;
        off[nseq] = total;
        if (total && !residues) return fail(ctx, KR_ERR_ARG, "kr_seqset_from_keys: null residues");
        for (u64 t = 0; t < total; ++t) {
            const u8 c = residues[res_offsets[0] + t];
            if (!((c >= 'A' && c <= 'Z') || c == '*' || c == '-'))
                return fail(ctx, KR_ERR_ALPHABET, "kr_seqset_from_keys: residue 0x%02x is outside 'A'..'Z', '*', '-'", c);
        }
    }
    std::unique_ptr<kr_seqset, void (*)(kr_seqset *)> s(new kr_seqset(), seqset_destroy);
    s->ctx = ctx; s->n = nseq; s->key_cap = kc; s->h_sizes = sizes;
    s->res = std::make_shared<DevMem>();
    s->res->stream = ctx->stream;
    CU(pool_alloc(ctx, &s->res->p, total + 16));
    if (total) { CU(cudaMemcpyAsync(s->res->p, residues + res_offsets[0], total, cudaMemcpyHostToDevice, ctx->stream)); ctx->st.h2d_bytes += total; }
    CU(dalloc(ctx, &s->d_off, nseq + 1)); CU(dalloc(ctx, &s->d_koff, nseq + 1));
    CU(dalloc(ctx, &s->d_keys, kc + 16)); CU(dalloc(ctx, &s->d_meta, nseq));
    CU(cudaMemcpyAsync(s->d_off, off.data(), (size_t)(nseq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(s->d_koff, koff.data(), (size_t)(nseq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (kc) CU(cudaMemcpyAsync(s->d_keys, keys + key_offsets[0], (size_t)kc * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->st.h2d_bytes += kc * 8 + 2ull * (nseq + 1) * 8;
    if (nseq) {
        StageTimer t(ctx, ST_OTHER);
        KR_LAUNCH(KR_T(k_meta_from_keys), (nseq + 7) / 8, 256, 0, ctx->stream, s->d_koff, res_offsets ? (const u8 *)s->res->p : (const u8 *)nullptr, s->d_off, nseq,
                  ctx->keyset_nonce, s->d_meta);
        LAUNCHED(ctx);
        CU(cudaGetLastError());
        ctx->keyset_nonce += nseq;
    }
    CU(cudaStreamSynchronize(ctx->stream));     // the host vectors above were staged by then
    *out = s.release();
    return KR_OK;
}

uint64_t kr_blocklist_capacity(const kr_ctx *ctx, const uint64_t *offsets, uint32_t nseq) {
    if (!ctx || (!offsets && nseq)) return 0;
    u64 kc = 0;
    for (u32 i = 0; i < nseq; ++i) kc += kmer_slots(ctx, offsets[i + 1] - offsets[i]);
    return blocklist_capacity(kc, (nseq + TB - 1) / TB);
}

kr_status kr_seqset_export_blocklists(kr_ctx *ctx, kr_seqset *set, kr_blocklists *out) {
    if (!ctx || !set || !out) return fail(ctx, KR_ERR_ARG, "kr_seqset_export_blocklists: null argument");
    if (set->ctx != ctx) return fail(ctx, KR_ERR_STATE, "kr_seqset_export_blocklists: set belongs to another context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    kr_status st = ensure_blocklists(ctx, set);
    if (st) return st;
    u64 total = 0;
    if (set->nblk) CU(cudaMemcpyAsync(&total, set->d_blkbase + set->nblk, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->err_pending && (st = check_alphabet(ctx))) return st;
    out->nseq = set->n; out->nblk = set->nblk; out->entries = total;
    out->lo_size = lo_bits(ctx->K) <= 16 ? 2 : 4; out->nchunk = NCHUNK;
    out->hi = set->d_hi; out->lo = set->d_lo; out->chunk_off = set->d_choff; out->chunk_cnt = set->d_chcnt;
    out->blk_base = set->d_blkbase; out->meta = set->d_meta;
    return KR_OK;
}

kr_status kr_seqset_import_blocklists(kr_ctx *ctx, const kr_blocklists *shards, uint32_t nshards, const uint8_t *d_residues,
                                      const uint64_t *offsets, uint32_t nseq, kr_seqset **out) {
    if (!ctx || !shards || !nshards || !out || !offsets) return fail(ctx, KR_ERR_ARG, "kr_seqset_import_blocklists: null argument");
    *out = nullptr;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    const u32 lo_size = lo_bits(ctx->K) <= 16 ? 2 : 4;
    u64 entries = 0; u32 nblk = 0, n = 0;
    for (u32 k = 0; k < nshards; ++k) {
        const kr_blocklists &sh = shards[k];
        if (sh.nchunk != NCHUNK || sh.lo_size != lo_size) return fail(ctx, KR_ERR_ARG, "import: shard %u was built with another chunking / K", k);
        if (k + 1 < nshards && sh.nseq % TB) return fail(ctx, KR_ERR_ARG, "import: shard %u has %u sequences; inner shards must hold multiples of %d", k, sh.nseq, TB);
        if (sh.nblk != (sh.nseq + TB - 1) / TB) return fail(ctx, KR_ERR_ARG, "import: shard %u block count", k);
        entries += (sh.entries + 63) / 64 * 64; nblk += sh.nblk; n += sh.nseq;   // every shard starts 256-byte aligned (TMA sources)
    }
    if (n != nseq) return fail(ctx, KR_ERR_ARG, "import: shards hold %u sequences, batch has %u", n, nseq);
    // The tile join feeds these arrays to cp.async.bulk, whose sources must be 16-byte aligned and in range: a shard that
    // breaks the layout is refused here (KR_ERR_ARG), before anything is built from it.
    {
        u32 *d_bad = nullptr;
        CU(dalloc(ctx, &d_bad, 2 * (size_t)nshards));
        CU(cudaMemsetAsync(d_bad, 0, 8 * (size_t)nshards, ctx->stream));
        for (u32 k = 0; k < nshards; ++k) {
            const kr_blocklists &sh = shards[k];
            if (sh.nseq == 0) continue;
            if (!sh.hi || !sh.lo || !sh.chunk_off || !sh.chunk_cnt || !sh.blk_base || !sh.meta) { dfree(ctx, d_bad); return fail(ctx, KR_ERR_ARG, "import: shard %u has a null array", k); }
            if (((reinterpret_cast<uintptr_t>(sh.hi) | reinterpret_cast<uintptr_t>(sh.lo)) & 15) || ((reinterpret_cast<uintptr_t>(sh.chunk_off) | reinterpret_cast<uintptr_t>(sh.chunk_cnt)) & 3)
                || (reinterpret_cast<uintptr_t>(sh.blk_base) & 7) || (reinterpret_cast<uintptr_t>(sh.meta) & 7)) {
                dfree(ctx, d_bad);
                return fail(ctx, KR_ERR_ARG, "import: shard %u: hi / lo must be 16-byte aligned device arrays (chunk tables 4-byte, block bases and records 8-byte)", k);
            }
            KR_LAUNCH(KR_T(k_validate_shard), (sh.nblk + 127) / 128, 128, 0, ctx->stream, sh.nblk, sh.entries, (const u32 *)sh.chunk_off, (const u32 *)sh.chunk_cnt,
                      (const u64 *)sh.blk_base, d_bad + 2 * k);
            LAUNCHED(ctx);
        }
        std::vector<u32> bad(2 * (size_t)nshards);
        CU(cudaMemcpyAsync(bad.data(), d_bad, 8 * (size_t)nshards, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        dfree(ctx, d_bad);
        for (u32 k = 0; k < nshards; ++k)
            if (bad[2 * k]) return fail(ctx, KR_ERR_ARG, "import: shard %u: %u block(s) break the block-list layout (first: block %u): chunk starts must be ascending multiples "
                                        "of %d entries inside the shard's arrays with %u spare entries after each block", k, bad[2 * k], bad[2 * k + 1], CHUNK_ALIGN, TAILPAD);
    }
    std::unique_ptr<kr_seqset, void (*)(kr_seqset *)> s(new kr_seqset(), seqset_destroy);
    s->ctx = ctx; s->n = nseq; s->nblk = nblk;
    const u64 total_res = offsets[nseq] - offsets[0];
    std::vector<u64> off(nseq + 1);
    for (u32 i = 0; i <= nseq; ++i) off[i] = offsets[i] - offsets[0];
    s->res = std::make_shared<DevMem>(); s->res->stream = ctx->stream;
    CU(pool_alloc(ctx, &s->res->p, total_res + 16));
    CU(cudaMemcpyAsync(s->res->p, d_residues + offsets[0], total_res, cudaMemcpyDeviceToDevice, ctx->stream));
    u8 *d_lo = nullptr;
    CU(dalloc(ctx, &s->d_off, nseq + 1)); CU(dalloc(ctx, &s->d_meta, nseq));
    CU(dalloc(ctx, &s->d_hi, entries + 16)); CU(dalloc(ctx, &d_lo, (entries + 16) * lo_size)); s->d_lo = d_lo;
    CU(dalloc(ctx, &s->d_choff, (size_t)nblk * (NCHUNK + 1))); CU(dalloc(ctx, &s->d_chcnt, (size_t)nblk * NCHUNK));
    CU(dalloc(ctx, &s->d_refs, nblk));
    CU(cudaMemcpyAsync(s->d_off, off.data(), (size_t)(nseq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    u64 e0 = 0; u32 b0 = 0, s0 = 0;
    for (u32 k = 0; k < nshards; ++k) {
        const kr_blocklists &sh = shards[k];
        if (sh.nseq == 0) continue;
        CU(cudaMemcpyAsync(s->d_hi + e0, sh.hi, sh.entries * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_lo + e0 * lo_size, sh.lo, sh.entries * lo_size, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(s->d_choff + (size_t)b0 * (NCHUNK + 1), sh.chunk_off, (size_t)sh.nblk * (NCHUNK + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(s->d_chcnt + (size_t)b0 * NCHUNK, sh.chunk_cnt, (size_t)sh.nblk * NCHUNK * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(s->d_meta + s0, sh.meta, (size_t)sh.nseq * sizeof(SeqMeta), cudaMemcpyDeviceToDevice, ctx->stream));
        KR_LAUNCH(KR_T(k_import_refs), (sh.nblk + 127) / 128, 128, 0, ctx->stream, sh.nblk, b0, s0, sh.nseq, s->d_hi, d_lo, lo_size, s->d_choff, s->d_chcnt,
                                                                     (const u64 *)sh.blk_base, e0, s->d_refs);
        LAUNCHED(ctx);
        e0 += (sh.entries + 63) / 64 * 64; b0 += sh.nblk; s0 += sh.nseq;
    }
    if (nseq) { KR_LAUNCH(KR_T(k_import_meta), (nseq + 255) / 256, 256, 0, ctx->stream, s->d_meta, nseq, (const u8 *)s->res->p, s->d_off); LAUNCHED(ctx); }
    CU(cudaGetLastError());
    s->bl_ready = true;
    *out = s.release();
    return KR_OK;
}

kr_status kr_sim_pairs(kr_ctx *ctx, const kr_seqset *A, const uint32_t *ia, const kr_seqset *B, const uint32_t *ib,
                       uint64_t npairs, int32_t *sim, uint8_t *identical) {
    if (!ctx || !A || !B || !sim || ((!ia || !ib) && npairs)) return fail(ctx, KR_ERR_ARG, "kr_sim_pairs: null argument");
    if (A->ctx != ctx || B->ctx != ctx) return fail(ctx, KR_ERR_STATE, "kr_sim_pairs: set belongs to another context");
    if ((!A->d_keys && A->n) || (!B->d_keys && B->n)) return fail(ctx, KR_ERR_STATE, "kr_sim_pairs: a set imported from block lists has no key lists");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    if (!npairs) return KR_OK;
    for (u64 p = 0; p < npairs; ++p)
        if (ia[p] >= A->n || ib[p] >= B->n) return fail(ctx, KR_ERR_ARG, "kr_sim_pairs: pair %llu index out of range", (unsigned long long)p);
    u32 *d_ia, *d_ib; int *d_sim; u8 *d_id;
    CU(dalloc(ctx, &d_ia, npairs)); CU(dalloc(ctx, &d_ib, npairs)); CU(dalloc(ctx, &d_sim, npairs)); CU(dalloc(ctx, &d_id, npairs));
    CU(cudaMemcpyAsync(d_ia, ia, npairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_ib, ib, npairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    {
        StageTimer t(ctx, ST_OTHER);
        u32 grid = (u32)std::min<u64>((npairs + 7) / 8, (u64)ctx->num_sms * 8);
        KR_LAUNCH(KR_T(k_sim_pairs), grid, 256, 0, ctx->stream, A->d_keys, A->d_koff, A->d_meta, d_ia, B->d_keys, B->d_koff, B->d_meta, d_ib, npairs,
                                                   ctx->flags, d_sim, d_id);
        LAUNCHED(ctx);
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(sim, d_sim, npairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (identical) CU(cudaMemcpyAsync(identical, d_id, npairs, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    dfree(ctx, d_ia); dfree(ctx, d_ib); dfree(ctx, d_sim); dfree(ctx, d_id);
    ctx->st.pairs += npairs; ctx->st.h2d_bytes += npairs * 8; ctx->st.d2h_bytes += npairs * 5;
    return KR_OK;
}

uint64_t kr_matrix_elems(uint32_t n_rows, uint32_t n_cols, uint32_t layout, uint32_t row_begin, uint32_t row_end) {
    if (row_end > n_rows) row_end = n_rows;
    if (row_begin >= row_end) return 0;
    return row_start(layout, row_end, n_rows, n_cols) - row_start(layout, row_begin, n_rows, n_cols);
}

kr_status kr_matrix_stripes(uint32_t n_rows, uint32_t n_cols, uint32_t layout, uint32_t nshards, uint32_t *row_bounds) {
    if (!row_bounds || nshards == 0 || layout > KR_LAYOUT_BLOCKTRI) return KR_ERR_ARG;
    u64 total = row_start(layout, n_rows, n_rows, n_cols);
    u32 nblk = (n_rows + TB - 1) / TB;
    row_bounds[0] = 0;
    u32 b = 0;
    for (u32 s = 1; s < nshards; ++s) {
        u64 want = total / nshards * s;
        // first block boundary whose cumulative cell count reaches the target, choosing the closer side
        while (b < nblk && row_start(layout, std::min<u64>((u64)(b + 1) * TB, n_rows), n_rows, n_cols) <= want) ++b;
        u64 lo = row_start(layout, std::min<u64>((u64)b * TB, n_rows), n_rows, n_cols);
        u64 hi = row_start(layout, std::min<u64>((u64)(b + 1) * TB, n_rows), n_rows, n_cols);
        u32 pick = (b < nblk && hi - want < want - lo) ? b + 1 : b;
        row_bounds[s] = (u32)std::min<u64>((u64)pick * TB, n_rows);
        if (row_bounds[s] < row_bounds[s - 1]) row_bounds[s] = row_bounds[s - 1];
    }
    row_bounds[nshards] = n_rows;
    return KR_OK;
}

kr_status kr_sim_matrix_device(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, void *out_dev) {
    if (!ctx) return KR_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    kr_status st = check_desc(ctx, A, B, desc);
    if (st) return st;
    if (!out_dev) return fail(ctx, KR_ERR_ARG, "kr_sim_matrix_device: null output");
    // a misaligned store is a device fault, not an error code: refuse the pointer here
    if (reinterpret_cast<uintptr_t>(out_dev) & (desc->elem == KR_ELEM_U16 ? 1 : 3)) return fail(ctx, KR_ERR_ARG, "kr_sim_matrix_device: output is not aligned to its element size");
    CU(cudaSetDevice(ctx->dev));
    bool self = (B == nullptr || B == A);
    return matrix_launch(ctx, const_cast<kr_seqset *>(A), const_cast<kr_seqset *>(B), self, desc, out_dev);
}

kr_status kr_sim_matrix(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, void *out_host) {
    if (!ctx) return KR_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    kr_status st = check_desc(ctx, A, B, desc);
    if (st) return st;
    if (!out_host) return fail(ctx, KR_ERR_ARG, "kr_sim_matrix: null output");
    CU(cudaSetDevice(ctx->dev));
    const bool self = (B == nullptr || B == A);
    const u32 ncols = self ? A->n : B->n;
    const size_t esz = desc->elem == KR_ELEM_U16 ? 2 : 4;
    if (!kr_matrix_elems(A->n, ncols, desc->layout, desc->row_begin, desc->row_end)) return KR_OK;
    // Row stripes of a few hundred MB are computed into two device buffers in turn while the previous stripe
    // travels to the host on the copy stream, so the D2H of a large matrix overlaps the join.
    const u64 target = 256ull << 20;
    u64 per_blockrow = (u64)TB * ncols * esz;
    u32 stripe_rows = (u32)std::max<u64>(1, target / std::max<u64>(per_blockrow, 1)) * TB;
    u64 max_elems = 0;
    for (u32 r = desc->row_begin; r < desc->row_end; r += stripe_rows)
        max_elems = std::max<u64>(max_elems, kr_matrix_elems(A->n, ncols, desc->layout, r, std::min(desc->row_end, r + stripe_rows)));
    u8 *d_buf[2] = {nullptr, nullptr};
    CU(dalloc(ctx, &d_buf[0], max_elems * esz + 16));
    CU(dalloc(ctx, &d_buf[1], max_elems * esz + 16));
    CU(cudaStreamSynchronize(ctx->stream));   // the buffers exist before the copy stream touches them
    u64 done_elems = 0;
    int k = 0;
    for (u32 r = desc->row_begin; r < desc->row_end && st == KR_OK; r += stripe_rows, ++k) {
        const int b = k & 1;
        kr_matrix_desc sub = *desc;
        sub.row_begin = r; sub.row_end = std::min(desc->row_end, r + stripe_rows);
        const u64 elems = kr_matrix_elems(A->n, ncols, desc->layout, sub.row_begin, sub.row_end);
        if (k >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));   // buffer b has left for the host
        st = matrix_launch(ctx, const_cast<kr_seqset *>(A), const_cast<kr_seqset *>(B), self, &sub, d_buf[b]);
        if (st) break;
        CU(cudaEventRecord(ctx->ev_done[b], ctx->stream));
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
        CU(cudaMemcpyAsync((u8 *)out_host + done_elems * esz, d_buf[b], elems * esz, cudaMemcpyDeviceToHost, ctx->copy_stream));
        CU(cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
        done_elems += elems;
        ctx->st.d2h_bytes += elems * esz;
    }
    cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess && st == KR_OK) st = fail(ctx, KR_ERR_CUDA, "matrix copy to host failed: %s", cudaGetErrorString(e));
    dfree(ctx, d_buf[0]); dfree(ctx, d_buf[1]);
    if (st == KR_OK && ctx->err_pending) st = check_alphabet(ctx);
    return st;
}


// the caller holds the context's lock
static kr_status sparse_impl(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, int32_t min_sim,
                             void *cells_host, uint64_t cap, uint64_t *count, void *cells_dev, bool packed = false) {
    kr_status st = check_desc(ctx, A, B, desc);
    if (st) return st;
    const size_t csz = packed ? 8 : 16;
    if (packed && (A->n > (1u << 24) || (B ? B->n : A->n) > (1u << 24)))
        return fail(ctx, KR_ERR_ARG, "kr_sim_matrix_sparse64: packed cells hold row and column numbers below 2^24");
    if (!count || (!cells_host && !cells_dev && cap) || min_sim < 1 || desc->layout == KR_LAYOUT_BLOCKTRI)
        return fail(ctx, KR_ERR_ARG, "kr_sim_matrix_sparse: bad argument (min_sim >= 1; layout FULL or TRIANGLE)");
    CU(cudaSetDevice(ctx->dev));
    *count = 0;
    const bool self = (B == nullptr || B == A);
    uint4 *d_cells = reinterpret_cast<uint4 *>(cells_dev); unsigned long long *d_count = nullptr;
    if (!cells_dev) { u8 *raw = nullptr; CU(dalloc(ctx, &raw, cap * csz + 16)); d_cells = reinterpret_cast<uint4 *>(raw); }
    CU(dalloc(ctx, &d_count, 1));
    CU(cudaMemsetAsync(d_count, 0, 8, ctx->stream));
    SparseSink sink{d_cells, cap, d_count, (u32)min_sim, packed};
    st = matrix_launch(ctx, const_cast<kr_seqset *>(A), const_cast<kr_seqset *>(B), self, desc, d_cells, &sink);
    if (st == KR_OK) {
        unsigned long long n = 0;
        cudaError_t e = cudaMemcpyAsync(&n, d_count, 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        const unsigned long long take = std::min<unsigned long long>(n, cap);
        if (e == cudaSuccess && take && cells_host) e = cudaMemcpyAsync(cells_host, d_cells, take * csz, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = fail(ctx, KR_ERR_CUDA, "sparse matrix copy to host failed: %s", cudaGetErrorString(e));
        ctx->st.d2h_bytes += 8 + (cells_host ? take * csz : 0);
        *count = n;
        if (st == KR_OK && ctx->err_pending) st = check_alphabet(ctx);
        if (st == KR_OK && n > cap) st = fail(ctx, KR_ERR_ARG, "kr_sim_matrix_sparse: %llu cells qualify, room for %llu", n, (unsigned long long)cap);
    }
    if (!cells_dev) dfree(ctx, reinterpret_cast<u8 *>(d_cells));
    dfree(ctx, d_count);
    return st;
}

kr_status kr_sim_matrix_sparse(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, int32_t min_sim,
                               uint32_t *cells_host, uint64_t cap, uint64_t *count) {
    if (!ctx) return KR_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return sparse_impl(ctx, A, B, desc, min_sim, cells_host, cap, count, nullptr);
}
kr_status kr_sim_matrix_sparse64(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, int32_t min_sim,
                                 uint64_t *cells_host, uint64_t cap, uint64_t *count) {
    if (!ctx) return KR_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return sparse_impl(ctx, A, B, desc, min_sim, cells_host, cap, count, nullptr, true);
}
kr_status kr_sim_matrix_sparse64_device(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, int32_t min_sim,
                                        void *cells_dev, uint64_t cap, uint64_t *count) {
    if (!ctx) return KR_ERR_ARG;
    if (!cells_dev && cap) return fail(ctx, KR_ERR_ARG, "kr_sim_matrix_sparse64_device: null output");
    if (reinterpret_cast<uintptr_t>(cells_dev) & 7) return fail(ctx, KR_ERR_ARG, "kr_sim_matrix_sparse64_device: cells must be 8-byte aligned");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return sparse_impl(ctx, A, B, desc, min_sim, nullptr, cap, count, cells_dev, true);
}
kr_status kr_sim_matrix_sparse_device(kr_ctx *ctx, const kr_seqset *A, const kr_seqset *B, const kr_matrix_desc *desc, int32_t min_sim,
                                      void *cells_dev, uint64_t cap, uint64_t *count) {
    if (!ctx) return KR_ERR_ARG;
    if (!cells_dev && cap) return fail(ctx, KR_ERR_ARG, "kr_sim_matrix_sparse_device: null output");
    if (reinterpret_cast<uintptr_t>(cells_dev) & 15) return fail(ctx, KR_ERR_ARG, "kr_sim_matrix_sparse_device: cells must be 16-byte aligned (one 16-byte store per cell)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return sparse_impl(ctx, A, B, desc, min_sim, nullptr, cap, count, cells_dev);
}

// ---- database ----------------------------------------------------------------------------------
kr_status kr_db_create(kr_ctx *ctx, kr_db **out) {
    if (!ctx || !out) return KR_ERR_ARG;
    kr_db *db = new kr_db(); db->ctx = ctx; *out = db;
    return KR_OK;
}
void kr_db_free(kr_db *db) {
    if (!db) return;
    kr_ctx *ctx = db->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->dev);
    for (auto &sg : db->segs) seqset_destroy(sg.set);
    db_free_arrays(db);
    db_free_index(db);
    delete db;
}
uint32_t kr_db_size(const kr_db *db) { return db ? db->n : 0; }

kr_status kr_db_append(kr_ctx *ctx, kr_db *db, const uint8_t *residues, const uint64_t *offsets, uint32_t nseq) {
    if (!ctx || !db) return KR_ERR_ARG;
    if (db->ctx != ctx) return fail(ctx, KR_ERR_STATE, "kr_db_append: database belongs to another context");
    kr_seqset *s = nullptr;
    kr_status st = encode_impl(ctx, residues, false, offsets, nseq, &s);
    if (st) return st;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    return db_append_owned(ctx, db, s);
}

kr_status kr_db_append_set(kr_ctx *ctx, kr_db *db, const kr_seqset *set, const uint32_t *idx, uint32_t n) {
    if (!ctx || !db || !set || (!idx && n)) return fail(ctx, KR_ERR_ARG, "kr_db_append_set: null argument");
    if (db->ctx != ctx || set->ctx != ctx) return fail(ctx, KR_ERR_STATE, "kr_db_append_set: handle belongs to another context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    kr_seqset *s = nullptr;
    kr_status st = gather_impl(ctx, set, idx, n, &s);
    if (st) return st;
    return db_append_owned(ctx, db, s);
}

kr_status kr_db_set_ranks(kr_ctx *ctx, kr_db *db, const uint32_t *rank, uint32_t n) {
    if (!ctx || !db || (!rank && n)) return KR_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n != db->n) return fail(ctx, KR_ERR_ARG, "kr_db_set_ranks: %u ranks for %u representatives", n, db->n);
    std::vector<u8> seen(n, 0);
    for (u32 i = 0; i < n; ++i) {
        if (rank[i] >= n || seen[rank[i]]) return fail(ctx, KR_ERR_ARG, "kr_db_set_ranks: not a permutation at %u", i);
        seen[rank[i]] = 1;
    }
    db->h_rank.assign(rank, rank + n);
    db->dirty = true;
    return KR_OK;
}

kr_status kr_db_find_closest_device(kr_ctx *ctx, const kr_db *db, const kr_seqset *queries, int32_t *d_best_sim,
                                    uint32_t *d_best_first, uint32_t *d_best_last) {
    if (!ctx || !db || !queries || !d_best_sim) return fail(ctx, KR_ERR_ARG, "kr_db_find_closest: null argument");
    if (db->ctx != ctx || queries->ctx != ctx) return fail(ctx, KR_ERR_STATE, "kr_db_find_closest: handle belongs to another context");
    if ((reinterpret_cast<uintptr_t>(d_best_sim) | reinterpret_cast<uintptr_t>(d_best_first) | reinterpret_cast<uintptr_t>(d_best_last)) & 3)
        return fail(ctx, KR_ERR_ARG, "kr_db_find_closest_device: outputs must be 4-byte aligned");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    return find_closest_launch(ctx, const_cast<kr_db *>(db), const_cast<kr_seqset *>(queries), d_best_sim, d_best_first, d_best_last);
}

kr_status kr_db_find_closest(kr_ctx *ctx, const kr_db *db, const kr_seqset *queries, int32_t *best_sim, uint32_t *best_first,
                             uint32_t *best_last) {
    if (!ctx || !db || !queries || !best_sim) return fail(ctx, KR_ERR_ARG, "kr_db_find_closest: null argument");
    if (db->ctx != ctx || queries->ctx != ctx) return fail(ctx, KR_ERR_STATE, "kr_db_find_closest: handle belongs to another context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    u32 nq = queries->n;
    if (!nq) return KR_OK;
    int *d_sim; u32 *d_first, *d_last;
    CU(dalloc(ctx, &d_sim, nq)); CU(dalloc(ctx, &d_first, nq)); CU(dalloc(ctx, &d_last, nq));
    kr_status st = find_closest_launch(ctx, const_cast<kr_db *>(db), const_cast<kr_seqset *>(queries), d_sim, d_first, d_last);
    if (st == KR_OK) {
        cudaError_t e = cudaMemcpyAsync(best_sim, d_sim, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && best_first) e = cudaMemcpyAsync(best_first, d_first, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && best_last) e = cudaMemcpyAsync(best_last, d_last, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = fail(ctx, KR_ERR_CUDA, "find_closest copy to host failed: %s", cudaGetErrorString(e));
        ctx->st.d2h_bytes += (u64)nq * 12;
        if (st == KR_OK && ctx->err_pending) st = check_alphabet(ctx);
    }
    dfree(ctx, d_sim); dfree(ctx, d_first); dfree(ctx, d_last);
    return st;
}

kr_status kr_db_any_ge(kr_ctx *ctx, const kr_db *db, const kr_seqset *cands, int32_t threshold, uint8_t *represented) {
    if (!ctx || !db || !cands || !represented) return fail(ctx, KR_ERR_ARG, "kr_db_any_ge: null argument");
    std::vector<int32_t> sim(cands->n);
    kr_status st = kr_db_find_closest(ctx, db, cands, sim.data(), nullptr, nullptr);
    if (st) return st;
    for (u32 i = 0; i < cands->n; ++i) represented[i] = sim[i] >= threshold;
    return KR_OK;
}

}  // extern "C"

// ---- the greedy build, in three steps that the single-GPU and the multi-GPU entry points share (callers hold the
//      context's lock) --------------------------------------------------------------------------------------------
// 1. best similarity of candidates [pos0, pos1) of `cands` against the representatives so far -> best[0 .. pos1 - pos0)
static kr_status build_score_range(kr_ctx *ctx, kr_db *db, const kr_seqset *cands, u32 pos0, u32 pos1, int32_t *best) {
    const u32 n = pos1 - pos0;
    if (!n) return KR_OK;
    if (!db->n) { for (u32 i = 0; i < n; ++i) best[i] = 0; return KR_OK; }
    kr_status st = KR_OK;
    int *d_sim = nullptr;
    CU(dalloc(ctx, &d_sim, n));
    bool use_index = false;
    if (cands->d_keys && !(ctx->flags & KR_FORCE_JOIN)) st = db_use_index(ctx, db, cands, &use_index);
    if (st == KR_OK) {
        if (use_index) {
            if ((st = db_refresh(ctx, db)) == KR_OK) st = find_closest_index(ctx, db, const_cast<kr_seqset *>(cands), pos0, pos1, d_sim, nullptr, nullptr);
        } else {
            std::vector<u32> idx(n);
            for (u32 i = 0; i < n; ++i) idx[i] = pos0 + i;
            kr_seqset *bs = nullptr;
            if ((st = gather_impl(ctx, cands, idx.data(), n, &bs)) == KR_OK) {
                st = find_closest_launch(ctx, db, bs, d_sim, nullptr, nullptr);
                seqset_destroy(bs);      // stream-ordered release
            }
        }
    }
    if (st == KR_OK) {
        cudaError_t e = cudaMemcpyAsync(best, d_sim, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = fail(ctx, KR_ERR_CUDA, "copy: %s", cudaGetErrorString(e));
        ctx->st.d2h_bytes += (u64)n * 4;
        if (st == KR_OK && ctx->err_pending) st = check_alphabet(ctx);
    }
    dfree(ctx, d_sim);
    return st;
}

// 2. the in-order sweep of checkGenome over the survivors (candidates no current representative covers): surv holds
//    their positions in `cands`, ascending; acc receives the indices INTO surv of the accepted ones
static kr_status build_resolve(kr_ctx *ctx, const kr_seqset *cands, const std::vector<u32> &surv, int32_t threshold, std::vector<u32> *acc) {
    const u32 S = (u32)surv.size();
    kr_status st = KR_OK;
    acc->clear();
    if (S == 0) return KR_OK;
    if (S == 1) { acc->push_back(0); return KR_OK; }
    kr_seqset *ss = nullptr;
    if ((st = gather_impl(ctx, cands, surv.data(), S, &ss))) return st;
    if (threshold >= 1) {
        // survivors against one another as the sparse cells with sim >= threshold: a few cells per survivor
        kr_matrix_desc d{KR_ELEM_I32, KR_LAYOUT_TRIANGLE, 0, S};
        uint64_t room = 8 * (uint64_t)S + 1024, ncell = 0;
        std::vector<u32> cells((size_t)room * 4);
        st = sparse_impl(ctx, ss, nullptr, &d, threshold, cells.data(), room, &ncell, nullptr);
        if (st == KR_ERR_ARG && ncell > room) {      // more qualifying pairs than guessed: once more with room for all
            room = ncell; cells.resize((size_t)room * 4);
            st = sparse_impl(ctx, ss, nullptr, &d, threshold, cells.data(), room, &ncell, nullptr);
        }
        seqset_destroy(ss);
        if (st) return st;
        // earlier[s] = survivors t < s with similarity >= threshold (CSR over the cells)
        std::vector<u32> start(S + 1, 0), adj((size_t)ncell);
        for (uint64_t k = 0; k < ncell; ++k) start[cells[4 * k + 1] + 1]++;
        for (u32 v = 0; v < S; ++v) start[v + 1] += start[v];
        std::vector<u32> fill(start.begin(), start.end() - 1);
        for (uint64_t k = 0; k < ncell; ++k) adj[fill[cells[4 * k + 1]]++] = cells[4 * k];
        std::vector<u8> accepted(S, 0);
        for (u32 v = 0; v < S; ++v) {
            bool repd = false;
            for (u32 k = start[v]; k < start[v + 1] && !repd; ++k) repd = accepted[adj[k]] != 0;
            if (!repd) { accepted[v] = 1; acc->push_back(v); }
        }
        return KR_OK;
    }
    // threshold <= 0: every similarity qualifies; the dense survivor matrix keeps the general code path honest
    kr_matrix_desc d{KR_ELEM_U16, KR_LAYOUT_FULL, 0, S};
    u16 *d_M = nullptr;
    cudaError_t e = dalloc(ctx, &d_M, (size_t)S * S + 8);
    if (e != cudaSuccess) { seqset_destroy(ss); return fail(ctx, KR_ERR_CUDA, "alloc: %s", cudaGetErrorString(e)); }
    st = matrix_launch(ctx, ss, nullptr, true, &d, d_M);
    std::vector<u16> M((size_t)S * S);
    if (st == KR_OK) {
        e = cudaMemcpyAsync(M.data(), d_M, (size_t)S * S * 2, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = fail(ctx, KR_ERR_CUDA, "copy: %s", cudaGetErrorString(e));
        ctx->st.d2h_bytes += (u64)S * S * 2;
        if (st == KR_OK && ctx->err_pending) st = check_alphabet(ctx);
    }
    dfree(ctx, d_M);
    seqset_destroy(ss);
    if (st) return st;
    for (u32 v = 0; v < S; ++v) {
        bool rep = false;
        const u16 *row = M.data() + (size_t)v * S;   // M is symmetric
        for (u32 t : *acc) if ((int)row[t] >= threshold) { rep = true; break; }
        if (!rep) acc->push_back(v);
    }
    return KR_OK;
}

// 3. append the accepted candidates (positions in `cands`) to the database
static kr_status build_accept(kr_ctx *ctx, kr_db *db, const kr_seqset *cands, const std::vector<u32> &acc_idx) {
    if (acc_idx.empty()) return KR_OK;
    kr_seqset *as = nullptr;
    kr_status st = gather_impl(ctx, cands, acc_idx.data(), (u32)acc_idx.size(), &as);
    if (st) return st;
    return db_append_owned(ctx, db, as);
}

static u32 build_default_batch(const kr_ctx *ctx, const kr_seqset *cands) {
    return (cands->d_keys && !(ctx->flags & KR_FORCE_JOIN)) ? 32768u : 4096u;
}

extern "C" {

kr_status kr_greedy_build(kr_ctx *ctx, kr_db *db, const kr_seqset *cands, int32_t threshold, uint32_t batch, uint8_t *added) {
    if (!ctx || !db || !cands || !added) return fail(ctx, KR_ERR_ARG, "kr_greedy_build: null argument");
    if (db->ctx != ctx || cands->ctx != ctx) return fail(ctx, KR_ERR_STATE, "kr_greedy_build: handle belongs to another context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->dev));
    // The accept / reject sweep of checkGenome is sequential; what is batched is the scoring.  Per batch of candidates:
    //   1. best similarity of every candidate against the representatives so far (index scan straight on the
    //      candidate set -- no gather -- or the tile join when the sets cannot be indexed);
    //   2. the candidates below the threshold ("survivors") against one another, as the sparse cells with
    //      sim >= threshold, and the sweep in candidate order over those adjacency lists on the host;
    //   3. accepted candidates are appended (the database's tail index is rebuilt by the next batch's scan).
    if (batch == 0) batch = build_default_batch(ctx, cands);
    const u32 C = cands->n;
    kr_status st = KR_OK;
    std::vector<u32> surv, acc, acc_idx;
    std::vector<int32_t> best;
    for (u32 pos = 0; pos < C && st == KR_OK; pos += batch) {
        const u32 n = std::min(batch, C - pos);
        best.resize(n);
        if ((st = build_score_range(ctx, db, cands, pos, pos + n, best.data()))) break;
        surv.clear();
        for (u32 i = 0; i < n; ++i) { added[pos + i] = 0; if (best[i] < threshold || !db->n) surv.push_back(pos + i); }
        if ((st = build_resolve(ctx, cands, surv, threshold, &acc))) break;
        acc_idx.resize(acc.size());
        for (size_t k = 0; k < acc.size(); ++k) { acc_idx[k] = surv[acc[k]]; added[acc_idx[k]] = 1; }
        st = build_accept(ctx, db, cands, acc_idx);
    }
    return st;
}

}  // extern "C"

#include "kr_api_multi.inc"
