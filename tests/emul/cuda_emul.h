// cuda_emul.h -- TEST INFRASTRUCTURE ONLY: a functional emulation of the small part of CUDA that
// kmers/reps_b200/csrc uses, so that the product's kernel and host sources can be compiled with g++
// (-DKR_EMUL) and EXECUTED on CPU threads by the tests in a container without a GPU.
//
// It is a simulator for checking logic (indexing, protocols, host orchestration), never a product path and never
// a fallback: only tests/ builds it (tests/emul/Makefile -> tests/emul/_build/libkmersreps_emul.so) and loads it
// through KR_LIB_PATH.  One CUDA thread = one fiber of the launching OS thread; CTAs of a launch run one after another;
// streams are synchronous.  What it models:
//   threadIdx/blockIdx/blockDim/gridDim, __shared__ (static and dynamic), __syncthreads, named barriers,
//   warp shuffles / votes (lock-step exchange through a per-warp buffer), global and shared atomics,
//   mbarrier (init / arrive / expect_tx / try_wait.parity) and 1-D bulk copies completing a transaction count
//   (with the alignment rules of cp.async.bulk CHECKED: 16-byte aligned source, destination and size),
//   cudaMalloc/Memcpy/Memset/streams/events as plain host operations.
// What it cannot show: performance, register pressure, hardware memory ordering beyond the happens-before model below.
//
// RACE DETECTION (-DKR_EMUL_TSAN, built with g++ -fsanitize=thread --param tsan-distinguish-volatile=1; tools/emul_race.sh):
// every CUDA thread of a CTA is a ThreadSanitizer fiber, switched WITHOUT synchronisation, and the only happens-before
// edges are the ones the hardware gives: __syncthreads / named barriers / __syncwarp (all-to-all among the arrivers),
// mbarrier arrive -> successful try_wait of that phase, CTA start and end.  Shuffles and votes order nothing; atomics are
// relaxed; volatile accesses count as relaxed atomics.  Two threads of a CTA touching the same bytes of shared OR global
// memory, one of them writing, with no such edge in between, is reported as a data race -- what compute-sanitizer's
// racecheck reports for shared memory on hardware, extended to global memory.  CTAs of a launch are ordered one after
// another in this mode.  KR_EMUL_RACE=grid drops that order instead: the CTAs of a launch are concurrent, as on the
// device, and a global-memory word that two CTAs touch (one writing, no atomics) is reported.  (Shared memory, which the
// emulated CTAs re-use, gets a fresh history per CTA in both modes.)
// STREAM ORDER is part of the same model: the emulator still executes every operation at once, in host program order,
// but each stream carries a clock of its own.  A kernel, an async copy / memset, a stream-ordered allocation or free is
// attributed to its stream (ordered after the host at the moment of the call and after the stream's earlier work); the
// host is ordered after a stream only by cudaStreamSynchronize / cudaEventSynchronize / a blocking copy to pageable
// memory / cudaFree, and one stream after another only by cudaEventRecord + cudaStreamWaitEvent.  A host read of a
// result before the synchronisation, two streams touching one buffer without an event, a free ordered before the last
// use: each is a report, although the emulated run itself computed the right thing.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>
#include <ucontext.h>

// (runtime calls that act as a stream run on the caller's stack: their own locals are not the detector's business)
#define KR_NOTSAN __attribute__((no_sanitize("thread")))
#ifdef KR_EMUL_TSAN
extern "C" {
void *__tsan_get_current_fiber(void);
void *__tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void *fiber);
void __tsan_switch_to_fiber(void *fiber, unsigned flags);
void __tsan_acquire(void *addr);
void __tsan_release(void *addr);
void AnnotateIgnoreReadsBegin(const char *f, int l);
void AnnotateIgnoreReadsEnd(const char *f, int l);
void AnnotateIgnoreWritesBegin(const char *f, int l);
void AnnotateIgnoreWritesEnd(const char *f, int l);
// volatile accesses (GCC --param tsan-distinguish-volatile=1 calls these instead of the plain hooks; libtsan itself does
// not define them): a relaxed atomic load, or a relaxed read-modify-write that leaves the value alone
#define KR_TSAN_VOL(N, T)                                                                                                      \
    __attribute__((weak)) void __tsan_volatile_read##N(void *a) { (void)__atomic_load_n((T *)a, __ATOMIC_RELAXED); }            \
    __attribute__((weak)) void __tsan_volatile_write##N(void *a) { (void)__atomic_fetch_add((T *)a, (T)0, __ATOMIC_RELAXED); }
KR_TSAN_VOL(1, uint8_t) KR_TSAN_VOL(2, uint16_t) KR_TSAN_VOL(4, uint32_t) KR_TSAN_VOL(8, uint64_t)
#undef KR_TSAN_VOL
}
namespace emu {
// the emulator's own bookkeeping (barrier counters, scheduler state, exchange buffers) is shared by all fibers by design:
// accesses inside a Quiet scope are not recorded
struct Quiet {
    Quiet() { AnnotateIgnoreReadsBegin(__FILE__, __LINE__); AnnotateIgnoreWritesBegin(__FILE__, __LINE__); }
    ~Quiet() { AnnotateIgnoreWritesEnd(__FILE__, __LINE__); AnnotateIgnoreReadsEnd(__FILE__, __LINE__); }
};
inline void hb_release(void *a) { __tsan_release(a); }
inline void hb_acquire(void *a) { __tsan_acquire(a); }
inline void *&tsan_main() { static void *m = nullptr; return m; }
// Who a CUDA thread is to the detector.  ThreadSanitizer multiplexes 256 thread slots and loses history when more threads
// than that are alive, so there are two granularities: a thread of its own (default: what finds a missing __syncwarp; the
// identities are given back at the end of every launch, so only one CTA's worth is ever alive), or -- KR_EMUL_RACE_ID=warp
// -- one identity per WARP, kept for the life of the stream: a few dozen identities in all, which is what makes conflicts
// BETWEEN launches, streams and CTAs (and between the warps of the 800-thread tile join) reliably visible.  The warp
// build needs --param tsan-instrument-func-entry-exit=0 (32 fibers share one shadow call stack otherwise).
inline bool warp_identity() { static const bool w = [] { const char *m = getenv("KR_EMUL_RACE_ID"); return m && !strcmp(m, "warp"); }(); return w; }
inline unsigned tsan_identity(bool grid_mode, unsigned b, unsigned nthreads, unsigned t) {
    const unsigned per = warp_identity() ? (nthreads + 31) / 32 : nthreads, me = warp_identity() ? t >> 5 : t;
    return grid_mode ? (b & 1) * per + me : me;
}
inline void tsan_to(void *f) { __tsan_switch_to_fiber(f, 1u /* no synchronisation */); }
}
#else
namespace emu {
struct Quiet { Quiet() {} ~Quiet() {} };
inline void hb_release(void *) {}
inline void hb_acquire(void *) {}
}
#endif

// streams and events (race detection: a stream carries a clock -- `tok` = everything it has completed, `tok_out` = what the
// kernel in flight has done so far -- and an "actor", the ThreadSanitizer identity its copies / memsets / frees run under)
typedef struct CUstream_st {
    int id; char *tok; char *tok_out; void *actor;
    // race detection: the CUDA threads of kernels on this stream are ThreadSanitizer fibers (and stacks) of this stream alone,
    // so that thread t of a kernel on one stream and thread t of a kernel on another are two threads to the detector
    std::vector<void *> fibers; std::vector<uint8_t *> stacks;
} *cudaStream_t;
typedef struct CUevent_st { std::chrono::steady_clock::time_point t; char *tok; } *cudaEvent_t;
namespace emu {
inline char *launch_tok() { static char t; return &t; }          // host -> stream work: released by the calling host thread
inline CUstream_st *stream_of(cudaStream_t s) {                   // the null stream is a stream like the others here
    static CUstream_st *null_stream = nullptr;
    if (s) return s;
    Quiet q;
    if (!null_stream) {
        null_stream = new CUstream_st{0, new char(0), new char(0), nullptr, {}, {}};
#ifdef KR_EMUL_TSAN
        null_stream->actor = __tsan_create_fiber(0);
#endif
    }
    return null_stream;
}
inline std::mutex &registry_lock() { static std::mutex m; return m; }
inline std::vector<CUstream_st *> &all_streams() { static std::vector<CUstream_st *> v; return v; }
inline std::map<const void *, size_t> &pinned_ranges() { static std::map<const void *, size_t> m; return m; }
inline bool is_pinned(const void *p) {
    std::lock_guard<std::mutex> lk(registry_lock());
    auto &m = pinned_ranges();
    auto it = m.upper_bound(p);
    if (it == m.begin()) return false;
    --it;
    return (const char *)p < (const char *)it->first + it->second;
}
#ifdef KR_EMUL_TSAN
// The calling host thread executes one operation of a stream: ordered after the host's own past and the stream's, and
// nothing of it flows back to the host (that takes a synchronisation).  Only ThreadSanitizer's idea of "who" changes; the
// code keeps running on the caller's stack, so the scope holds nothing but calls into intercepted libc functions.
struct OnStream {
    void *prev; CUstream_st *s;
    KR_NOTSAN explicit OnStream(cudaStream_t st) : s(stream_of(st)) {
        hb_release(launch_tok());
        prev = __tsan_get_current_fiber();
        __tsan_switch_to_fiber(s->actor, 1u);
        hb_acquire(launch_tok()); hb_acquire(s->tok);
    }
    KR_NOTSAN ~OnStream() { hb_release(s->tok); __tsan_switch_to_fiber(prev, 1u); }
};
inline void host_waits_for(char *tok) { if (tok) hb_acquire(tok); }
#else
struct OnStream { explicit OnStream(cudaStream_t) {} };
inline void host_waits_for(char *) {}
#endif
}

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
// static __shared__ variables are statics placed in one ELF section of their own, so that the launcher can poison them all
// before every CTA: on hardware shared memory starts undefined in every CTA, and a kernel that relies on zeros (or on what
// the previous CTA left there) must fail here too.  (GCC ignores the section for statics of a TEMPLATE function -- they are
// COMDAT -- so kernels that are templates declare their __shared__ arrays at namespace scope, as kr_encode.cuh does.)
#define __shared__ static __attribute__((section("kr_emu_smem")))
extern "C" { extern char __start_kr_emu_smem[] __attribute__((weak)), __stop_kr_emu_smem[] __attribute__((weak)); }
#define __restrict__

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a) : x(a) {} };

#ifdef KR_EMUL_TSAN
#include <sys/mman.h>
// first object of the static shared-memory section, page-aligned; race_check.cpp defines the last one (kr_emu_smem_tail)
__attribute__((section("kr_emu_smem"), aligned(4096), used)) static char kr_emu_smem_head[8192];
extern "C" char kr_emu_smem_tail[];
namespace emu {
inline void static_smem_renew() {      // (the compiler may lay the section out in either order of definition)
    char *lo = __start_kr_emu_smem, *hi = __stop_kr_emu_smem;
    char *h = kr_emu_smem_head, *t = kr_emu_smem_tail;
    char *first = h < t ? h : t, *last = h < t ? t : h;
    if (lo != first || (reinterpret_cast<uintptr_t>(lo) & 4095) || last + 8192 != hi) {
        fprintf(stderr, "cuda_emul: the static shared-memory section is not bracketed by its head and tail objects\n"); abort();
    }
    if (mmap(lo, (size_t)(hi - lo), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED, -1, 0) != lo) { perror("cuda_emul: mmap"); abort(); }
}
}
#endif

namespace emu {

// One CUDA thread = one fiber (ucontext); the fibers of a CTA are run round-robin by the launching OS thread, so
// execution is single-threaded and deterministic, a blocked fiber just yields, and a CTA in which every fiber is
// blocked is reported as a deadlock instead of hanging.
// Context switch: on x86-64 a few instructions of assembly (callee-saved registers + stack pointer); swapcontext()
// elsewhere (it makes a signal-mask system call per switch, which dominates the run time of barrier-heavy kernels).
#if defined(__x86_64__)
#define KR_EMU_ASM_SWITCH 1
extern "C" void kr_emu_switch(void **save_sp, void *new_sp);
__asm__(".text\n.weak kr_emu_switch\n.type kr_emu_switch,@function\nkr_emu_switch:\n"
        "pushq %rbp\npushq %rbx\npushq %r12\npushq %r13\npushq %r14\npushq %r15\n"
        "movq %rsp, (%rdi)\nmovq %rsi, %rsp\n"
        "popq %r15\npopq %r14\npopq %r13\npopq %r12\npopq %rbx\npopq %rbp\nret\n"
        ".size kr_emu_switch, .-kr_emu_switch\n");
struct Ctx { void *sp = nullptr; };
KR_NOTSAN inline void ctx_switch(Ctx *from, Ctx *to) { kr_emu_switch(&from->sp, to->sp); }
inline void ctx_make(Ctx *c, uint8_t *stack, size_t bytes, void (*entry)()) {
    uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~uintptr_t(15);
    void **sp = reinterpret_cast<void **>(top);
    *--sp = nullptr;                                   // return address of `entry` (it never returns)
    *--sp = reinterpret_cast<void *>(entry);           // popped by the `ret` of the first switch
    for (int i = 0; i < 6; ++i) *--sp = nullptr;       // rbp rbx r12 r13 r14 r15
    c->sp = sp;
}
#else
struct Ctx { ucontext_t uc; };
KR_NOTSAN inline void ctx_switch(Ctx *from, Ctx *to) { swapcontext(&from->uc, &to->uc); }
inline void ctx_make(Ctx *c, uint8_t *stack, size_t bytes, void (*entry)()) {
    getcontext(&c->uc); c->uc.uc_stack.ss_sp = stack; c->uc.uc_stack.ss_size = bytes; c->uc.uc_link = nullptr;
    makecontext(&c->uc, entry, 0);
}
#endif
struct Fiber { Ctx ctx; dim3 tid; bool done = false; uint8_t *stack = nullptr; };
// fiber stacks are kept for the life of the process and reused by every launch (no page faults per launch)
inline uint8_t *fiber_stack(unsigned t, size_t bytes) {
    static std::vector<uint8_t *> pool;
    if (pool.size() <= t) pool.resize(t + 1, nullptr);
    if (!pool[t]) pool[t] = static_cast<uint8_t *>(malloc(bytes));
    return pool[t];
}
struct Sched {
    Ctx main;
    std::vector<Fiber> fibers;
    unsigned cur = 0;
    unsigned long long idle_yields = 0;   // yields since the last sign of progress
};
inline Sched *&sched() { static Sched *s = nullptr; return s; }
inline void progress() { Quiet q; sched()->idle_yields = 0; }
inline void yield() {
    Quiet q;
    Sched *s = sched();
    if (++s->idle_yields > 50000000ull) { fprintf(stderr, "cuda_emul: deadlock (every thread of the CTA is waiting)\n"); abort(); }
#ifdef KR_EMUL_TSAN
    tsan_to(tsan_main());
#endif
    ctx_switch(&s->fibers[s->cur].ctx, &s->main);
}

// a reusable counting barrier whose party count can shrink (threads that leave the kernel early)
class Barrier {
  public:
    explicit Barrier(unsigned n) : parties_(n), waiting_(0), gen_(0) {}
    // orders = true: everything the arrivers did before happens-before everything they do after (__syncthreads, named
    // barriers, __syncwarp); false: a lock-step exchange that orders no memory (shuffles, votes).  One token per
    // generation parity, so that a thread already arriving at the NEXT barrier adds nothing to what a slow thread still
    // leaving this one acquires.
    void arrive_and_wait(bool orders = true) {
        unsigned g;
        { Quiet q; g = gen_; }
        if (orders) hb_release(&tok_[g & 1]);
        {
            Quiet q;
            if (++waiting_ >= parties_) { waiting_ = 0; ++gen_; progress(); }
            else while (gen_ == g) yield();
        }
        if (orders) hb_acquire(&tok_[g & 1]);
    }
    void drop() {      // a thread that leaves the kernel: what it did is visible to the others from their next barrier on
        hb_release(&tok_[0]); hb_release(&tok_[1]);
        Quiet q;
        --parties_;
        if (parties_ && waiting_ >= parties_) { waiting_ = 0; ++gen_; progress(); }
    }
  private:
    unsigned parties_, waiting_, gen_;
    char tok_[2] = {0, 0};
};

struct Warp { Barrier bar; uint64_t buf[32]; unsigned n; explicit Warp(unsigned k) : bar(k), n(k) {} };

struct Block {
    unsigned nthreads;
    Barrier sync;
    std::vector<std::unique_ptr<Warp>> warps;
    std::map<int, std::unique_ptr<Barrier>> named;
    struct Vote { unsigned acc[2] = {0, 0}, arrived = 0, gen = 0; };
    std::map<int, Vote> votes;
    explicit Block(unsigned n) : nthreads(n), sync(n) {
        for (unsigned w = 0; w * 32 < n; ++w) warps.emplace_back(new Warp(std::min(32u, n - w * 32)));
    }
};

inline Block *&cur_block() { static Block *b = nullptr; return b; }
inline uint8_t *&dyn_smem() { static uint8_t *p = nullptr; return p; }
inline size_t &dyn_smem_size() { static size_t n = 0; return n; }
inline std::mutex &launch_lock() { static std::mutex m; return m; }   // one launch at a time per process
struct Trap {};
inline std::function<void()> *&cur_body() { static std::function<void()> *f = nullptr; return f; }
inline bool &trapped() { static bool t = false; return t; }

inline CUstream_st *&cur_stream();
inline void fiber_main() {
    CUstream_st *st;
    { Quiet q; st = cur_stream(); }
    hb_acquire(launch_tok()); hb_acquire(st->tok);     // after the host at launch time and after the stream's earlier work
    Sched *s; Fiber *f; std::function<void()> *body;
    { Quiet q; s = sched(); f = &s->fibers[s->cur]; body = cur_body(); }
    try { (*body)(); } catch (Trap &) { trapped() = true; }
    { Quiet q; f->done = true; }
    progress();
    Barrier *b0, *b1;
    { Quiet q; Block *blk = cur_block(); b0 = &blk->sync; b1 = &blk->warps[f->tid.x >> 5]->bar; }
    b0->drop();                                    // a thread that has left the kernel no longer takes part
    b1->drop();
    hb_release(st->tok_out);                       // (last: nothing this thread touches comes after it)
    Ctx *mine, *back;
    { Quiet q; mine = &f->ctx; back = &s->main; }  // (no ignore scope stays open: the fiber is destroyed with its stream)
#ifdef KR_EMUL_TSAN
    tsan_to(tsan_main());
#endif
    ctx_switch(mine, back);
    abort();                                       // a finished fiber is never resumed
}

}  // namespace emu

inline dim3 blockIdx, blockDim, gridDim;
namespace emu { inline dim3 tid() { Quiet q; Sched *s = sched(); return s->fibers[s->cur].tid; } }
#define threadIdx (emu::tid())

namespace emu {

inline Warp &my_warp() { const unsigned w = threadIdx.x >> 5; Quiet q; return *cur_block()->warps[w]; }

inline CUstream_st *&cur_stream() { static CUstream_st *s = nullptr; return s; }
template <class F> void launch(dim3 grid, dim3 block, size_t smem_bytes, cudaStream_t stream, F body) {
    std::lock_guard<std::mutex> lk(launch_lock());
    hb_release(launch_tok());                          // the kernel comes after everything this host thread has done
    Quiet quiet;                                       // the launcher's bookkeeping is not the kernel's memory traffic
    cur_stream() = stream_of(stream);
    gridDim = grid; blockDim = block;
#ifdef KR_EMUL_TSAN
    static const bool grid_mode = [] { const char *m = getenv("KR_EMUL_RACE"); return m && !strcmp(m, "grid"); }();
#else
    const bool grid_mode = false;
#endif
    std::vector<uint8_t> smem(smem_bytes + 2048);
    uint8_t *base = smem.data() + (1024 - (reinterpret_cast<uintptr_t>(smem.data()) & 1023));
    dyn_smem() = base; dyn_smem_size() = smem_bytes;
#ifdef KR_EMUL_TSAN
    // (never freed: the kernel is asynchronous to the host in the detector's eyes, so the host may not touch -- or
    // re-use the memory of -- what the kernel's threads read; the arguments were captured by value into this object)
    std::function<void()> &fn = *new std::function<void()>(body);
#else
    std::function<void()> fn = body;
#endif
    cur_body() = &fn;
    Sched sc;
    sc.fibers.resize(block.x);
    const size_t stack_bytes = 96 * 1024;
    for (unsigned t = 0; t < block.x; ++t) sc.fibers[t].stack = fiber_stack(t, stack_bytes);
    sched() = &sc;
    trapped() = false;
#ifdef KR_EMUL_TSAN
    tsan_main() = __tsan_get_current_fiber();
#endif
    for (unsigned b = 0; b < grid.x; ++b) {
        std::unique_ptr<Block> blkp(new Block(block.x));   // (heap: ThreadSanitizer forgets a barrier's history when its memory is freed)
        Block &blk = *blkp;
        cur_block() = &blk;
        blockIdx.x = b;
#ifdef KR_EMUL_TSAN
        // Every CTA has shared memory of its own on the device; here all CTAs of all launches (of all streams) re-use the same
        // bytes.  So that the detector does not take that re-use for a conflict, both kinds of shared memory get a fresh
        // history per CTA: allocating inside the launcher's ignore scope resets ThreadSanitizer's shadow for the range
        // (malloc for the dynamic part, an anonymous mapping laid over the page-aligned section of the static part).
        smem = std::vector<uint8_t>(smem_bytes + 2048);
        base = smem.data() + (1024 - (reinterpret_cast<uintptr_t>(smem.data()) & 1023));
        dyn_smem() = base;
        static_smem_renew();
#endif
        memset(base, 0xCD, smem_bytes);                // shared memory starts undefined in every CTA: make reliance on zeros show
        if (__start_kr_emu_smem && __stop_kr_emu_smem > __start_kr_emu_smem) memset(__start_kr_emu_smem, 0xCD, (size_t)(__stop_kr_emu_smem - __start_kr_emu_smem));
        for (unsigned t = 0; t < block.x; ++t) {
            Fiber &f = sc.fibers[t];
            f.done = false; f.tid = dim3(t);
            // grid mode: the threads of two consecutive CTAs are different ThreadSanitizer fibers (with stacks of their
            // own), so that thread t of one CTA and thread t of the next are not "the same thread" to the detector
#ifdef KR_EMUL_TSAN
            {
                CUstream_st *st = cur_stream();
                const unsigned sidx = grid_mode ? (b & 1) * block.x + t : t, idx = tsan_identity(grid_mode, b, block.x, t);
                if (st->stacks.size() <= sidx) st->stacks.resize(sidx + 1, nullptr);
                if (st->fibers.size() <= idx) st->fibers.resize(idx + 1, nullptr);
                if (!st->stacks[sidx]) st->stacks[sidx] = static_cast<uint8_t *>(malloc(stack_bytes));
                if (!st->fibers[idx]) st->fibers[idx] = __tsan_create_fiber(0);
                f.stack = st->stacks[sidx];
            }
#endif
            ctx_make(&f.ctx, f.stack, stack_bytes, fiber_main);
        }
        sc.idle_yields = 0;
        // Between two synchronisation points a fiber runs undisturbed, so the ORDER in which the fibers of a CTA get
        // their turn decides which side of a missing barrier wins.  KR_EMUL_ORDER=reverse / random runs them last-to-first /
        // shuffled per CTA: a kernel that only works because thread 0 happens to run first fails under one of them.
        static const char *order_mode = getenv("KR_EMUL_ORDER");
        std::vector<unsigned> order(block.x);
        for (unsigned t = 0; t < block.x; ++t) order[t] = t;
        if (order_mode && !strcmp(order_mode, "reverse")) std::reverse(order.begin(), order.end());
        else if (order_mode && !strcmp(order_mode, "random")) {
            uint64_t x = 0x9E3779B97F4A7C15ull * (b + 1) + block.x;
            for (unsigned t = block.x; t > 1; --t) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; std::swap(order[t - 1], order[x % t]); }
        }
        for (unsigned live = block.x; live;) {
            live = 0;
            for (unsigned ti = 0; ti < block.x; ++ti) {
                const unsigned t = order[ti];
                if (sc.fibers[t].done) continue;
                sc.cur = t;
                Ctx *to = &sc.fibers[t].ctx;           // (nothing of the launcher's is touched under the thread's identity)
#ifdef KR_EMUL_TSAN
                tsan_to(cur_stream()->fibers[tsan_identity(grid_mode, b, block.x, t)]);
#endif
                ctx_switch(&sc.main, to);
                if (!sc.fibers[t].done) ++live;
            }
        }
#ifdef KR_EMUL_TSAN
        // CTA mode: what this CTA did is the stream's past for the next CTA (nothing flows to the host: a launch is asynchronous)
        if (!grid_mode) { void *me = __tsan_get_current_fiber(); CUstream_st *st = cur_stream(); __tsan_switch_to_fiber(st->actor, 1u);
                          hb_acquire(st->tok_out); hb_release(st->tok); __tsan_switch_to_fiber(me, 1u); }
#endif
        cur_block() = nullptr;
        if (trapped()) { fprintf(stderr, "cuda_emul: kernel trapped in block %u\n", b); abort(); }
    }
#ifdef KR_EMUL_TSAN
    // grid mode: only the end of the launch orders its CTAs before what follows on the stream
    if (grid_mode) { void *me = __tsan_get_current_fiber(); CUstream_st *st = cur_stream(); __tsan_switch_to_fiber(st->actor, 1u);
                     hb_acquire(st->tok_out); hb_release(st->tok); __tsan_switch_to_fiber(me, 1u); }
#endif
#ifdef KR_EMUL_TSAN
    // per-thread identities are given back as soon as the launch has ended (see tsan_identity; the stacks stay: the next
    // launch on this stream is ordered after this one through the stream's clock)
    if (!warp_identity()) {
        for (void *f : cur_stream()->fibers) if (f) __tsan_destroy_fiber(f);
        std::fill(cur_stream()->fibers.begin(), cur_stream()->fibers.end(), nullptr);
    }
#endif
    sched() = nullptr;
}

}  // namespace emu

// ---- device-side intrinsics ------------------------------------------------------------------------------------
inline void __syncthreads() { emu::Block *b; { emu::Quiet q; b = emu::cur_block(); } b->sync.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::my_warp().bar.arrive_and_wait(); }
inline void __trap() { throw emu::Trap(); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }
template <class T> inline T min(T a, T b) { return a < b ? a : b; }
template <class T> inline T max(T a, T b) { return a > b ? a : b; }
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }

namespace emu {
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
// every thread of the warp publishes its value, then reads the lane it wants
template <class T> inline T exchange(T v, unsigned src_lane, bool keep_own_if_out_of_range) {
    Warp &w = my_warp();
    const unsigned lane = threadIdx.x & 31;
    { Quiet q; w.buf[lane] = to_bits(v); }
    w.bar.arrive_and_wait(false);                  // a shuffle is not a memory barrier
    T r;
    { Quiet q; r = (src_lane < w.n) ? from_bits<T>(w.buf[src_lane]) : v; }
    (void)keep_own_if_out_of_range;
    w.bar.arrive_and_wait(false);
    return r;
}
}  // namespace emu
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu::exchange(v, (unsigned)src & 31, true); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int mask) { return emu::exchange(v, (threadIdx.x & 31) ^ (unsigned)mask, true); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const unsigned lane = threadIdx.x & 31;
    return emu::exchange(v, lane >= d ? lane - d : lane, true);
}
inline unsigned __ballot_sync(unsigned, bool pred) {
    emu::Warp &w = emu::my_warp();
    const unsigned lane = threadIdx.x & 31;
    { emu::Quiet q; w.buf[lane] = pred ? 1 : 0; }
    w.bar.arrive_and_wait(false);                  // a vote is not a memory barrier
    unsigned r = 0;
    { emu::Quiet q; for (unsigned l = 0; l < w.n; ++l) if (w.buf[l]) r |= 1u << l; }
    w.bar.arrive_and_wait(false);
    return r;
}
inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0; }

// Lanes of a warp run concurrently on hardware; here a fiber runs undisturbed between synchronisation points.  With
// KR_EMUL_PREEMPT=<seed> a fiber is also suspended, pseudo-randomly, just BEFORE an atomic operation executes, so the
// other threads of the CTA interleave inside lock-free sections (a slot seen empty and claimed by somebody else before
// the compare-and-swap lands, counters bumped between a read and its use).
namespace emu {
inline void preempt() {
    static const long mode = [] { const char *s = getenv("KR_EMUL_PREEMPT"); return s ? atol(s) : 0L; }();
    bool go;
    {
        Quiet q;
        if (!mode || !sched()) return;
        static uint64_t x = 0x2545F4914F6CDD1Dull * (uint64_t)mode + 1;
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        go = (x & 3) == 0;
    }
    if (go) { progress(); yield(); }
}
}  // namespace emu
// atomics (global or shared memory alike).  RELAXED, as on the device: an atomic orders nothing but itself.
template <class T> inline T atomicAdd(T *p, T v) { emu::preempt(); return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned *p, int v) { emu::preempt(); return __atomic_fetch_add(p, (unsigned)v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long v) { emu::preempt(); return __atomic_fetch_add(p, (unsigned long long)v, __ATOMIC_RELAXED); }
template <class T> inline T atomicMax(T *p, T v) {
    emu::preempt();
    T o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
template <class T> inline T atomicExch(T *p, T v) { emu::preempt(); return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicCAS(T *p, T cmp, T v) { emu::preempt(); __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp; }

// ---- mbarrier + bulk copy (the PTX wrappers of kr_kernels.cuh call these under KR_EMUL) -------------------------------
namespace emu {
struct MBar { uint32_t phase : 1, pending : 15, expected : 15; int32_t tx; };
static_assert(sizeof(MBar) == 8, "an mbarrier is 64 bits");
// happens-before: an arrival (or a completed transaction) of the phase with parity P releases token P of the barrier, a
// successful try_wait.parity(P) acquires it (the tokens are the first two bytes of the barrier word itself, as addresses)
inline void mbar_hb_release(void *bar) { unsigned ph; { Quiet q; ph = ((MBar *)bar)->phase; } hb_release((char *)bar + ph); }
inline void mbar_check(MBar *b) { if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->expected; progress(); } }
inline void mbar_init(void *bar, uint32_t count) { Quiet q; MBar *b = (MBar *)bar; b->phase = 0; b->pending = b->expected = count; b->tx = 0; }
inline void mbar_arrive(void *bar, uint32_t n) {
    mbar_hb_release(bar);
    Quiet q;
    MBar *b = (MBar *)bar;
    if (b->pending < n) { fprintf(stderr, "cuda_emul: mbarrier over-arrival\n"); abort(); }
    b->pending -= n; mbar_check(b);
}
inline void mbar_expect_tx(void *bar, uint32_t bytes) {   // arrive.expect_tx: one arrival + a transaction count
    mbar_hb_release(bar);
    Quiet q;
    MBar *b = (MBar *)bar;
    b->tx += (int32_t)bytes;
    if (b->pending < 1) { fprintf(stderr, "cuda_emul: mbarrier over-arrival\n"); abort(); }
    b->pending -= 1; mbar_check(b);
}
inline void mbar_complete_tx(void *bar, uint32_t bytes) { mbar_hb_release(bar); Quiet q; MBar *b = (MBar *)bar; b->tx -= (int32_t)bytes; mbar_check(b); }
inline bool mbar_try_wait(void *bar, uint32_t parity) {
    bool done;
    { Quiet q; done = ((MBar *)bar)->phase != (parity & 1); }
    if (!done) yield();
    else hb_acquire((char *)bar + (parity & 1));
    return done;
}
inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, void *bar) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (bytes & 15)) {
        fprintf(stderr, "cuda_emul: cp.async.bulk needs 16-byte aligned dst/src/size (dst %p src %p bytes %u): this faults on hardware\n", dst, src, bytes);
        abort();
    }
    if ((uint8_t *)dst < dyn_smem() || (uint8_t *)dst + bytes > dyn_smem() + dyn_smem_size()) { fprintf(stderr, "cuda_emul: bulk copy outside shared memory\n"); abort(); }
    memcpy(dst, src, bytes);
    mbar_complete_tx(bar, bytes);
}
inline void named_barrier(int id, unsigned count) {
    Barrier *b;
    { Quiet q; Block *blk = cur_block(); auto &slot = blk->named[id]; if (!slot) slot.reset(new Barrier(count)); b = slot.get(); }
    b->arrive_and_wait();
}
// barrier.cta.red.or.pred: the same barrier with an OR reduction of a predicate; every thread gets the same answer
inline bool named_barrier_or(int id, unsigned count, bool pred) {
    Block::Vote *v; unsigned g;
    {
        Quiet q;
        v = &cur_block()->votes[id];
        g = v->gen & 1;
        if (pred) v->acc[g] = 1;
        if (++v->arrived == count) { v->arrived = 0; ++v->gen; v->acc[v->gen & 1] = 0; }   // the last arriver opens the next generation
    }
    named_barrier(id, count);
    Quiet q;
    return v->acc[g] != 0;
}
// shared-memory "addresses" are offsets into the dynamic buffer (+ a bias so that 0 is never valid)
inline uint32_t smem_addr(const void *p) {
    const uint8_t *q = (const uint8_t *)p;
    if (q < dyn_smem() || q > dyn_smem() + dyn_smem_size()) { fprintf(stderr, "cuda_emul: address is not in dynamic shared memory\n"); abort(); }
    return (uint32_t)(q - dyn_smem()) + 4096;
}
inline void *smem_ptr(uint32_t a) {
    if (a < 4096 || a - 4096 + 4 > dyn_smem_size() + 16) { fprintf(stderr, "cuda_emul: shared load outside the CTA's shared memory (offset %u of %zu)\n", a - 4096, dyn_smem_size()); abort(); }
    return dyn_smem() + (a - 4096);
}
}  // namespace emu

// ---- the runtime API subset kr_api.cu uses (single device, synchronous streams) -------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmul = 1 };
typedef struct CUmempool_st { int id; } *cudaMemPool_t;
enum { cudaMemAllocationTypePinned = 1, cudaMemLocationTypeDevice = 1 };
struct cudaMemLocation { int type = 0, id = 0; };
struct cudaMemPoolProps { int allocType = 0; cudaMemLocation location; };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaMemPoolAttrReleaseThreshold = 4,
       cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 4; size_t sharedMemPerBlockOptin = 232448; char name[64] = "cuda_emul"; };
inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emulated error" : "no error"; }
inline cudaError_t cudaGetDeviceCount(int *n) { const char *s = getenv("KR_EMUL_DEVICES"); *n = s ? atoi(s) : 1; return cudaSuccess; }   // all "devices" are this host
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); if (const char *s = getenv("KR_EMUL_SMS")) p->multiProcessorCount = atoi(s); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    CUstream_st *st = new CUstream_st{1, new char(0), new char(0), nullptr, {}, {}};
#ifdef KR_EMUL_TSAN
    st->actor = __tsan_create_fiber(0);
#endif
    { std::lock_guard<std::mutex> lk(emu::registry_lock()); emu::all_streams().push_back(st); }
    *s = st;
    return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t s) {      // (the record is kept: an event may still name the stream's clock)
    std::lock_guard<std::mutex> lk(emu::registry_lock());
    auto &v = emu::all_streams();
    v.erase(std::remove(v.begin(), v.end(), s), v.end());
#ifdef KR_EMUL_TSAN
    for (void *f : s->fibers) if (f) __tsan_destroy_fiber(f);
    for (uint8_t *st : s->stacks) free(st);
    s->fibers.clear(); s->stacks.clear();
#endif
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t s) { emu::host_waits_for(emu::stream_of(s)->tok); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronizeEmul() {              // what a blocking cudaFree / cudaMalloc implies
    std::vector<CUstream_st *> v;
    { std::lock_guard<std::mutex> lk(emu::registry_lock()); v = emu::all_streams(); }
    for (CUstream_st *st : v) emu::host_waits_for(st->tok);
    emu::host_waits_for(emu::stream_of(nullptr)->tok);
    return cudaSuccess;
}
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new CUevent_st(); (*e)->tok = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
KR_NOTSAN inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {      // the event takes the stream's clock as of now (a fresh token per record)
    char *tok = new char(0);
    { emu::OnStream on(s); emu::hb_release(tok); }
    emu::Quiet q;
    e->t = std::chrono::steady_clock::now(); e->tok = tok;
    return cudaSuccess;
}
KR_NOTSAN inline cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    char *tok;
    { emu::Quiet q; tok = e->tok; }
    if (tok) { emu::OnStream on(s); emu::hb_acquire(tok); }
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t e) { char *tok; { emu::Quiet q; tok = e->tok; } emu::host_waits_for(tok); return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { emu::Quiet q; *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
inline cudaError_t cudaMemPoolCreate(cudaMemPool_t *p, const cudaMemPoolProps *) { *p = new CUmempool_st{1}; return cudaSuccess; }
inline cudaError_t cudaMemPoolDestroy(cudaMemPool_t p) { delete p; return cudaSuccess; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void *) { return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
// allocations are 256-byte aligned like cudaMalloc's and poisoned so that reads of uninitialised device memory show
KR_NOTSAN inline cudaError_t cudaMalloc(void **p, size_t n) { if (posix_memalign(p, 256, n ? n : 1)) return cudaErrorEmul; memset(*p, 0xCD, n); return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
// stream-ordered allocation and release: the memory exists from this point of the stream on, and is gone after that one
KR_NOTSAN inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t s) {
    void *q = nullptr; cudaError_t r;
    { emu::OnStream on(s); r = cudaMalloc(&q, n); }
    *p = q;
    return r;
}
inline cudaError_t cudaMallocFromPoolAsync(void **p, size_t n, cudaMemPool_t, cudaStream_t s) { return cudaMallocAsync(p, n, s); }
inline cudaError_t cudaFree(void *p) { cudaDeviceSynchronizeEmul(); free(p); return cudaSuccess; }      // cudaFree synchronises the device
KR_NOTSAN inline cudaError_t cudaFreeAsync(void *p, cudaStream_t s) { emu::OnStream on(s); free(p); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) {
    *p = malloc(n ? n : 1);
    if (!*p) return cudaErrorEmul;
    std::lock_guard<std::mutex> lk(emu::registry_lock());
    emu::pinned_ranges()[*p] = n ? n : 1;
    return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void *p) { { std::lock_guard<std::mutex> lk(emu::registry_lock()); emu::pinned_ranges().erase(p); } free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { cudaDeviceSynchronizeEmul(); if (n) memmove(d, s, n); return cudaSuccess; }
// An asynchronous copy belongs to its stream.  Pageable host memory makes it partly synchronous, as on the device: a
// pageable SOURCE is staged before the call returns (the host may re-use it at once), a pageable DESTINATION is complete
// when the call returns.  Pinned memory (cudaHostAlloc) gets neither favour.
KR_NOTSAN inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st) {
    if (!n) return cudaSuccess;
    if (k == cudaMemcpyHostToDevice && !emu::is_pinned(s)) {
        void *stage = malloc(n);
        memcpy(stage, s, n);                               // read by the host, now
        { emu::OnStream on(st); memcpy(d, stage, n); free(stage); }
        return cudaSuccess;
    }
    { emu::OnStream on(st); memmove(d, s, n); }
    if (k == cudaMemcpyDeviceToHost && !emu::is_pinned(d)) emu::host_waits_for(emu::stream_of(st)->tok);
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void *d, int v, size_t n) { cudaDeviceSynchronizeEmul(); if (n) memset(d, v, n); return cudaSuccess; }
KR_NOTSAN inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t s) { if (n) { emu::OnStream on(s); memset(d, v, n); } return cudaSuccess; }
