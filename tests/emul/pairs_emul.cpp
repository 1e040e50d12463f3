// CPU shim test of kmers/reps_b200/csrc/kr_pairs.cuh (EXPERIMENTAL posting-list all-vs-all path).
// The kernels of that file use no __syncthreads and no warp intrinsics, so running their threads one after another
// is a faithful execution of their logic (cell numbering in the three layouts, mirrored cells, 16-bit packing, row
// stripes, the equal-sequence rule, the diagonal).  This checks that logic against a brute-force count; it says
// nothing about performance and it is not a product path (TEST INFRASTRUCTURE: compiled and run only by
// tests/test_pairs_emul.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <set>
#include <tuple>
#include <vector>

struct Dim3 { unsigned x = 1, y = 1, z = 1; };
static Dim3 threadIdx, blockIdx, blockDim, gridDim;
template <class T> static T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> static T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

struct uint4 { uint32_t x, y, z, w; };
static uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
#define KR_PAIRS_EMUL 1
#include "../../kmers/reps_b200/csrc/kr_pairs.cuh"
using namespace kr;

template <class K, class P> static void launch(K kernel, unsigned grid, unsigned block, P p) {
    gridDim.x = grid; blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b) for (unsigned t = 0; t < block; ++t) { blockIdx.x = b; threadIdx.x = t; kernel(p); }
}

static uint64_t row_start(uint32_t layout, uint64_t r, uint64_t n) {
    if (layout == 0) return r * n;
    if (layout == 1) return r * n - r * (r + 1) / 2;
    uint64_t I = r >> 8, c0 = I << 8;
    return 256ull * (I * n - 128ull * I * (I - 1)) + (r - c0) * (n - c0);
}

int main(int argc, char **argv) {
    const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1;
    std::mt19937_64 rng(seed);
    int checked = 0;
    for (int trial = 0; trial < 3; ++trial) {
        const uint32_t N = trial == 0 ? 1 : trial == 1 ? 257 : 258 + (uint32_t)(rng() % 330);
        // sequences as sets of small "keys"; a few exact duplicates (equal residues) and near-duplicates (equal sets)
        std::vector<std::set<uint32_t>> keys(N);
        std::vector<std::vector<uint8_t>> resid(N);
        for (uint32_t i = 0; i < N; ++i) {
            if (i > 3 && rng() % 10 == 0) { uint32_t src = (uint32_t)(rng() % i); keys[i] = keys[src]; resid[i] = resid[src]; if (rng() % 3 == 0) resid[i].push_back(7); continue; }
            uint32_t fam = (uint32_t)(rng() % 9), cnt = (uint32_t)(rng() % 40);
            for (uint32_t k = 0; k < cnt; ++k) keys[i].insert(fam * 50 + (uint32_t)(rng() % 60));
            resid[i].resize(5 + rng() % 20);
            for (auto &c : resid[i]) c = (uint8_t)('A' + rng() % 20);
        }
        std::vector<SeqMeta> meta(N);
        for (uint32_t i = 0; i < N; ++i) {
            meta[i].size = (uint32_t)keys[i].size(); meta[i].len = (uint32_t)resid[i].size(); meta[i].res = resid[i].data();
            uint64_t h = 1469598103934665603ull; for (uint8_t c : resid[i]) h = (h ^ c) * 1099511628211ull;
            meta[i].hash = (rng() % 4 == 0) ? 42 : h;    // forced hash collisions between different sequences too
            for (uint32_t j = 0; j < i; ++j) if (resid[j] == resid[i]) meta[i].hash = meta[j].hash;
        }
        // posting lists in a sparse slot table, lists in arbitrary order
        std::map<uint32_t, std::vector<uint32_t>> inv;
        for (uint32_t i = 0; i < N; ++i) for (uint32_t k : keys[i]) inv[k].push_back(i);
        const uint64_t nslots = 1021;   // not a multiple of 8: exercises the tail group
        std::vector<uint32_t> tab_cnt(nslots, 0), tab_off(nslots, 0), postings;
        std::vector<uint64_t> free_slots(nslots);
        for (uint64_t s = 0; s < nslots; ++s) free_slots[s] = s;
        std::shuffle(free_slots.begin(), free_slots.end(), rng);
        size_t fs = 0;
        for (auto &kv : inv) {
            auto lst = kv.second; std::shuffle(lst.begin(), lst.end(), rng);
            uint64_t s = free_slots[fs++];
            tab_cnt[s] = (uint32_t)lst.size(); tab_off[s] = (uint32_t)postings.size();
            postings.insert(postings.end(), lst.begin(), lst.end());
        }
        postings.push_back(0);
        // brute force
        auto sim = [&](uint32_t a, uint32_t b, bool ident_inf) -> uint32_t {
            if (ident_inf && resid[a] == resid[b]) return 9999;
            uint32_t c = 0; for (uint32_t k : keys[a]) c += (uint32_t)keys[b].count(k); return c;
        };
        for (uint32_t layout = 0; layout < 3; ++layout) for (uint32_t elem = 0; elem < 2; ++elem) for (uint32_t flags = 0; flags < 2; ++flags) {
            std::vector<std::pair<uint32_t, uint32_t>> stripes = {{0, N}};
            if (N > 256) { stripes.push_back({256, N}); stripes.push_back({0, 256}); stripes.push_back({256, 512 < N ? 512 : N}); }
            for (auto st : stripes) {
                PairsParams p{};
                p.tab_cnt = tab_cnt.data(); p.tab_off = tab_off.data(); p.postings = postings.data(); p.nslots = nslots;
                p.elem = elem; p.layout = layout; p.N = N; p.row_begin = st.first; p.row_end = st.second; p.flags = flags;
                p.origin = row_start(layout, st.first, N); p.total = row_start(layout, st.second, N) - p.origin;
                std::vector<uint8_t> out((p.total + 2) * (elem ? 2 : 4) + 8, 0);
                const size_t guard = p.total * (elem ? 2 : 4);
                memset(out.data() + guard, 0xEE, out.size() - guard);          // nothing may be written past the last cell
                p.out = out.data();
                uint32_t hm = 1; while (hm < 2 * N) hm <<= 1;
                std::vector<uint32_t> head(hm, PAIRS_NIL), next(N, PAIRS_NIL);
                p.meta = meta.data(); p.head = head.data(); p.next = next.data(); p.head_mask = hm - 1;
                launch(k_pairs_emit, 3, 64, p);
                launch(k_pairs_dup_insert, (N + 127) / 128, 128, p);
                launch(k_pairs_dup_fix, (N + 127) / 128, 128, p);
                for (size_t k = guard; k < out.size(); ++k) if (out[k] != 0xEE) { printf("FAIL: wrote past the result (N=%u layout=%u elem=%u)\n", N, layout, elem); return 1; }
                uint64_t pos = 0;
                for (uint32_t i = st.first; i < st.second; ++i) {
                    uint32_t j0 = layout == 0 ? 0 : layout == 1 ? i + 1 : (i / 256) * 256;
                    for (uint32_t j = j0; j < N; ++j, ++pos) {
                        uint32_t want = (i == j) ? (flags ? 9999u : meta[i].size) : sim(i, j, flags != 0);
                        uint32_t got = elem ? ((uint16_t *)out.data())[pos] : ((uint32_t *)out.data())[pos];
                        if (got != want) { printf("FAIL: N=%u layout=%u elem=%u flags=%u rows[%u,%u) cell(%u,%u) got %u want %u\n", N, layout, elem, flags, st.first, st.second, i, j, got, want); return 1; }
                    }
                }
                if (pos != p.total) { printf("FAIL: cell count\n"); return 1; }
                ++checked;
            }
        }
        // sparse result through the pair table: TRIANGLE cells with sim >= min_sim, per row stripe
        for (uint32_t flags = 0; flags < 2; ++flags) for (uint32_t min_sim : {1u, 3u, 9999u}) {
            std::vector<std::pair<uint32_t, uint32_t>> stripes = {{0, N}};
            if (N > 256) { stripes.push_back({256, N}); stripes.push_back({0, 256}); }
            for (auto st : stripes) {
                PairTabParams p{};
                p.tab_cnt = tab_cnt.data(); p.tab_off = tab_off.data(); p.postings = postings.data(); p.nslots = nslots;
                uint64_t work = 0; for (auto &kv : inv) work += (uint64_t)kv.second.size() * (kv.second.size() - 1) / 2;
                uint64_t pcap = 16; while (pcap < 2 * work + 2ull * N) pcap <<= 1;
                std::vector<unsigned long long> pt_key(pcap, PAIRTAB_EMPTY); std::vector<uint32_t> pt_cnt(pcap, 0);
                p.pt_key = pt_key.data(); p.pt_cnt = pt_cnt.data(); p.pt_mask = pcap - 1;
                p.N = N; p.row_begin = st.first; p.row_end = st.second; p.flags = flags; p.min_sim = min_sim;
                uint32_t hm = 1; while (hm < 2 * N) hm <<= 1;
                std::vector<uint32_t> head(hm, PAIRS_NIL), next(N, PAIRS_NIL);
                p.meta = meta.data(); p.head = head.data(); p.next = next.data(); p.head_mask = hm - 1;
                std::vector<uint4> cells(64); unsigned long long count = 0;
                p.cells = cells.data(); p.cells_cap = cells.size(); p.cells_count = &count;
                launch(k_pairs_emit_tab, 3, 64, p);
                launch(k_pairs_dup_insert_tab, (N + 127) / 128, 128, p);
                launch(k_pairs_dup_tab, (N + 127) / 128, 128, p);
                launch(k_pairs_compact, 2, 96, p);
                if (count > cells.size()) {          // too little room: the count is right, run the compaction again with enough
                    cells.assign(count, uint4{}); p.cells = cells.data(); p.cells_cap = cells.size(); count = 0;
                    launch(k_pairs_compact, 2, 96, p);
                }
                std::set<std::tuple<uint32_t, uint32_t, uint32_t>> got, want;
                for (unsigned long long c = 0; c < count; ++c) got.insert({cells[c].x, cells[c].y, cells[c].z});
                if (got.size() != count) { printf("FAIL: sparse cell listed twice\n"); return 1; }
                for (uint32_t i = st.first; i < st.second; ++i) for (uint32_t j = i + 1; j < N; ++j) { uint32_t v = sim(i, j, flags != 0); if (v >= min_sim) want.insert({i, j, v}); }
                if (got != want) { printf("FAIL: sparse cells differ N=%u flags=%u min_sim=%u rows[%u,%u): got %zu want %zu\n", N, flags, min_sim, st.first, st.second, got.size(), want.size()); return 1; }
                ++checked;
            }
        }
        // work estimate
        unsigned long long work = 0, want = 0;
        gridDim.x = 2; blockDim.x = 32;
        for (unsigned b = 0; b < 2; ++b) for (unsigned t = 0; t < 32; ++t) { blockIdx.x = b; threadIdx.x = t; k_pairs_work(tab_cnt.data(), nslots, &work); }
        for (auto &kv : inv) want += (unsigned long long)kv.second.size() * (kv.second.size() - 1) / 2;
        if (work != want) { printf("FAIL: work %llu want %llu\n", work, want); return 1; }
    }
    for (uint64_t t = 0; t < 3000000; t += (t < 5000 ? 1 : 997)) {
        uint32_t b = pairs_tri_row(t);
        if (!((uint64_t)b * (b - 1) / 2 <= t && t < (uint64_t)(b + 1) * b / 2)) { printf("FAIL: tri_row(%llu)\n", (unsigned long long)t); return 1; }
    }
    for (uint64_t m : {65535ull, 100000ull, 4000000000ull}) for (int d = -2; d <= 2; ++d) {
        uint64_t t = m * (m - 1) / 2 + d; uint32_t b = pairs_tri_row(t);
        if (!((uint64_t)b * (b - 1) / 2 <= t && t < (uint64_t)(b + 1) * b / 2)) { printf("FAIL: tri_row big\n"); return 1; }
    }
    printf("OK %d configurations\n", checked);
    return 0;
}
