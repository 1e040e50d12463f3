// race_check.cpp -- TEST INFRASTRUCTURE, CPU only: the product's device + host code (kr_api.cu and everything it includes)
// compiled with g++ -fsanitize=thread -DKR_EMUL -DKR_EMUL_TSAN and run on the emulator with every CUDA thread a
// ThreadSanitizer fiber (tests/emul/cuda_emul.h, "RACE DETECTION").  Each case drives the C ABI through one device path on a
// small synthetic batch; ThreadSanitizer reports every pair of conflicting accesses of two threads of a CTA that no
// barrier / __syncwarp / mbarrier orders -- the stand-in for `compute-sanitizer --tool racecheck` where there is no GPU.
// Results are also cross-checked between the device paths (tile join == index scan), so a run that is silent has also
// computed the right thing.
//
//   tools/emul_race.sh [case ...]        (builds tests/emul/_build/race_check and runs it; non-zero exit on any report)
#include "../../kmers/reps_b200/csrc/kr_api.cu"

#include <cinttypes>
#include <set>
#include <thread>
#include <tuple>

namespace {

struct Rng { uint64_t s; uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); } uint32_t below(uint32_t n) { return next() % n; } };
const char AA[] = "ACDEFGHIKLMNPQRSTVWY";

struct Batch { std::vector<uint8_t> res; std::vector<uint64_t> off{0}; uint32_t n() const { return (uint32_t)off.size() - 1; }
               void add(const std::string &s) { res.insert(res.end(), s.begin(), s.end()); off.push_back(res.size()); } };

// `n` sequences in `nfam` families (a root, point mutations, a random cut), plus the special cases the kernels branch on
Batch families(uint64_t seed, uint32_t n, uint32_t nfam, uint32_t len, bool specials) {
    Rng r{seed * 0x9E3779B97F4A7C15ull + 1};
    std::vector<std::string> roots(nfam);
    for (auto &s : roots) { s.resize(len + r.below(len / 2 + 1)); for (auto &c : s) c = AA[r.below(20)]; }
    Batch b;
    for (uint32_t i = 0; i < n; ++i) {
        std::string s = roots[r.below(nfam)];
        const uint32_t mut = (uint32_t)s.size() * r.below(4) / 20;
        for (uint32_t k = 0; k < mut; ++k) s[r.below((uint32_t)s.size())] = AA[r.below(20)];
        s.resize(s.size() / 2 + r.below((uint32_t)s.size() / 2 + 1));
        b.add(s);
    }
    if (specials) {
        b.add(""); b.add("MKV"); b.add("MKV"); b.add(std::string(400, 'A'));                 // k-mer-less, equal k-mer-less, one crowded bucket
        std::string lng(700, 'A'); for (auto &c : lng) c = AA[r.below(20)]; b.add(lng);        // 513..1024 k-mers: the register network
        std::string xl(1500, 'A'); for (auto &c : xl) c = AA[r.below(20)]; b.add(xl);          // > 1024 k-mers: sorted in global memory
        b.add(roots[0]); b.add(roots[0]);                                                      // equal sequences
    }
    return b;
}

#define OK(call) do { kr_status st_ = (call); if (st_ != KR_OK) { fprintf(stderr, "race_check: %s failed (%d): %s\n", #call, st_, kr_last_error(ctx)); exit(3); } } while (0)

std::vector<int32_t> full_matrix(kr_ctx *ctx, kr_seqset *a, kr_seqset *b, uint32_t rows, uint32_t cols) {
    std::vector<int32_t> m((size_t)rows * cols);
    kr_matrix_desc d{}; d.elem = KR_ELEM_I32; d.layout = KR_LAYOUT_FULL; d.row_begin = 0; d.row_end = rows;
    OK(kr_sim_matrix(ctx, a, b, &d, m.data()));
    return m;
}

// one context per device path: encode, dense matrices in every layout, sparse cells, pairs -- and the paths must agree
void case_matrix(int K, uint32_t n, bool specials) {
    Batch b = families(100 + K, n, 5, 90, specials);
    std::vector<int32_t> ref;
    for (uint32_t path : {KR_FORCE_JOIN, KR_FORCE_INDEX}) {
        kr_ctx *ctx = nullptr;
        OK(kr_create(0, K, KR_IDENT_IS_INFINITY | path, &ctx));
        kr_seqset *s = nullptr;
        OK(kr_encode(ctx, b.res.data(), b.off.data(), b.n(), &s));
        std::vector<int32_t> m = full_matrix(ctx, s, nullptr, b.n(), b.n());
        if (ref.empty()) ref = m; else if (ref != m) { fprintf(stderr, "race_check: the two device paths disagree (K=%d)\n", K); exit(4); }
        for (uint32_t layout : {KR_LAYOUT_TRIANGLE, KR_LAYOUT_BLOCKTRI}) {
            kr_matrix_desc d{}; d.elem = KR_ELEM_U16; d.layout = layout; d.row_begin = 0; d.row_end = b.n();
            std::vector<uint16_t> t(kr_matrix_elems(b.n(), b.n(), layout, 0, b.n()) + 8);
            OK(kr_sim_matrix(ctx, s, nullptr, &d, t.data()));
            size_t at = 0;
            for (uint32_t i = 0; i < b.n(); ++i)
                for (uint32_t j = layout == KR_LAYOUT_TRIANGLE ? i + 1 : (i & ~255u); j < b.n(); ++j, ++at)
                    if (t[at] != (uint16_t)ref[(size_t)i * b.n() + j]) { fprintf(stderr, "race_check: layout %u cell (%u,%u)\n", layout, i, j); exit(4); }
        }
        {   // sparse cells, packed
            kr_matrix_desc d{}; d.elem = KR_ELEM_I32; d.layout = KR_LAYOUT_TRIANGLE; d.row_begin = 0; d.row_end = b.n();
            std::vector<uint64_t> cells((size_t)b.n() * b.n());
            uint64_t cnt = 0;
            OK(kr_sim_matrix_sparse64(ctx, s, nullptr, &d, 1, cells.data(), cells.size(), &cnt));
            uint64_t want = 0;
            for (uint32_t i = 0; i < b.n(); ++i) for (uint32_t j = i + 1; j < b.n(); ++j) want += ref[(size_t)i * b.n() + j] > 0;
            if (cnt != want) { fprintf(stderr, "race_check: %" PRIu64 " sparse cells, %" PRIu64 " non-zero pairs\n", cnt, want); exit(4); }
            for (uint64_t k = 0; k < cnt; ++k)
                if (KR_CELL_SIM(cells[k]) != ref[(size_t)KR_CELL_ROW(cells[k]) * b.n() + KR_CELL_COL(cells[k])]) { fprintf(stderr, "race_check: sparse cell value\n"); exit(4); }
        }
        {   // pairs
            std::vector<uint32_t> ia, ib;
            for (uint32_t i = 0; i < b.n(); i += 3) { ia.push_back(i); ib.push_back((i * 7 + 1) % b.n()); }
            std::vector<int32_t> sim(ia.size()); std::vector<uint8_t> ident(ia.size());
            OK(kr_sim_pairs(ctx, s, ia.data(), s, ib.data(), ia.size(), sim.data(), ident.data()));
            for (size_t k = 0; k < ia.size(); ++k)
                if (sim[k] != ref[(size_t)ia[k] * b.n() + ib[k]]) { fprintf(stderr, "race_check: kr_sim_pairs (%u,%u)\n", ia[k], ib[k]); exit(4); }
        }
        kr_seqset_free(s);
        kr_destroy(ctx);
    }
}

// findClosest with ranks + greedy build on both paths; the results must agree
void case_db(int K) {
    Batch reps = families(7, 300, 12, 100, true), q = families(8, 120, 12, 100, true);
    std::vector<int32_t> sim0; std::vector<uint32_t> first0, last0; std::vector<uint8_t> added0;
    for (uint32_t path : {KR_FORCE_JOIN, KR_FORCE_INDEX}) {
        kr_ctx *ctx = nullptr;
        OK(kr_create(0, K, KR_IDENT_IS_INFINITY | path, &ctx));
        kr_db *db = nullptr;
        OK(kr_db_create(ctx, &db));
        OK(kr_db_append(ctx, db, reps.res.data(), reps.off.data(), 200));
        OK(kr_db_append(ctx, db, reps.res.data(), reps.off.data() + 200, reps.n() - 200));      // second segment / tail level
        std::vector<uint32_t> rank(reps.n());
        for (uint32_t i = 0; i < reps.n(); ++i) rank[i] = (i * 131u) % reps.n();      // 131 is coprime to every n used here? checked below
        { std::set<uint32_t> u(rank.begin(), rank.end()); if (u.size() != rank.size()) for (uint32_t i = 0; i < reps.n(); ++i) rank[i] = reps.n() - 1 - i; }
        OK(kr_db_set_ranks(ctx, db, rank.data(), reps.n()));
        kr_seqset *qs = nullptr;
        OK(kr_encode(ctx, q.res.data(), q.off.data(), q.n(), &qs));
        std::vector<int32_t> sim(q.n()); std::vector<uint32_t> first(q.n()), last(q.n());
        OK(kr_db_find_closest(ctx, db, qs, sim.data(), first.data(), last.data()));
        std::vector<uint8_t> rep(q.n());
        OK(kr_db_any_ge(ctx, db, qs, 20, rep.data()));
        for (uint32_t i = 0; i < q.n(); ++i) if ((sim[i] >= 20) != (rep[i] != 0)) { fprintf(stderr, "race_check: kr_db_any_ge vs findClosest at %u\n", i); exit(4); }
        // greedy build of the queries into a fresh database, small batches (score / resolve / accept, tail level, compaction)
        kr_db *db2 = nullptr;
        OK(kr_db_create(ctx, &db2));
        std::vector<uint8_t> added(q.n());
        OK(kr_greedy_build(ctx, db2, qs, 15, 16, added.data()));
        if (sim0.empty()) { sim0 = sim; first0 = first; last0 = last; added0 = added; }
        else if (sim0 != sim || first0 != first || last0 != last || added0 != added) { fprintf(stderr, "race_check: the two device paths disagree (db, K=%d)\n", K); exit(4); }
        kr_db_free(db2);
        kr_seqset_free(qs);
        kr_db_free(db);
        kr_destroy(ctx);
    }
}

// export block lists, import them as 3 shards, join the imported set (the path that faulted in round 1)
void case_shards() {
    Batch b = families(21, 700, 9, 80, true);
    kr_ctx *ctx = nullptr;
    OK(kr_create(0, 8, KR_IDENT_IS_INFINITY | KR_FORCE_JOIN, &ctx));
    kr_seqset *whole = nullptr;
    OK(kr_encode(ctx, b.res.data(), b.off.data(), b.n(), &whole));
    std::vector<int32_t> ref = full_matrix(ctx, whole, nullptr, b.n(), b.n());
    const uint32_t cut[4] = {0, 256, 512, b.n()};
    kr_seqset *part[3]; kr_blocklists bl[3];
    for (int k = 0; k < 3; ++k) {
        OK(kr_encode(ctx, b.res.data(), b.off.data() + cut[k], cut[k + 1] - cut[k], &part[k]));   // (offsets are absolute)
        OK(kr_seqset_export_blocklists(ctx, part[k], &bl[k]));
    }
    uint8_t *d_res = nullptr;
    cudaMalloc((void **)&d_res, b.res.size() + 16);
    cudaMemcpy(d_res, b.res.data(), b.res.size(), cudaMemcpyHostToDevice);
    kr_seqset *imp = nullptr;
    OK(kr_seqset_import_blocklists(ctx, bl, 3, d_res, b.off.data(), b.n(), &imp));
    if (full_matrix(ctx, imp, nullptr, b.n(), b.n()) != ref) { fprintf(stderr, "race_check: imported shards differ\n"); exit(4); }
    kr_seqset_free(imp);
    for (int k = 0; k < 3; ++k) kr_seqset_free(part[k]);
    kr_seqset_free(whole);
    cudaFree(d_res);
    kr_destroy(ctx);
}

// The rest of the ABI, each entry point with a synchronisation discipline of its own: rectangular matrices, row stripes,
// 16-byte sparse cells, gathered sets, key export / import, and an alphabet error in the middle (the context stays usable).
void case_misc() {
    Batch a = families(51, 300, 6, 90, true), b = families(52, 140, 6, 90, true);
    std::vector<int32_t> ref_rect, ref_self;
    for (uint32_t path : {KR_FORCE_JOIN, KR_FORCE_INDEX, 0u}) {
        kr_ctx *ctx = nullptr;
        OK(kr_create(0, 8, KR_IDENT_IS_INFINITY | path, &ctx));
        kr_seqset *sa = nullptr, *sb = nullptr;
        OK(kr_encode(ctx, a.res.data(), a.off.data(), a.n(), &sa));
        OK(kr_encode(ctx, b.res.data(), b.off.data(), b.n(), &sb));
        std::vector<int32_t> rect = full_matrix(ctx, sa, sb, a.n(), b.n());
        if (ref_rect.empty()) ref_rect = rect; else if (rect != ref_rect) { fprintf(stderr, "race_check: rectangular matrix differs between paths\n"); exit(4); }
        std::vector<int32_t> self = full_matrix(ctx, sa, nullptr, a.n(), a.n());
        if (ref_self.empty()) ref_self = self; else if (self != ref_self) { fprintf(stderr, "race_check: self matrix differs between paths\n"); exit(4); }
        {   // second stripe of the block-triangle (the index exists by now: the scan writes the zeros itself)
            kr_matrix_desc d{}; d.elem = KR_ELEM_U16; d.layout = KR_LAYOUT_BLOCKTRI; d.row_begin = 256; d.row_end = a.n();
            std::vector<uint16_t> t(kr_matrix_elems(a.n(), a.n(), KR_LAYOUT_BLOCKTRI, 256, a.n()) + 8);
            OK(kr_sim_matrix(ctx, sa, nullptr, &d, t.data()));
            size_t at = 0;
            for (uint32_t i = 256; i < a.n(); ++i) for (uint32_t j = 256; j < a.n(); ++j, ++at)
                if (t[at] != (uint16_t)ref_self[(size_t)i * a.n() + j]) { fprintf(stderr, "race_check: stripe cell (%u,%u)\n", i, j); exit(4); }
        }
        {   // 16-byte sparse cells of the rectangular matrix
            kr_matrix_desc d{}; d.elem = KR_ELEM_I32; d.layout = KR_LAYOUT_FULL; d.row_begin = 0; d.row_end = a.n();
            std::vector<uint32_t> cells((size_t)4 * a.n() * b.n());
            uint64_t cnt = 0, want = 0;
            OK(kr_sim_matrix_sparse(ctx, sa, sb, &d, 2, cells.data(), (uint64_t)a.n() * b.n(), &cnt));
            for (int32_t v : ref_rect) want += v >= 2;
            if (cnt != want) { fprintf(stderr, "race_check: %" PRIu64 " sparse cells, want %" PRIu64 "\n", cnt, want); exit(4); }
            for (uint64_t k = 0; k < cnt; ++k)
                if ((int32_t)cells[4 * k + 2] != ref_rect[(size_t)cells[4 * k] * b.n() + cells[4 * k + 1]]) { fprintf(stderr, "race_check: sparse cell\n"); exit(4); }
        }
        {   // a gathered set, its keys exported and imported again
            std::vector<uint32_t> idx;
            for (uint32_t i = 0; i < a.n(); i += 7) idx.push_back(a.n() - 1 - i);
            kr_seqset *g = nullptr;
            OK(kr_seqset_gather(ctx, sa, idx.data(), (uint32_t)idx.size(), &g));
            std::vector<int32_t> gm = full_matrix(ctx, g, sb, (uint32_t)idx.size(), b.n());
            for (size_t r = 0; r < idx.size(); ++r) for (uint32_t c = 0; c < b.n(); ++c)
                if (gm[r * b.n() + c] != ref_rect[(size_t)idx[r] * b.n() + c]) { fprintf(stderr, "race_check: gathered row %zu\n", r); exit(4); }
            std::vector<uint64_t> keys, koff{0}, roff{0}; std::vector<uint8_t> res;
            for (uint32_t i = 0; i < 24 && i < idx.size(); ++i) {
                uint32_t n = 0;
                OK(kr_seqset_keys(g, i, nullptr, 0, &n) == KR_OK || n ? KR_OK : KR_OK);
                std::vector<uint64_t> k(n + 1);
                if (n) OK(kr_seqset_keys(g, i, k.data(), n, &n));
                keys.insert(keys.end(), k.begin(), k.begin() + n); koff.push_back(keys.size());
                res.insert(res.end(), a.res.begin() + a.off[idx[i]], a.res.begin() + a.off[idx[i] + 1]); roff.push_back(res.size());
            }
            kr_seqset *fk = nullptr;
            OK(kr_seqset_from_keys(ctx, keys.data(), koff.data(), res.data(), roff.data(), (uint32_t)koff.size() - 1, &fk));
            std::vector<int32_t> fm = full_matrix(ctx, fk, sb, (uint32_t)koff.size() - 1, b.n());
            for (size_t r = 0; r + 1 < koff.size(); ++r) for (uint32_t c = 0; c < b.n(); ++c)
                if (fm[r * b.n() + c] != ref_rect[(size_t)idx[r] * b.n() + c]) { fprintf(stderr, "race_check: key-list set row %zu col %u\n", r, c); exit(4); }
            kr_seqset_free(fk); kr_seqset_free(g);
        }
        {   // a residue outside the alphabet: the call fails, the context goes on
            Batch bad = families(53, 40, 3, 60, false);
            bad.res[bad.off[17] + 3] = '#';
            kr_seqset *x = nullptr;
            if (kr_encode(ctx, bad.res.data(), bad.off.data(), bad.n(), &x) != KR_ERR_ALPHABET) { fprintf(stderr, "race_check: the bad residue went through\n"); exit(4); }
            std::vector<int32_t> again = full_matrix(ctx, sa, sb, a.n(), b.n());
            if (again != ref_rect) { fprintf(stderr, "race_check: the context is off after an alphabet error\n"); exit(4); }
        }
        kr_seqset_free(sa); kr_seqset_free(sb);
        kr_destroy(ctx);
    }
}

// What bench.py does: device-resident inputs and outputs through the _device entry points (no host synchronisation inside
// the step: encode -> matrix -> free the previous set, stream-ordered), then the e2e step with PINNED host buffers, whose
// copies are truly asynchronous -- the library has to synchronise before it returns, or the host reads below are races.
void case_bench_like() {
    Batch b = families(41, 520, 8, 90, true), q = families(42, 64, 8, 90, false);
    for (uint32_t path : {KR_FORCE_INDEX, KR_FORCE_JOIN}) {
        kr_ctx *ctx = nullptr;
        OK(kr_create(0, 8, KR_IDENT_IS_INFINITY | path, &ctx));
        uint8_t *d_res = nullptr;
        cudaMalloc((void **)&d_res, b.res.size() + 16);
        cudaMemcpy(d_res, b.res.data(), b.res.size(), cudaMemcpyHostToDevice);
        const uint64_t elems = kr_matrix_elems(b.n(), b.n(), KR_LAYOUT_BLOCKTRI, 0, b.n());
        uint16_t *d_out = nullptr;
        cudaMalloc((void **)&d_out, elems * 2 + 16);
        kr_matrix_desc d{}; d.elem = KR_ELEM_U16; d.layout = KR_LAYOUT_BLOCKTRI; d.row_begin = 0; d.row_end = b.n();
        kr_seqset *prev = nullptr;
        OK(kr_timer_begin(ctx));
        for (int step = 0; step < 3; ++step) {
            kr_seqset *cur = nullptr;
            OK(kr_encode_device(ctx, d_res, b.off.data(), b.n(), &cur));
            OK(kr_sim_matrix_device(ctx, cur, nullptr, &d, d_out));
            if (prev) kr_seqset_free(prev);
            prev = cur;
        }
        double ms = 0;
        OK(kr_timer_end(ctx, &ms));                    // synchronises
        std::vector<uint16_t> dense(elems);
        cudaMemcpy(dense.data(), d_out, elems * 2, cudaMemcpyDeviceToHost);
        // findClosest through the device entry points, chunked as the bench does
        kr_db *db = nullptr;
        OK(kr_db_create(ctx, &db));
        OK(kr_db_append_set(ctx, db, prev, nullptr, 0));
        std::vector<uint32_t> idx(300); for (uint32_t i = 0; i < 300; ++i) idx[i] = i;
        OK(kr_db_append_set(ctx, db, prev, idx.data(), 300));
        uint8_t *d_q = nullptr; int32_t *d_sim = nullptr; uint32_t *d_first = nullptr, *d_last = nullptr;
        cudaMalloc((void **)&d_q, q.res.size() + 16); cudaMemcpy(d_q, q.res.data(), q.res.size(), cudaMemcpyHostToDevice);
        cudaMalloc((void **)&d_sim, q.n() * 4); cudaMalloc((void **)&d_first, q.n() * 4); cudaMalloc((void **)&d_last, q.n() * 4);
        for (int step = 0; step < 2; ++step)
            for (uint32_t a0 = 0; a0 < q.n(); a0 += 32) {
                kr_seqset *qs = nullptr;
                OK(kr_encode_device(ctx, d_q, q.off.data() + a0, std::min(32u, q.n() - a0), &qs));
                OK(kr_db_find_closest_device(ctx, db, qs, d_sim + a0, d_first + a0, d_last + a0));
                kr_seqset_free(qs);
            }
        OK(kr_synchronize(ctx));
        std::vector<int32_t> sim(q.n());
        cudaMemcpy(sim.data(), d_sim, q.n() * 4, cudaMemcpyDeviceToHost);
        // e2e: pinned buffers
        uint8_t *h_res = nullptr; uint64_t *h_cells = nullptr;
        OK(kr_host_alloc(ctx, b.res.size() + 1, (void **)&h_res));
        OK(kr_host_alloc(ctx, 8 * 200000, (void **)&h_cells));
        memcpy(h_res, b.res.data(), b.res.size());
        uint64_t total = 0;
        for (int step = 0; step < 2; ++step) {
            kr_seqset *s = nullptr;
            OK(kr_encode(ctx, h_res, b.off.data(), b.n(), &s));
            uint64_t cnt = 0;
            kr_matrix_desc t{}; t.elem = KR_ELEM_I32; t.layout = KR_LAYOUT_TRIANGLE; t.row_begin = 0; t.row_end = b.n();
            OK(kr_sim_matrix_sparse64(ctx, s, nullptr, &t, 1, h_cells, 200000, &cnt));
            for (uint64_t k = 0; k < cnt && k < 200000; ++k) total += KR_CELL_SIM(h_cells[k]);      // the host reads the pinned result at once
            h_res[0] = h_res[0];                                                                     // ... and may write its input again
            kr_seqset_free(s);
        }
        // the dense result and the sparse cells describe the same matrix
        uint64_t dsum = 0; size_t at = 0;
        for (uint32_t i = 0; i < b.n(); ++i) for (uint32_t j = (i & ~255u); j < b.n(); ++j, ++at) if (j > i) dsum += dense[at];
        if (dsum * 2 != total) { fprintf(stderr, "race_check: dense stripe and sparse cells disagree (%" PRIu64 " vs %" PRIu64 ")\n", dsum * 2, total); exit(4); }
        kr_host_free(ctx, h_res); kr_host_free(ctx, h_cells);
        kr_db_free(db);
        kr_seqset_free(prev);
        cudaFree(d_q); cudaFree(d_sim); cudaFree(d_first); cudaFree(d_last); cudaFree(d_res); cudaFree(d_out);
        kr_destroy(ctx);
    }
}

// kr_multi: three emulated devices behind one handle.  The worker threads are real OS threads, so here ThreadSanitizer
// also checks the HOST side of the library -- the hand-over of jobs to the workers, their writes into slices of the
// caller's buffers, the shared error string -- and concurrent callers of one handle (NeighborProcessor --para).
void case_multi() {
    setenv("KR_EMUL_DEVICES", "3", 1);
    Batch b = families(31, 600, 8, 90, true), q = families(32, 90, 8, 90, true);
    kr_ctx *ctx = nullptr;       // (for the OK macro's message only)
    kr_multi *m = nullptr;
    const int devs[3] = {0, 1, 2};
    if (kr_multi_create(devs, 3, 8, KR_IDENT_IS_INFINITY, &m) != KR_OK) { fprintf(stderr, "race_check: kr_multi_create: %s\n", kr_multi_last_error(nullptr)); exit(3); }
    // reference: one plain context
    OK(kr_create(0, 8, KR_IDENT_IS_INFINITY, &ctx));
    kr_seqset *s = nullptr;
    OK(kr_encode(ctx, b.res.data(), b.off.data(), b.n(), &s));
    std::vector<int32_t> ref = full_matrix(ctx, s, nullptr, b.n(), b.n());
    std::vector<int32_t> got((size_t)b.n() * b.n());
    if (kr_multi_sim_matrix(m, b.res.data(), b.off.data(), b.n(), nullptr, nullptr, 0, KR_ELEM_I32, KR_LAYOUT_FULL, got.data()) != KR_OK || got != ref) {
        fprintf(stderr, "race_check: kr_multi_sim_matrix: %s\n", kr_multi_last_error(m)); exit(4);
    }
    kr_multi_db *mdb = nullptr;
    if (kr_multi_db_create(m, &mdb) != KR_OK || kr_multi_db_append(m, mdb, b.res.data(), b.off.data(), b.n()) != KR_OK) { fprintf(stderr, "race_check: kr_multi_db: %s\n", kr_multi_last_error(m)); exit(3); }
    kr_db *db = nullptr;
    OK(kr_db_create(ctx, &db));
    OK(kr_db_append(ctx, db, b.res.data(), b.off.data(), b.n()));
    kr_seqset *qs = nullptr;
    OK(kr_encode(ctx, q.res.data(), q.off.data(), q.n(), &qs));
    std::vector<int32_t> sim(q.n()); std::vector<uint32_t> first(q.n()), last(q.n());
    OK(kr_db_find_closest(ctx, db, qs, sim.data(), first.data(), last.data()));
    // two caller threads on the one handle at the same time
    auto caller = [&](int who) {
        std::vector<int32_t> msim(q.n()); std::vector<uint32_t> mfirst(q.n()), mlast(q.n());
        for (int rep = 0; rep < 2; ++rep) {
            if (kr_multi_db_find_closest(m, mdb, q.res.data(), q.off.data(), q.n(), msim.data(), mfirst.data(), mlast.data()) != KR_OK ||
                msim != sim || mfirst != first || mlast != last) { fprintf(stderr, "race_check: kr_multi_db_find_closest (caller %d): %s\n", who, kr_multi_last_error(m)); exit(4); }
        }
    };
    std::thread t1(caller, 1), t2(caller, 2);
    t1.join(); t2.join();
    // greedy build over the handle == over one context
    std::vector<uint8_t> a1(q.n()), a2(q.n());
    kr_db *db2 = nullptr; kr_multi_db *mdb2 = nullptr;
    OK(kr_db_create(ctx, &db2));
    OK(kr_greedy_build(ctx, db2, qs, 15, 16, a1.data()));
    if (kr_multi_db_create(m, &mdb2) != KR_OK || kr_multi_greedy_build(m, mdb2, q.res.data(), q.off.data(), q.n(), 15, 16, a2.data()) != KR_OK || a1 != a2) {
        fprintf(stderr, "race_check: kr_multi_greedy_build: %s\n", kr_multi_last_error(m)); exit(4);
    }
    kr_multi_db_free(mdb2); kr_multi_db_free(mdb);
    kr_db_free(db2); kr_db_free(db); kr_seqset_free(qs); kr_seqset_free(s);
    kr_destroy(ctx);
    kr_multi_destroy(m);
}

// Positive controls (kernels of this file, not of the product): the detector must report each of these, or a silent run
// of the product cases means nothing.  `tools/emul_race.sh` runs them separately and requires a report from every one.
__global__ void k_ctl_missing_syncthreads(u32 *out) {
    __shared__ u32 buf[64];
    buf[threadIdx.x] = threadIdx.x;
    out[threadIdx.x] = buf[threadIdx.x ^ 32];          // another warp's word, no __syncthreads in between
}
__global__ void k_ctl_missing_syncwarp(u32 *out) {
    __shared__ u32 buf[32];
    buf[threadIdx.x] = threadIdx.x;
    const u32 v = __shfl_xor_sync(0xffffffffu, threadIdx.x, 1);   // a shuffle orders no memory
    out[threadIdx.x] = buf[threadIdx.x ^ 1] + v;
}
__global__ void k_ctl_global_conflict(u32 *out) { out[threadIdx.x >> 1] = threadIdx.x; }   // two threads, one word, different values
__global__ void k_ctl_clean(u32 *out) {                // the same exchanges, ordered: must stay silent
    __shared__ u32 buf[64];
    buf[threadIdx.x] = threadIdx.x;
    __syncthreads();
    u32 v = buf[threadIdx.x ^ 32];
    __syncthreads();
    buf[threadIdx.x] = v;
    __syncwarp();
    v += buf[threadIdx.x ^ 1];
    atomicAdd(&out[0], v);
    volatile u32 *flag = out + 1;
    if (threadIdx.x == 0) *flag = 1; else (void)*flag;  // volatile accesses are relaxed atomics
}
__global__ void k_ctl_grid_conflict(u32 *out) { if (threadIdx.x == 0) out[0] = blockIdx.x; }   // two CTAs, one word: a race only between CTAs
void case_control(int which) {
    u32 *d = nullptr;
    cudaMalloc((void **)&d, 64 * 4);
    cudaMemset(d, 0, 64 * 4);
    if (which == 0) KR_LAUNCH(k_ctl_missing_syncthreads, 1, 64, 0, nullptr, d);
    if (which == 1) KR_LAUNCH(k_ctl_missing_syncwarp, 1, 32, 0, nullptr, d);
    if (which == 2) KR_LAUNCH(k_ctl_global_conflict, 1, 64, 0, nullptr, d);
    if (which == 3) KR_LAUNCH(k_ctl_clean, 2, 64, 0, nullptr, d);
    if (which == 4) KR_LAUNCH(k_ctl_grid_conflict, 2, 32, 0, nullptr, d);
    cudaFree(d);       // (synchronises the device)
}

// Stream-order controls: the emulator executes everything at once, so each of these computes the "right" values -- and each
// is a bug on a device that the detector must report: a host read before the synchronisation, two streams on one buffer
// without an event, a stream-ordered free that is not ordered after the last use.  The last one is their ordered twin.
__global__ void k_ctl_fill(u32 *out, u32 v) { out[threadIdx.x] = v + threadIdx.x; }
u32 case_stream_control(int which) {
    cudaStream_t a, b;
    cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    cudaEvent_t ev; cudaEventCreate(&ev);
    u32 *d = nullptr, *h = nullptr;
    cudaMallocAsync((void **)&d, 64 * 4, a);
    cudaHostAlloc((void **)&h, 64 * 4, cudaHostAllocDefault);
    u32 sum = 0;
    KR_LAUNCH(k_ctl_fill, 1, 64, 0, a, d, 7u);
    if (which == 0) {            // pinned destination: the copy is asynchronous, the host reads too early
        cudaMemcpyAsync(h, d, 64 * 4, cudaMemcpyDeviceToHost, a);
        for (int i = 0; i < 64; ++i) sum += h[i];
        cudaStreamSynchronize(a);
    } else if (which == 1) {     // a second stream writes the buffer with no event between the two
        KR_LAUNCH(k_ctl_fill, 1, 64, 0, b, d, 9u);
        cudaStreamSynchronize(a); cudaStreamSynchronize(b);
    } else if (which == 4) {     // the same with wide launches (NT threads x NG CTAs per stream: far more CUDA threads than the detector has slots)
        const char *e = getenv("NT"), *g = getenv("NG");
        const unsigned nt = e ? atoi(e) : 512, ng = g ? atoi(g) : 4;
        u32 *big = nullptr;
        cudaMallocAsync((void **)&big, 4096 * 4, a);
        KR_LAUNCH(k_ctl_fill, ng, nt, 0, a, big, 7u);
        KR_LAUNCH(k_ctl_fill, ng, nt, 0, b, big, 9u);
        cudaStreamSynchronize(a); cudaStreamSynchronize(b);
        cudaFreeAsync(big, a);
    } else if (which == 2) {     // freed in stream b's order while stream a's kernel may still be running
        cudaFreeAsync(d, b); d = nullptr;
        cudaStreamSynchronize(a); cudaStreamSynchronize(b);
    } else {                     // the ordered twin of all three
        cudaEventRecord(ev, a);
        cudaStreamWaitEvent(b, ev, 0);
        KR_LAUNCH(k_ctl_fill, 1, 64, 0, b, d, 9u);
        cudaMemcpyAsync(h, d, 64 * 4, cudaMemcpyDeviceToHost, b);
        cudaStreamSynchronize(b);
        for (int i = 0; i < 64; ++i) sum += h[i];
        cudaFreeAsync(d, b); d = nullptr;
        cudaStreamSynchronize(b);
    }
    if (d) { cudaFreeAsync(d, a); cudaStreamSynchronize(a); }
    cudaFreeHost(h);
    cudaEventDestroy(ev); cudaStreamDestroy(a); cudaStreamDestroy(b);
    return sum;
}

}  // namespace

// last object of the static shared-memory section (see static_smem_renew in cuda_emul.h)
extern "C" { __attribute__((section("kr_emu_smem"), aligned(4096), used)) char kr_emu_smem_tail[8192]; }

int main(int argc, char **argv) {
    std::set<std::string> want;
    for (int i = 1; i < argc; ++i) want.insert(argv[i]);
    auto run = [&](const char *name, auto fn) {
        if (!want.empty() && !want.count(name)) return;
        fprintf(stderr, "race_check: %s\n", name);
        fn();
    };
    if (want.count("smoke")) { fprintf(stderr, "race_check: smoke\n"); case_matrix(8, 60, true); }   // the CPU suite's quick pass
    run("matrix_k8", [] { case_matrix(8, 300, true); });
    run("matrix_k10", [] { case_matrix(10, 120, true); });      // 32-bit low words in the tile join
    run("matrix_k3", [] { case_matrix(3, 270, false); });       // few distinct keys: long posting lists, crowded buckets, multi-pass queries
    run("db_k8", [] { case_db(8); });
    run("shards", [] { case_shards(); });
    run("multi", [] { case_multi(); });
    run("bench_like", [] { case_bench_like(); });
    run("misc", [] { case_misc(); });
    // controls run only when named (each of the first three MUST be reported)
    if (want.count("ctl_syncthreads")) case_control(0);
    if (want.count("ctl_syncwarp")) case_control(1);
    if (want.count("ctl_global")) case_control(2);
    if (want.count("ctl_clean")) case_control(3);
    if (want.count("ctl_host_early")) case_stream_control(0);
    if (want.count("ctl_two_streams")) case_stream_control(1);
    if (want.count("ctl_free_order")) case_stream_control(2);
    if (want.count("ctl_two_streams_wide")) case_stream_control(4);
    if (want.count("ctl_streams_clean")) case_stream_control(3);
    if (want.count("ctl_grid")) case_control(4);       // reported under KR_EMUL_RACE=grid only
    fprintf(stderr, "race_check: done\n");
    return 0;
}
