/* synth.c -- deterministic synthetic seed-protein workloads (BASELINE.json configs 3-5,
 * SURVEY.md 8d): families of mutated ~300-aa proteins so that similarities span the
 * representation threshold.  Integer arithmetic only (SplitMix64), so every platform
 * generates the same bytes.  Built as libkrsynth.so; used by bench.py and tests. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rnext(rng_t *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint32_t rbelow(rng_t *r, uint32_t n) { return (uint32_t)(((rnext(r) >> 32) * (uint64_t)n) >> 32); }

/* residue counts of data/rep50.db, total 518,976 */
static const char AA[21] = "LEGARIVDKFTPSNQMHYCW";
static const uint32_t CNT[20] = {48827, 42607, 37306, 35185, 33689, 32513, 31252, 30691, 30145, 29359,
                                 25571, 25444, 25220, 19059, 18890, 15768, 13392, 13332, 6459, 4267};
static uint32_t CUM[20];
static void init_cum(void) { uint32_t c = 0; for (int i = 0; i < 20; i++) { c += CNT[i]; CUM[i] = c; } }
static inline char raa(rng_t *r) {
    uint32_t x = rbelow(r, 518976);
    int lo = 0, hi = 19;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (x < CUM[mid]) hi = mid; else lo = mid + 1; }
    return AA[lo];
}

#define KRS_MAXLEN 640
/* worst-case bytes for nseq sequences */
uint64_t krs_capacity(uint32_t nseq) { return (uint64_t)nseq * KRS_MAXLEN; }

/* Generates nseq sequences in nfam families.  Sequence i < nfam is the root of family i; every later
 * one picks a family uniformly, then a uniformly chosen earlier member of it, and is either an exact
 * duplicate (dup_per_10k / 10000) or a mutated copy: per-residue substitution rate drawn from
 * {1,3,6,10,15,25}%, indels with probability 0.2% per position (length 1-3).  xrun_per_10k of the
 * sequences get a run of 20-30 'X' (exercises the deferral rule, RepGenomeDbProcessor.java:136-137).
 * Returns total residues, or 0 when cap is too small.  family[i] (may be NULL) = family of i. */
uint64_t krs_generate(uint64_t seed, uint32_t nseq, uint32_t nfam, uint32_t dup_per_10k, uint32_t xrun_per_10k,
                      uint8_t *res, uint64_t cap, uint64_t *off, uint32_t *family) {
    static const uint32_t RATE[6] = {655, 1966, 3932, 6554, 9830, 16384}; /* p * 65536 */
    init_cum();
    if (nfam == 0) nfam = 1;
    if (nfam > nseq) nfam = nseq;
    rng_t r = {seed};
    uint32_t *fam = (uint32_t *)malloc(sizeof(uint32_t) * (nseq ? nseq : 1));
    uint32_t *fcount = (uint32_t *)calloc(nfam + 1, sizeof(uint32_t));
    for (uint32_t i = 0; i < nseq; i++) { fam[i] = i < nfam ? i : rbelow(&r, nfam); fcount[fam[i] + 1]++; }
    for (uint32_t f = 0; f < nfam; f++) fcount[f + 1] += fcount[f];
    uint32_t *members = (uint32_t *)malloc(sizeof(uint32_t) * (nseq ? nseq : 1));
    uint32_t *filled = (uint32_t *)calloc(nfam, sizeof(uint32_t));
    uint64_t pos = 0;
    off[0] = 0;
    for (uint32_t i = 0; i < nseq; i++) {
        if (pos + KRS_MAXLEN > cap) { pos = 0; goto done; }
        uint8_t *out = res + pos;
        uint32_t f = fam[i], len = 0;
        if (filled[f] == 0) {
            uint32_t s = 0;
            uint64_t x = 0;
            for (int k = 0; k < 12; k++) { if ((k & 3) == 0) x = rnext(&r); s += (uint32_t)(x & 0xFFFF); x >>= 16; }
            int32_t L = 300 + ((int32_t)s - 393210) * 25 / 65536;
            if (L < 220) L = 220;
            if (L > 400) L = 400;
            for (int32_t k = 0; k < L; k++) out[len++] = (uint8_t)raa(&r);
        } else {
            uint32_t parent = members[fcount[f] + rbelow(&r, filled[f])];
            const uint8_t *src = res + off[parent];
            uint32_t plen = (uint32_t)(off[parent + 1] - off[parent]);
            if (rbelow(&r, 10000) < dup_per_10k) {
                memcpy(out, src, plen); len = plen;
            } else {
                uint32_t rate = RATE[rbelow(&r, 6)];
                for (uint32_t k = 0; k < plen && len < KRS_MAXLEN - 4; k++) {
                    uint64_t x = rnext(&r);
                    uint32_t a = (uint32_t)(x & 0xFFFF), b = (uint32_t)((x >> 16) & 0xFFFF);
                    if (b < 131) {                                   /* 0.2% indel */
                        if (x & (1ull << 32)) continue;              /* deletion */
                        uint32_t n = 1 + (uint32_t)((x >> 33) % 3);  /* insertion of 1-3 */
                        for (uint32_t t = 0; t < n; t++) out[len++] = (uint8_t)raa(&r);
                    }
                    out[len++] = (a < rate) ? (uint8_t)raa(&r) : src[k];
                }
            }
        }
        if (xrun_per_10k && rbelow(&r, 10000) < xrun_per_10k && len > 60) {
            uint32_t n = 20 + rbelow(&r, 11), at = rbelow(&r, len - n);
            memset(out + at, 'X', n);
        }
        members[fcount[f] + filled[f]++] = i;
        if (family) family[i] = f;
        pos += len;
        off[i + 1] = pos;
    }
done:
    free(fam); free(fcount); free(members); free(filled);
    return pos;
}
