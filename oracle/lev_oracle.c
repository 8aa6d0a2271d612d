/*
 * oracle/lev_oracle.c — CPU restatement of the StringWars `similarities` hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under stringwars_b200/ may link, import or call this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker or as the timed CPU baseline — never as the product path.
 *
 * PARITY STATUS: **parity unpinned by the reference.** /root/reference holds no golden vectors,
 * known-answer tests or fixtures for this path, and the algorithm itself lives in
 * third-party crates that are absent from /root/reference and cannot be built here (no Rust, no network):
 *   rapidfuzz 0.5.0   (Cargo.lock:3939-3942)  — `rapidfuzz::distance::levenshtein::distance`
 *   bio       4.0.1   (Cargo.lock:510-513)    — `bio::alignment::distance::levenshtein`, `pairwise::Aligner`
 *   stringzilla 5.0.1 (Cargo.lock:4678-4685)  — `szs::{LevenshteinDistances, NeedlemanWunschScores, ...}`
 * What pins this oracle instead (tests/test_oracle.py): literature known answers, an independent pure
 * Python implementation (oracle/pyref.py), a third-party implementation that IS in this image
 * (torchaudio 2.11 `functional.edit_distance`, fixtures in tests/golden/ made by scripts/make_golden.py),
 * cross-agreement of the two algorithm families below, and algebraic properties (hypothesis).
 *
 * Two algorithm families, because the reference times both kinds:
 *   orc_lev_*      textbook Wagner-Fischer two-row DP — the definition; same algorithm class as
 *                  `bio_distance::levenshtein(a, b)`            (similarities/bench.rs:455)
 *   orc_rf_lev_*   "rapidfuzz-class" restatement of the published algorithm behind
 *                  `levenshtein::distance(a.iter().copied(), b.iter().copied())` (similarities/bench.rs:416-419):
 *                  strip common affix, build a 256-entry pattern-match bit table for the shorter string,
 *                  Hyyrö-2003 bit-parallel single word when it fits 64 symbols, Myers/Hyyrö blocks
 *                  otherwise. rapidfuzz's extra speed paths (mbleven for cutoffs < 4, Ukkonen band +
 *                  score-hint doubling) only change speed, not results, and are NOT restated; this
 *                  restatement keeps scratch thread-local instead of allocating per call, so it is a
 *                  conservative (faster-than-real) stand-in when used as the timed CPU baseline.
 *   orc_nw_/orc_sw_ Needleman-Wunsch / Smith-Waterman scores with a 32-class substitution table and
 *                  linear or affine (Gotoh) gaps — semantics of `NeedlemanWunschScores::new(&scope,
 *                  &byte_to_class, &class_costs, open, extend)` (similarities/bench.rs:655-674, 98-108)
 *                  and of the `bio` baseline `Aligner::with_capacity(..).global(a,b).score` (:746-768).
 *                  The two contenders disagree on what a length-k gap costs; `convention` selects:
 *                    0 = stringzilla: open + (k-1)*extend      1 = rust-bio: open + k*extend
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ textbook Wagner-Fischer */

/* Unit-cost Levenshtein over bytes: D[i][j] = min(D[i-1][j-1] + (a!=b), D[i-1][j] + 1, D[i][j-1] + 1).
 * Definition of the value returned at similarities/bench.rs:416-419 and :455. */
ORC_EXPORT uint64_t orc_lev_bytes(const uint8_t *a, size_t la, const uint8_t *b, size_t lb) {
    if (la == 0) return lb;
    if (lb == 0) return la;
    uint32_t stack_row[1025];
    uint32_t *row = lb + 1 <= 1025 ? stack_row : (uint32_t *)malloc((lb + 1) * sizeof(uint32_t));
    for (size_t j = 0; j <= lb; ++j) row[j] = (uint32_t)j;
    for (size_t i = 1; i <= la; ++i) {
        uint32_t diag = row[0];
        row[0] = (uint32_t)i;
        uint8_t ca = a[i - 1];
        for (size_t j = 1; j <= lb; ++j) {
            uint32_t up = row[j];
            uint32_t sub = diag + (ca != b[j - 1]);
            uint32_t best = up < row[j - 1] ? up : row[j - 1];
            best += 1;
            row[j] = sub < best ? sub : best;
            diag = up;
        }
    }
    uint64_t result = row[lb];
    if (row != stack_row) free(row);
    return result;
}

/* Lenient UTF-8 symbol split, identical to the device decoder: a symbol starts at every byte that is not a
 * continuation byte (10xxxxxx) and its value is the big-endian packing of that lead byte and up to three
 * following continuation bytes. For valid UTF-8 this is a bijection with Unicode scalar values, so
 * comparing packed symbols == comparing `str::chars()` items (similarities/bench.rs:432-437). */
ORC_EXPORT size_t orc_utf8_symbols(const uint8_t *s, size_t n, uint32_t *out) {
    size_t count = 0;
    size_t i = 0;
    while (i < n && (s[i] & 0xC0) == 0x80) ++i; /* stray leading continuation bytes: ignored */
    while (i < n) {
        uint32_t sym = s[i++];
        int taken = 1;
        while (i < n && (s[i] & 0xC0) == 0x80) {
            if (taken < 4) sym = (sym << 8) | s[i];
            ++taken;
            ++i;
        }
        if (out) out[count] = sym;
        ++count;
    }
    return count;
}

static uint64_t lev_u32(const uint32_t *a, size_t la, const uint32_t *b, size_t lb) {
    if (la == 0) return lb;
    if (lb == 0) return la;
    uint32_t *row = (uint32_t *)malloc((lb + 1) * sizeof(uint32_t));
    for (size_t j = 0; j <= lb; ++j) row[j] = (uint32_t)j;
    for (size_t i = 1; i <= la; ++i) {
        uint32_t diag = row[0];
        row[0] = (uint32_t)i;
        for (size_t j = 1; j <= lb; ++j) {
            uint32_t up = row[j];
            uint32_t sub = diag + (a[i - 1] != b[j - 1]);
            uint32_t best = (up < row[j - 1] ? up : row[j - 1]) + 1;
            row[j] = sub < best ? sub : best;
            diag = up;
        }
    }
    uint64_t result = row[lb];
    free(row);
    return result;
}

/* Codepoint-level Levenshtein (`levenshtein::distance(a_str.chars(), b_str.chars())`, bench.rs:437). */
ORC_EXPORT uint64_t orc_lev_utf8(const uint8_t *a, size_t la, const uint8_t *b, size_t lb) {
    uint32_t *sa = (uint32_t *)malloc((la + 1) * sizeof(uint32_t));
    uint32_t *sb = (uint32_t *)malloc((lb + 1) * sizeof(uint32_t));
    size_t na = orc_utf8_symbols(a, la, sa);
    size_t nb = orc_utf8_symbols(b, lb, sb);
    uint64_t d = lev_u32(sa, na, sb, nb);
    free(sa);
    free(sb);
    return d;
}

/* ------------------------------------------------------------------ rapidfuzz-class bit-parallel */

static __thread uint64_t *tls_pm = NULL; /* words x 256 pattern-match table */
static __thread size_t tls_pm_words = 0;
static __thread uint64_t *tls_vecs = NULL; /* 2 x words (VP, VN) */

static void rf_reserve(size_t words) {
    if (words <= tls_pm_words) return;
    free(tls_pm);
    free(tls_vecs);
    tls_pm = (uint64_t *)malloc(words * 256 * sizeof(uint64_t));
    tls_vecs = (uint64_t *)malloc(words * 2 * sizeof(uint64_t));
    tls_pm_words = words;
}

/* Hyyrö 2003, pattern fits one 64-bit word. `pm` is the 256-entry table of the pattern. */
static uint64_t rf_hyrroe2003_word(const uint64_t *pm, size_t m, const uint8_t *t, size_t n) {
    uint64_t vp = ~0ULL, vn = 0, dist = m;
    const uint64_t last = 1ULL << (m - 1);
    for (size_t j = 0; j < n; ++j) {
        uint64_t x = pm[t[j]];
        uint64_t d0 = (((x & vp) + vp) ^ vp) | x | vn;
        uint64_t hp = vn | ~(d0 | vp);
        uint64_t hn = d0 & vp;
        dist += (hp & last) != 0;
        dist -= (hn & last) != 0;
        hp = (hp << 1) | 1;
        hn <<= 1;
        vp = hn | ~(d0 | hp);
        vn = hp & d0;
    }
    return dist;
}

/* Myers 1999 / Hyyrö 2003 blocks, full band. `pm` is words x 256. */
static uint64_t rf_hyrroe2003_block(const uint64_t *pm, size_t words, size_t m, const uint8_t *t, size_t n) {
    uint64_t *vp = tls_vecs, *vn = tls_vecs + words;
    for (size_t w = 0; w < words; ++w) vp[w] = ~0ULL, vn[w] = 0;
    uint64_t dist = m;
    const uint64_t last = 1ULL << ((m - 1) & 63);
    for (size_t j = 0; j < n; ++j) {
        uint64_t hp_carry = 1, hn_carry = 0;
        const uint8_t c = t[j];
        for (size_t w = 0; w < words; ++w) {
            uint64_t x = pm[w * 256 + c] | hn_carry;
            uint64_t d0 = (((x & vp[w]) + vp[w]) ^ vp[w]) | x | vn[w];
            uint64_t hp = vn[w] | ~(d0 | vp[w]);
            uint64_t hn = d0 & vp[w];
            if (w + 1 == words) {
                dist += (hp & last) != 0;
                dist -= (hn & last) != 0;
            }
            uint64_t hp_out = hp >> 63, hn_out = hn >> 63;
            hp = (hp << 1) | hp_carry;
            hn = (hn << 1) | hn_carry;
            hp_carry = hp_out;
            hn_carry = hn_out;
            vp[w] = hn | ~(d0 | hp);
            vn[w] = hp & d0;
        }
    }
    return dist;
}

/* Restatement of the call at similarities/bench.rs:416-419 (see header for what is and is not restated). */
ORC_EXPORT uint64_t orc_rf_lev_bytes(const uint8_t *a, size_t la, const uint8_t *b, size_t lb) {
    /* common prefix / suffix removal */
    while (la && lb && a[0] == b[0]) ++a, ++b, --la, --lb;
    while (la && lb && a[la - 1] == b[lb - 1]) --la, --lb;
    if (la == 0) return lb;
    if (lb == 0) return la;
    /* pattern = shorter string */
    const uint8_t *p = a, *t = b;
    size_t m = la, n = lb;
    if (m > n) p = b, t = a, m = lb, n = la;
    size_t words = (m + 63) / 64;
    rf_reserve(words);
    memset(tls_pm, 0, words * 256 * sizeof(uint64_t));
    for (size_t i = 0; i < m; ++i) tls_pm[(i >> 6) * 256 + p[i]] |= 1ULL << (i & 63);
    if (words == 1) return rf_hyrroe2003_word(tls_pm, m, t, n);
    return rf_hyrroe2003_block(tls_pm, words, m, t, n);
}

/* ------------------------------------------------------------------ NW / SW scores */

#define ORC_NEG (-(1LL << 40))

/* Global (local=0) or local (local=1) alignment SCORE, maximising.
 * byte_to_class: 256 entries (bench.rs:99-102), costs: 32x32 row-major i8 (bench.rs:103-106).
 * Gap of length k costs  convention==0: open + (k-1)*extend   (stringzilla wording, bench.rs:7, 342, 353)
 *                        convention==1: open + k*extend       (rust-bio `Aligner`, bench.rs:746-753)      */
ORC_EXPORT int64_t orc_align_score(const uint8_t *a, size_t la, const uint8_t *b, size_t lb,
                                   const uint8_t *byte_to_class, const int8_t *costs, int open, int extend,
                                   int convention, int local) {
    const int64_t first = convention == 0 ? open : (int64_t)open + extend; /* cost of the first gap symbol */
    const int64_t next = extend;                                           /* cost of each further one */
    int64_t *H = (int64_t *)malloc((lb + 1) * sizeof(int64_t));          /* best ending anywhere */
    int64_t *E = (int64_t *)malloc((lb + 1) * sizeof(int64_t));          /* best ending with a gap in `a` (vertical) */
    int64_t best_local = 0;
    H[0] = 0;
    E[0] = ORC_NEG;
    for (size_t j = 1; j <= lb; ++j) {
        H[j] = local ? 0 : first + (int64_t)(j - 1) * next;
        E[j] = ORC_NEG;
    }
    for (size_t i = 1; i <= la; ++i) {
        int64_t diag = H[0];
        H[0] = local ? 0 : first + (int64_t)(i - 1) * next;
        int64_t F = ORC_NEG; /* best ending with a gap in `b` (horizontal) */
        const int8_t *row = costs + 32 * byte_to_class[a[i - 1]];
        for (size_t j = 1; j <= lb; ++j) {
            int64_t e1 = E[j] + next, e2 = H[j] + first;
            E[j] = e1 > e2 ? e1 : e2;
            int64_t f1 = F + next, f2 = H[j - 1] + first;
            F = f1 > f2 ? f1 : f2;
            int64_t h = diag + row[byte_to_class[b[j - 1]]];
            if (E[j] > h) h = E[j];
            if (F > h) h = F;
            if (local && h < 0) h = 0;
            diag = H[j];
            H[j] = h;
            if (h > best_local) best_local = h;
        }
    }
    int64_t result = local ? best_local : H[lb];
    free(H);
    free(E);
    return result;
}

/* ------------------------------------------------------------------ batched drivers (tapes) */

typedef struct {
    int kind; /* 0 lev WF, 1 lev rapidfuzz-class, 2 lev utf8, 3 align score */
    const uint8_t *a_data, *b_data;
    const uint64_t *a_offs, *b_offs;
    size_t lo, hi;   /* pair range (pairs mode) or flat cell range (cross mode) */
    int repeats;     /* timing runs: walk the range this many times (>= 1) */
    size_t n_c;      /* cross mode: candidates per query; 0 = element-wise pairs */
    size_t stride;   /* cross mode: output row stride in elements */
    int64_t *out;
    const uint8_t *byte_to_class;
    const int8_t *costs;
    int open, extend, convention, local;
} orc_job_t;

static void *orc_worker(void *arg) {
    orc_job_t *job = (orc_job_t *)arg;
    for (int rep = 0; rep < (job->repeats > 0 ? job->repeats : 1); ++rep)
    for (size_t k = job->lo; k < job->hi; ++k) {
        size_t ia = job->n_c ? k / job->n_c : k, ib = job->n_c ? k % job->n_c : k;
        const uint8_t *a = job->a_data + job->a_offs[ia];
        const uint8_t *b = job->b_data + job->b_offs[ib];
        size_t la = job->a_offs[ia + 1] - job->a_offs[ia], lb = job->b_offs[ib + 1] - job->b_offs[ib];
        int64_t r;
        switch (job->kind) {
        case 0: r = (int64_t)orc_lev_bytes(a, la, b, lb); break;
        case 1: r = (int64_t)orc_rf_lev_bytes(a, la, b, lb); break;
        case 2: r = (int64_t)orc_lev_utf8(a, la, b, lb); break;
        default:
            r = orc_align_score(a, la, b, lb, job->byte_to_class, job->costs, job->open, job->extend,
                                job->convention, job->local);
        }
        job->out[job->n_c ? ia * job->stride + ib : k] = r;
    }
    return NULL;
}

static void orc_run(orc_job_t proto, size_t total, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > total) threads = total ? (int)total : 1;
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    orc_job_t *jobs = (orc_job_t *)malloc(sizeof(orc_job_t) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = proto;
        jobs[t].lo = total * t / threads;
        jobs[t].hi = total * (t + 1) / threads;
        if (threads == 1) orc_worker(&jobs[t]);
        else pthread_create(&tids[t], NULL, orc_worker, &jobs[t]);
    }
    if (threads > 1)
        for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
    free(tids);
    free(jobs);
}

/* out[i] = distance(a[i], b[i]); kind: 0 Wagner-Fischer, 1 rapidfuzz-class, 2 codepoint-level. */
ORC_EXPORT void orc_lev_pairs(int kind, const uint8_t *a_data, const uint64_t *a_offs, const uint8_t *b_data,
                              const uint64_t *b_offs, size_t n_pairs, uint64_t *out, int threads) {
    orc_job_t job = {0};
    job.kind = kind, job.a_data = a_data, job.a_offs = a_offs, job.b_data = b_data, job.b_offs = b_offs;
    job.out = (int64_t *)out;
    orc_run(job, n_pairs, threads);
}

/* out[q * stride + c] = distance(queries[q], candidates[c]) — the `compute_into` shape (bench.rs:599-603). */
ORC_EXPORT void orc_lev_cross(int kind, const uint8_t *q_data, const uint64_t *q_offs, size_t n_q,
                              const uint8_t *c_data, const uint64_t *c_offs, size_t n_c, uint64_t *out,
                              size_t stride, int threads) {
    orc_job_t job = {0};
    job.kind = kind, job.a_data = q_data, job.a_offs = q_offs, job.b_data = c_data, job.b_offs = c_offs;
    job.out = (int64_t *)out, job.n_c = n_c, job.stride = stride;
    if (n_c) orc_run(job, n_q * n_c, threads);
}

ORC_EXPORT void orc_align_pairs(const uint8_t *a_data, const uint64_t *a_offs, const uint8_t *b_data,
                                const uint64_t *b_offs, size_t n_pairs, const uint8_t *byte_to_class,
                                const int8_t *costs, int open, int extend, int convention, int local,
                                int64_t *out, int threads) {
    orc_job_t job = {0};
    job.kind = 3, job.a_data = a_data, job.a_offs = a_offs, job.b_data = b_data, job.b_offs = b_offs;
    job.out = out, job.byte_to_class = byte_to_class, job.costs = costs;
    job.open = open, job.extend = extend, job.convention = convention, job.local = local;
    orc_run(job, n_pairs, threads);
}

ORC_EXPORT void orc_align_cross(const uint8_t *q_data, const uint64_t *q_offs, size_t n_q, const uint8_t *c_data,
                                const uint64_t *c_offs, size_t n_c, const uint8_t *byte_to_class,
                                const int8_t *costs, int open, int extend, int convention, int local,
                                int64_t *out, size_t stride, int threads) {
    orc_job_t job = {0};
    job.kind = 3, job.a_data = q_data, job.a_offs = q_offs, job.b_data = c_data, job.b_offs = c_offs;
    job.out = out, job.n_c = n_c, job.stride = stride, job.byte_to_class = byte_to_class, job.costs = costs;
    job.open = open, job.extend = extend, job.convention = convention, job.local = local;
    if (n_c) orc_run(job, n_q * n_c, threads);
}

/* Timed variant for the CPU baseline: every thread walks its share of the batch `repeats` times; returns the
 * wall seconds between thread creation and the last join (so start-up cost is amortised, not hidden). */
ORC_EXPORT double orc_lev_pairs_timed(int kind, const uint8_t *a_data, const uint64_t *a_offs, const uint8_t *b_data,
                                      const uint64_t *b_offs, size_t n_pairs, uint64_t *out, int threads, int repeats) {
    orc_job_t job = {0};
    job.kind = kind, job.a_data = a_data, job.a_offs = a_offs, job.b_data = b_data, job.b_offs = b_offs;
    job.out = (int64_t *)out, job.repeats = repeats;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    orc_run(job, n_pairs, threads);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
