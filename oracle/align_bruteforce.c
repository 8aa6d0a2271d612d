/*
 * align_bruteforce.c — DEFINITION-LEVEL checker for alignment scores. TEST INFRASTRUCTURE ONLY.
 *
 * Shares nothing with the Gotoh recurrences of lev_oracle.c or of the CUDA kernels: it ENUMERATES every alignment
 * of two short strings column by column (substitution, gap in a, gap in b) and scores each one by the definition
 *     score = sum of substitution scores + sum over maximal gap runs of (first + (k - 1) * next),
 * where a run is a maximal sequence of consecutive columns with a gap in the SAME string. The two gap conventions of
 * the reference's contenders differ only in `first`:
 *     stringzilla  `NeedlemanWunschScores::new(.., open, extend)`       first = open,          next = extend
 *                  (StringWars similarities/bench.rs:658-662)
 *     bio          `Aligner::with_capacity(.., open, extend, ..)`       first = open + extend, next = extend
 *                  (bench.rs:746)
 * Global: the best alignment of the whole strings. Local: the best global alignment over every pair of substrings,
 * never below 0 (the empty alignment). Exponential time: strings of up to ~8 symbols.
 * Valid for |next| <= |first| (every model the harness uses): when extending is dearer than opening, a recurrence may
 * re-open inside a run, which "maximal run" scoring by definition does not.
 */
#include <stddef.h>
#include <stdint.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct {
    const uint8_t *a, *b;
    size_t la, lb;
    const uint8_t *byte_to_class;
    const int8_t *costs; /* 32 x 32 */
    int first, next;
    int64_t best;
} brute_t;

enum { COL_NONE = 0, COL_SUBST = 1, COL_GAP_IN_A = 2, COL_GAP_IN_B = 3 };

static void enumerate(brute_t *ctx, size_t i, size_t j, int prev, int64_t score) {
    if (i == ctx->la && j == ctx->lb) {
        if (score > ctx->best) ctx->best = score;
        return;
    }
    if (i < ctx->la && j < ctx->lb) {
        const int s = ctx->costs[32 * ctx->byte_to_class[ctx->a[i]] + ctx->byte_to_class[ctx->b[j]]];
        enumerate(ctx, i + 1, j + 1, COL_SUBST, score + s);
    }
    if (i < ctx->la) /* a[i] faces a gap in b */
        enumerate(ctx, i + 1, j, COL_GAP_IN_B, score + (prev == COL_GAP_IN_B ? ctx->next : ctx->first));
    if (j < ctx->lb) /* b[j] faces a gap in a */
        enumerate(ctx, i, j + 1, COL_GAP_IN_A, score + (prev == COL_GAP_IN_A ? ctx->next : ctx->first));
}

static int64_t brute_global(const uint8_t *a, size_t la, const uint8_t *b, size_t lb, const uint8_t *byte_to_class, const int8_t *costs,
                            int first, int next) {
    brute_t ctx = {a, b, la, lb, byte_to_class, costs, first, next, INT64_MIN};
    enumerate(&ctx, 0, 0, COL_NONE, 0);
    return ctx.best;
}

/* convention: 0 = stringzilla (first = open), 1 = bio (first = open + extend). local: 0 = global score, 1 = local. */
ORC_API int64_t orc_align_bruteforce(const uint8_t *a, size_t la, const uint8_t *b, size_t lb, const uint8_t *byte_to_class,
                                     const int8_t *costs, int open, int extend, int convention, int local) {
    const int first = convention == 1 ? open + extend : open;
    if (!local) return brute_global(a, la, b, lb, byte_to_class, costs, first, extend);
    int64_t best = 0; /* the empty local alignment */
    for (size_t i0 = 0; i0 < la; ++i0)
        for (size_t i1 = i0 + 1; i1 <= la; ++i1)
            for (size_t j0 = 0; j0 < lb; ++j0)
                for (size_t j1 = j0 + 1; j1 <= lb; ++j1) {
                    const int64_t s = brute_global(a + i0, i1 - i0, b + j0, j1 - j0, byte_to_class, costs, first, extend);
                    if (s > best) best = s;
                }
    return best;
}
