/*
 * hw2_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the reference pairwise aligner
 * (/root/reference/Local_Global_Alignment/hw2.cpp) used only as the parity
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * Nothing under bioinformatics-algorithms_b200/ links, imports or executes it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   - the reference's shipped golden files (global.txt / local.txt),
 *   - SURVEY.md Appendix B known-answer vectors,
 *   - tests/golden/*.json, produced by the UNMODIFIED reference binary
 *     (oracle/_ref/hw2, built by oracle/Makefile from the reference sources
 *     where they lie) via tests/golden/make_golden.py,
 *   - and, when oracle/_ref/hw2 is present, thousands of live 1-pair runs.
 *
 * Unlike the reference (which only prints the batch winner) every function
 * here reports per-pair results so the CUDA path can be checked pair by pair.
 *
 * Conventions shared with include/b2align.h:
 *   ops are produced in TRACEBACK order (alignment end -> start), one byte
 *   each: 'M' diagonal, 'D' pattern base over '-', 'I' '-' over text base
 *   (hw2.cpp:164-180, :240-256 -- the reference's own letters).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef struct {
    int32_t  score;     /* hw2.cpp:186 (global), :225-229 (local)            */
    uint32_t end_i;     /* cell where the traceback starts (1-based rows)    */
    uint32_t end_j;
    uint32_t start_i;   /* cell where the traceback stops                    */
    uint32_t start_j;
    int32_t  overlap;   /* overlapLongestExactMatch, hw2.cpp:267-278         */
    uint32_t n_ops;
} orc_result;

/* Needleman-Wunsch fill + traceback: hw2.cpp:118-190. Tie order d > l > u. */
static int orc_global(const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n,
                      int match, int mismatch, int gap, orc_result* res, uint8_t* ops)
{
    size_t W = (size_t)n + 1;
    int32_t* dp = (int32_t*)malloc(sizeof(int32_t) * (m + 1) * W);
    char*    tb = (char*)malloc((size_t)(m + 1) * W);
    if (!dp || !tb) { free(dp); free(tb); return -1; }
    memset(tb, ' ', (size_t)(m + 1) * W);
    /* hw2.cpp:125-136 (size_t * int wraps back to the right int, :126) */
    for (uint32_t i = 0; i <= m; ++i) { dp[i * W] = (int32_t)((int64_t)i * gap); if (i) tb[i * W] = 'u'; }
    for (uint32_t j = 0; j <= n; ++j) { dp[j] = (int32_t)((int64_t)j * gap);     if (j) tb[j] = 'l'; }
    /* hw2.cpp:138-156 */
    for (uint32_t i = 1; i <= m; ++i) {
        for (uint32_t j = 1; j <= n; ++j) {
            int up   = dp[(i - 1) * W + j] + gap;
            int left = dp[i * W + j - 1] + gap;
            int v    = dp[(i - 1) * W + j - 1] + (p[i - 1] == t[j - 1] ? match : mismatch);
            char d = 'd';
            if (left > v) { v = left; d = 'l'; }
            if (up > v)   { v = up;   d = 'u'; }
            dp[i * W + j] = v; tb[i * W + j] = d;
        }
    }
    /* hw2.cpp:158-181 */
    uint32_t ti = m, tj = n, k = 0;
    int cur = 0, best = 0;
    while (ti > 0 || tj > 0) {
        char d = tb[ti * W + tj];
        if (ti > 0 && tj > 0 && d == 'd') {
            ops[k++] = 'M';
            /* overlapLongestExactMatch hw2.cpp:267-278 (direction-agnostic) */
            if (p[ti - 1] == t[tj - 1] && p[ti - 1] != '-') { if (++cur > best) best = cur; } else cur = 0;
            --ti; --tj;
        } else if (ti > 0 && d == 'u') { ops[k++] = 'D'; cur = 0; --ti; }
        else if (tj > 0 && d == 'l')   { ops[k++] = 'I'; cur = 0; --tj; }
        else { free(dp); free(tb); return -2; } /* reference would spin forever */
    }
    res->score = dp[(size_t)m * W + n];          /* hw2.cpp:186 */
    res->end_i = m; res->end_j = n; res->start_i = 0; res->start_j = 0;
    res->overlap = best; res->n_ops = k;
    free(dp); free(tb);
    return 0;
}

/* Smith-Waterman fill + arg-max + traceback: hw2.cpp:192-265. Tie order 0 > d > u > l. */
static int orc_local(const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n,
                     int match, int mismatch, int gap, orc_result* res, uint8_t* ops)
{
    size_t W = (size_t)n + 1;
    int32_t* dp = (int32_t*)calloc((size_t)(m + 1) * W, sizeof(int32_t));
    char*    tb = (char*)malloc((size_t)(m + 1) * W);
    if (!dp || !tb) { free(dp); free(tb); return -1; }
    memset(tb, ' ', (size_t)(m + 1) * W);
    int score = 0; uint32_t bi = 0, bj = 0;       /* hw2.cpp:202-203 */
    for (uint32_t i = 1; i <= m; ++i) {
        for (uint32_t j = 1; j <= n; ++j) {
            int diag = dp[(i - 1) * W + j - 1] + (p[i - 1] == t[j - 1] ? match : mismatch);
            int up   = dp[(i - 1) * W + j] + gap;
            int left = dp[i * W + j - 1] + gap;
            int ul = up > left ? up : left;
            int v  = diag > ul ? diag : ul;
            if (v < 0) v = 0;                      /* hw2.cpp:211 */
            dp[i * W + j] = v;
            char d;                                /* hw2.cpp:214-222 */
            if (v == 0) d = '0'; else if (v == diag) d = 'd'; else if (v == up) d = 'u'; else d = 'l';
            tb[i * W + j] = d;
            if (v > score) { score = v; bi = i; bj = j; } /* hw2.cpp:225-229 */
        }
    }
    uint32_t ti = bi, tj = bj, k = 0;
    int cur = 0, best = 0;
    /* hw2.cpp:239-257 */
    while (ti > 0 && tj > 0 && dp[ti * W + tj] != 0) {
        char d = tb[ti * W + tj];
        if (d == 'd') {
            ops[k++] = 'M';
            if (p[ti - 1] == t[tj - 1] && p[ti - 1] != '-') { if (++cur > best) best = cur; } else cur = 0;
            --ti; --tj;
        } else if (d == 'u') { ops[k++] = 'D'; cur = 0; --ti; }
        else if (d == 'l')   { ops[k++] = 'I'; cur = 0; --tj; }
        else { free(dp); free(tb); return -2; }
    }
    res->score = score; res->end_i = bi; res->end_j = bj; res->start_i = ti; res->start_j = tj;
    res->overlap = best; res->n_ops = k;
    free(dp); free(tb);
    return 0;
}

/* mode 0 = global (-g), 1 = local (-l). ops must hold m+n bytes. */
int orc_align(int mode, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n,
              int match, int mismatch, int gap, orc_result* res, uint8_t* ops)
{
    return mode == 0 ? orc_global(p, m, t, n, match, mismatch, gap, res, ops)
                     : orc_local(p, m, t, n, match, mismatch, gap, res, ops);
}

/* Row-checkpointed restatement of orc_global / orc_local for pairs whose (m+1) x (n+1) matrices do not fit in memory
 * (config 4: the reference itself needs 49 GB for 100 kb x 100 kb).  Pass 1 is the fill loop hw2.cpp:138-156 / :205-231 on two
 * rows, keeping every ck-th row (and, local, the first row-major maximum, hw2.cpp:225-229).  Pass 2 re-fills one block of <= ck
 * rows from the checkpoint above it, bottom-up and only up to the path's current column, re-derives the direction letter of every
 * VISITED cell with the reference's own comparisons on the re-filled values (hw2.cpp:145-153: d, then l if left > diag, then u if
 * up > max; hw2.cpp:214-222: 0, d, u, l by equality) and walks exactly as hw2.cpp:158-181 / :239-257 do.  Same outputs as orc_align
 * (tests/test_oracle.py checks that on random, tandem-repeat and odd-scoring cases with tiny ck). */
int orc_align_ckpt(int mode, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n,
                   int match, int mismatch, int gap, uint32_t ck, orc_result* res, uint8_t* ops)
{
    if (ck == 0) ck = 1024;
    const size_t W = (size_t)n + 1;
    const uint32_t nck = m / ck + 1;                       /* checkpoint b holds DP row b*ck */
    int32_t* cp = (int32_t*)malloc(sizeof(int32_t) * W * nck);
    int32_t* row = (int32_t*)malloc(sizeof(int32_t) * W);
    int32_t* blk = (int32_t*)malloc(sizeof(int32_t) * W * ((size_t)ck + 1));
    if (!cp || !row || !blk) { free(cp); free(row); free(blk); return -1; }
    int score = 0; uint32_t bi = 0, bj = 0;
    for (uint32_t j = 0; j <= n; ++j) row[j] = mode == 0 ? (int32_t)((int64_t)j * gap) : 0;
    memcpy(cp, row, sizeof(int32_t) * W);
    int32_t* tmp = blk;                                     /* scratch for pass 1 (blk is not in use yet) */
    for (uint32_t i = 1; i <= m; ++i) {
        /* as in orc_score_only: max(diag + s, up + gap) from the previous row (vectorisable), then the left neighbour along the row */
        const uint8_t pc = p[i - 1];
        for (uint32_t j = 1; j <= n; ++j) {
            const int d = row[j - 1] + (pc == t[j - 1] ? match : mismatch);
            const int u = row[j] + gap;
            tmp[j] = d > u ? d : u;
        }
        int32_t left = mode == 0 ? (int32_t)((int64_t)i * gap) : 0;
        row[0] = left;
        for (uint32_t j = 1; j <= n; ++j) {
            int v = tmp[j]; const int l = left + gap; if (l > v) v = l;
            if (mode != 0) { if (v < 0) v = 0; if (v > score) { score = v; bi = i; bj = j; } }
            row[j] = v; left = v;
        }
        if (i % ck == 0) memcpy(cp + (size_t)(i / ck) * W, row, sizeof(int32_t) * W);
    }
    uint32_t ti = mode == 0 ? m : bi, tj = mode == 0 ? n : bj, k = 0;
    int cur = 0, best = 0, stop = 0;
    res->score = mode == 0 ? row[n] : score;
    res->end_i = ti; res->end_j = tj;
    while (!stop && ti > 0 && tj > 0) {
        const uint32_t b = (ti - 1) / ck, r0 = b * ck;     /* block rows r0+1 .. ti, columns 0 .. tj */
        const size_t BW = (size_t)tj + 1;
        memcpy(blk, cp + (size_t)b * W, sizeof(int32_t) * BW);
        for (uint32_t i = r0 + 1; i <= ti; ++i) {
            int32_t* prev = blk + (size_t)(i - 1 - r0) * BW; int32_t* curr = blk + (size_t)(i - r0) * BW;
            const uint8_t pc = p[i - 1];
            for (uint32_t j = 1; j <= tj; ++j) {
                const int d = prev[j - 1] + (pc == t[j - 1] ? match : mismatch);
                const int u = prev[j] + gap;
                curr[j] = d > u ? d : u;
            }
            int32_t left = mode == 0 ? (int32_t)((int64_t)i * gap) : 0;
            curr[0] = left;
            for (uint32_t j = 1; j <= tj; ++j) {
                int v = curr[j]; const int l = left + gap; if (l > v) v = l; if (mode != 0 && v < 0) v = 0;
                curr[j] = v; left = v;
            }
        }
        while (ti > r0 && tj > 0) {
            const int32_t* prev = blk + (size_t)(ti - 1 - r0) * BW; const int32_t* curr = blk + (size_t)(ti - r0) * BW;
            const int diag = prev[tj - 1] + (p[ti - 1] == t[tj - 1] ? match : mismatch);
            const int up = prev[tj] + gap, left = curr[tj - 1] + gap, v = curr[tj];
            char d;
            if (mode == 0) { int x = diag; d = 'd'; if (left > x) { x = left; d = 'l'; } if (up > x) d = 'u'; }      /* hw2.cpp:145-153 */
            else { if (v == 0) { stop = 1; break; } d = v == diag ? 'd' : (v == up ? 'u' : 'l'); }                        /* hw2.cpp:239, :214-222 */
            if (d == 'd') {
                ops[k++] = 'M';
                if (p[ti - 1] == t[tj - 1] && p[ti - 1] != '-') { if (++cur > best) best = cur; } else cur = 0;
                --ti; --tj;
            } else if (d == 'u') { ops[k++] = 'D'; cur = 0; --ti; }
            else { ops[k++] = 'I'; cur = 0; --tj; }
        }
    }
    if (mode == 0) {                                        /* borders: column 0 holds 'u', row 0 holds 'l' (hw2.cpp:128, :134) */
        while (ti > 0) { ops[k++] = 'D'; --ti; }
        while (tj > 0) { ops[k++] = 'I'; --tj; }
    }
    res->start_i = ti; res->start_j = tj; res->overlap = best; res->n_ops = k;
    free(cp); free(row); free(blk);
    return 0;
}

/* Linear-memory score(+end cell) only, for sizes where the full matrices do
 * not fit: same recurrences (hw2.cpp:138-156 / :205-231), two rows. */
int orc_score_only(int mode, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n,
                   int match, int mismatch, int gap, orc_result* res)
{
    /* Per row the cell max(diag + s, up + gap) depends on the previous row only (first loop: no loop-carried dependency, the compiler
     * vectorises it); the left neighbour is folded in by a second, sequential loop.  max is associative, so every H equals the
     * reference's; the local arg-max scans the finished row left to right with the same strict '>' (hw2.cpp:225-229). */
    int32_t* row = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1) * 2);
    if (!row) return -1;
    int32_t* tmp = row + (size_t)n + 1;
    int score = 0; uint32_t bi = 0, bj = 0;
    for (uint32_t j = 0; j <= n; ++j) row[j] = mode == 0 ? (int32_t)((int64_t)j * gap) : 0;
    for (uint32_t i = 1; i <= m; ++i) {
        const uint8_t pc = p[i - 1];
        for (uint32_t j = 1; j <= n; ++j) {
            const int d = row[j - 1] + (pc == t[j - 1] ? match : mismatch);
            const int u = row[j] + gap;
            tmp[j] = d > u ? d : u;
        }
        int32_t left = mode == 0 ? (int32_t)((int64_t)i * gap) : 0;
        row[0] = left;
        if (mode == 0) {
            for (uint32_t j = 1; j <= n; ++j) { int v = tmp[j]; const int l = left + gap; if (l > v) v = l; row[j] = v; left = v; }
        } else {
            for (uint32_t j = 1; j <= n; ++j) {
                int v = tmp[j]; const int l = left + gap; if (l > v) v = l; if (v < 0) v = 0;
                if (v > score) { score = v; bi = i; bj = j; }
                row[j] = v; left = v;
            }
        }
    }
    memset(res, 0, sizeof(*res));
    if (mode == 0) { res->score = row[n]; res->end_i = m; res->end_j = n; }
    else { res->score = score; res->end_i = bi; res->end_j = bj; }
    free(row);
    return 0;
}

/* prepareCigarString, hw2.cpp:59-78: RLE of the op list read back to front. */
size_t orc_cigar(const uint8_t* ops, uint32_t n_ops, char* out, size_t cap)
{
    size_t w = 0;
    if (cap) out[0] = 0;
    if (n_ops == 0) return 0;
    uint32_t count = 1; uint8_t cur = ops[n_ops - 1];
    for (int64_t i = (int64_t)n_ops - 2; i >= -1; --i) {
        if (i >= 0 && ops[i] == cur) { ++count; continue; }
        w += (size_t)snprintf(out + (w < cap ? w : cap), w < cap ? cap - w : 0, "%u%c", count, cur);
        if (i >= 0) { cur = ops[i]; count = 1; }
    }
    return w;
}

/* prepareMDZString, hw2.cpp:80-116, restated on (ops, raw sequences, start cell)
 * instead of the two aligned strings: column c of the alignment holds pattern
 * base p[pi] for M/D and text base t[tj] for M/I. */
size_t orc_mdz(const uint8_t* ops, uint32_t n_ops, const uint8_t* p, const uint8_t* t,
               uint32_t start_i, uint32_t start_j, char* out, size_t cap)
{
    size_t w = 0; int run = 0;
    uint32_t pi = start_i, tj = start_j;
    int64_t c = (int64_t)n_ops - 1;                 /* forward order = reversed list (hw2.cpp:81) */
#define EMIT(...) do { w += (size_t)snprintf(out + (w < cap ? w : cap), w < cap ? cap - w : 0, __VA_ARGS__); } while (0)
    while (c >= 0) {
        if (ops[c] == 'M') {
            if (p[pi] == t[tj]) ++run;              /* hw2.cpp:89-90 */
            else { EMIT("%d%c", run, t[tj]); run = 0; } /* hw2.cpp:93-95: reference base */
            ++pi; ++tj; --c;
        } else if (ops[c] == 'D') {                 /* hw2.cpp:98-108 */
            EMIT("%d^", run); run = 0;
            while (c >= 0 && ops[c] == 'D') { EMIT("%c", p[pi]); ++pi; --c; }
        } else { ++tj; --c; }                       /* hw2.cpp:109-112: I does not break the run */
    }
    EMIT("%d", run);                                /* hw2.cpp:114 */
#undef EMIT
    return w;
}

/* hw3.cpp:23-98 score path (3-state affine global, no E<->F transitions), two
 * rows per state instead of six full matrices. */
int orc_affine_score(const uint8_t* s1, uint32_t m, const uint8_t* s2, uint32_t n,
                     int match, int mismatch, int gopen, int gext, int32_t* score)
{
    const int32_t NEG = INT32_MIN / 2;             /* hw3.cpp:16 */
    size_t W = (size_t)n + 1;
    int32_t* V = (int32_t*)malloc(sizeof(int32_t) * W * 6);
    if (!V) return -1;
    int32_t *Vp = V, *Fp = V + W, *Ep = V + 2 * W, *Vc = V + 3 * W, *Fc = V + 4 * W, *Ec = V + 5 * W;
    Vp[0] = 0; Fp[0] = NEG; Ep[0] = NEG;            /* hw3.cpp:40-41 */
    for (uint32_t j = 1; j <= n; ++j) { Vp[j] = NEG; Fp[j] = NEG; Ep[j] = gopen + gext * (int32_t)(j - 1); } /* :48-53 */
    /* V and F of a row depend on the previous row only (first loop, vectorisable); E runs along the row (second loop).  Same values. */
    for (uint32_t i = 1; i <= m; ++i) {
        Vc[0] = NEG; Ec[0] = NEG; Fc[0] = gopen + gext * (int32_t)(i - 1);                                   /* :42-47 */
        const uint8_t c1 = s1[i - 1];
        {
            const int32_t* restrict vp = Vp; const int32_t* restrict fp = Fp; const int32_t* restrict ep = Ep;   /* six disjoint rows */
            int32_t* restrict vc = Vc; int32_t* restrict fc = Fc;
            for (uint32_t j = 1; j <= n; ++j) {
                const int s = c1 == s2[j - 1] ? match : mismatch;
                int v = vp[j - 1];                                         /* :59-68: max(V, F, E)[i-1][j-1] + s */
                if (fp[j - 1] > v) v = fp[j - 1];
                if (ep[j - 1] > v) v = ep[j - 1];
                int f = vp[j] + gopen + gext;                              /* :70-75 */
                if (fp[j] + gext > f) f = fp[j] + gext;
                vc[j] = v + s; fc[j] = f;
            }
        }
        int32_t e = Ec[0];
        for (uint32_t j = 1; j <= n; ++j) {
            int open = Vc[j - 1] + gopen + gext;                           /* :77-82 */
            e = e + gext > open ? e + gext : open;
            Ec[j] = e;
        }
        int32_t* x;
        x = Vp; Vp = Vc; Vc = x; x = Fp; Fp = Fc; Fc = x; x = Ep; Ep = Ec; Ec = x;
    }
    int best = Vp[n];                                                      /* :86-98 */
    if (Fp[n] > best) best = Fp[n];
    if (Ep[n] > best) best = Ep[n];
    *score = best;
    free(V);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * hw4's own Needleman-Wunsch (/root/reference/hw4/hw4.cpp:16-72) and its distance (hw4.cpp:141-152).
 * Same recurrence as hw2's, but the tie order is d > u > l (hw4.cpp:37-46: 'U' overrides the diagonal if
 * strictly larger, then 'L' overrides if strictly larger) -- NOT hw2's d > l > u (SURVEY.md trap T3).
 * distance = number of alignment columns holding a gap or a mismatch.  ops in traceback order, hw2's
 * letters: 'M' diagonal, 'D' sequence1 base over '-', 'I' '-' over sequence2 base.
 * ------------------------------------------------------------------------------------------------ */
int orc_hw4_nw(const uint8_t* s1, uint32_t m, const uint8_t* s2, uint32_t n, int match, int mismatch, int gap,
               int32_t* score, int32_t* distance, uint32_t* n_ops, uint8_t* ops)
{
    size_t W = (size_t)n + 1;
    int32_t* dp = (int32_t*)malloc(sizeof(int32_t) * (m + 1) * W);
    char*    tb = (char*)malloc((size_t)(m + 1) * W);
    if (!dp || !tb) { free(dp); free(tb); return -1; }
    dp[0] = 0; tb[0] = 0;
    for (uint32_t i = 1; i <= m; ++i) { dp[i * W] = dp[(i - 1) * W] + gap; tb[i * W] = 'U'; }   /* hw4.cpp:21-24 */
    for (uint32_t j = 1; j <= n; ++j) { dp[j] = dp[j - 1] + gap; tb[j] = 'L'; }                 /* hw4.cpp:25-28 */
    for (uint32_t i = 1; i <= m; ++i)
        for (uint32_t j = 1; j <= n; ++j) {                                                     /* hw4.cpp:30-48 */
            int up = dp[(i - 1) * W + j] + gap, left = dp[i * W + j - 1] + gap;
            int v = dp[(i - 1) * W + j - 1] + (s1[i - 1] == s2[j - 1] ? match : mismatch);
            char d = 'D';
            if (up > v)   { v = up;   d = 'U'; }
            if (left > v) { v = left; d = 'L'; }
            dp[i * W + j] = v; tb[i * W + j] = d;
        }
    uint32_t i = m, j = n, k = 0; int32_t dist = 0;
    while (i > 0 || j > 0) {                                                                    /* hw4.cpp:52-68 */
        if (i > 0 && j > 0 && tb[i * W + j] == 'D') { if (s1[i - 1] != s2[j - 1]) ++dist; if (ops) ops[k] = 'M'; ++k; --i; --j; }
        else if (i > 0 && tb[i * W + j] == 'U')     { ++dist; if (ops) ops[k] = 'D'; ++k; --i; }
        else                                         { ++dist; if (ops) ops[k] = 'I'; ++k; --j; }
    }
    if (score) *score = dp[m * W + n];
    if (distance) *distance = dist;
    if (n_ops) *n_ops = k;
    free(dp); free(tb);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * hw3's affine_alignment WITH its traceback (/root/reference/Multiple_Sequence_Alignment/hw3.cpp:23-135).
 * Full V/F/E + trace matrices, so for test sizes only.  ops in traceback order, hw2's letters:
 * 'M' column of two bases (state V), 'D' string1 base over '-' (state F), 'I' '-' over string2 base (state E).
 * ------------------------------------------------------------------------------------------------ */
int orc_affine_align(const uint8_t* s1, uint32_t m, const uint8_t* s2, uint32_t n, int match, int mismatch,
                     int gopen, int gext, int32_t* score, uint32_t* n_ops, uint8_t* ops)
{
    const int32_t NEG = INT32_MIN / 2;
    size_t W = (size_t)n + 1, cells = (size_t)(m + 1) * W;
    int32_t* V = (int32_t*)malloc(sizeof(int32_t) * cells * 3);
    int8_t*  T = (int8_t*)malloc(cells * 3);
    if (!V || !T) { free(V); free(T); return -1; }
    int32_t *F = V + cells, *E = V + 2 * cells;
    int8_t *tV = T, *tF = T + cells, *tE = T + 2 * cells;
    for (size_t k = 0; k < cells; ++k) { V[k] = F[k] = E[k] = NEG; tV[k] = tF[k] = tE[k] = -1; }      /* hw3.cpp:28-37 */
    V[0] = 0;                                                                                          /* hw3.cpp:40 */
    for (uint32_t i = 1; i <= m; ++i) { F[i * W] = gopen + gext * (int32_t)(i - 1); tF[i * W] = i == 1 ? 0 : 1; }   /* :42-47 */
    for (uint32_t j = 1; j <= n; ++j) { E[j] = gopen + gext * (int32_t)(j - 1); tE[j] = j == 1 ? 0 : 1; }           /* :48-53 */
    for (uint32_t i = 1; i <= m; ++i)
        for (uint32_t j = 1; j <= n; ++j) {                                                            /* hw3.cpp:55-84 */
            size_t c = i * W + j, d = (i - 1) * W + j - 1, u = (i - 1) * W + j, l = i * W + j - 1;
            int s = s1[i - 1] == s2[j - 1] ? match : mismatch;
            V[c] = V[d] + s; tV[c] = 0;
            if (F[d] + s > V[c]) { V[c] = F[d] + s; tV[c] = 1; }
            if (E[d] + s > V[c]) { V[c] = E[d] + s; tV[c] = 2; }
            F[c] = V[u] + gopen + gext; tF[c] = 0;
            if (F[u] + gext > F[c]) { F[c] = F[u] + gext; tF[c] = 1; }
            E[c] = V[l] + gopen + gext; tE[c] = 0;
            if (E[l] + gext > E[c]) { E[c] = E[l] + gext; tE[c] = 1; }
        }
    size_t end = (size_t)m * W + n;
    int state = 0; int32_t best = V[end];                                                              /* hw3.cpp:86-98 */
    if (F[end] > best) { best = F[end]; state = 1; }
    if (E[end] > best) { best = E[end]; state = 2; }
    if (score) *score = best;
    uint32_t i = m, j = n, k = 0;
    while (i > 0 || j > 0) {                                                                           /* hw3.cpp:105-131 */
        size_t c = (size_t)i * W + j;
        if (state == 0) { int prev = tV[c]; if (ops) ops[k] = 'M'; ++k; --i; --j; state = prev; }
        else if (state == 1) { state = tF[c] == 0 ? 0 : 1; if (ops) ops[k] = 'D'; ++k; --i; }
        else { state = tE[c] == 0 ? 0 : 2; if (ops) ops[k] = 'I'; ++k; --j; }
        if (k > m + n) break;                                                                          /* (the reference would run off the matrix) */
    }
    if (n_ops) *n_ops = k;
    free(V); free(T);
    return 0;
}
