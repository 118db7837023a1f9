/*
 * oracle/dp_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of SOAP3-dp's semi-global affine-gap DP (scheme 1, full
 * table): score pass + traceback for a batch in the reference's packed batch
 * format.  Follows
 *   recurrence, soft-clip restart, anchors, best-cell rule   DV-DPfunctions.cu:146-241
 *   per-alignment parameter defaults (NULL arrays)            DV-DPfunctions.cu:254-266
 *   traceback state machine and pattern bytes                 DV-DPfunctions.cu:316-512
 *   packed layouts (1-based, MSB first, 32-interleaved)       DV-DPfunctions.cu:55-59
 *
 * Unlike the reference, which keeps the full H and E tables (8 B per cell) and
 * re-derives every traceback decision from table lookups, this restatement
 * records ONE byte per cell during the score pass -- the outcome of each
 * comparison the reference's traceback would make at that cell -- and the
 * traceback only reads those bytes.  That is the formulation the CUDA kernel
 * uses; tests/test_oracle_vs_ref.py pins it against the reference's own kernels
 * compiled for the host (oracle/_ref/libref_dp.so).
 *
 * Domain of bit-exactness: cutoff thresholds large enough that no cell on a
 * reported path is saturated at -32000 (any cutoff >= 0, the only values the
 * reference's callers produce: CPUfunctions.cpp:65-66).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-32000)

/* traceback byte */
#define TB_DIAG   0
#define TB_DOPEN  1
#define TB_DEXT   2
#define TB_SMEXIT 3
#define TB_SIEXIT 4
#define TB_IOPEN  5
#define TB_IEXT   6
#define TB_CHOICE_MASK 7
#define TB_MATCH  8      /* refChar == readChar                              */
#define TB_EOPEN  16     /* E[j][i] == open + H[j-1][i]                      */
#define TB_FOPEN  32     /* F carried out of (j,i) == open + H[j][i-1]       */
#define TB_FEXIT  64     /* i <= clipLt+1 and F carried out == init_j + open */

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int clampS(int x) { return x > NEG_INF ? x : NEG_INF; }

static inline uint32_t unpack1(const uint32_t *seq, uint32_t lane_base, uint32_t i)
{
    /* base i >= 1 in bits 2*(15-(i&15)) of word i>>4, words 32-interleaved */
    return (seq[lane_base + ((i >> 4) << 5)] >> ((15 - (i & 15)) << 1)) & 3;
}

typedef struct {
    int match, mismatch, open, ext;
} dp_scores;

typedef struct { int best; uint32_t hit, scRight, count; } dp_best;

/* column 0 (DV-DPfunctions.cu:167-184) */
static void dp_column0(uint32_t m, dp_scores sc, uint32_t clipLt, int *Hprev, int *Eprev)
{
    const int gapInit = sc.open - sc.ext;
    Hprev[0] = clampS(0); Eprev[0] = clampS(gapInit);
    int up = gapInit;
    for (uint32_t i = 1; i <= m; ++i) {
        if (i <= clipLt) { Hprev[i] = clampS(sc.open); Eprev[i] = clampS(sc.open + gapInit); }
        else { up += sc.ext; Hprev[i] = clampS(up); Eprev[i] = clampS(up + gapInit); }
    }
}

/* column j >= 1 from column j - 1 (Hprev / Eprev in, out); prevInit = the score pass's prevInitScore (0 before column 1,
 * :186); tbc = this column's traceback bytes (m + 1) or NULL; the best-cell rule of :225-235 */
static void dp_column(const uint32_t *dna, uint32_t dna_base, const uint32_t *read, uint32_t read_base, uint32_t m, uint32_t j,
                      dp_scores sc, uint32_t clipLt, uint32_t clipRt, uint32_t anchorLeft, uint32_t anchorRight,
                      int prevInit, int *Hprev, int *Eprev, uint8_t *tbc, dp_best *bst)
{
    const int gapInit = sc.open - sc.ext;
    const int clipRtCheck = (int)(m - clipRt);
    const uint32_t refChar = unpack1(dna, dna_base, j);
    const int init = (j >= anchorLeft) ? NEG_INF : 0;
    /* GPUBacktrack recomputes "the previous column's init" as (j > anchorLeft) (:352-357,:368),
     * which differs from the score pass's prevInitScore (0 before column 1, :186) when
     * anchorLeft == 0 and j == 1; the traceback test must use the traceback's value */
    const int prevInitTb = (j > anchorLeft) ? NEG_INF : 0;
    int upScore = init, F = init + gapInit;
    int diag = prevInit;            /* with soft-clip restart applied */
    int diagRaw = Hprev[0];         /* stored H[j-1][i-1]             */
    int upStored = clampS(upScore); /* stored H[j][i-1]               */
    Hprev[0] = clampS(upScore);
    Eprev[0] = clampS(F);
    for (uint32_t i = 1; i <= m; ++i) {
        const uint32_t readChar = unpack1(read, read_base, i);
        const int d = (refChar == readChar) ? sc.match : sc.mismatch;
        const int left = Hprev[i], eLeft = Eprev[i];
        const int e = imax(sc.open + left, sc.ext + eLeft);
        const int eStored = clampS(e);
        F = imax(sc.ext + F, sc.open + upScore);
        upScore = imax(imax(F, e), diag + d);
        const int hStored = clampS(upScore);
        /* what GPUBacktrack would decide in its NORMAL state here (:379-440) */
        uint8_t b;
        if (hStored == d + diagRaw) b = TB_DIAG;
        else if (hStored == sc.open + left) b = TB_DOPEN;
        else if (hStored == sc.ext + eLeft) b = TB_DEXT;
        else if (i <= clipLt + 1 && hStored == prevInitTb + d) b = TB_SMEXIT;
        else if (i <= clipLt + 1 && hStored == init + sc.open) b = TB_SIEXIT;
        else if (hStored == sc.open + upStored) b = TB_IOPEN;
        else b = TB_IEXT;
        if (refChar == readChar) b |= TB_MATCH;
        if (eStored == sc.open + left) b |= TB_EOPEN;
        /* carried state for the next row (:215-219) */
        diag = left; diagRaw = left;
        if (i <= clipLt) { F = imax(init + gapInit, F); diag = imax(prevInit, diag); }
        if (F == sc.open + upStored) b |= TB_FOPEN;
        if (i <= clipLt + 1 && F == init + sc.open) b |= TB_FEXIT;
        if (tbc) tbc[i] = b;
        Hprev[i] = hStored; Eprev[i] = eStored;
        upStored = hStored;
        if (bst && (int)i >= clipRtCheck && j >= anchorRight) {
            if (upScore > bst->best) { bst->best = upScore; bst->hit = j; bst->scRight = m - i; bst->count = 1; }
            else if (upScore == bst->best) ++bst->count;
        }
    }
}

/* traceback (DV-DPfunctions.cu:330-508) driven by the recorded bytes of columns firstCol + 1 .. hit (tb = the bytes of
 * column firstCol + 1 on).  Returns 0 when the path wants a column at or before firstCol > 0 (window too short). */
static int dp_traceback(const uint8_t *tb, uint32_t firstCol, uint32_t m, uint32_t clipLt, dp_best bst, uint8_t *pattern, uint32_t *hitOut)
{
    const uint32_t scRight = bst.scRight;
    uint32_t p = 0;
    if (scRight > 0) { pattern[p++] = 'S'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)scRight; }
    uint32_t readPos = m - scRight, refIndex = bst.hit;
    enum { NORMAL, I_EXT, D_EXT, SM_EXIT, SI_EXIT } state = NORMAL;
    uint8_t lastCell = 0;
    while (readPos > 0 && refIndex > firstCol) {
        const uint8_t b = tb[(size_t)(refIndex - firstCol - 1) * (m + 1) + readPos];
        lastCell = b;
        if (state == NORMAL) {
            const int ch = b & TB_CHOICE_MASK;
            if (ch == TB_DIAG) { pattern[p++] = (b & TB_MATCH) ? 'M' : 'm'; --refIndex; --readPos; }
            else if (ch == TB_DOPEN) { pattern[p++] = 'D'; --refIndex; }
            else if (ch == TB_DEXT) { pattern[p++] = 'D'; --refIndex; state = D_EXT; }
            else if (ch == TB_SMEXIT) { state = SM_EXIT; break; }
            else if (ch == TB_SIEXIT) { state = SI_EXIT; break; }
            else if (ch == TB_IOPEN) { pattern[p++] = 'I'; --readPos; }
            else { pattern[p++] = 'I'; --readPos; state = I_EXT; }
        } else if (state == D_EXT) {
            pattern[p++] = 'D'; --refIndex;
            if (b & TB_EOPEN) state = NORMAL;
        } else { /* I_EXT */
            if (b & TB_FEXIT) { state = SI_EXIT; break; }
            pattern[p++] = 'I'; --readPos;
            if (b & TB_FOPEN) state = NORMAL;
        }
    }
    if (firstCol > 0 && refIndex == firstCol && readPos > 0 && state != SM_EXIT && state != SI_EXIT) return 0;
    if (refIndex == 0) {
        const uint32_t scNum = clipLt < readPos ? clipLt : readPos;
        if (scNum < readPos) { pattern[p++] = 'I'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)(readPos - scNum); }
        pattern[p++] = 'S'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)scNum;
    } else if (state == SI_EXIT) {
        pattern[p++] = 'I'; pattern[p++] = 'S'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)(readPos - 1);
    } else if (state == SM_EXIT) {
        pattern[p++] = (lastCell & TB_MATCH) ? 'M' : 'm';
        pattern[p++] = 'S'; pattern[p++] = 'V'; pattern[p++] = (uint8_t)(readPos - 1);
        refIndex -= 1;
    }
    pattern[p++] = 0;
    *hitOut = refIndex;
    return 1;
}

/* one alignment; tb holds (n + 1) x (m + 1) bytes, column by column */
static void dp_one(const uint32_t *dna, uint32_t dna_base, uint32_t n,
                   const uint32_t *read, uint32_t read_base, uint32_t m,
                   dp_scores sc, uint32_t clipLt, uint32_t clipRt, uint32_t anchorLeft, uint32_t anchorRight,
                   int cutoff, int *scoreOut, uint32_t *hitOut, uint32_t *countOut, uint8_t *pattern,
                   uint8_t *tb, int *Hprev, int *Eprev)
{
    dp_best bst = {NEG_INF, 0, 0, 0};
    dp_column0(m, sc, clipLt, Hprev, Eprev);
    for (uint32_t j = 1; j <= n; ++j)
        dp_column(dna, dna_base, read, read_base, m, j, sc, clipLt, clipRt, anchorLeft, anchorRight,
                  j == 1 ? 0 : ((j - 1 >= anchorLeft) ? NEG_INF : 0), Hprev, Eprev, tb + (size_t)(j - 1) * (m + 1), &bst);
    *scoreOut = bst.best; *hitOut = bst.hit; *countOut = bst.count;
    if (bst.best < cutoff) return;
    dp_traceback(tb, 0, m, clipLt, bst, pattern, hitOut);
}

/* The same alignment WITHOUT the full table: a score pass that keeps (H, E) of every `every`-th column only, then a
 * second sweep from the last kept column at least m + slack columns before the best cell up to the best cell's column,
 * recording traceback bytes for those columns alone; if the path wants an earlier column the sweep restarts from an
 * earlier kept column (column 0 in the end).  Design check for a DP kernel that never writes the whole plane: results
 * must equal dp_one's.  resweptCols / restarts accumulate what that costs. */
static void dp_one_resweep(const uint32_t *dna, uint32_t dna_base, uint32_t n,
                           const uint32_t *read, uint32_t read_base, uint32_t m,
                           dp_scores sc, uint32_t clipLt, uint32_t clipRt, uint32_t anchorLeft, uint32_t anchorRight,
                           int cutoff, int *scoreOut, uint32_t *hitOut, uint32_t *countOut, uint8_t *pattern,
                           uint8_t *tb, int *Hprev, int *Eprev, int *ckpt, uint32_t every, uint32_t slack,
                           unsigned long long *resweptCols, unsigned long long *restarts)
{
    dp_best bst = {NEG_INF, 0, 0, 0};
    const size_t w = (size_t)m + 1;
    dp_column0(m, sc, clipLt, Hprev, Eprev);
    memcpy(ckpt, Hprev, w * sizeof(int)); memcpy(ckpt + w, Eprev, w * sizeof(int));
    for (uint32_t j = 1; j <= n; ++j) {
        dp_column(dna, dna_base, read, read_base, m, j, sc, clipLt, clipRt, anchorLeft, anchorRight,
                  j == 1 ? 0 : ((j - 1 >= anchorLeft) ? NEG_INF : 0), Hprev, Eprev, NULL, &bst);
        if (j % every == 0) { int *c = ckpt + (size_t)(j / every) * 2 * w; memcpy(c, Hprev, w * sizeof(int)); memcpy(c + w, Eprev, w * sizeof(int)); }
    }
    *scoreOut = bst.best; *hitOut = bst.hit; *countOut = bst.count;
    if (bst.best < cutoff) return;
    uint32_t want = bst.hit > m + slack ? bst.hit - (m + slack) : 0;
    for (;;) {
        const uint32_t c0 = want / every * every;
        const int *c = ckpt + (size_t)(c0 / every) * 2 * w;
        memcpy(Hprev, c, w * sizeof(int)); memcpy(Eprev, c + w, w * sizeof(int));
        for (uint32_t j = c0 + 1; j <= bst.hit; ++j)
            dp_column(dna, dna_base, read, read_base, m, j, sc, clipLt, clipRt, anchorLeft, anchorRight,
                      j == 1 ? 0 : ((j - 1 >= anchorLeft) ? NEG_INF : 0), Hprev, Eprev, tb + (size_t)(j - c0 - 1) * w, NULL);
        *resweptCols += bst.hit - c0;
        if (dp_traceback(tb, c0, m, clipLt, bst, pattern, hitOut)) return;
        ++*restarts;
        want = c0 > 4 * every ? c0 - 4 * every : 0;
    }
}

/* Batch entry: same arrays as SemiGlobalAligner::performAlignment (DV-DPfunctions.cu:669).
 * maxDPTableLength == maxDNALength (scheme 1); patternLength = maxReadLength + maxDNALength.
 * Returns the number of DP cells (sum readLength*DNALength). */
unsigned long long s3o_dp_align(const uint32_t *packedDNASequence, const uint32_t *DNALengths, uint32_t maxDNALength,
                                const uint32_t *packedReadSequence, const uint32_t *readLengths, uint32_t maxReadLength,
                                const int32_t *cutoffThresholds, int32_t *scores, uint32_t *hitLocs,
                                uint32_t *maxScoreCounts, uint8_t *pattern, uint32_t numOfThreads,
                                const uint32_t *clipLtSizes, const uint32_t *clipRtSizes,
                                const uint32_t *anchorLeftLocs, const uint32_t *anchorRightLocs,
                                int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore)
{
    const uint32_t dnaW = ((maxDNALength + 15) >> 4) << 5, readW = ((maxReadLength + 15) >> 4) << 5;
    const uint32_t patLen = maxReadLength + maxDNALength;
    dp_scores sc = {matchScore, mismatchScore, gapOpenScore, gapExtendScore};
    unsigned long long cells = 0;
    #pragma omp parallel reduction(+:cells)
    {
        uint8_t *tb = (uint8_t *)malloc((size_t)(maxDNALength + 2) * (maxReadLength + 2));
        int *H = (int *)malloc(sizeof(int) * (maxReadLength + 2)), *E = (int *)malloc(sizeof(int) * (maxReadLength + 2));
        #pragma omp for schedule(dynamic, 8)
        for (long long t = 0; t < (long long)numOfThreads; ++t) {
            const uint32_t g = (uint32_t)t >> 5, lane = (uint32_t)t & 31;
            dp_one(packedDNASequence, g * dnaW + lane, DNALengths[t], packedReadSequence, g * readW + lane, readLengths[t],
                   sc, clipLtSizes ? clipLtSizes[t] : 0, clipRtSizes ? clipRtSizes[t] : 0,
                   anchorLeftLocs ? anchorLeftLocs[t] : maxDNALength, anchorRightLocs ? anchorRightLocs[t] : 0,
                   cutoffThresholds[t], &scores[t], &hitLocs[t], &maxScoreCounts[t], pattern + (size_t)t * patLen, tb, H, E);
            cells += (unsigned long long)readLengths[t] * DNALengths[t];
        }
        free(tb); free(H); free(E);
    }
    return cells;
}


/* s3o_dp_align computed without the full table (dp_one_resweep): same arrays, same results; stats[0] += columns swept a
 * second time, stats[1] += restarts, stats[2] += columns of the alignments that were traced (what the full table holds). */
unsigned long long s3o_dp_align_resweep(const uint32_t *packedDNASequence, const uint32_t *DNALengths, uint32_t maxDNALength,
                                        const uint32_t *packedReadSequence, const uint32_t *readLengths, uint32_t maxReadLength,
                                        const int32_t *cutoffThresholds, int32_t *scores, uint32_t *hitLocs,
                                        uint32_t *maxScoreCounts, uint8_t *pattern, uint32_t numOfThreads,
                                        const uint32_t *clipLtSizes, const uint32_t *clipRtSizes,
                                        const uint32_t *anchorLeftLocs, const uint32_t *anchorRightLocs,
                                        int matchScore, int mismatchScore, int gapOpenScore, int gapExtendScore,
                                        uint32_t checkpointEvery, uint32_t slack, unsigned long long *stats)
{
    const uint32_t dnaW = ((maxDNALength + 15) >> 4) << 5, readW = ((maxReadLength + 15) >> 4) << 5;
    const uint32_t patLen = maxReadLength + maxDNALength;
    dp_scores sc = {matchScore, mismatchScore, gapOpenScore, gapExtendScore};
    unsigned long long cells = 0, reswept = 0, restarts = 0, tracedCols = 0;
    uint8_t *tb = (uint8_t *)malloc((size_t)(maxDNALength + 2) * (maxReadLength + 2));
    int *H = (int *)malloc(sizeof(int) * (maxReadLength + 2)), *E = (int *)malloc(sizeof(int) * (maxReadLength + 2));
    int *ckpt = (int *)malloc(sizeof(int) * 2 * (maxReadLength + 2) * (maxDNALength / checkpointEvery + 2));
    for (uint32_t t = 0; t < numOfThreads; ++t) {
        const uint32_t g = t >> 5, lane = t & 31;
        dp_one_resweep(packedDNASequence, g * dnaW + lane, DNALengths[t], packedReadSequence, g * readW + lane, readLengths[t],
                       sc, clipLtSizes ? clipLtSizes[t] : 0, clipRtSizes ? clipRtSizes[t] : 0,
                       anchorLeftLocs ? anchorLeftLocs[t] : maxDNALength, anchorRightLocs ? anchorRightLocs[t] : 0,
                       cutoffThresholds[t], &scores[t], &hitLocs[t], &maxScoreCounts[t], pattern + (size_t)t * patLen, tb, H, E,
                       ckpt, checkpointEvery, slack, &reswept, &restarts);
        cells += (unsigned long long)readLengths[t] * DNALengths[t];
        if (scores[t] >= cutoffThresholds[t]) tracedCols += DNALengths[t];
    }
    free(tb); free(H); free(E); free(ckpt);
    if (stats) { stats[0] += reswept; stats[1] += restarts; stats[2] += tracedCols; }
    return cells;
}
