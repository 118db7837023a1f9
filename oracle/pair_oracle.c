/* TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's paired-end pairing of the two reads' occurrence lists
 * (SURVEY.md 8f row 3), batched over read pairs, for checking s3_pair_occurrences.  Follows
 *   PEMappingOccurrences  PEAlgnmt.cpp:480-547   both lists sorted by position (PERadixSort :114-199, eight stable 4-bit
 *                                                passes over the low 32 bits: ties keep their arrival order), then
 *   PEMappingCore         PEAlgnmt.cpp:229-291   merge walk: the element with the smaller position (list 1 on ties) is the
 *                                                left leg if it has the left-leg strand; it is tried against the other
 *                                                list from that list's cursor until PEIsPairOutOfRange
 *   PEIsPairEndMatch      PEAlgnmt.cpp:566-597   gap = right.pos + patternLength - 1 - left.pos + 1 in uint arithmetic,
 *                                                insertLbound <= gap <= insertUbound compared unsigned, strands = legs
 *   PEIsPairOutOfRange    PEAlgnmt.cpp:599-606   only for different strands: left.pos + insertUbound < right.pos + patternLength - 1
 *   PEReportPairResult    PEAlgnmt.cpp:608-637   fields _1 always from list 1, _2 from list 2
 *   PEStatsPEPairList     PEAlgnmt.cpp:777-831   optimal / suboptimal pair and the histogram of total mismatches
 * patternLength is the second read's length for both orientations (CPUfunctions.cpp:2284).
 * Pinned against those functions compiled from the reference by oracle/build_ref.sh (libref_pair.so):
 * tests/test_cpu_oracle_vs_ref.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t pos; uint8_t strand, mism; uint32_t order; } occ_t;

static int cmp_occ(const void *a, const void *b)
{
    const occ_t *x = (const occ_t *)a, *y = (const occ_t *)b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return x->order < y->order ? -1 : (x->order > y->order);
}

static occ_t *sorted_list(const uint32_t *pos, const uint8_t *strand, const uint8_t *mism, uint64_t lo, uint64_t hi)
{
    uint64_t n = hi - lo;
    occ_t *o = (occ_t *)malloc((n + 1) * sizeof(occ_t));
    for (uint64_t i = 0; i < n; ++i) { o[i].pos = pos[lo + i]; o[i].strand = strand[lo + i]; o[i].mism = mism[lo + i]; o[i].order = (uint32_t)i; }
    qsort(o, n, sizeof(occ_t), cmp_occ);
    return o;
}

static int is_match(const occ_t *l, const occ_t *r, uint32_t patLen, uint32_t lb, uint32_t ub, int sl, int sr, uint32_t *ins)
{
    uint32_t right_end = r->pos + patLen - 1u;
    uint32_t gap = right_end - l->pos + 1u;
    *ins = gap;
    return lb <= gap && gap <= ub && l->strand == sl && r->strand == sr;
}

static int out_of_range(const occ_t *l, const occ_t *r, uint32_t patLen, uint32_t ub)
{
    if (l->strand == r->strand) return 0;
    return (uint32_t)(l->pos + ub) < (uint32_t)(r->pos + patLen - 1u);
}

/* One record per valid pair, in the reference's emission order.  out* may be NULL (count only).
 * pairOffsets[numPairs + 1]; optimal / suboptimal: index of that pair inside its read pair's records, 0xFFFFFFFF = none;
 * mismatchStats: numPairs x 32 counters (the reference's array holds 2 * MAX_NUM_OF_ERROR = 30).  Returns the total. */
uint64_t s3o_pair_occurrences(const uint32_t *pos1, const uint8_t *strand1, const uint8_t *mism1, const uint64_t *off1,
                              const uint32_t *pos2, const uint8_t *strand2, const uint8_t *mism2, const uint64_t *off2,
                              const uint32_t *patternLengths, uint64_t numPairs,
                              int32_t insertLbound, int32_t insertUbound, int strandLeftLeg, int strandRightLeg, int reportOne,
                              uint64_t *pairOffsets, uint32_t *outPos1, uint32_t *outPos2, uint32_t *outInsertion,
                              uint8_t *outFlags /* 4 per record: strand_1, mismatch_1, strand_2, mismatch_2 */,
                              uint64_t outCap, uint32_t *optimal, uint32_t *suboptimal, uint32_t *mismatchStats)
{
    uint64_t total = 0;
    const uint32_t lb = (uint32_t)insertLbound, ub = (uint32_t)insertUbound;
    for (uint64_t p = 0; p < numPairs; ++p) {
        uint64_t n1 = off1[p + 1] - off1[p], n2 = off2[p + 1] - off2[p];
        occ_t *a = sorted_list(pos1, strand1, mism1, off1[p], off1[p + 1]);
        occ_t *b = sorted_list(pos2, strand2, mism2, off2[p], off2[p + 1]);
        const uint32_t pl = patternLengths[p];
        uint64_t i1 = 0, i2 = 0, first = total;
        pairOffsets[p] = total;
        uint32_t opt = 0xFFFFFFFFu, sub = 0xFFFFFFFFu;
        uint8_t optCount = 255, optDiff = 255;
        if (mismatchStats) memset(mismatchStats + p * 32, 0, 32 * sizeof(uint32_t));
#define EMIT(x1, x2, ins) do {                                                                   \
            if (outPos1 && total < outCap) {                                                     \
                outPos1[total] = (x1)->pos; outPos2[total] = (x2)->pos; outInsertion[total] = (ins); \
                outFlags[4 * total] = (x1)->strand; outFlags[4 * total + 1] = (x1)->mism;         \
                outFlags[4 * total + 2] = (x2)->strand; outFlags[4 * total + 3] = (x2)->mism;     \
            }                                                                                    \
            int tot_ = (int8_t)((x1)->mism + (x2)->mism);   /* char totalMismatchCount */          \
            if (mismatchStats && tot_ >= 0 && tot_ < 32) mismatchStats[p * 32 + tot_]++;          \
            int d_ = (int8_t)(x1)->mism - (int8_t)(x2)->mism;                                     \
            if ((int8_t)(x2)->mism > (int8_t)(x1)->mism) d_ = -d_;                                \
            if (tot_ < optCount) { sub = opt; opt = (uint32_t)(total - first); optCount = (uint8_t)tot_; optDiff = (uint8_t)d_; } \
            else if (tot_ == optCount && d_ < optDiff) { opt = (uint32_t)(total - first); optCount = (uint8_t)tot_; optDiff = (uint8_t)d_; } \
            ++total;                                                                             \
        } while (0)
        while (i1 < n1 && i2 < n2) {
            uint32_t ins;
            if (a[i1].pos <= b[i2].pos) {
                if (a[i1].strand == strandLeftLeg)
                    for (uint64_t i = i2; i < n2; ++i) {
                        if (is_match(&a[i1], &b[i], pl, lb, ub, strandLeftLeg, strandRightLeg, &ins)) {
                            EMIT(&a[i1], &b[i], ins);
                            if (reportOne) break;
                        }
                        if (out_of_range(&a[i1], &b[i], pl, ub)) break;
                    }
                ++i1;
            } else {
                if (b[i2].strand == strandLeftLeg)
                    for (uint64_t i = i1; i < n1; ++i) {
                        if (is_match(&b[i2], &a[i], pl, lb, ub, strandLeftLeg, strandRightLeg, &ins)) {
                            EMIT(&a[i], &b[i2], ins);
                            if (reportOne) break;
                        }
                        if (out_of_range(&b[i2], &a[i], pl, ub)) break;
                    }
                ++i2;
            }
        }
#undef EMIT
        if (optimal) optimal[p] = opt;
        if (suboptimal) suboptimal[p] = sub;
        free(a); free(b);
    }
    pairOffsets[numPairs] = total;
    return total;
}
