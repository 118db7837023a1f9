/* TEST INFRASTRUCTURE ONLY: CPU restatement of the single-end DP seeding step "seed hits -> candidate positions"
 * of the reference, for checking s3_seed_candidates.  Follows
 *   SingleEndSeedingBatch::decodePositions  DV-DPfunctions.cu:1143-1219  (SA ranges -> estimated read starts:
 *       strand 1: SA[k] - offset;  strand 2: SA[k] + seedLength + offset - readLength, in uint arithmetic; then three
 *       radix sorts by pos, readID, strand -- the last one, MC_RadixSort_8_8 (DV-DPfunctions.h:90-95), sorts into the
 *       auxiliary array that is freed right after, so the hits stay ordered by (readID, pos), ties in arrival order)
 *   SingleEndSeedingBatch::singleMerge      DV-DPfunctions.cu:1101-1141  (per read: the first hit, then every hit more
 *       than DPS_DIVIDE_GAP = 50 beyond the last one kept, `prevLoc + 50 < curLoc` in uint arithmetic)
 * Pinned against the reference's own macros and singleMerge body compiled by oracle/build_ref.sh (libref_seed.so):
 * tests/test_cpu_oracle_vs_ref.py.  maxPerRange caps the positions taken from one range like s3_locate (the reference
 * takes all; its caps sit upstream in the seed-hit limits).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t readID, pos; int32_t strand; uint32_t order; } hit_t;

static int cmp_hit(const void *a, const void *b)
{
    const hit_t *x = (const hit_t *)a, *y = (const hit_t *)b;
    if (x->readID != y->readID) return x->readID < y->readID ? -1 : 1;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return x->order < y->order ? -1 : (x->order > y->order);       /* stable */
}

uint64_t s3o_seed_candidates(const uint32_t *sa, const uint32_t *saL, const uint32_t *saR, const int32_t *strands,
                             const uint32_t *readIDs, const uint32_t *offsets, const uint32_t *seedLengths,
                             const uint32_t *readLengths, uint64_t numRanges, uint32_t maxPerRange,
                             uint32_t *outReadID, uint32_t *outPos, int32_t *outStrand, uint64_t outCap)
{
    uint64_t total = 0;
    for (uint64_t g = 0; g < numRanges; ++g)
        if (saR[g] >= saL[g]) { uint64_t c = (uint64_t)(saR[g] - saL[g]) + 1; total += c > maxPerRange ? maxPerRange : c; }
    hit_t *h = (hit_t *)malloc((total + 1) * sizeof(hit_t));
    uint64_t n = 0;
    for (uint64_t g = 0; g < numRanges; ++g) {
        if (saR[g] < saL[g]) continue;
        uint64_t c = (uint64_t)(saR[g] - saL[g]) + 1;
        if (c > maxPerRange) c = maxPerRange;
        for (uint64_t k = 0; k < c; ++k) {
            const uint32_t x = sa[(uint64_t)saL[g] + k];
            h[n].pos = strands[g] == 1 ? x - offsets[g] : x + seedLengths[g] + offsets[g] - readLengths[g];
            h[n].readID = readIDs[g]; h[n].strand = strands[g]; h[n].order = (uint32_t)n;
            ++n;
        }
    }
    qsort(h, n, sizeof(hit_t), cmp_hit);
    uint64_t m = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i;
        while (j < n && h[j].readID == h[i].readID) ++j;
        uint32_t prev = h[i].pos;
        if (m < outCap) { outReadID[m] = h[i].readID; outPos[m] = h[i].pos; outStrand[m] = h[i].strand; }
        ++m;
        for (uint64_t k = i + 1; k < j; ++k)
            if ((uint32_t)(prev + 50u) < h[k].pos) {
                if (m < outCap) { outReadID[m] = h[k].readID; outPos[m] = h[k].pos; outStrand[m] = h[k].strand; }
                ++m;
                prev = h[k].pos;
            }
        i = j;
    }
    free(h);
    return m;
}

/* ---- paired-end seeding: seed hits of both ends -> candidate (left start, right start) pairs -------------------------
 * Follows PairEndSeedingBatch::decodePositions (DV-DPfunctions.cu:2882-2969: estimated starts as above with
 * strandIndex = strand - 1, key = readID | strandIndex << 31, two guards, sorted by pos then by key),
 * findRevStart (:2626-2653), pairEndMerge (:2780-2880) and the tail of decodeMergePositions (:2976-2999).
 * Quirks kept: the left group is thinned IN PLACE (50-base gap) before the join, so when both legs use the same strand
 * the second call joins against the array the first call left behind; the window is computed from
 * readLengths[pair id] whichever end is on the left; after a match only the left pointer advances; all position
 * arithmetic is uint.  Pinned against the reference's own functions (libref_seed_pair.so) in the CPU tier. */
typedef struct { uint32_t pos, key; } phit_t;
typedef struct { uint32_t id, posL, posR, order; } pcand_t;

static int cmp_phit(const void *a, const void *b)
{
    const phit_t *x = (const phit_t *)a, *y = (const phit_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return 0;                                                      /* equal hits are indistinguishable */
}

static int cmp_pcand(const void *a, const void *b)
{
    const pcand_t *x = (const pcand_t *)a, *y = (const pcand_t *)b;
    if (x->id != y->id) return x->id < y->id ? -1 : 1;
    return x->order < y->order ? -1 : (x->order > y->order);       /* stable, like the reference's radix sort */
}

static void pair_merge(pcand_t **out, uint64_t *m, uint64_t *cap, phit_t *left, phit_t *right, uint32_t leftReadOrMate,
                       const uint32_t *lengthsByReadID, int insLow, int insHigh)
{
    phit_t *li = left, *ri = right;
    for (;;) {
        uint32_t rid = ri->key & 0x7FFFFFFFu;
        while ((li->key & 0x7FFFFFFFu) < rid) ++li;
        const uint32_t lid = li->key & 0x7FFFFFFFu;
        while ((ri->key & 0x7FFFFFFFu) < lid) ++ri;
        rid = ri->key & 0x7FFFFFFFu;
        if (rid == 0x7FFFFFFFu) break;
        if (lid < rid) continue;
        phit_t *ls = li, *rs = ri;
        while ((li->key & 0x7FFFFFFFu) == lid) ++li;
        while ((ri->key & 0x7FFFFFFFu) == rid) ++ri;
        phit_t *le = li, *re = ri;
        {   /* thin the left group in place */
            phit_t *c = ls;
            uint32_t prev = c->pos;
            for (phit_t *p = ls + 1; p < le; ++p)
                if ((uint32_t)(prev + 50u) < p->pos) { *(++c) = *p; prev = p->pos; }
            le = c + 1;
        }
        const int readLength = (int)lengthsByReadID[lid];
        const int margin = readLength > 100 ? (readLength >> 2) : 25;
        int lengthLow = insLow - readLength - margin;
        if (lengthLow < 0) lengthLow = 0;
        const int lengthHigh = insHigh - readLength + margin;
        phit_t *lp = ls, *rp = rs;
        uint32_t lloc = lp->pos, rloc = rp->pos;
        while (lp < le && rp < re) {
            if ((uint32_t)(lloc + (uint32_t)lengthLow) > rloc) { ++rp; rloc = rp->pos; }
            else if ((uint32_t)(lloc + (uint32_t)lengthHigh) < rloc) { ++lp; lloc = lp->pos; }
            else {
                if (*m == *cap) { *cap = *cap * 2 + 1024; *out = (pcand_t *)realloc(*out, *cap * sizeof(pcand_t)); }
                (*out)[*m].id = (lp->key & 0x7FFFFFFFu) + leftReadOrMate; (*out)[*m].posL = lloc; (*out)[*m].posR = rloc; (*out)[*m].order = (uint32_t)*m;
                ++*m;
                ++lp; lloc = lp->pos;
            }
        }
    }
}

static phit_t *pair_side(const uint32_t *sa, const uint32_t *saL, const uint32_t *saR, const int32_t *strands,
                         const uint32_t *readIDs, const uint32_t *offsets, const uint32_t *seedLengths,
                         const uint32_t *readLengths, uint64_t numRanges, uint32_t maxPerRange, uint64_t *len)
{
    uint64_t total = 0;
    for (uint64_t g = 0; g < numRanges; ++g)
        if (saR[g] >= saL[g]) { uint64_t c = (uint64_t)(saR[g] - saL[g]) + 1; total += c > maxPerRange ? maxPerRange : c; }
    phit_t *h = (phit_t *)malloc((total + 2) * sizeof(phit_t));
    uint64_t n = 0;
    for (uint64_t g = 0; g < numRanges; ++g) {
        if (saR[g] < saL[g]) continue;
        uint64_t c = (uint64_t)(saR[g] - saL[g]) + 1;
        if (c > maxPerRange) c = maxPerRange;
        const uint32_t si = (uint32_t)strands[g] - 1u;
        for (uint64_t k = 0; k < c; ++k) {
            const uint32_t x = sa[(uint64_t)saL[g] + k];
            h[n].pos = si == 0 ? x - offsets[g] : x + seedLengths[g] + offsets[g] - readLengths[g];
            h[n].key = readIDs[g] | (si << 31);
            ++n;
        }
    }
    h[n].key = 0x7FFFFFFFu; h[n].pos = 0xFFFFFFFFu; ++n;
    h[n].key = 0x7FFFFFFFu | (1u << 31); h[n].pos = 0xFFFFFFFFu; ++n;
    qsort(h, n, sizeof(phit_t), cmp_phit);
    *len = n;
    return h;
}

uint64_t s3o_seed_pair_candidates(const uint32_t *sa,
                                  const uint32_t *saL0, const uint32_t *saR0, const int32_t *strands0, const uint32_t *readIDs0,
                                  const uint32_t *offsets0, const uint32_t *seedLengths0, const uint32_t *readLengths0, uint64_t n0,
                                  const uint32_t *saL1, const uint32_t *saR1, const int32_t *strands1, const uint32_t *readIDs1,
                                  const uint32_t *offsets1, const uint32_t *seedLengths1, const uint32_t *readLengths1, uint64_t n1,
                                  uint32_t maxPerRange, const uint32_t *lengthsByReadID, int insertLow, int insertHigh,
                                  int peStrandLeftLeg, int peStrandRightLeg,
                                  uint32_t *outID, uint32_t *outPosL, uint32_t *outPosR, uint64_t outCap)
{
    uint64_t len[2];
    phit_t *side[2];
    side[0] = pair_side(sa, saL0, saR0, strands0, readIDs0, offsets0, seedLengths0, readLengths0, n0, maxPerRange, &len[0]);
    side[1] = pair_side(sa, saL1, saR1, strands1, readIDs1, offsets1, seedLengths1, readLengths1, n1, maxPerRange, &len[1]);
    phit_t *arr[2][2];
    for (int s = 0; s < 2; ++s) {
        uint64_t rev = 0;
        while (rev < len[s] && !(side[s][rev].key >> 31)) ++rev;
        arr[s][0] = side[s]; arr[s][1] = side[s] + rev;
    }
    pcand_t *c = NULL;
    uint64_t m = 0, cap = 0;
    pair_merge(&c, &m, &cap, arr[0][peStrandLeftLeg - 1], arr[1][peStrandRightLeg - 1], 0, lengthsByReadID, insertLow, insertHigh);
    pair_merge(&c, &m, &cap, arr[1][peStrandLeftLeg - 1], arr[0][peStrandRightLeg - 1], 1, lengthsByReadID, insertLow, insertHigh);
    if (m) qsort(c, m, sizeof(pcand_t), cmp_pcand);
    for (uint64_t i = 0; i < m && i < outCap; ++i) { outID[i] = c[i].id; outPosL[i] = c[i].posL; outPosR[i] = c[i].posR; }
    free(c); free(side[0]); free(side[1]);
    return m;
}
