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
