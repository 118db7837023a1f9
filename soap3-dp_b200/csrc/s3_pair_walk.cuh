// The per-read-pair walk of s3_pair_occurrences, shared by its count and fill kernels (csrc/s3_pair.cu).
//
// Follows PEMappingCore / PEIsPairEndMatch / PEIsPairOutOfRange / PEReportPairResult / PEStatsPEPairList
// (PEAlgnmt.cpp:229-291, 566-637, 777-831) over the two lists of one read pair, already ordered by position with ties in
// arrival order (what PERadixSort leaves).  Plain integer code with no CUDA in it besides the S3_HD qualifier, so the
// CPU tier compiles this very file with a host compiler and checks it against the pairing oracle
// (tests/pair_walk_harness.cpp); the kernels add nothing but the loop over read pairs.
#pragma once
#include <stdint.h>

#ifndef S3_HD
#ifdef __CUDACC__
#define S3_HD __host__ __device__ __forceinline__
#else
#define S3_HD inline
#endif
#endif

// sorted lists: key = pair index << 32 | position, val = index of the occurrence in the caller's arrays
struct S3PairLists {
    const unsigned long long *key1, *key2;
    const uint32_t *val1, *val2;
    const uint8_t *strand1, *mism1, *strand2, *mism2;      // caller order, indexed through val
};

struct S3PairParams {
    uint32_t lbound, ubound;        // insertLbound / insertUbound, compared unsigned like the reference
    int leftLeg, rightLeg;          // strandLeftLeg / strandRightLeg (1 or 2)
    int reportOne;                  // PE_REPORT_ONE: stop at a left leg's first partner
};

struct S3PairOut {                  // record arrays (fill pass only)
    uint32_t *pos1, *pos2, *insertion;
    uint8_t *flags;                 // 4 per record: strand_1, mismatch_1, strand_2, mismatch_2
    uint32_t *optimal, *suboptimal; // per read pair: record index inside the pair's records, 0xFFFFFFFF = none
    uint32_t *stats;                // per read pair: 32 counters by total mismatches
};

// Walks read pair p: lists [a0, a1) and [b0, b1), records from `base` on.  Returns the number of records.
template <bool FILL>
S3_HD uint32_t s3_pair_walk(const S3PairLists &L, const S3PairParams &P, uint64_t p, uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1,
                            uint32_t patternLength, uint64_t base, const S3PairOut &O)
{
    uint32_t n = 0, opt = 0xFFFFFFFFu, sub = 0xFFFFFFFFu;
    uint32_t optCount = 255, optDiff = 255;
    uint64_t i1 = a0, i2 = b0;
    while (i1 < a1 && i2 < b1) {
        const uint32_t pa = (uint32_t)L.key1[i1], pb = (uint32_t)L.key2[i2];
        const bool firstIsLeft = pa <= pb;                                  // list 1 goes first on equal positions
        // the element that goes first is the left leg; it is tried against the other list from that list's cursor
        const uint32_t lv = firstIsLeft ? L.val1[i1] : L.val2[i2];
        const uint32_t lpos = firstIsLeft ? pa : pb;
        const uint8_t lstrand = firstIsLeft ? L.strand1[lv] : L.strand2[lv];
        if (lstrand == P.leftLeg) {
            const uint64_t end = firstIsLeft ? b1 : a1;
            for (uint64_t i = firstIsLeft ? i2 : i1; i < end; ++i) {
                const uint32_t rv = firstIsLeft ? L.val2[i] : L.val1[i];
                const uint32_t rpos = (uint32_t)(firstIsLeft ? L.key2[i] : L.key1[i]);
                const uint8_t rstrand = firstIsLeft ? L.strand2[rv] : L.strand1[rv];
                const uint32_t rightEnd = rpos + patternLength - 1u;
                const uint32_t gap = rightEnd - lpos + 1u;
                bool stop = false;
                if (P.lbound <= gap && gap <= P.ubound && rstrand == P.rightLeg) {
                    // fields _1 come from list 1 and _2 from list 2 whichever is the left leg
                    const uint32_t v1 = firstIsLeft ? lv : rv, v2 = firstIsLeft ? rv : lv;
                    const uint8_t m1 = L.mism1[v1], m2 = L.mism2[v2];
                    if (FILL) {
                        const uint64_t r = base + n;
                        O.pos1[r] = firstIsLeft ? lpos : rpos;
                        O.pos2[r] = firstIsLeft ? rpos : lpos;
                        O.insertion[r] = gap;
                        O.flags[4 * r] = L.strand1[v1]; O.flags[4 * r + 1] = m1;
                        O.flags[4 * r + 2] = L.strand2[v2]; O.flags[4 * r + 3] = m2;
                        const int tot = (int8_t)(uint8_t)(m1 + m2);                 // char totalMismatchCount
                        if (tot >= 0 && tot < 32) O.stats[p * 32 + tot]++;
                        int d = (int)(int8_t)m1 - (int)(int8_t)m2;
                        if ((int8_t)m2 > (int8_t)m1) d = -d;
                        if (tot < (int)optCount) { sub = opt; opt = n; optCount = (uint8_t)tot; optDiff = (uint8_t)d; }
                        else if (tot == (int)optCount && d < (int)optDiff) { opt = n; optCount = (uint8_t)tot; optDiff = (uint8_t)d; }
                    }
                    ++n;
                    stop = P.reportOne != 0;
                }
                if (stop) break;
                if (lstrand != rstrand && (uint32_t)(lpos + P.ubound) < rightEnd) break;     // PEIsPairOutOfRange
            }
        }
        if (firstIsLeft) ++i1; else ++i2;
    }
    if (FILL) { O.optimal[p] = opt; O.suboptimal[p] = sub; }
    return n;
}
