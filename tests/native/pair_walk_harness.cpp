/* CPU-tier harness: compiles the product's csrc/s3_pair_walk.cuh -- the per-read-pair walk both kernels of
 * s3_pair_occurrences run -- with a host compiler and drives it the way csrc/s3_pair.cu does: keys read pair << 32 |
 * position with the caller's index as value, one stable sort per list, a count pass, an exclusive sum, a fill pass.
 * Built by tests/test_cpu_pair_walk.py into tests/native/_build/; not part of the product library.
 */
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>
#include "s3_pair_walk.cuh"

static void sorted_keys(const uint32_t *pos, const uint64_t *off, uint64_t numPairs, std::vector<unsigned long long> &key, std::vector<uint32_t> &val)
{
    const uint64_t n = off[numPairs];
    std::vector<unsigned long long> k(n);
    std::vector<uint32_t> order(n);
    for (uint64_t p = 0; p < numPairs; ++p)
        for (uint64_t e = off[p]; e < off[p + 1]; ++e) k[e] = ((unsigned long long)p << 32) | pos[e];
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return k[a] < k[b]; });
    key.resize(n); val.resize(n);
    for (uint64_t i = 0; i < n; ++i) { key[i] = k[order[i]]; val[i] = order[i]; }
}

extern "C" uint64_t harness_pair_occurrences(const uint32_t *pos1, const uint8_t *strand1, const uint8_t *mism1, const uint64_t *off1,
                                             const uint32_t *pos2, const uint8_t *strand2, const uint8_t *mism2, const uint64_t *off2,
                                             const uint32_t *patternLengths, uint64_t numPairs,
                                             int32_t insertLbound, int32_t insertUbound, int strandLeftLeg, int strandRightLeg, int reportOne,
                                             uint64_t *pairOffsets, uint32_t *outPos1, uint32_t *outPos2, uint32_t *outInsertion, uint8_t *outFlags,
                                             uint64_t outCap, uint32_t *optimal, uint32_t *suboptimal, uint32_t *mismatchStats)
{
    std::vector<unsigned long long> k1, k2;
    std::vector<uint32_t> v1, v2;
    sorted_keys(pos1, off1, numPairs, k1, v1);
    sorted_keys(pos2, off2, numPairs, k2, v2);
    S3PairLists L = {k1.data(), k2.data(), v1.data(), v2.data(), strand1, mism1, strand2, mism2};
    S3PairParams P = {(uint32_t)insertLbound, (uint32_t)insertUbound, strandLeftLeg, strandRightLeg, reportOne};
    S3PairOut O = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t total = 0;
    for (uint64_t p = 0; p < numPairs; ++p) {
        pairOffsets[p] = total;
        total += s3_pair_walk<false>(L, P, p, off1[p], off1[p + 1], off2[p], off2[p + 1], patternLengths[p], 0, O);
    }
    pairOffsets[numPairs] = total;
    if (!outPos1 || total > outCap) return total;
    memset(mismatchStats, 0, numPairs * 32 * sizeof(uint32_t));
    O = S3PairOut{outPos1, outPos2, outInsertion, outFlags, optimal, suboptimal, mismatchStats};
    for (uint64_t p = 0; p < numPairs; ++p)
        s3_pair_walk<true>(L, P, p, off1[p], off1[p + 1], off2[p], off2[p + 1], patternLengths[p], pairOffsets[p], O);
    return total;
}

/* ---- the per-read step of s3_retain_best (csrc/s3_retain_walk.cuh), driven like csrc/s3_pair.cu drives it: a count
 * pass, exclusive sums of the kept entries, a fill pass ---- */
#include "s3_retain_walk.cuh"

extern "C" void harness_retain_best(int mode, int32_t maxNum,
                                    const uint32_t *saL, const uint32_t *saR, const uint8_t *saStrand, const uint8_t *saMism, const uint64_t *saOff,
                                    const uint32_t *occPos, const uint8_t *occStrand, const uint8_t *occMism, const uint64_t *occOff, uint64_t numReads,
                                    uint64_t *outSaOff, uint32_t *outSaL, uint32_t *outSaR, uint8_t *outSaFlags,
                                    uint64_t *outOccOff, uint32_t *outOccPos, uint8_t *outOccFlags, uint32_t *num)
{
    S3RetainIn I = {saL, saR, saStrand, saMism, occPos, occStrand, occMism};
    S3RetainOut O = {nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t s = 0, o = 0;
    for (uint64_t r = 0; r < numReads; ++r) {
        uint32_t ks, ko;
        s3_retain_walk<false>(I, mode, maxNum, saOff[r], saOff[r + 1], occOff[r], occOff[r + 1], 0, 0, O, &ks, &ko);
        outSaOff[r] = s; outOccOff[r] = o;
        s += ks; o += ko;
    }
    outSaOff[numReads] = s; outOccOff[numReads] = o;
    O = S3RetainOut{outSaL, outSaR, outSaFlags, outOccPos, outOccFlags};
    for (uint64_t r = 0; r < numReads; ++r) {
        uint32_t ks, ko;
        num[r] = s3_retain_walk<true>(I, mode, maxNum, saOff[r], saOff[r + 1], occOff[r], occOff[r + 1], outSaOff[r], outOccOff[r], O, &ks, &ko);
    }
}
