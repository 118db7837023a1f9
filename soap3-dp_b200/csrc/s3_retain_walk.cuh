// The per-read step of s3_retain_best, shared by its count and fill kernels (csrc/s3_pair.cu).
//
// Follows retainAllBest / retainAllBestWithCap / retainAllBestAndSecBest (SAList.cpp:140-348) on one read's SA-range
// list and occurrence list.  The reference filters in place while its minimum is still moving (a smaller count met later
// empties what was kept so far); what survives is: the SA ranges whose mismatch count is the SA list's minimum -- unless
// an occurrence has fewer, which empties the SA side -- and the occurrences that match the overall minimum, the capped
// variant cutting ranges short / skipping entries once maxNum occurrences are in, in list order.  Plain integer code,
// compiled for the host as well by the CPU tier (tests/native/pair_walk_harness.cpp).
#pragma once
#include <stdint.h>

#ifndef S3_HD
#ifdef __CUDACC__
#define S3_HD __host__ __device__ __forceinline__
#else
#define S3_HD inline
#endif
#endif

#define S3_RETAIN_ALL_BEST        0
#define S3_RETAIN_BEST_WITH_CAP   1
#define S3_RETAIN_BEST_AND_SECOND 2

struct S3RetainIn {
    const uint32_t *saL, *saR; const uint8_t *saStrand, *saMism;
    const uint32_t *occPos; const uint8_t *occStrand, *occMism;
};

struct S3RetainOut {                    // fill pass only
    uint32_t *saL, *saR; uint8_t *saFlags;      // 2 per kept range: strand, mismatchCount
    uint32_t *occPos; uint8_t *occFlags;        // 2 per kept occurrence
};

// Read with SA ranges [s0, s1) and occurrences [o0, o1); kept entries go to saBase.. / occBase.. (FILL).
// Returns the function's return value (number of occurrences retained); *keptSa / *keptOcc = entries kept.
template <bool FILL>
S3_HD uint32_t s3_retain_walk(const S3RetainIn &I, int mode, int32_t maxNum, uint64_t s0, uint64_t s1, uint64_t o0, uint64_t o1,
                              uint64_t saBase, uint64_t occBase, const S3RetainOut &O, uint32_t *keptSa, uint32_t *keptOcc)
{
    int mSa = 999, mOcc = 999;
    for (uint64_t i = s0; i < s1; ++i) if ((int)I.saMism[i] < mSa) mSa = I.saMism[i];
    for (uint64_t i = o0; i < o1; ++i) if ((int)I.occMism[i] < mOcc) mOcc = I.occMism[i];     // uint8_t mismatchCount (2bwt-flex/SRACore.h:86-95)
    uint32_t num = 0, nSa = 0, nOcc = 0;
    if (mode == S3_RETAIN_BEST_AND_SECOND) {
        const int m = mSa < mOcc ? mSa : mOcc;
        for (uint64_t i = s0; i < s1; ++i)
            if ((int)I.saMism[i] <= m + 1) {
                if (FILL) { O.saL[saBase + nSa] = I.saL[i]; O.saR[saBase + nSa] = I.saR[i]; O.saFlags[2 * (saBase + nSa)] = I.saStrand[i]; O.saFlags[2 * (saBase + nSa) + 1] = I.saMism[i]; }
                num += I.saR[i] - I.saL[i] + 1; ++nSa;
            }
        for (uint64_t i = o0; i < o1; ++i)
            if ((int)I.occMism[i] <= m + 1) {
                if (FILL) { O.occPos[occBase + nOcc] = I.occPos[i]; O.occFlags[2 * (occBase + nOcc)] = I.occStrand[i]; O.occFlags[2 * (occBase + nOcc) + 1] = I.occMism[i]; }
                ++num; ++nOcc;
            }
    } else {
        const bool cap = mode == S3_RETAIN_BEST_WITH_CAP;
        if (mSa <= mOcc)                            // an occurrence with fewer mismatches empties the SA side
            for (uint64_t i = s0; i < s1; ++i) {
                if ((int)I.saMism[i] != mSa) continue;
                int cur = (int)(I.saR[i] - I.saL[i] + 1);
                if (cap) {
                    if (nSa == 0) { if (cur > maxNum) cur = maxNum; }              // the entry that sets the minimum
                    else if (!(num < (uint32_t)maxNum)) continue;
                    else if (num + cur > (uint32_t)maxNum) cur = maxNum - (int)num;
                }
                if (FILL) { O.saL[saBase + nSa] = I.saL[i]; O.saR[saBase + nSa] = (uint32_t)cur + I.saL[i] - 1; O.saFlags[2 * (saBase + nSa)] = I.saStrand[i]; O.saFlags[2 * (saBase + nSa) + 1] = I.saMism[i]; }
                num += (uint32_t)cur; ++nSa;
            }
        if (mOcc <= mSa && mOcc != 999) {
            const bool reset = mOcc < mSa;          // the first best occurrence restarts the count at 1, cap or not
            for (uint64_t i = o0; i < o1; ++i) {
                if ((int)I.occMism[i] != mOcc) continue;
                if (cap && !(reset && nOcc == 0) && !(num < (uint32_t)maxNum)) continue;
                if (FILL) { O.occPos[occBase + nOcc] = I.occPos[i]; O.occFlags[2 * (occBase + nOcc)] = I.occStrand[i]; O.occFlags[2 * (occBase + nOcc) + 1] = I.occMism[i]; }
                ++num; ++nOcc;
            }
        }
    }
    *keptSa = nSa; *keptOcc = nOcc;
    return num;
}
