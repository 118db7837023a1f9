// s3_common.cuh -- shared device structures of the B200 hot path.
//
// Index layout in HBM (DESIGN.md "data layout"): per direction an array of
// 64-byte, 64-byte-aligned buckets, one per 192 BWT positions:
//     bytes  0..15  cnt[4]    cnt[c] = cumulativeFreq[c] + Occ(c, 192*b + 96)   (count at the bucket MIDDLE)
//     bytes 16..39  half A    bases 0..95  as bit planes: hi0 hi1 hi2 lo0 | lo1 lo2
//     bytes 40..63  half B    bases 96..191 as bit planes: lo1 lo2 | hi0 hi1 hi2 lo0
// A base's 2-bit code is split into a "hi" and a "lo" plane (base k of a half is bit
// k%32 of plane word k/32), so the four symbol counts of up to 96 bases cost 9 POPC
// (hi, lo, hi&lo per word) and rank'(c, i) counts forward or backward from the middle:
// at most 96 bases, one 16 B + one 16 B + one 8 B load inside ONE 64-byte line.  The
// reference needs a 16 B occ load and a 16 B BWT load from two different cache lines and
// counts 2-bit codes with 64-bit masks (DV-Kernel.cu:27-280).  Positions past the end of
// the text are padded with code 0 and counted as such on both sides of the middle, so
// they cancel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define S3_BUCKET_BASES 192u
#define S3_THREADS 128

struct S3Half {
    const uint4 *buckets;     // 4 x uint4 per bucket
    uint32_t inverseSa0;
    uint32_t numBuckets;
};

struct s3_index {
    int device;
    cudaStream_t stream;
    S3Half fwd, rev;
    uint32_t textLength;
    uint4 *d_fwd, *d_rev;
    uint32_t *d_packedDNA;    // optional
    uint32_t *d_sa;           // optional
    size_t bytes;
    // scratch reused across calls (grown on demand)
    void *scratch; size_t scratchBytes;
    void *pinned; size_t pinnedBytes;
    // persistent search launches
    uint32_t *d_workCounter;
    int numSms;
    size_t searchSmem; int searchBlocksPerSm;
};

void s3_set_error(const char *fmt, ...);
extern unsigned long long g_s3_launches;
#define S3_LAUNCHED(n) (g_s3_launches += (n))
#define S3_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            s3_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,             \
                         cudaGetErrorString(e__));                                      \
            return S3_ECUDA;                                                            \
        }                                                                               \
    } while (0)

int s3_scratch(s3_index *ix, size_t bytes, void **out);
int s3_pinned(s3_index *ix, size_t bytes, void **out);

// ---- rank'(., idx) for all four symbols ------------------------------------
// Mathematically the reference's GPUBWTAllOccValue (DV-Kernel.cu:282-299):
// C[c] + #{c in BWT[0, idx)} with the "$ is not stored" shift.
__device__ __forceinline__ uint32_t s3_shl_clamp(uint32_t a, uint32_t s)
{
    uint32_t d;                                   // PTX shl clamps shift amounts > 32 to 32 (result 0)
    asm("shl.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(s));
    return d;
}

__device__ __forceinline__ void s3_rank4(const S3Half &h, uint32_t idx, uint32_t out[4])
{
    idx -= (idx > h.inverseSa0);
    const uint32_t b = __umulhi(idx, 0xAAAAAAABu) >> 7;          // idx / 192
    const uint32_t rem = idx - b * S3_BUCKET_BASES;
    const bool isB = rem >= 96u;
    const uint32_t x = isB ? rem - 96u : rem;                     // position inside the half
    const uint4 *p = h.buckets + (size_t)b * 4;
    const uint4 cnt = __ldg(p);
    const uint4 hv = __ldg(p + (isB ? 3 : 1));                    // hi0 hi1 hi2 lo0
    const uint2 lv = __ldg(reinterpret_cast<const uint2 *>(p) + (isB ? 5 : 4));   // lo1 lo2
    // half A counts bases [x, 96) (subtracted), half B counts bases [0, x) (added)
    const uint32_t flip = isB ? 0xFFFFFFFFu : 0u;
    const uint32_t m0 = s3_shl_clamp(0xFFFFFFFFu, x) ^ flip;
    const uint32_t m1 = s3_shl_clamp(0xFFFFFFFFu, (uint32_t)max((int)x - 32, 0)) ^ flip;
    const uint32_t m2 = s3_shl_clamp(0xFFFFFFFFu, (uint32_t)max((int)x - 64, 0)) ^ flip;
    const uint32_t h0 = hv.x & m0, h1 = hv.y & m1, h2 = hv.z & m2;
    const uint32_t l0 = hv.w & m0, l1 = lv.x & m1, l2 = lv.y & m2;
    const uint32_t nHi = __popc(h0) + __popc(h1) + __popc(h2);
    const uint32_t nLo = __popc(l0) + __popc(l1) + __popc(l2);
    const uint32_t nT = __popc(h0 & l0) + __popc(h1 & l1) + __popc(h2 & l2);
    const uint32_t len = isB ? x : 96u - x;
    const uint32_t cA = len - nHi - nLo + nT, cC = nLo - nT, cG = nHi - nT;
    out[0] = isB ? cnt.x + cA : cnt.x - cA;
    out[1] = isB ? cnt.y + cC : cnt.y - cC;
    out[2] = isB ? cnt.z + cG : cnt.z - cG;
    out[3] = isB ? cnt.w + nT : cnt.w - nT;
}
