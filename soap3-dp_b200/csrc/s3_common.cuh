// s3_common.cuh -- shared device structures of the B200 hot path.
//
// Index layout in HBM (DESIGN.md "data layout"): per direction an array of
// 64-byte, 64-byte-aligned buckets
//     struct { uint32 cnt[4]; uint32 bwt[12]; }
// cnt[c] = cumulativeFreq[c] + Occ(c, 192*b) on the $-less BWT, bwt = the next
// 192 bases (2 bit/base, 16 per word, MSB first -- the reference's own word
// format, 2bwt-lib/BWT.c:119-175).  One rank evaluation = one 64-byte bucket =
// one 32-byte-sector pair, instead of the reference's 16 B occ + 16 B BWT loads
// from two different cache lines (DV-Kernel.cu:256-280).
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
__device__ __forceinline__ void s3_rank4(const S3Half &h, uint32_t idx, uint32_t out[4])
{
    idx -= (idx > h.inverseSa0);
    const uint32_t b = __umulhi(idx, 0xAAAAAAABu) >> 7;          // idx / 192
    const uint32_t rem = idx - b * S3_BUCKET_BASES;
    const uint4 *p = h.buckets + (size_t)b * 4;
    const uint4 cnt = __ldg(p);
    const uint4 w0 = __ldg(p + 1), w1 = __ldg(p + 2), w2 = __ldg(p + 3);
    const unsigned long long chunk[6] = {
        ((unsigned long long)w0.x << 32) | w0.y, ((unsigned long long)w0.z << 32) | w0.w,
        ((unsigned long long)w1.x << 32) | w1.y, ((unsigned long long)w1.z << 32) | w1.w,
        ((unsigned long long)w2.x << 32) | w2.y, ((unsigned long long)w2.z << 32) | w2.w};
    uint32_t nT = 0, nHi = 0, nLo = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int nb = min(max((int)rem - 32 * k, 0), 32);
        const unsigned long long mask = (nb >= 32) ? ~0ull : ~(~0ull >> (2 * nb));
        const unsigned long long vm = mask & 0x5555555555555555ull;
        const unsigned long long hi = (chunk[k] >> 1) & vm;
        const unsigned long long lo = chunk[k] & vm;
        nT += __popcll(hi & lo);
        nHi += __popcll(hi);
        nLo += __popcll(lo);
    }
    const uint32_t cT = nT, cG = nHi - nT, cC = nLo - nT, cA = rem - nHi - nLo + nT;
    out[0] = cnt.x + cA;
    out[1] = cnt.y + cC;
    out[2] = cnt.z + cG;
    out[3] = cnt.w + cT;
}
