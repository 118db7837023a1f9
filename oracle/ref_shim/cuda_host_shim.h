/* TEST INFRASTRUCTURE ONLY.
 * Minimal CUDA-on-host shim so that the reference's device code
 * (DV-Kernel.cu, DV-DPfunctions.cu:35-512) compiles with g++ and can be run,
 * one "thread" at a time, as the bit-exact oracle for answer slots / DP output.
 * No reference source is copied: the driver .cpp files #include the reference
 * files from where they lie ($REF).
 */
#ifndef S3_CUDA_HOST_SHIM_H
#define S3_CUDA_HOST_SHIM_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#define __device__
#define __global__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __align__(x) __attribute__((aligned(x)))

struct s3_dim3 { unsigned x, y, z; };
static thread_local s3_dim3 blockIdx  = {0, 0, 0};
static thread_local s3_dim3 threadIdx = {0, 0, 0};
static thread_local s3_dim3 blockDim  = {1, 1, 1};

struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };

/* rank-query counter: the unit of SURVEY.md 8(d)'s roofline (one Occ evaluation) */
static thread_local unsigned long long s3_rank_queries = 0;

static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline unsigned long long __brevll(unsigned long long v)
{
    v = ((v >> 1) & 0x5555555555555555ULL) | ((v & 0x5555555555555555ULL) << 1);
    v = ((v >> 2) & 0x3333333333333333ULL) | ((v & 0x3333333333333333ULL) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((v & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(v);
}
static inline unsigned __usad(unsigned x, unsigned y, unsigned z) { return (x > y ? x - y : y - x) + z; }
using std::max;
using std::min;
#endif
