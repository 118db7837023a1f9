/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * Runs the reference's own DP kernels (SemiGlobalAligntment / GPUBacktrack,
 * DV-DPfunctions.cu:243,316 and their helpers :35-241), compiled for the host
 * through cuda_host_shim.h.  dp_kernels.inc is generated at build time by sed
 * from DV-DPfunctions.cu lines 35-512 with the two CUDA-12-removed texture
 * references replaced by plain array reads.
 */
#include "cuda_host_shim.h"
#include <stdio.h>
#include <omp.h>
typedef unsigned int uint;
typedef unsigned char uchar;
#define DP_THREADS_PER_BLOCK 128
#define MC_CeilDivide16(x) ((x+15)>>4)
#include "dp_kernels.inc"

extern "C" {

/* One SemiGlobalAligner::performAlignment (DV-DPfunctions.cu:669-725) in scheme
 * `scheme` over numOfThreads alignments; arrays are the reference's host batch
 * arrays.  clipRtSizes is copied (the device copy is overwritten by the kernel). */
int ref_dp_align(const uint *packedDNASequence, const uint *DNALengths, uint maxDNALength, uint maxDPTableLength,
                 const uint *packedReadSequence, const uint *readLengths, uint maxReadLength,
                 const int *cutoffThresholds, int *scores, uint *hitLocs, uint *maxScoreCounts, uchar *pattern,
                 uint numOfThreads, const uint *clipLtSizes, const uint *clipRtSizes,
                 const uint *anchorLeftLocs, const uint *anchorRightLocs,
                 int MatchScore, int MismatchScore, int GapOpenScore, int GapExtendScore, int scheme, int nthreads)
{
    const uint groups = (numOfThreads + 31) / 32;
    const uint dnaW = MC_CeilDivide16(maxDNALength) * 32, readW = MC_CeilDivide16(maxReadLength) * 32;
    const size_t tableShorts = (size_t)2 * maxDPTableLength * maxReadLength * 32;
    const uint patLen = maxReadLength + maxDPTableLength;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    #pragma omp parallel num_threads(nthreads)
    {
        short *table = (short *)malloc(tableShorts * sizeof(short) + 64);
        uint startOffsets[32], clipRt[32];
        #pragma omp for schedule(dynamic, 1)
        for (long long g = 0; g < (long long)groups; ++g) {
            const uint base = (uint)g * 32;
            const uint cnt = numOfThreads - base < 32 ? numOfThreads - base : 32;
            if (clipRtSizes) for (uint t = 0; t < cnt; ++t) clipRt[t] = clipRtSizes[base + t];
            for (int pass = 0; pass < 2; ++pass)
                for (uint t = 0; t < cnt; ++t) {
                    blockIdx.x = 0; threadIdx.x = t;
                    if (pass == 0)
                        SemiGlobalAligntment((uint *)packedDNASequence + (size_t)g * dnaW, (uint *)DNALengths + base,
                                             maxDNALength, maxDPTableLength,
                                             (uint *)packedReadSequence + (size_t)g * readW, (uint *)readLengths + base,
                                             maxReadLength, scores + base, hitLocs + base, startOffsets,
                                             clipLtSizes ? (uint *)clipLtSizes + base : NULL, clipRtSizes ? clipRt : NULL,
                                             anchorLeftLocs ? (uint *)anchorLeftLocs + base : NULL,
                                             anchorRightLocs ? (uint *)anchorRightLocs + base : NULL, cnt,
                                             MatchScore, MismatchScore, GapOpenScore, GapExtendScore,
                                             table, maxScoreCounts + base, scheme);
                    else
                        GPUBacktrack((uint *)packedDNASequence + (size_t)g * dnaW, (uint *)DNALengths + base,
                                     maxDNALength, maxDPTableLength,
                                     (uint *)packedReadSequence + (size_t)g * readW, (uint *)readLengths + base,
                                     maxReadLength, scores + base, hitLocs + base, startOffsets,
                                     clipLtSizes ? (uint *)clipLtSizes + base : NULL, clipRtSizes ? clipRt : NULL,
                                     anchorLeftLocs ? (uint *)anchorLeftLocs + base : NULL, cnt,
                                     MatchScore, MismatchScore, GapOpenScore, GapExtendScore,
                                     (int *)cutoffThresholds + base, table, pattern + (size_t)base * patLen);
                }
        }
        free(table);
    }
    return 0;
}

} /* extern "C" */
