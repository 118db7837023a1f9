/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * Runs the reference's own search kernels (DV-Kernel.cu:4249,4505,4741),
 * compiled for the host through cuda_host_shim.h, over a batch of reads:
 * one call == one kernel launch of the reference (one case, both strands).
 */
#include "cuda_host_shim.h"
typedef unsigned int uint;
#ifdef S3_COUNT_RANK_QUERIES
#include "DV-Kernel.counted.cu"   /* build_ref.sh: sed-generated copy with ++s3_rank_queries in the 3 Occ functions */
#else
#include "DV-Kernel.cu"           /* found through -I$REF */
#endif
#include <omp.h>

extern "C" {

/* Mirrors the <<<blocksNeeded,128>>> launch in alignment.cu:170-199.
 * queries is modified in place (reverse-complemented) exactly like the device
 * buffer of the reference.  Returns the number of rank queries executed. */
unsigned long long ref_search_launch(uint whichCase, uint *queries, uint *readLengths, uint numQueries,
                                     uint wordPerQuery, uint *bwt, uint *occ, uint inverseSa0,
                                     uint *revBwt, uint *revOcc, uint revInverseSa0, uint textLength,
                                     uint *answers, unsigned char *isBad, uint round, uint numMismatch,
                                     uint sa_range_allowed, uint wordPerAnswer, int isExactNumMismatch,
                                     int nthreads)
{
    unsigned long long total = 0;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    #pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) reduction(+:total)
    for (long long q = 0; q < (long long)numQueries; ++q) {
        blockIdx.x = (unsigned)(q / THREADS_PER_BLOCK);
        threadIdx.x = (unsigned)(q % THREADS_PER_BLOCK);
        s3_rank_queries = 0;
        if (numMismatch <= 3)
            kernel(whichCase, queries, readLengths, numQueries, wordPerQuery, bwt, occ, inverseSa0,
                   revBwt, revOcc, revInverseSa0, textLength, answers, (bool *)isBad, round, numMismatch,
                   sa_range_allowed, wordPerAnswer, isExactNumMismatch != 0);
        else if (whichCase < 5)
            kernel_4mismatch_1(whichCase, queries, readLengths, numQueries, wordPerQuery, bwt, occ, inverseSa0,
                   revBwt, revOcc, revInverseSa0, textLength, answers, (bool *)isBad, round,
                   sa_range_allowed, wordPerAnswer, isExactNumMismatch != 0);
        else
            kernel_4mismatch_2(whichCase, queries, readLengths, numQueries, wordPerQuery, bwt, occ, inverseSa0,
                   revBwt, revOcc, revInverseSa0, textLength, answers, (bool *)isBad, round,
                   sa_range_allowed, wordPerAnswer, isExactNumMismatch != 0);
        total += s3_rank_queries;
    }
    return total;
}

/* single rank probe of the reference's GPUBWTOccValue (DV-Kernel.cu:256) */
uint ref_rank(uint *bwt, uint *occ, uint index, int c, uint inverseSa0)
{
    return GPUBWTOccValue(bwt, occ, index, (char)c, inverseSa0);
}

} /* extern "C" */
