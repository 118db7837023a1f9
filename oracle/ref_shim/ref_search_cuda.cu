/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref_search_cuda.sh.
 * The reference's own search kernels (DV-Kernel.cu:4249 kernel, :4505 kernel_4mismatch_1, :4741 kernel_4mismatch_2),
 * unmodified, compiled for sm_100a and run ON THE GPU, launched the way perform_round1_alignment launches them
 * (alignment.cu:118-215: one launch per case over the whole batch, blocks of THREADS_PER_BLOCK threads, one read per
 * thread, answers copied back after every case).  The "kernel to beat" on the same B200 for bench.py's baseline leg,
 * and a second opinion for parity.  Nothing of the product links or loads this file.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef unsigned int uint;
#include "DV-Kernel.cu"           /* found through -I$REF */

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fprintf(stderr, "[ref_search_cuda] %s: %s\n", #call, cudaGetErrorString(e__)); return -1; } } while (0)

static uint *g_bwt, *g_occ, *g_rbwt, *g_rocc;

extern "C" int ref_search_cuda_upload(const uint *bwt, const uint *revBwt, size_t bwtWords, const uint *occ, const uint *revOcc, size_t occWords)
{
    CK(cudaMalloc(&g_bwt, bwtWords * 4)); CK(cudaMalloc(&g_rbwt, bwtWords * 4));
    CK(cudaMalloc(&g_occ, occWords * 4)); CK(cudaMalloc(&g_rocc, occWords * 4));
    CK(cudaMemcpy(g_bwt, bwt, bwtWords * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(g_rbwt, revBwt, bwtWords * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g_occ, occ, occWords * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(g_rocc, revOcc, occWords * 4, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" void ref_search_cuda_free(void)
{
    cudaFree(g_bwt); cudaFree(g_rbwt); cudaFree(g_occ); cudaFree(g_rocc);
    g_bwt = g_rbwt = g_occ = g_rocc = NULL;
}

/* round 1 of perform_round1_alignment; kernelMs = the case launches only (copies excluded) */
extern "C" int ref_search_cuda_round1(const uint *queries, const uint *readLengths, uint batchSize, uint wordPerQuery,
                                      uint inverseSa0, uint revInverseSa0, uint textLength, uint *const *answers,
                                      uint numMismatch, uint numCases, uint sa_range_allowed, uint word_per_ans,
                                      int isExactNumMismatch, float *kernelMs)
{
    const size_t roundUp = ((size_t)batchSize + 31) / 32 * 32;
    const uint blocksNeeded = (batchSize + THREADS_PER_BLOCK * QUERIES_PER_THREAD - 1) / (THREADS_PER_BLOCK * QUERIES_PER_THREAD);
    uint *_queries, *_readLengths, *_answers;
    bool *_isBad;
    CK(cudaMalloc(&_queries, roundUp * wordPerQuery * 4)); CK(cudaMalloc(&_readLengths, roundUp * 4));
    CK(cudaMalloc(&_answers, roundUp * word_per_ans * 4)); CK(cudaMalloc(&_isBad, roundUp));
    CK(cudaMemset(_isBad, 0, roundUp));
    CK(cudaMemcpy(_queries, queries, roundUp * wordPerQuery * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(_readLengths, readLengths, roundUp * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float total = 0.f;
    for (uint caseno = 0; caseno < numCases; ++caseno) {
        CK(cudaEventRecord(e0));
        if (numMismatch <= 3)
            kernel<<<blocksNeeded, THREADS_PER_BLOCK>>>(caseno, _queries, _readLengths, batchSize, wordPerQuery, g_bwt, g_occ, inverseSa0,
                                                        g_rbwt, g_rocc, revInverseSa0, textLength, _answers, _isBad, 0, numMismatch,
                                                        sa_range_allowed, word_per_ans, isExactNumMismatch != 0);
        else if (caseno < 5)
            kernel_4mismatch_1<<<blocksNeeded, THREADS_PER_BLOCK>>>(caseno, _queries, _readLengths, batchSize, wordPerQuery, g_bwt, g_occ, inverseSa0,
                                                                    g_rbwt, g_rocc, revInverseSa0, textLength, _answers, _isBad, 0,
                                                                    sa_range_allowed, word_per_ans, isExactNumMismatch != 0);
        else
            kernel_4mismatch_2<<<blocksNeeded, THREADS_PER_BLOCK>>>(caseno, _queries, _readLengths, batchSize, wordPerQuery, g_bwt, g_occ, inverseSa0,
                                                                    g_rbwt, g_rocc, revInverseSa0, textLength, _answers, _isBad, 0,
                                                                    sa_range_allowed, word_per_ans, isExactNumMismatch != 0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
        CK(cudaMemcpy(answers[caseno], _answers, roundUp * word_per_ans * 4, cudaMemcpyDeviceToHost));
    }
    if (kernelMs) *kernelMs = total;
    cudaFree(_queries); cudaFree(_readLengths); cudaFree(_answers); cudaFree(_isBad);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
