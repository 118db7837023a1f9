/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * The reference's own DP kernels (SemiGlobalAligntment / GPUBacktrack, DV-DPfunctions.cu:243,316 and their helpers
 * :35-241) compiled for sm_100a and run ON THE GPU, launched the way SemiGlobalAligner::performAlignment launches
 * them (DV-DPfunctions.cu:669-725: blocks of DP_THREADS_PER_BLOCK = 128 threads, one alignment per thread, at most
 * numOfBlocks = 64 blocks per call, full H/E tables in global memory).  dp_kernels.inc is generated at build time by
 * sed from DV-DPfunctions.cu lines 35-512 with the two CUDA-12-removed texture references replaced by plain array
 * reads.  Used by bench.py's baseline leg as the "kernel to beat" on the same B200 and as a
 * second opinion for parity; nothing of the product links or loads this file.
 */
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned int uint;
typedef unsigned char uchar;
#define DP_THREADS_PER_BLOCK 128
#define MC_CeilDivide16(x) ((x+15)>>4)
#include "dp_kernels.inc"

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fprintf(stderr, "[ref_dp_cuda] %s: %s\n", #call, cudaGetErrorString(e__)); return -1; } } while (0)

extern "C" int ref_dp_cuda_align(const uint *packedDNASequence, const uint *DNALengths, uint maxDNALength, uint maxDPTableLength,
                                 const uint *packedReadSequence, const uint *readLengths, uint maxReadLength,
                                 const int *cutoffThresholds, int *scores, uint *hitLocs, uint *maxScoreCounts, uchar *pattern,
                                 uint numOfThreads, const uint *clipLtSizes, const uint *clipRtSizes,
                                 const uint *anchorLeftLocs, const uint *anchorRightLocs,
                                 int MatchScore, int MismatchScore, int GapOpenScore, int GapExtendScore,
                                 int numOfBlocks, float *kernelMs)
{
    const uint batch = (uint)numOfBlocks * DP_THREADS_PER_BLOCK;
    const size_t dnaW = MC_CeilDivide16(maxDNALength), readW = MC_CeilDivide16(maxReadLength);
    const size_t patLen = maxReadLength + maxDPTableLength;
    const size_t tableBytes = (size_t)2 * maxDPTableLength * maxReadLength * sizeof(short) * batch;
    uint *d_dna, *d_dnaLen, *d_read, *d_readLen, *d_hit, *d_start, *d_clipLt = NULL, *d_clipRt = NULL, *d_ancL = NULL, *d_ancR = NULL, *d_cnt;
    int *d_score, *d_cutoff;
    uchar *d_pat;
    void *d_table;
    CK(cudaMalloc(&d_dna, batch * dnaW * 4)); CK(cudaMalloc(&d_read, batch * readW * 4));
    CK(cudaMalloc(&d_dnaLen, batch * 4)); CK(cudaMalloc(&d_readLen, batch * 4)); CK(cudaMalloc(&d_hit, batch * 4));
    CK(cudaMalloc(&d_start, batch * 4)); CK(cudaMalloc(&d_cnt, batch * 4)); CK(cudaMalloc(&d_score, batch * 4));
    CK(cudaMalloc(&d_cutoff, batch * 4)); CK(cudaMalloc(&d_pat, batch * patLen)); CK(cudaMalloc(&d_table, tableBytes));
    if (clipLtSizes) CK(cudaMalloc(&d_clipLt, batch * 4));
    if (clipRtSizes) CK(cudaMalloc(&d_clipRt, batch * 4));
    if (anchorLeftLocs) CK(cudaMalloc(&d_ancL, batch * 4));
    if (anchorRightLocs) CK(cudaMalloc(&d_ancR, batch * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float total = 0.f;
    for (uint first = 0; first < numOfThreads; first += batch) {
        const uint n = numOfThreads - first < batch ? numOfThreads - first : batch, up = (n + 31) / 32 * 32;
        CK(cudaMemcpy(d_dna, packedDNASequence + (size_t)first * dnaW, up * dnaW * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_read, packedReadSequence + (size_t)first * readW, up * readW * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_dnaLen, DNALengths + first, n * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_readLen, readLengths + first, n * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_cutoff, cutoffThresholds + first, n * 4, cudaMemcpyHostToDevice));
        if (clipLtSizes) CK(cudaMemcpy(d_clipLt, clipLtSizes + first, n * 4, cudaMemcpyHostToDevice));
        if (clipRtSizes) CK(cudaMemcpy(d_clipRt, clipRtSizes + first, n * 4, cudaMemcpyHostToDevice));
        if (anchorLeftLocs) CK(cudaMemcpy(d_ancL, anchorLeftLocs + first, n * 4, cudaMemcpyHostToDevice));
        if (anchorRightLocs) CK(cudaMemcpy(d_ancR, anchorRightLocs + first, n * 4, cudaMemcpyHostToDevice));
        const int blocksNeeded = (n + DP_THREADS_PER_BLOCK - 1) / DP_THREADS_PER_BLOCK;
        CK(cudaEventRecord(e0));
        SemiGlobalAligntment<<<blocksNeeded, DP_THREADS_PER_BLOCK>>>(d_dna, d_dnaLen, maxDNALength, maxDPTableLength, d_read, d_readLen,
                                                                      maxReadLength, d_score, d_hit, d_start, d_clipLt, d_clipRt, d_ancL, d_ancR, n,
                                                                      MatchScore, MismatchScore, GapOpenScore, GapExtendScore, d_table, d_cnt, 1);
        GPUBacktrack<<<blocksNeeded, DP_THREADS_PER_BLOCK>>>(d_dna, d_dnaLen, maxDNALength, maxDPTableLength, d_read, d_readLen, maxReadLength,
                                                              d_score, d_hit, d_start, d_clipLt, d_clipRt, d_ancL, n,
                                                              MatchScore, MismatchScore, GapOpenScore, GapExtendScore, d_cutoff, d_table, d_pat);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
        CK(cudaMemcpy(scores + first, d_score, n * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hitLocs + first, d_hit, n * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(maxScoreCounts + first, d_cnt, n * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pattern + (size_t)first * patLen, d_pat, (size_t)n * patLen, cudaMemcpyDeviceToHost));
    }
    if (kernelMs) *kernelMs = total;
    cudaFree(d_dna); cudaFree(d_read); cudaFree(d_dnaLen); cudaFree(d_readLen); cudaFree(d_hit); cudaFree(d_start); cudaFree(d_cnt);
    cudaFree(d_score); cudaFree(d_cutoff); cudaFree(d_pat); cudaFree(d_table);
    if (d_clipLt) cudaFree(d_clipLt); if (d_clipRt) cudaFree(d_clipRt); if (d_ancL) cudaFree(d_ancL); if (d_ancR) cudaFree(d_ancR);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
