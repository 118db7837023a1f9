// s3_dp_make_windows: candidates -> DP windows on the device (see s3_windows.cuh for what is decided and where the
// reference decides it).  One thread per candidate; the half-end mode, which yields 0..2 windows per candidate, runs a
// count pass, a prefix sum and a fill pass so that the windows come out in candidate order like the reference's batches.
#include "s3_common.cuh"
#include "s3_windows.cuh"
#include "../../include/soap3dp_b200.h"

#include <cub/cub.cuh>
#include <stdlib.h>
#include <string.h>

struct S3WinIn {
    const uint32_t *readIDs, *pos, *pos2, *leftStarts, *leftHitLocs, *readLengths;
    const uint8_t *strands;
    const int32_t *leftScores;
};
struct S3WinOut {
    uint32_t *candidate, *readIDs, *start, *dnaLen, *clipLt, *clipRt, *ancL, *ancR;
    uint8_t *strands, *leftOrRight;
    int32_t *cutoff;
};

__device__ __forceinline__ void s3_win_store(const S3WinOut &o, uint32_t k, uint32_t cand, const S3Window &x)
{
    o.candidate[k] = cand; o.readIDs[k] = x.readID; o.start[k] = x.start; o.dnaLen[k] = x.dnaLen; o.clipLt[k] = x.clipLt; o.clipRt[k] = x.clipRt;
    o.ancL[k] = x.ancL; o.ancR[k] = x.ancR; o.strands[k] = x.strand; o.leftOrRight[k] = x.leftOrRight; o.cutoff[k] = x.cutoff;
}

template <bool FILL>
__global__ void s3_make_windows_kernel(int mode, S3WinParams w, S3WinIn in, uint32_t n, uint32_t *__restrict__ count, const uint32_t *__restrict__ off, S3WinOut out)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    S3Window x[2];
    int k = 0;
    const uint32_t id = in.readIDs[c];
    if (mode == S3_WIN_SINGLE) { s3_win_single(w, id, in.pos[c], in.strands[c], in.readLengths[id], x[0]); k = 1; }
    else if (mode == S3_WIN_HALF) k = s3_win_half(w, id, in.pos[c], in.strands[c], in.readLengths[id], in.readLengths[id ^ 1u], x);
    else if (mode == S3_WIN_PAIR_LEFT) { s3_win_pair_left(w, id, in.pos[c], in.readLengths[id], x[0]); k = 1; }
    else {
        // the right read is aligned only when the left one reached its cutoff (packRight, DV-DPfunctions.cu:3427)
        if (in.leftScores[c] >= s3_win_cutoff(w, id, in.readLengths[id])) {
            s3_win_pair_right(w, id, in.pos2[c], in.leftStarts[c] + in.leftHitLocs[c], in.readLengths[id ^ 1u], x[0]); k = 1;
        }
    }
    if (!FILL) { count[c] = (uint32_t)k; return; }
    for (int j = 0; j < k; ++j) s3_win_store(out, off[c] + j, c, x[j]);
}

#define S3_TRYW(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { s3_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); rc = S3_ECUDA; goto done; } } while (0)

extern "C" int s3_dp_make_windows(s3_index *ix, int mode, const s3_window_params *par, const uint32_t *readLengths, uint64_t numReads,
                                  const uint32_t *readIDs, const uint32_t *positions, const uint32_t *positions2, const uint8_t *strands,
                                  const int32_t *leftScores, const uint32_t *leftStarts, const uint32_t *leftHitLocs, uint64_t numCandidates,
                                  uint32_t *outCandidate, uint32_t *outReadIDs, uint8_t *outStrands, uint8_t *outLeftOrRight,
                                  uint32_t *DNAStarts, uint32_t *DNALengths, uint32_t *clipLtSizes, uint32_t *clipRtSizes,
                                  uint32_t *anchorLeftLocs, uint32_t *anchorRightLocs, int32_t *cutoffThresholds, uint64_t *numWindows)
{
    if (!ix || !par || !readLengths || !numWindows || (numCandidates && (!readIDs || !positions || !outCandidate || !outReadIDs || !outStrands ||
        !outLeftOrRight || !DNAStarts || !DNALengths || !clipLtSizes || !clipRtSizes || !anchorLeftLocs || !anchorRightLocs || !cutoffThresholds))) {
        s3_set_error("s3_dp_make_windows: NULL argument"); return S3_EINVAL;
    }
    if (mode < S3_WIN_SINGLE || mode > S3_WIN_PAIR_RIGHT) { s3_set_error("s3_dp_make_windows: mode %d", mode); return S3_EINVAL; }
    if ((mode == S3_WIN_SINGLE || mode == S3_WIN_HALF) && numCandidates && !strands) { s3_set_error("s3_dp_make_windows: strands needed"); return S3_EINVAL; }
    if (mode == S3_WIN_PAIR_RIGHT && numCandidates && (!positions2 || !leftScores || !leftStarts || !leftHitLocs)) { s3_set_error("s3_dp_make_windows: the left alignments are needed"); return S3_EINVAL; }
    if (numCandidates >= 0x7FFFFFFFull || numReads >= 0xFFFFFFFFull) { s3_set_error("s3_dp_make_windows: batch too large"); return S3_EINVAL; }
    *numWindows = 0;
    if (numCandidates == 0) return S3_OK;
    for (uint64_t c = 0; c < numCandidates; ++c)
        if (readIDs[c] >= numReads || (mode != S3_WIN_SINGLE && (readIDs[c] ^ 1u) >= numReads)) { s3_set_error("s3_dp_make_windows: read id %u out of range", readIDs[c]); return S3_EINVAL; }
    if (cudaSetDevice(ix->device) != cudaSuccess) { s3_set_error("s3_dp_make_windows: cudaSetDevice failed"); return S3_ECUDA; }
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    const uint32_t n = (uint32_t)numCandidates;
    const size_t cap = (mode == S3_WIN_HALF) ? 2 * (size_t)n : n;
    char *d = NULL;
    size_t scanTemp = 0;
    cub::DeviceScan::ExclusiveSum(NULL, scanTemp, (uint32_t *)NULL, (uint32_t *)NULL, (int)(n + 1), st);
    const size_t bytes = (numReads + 7 * (size_t)n + 2 * ((size_t)n + 1) + 9 * cap) * 4 + (size_t)n + 2 * cap + scanTemp + 4096;
    uint32_t total = 0;
    S3_TRYW(cudaMalloc(&d, bytes));
    {
        uint32_t *p = (uint32_t *)d;
        auto take = [&](size_t k) { uint32_t *q = p; p += (k + 63) / 64 * 64; return q; };
        uint32_t *d_len = take(numReads), *d_id = take(n), *d_pos = take(n), *d_pos2 = take(n), *d_ls = take(n), *d_lh = take(n);
        int32_t *d_sc = (int32_t *)take(n);
        uint32_t *d_cnt = take(n + 1), *d_off = take(n + 1);
        S3WinOut o;
        o.candidate = take(cap); o.readIDs = take(cap); o.start = take(cap); o.dnaLen = take(cap); o.clipLt = take(cap); o.clipRt = take(cap);
        o.ancL = take(cap); o.ancR = take(cap); o.cutoff = (int32_t *)take(cap);
        uint8_t *d_str = (uint8_t *)take((n + 3) / 4);
        o.strands = (uint8_t *)take((cap + 3) / 4); o.leftOrRight = (uint8_t *)take((cap + 3) / 4);
        void *d_tmp = take((scanTemp + 3) / 4);
        S3_TRYW(cudaMemcpyAsync(d_len, readLengths, numReads * 4, cudaMemcpyHostToDevice, st));
        S3_TRYW(cudaMemcpyAsync(d_id, readIDs, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        S3_TRYW(cudaMemcpyAsync(d_pos, positions, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        if (strands) S3_TRYW(cudaMemcpyAsync(d_str, strands, n, cudaMemcpyHostToDevice, st));
        if (mode == S3_WIN_PAIR_RIGHT) {
            S3_TRYW(cudaMemcpyAsync(d_pos2, positions2, (size_t)n * 4, cudaMemcpyHostToDevice, st));
            S3_TRYW(cudaMemcpyAsync(d_ls, leftStarts, (size_t)n * 4, cudaMemcpyHostToDevice, st));
            S3_TRYW(cudaMemcpyAsync(d_lh, leftHitLocs, (size_t)n * 4, cudaMemcpyHostToDevice, st));
            S3_TRYW(cudaMemcpyAsync(d_sc, leftScores, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        }
        S3WinParams w;
        w.insertLow = par->insertLow; w.insertHigh = par->insertHigh; w.leftLeg = par->strandLeftLeg; w.rightLeg = par->strandRightLeg;
        w.softClipLeft = par->softClipLeft; w.softClipRight = par->softClipRight; w.cutoff[0] = par->cutoffThreshold[0]; w.cutoff[1] = par->cutoffThreshold[1];
        w.maxDNALength = par->maxDNALength; w.textLength = ix->textLength;
        S3WinIn in = {d_id, d_pos, d_pos2, d_ls, d_lh, d_len, d_str, d_sc};
        const unsigned blocks = (n + 255) / 256;
        S3_TRYW(cudaMemsetAsync(d_cnt + n, 0, 4, st));
        s3_make_windows_kernel<false><<<blocks, 256, 0, st>>>(mode, w, in, n, d_cnt, NULL, o);
        S3_TRYW(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_cnt, d_off, (int)(n + 1), st));
        s3_make_windows_kernel<true><<<blocks, 256, 0, st>>>(mode, w, in, n, NULL, d_off, o);
        S3_LAUNCHED(2);
        S3_TRYW(cudaGetLastError());
        S3_TRYW(cudaMemcpyAsync(&total, d_off + n, 4, cudaMemcpyDeviceToHost, st));
        S3_TRYW(cudaStreamSynchronize(st));
        if (total) {
            S3_TRYW(cudaMemcpyAsync(outCandidate, o.candidate, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaMemcpyAsync(outReadIDs, o.readIDs, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaMemcpyAsync(DNAStarts, o.start, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaMemcpyAsync(DNALengths, o.dnaLen, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaMemcpyAsync(clipLtSizes, o.clipLt, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaMemcpyAsync(clipRtSizes, o.clipRt, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaMemcpyAsync(anchorLeftLocs, o.ancL, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaMemcpyAsync(anchorRightLocs, o.ancR, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaMemcpyAsync(cutoffThresholds, o.cutoff, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaMemcpyAsync(outStrands, o.strands, total, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaMemcpyAsync(outLeftOrRight, o.leftOrRight, total, cudaMemcpyDeviceToHost, st));
            S3_TRYW(cudaStreamSynchronize(st));
        }
        *numWindows = total;
    }
done:
    if (d) cudaFree(d);
    return rc;
}
