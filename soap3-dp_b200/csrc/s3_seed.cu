// s3_seed.cu -- seed hits -> candidate positions for single-end DP seeding, on the device.
//
// Replaces SingleEndSeedingBatch::decodePositions + singleMerge (DV-DPfunctions.cu:1101-1219): the host loops over
// (*bwt->_bwtSaValue)(bwt, k), three host radix sorts and the merge walk.  Here: positions gathered from the suffix
// array in HBM, one stable 64-bit radix sort by (readID, estimated read start) -- the order the reference's three
// passes really leave, its 8-bit strand pass sorting into an array that is freed (DV-DPfunctions.h:90-95,
// .cu:1214-1217) --, one thread per read for the merge walk, a stream compaction.
#include "s3_common.cuh"
#include "../../include/soap3dp_b200.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <stdlib.h>

#define S3_DIVIDE_GAP 50u            // DPS_DIVIDE_GAP, DV-DPfunctions.h:944

__global__ void s3_seed_count_kernel(const uint32_t *__restrict__ saL, const uint32_t *__restrict__ saR, uint64_t n,
                                     uint32_t maxPerRange, unsigned long long *__restrict__ cnt)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n) return;
    unsigned long long c = 0;
    if (g < n && saR[g] >= saL[g]) { c = (unsigned long long)(saR[g] - saL[g]) + 1; if (c > maxPerRange) c = maxPerRange; }
    cnt[g] = c;
}

// one warp per range: estimated read start of every position (DV-DPfunctions.cu:1162-1164), key = readID << 32 | start
__global__ void s3_seed_fill_kernel(const uint32_t *__restrict__ sa, const uint32_t *__restrict__ saL, const uint32_t *__restrict__ saR,
                                    const int32_t *__restrict__ strands, const uint32_t *__restrict__ readIDs,
                                    const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ seedLengths,
                                    const uint32_t *__restrict__ readLengths, uint64_t n, uint32_t maxPerRange,
                                    const unsigned long long *__restrict__ start, unsigned long long *__restrict__ keys,
                                    int32_t *__restrict__ vals)
{
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (g >= n || saR[g] < saL[g]) return;
    unsigned long long c = (unsigned long long)(saR[g] - saL[g]) + 1;
    if (c > maxPerRange) c = maxPerRange;
    const int32_t strand = strands[g];
    const uint32_t off = offsets[g], add = seedLengths[g] + off - readLengths[g];
    const unsigned long long hi = (unsigned long long)readIDs[g] << 32;
    for (unsigned long long k = lane; k < c; k += 32) {
        const uint32_t x = sa[(size_t)saL[g] + k];
        keys[start[g] + k] = hi | (strand == 1 ? x - off : x + add);
        vals[start[g] + k] = strand;
    }
}

// one thread per read (the element where a new readID starts walks its read): the first hit and every hit more than
// DPS_DIVIDE_GAP past the last one kept, in uint arithmetic like the reference (DV-DPfunctions.cu:1121-1134)
__global__ void s3_seed_merge_kernel(const unsigned long long *__restrict__ keys, uint64_t n, uint8_t *__restrict__ keep)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t rid = (uint32_t)(keys[i] >> 32);
    if (i > 0 && (uint32_t)(keys[i - 1] >> 32) == rid) return;
    keep[i] = 1;
    uint32_t prev = (uint32_t)keys[i];
    for (uint64_t k = i + 1; k < n && (uint32_t)(keys[k] >> 32) == rid; ++k) {
        const uint32_t cur = (uint32_t)keys[k];
        if ((uint32_t)(prev + S3_DIVIDE_GAP) < cur) { keep[k] = 1; prev = cur; }
    }
}

__global__ void s3_seed_gather_kernel(const unsigned long long *__restrict__ keys, const int32_t *__restrict__ vals,
                                      const uint32_t *__restrict__ sel, uint64_t m, uint32_t *__restrict__ outReadID,
                                      uint32_t *__restrict__ outPos, int32_t *__restrict__ outStrand)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned long long k = keys[sel[i]];
    outReadID[i] = (uint32_t)(k >> 32); outPos[i] = (uint32_t)k; outStrand[i] = vals[sel[i]];
}

#define S3_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { s3_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); rc = S3_ECUDA; goto done; } } while (0)

extern "C" int s3_seed_candidates(s3_index *ix, const uint32_t *saL, const uint32_t *saR, const int32_t *strands,
                                  const uint32_t *readIDs, const uint32_t *offsets, const uint32_t *seedLengths,
                                  const uint32_t *readLengths, uint64_t numRanges, uint32_t maxPerRange,
                                  uint32_t **candReadIDs, uint32_t **candPositions, int32_t **candStrands, uint64_t *numCandidates)
{
    if (!ix || !candReadIDs || !candPositions || !candStrands || !numCandidates ||
        (numRanges && (!saL || !saR || !strands || !readIDs || !offsets || !seedLengths || !readLengths))) {
        s3_set_error("s3_seed_candidates: NULL argument"); return S3_EINVAL;
    }
    if (!ix->loc.sa) { s3_set_error("s3_seed_candidates: the index was uploaded without its suffix array"); return S3_EINVAL; }
    if (numRanges >= 0x7FFFFFFFull || maxPerRange == 0) { s3_set_error("s3_seed_candidates: numRanges / maxPerRange out of range"); return S3_EINVAL; }
    *candReadIDs = *candPositions = NULL; *candStrands = NULL; *numCandidates = 0;
    if (numRanges == 0) return S3_OK;
    if (cudaSetDevice(ix->device) != cudaSuccess) { s3_set_error("s3_seed_candidates: cudaSetDevice failed"); return S3_ECUDA; }
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    const size_t rB = numRanges * 4;
    char *d_in = NULL, *d_work = NULL;
    void *d_tmp = NULL;
    unsigned long long total = 0;
    uint32_t m = 0;
    uint32_t *h_r = NULL, *h_p = NULL; int32_t *h_s = NULL;
    size_t tmpBytes = 0, t1 = 0, t2 = 0, t3 = 0;
    S3_TRY(cudaMalloc(&d_in, 7 * rB + (numRanges + 1) * 8 + 8));
    {
        uint32_t *d_l = (uint32_t *)d_in, *d_r = d_l + numRanges, *d_rid = d_r + numRanges, *d_off = d_rid + numRanges,
                 *d_sl = d_off + numRanges, *d_rl = d_sl + numRanges;
        int32_t *d_st = (int32_t *)(d_rl + numRanges);
        unsigned long long *d_cnt = (unsigned long long *)(d_st + numRanges + ((numRanges & 1) ? 1 : 0));
        S3_TRY(cudaMemcpyAsync(d_l, saL, rB, cudaMemcpyHostToDevice, st)); S3_TRY(cudaMemcpyAsync(d_r, saR, rB, cudaMemcpyHostToDevice, st));
        S3_TRY(cudaMemcpyAsync(d_rid, readIDs, rB, cudaMemcpyHostToDevice, st)); S3_TRY(cudaMemcpyAsync(d_off, offsets, rB, cudaMemcpyHostToDevice, st));
        S3_TRY(cudaMemcpyAsync(d_sl, seedLengths, rB, cudaMemcpyHostToDevice, st)); S3_TRY(cudaMemcpyAsync(d_rl, readLengths, rB, cudaMemcpyHostToDevice, st));
        S3_TRY(cudaMemcpyAsync(d_st, strands, rB, cudaMemcpyHostToDevice, st));
        s3_seed_count_kernel<<<(unsigned)((numRanges + 256) / 256), 256, 0, st>>>(d_l, d_r, numRanges, maxPerRange, d_cnt);
        S3_LAUNCHED(1);
        cub::DeviceScan::ExclusiveSum(NULL, t1, d_cnt, d_cnt, (int)(numRanges + 1), st);
        S3_TRY(cudaMalloc(&d_tmp, t1));
        S3_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, t1, d_cnt, d_cnt, (int)(numRanges + 1), st));
        S3_TRY(cudaMemcpyAsync(&total, d_cnt + numRanges, 8, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaStreamSynchronize(st));
        cudaFree(d_tmp); d_tmp = NULL;
        if (total == 0) goto done;
        if (total >= 0x7FFFFFFFull) { s3_set_error("s3_seed_candidates: %llu positions in one call", total); rc = S3_EINVAL; goto done; }
        // keys x2, vals x2, keep flags, selected indices, outputs
        const size_t T = (size_t)total;
        S3_TRY(cudaMalloc(&d_work, T * (8 + 8 + 4 + 4 + 4 + 4 + 4 + 4) + T + 1024));
        unsigned long long *k0 = (unsigned long long *)d_work, *k1 = k0 + T;
        int32_t *v0 = (int32_t *)(k1 + T), *v1 = v0 + T;
        uint32_t *sel = (uint32_t *)(v1 + T), *o_r = sel + T, *o_p = o_r + T;
        int32_t *o_s = (int32_t *)(o_p + T);
        uint8_t *keep = (uint8_t *)(o_s + T);
        uint32_t *d_m = (uint32_t *)d_in;                 // the range arrays are done with after the fill: reuse a word for the count
        s3_seed_fill_kernel<<<(unsigned)((numRanges * 32 + 255) / 256), 256, 0, st>>>(ix->loc.sa, d_l, d_r, d_st, d_rid, d_off, d_sl, d_rl,
                                                                                         numRanges, maxPerRange, d_cnt, k0, v0);
        S3_LAUNCHED(1);
        cub::DeviceRadixSort::SortPairs(NULL, t2, k0, k1, v0, v1, (int)T, 0, 64, st);
        cub::DeviceSelect::Flagged(NULL, t3, cub::CountingInputIterator<uint32_t>(0), keep, sel, d_m, (int)T, st);
        tmpBytes = t2 > t3 ? t2 : t3;
        S3_TRY(cudaMalloc(&d_tmp, tmpBytes));
        S3_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, t2, k0, k1, v0, v1, (int)T, 0, 64, st));      // stable: ties stay in arrival order
        S3_TRY(cudaMemsetAsync(keep, 0, T, st));
        s3_seed_merge_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(k1, T, keep);
        S3_LAUNCHED(1);
        S3_TRY(cub::DeviceSelect::Flagged(d_tmp, t3, cub::CountingInputIterator<uint32_t>(0), keep, sel, d_m, (int)T, st));
        S3_TRY(cudaMemcpyAsync(&m, d_m, 4, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaStreamSynchronize(st));
        s3_seed_gather_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(k1, v1, sel, m, o_r, o_p, o_s);
        S3_LAUNCHED(1);
        S3_TRY(cudaGetLastError());
        h_r = (uint32_t *)malloc((size_t)m * 4); h_p = (uint32_t *)malloc((size_t)m * 4); h_s = (int32_t *)malloc((size_t)m * 4);
        if (!h_r || !h_p || !h_s) { s3_set_error("s3_seed_candidates: out of host memory"); rc = S3_ENOMEM; goto done; }
        S3_TRY(cudaMemcpyAsync(h_r, o_r, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaMemcpyAsync(h_p, o_p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaMemcpyAsync(h_s, o_s, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaStreamSynchronize(st));
        *candReadIDs = h_r; *candPositions = h_p; *candStrands = h_s; *numCandidates = m;
        h_r = h_p = NULL; h_s = NULL;
    }
done:
    if (d_in) cudaFree(d_in);
    if (d_work) cudaFree(d_work);
    if (d_tmp) cudaFree(d_tmp);
    free(h_r); free(h_p); free(h_s);
    return rc;
}
