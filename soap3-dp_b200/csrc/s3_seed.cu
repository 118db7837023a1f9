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
#include <chrono>

// S3_STAGE_TIMING=1: wall-clock laps on stderr (tuning aid, like s3_stages.cu)
struct S3Lap {
    bool on; std::chrono::steady_clock::time_point t; const char *what;
    explicit S3Lap(const char *w) : on(getenv("S3_STAGE_TIMING") != NULL), t(std::chrono::steady_clock::now()), what(w) {}
    void lap(const char *step)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "    [%s] %-24s %8.2f ms\n", what, step, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

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

// Device form: every pointer is a device pointer; *d_out is a stream-ordered allocation of 3 * m words (read id | estimated
// start | strand) the caller returns with cudaFreeAsync on the index stream (NULL when m == 0).
int s3_seed_candidates_device(s3_index *ix, const uint32_t *d_l, const uint32_t *d_r, const int32_t *d_st, const uint32_t *d_rid,
                              const uint32_t *d_off, const uint32_t *d_sl, const uint32_t *d_rl, uint64_t numRanges, uint32_t maxPerRange,
                              uint32_t **d_out, uint32_t *numCandidates)
{
    *d_out = NULL; *numCandidates = 0;
    if (numRanges == 0) return S3_OK;
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    unsigned long long *d_cnt = NULL;
    char *d_work = NULL;
    void *d_tmp = NULL;
    size_t t1 = 0, t2 = 0, t3 = 0;
    unsigned long long *h_total = (unsigned long long *)ix->pinnedCount;
    uint32_t *h_m = (uint32_t *)ix->pinnedCount + 4;
    S3_TRY(cudaMallocAsync((void **)&d_cnt, (numRanges + 1) * 8 + 8, st));
    s3_seed_count_kernel<<<(unsigned)((numRanges + 256) / 256), 256, 0, st>>>(d_l, d_r, numRanges, maxPerRange, d_cnt);
    S3_LAUNCHED(1);
    cub::DeviceScan::ExclusiveSum(NULL, t1, d_cnt, d_cnt, (int)(numRanges + 1), st);
    S3_TRY(cudaMallocAsync(&d_tmp, t1 + 16, st));
    S3_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, t1, d_cnt, d_cnt, (int)(numRanges + 1), st));
    S3_TRY(cudaMemcpyAsync(h_total, d_cnt + numRanges, 8, cudaMemcpyDeviceToHost, st));
    S3_TRY(cudaStreamSynchronize(st));
    cudaFreeAsync(d_tmp, st); d_tmp = NULL;
    {
        const unsigned long long total = *h_total;
        if (total == 0) goto done;
        if (total >= 0x7FFFFFFFull) { s3_set_error("s3_seed_candidates: %llu positions in one call", total); rc = S3_EINVAL; goto done; }
        // keys x2, vals x2, selected indices, keep flags, the count
        const size_t T = (size_t)total;
        S3_TRY(cudaMallocAsync((void **)&d_work, T * (8 + 8 + 4 + 4 + 4) + T + 1024, st));
        unsigned long long *k0 = (unsigned long long *)d_work, *k1 = k0 + T;
        int32_t *v0 = (int32_t *)(k1 + T), *v1 = v0 + T;
        uint32_t *sel = (uint32_t *)(v1 + T);
        uint8_t *keep = (uint8_t *)(sel + T);
        uint32_t *d_m = (uint32_t *)(((uintptr_t)(keep + T) + 15) & ~(uintptr_t)15);
        s3_seed_fill_kernel<<<(unsigned)((numRanges * 32 + 255) / 256), 256, 0, st>>>(ix->loc.sa, d_l, d_r, d_st, d_rid, d_off, d_sl, d_rl,
                                                                                         numRanges, maxPerRange, d_cnt, k0, v0);
        S3_LAUNCHED(1);
        cub::DeviceRadixSort::SortPairs(NULL, t2, k0, k1, v0, v1, (int)T, 0, 64, st);
        cub::DeviceSelect::Flagged(NULL, t3, cub::CountingInputIterator<uint32_t>(0), keep, sel, d_m, (int)T, st);
        S3_TRY(cudaMallocAsync(&d_tmp, (t2 > t3 ? t2 : t3) + 16, st));
        S3_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, t2, k0, k1, v0, v1, (int)T, 0, 64, st));      // stable: ties stay in arrival order
        S3_TRY(cudaMemsetAsync(keep, 0, T, st));
        s3_seed_merge_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(k1, T, keep);
        S3_LAUNCHED(1);
        S3_TRY(cub::DeviceSelect::Flagged(d_tmp, t3, cub::CountingInputIterator<uint32_t>(0), keep, sel, d_m, (int)T, st));
        S3_TRY(cudaMemcpyAsync(h_m, d_m, 4, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaStreamSynchronize(st));
        const uint32_t m = *h_m;
        if (m) {
            S3_TRY(cudaMallocAsync((void **)d_out, (size_t)m * 12, st));
            s3_seed_gather_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(k1, v1, sel, m, *d_out, *d_out + m, (int32_t *)(*d_out + 2 * (size_t)m));
            S3_LAUNCHED(1);
            S3_TRY(cudaGetLastError());
        }
        *numCandidates = m;
    }
done:
    if (d_cnt) cudaFreeAsync(d_cnt, st);
    if (d_work) cudaFreeAsync(d_work, st);
    if (d_tmp) cudaFreeAsync(d_tmp, st);
    if (rc && *d_out) { cudaFreeAsync(*d_out, st); *d_out = NULL; }
    return rc;
}

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
    // the suffix array has textLength + 1 rows: a stale or overflow-marker answer word passed as a range must not become a read past it
    for (uint64_t g = 0; g < numRanges; ++g)
        if (saR[g] >= saL[g] && saR[g] > ix->textLength) { s3_set_error("s3_seed_candidates: range %llu (%u..%u) lies outside the suffix array", (unsigned long long)g, saL[g], saR[g]); return S3_EINVAL; }
    *candReadIDs = *candPositions = NULL; *candStrands = NULL; *numCandidates = 0;
    if (numRanges == 0) return S3_OK;
    if (cudaSetDevice(ix->device) != cudaSuccess) { s3_set_error("s3_seed_candidates: cudaSetDevice failed"); return S3_ECUDA; }
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    const size_t rB = numRanges * 4;
    uint32_t *d_in = NULL, *d_out = NULL;
    uint32_t m = 0;
    uint32_t *h_r = NULL, *h_p = NULL; int32_t *h_s = NULL;
    S3_TRY(cudaMallocAsync((void **)&d_in, 7 * rB + 64, st));
    {
        const uint32_t *src[7] = {saL, saR, (const uint32_t *)strands, readIDs, offsets, seedLengths, readLengths};
        for (int a = 0; a < 7; ++a) S3_TRY(cudaMemcpyAsync(d_in + a * numRanges, src[a], rB, cudaMemcpyHostToDevice, st));
        if ((rc = s3_seed_candidates_device(ix, d_in, d_in + numRanges, (const int32_t *)(d_in + 2 * numRanges), d_in + 3 * numRanges, d_in + 4 * numRanges,
                                            d_in + 5 * numRanges, d_in + 6 * numRanges, numRanges, maxPerRange, &d_out, &m))) goto done;
        if (m) {
            h_r = (uint32_t *)malloc((size_t)m * 4); h_p = (uint32_t *)malloc((size_t)m * 4); h_s = (int32_t *)malloc((size_t)m * 4);
            if (!h_r || !h_p || !h_s) { s3_set_error("s3_seed_candidates: out of host memory"); rc = S3_ENOMEM; goto done; }
            S3_TRY(cudaMemcpyAsync(h_r, d_out, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
            S3_TRY(cudaMemcpyAsync(h_p, d_out + m, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
            S3_TRY(cudaMemcpyAsync(h_s, d_out + 2 * (size_t)m, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
            S3_TRY(cudaStreamSynchronize(st));
            *candReadIDs = h_r; *candPositions = h_p; *candStrands = h_s; *numCandidates = m;
            h_r = h_p = NULL; h_s = NULL;
        }
    }
done:
    if (d_in) cudaFreeAsync(d_in, st);
    if (d_out) cudaFreeAsync(d_out, st);
    free(h_r); free(h_p); free(h_s);
    return rc;
}

// =====================================================================================================================
// Paired-end seeding: seed hits of both ends -> candidate (left start, right start) pairs.
// Replaces PairEndSeedingBatch::decodePositions (both ends), findRevStart, the two pairEndMerge calls and the final sort
// of decodeMergePositions (DV-DPfunctions.cu:2626-2653,2780-2999).  Per end: key = (strandIndex << 31 | pair id) << 32
// | estimated start, two guards, one 64-bit sort (= the reference's sort by pos, then by strand_readID).  Per call:
// one thread per left group that has a partner group thins it IN PLACE (the reference does, and when both legs use the
// same strand its second call joins against what the first left behind) and then walks the two-pointer join, once to
// count and once to write; the calls' candidates, first call first, get one stable sort by readIDLeft.
// =====================================================================================================================
typedef unsigned long long s3_u64;

__global__ void s3_pair_fill_kernel(const uint32_t *__restrict__ sa, const uint32_t *__restrict__ saL, const uint32_t *__restrict__ saR,
                                    const int32_t *__restrict__ strands, const uint32_t *__restrict__ readIDs,
                                    const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ seedLengths,
                                    const uint32_t *__restrict__ readLengths, uint64_t n, uint32_t maxPerRange,
                                    const s3_u64 *__restrict__ start, s3_u64 total, s3_u64 *__restrict__ keys)
{
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (g == n && lane < 2) {                                           // array guards (DV-DPfunctions.cu:2961-2962)
        keys[total + lane] = ((s3_u64)(0x7FFFFFFFu | (lane << 31)) << 32) | 0xFFFFFFFFull;
        return;
    }
    if (g >= n || saR[g] < saL[g]) return;
    s3_u64 c = (s3_u64)(saR[g] - saL[g]) + 1;
    if (c > maxPerRange) c = maxPerRange;
    const uint32_t si = (uint32_t)strands[g] - 1u, off = offsets[g], add = seedLengths[g] + off - readLengths[g];
    const s3_u64 hi = (s3_u64)(readIDs[g] | (si << 31)) << 32;
    for (s3_u64 k = lane; k < c; k += 32) {
        const uint32_t x = sa[(size_t)saL[g] + k];
        keys[start[g] + k] = hi | (si == 0 ? x - off : x + add);
    }
}

__device__ __forceinline__ uint32_t s3_pair_lower(const s3_u64 *k, uint32_t lo, uint32_t hi, uint32_t key32)
{
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if ((uint32_t)(k[mid] >> 32) < key32) lo = mid + 1; else hi = mid; }
    return lo;
}

// left groups with a partner: thinned in place (MC_Compress, DV-DPfunctions.cu:2826-2838); their new end and the start
// of the partner group are noted at the group's first element
__global__ void s3_pair_thin_kernel(s3_u64 *__restrict__ L, uint32_t lo, uint32_t hi, const s3_u64 *__restrict__ R, uint32_t rlo, uint32_t rhi,
                                    uint32_t rightStrandBit, uint32_t *__restrict__ groupEnd, uint32_t *__restrict__ rightStart)
{
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    groupEnd[i] = 0;
    const uint32_t key32 = (uint32_t)(L[i] >> 32), id = key32 & 0x7FFFFFFFu;
    if (id == 0x7FFFFFFFu || (i > lo && (uint32_t)(L[i - 1] >> 32) == key32)) return;
    const uint32_t rkey = id | rightStrandBit, j = s3_pair_lower(R, rlo, rhi, rkey);
    if (j >= rhi || (uint32_t)(R[j] >> 32) != rkey) return;
    uint32_t c = i, prev = (uint32_t)L[i];
    for (uint32_t p = i + 1; p < hi && (uint32_t)(L[p] >> 32) == key32; ++p) {
        const uint32_t cur = (uint32_t)L[p];
        if ((uint32_t)(prev + S3_DIVIDE_GAP) < cur) { L[++c] = L[p]; prev = cur; }
    }
    groupEnd[i] = c + 1;
    rightStart[i] = j;
}

// the two-pointer walk of pairEndMerge (DV-DPfunctions.cu:2839-2877); FILL = false counts, true writes
template <bool FILL>
__global__ void s3_pair_join_kernel(const s3_u64 *__restrict__ L, uint32_t lo, uint32_t hi, const s3_u64 *__restrict__ R,
                                    const uint32_t *__restrict__ groupEnd, const uint32_t *__restrict__ rightStart,
                                    const uint32_t *__restrict__ lengthsByReadID, int insertLow, int insertHigh, uint32_t leftReadOrMate,
                                    s3_u64 *__restrict__ counts, uint32_t *__restrict__ outID, uint32_t *__restrict__ outL, uint32_t *__restrict__ outR)
{
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > hi) return;
    if (i == hi) { if (!FILL) counts[i - lo] = 0; return; }            // one extra slot: the scan leaves the total there
    const uint32_t le = groupEnd[i];
    if (le == 0) { if (!FILL) counts[i - lo] = 0; return; }
    const uint32_t id = (uint32_t)(L[i] >> 32) & 0x7FFFFFFFu;
    uint32_t rp = rightStart[i];
    const uint32_t rkey = (uint32_t)(R[rp] >> 32);
    uint32_t re = rp;
    while ((uint32_t)(R[re] >> 32) == rkey) ++re;                       // ends at the next group or a guard
    const int readLength = (int)lengthsByReadID[id];
    const int margin = readLength > 100 ? (readLength >> 2) : 25;       // DP2_MARGIN, DV-DPfunctions.cu:2549
    int lengthLow = insertLow - readLength - margin;
    if (lengthLow < 0) lengthLow = 0;
    const int lengthHigh = insertHigh - readLength + margin;
    uint32_t lp = i, lloc = (uint32_t)L[lp], rloc = (uint32_t)R[rp];
    s3_u64 n = 0;
    const s3_u64 base = FILL ? counts[i - lo] : 0;
    while (lp < le && rp < re) {
        if ((uint32_t)(lloc + (uint32_t)lengthLow) > rloc) { ++rp; rloc = (uint32_t)R[rp]; }
        else if ((uint32_t)(lloc + (uint32_t)lengthHigh) < rloc) { ++lp; lloc = (uint32_t)L[lp]; }
        else {
            if (FILL) { outID[base + n] = id + leftReadOrMate; outL[base + n] = lloc; outR[base + n] = rloc; }
            ++n;
            ++lp; lloc = (uint32_t)L[lp];
        }
    }
    if (!FILL) counts[i - lo] = n;
}

__global__ void s3_pair_gather_kernel(const uint32_t *__restrict__ order, uint64_t m, const uint32_t *__restrict__ inL, const uint32_t *__restrict__ inR,
                                      uint32_t *__restrict__ outL, uint32_t *__restrict__ outR)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    outL[i] = inL[order[i]]; outR[i] = inR[order[i]];
}

__global__ void s3_iota_kernel(uint32_t *a, uint64_t m)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) a[i] = (uint32_t)i;
}

struct S3PairSide { s3_u64 *keys; uint32_t len, rev; };

// one end's hits: gathered, keyed, sorted; rev = first element of the reverse-strand part (findRevStart)
static int s3_pair_side(s3_index *ix, const uint32_t *const in[7], uint64_t n, uint32_t maxPerRange, int onDevice, S3PairSide *out)
{
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    const size_t rB = n * 4;
    char *d_in = NULL;
    void *d_tmp = NULL;
    s3_u64 *k0 = NULL, *k1 = NULL;
    size_t t1 = 0, t2 = 0;
    s3_u64 total = 0;
    out->keys = NULL; out->len = out->rev = 0;
    S3_TRY(cudaMallocAsync(&d_in, 7 * rB + (n + 1) * 8 + 16, st));
    {
        uint32_t *d[7];
        for (int a = 0; a < 7; ++a) {
            if (onDevice) { d[a] = const_cast<uint32_t *>(in[a]); continue; }
            d[a] = (uint32_t *)d_in + a * n;
            if (n) S3_TRY(cudaMemcpyAsync(d[a], in[a], rB, cudaMemcpyHostToDevice, st));
        }
        s3_u64 *d_cnt = (s3_u64 *)(((uintptr_t)((uint32_t *)d_in + 7 * n) + 7) & ~(uintptr_t)7);
        s3_seed_count_kernel<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(d[0], d[1], n, maxPerRange, d_cnt);
        S3_LAUNCHED(1);
        cub::DeviceScan::ExclusiveSum(NULL, t1, d_cnt, d_cnt, (int)(n + 1), st);
        S3_TRY(cudaMallocAsync(&d_tmp, t1, st));
        S3_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, t1, d_cnt, d_cnt, (int)(n + 1), st));
        S3_TRY(cudaMemcpyAsync(&total, d_cnt + n, 8, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaStreamSynchronize(st));
        cudaFreeAsync(d_tmp, st); d_tmp = NULL;
        if (total + 2 >= 0x7FFFFFFFull) { s3_set_error("s3_seed_pair_candidates: %llu positions in one call", total); rc = S3_EINVAL; goto done; }
        const size_t T = (size_t)total + 2;
        S3_TRY(cudaMallocAsync(&k0, T * 8, st)); S3_TRY(cudaMallocAsync(&k1, T * 8, st));
        s3_pair_fill_kernel<<<(unsigned)(((n + 1) * 32 + 255) / 256), 256, 0, st>>>(ix->loc.sa, d[0], d[1], (const int32_t *)d[2], d[3], d[4], d[5], d[6],
                                                                                       n, maxPerRange, d_cnt, total, k0);
        S3_LAUNCHED(1);
        cub::DeviceRadixSort::SortKeys(NULL, t2, k0, k1, (int)T, 0, 64, st);
        S3_TRY(cudaMallocAsync(&d_tmp, t2, st));
        S3_TRY(cub::DeviceRadixSort::SortKeys(d_tmp, t2, k0, k1, (int)T, 0, 64, st));
        S3_TRY(cudaStreamSynchronize(st));
        out->keys = k1; k1 = NULL; out->len = (uint32_t)T;
    }
done:
    if (d_in) cudaFreeAsync(d_in, st);
    if (d_tmp) cudaFreeAsync(d_tmp, st);
    if (k0) cudaFreeAsync(k0, st);
    if (k1) cudaFreeAsync(k1, st);
    return rc;
}

__global__ void s3_pair_revstart_kernel(const s3_u64 *k, uint32_t len, uint32_t *out)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *out = s3_pair_lower(k, 0, len, 0x80000000u);
}

int s3_seed_pair_candidates_any(s3_index *ix, const uint32_t *const in0[7], uint64_t n0, const uint32_t *const in1[7], uint64_t n1, int onDevice,
                                uint32_t maxPerRange, const uint32_t *lengthsByReadID, uint64_t numReadIDs,
                                int insertLow, int insertHigh, int peStrandLeftLeg, int peStrandRightLeg,
                                uint32_t **candReadIDLeft, uint32_t **candPosLeft, uint32_t **candPosRight, uint64_t *numCandidates);

extern "C" int s3_seed_pair_candidates(s3_index *ix,
                                       const uint32_t *saL0, const uint32_t *saR0, const int32_t *strands0, const uint32_t *readIDs0,
                                       const uint32_t *offsets0, const uint32_t *seedLengths0, const uint32_t *readLengths0, uint64_t n0,
                                       const uint32_t *saL1, const uint32_t *saR1, const int32_t *strands1, const uint32_t *readIDs1,
                                       const uint32_t *offsets1, const uint32_t *seedLengths1, const uint32_t *readLengths1, uint64_t n1,
                                       uint32_t maxPerRange, const uint32_t *lengthsByReadID, uint64_t numReadIDs,
                                       int insertLow, int insertHigh, int peStrandLeftLeg, int peStrandRightLeg,
                                       uint32_t **candReadIDLeft, uint32_t **candPosLeft, uint32_t **candPosRight, uint64_t *numCandidates)
{
    if (!ix || !candReadIDLeft || !candPosLeft || !candPosRight || !numCandidates || !lengthsByReadID ||
        (n0 && (!saL0 || !saR0 || !strands0 || !readIDs0 || !offsets0 || !seedLengths0 || !readLengths0)) ||
        (n1 && (!saL1 || !saR1 || !strands1 || !readIDs1 || !offsets1 || !seedLengths1 || !readLengths1))) {
        s3_set_error("s3_seed_pair_candidates: NULL argument"); return S3_EINVAL;
    }
    if (!ix->loc.sa) { s3_set_error("s3_seed_pair_candidates: the index was uploaded without its suffix array"); return S3_EINVAL; }
    if (peStrandLeftLeg < 1 || peStrandLeftLeg > 2 || peStrandRightLeg < 1 || peStrandRightLeg > 2 || maxPerRange == 0 ||
        n0 >= 0x7FFFFFF0ull || n1 >= 0x7FFFFFF0ull) { s3_set_error("s3_seed_pair_candidates: argument out of range"); return S3_EINVAL; }
    for (uint64_t g = 0; g < n0; ++g) if (readIDs0[g] >= numReadIDs) { s3_set_error("s3_seed_pair_candidates: read id %u out of range", readIDs0[g]); return S3_EINVAL; }
    for (uint64_t g = 0; g < n1; ++g) if (readIDs1[g] >= numReadIDs) { s3_set_error("s3_seed_pair_candidates: read id %u out of range", readIDs1[g]); return S3_EINVAL; }
    for (uint64_t g = 0; g < n0; ++g) if (saR0[g] >= saL0[g] && saR0[g] > ix->textLength) { s3_set_error("s3_seed_pair_candidates: range %u..%u lies outside the suffix array", saL0[g], saR0[g]); return S3_EINVAL; }
    for (uint64_t g = 0; g < n1; ++g) if (saR1[g] >= saL1[g] && saR1[g] > ix->textLength) { s3_set_error("s3_seed_pair_candidates: range %u..%u lies outside the suffix array", saL1[g], saR1[g]); return S3_EINVAL; }
    const uint32_t *const in0[7] = {saL0, saR0, (const uint32_t *)strands0, readIDs0, offsets0, seedLengths0, readLengths0};
    const uint32_t *const in1[7] = {saL1, saR1, (const uint32_t *)strands1, readIDs1, offsets1, seedLengths1, readLengths1};
    return s3_seed_pair_candidates_any(ix, in0, n0, in1, n1, 0, maxPerRange, lengthsByReadID, numReadIDs, insertLow, insertHigh, peStrandLeftLeg, peStrandRightLeg,
                                       candReadIDLeft, candPosLeft, candPosRight, numCandidates);
}

// onDevice: the range arrays and lengthsByReadID are device arrays, and the three output pointers receive device arrays: the first
// one (*candReadIDLeft) is the base of one allocation the caller returns with cudaFreeAsync on the index stream.
int s3_seed_pair_candidates_any(s3_index *ix, const uint32_t *const in0[7], uint64_t n0, const uint32_t *const in1[7], uint64_t n1, int onDevice,
                                uint32_t maxPerRange, const uint32_t *lengthsByReadID, uint64_t numReadIDs,
                                int insertLow, int insertHigh, int peStrandLeftLeg, int peStrandRightLeg,
                                uint32_t **candReadIDLeft, uint32_t **candPosLeft, uint32_t **candPosRight, uint64_t *numCandidates)
{
    *candReadIDLeft = *candPosLeft = *candPosRight = NULL; *numCandidates = 0;
    if (cudaSetDevice(ix->device) != cudaSuccess) { s3_set_error("s3_seed_pair_candidates: cudaSetDevice failed"); return S3_ECUDA; }
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    S3PairSide side[2] = {{NULL, 0, 0}, {NULL, 0, 0}};
    uint32_t *d_len = NULL, *d_aux = NULL, *d_out = NULL, *d_sorted = NULL;
    s3_u64 *d_counts = NULL;
    void *d_tmp = NULL;
    uint32_t *h[3] = {NULL, NULL, NULL};
    s3_u64 tot[2] = {0, 0};
    S3Lap clk("s3_seed_pair_candidates");
    if ((rc = s3_pair_side(ix, in0, n0, maxPerRange, onDevice, &side[0]))) goto done;
    if ((rc = s3_pair_side(ix, in1, n1, maxPerRange, onDevice, &side[1]))) goto done;
    clk.lap("two sides: gather + sort");
    {
        S3_TRY(cudaMallocAsync(&d_len, numReadIDs * 4 + 16, st));
        S3_TRY(cudaMemcpyAsync(d_len, lengthsByReadID, numReadIDs * 4, onDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        uint32_t *d_rev = d_len + numReadIDs;
        for (int s = 0; s < 2; ++s) s3_pair_revstart_kernel<<<1, 1, 0, st>>>(side[s].keys, side[s].len, d_rev + s);
        uint32_t rev[2];
        S3_TRY(cudaMemcpyAsync(rev, d_rev, 8, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaStreamSynchronize(st));
        side[0].rev = rev[0]; side[1].rev = rev[1];
        const uint32_t maxLen = side[0].len > side[1].len ? side[0].len : side[1].len;
        S3_TRY(cudaMallocAsync(&d_counts, ((size_t)maxLen + 1) * 8, st));                   // only to size the scan's temporary storage
        size_t tScan = 0;
        cub::DeviceScan::ExclusiveSum(NULL, tScan, d_counts, d_counts, (int)(maxLen + 1), st);
        S3_TRY(cudaMallocAsync(&d_tmp, tScan, st));
        // segment of an end's array that holds strand index si: [0, rev) or [rev, len)
        auto seg = [&](int s, int si, uint32_t &lo, uint32_t &hi) { lo = si ? side[s].rev : 0u; hi = si ? side[s].len : side[s].rev; };
        // the two calls (DV-DPfunctions.cu:2987-2988): read left / mate right, then mate left / read right.  Counting of
        // call 1 must see the arrays as call 0's thinning left them, so: thin + count, call after call, then fill.
        uint32_t lo[2], hi[2], rlo[2], rhi[2];
        uint32_t *groupEnd[2], *rightStart[2];
        s3_u64 *counts[2] = {NULL, NULL};
        uint32_t *ge_all = NULL;
        S3_TRY(cudaMallocAsync(&ge_all, ((size_t)side[0].len + side[1].len) * 8 + 16, st));
        groupEnd[0] = ge_all; rightStart[0] = ge_all + side[0].len;
        groupEnd[1] = ge_all + 2 * (size_t)side[0].len; rightStart[1] = groupEnd[1] + side[1].len;
        s3_u64 *cnt_all = NULL;
        cudaError_t e2 = cudaMallocAsync((void **)&cnt_all, ((size_t)side[0].len + side[1].len + 2) * 8, st);
        if (e2 != cudaSuccess) { cudaFreeAsync(ge_all, st); s3_set_error("s3_seed_pair_candidates: %s", cudaGetErrorString(e2)); rc = S3_ECUDA; goto done; }
        counts[0] = cnt_all; counts[1] = cnt_all + side[0].len + 1;
        bool ok = true;
        for (int call = 0; call < 2 && ok; ++call) {
            const int ls = call, rs = 1 - call;                                    // which end is on the left
            seg(ls, peStrandLeftLeg - 1, lo[call], hi[call]);
            seg(rs, peStrandRightLeg - 1, rlo[call], rhi[call]);
            const uint32_t nL = hi[call] - lo[call];
            if (nL) {
                s3_pair_thin_kernel<<<(nL + 255) / 256, 256, 0, st>>>(side[ls].keys, lo[call], hi[call], side[rs].keys, rlo[call], rhi[call],
                                                                      (uint32_t)(peStrandRightLeg - 1) << 31, groupEnd[call] - 0, rightStart[call]);
                S3_LAUNCHED(1);
            }
            // the fill pass of call 0 must read call 0's LEFT array as thinned and its RIGHT array untouched: write call 0's
            // candidates before call 1 thins that right array
            s3_pair_join_kernel<false><<<(nL + 1 + 255) / 256, 256, 0, st>>>(side[ls].keys, lo[call], hi[call], side[rs].keys, groupEnd[call], rightStart[call],
                                                                            d_len, insertLow, insertHigh, (uint32_t)call, counts[call], NULL, NULL, NULL);
            S3_LAUNCHED(1);
            if (cub::DeviceScan::ExclusiveSum(d_tmp, tScan, counts[call], counts[call], (int)(nL + 1), st) != cudaSuccess) { ok = false; break; }
            if (cudaMemcpyAsync(&tot[call], counts[call] + nL, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) { ok = false; break; }
            if (cudaStreamSynchronize(st) != cudaSuccess) { ok = false; break; }
            if (call == 0) {
                // room for call 0's candidates now; call 1's are appended after its own count
                if (tot[0] && cudaMallocAsync((void **)&d_out, (size_t)tot[0] * 12, st) != cudaSuccess) { ok = false; break; }
                if (tot[0]) {
                    s3_pair_join_kernel<true><<<(nL + 1 + 255) / 256, 256, 0, st>>>(side[ls].keys, lo[0], hi[0], side[rs].keys, groupEnd[0], rightStart[0], d_len,
                                                                                   insertLow, insertHigh, 0u, counts[0], d_out, d_out + tot[0], d_out + 2 * tot[0]);
                    S3_LAUNCHED(1);
                }
            }
        }
        clk.lap("thin + join, two calls");
        uint32_t *d_all = NULL;
        const s3_u64 m = tot[0] + tot[1];
        if (ok && m >= 0x7FFFFFFFull) { s3_set_error("s3_seed_pair_candidates: %llu candidates in one call", m); rc = S3_EINVAL; ok = false; }
        if (ok && m) {
            // all candidates in one place: ids | left | right, call 0's first
            ok = cudaMallocAsync((void **)&d_all, (size_t)m * 12, st) == cudaSuccess;
            if (ok && tot[0]) {
                ok = cudaMemcpyAsync(d_all, d_out, (size_t)tot[0] * 4, cudaMemcpyDeviceToDevice, st) == cudaSuccess &&
                     cudaMemcpyAsync(d_all + m, d_out + tot[0], (size_t)tot[0] * 4, cudaMemcpyDeviceToDevice, st) == cudaSuccess &&
                     cudaMemcpyAsync(d_all + 2 * m, d_out + 2 * tot[0], (size_t)tot[0] * 4, cudaMemcpyDeviceToDevice, st) == cudaSuccess;
            }
            if (ok && tot[1]) {
                const uint32_t nL = hi[1] - lo[1];
                s3_pair_join_kernel<true><<<(nL + 1 + 255) / 256, 256, 0, st>>>(side[1].keys, lo[1], hi[1], side[0].keys, groupEnd[1], rightStart[1], d_len,
                                                                               insertLow, insertHigh, 1u, counts[1], d_all + tot[0], d_all + m + tot[0], d_all + 2 * m + tot[0]);
                S3_LAUNCHED(1);
            }
            // stable sort by readIDLeft (MC_RadixSort_32_16 on candArr, DV-DPfunctions.cu:2997)
            size_t tSort = 0;
            void *d_t2 = NULL;
            if (ok) ok = cudaMallocAsync((void **)&d_sorted, (size_t)m * 20, st) == cudaSuccess;      // ids out | order in | order out | left out | right out
            if (ok) {
                uint32_t *idsOut = d_sorted, *ordIn = d_sorted + m, *ordOut = d_sorted + 2 * m, *lOut = d_sorted + 3 * m, *rOut = d_sorted + 4 * m;
                s3_iota_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(ordIn, m);
                cub::DeviceRadixSort::SortPairs(NULL, tSort, d_all, idsOut, ordIn, ordOut, (int)m, 0, 32, st);
                ok = cudaMallocAsync((void **)&d_t2, tSort, st) == cudaSuccess &&
                     cub::DeviceRadixSort::SortPairs(d_t2, tSort, d_all, idsOut, ordIn, ordOut, (int)m, 0, 32, st) == cudaSuccess;
                if (ok) {
                    s3_pair_gather_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(ordOut, m, d_all + m, d_all + 2 * m, lOut, rOut);
                    S3_LAUNCHED(2);
                    if (onDevice) {
                        ok = cudaStreamSynchronize(st) == cudaSuccess;
                        if (ok) { *candReadIDLeft = idsOut; *candPosLeft = lOut; *candPosRight = rOut; d_sorted = NULL; }     // the caller's now
                    }
                    for (int a = 0; a < 3 && ok && !onDevice; ++a) { h[a] = (uint32_t *)malloc((size_t)m * 4); ok = h[a] != NULL; }
                    if (onDevice) { /* nothing to copy back */ } else
                    if (ok) ok = cudaMemcpyAsync(h[0], idsOut, (size_t)m * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                                 cudaMemcpyAsync(h[1], lOut, (size_t)m * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                                 cudaMemcpyAsync(h[2], rOut, (size_t)m * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                                 cudaStreamSynchronize(st) == cudaSuccess;
                }
                if (d_t2) cudaFreeAsync(d_t2, st);
            }
        }
        clk.lap("fill + final sort");
        if (ok && cudaGetLastError() != cudaSuccess) ok = false;
        cudaFreeAsync(ge_all, st); cudaFreeAsync(cnt_all, st);
        if (d_all) cudaFreeAsync(d_all, st);
        if (!ok) { if (rc == S3_OK) { s3_set_error("s3_seed_pair_candidates: a CUDA call failed: %s", cudaGetErrorString(cudaGetLastError())); rc = S3_ECUDA; } goto done; }
        if (m && !onDevice) { *candReadIDLeft = h[0]; *candPosLeft = h[1]; *candPosRight = h[2]; h[0] = h[1] = h[2] = NULL; }
        *numCandidates = m;
    }
done:
    clk.lap("frees");
    for (int s = 0; s < 2; ++s) if (side[s].keys) cudaFreeAsync(side[s].keys, st);
    if (d_len) cudaFreeAsync(d_len, st);
    if (d_aux) cudaFreeAsync(d_aux, st);
    if (d_counts) cudaFreeAsync(d_counts, st);
    if (d_out) cudaFreeAsync(d_out, st);
    if (d_sorted) cudaFreeAsync(d_sorted, st);
    if (d_tmp) cudaFreeAsync(d_tmp, st);
    for (int a = 0; a < 3; ++a) free(h[a]);
    return rc;
}

// =====================================================================================================================
// The seeding driver.  Replaces single_1_mismatch_alignment2 (alignment.cu:1839-1893) with the per-seed bookkeeping of
// hostKernelSingle (CPUfunctions.cpp:2700-2870) for a batch of seeds: every seed is searched exactly
// (single_all_valid_seed_alignment with 0 mismatches); the seeds without any alignment -- noAlignment, every case's status
// word "no hit" (alignment.cu:1643-1661) -- are searched again with 1 mismatch (both cases of that scheme, the "at most"
// variant like every caller).  A seed keeps its SA ranges when they hold at most maxHitNum occurrences; with more it keeps
// none and is marked too-many (status 4, ProceedDPForTooManyHits = 0 as shipped).  The reference reaches a seed's complete
// range list through round 1, round 2 and its CPU search; s3_search enumerates it without caps in the same order (cases
// ascending, enumeration order inside a case).
// =====================================================================================================================
extern "C" void s3_seed_search_result_free(s3_seed_search_result *r)
{
    if (!r) return;
    free(r->offsets); free(r->saL); free(r->saR); free(r->strand); free(r->status);
    memset(r, 0, sizeof *r);
}

extern "C" int s3_seed_search(s3_index *ix, const uint32_t *seeds, const uint32_t *seedLengths, uint64_t numSeeds, uint32_t wordPerSeed,
                              const uint32_t *maxHitNum, s3_seed_search_result *out)
{
    if (!out) { s3_set_error("s3_seed_search: NULL result"); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    if (!ix || (numSeeds && (!seeds || !seedLengths || !maxHitNum))) { s3_set_error("s3_seed_search: NULL argument"); return S3_EINVAL; }
    out->numSeeds = numSeeds;
    out->offsets = (uint64_t *)calloc(numSeeds + 1, 8);
    out->status = (uint8_t *)calloc(numSeeds ? numSeeds : 1, 1);
    if (!out->offsets || !out->status) { s3_seed_search_result_free(out); s3_set_error("s3_seed_search: out of host memory"); return S3_ENOMEM; }
    if (numSeeds == 0) return S3_OK;
    int rc;
    S3Lap clk("s3_seed_search");
    s3_search_result exact, one;
    memset(&one, 0, sizeof one);
    if ((rc = s3_search(ix, seeds, seedLengths, numSeeds, wordPerSeed, 0, 0, &exact))) { s3_seed_search_result_free(out); return rc; }
    clk.lap("exact search");
    // the seeds without an exact alignment, packed like packReads2 (CPUfunctions.cpp:303-338)
    uint64_t n1 = 0;
    for (uint64_t s = 0; s < numSeeds; ++s) n1 += exact.offsets[s + 1] == exact.offsets[s];
    uint32_t *ids1 = NULL, *q1 = NULL, *l1 = NULL;
    if (n1) {
        const uint64_t up1 = (n1 + 31) / 32 * 32;
        ids1 = (uint32_t *)malloc(n1 * 4); q1 = (uint32_t *)calloc(up1 * wordPerSeed, 4); l1 = (uint32_t *)calloc(up1, 4);
        if (!ids1 || !q1 || !l1) { rc = S3_ENOMEM; s3_set_error("s3_seed_search: out of host memory"); }
        else {
            uint64_t k = 0;
            for (uint64_t s = 0; s < numSeeds; ++s) {
                if (exact.offsets[s + 1] != exact.offsets[s]) continue;
                const uint32_t *src = seeds + (s / 32) * 32 * wordPerSeed + s % 32;
                uint32_t *dst = q1 + (k / 32) * 32 * wordPerSeed + k % 32;
                for (uint32_t w = 0; w < wordPerSeed; ++w) dst[(size_t)w * 32] = src[(size_t)w * 32];
                l1[k] = seedLengths[s]; ids1[k] = (uint32_t)s; ++k;
            }
            clk.lap("pack seeds without a hit");
            rc = s3_search(ix, q1, l1, n1, wordPerSeed, 1, 0, &one);
            clk.lap("1-mismatch search");
        }
    }
    if (rc == S3_OK) {
        // count, then fill: a seed's ranges come from the pass that found it
        uint64_t total = 0, k = 0;
        for (uint64_t s = 0; s < numSeeds; ++s) {
            const bool second = exact.offsets[s + 1] == exact.offsets[s];
            const s3_search_result &R = second ? one : exact;
            const uint64_t a = second ? R.offsets[k] : R.offsets[s], b = second ? R.offsets[k + 1] : R.offsets[s + 1];
            if (second) ++k;
            unsigned long long occ = 0;
            for (uint64_t g = a; g < b; ++g) occ += (unsigned long long)(R.saR[g] - R.saL[g]) + 1;
            out->offsets[s] = total;
            if (occ == 0) out->status[s] = 0;
            else if (occ <= maxHitNum[s]) { out->status[s] = 1; total += b - a; }
            else out->status[s] = 4;
        }
        out->offsets[numSeeds] = total; out->total = total;
        const uint64_t T = total ? total : 1;
        out->saL = (uint32_t *)malloc(T * 4); out->saR = (uint32_t *)malloc(T * 4); out->strand = (uint8_t *)malloc(T);
        if (!out->saL || !out->saR || !out->strand) { rc = S3_ENOMEM; s3_set_error("s3_seed_search: out of host memory"); }
        else {
            k = 0;
            for (uint64_t s = 0; s < numSeeds; ++s) {
                const bool second = exact.offsets[s + 1] == exact.offsets[s];
                const s3_search_result &R = second ? one : exact;
                const uint64_t a = second ? R.offsets[k] : R.offsets[s];
                if (second) ++k;
                if (out->status[s] != 1) continue;
                const uint64_t cnt = out->offsets[s + 1] - out->offsets[s];
                for (uint64_t g = 0; g < cnt; ++g) {
                    out->saL[out->offsets[s] + g] = R.saL[a + g]; out->saR[out->offsets[s] + g] = R.saR[a + g];
                    out->strand[out->offsets[s] + g] = (uint8_t)((R.info[a + g] & 1u) + 1u);          // SARecord.strand 1 / 2
                }
            }
        }
    }
    free(ids1); free(q1); free(l1);
    s3_search_result_free(&exact); s3_search_result_free(&one);
    clk.lap("merge");
    if (rc) s3_seed_search_result_free(out);
    return rc;
}

// ---- the seeding driver on device arrays (the seeded DP stages) ----------------------------------------------------------
__global__ void s3_seedsrch_nohit_kernel(const unsigned long long *__restrict__ starts0, uint32_t numSeeds, uint8_t *__restrict__ flags)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < numSeeds) flags[s] = starts0[s + 1] == starts0[s];
}

// the seeds without an exact alignment, packed like packReads2 (CPUfunctions.cpp:303-338); rank[seed] = its place in that batch
__global__ void s3_seedsrch_gather_kernel(const uint32_t *__restrict__ ids1, uint32_t n1, const uint32_t *__restrict__ seeds,
                                          const uint32_t *__restrict__ seedLengths, uint32_t wordPerSeed, uint32_t *__restrict__ q1,
                                          uint32_t *__restrict__ l1, uint32_t *__restrict__ rank)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n1) return;
    const uint32_t s = ids1[k];
    const uint32_t *src = seeds + (size_t)(s / 32) * 32 * wordPerSeed + s % 32;
    uint32_t *dst = q1 + (size_t)(k / 32) * 32 * wordPerSeed + k % 32;
    for (uint32_t w = 0; w < wordPerSeed; ++w) dst[(size_t)w * 32] = src[(size_t)w * 32];
    l1[k] = seedLengths[s];
    rank[s] = k;
}

// per seed: the pass its ranges come from, how many occurrences they hold, what it keeps (count pass), then the kept ranges
// with their seed's read id, offset, length and read length (fill pass)
template <bool FILL>
__global__ void s3_seedsrch_merge_kernel(uint32_t numSeeds, const unsigned long long *__restrict__ starts0, const uint32_t *__restrict__ out0,
                                         unsigned long long total0, const uint32_t *__restrict__ rank, const unsigned long long *__restrict__ starts1,
                                         const uint32_t *__restrict__ out1, unsigned long long total1, const uint32_t *__restrict__ maxHit,
                                         uint8_t *__restrict__ status, uint32_t *__restrict__ keptCount, const uint32_t *__restrict__ keptOff,
                                         const uint32_t *__restrict__ seedReadID, const uint32_t *__restrict__ seedOffset, const uint32_t *__restrict__ seedLength,
                                         const uint32_t *__restrict__ seedReadLength, uint32_t *__restrict__ dst, uint32_t numKept)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= numSeeds) return;
    const bool second = starts0[s + 1] == starts0[s];
    unsigned long long a, b, tot;
    const uint32_t *src;
    if (second) { const uint32_t k = rank[s]; a = starts1[2 * (size_t)k]; b = starts1[2 * (size_t)k + 2]; src = out1; tot = total1; }
    else { a = starts0[s]; b = starts0[s + 1]; src = out0; tot = total0; }
    if (!FILL) {
        unsigned long long occ = 0;
        for (unsigned long long g = a; g < b; ++g) occ += (unsigned long long)(src[tot + g] - src[g]) + 1;
        const uint8_t st = occ == 0 ? 0 : (occ <= maxHit[s] ? 1 : 4);
        status[s] = st;
        keptCount[s] = st == 1 ? (uint32_t)(b - a) : 0u;
        return;
    }
    if (status[s] != 1) return;
    const uint32_t o = keptOff[s], n = (uint32_t)(b - a);
    for (uint32_t g = 0; g < n; ++g) {
        dst[o + g] = src[a + g];                                                      // saL
        dst[(size_t)numKept + o + g] = src[tot + a + g];                              // saR
        dst[2 * (size_t)numKept + o + g] = (src[2 * tot + a + g] & 1u) + 1u;          // SARecord.strand 1 / 2
        dst[3 * (size_t)numKept + o + g] = seedReadID[s];
        dst[4 * (size_t)numKept + o + g] = seedOffset[s];
        dst[5 * (size_t)numKept + o + g] = seedLength[s];
        dst[6 * (size_t)numKept + o + g] = seedReadLength[s];
    }
}

int s3_seed_search_device(s3_index *ix, const uint32_t *d_seeds, const uint32_t *d_seedLengths, uint32_t numSeeds, uint32_t wordPerSeed,
                          const uint32_t *d_maxHit, const uint32_t *d_seedReadID, const uint32_t *d_seedOffset, const uint32_t *d_seedReadLength,
                          S3SeedRangesDev *out)
{
    const uint32_t splitSeed = out->splitSeed <= numSeeds ? out->splitSeed : numSeeds;
    memset(out, 0, sizeof *out);
    if (numSeeds == 0) return S3_OK;
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    unsigned long long *d_starts0 = NULL, *d_starts1 = NULL;
    uint32_t *d_out0 = NULL, *d_out1 = NULL, *d_ids1 = NULL, *d_q1 = NULL, *d_l1 = NULL, *d_rank = NULL, *d_keptCount = NULL, *d_keptOff = NULL, *d_n1 = NULL;
    uint8_t *d_flags = NULL;
    void *d_tmp = NULL;
    unsigned long long total0 = 0, total1 = 0;
    uint32_t n1 = 0;
    uint32_t *h_cnt = (uint32_t *)ix->pinnedCount + 8;
    size_t t1 = 0, t2 = 0;
    const unsigned nb = (numSeeds + 255) / 256;
    S3_TRY(cudaMallocAsync((void **)&d_starts0, ((size_t)numSeeds + 1) * 8, st));
    if ((rc = s3_search_csr_device(ix, d_seeds, d_seedLengths, numSeeds, wordPerSeed, 0, 0, d_starts0, &d_out0, &total0))) goto done;
    S3_TRY(cudaMallocAsync((void **)&d_flags, numSeeds, st));
    S3_TRY(cudaMallocAsync((void **)&d_ids1, (size_t)numSeeds * 4 + 16, st));
    S3_TRY(cudaMallocAsync((void **)&d_rank, (size_t)numSeeds * 4, st));
    S3_TRY(cudaMallocAsync((void **)&d_n1, 16, st));
    s3_seedsrch_nohit_kernel<<<nb, 256, 0, st>>>(d_starts0, numSeeds, d_flags);
    S3_LAUNCHED(1);
    cub::DeviceSelect::Flagged(NULL, t1, cub::CountingInputIterator<uint32_t>(0), d_flags, d_ids1, d_n1, (int)numSeeds, st);
    cub::DeviceScan::ExclusiveSum(NULL, t2, (uint32_t *)NULL, (uint32_t *)NULL, (int)(numSeeds + 1), st);
    S3_TRY(cudaMallocAsync(&d_tmp, (t1 > t2 ? t1 : t2) + 16, st));
    S3_TRY(cub::DeviceSelect::Flagged(d_tmp, t1, cub::CountingInputIterator<uint32_t>(0), d_flags, d_ids1, d_n1, (int)numSeeds, st));
    S3_TRY(cudaMemcpyAsync(h_cnt, d_n1, 4, cudaMemcpyDeviceToHost, st));
    S3_TRY(cudaStreamSynchronize(st));
    n1 = h_cnt[0];
    if (n1) {
        const size_t up1 = ((size_t)n1 + 31) / 32 * 32;
        S3_TRY(cudaMallocAsync((void **)&d_q1, up1 * wordPerSeed * 4, st));
        S3_TRY(cudaMallocAsync((void **)&d_l1, up1 * 4, st));
        S3_TRY(cudaMallocAsync((void **)&d_starts1, (2 * (size_t)n1 + 1) * 8, st));
        S3_TRY(cudaMemsetAsync(d_q1, 0, up1 * wordPerSeed * 4, st));
        S3_TRY(cudaMemsetAsync(d_l1, 0, up1 * 4, st));
        s3_seedsrch_gather_kernel<<<(n1 + 255) / 256, 256, 0, st>>>(d_ids1, n1, d_seeds, d_seedLengths, wordPerSeed, d_q1, d_l1, d_rank);
        S3_LAUNCHED(1);
        if ((rc = s3_search_csr_device(ix, d_q1, d_l1, n1, wordPerSeed, 1, 0, d_starts1, &d_out1, &total1))) goto done;
    }
    S3_TRY(cudaMallocAsync((void **)&out->d_status, numSeeds, st));
    S3_TRY(cudaMallocAsync((void **)&d_keptCount, ((size_t)numSeeds + 1) * 4, st));
    S3_TRY(cudaMallocAsync((void **)&d_keptOff, ((size_t)numSeeds + 1) * 4, st));
    S3_TRY(cudaMemsetAsync(d_keptCount + numSeeds, 0, 4, st));
    s3_seedsrch_merge_kernel<false><<<nb, 256, 0, st>>>(numSeeds, d_starts0, d_out0, total0, d_rank, d_starts1, d_out1, total1, d_maxHit, out->d_status,
                                                       d_keptCount, NULL, NULL, NULL, NULL, NULL, NULL, 0);
    S3_LAUNCHED(1);
    S3_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, t2, d_keptCount, d_keptOff, (int)(numSeeds + 1), st));
    S3_TRY(cudaMemcpyAsync(h_cnt, d_keptOff + numSeeds, 4, cudaMemcpyDeviceToHost, st));
    S3_TRY(cudaMemcpyAsync(h_cnt + 1, d_keptOff + splitSeed, 4, cudaMemcpyDeviceToHost, st));
    S3_TRY(cudaStreamSynchronize(st));
    out->numRanges = h_cnt[0];
    out->splitSeed = splitSeed; out->rangesBeforeSplit = h_cnt[1];
    if (out->numRanges) {
        S3_TRY(cudaMallocAsync((void **)&out->d_buf, (size_t)out->numRanges * 28, st));
        s3_seedsrch_merge_kernel<true><<<nb, 256, 0, st>>>(numSeeds, d_starts0, d_out0, total0, d_rank, d_starts1, d_out1, total1, d_maxHit, out->d_status,
                                                          NULL, d_keptOff, d_seedReadID, d_seedOffset, d_seedLengths, d_seedReadLength, out->d_buf,
                                                          (uint32_t)out->numRanges);
        S3_LAUNCHED(1);
        S3_TRY(cudaGetLastError());
    }
done:
    {
        void *tmp[] = {d_starts0, d_starts1, d_out0, d_out1, d_ids1, d_q1, d_l1, d_rank, d_keptCount, d_keptOff, d_n1, d_flags, d_tmp};
        for (void *p : tmp) if (p) cudaFreeAsync(p, st);
    }
    if (rc) { if (out->d_buf) cudaFreeAsync(out->d_buf, st); if (out->d_status) cudaFreeAsync(out->d_status, st); memset(out, 0, sizeof *out); }
    return rc;
}
