// Paired-end pairing of the two reads' occurrence lists, batched over read pairs.
//
// Replaces, for a whole batch at once, what hostKernel does pair by pair after both reads have hits
// (CPUfunctions.cpp:2281-2310): PEMappingOccurrences (PEAlgnmt.cpp:480-547: PERadixSort of both lists, PEMappingCore)
// and PEStatsPEOutput (:777-838).  Both lists get ONE stable 64-bit radix sort each, key = read pair << 32 | position
// (= the reference's per-pair stable sort by position, ties in arrival order); then one thread per read pair walks the
// merge (s3_pair_walk.cuh), once to count and once to write records, optimal / suboptimal pair and the histogram.
#include "s3_common.cuh"
#include "s3_pair_walk.cuh"
#include "../../include/soap3dp_b200.h"

#include <cub/cub.cuh>

typedef unsigned long long s3_u64;

#define S3_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { s3_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); rc = S3_ECUDA; goto done; } } while (0)

// key = read pair << 32 | position, val = index in the caller's arrays; the read pair of element e by bisection of the CSR offsets
__global__ void s3_pair_key_kernel(const uint32_t *__restrict__ pos, const s3_u64 *__restrict__ off, uint64_t numPairs, uint64_t n,
                                   s3_u64 *__restrict__ key, uint32_t *__restrict__ val)
{
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint64_t lo = 0, hi = numPairs;                  // largest p with off[p] <= e
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (off[mid] <= e) lo = mid; else hi = mid;
    }
    key[e] = ((s3_u64)lo << 32) | pos[e];
    val[e] = (uint32_t)e;
}

template <bool FILL>
__global__ void s3_pair_walk_kernel(S3PairLists L, S3PairParams P, const s3_u64 *__restrict__ off1, const s3_u64 *__restrict__ off2,
                                    const uint32_t *__restrict__ patternLengths, uint64_t numPairs, s3_u64 *__restrict__ counts, S3PairOut O)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= numPairs) return;
    const uint32_t n = s3_pair_walk<FILL>(L, P, p, off1[p], off1[p + 1], off2[p], off2[p + 1], patternLengths[p], FILL ? counts[p] : 0, O);
    if (!FILL) counts[p] = n;
}

extern "C" int s3_pair_occurrences(s3_index *ix,
                                   const uint32_t *pos1, const uint8_t *strand1, const uint8_t *mism1, const uint64_t *off1,
                                   const uint32_t *pos2, const uint8_t *strand2, const uint8_t *mism2, const uint64_t *off2,
                                   const uint32_t *patternLengths, uint64_t numPairs,
                                   int32_t insertLbound, int32_t insertUbound, int strandLeftLeg, int strandRightLeg, int reportOne,
                                   uint64_t *pairOffsets, uint32_t **outPos1, uint32_t **outPos2, uint32_t **outInsertion, uint8_t **outFlags,
                                   uint32_t *optimal, uint32_t *suboptimal, uint32_t *mismatchStats)
{
    if (!ix || !off1 || !off2 || !pairOffsets || !outPos1 || !outPos2 || !outInsertion || !outFlags || !optimal || !suboptimal || !mismatchStats ||
        (numPairs && !patternLengths)) {
        s3_set_error("s3_pair_occurrences: NULL argument"); return S3_EINVAL;
    }
    *outPos1 = *outPos2 = *outInsertion = NULL; *outFlags = NULL;
    pairOffsets[0] = 0;
    if (numPairs == 0) return S3_OK;
    const uint64_t n1 = off1[numPairs], n2 = off2[numPairs];
    if (off1[0] != 0 || off2[0] != 0) { s3_set_error("s3_pair_occurrences: offsets must start at 0"); return S3_EINVAL; }
    if (numPairs >= 0x7FFFFFFFull || n1 >= 0x7FFFFFFFull || n2 >= 0x7FFFFFFFull) { s3_set_error("s3_pair_occurrences: batch too large"); return S3_EINVAL; }
    if ((n1 && (!pos1 || !strand1 || !mism1)) || (n2 && (!pos2 || !strand2 || !mism2))) { s3_set_error("s3_pair_occurrences: NULL list"); return S3_EINVAL; }
    if (cudaSetDevice(ix->device) != cudaSuccess) { s3_set_error("s3_pair_occurrences: cudaSetDevice failed"); return S3_ECUDA; }
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    char *d_in = NULL, *d_out = NULL;
    void *d_tmp = NULL;
    s3_u64 total = 0;
    uint32_t *h_a = NULL, *h_b = NULL, *h_i = NULL; uint8_t *h_f = NULL;
    size_t t1 = 0, t2 = 0, t3 = 0, tmpBytes = 0;
    const size_t P8 = (numPairs + 1) * 8, m1 = n1 ? n1 : 1, m2 = n2 ? n2 : 1;
    // 8-byte arrays first, then 4-byte, then bytes
    const size_t inBytes = 3 * P8 + (m1 + m2) * 2 * 8 + (m1 + m2) * 2 * 4 + (m1 + m2) * 4 + numPairs * 4 + numPairs * (4 + 4 + 32 * 4) + (m1 + m2) * 2 + 64;
    S3_TRY(cudaMalloc(&d_in, inBytes));
    {
        s3_u64 *d_off1 = (s3_u64 *)d_in, *d_off2 = d_off1 + numPairs + 1, *d_cnt = d_off2 + numPairs + 1;
        s3_u64 *k1a = d_cnt + numPairs + 1, *k1b = k1a + m1, *k2a = k1b + m1, *k2b = k2a + m2;
        uint32_t *v1a = (uint32_t *)(k2b + m2), *v1b = v1a + m1, *v2a = v1b + m1, *v2b = v2a + m2;
        uint32_t *d_pos1 = v2b + m2, *d_pos2 = d_pos1 + m1, *d_pl = d_pos2 + m2;
        uint32_t *d_opt = d_pl + numPairs, *d_sub = d_opt + numPairs, *d_stats = d_sub + numPairs;
        uint8_t *d_s1 = (uint8_t *)(d_stats + numPairs * 32), *d_m1 = d_s1 + m1, *d_s2 = d_m1 + m1, *d_m2 = d_s2 + m2;
        S3_TRY(cudaMemcpyAsync(d_off1, off1, P8, cudaMemcpyHostToDevice, st));
        S3_TRY(cudaMemcpyAsync(d_off2, off2, P8, cudaMemcpyHostToDevice, st));
        S3_TRY(cudaMemcpyAsync(d_pl, patternLengths, numPairs * 4, cudaMemcpyHostToDevice, st));
        if (n1) {
            S3_TRY(cudaMemcpyAsync(d_pos1, pos1, n1 * 4, cudaMemcpyHostToDevice, st));
            S3_TRY(cudaMemcpyAsync(d_s1, strand1, n1, cudaMemcpyHostToDevice, st));
            S3_TRY(cudaMemcpyAsync(d_m1, mism1, n1, cudaMemcpyHostToDevice, st));
        }
        if (n2) {
            S3_TRY(cudaMemcpyAsync(d_pos2, pos2, n2 * 4, cudaMemcpyHostToDevice, st));
            S3_TRY(cudaMemcpyAsync(d_s2, strand2, n2, cudaMemcpyHostToDevice, st));
            S3_TRY(cudaMemcpyAsync(d_m2, mism2, n2, cudaMemcpyHostToDevice, st));
        }
        S3_TRY(cudaMemsetAsync(d_cnt, 0, P8, st));
        S3_TRY(cudaMemsetAsync(d_stats, 0, numPairs * 32 * 4, st));
        cub::DeviceRadixSort::SortPairs(NULL, t1, k1a, k1b, v1a, v1b, (int)m1, 0, 64, st);
        cub::DeviceRadixSort::SortPairs(NULL, t2, k2a, k2b, v2a, v2b, (int)m2, 0, 64, st);
        cub::DeviceScan::ExclusiveSum(NULL, t3, d_cnt, d_cnt, (int)(numPairs + 1), st);
        tmpBytes = t1 > t2 ? t1 : t2;
        if (t3 > tmpBytes) tmpBytes = t3;
        S3_TRY(cudaMalloc(&d_tmp, tmpBytes));
        if (n1) {
            s3_pair_key_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(d_pos1, d_off1, numPairs, n1, k1a, v1a);
            S3_LAUNCHED(1);
            S3_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, t1, k1a, k1b, v1a, v1b, (int)n1, 0, 64, st));   // stable: ties stay in arrival order
        }
        if (n2) {
            s3_pair_key_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(d_pos2, d_off2, numPairs, n2, k2a, v2a);
            S3_LAUNCHED(1);
            S3_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, t2, k2a, k2b, v2a, v2b, (int)n2, 0, 64, st));
        }
        S3PairLists L = {k1b, k2b, v1b, v2b, d_s1, d_m1, d_s2, d_m2};
        S3PairParams P = {(uint32_t)insertLbound, (uint32_t)insertUbound, strandLeftLeg, strandRightLeg, reportOne};
        S3PairOut O = {NULL, NULL, NULL, NULL, d_opt, d_sub, d_stats};
        const unsigned blocks = (unsigned)((numPairs + 127) / 128);
        s3_pair_walk_kernel<false><<<blocks, 128, 0, st>>>(L, P, d_off1, d_off2, d_pl, numPairs, d_cnt, O);
        S3_LAUNCHED(1);
        S3_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, t3, d_cnt, d_cnt, (int)(numPairs + 1), st));
        S3_TRY(cudaMemcpyAsync(pairOffsets, d_cnt, P8, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaStreamSynchronize(st));
        total = pairOffsets[numPairs];
        if (total >= 0x7FFFFFFFull) { s3_set_error("s3_pair_occurrences: %llu pairs in one call", total); rc = S3_EINVAL; goto done; }
        const size_t T = total ? (size_t)total : 1;
        S3_TRY(cudaMalloc(&d_out, T * 16 + 64));
        O.pos1 = (uint32_t *)d_out; O.pos2 = O.pos1 + T; O.insertion = O.pos2 + T; O.flags = (uint8_t *)(O.insertion + T);
        s3_pair_walk_kernel<true><<<blocks, 128, 0, st>>>(L, P, d_off1, d_off2, d_pl, numPairs, d_cnt, O);
        S3_LAUNCHED(1);
        S3_TRY(cudaGetLastError());
        S3_TRY(cudaMemcpyAsync(optimal, d_opt, numPairs * 4, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaMemcpyAsync(suboptimal, d_sub, numPairs * 4, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaMemcpyAsync(mismatchStats, d_stats, numPairs * 32 * 4, cudaMemcpyDeviceToHost, st));
        if (total) {
            h_a = (uint32_t *)malloc(T * 4); h_b = (uint32_t *)malloc(T * 4); h_i = (uint32_t *)malloc(T * 4); h_f = (uint8_t *)malloc(T * 4);
            if (!h_a || !h_b || !h_i || !h_f) { s3_set_error("s3_pair_occurrences: out of host memory"); rc = S3_ENOMEM; goto done; }
            S3_TRY(cudaMemcpyAsync(h_a, O.pos1, T * 4, cudaMemcpyDeviceToHost, st));
            S3_TRY(cudaMemcpyAsync(h_b, O.pos2, T * 4, cudaMemcpyDeviceToHost, st));
            S3_TRY(cudaMemcpyAsync(h_i, O.insertion, T * 4, cudaMemcpyDeviceToHost, st));
            S3_TRY(cudaMemcpyAsync(h_f, O.flags, T * 4, cudaMemcpyDeviceToHost, st));
        }
        S3_TRY(cudaStreamSynchronize(st));
        *outPos1 = h_a; *outPos2 = h_b; *outInsertion = h_i; *outFlags = h_f;
        h_a = h_b = h_i = NULL; h_f = NULL;
    }
done:
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    if (d_tmp) cudaFree(d_tmp);
    free(h_a); free(h_b); free(h_i); free(h_f);
    return rc;
}

// =====================================================================================================================
// Best-hit filters on every read's SA-range list and occurrence list.  Replaces retainAllBest / retainAllBestWithCap /
// retainAllBestAndSecBest (SAList.cpp:140-348) as hostKernel applies them read by read before DP and pairing
// (CPUfunctions.cpp:2170-2255): one thread per read, a count pass, two prefix sums, a fill pass (s3_retain_walk.cuh).
// =====================================================================================================================
#include "s3_retain_walk.cuh"

template <bool FILL>
__global__ void s3_retain_kernel(S3RetainIn I, int mode, int32_t maxNum, const s3_u64 *__restrict__ saOff, const s3_u64 *__restrict__ occOff,
                                 uint64_t numReads, s3_u64 *__restrict__ keptSa, s3_u64 *__restrict__ keptOcc, S3RetainOut O, uint32_t *__restrict__ num)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= numReads) return;
    uint32_t ks, ko;
    const uint32_t n = s3_retain_walk<FILL>(I, mode, maxNum, saOff[r], saOff[r + 1], occOff[r], occOff[r + 1],
                                            FILL ? keptSa[r] : 0, FILL ? keptOcc[r] : 0, O, &ks, &ko);
    if (FILL) num[r] = n;
    else { keptSa[r] = ks; keptOcc[r] = ko; }
}

extern "C" int s3_retain_best(s3_index *ix, int mode, int32_t maxNum,
                              const uint32_t *saL, const uint32_t *saR, const uint8_t *saStrand, const uint8_t *saMism, const uint64_t *saOff,
                              const uint32_t *occPos, const uint8_t *occStrand, const uint8_t *occMism, const uint64_t *occOff, uint64_t numReads,
                              uint64_t *outSaOff, uint32_t *outSaL, uint32_t *outSaR, uint8_t *outSaFlags,
                              uint64_t *outOccOff, uint32_t *outOccPos, uint8_t *outOccFlags, uint32_t *num)
{
    if (!ix || !saOff || !occOff || !outSaOff || !outOccOff || (numReads && !num)) { s3_set_error("s3_retain_best: NULL argument"); return S3_EINVAL; }
    if (mode < S3_RETAIN_ALL_BEST || mode > S3_RETAIN_BEST_AND_SECOND || (mode == S3_RETAIN_BEST_WITH_CAP && maxNum < 1)) {
        s3_set_error("s3_retain_best: mode %d / maxNum %d", mode, maxNum); return S3_EINVAL;
    }
    outSaOff[0] = outOccOff[0] = 0;
    if (numReads == 0) return S3_OK;
    if (saOff[0] != 0 || occOff[0] != 0) { s3_set_error("s3_retain_best: offsets must start at 0"); return S3_EINVAL; }
    const uint64_t nS = saOff[numReads], nO = occOff[numReads];
    if (numReads >= 0x7FFFFFFFull || nS >= 0x7FFFFFFFull || nO >= 0x7FFFFFFFull) { s3_set_error("s3_retain_best: batch too large"); return S3_EINVAL; }
    if ((nS && (!saL || !saR || !saStrand || !saMism || !outSaL || !outSaR || !outSaFlags)) || (nO && (!occPos || !occStrand || !occMism || !outOccPos || !outOccFlags))) {
        s3_set_error("s3_retain_best: NULL list"); return S3_EINVAL;
    }
    if (cudaSetDevice(ix->device) != cudaSuccess) { s3_set_error("s3_retain_best: cudaSetDevice failed"); return S3_ECUDA; }
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    char *d_buf = NULL;
    void *d_tmp = NULL;
    size_t t1 = 0;
    const size_t R8 = (numReads + 1) * 8, mS = nS ? nS : 1, mO = nO ? nO : 1;
    // 8-byte arrays, then 4-byte, then bytes
    const size_t bytes = 4 * R8 + (4 * mS + 2 * mO + numReads) * 4 + (2 * mS + 2 * mO) + (2 * mS + 2 * mO) + 64;
    S3_TRY(cudaMalloc(&d_buf, bytes));
    {
        s3_u64 *d_saOff = (s3_u64 *)d_buf, *d_occOff = d_saOff + numReads + 1, *d_kS = d_occOff + numReads + 1, *d_kO = d_kS + numReads + 1;
        uint32_t *d_saL = (uint32_t *)(d_kO + numReads + 1), *d_saR = d_saL + mS, *d_oL = d_saR + mS, *d_oR = d_oL + mS;
        uint32_t *d_occ = d_oR + mS, *d_oOcc = d_occ + mO, *d_num = d_oOcc + mO;
        uint8_t *d_sS = (uint8_t *)(d_num + numReads), *d_sM = d_sS + mS, *d_cS = d_sM + mS, *d_cM = d_cS + mO;
        uint8_t *d_oSF = d_cM + mO, *d_oOF = d_oSF + 2 * mS;
        S3_TRY(cudaMemcpyAsync(d_saOff, saOff, R8, cudaMemcpyHostToDevice, st));
        S3_TRY(cudaMemcpyAsync(d_occOff, occOff, R8, cudaMemcpyHostToDevice, st));
        if (nS) {
            S3_TRY(cudaMemcpyAsync(d_saL, saL, nS * 4, cudaMemcpyHostToDevice, st)); S3_TRY(cudaMemcpyAsync(d_saR, saR, nS * 4, cudaMemcpyHostToDevice, st));
            S3_TRY(cudaMemcpyAsync(d_sS, saStrand, nS, cudaMemcpyHostToDevice, st)); S3_TRY(cudaMemcpyAsync(d_sM, saMism, nS, cudaMemcpyHostToDevice, st));
        }
        if (nO) {
            S3_TRY(cudaMemcpyAsync(d_occ, occPos, nO * 4, cudaMemcpyHostToDevice, st));
            S3_TRY(cudaMemcpyAsync(d_cS, occStrand, nO, cudaMemcpyHostToDevice, st)); S3_TRY(cudaMemcpyAsync(d_cM, occMism, nO, cudaMemcpyHostToDevice, st));
        }
        S3_TRY(cudaMemsetAsync(d_kS, 0, 2 * R8, st));
        cub::DeviceScan::ExclusiveSum(NULL, t1, d_kS, d_kS, (int)(numReads + 1), st);
        S3_TRY(cudaMalloc(&d_tmp, t1));
        S3RetainIn I = {d_saL, d_saR, d_sS, d_sM, d_occ, d_cS, d_cM};
        S3RetainOut O = {d_oL, d_oR, d_oSF, d_oOcc, d_oOF};
        const unsigned blocks = (unsigned)((numReads + 127) / 128);
        s3_retain_kernel<false><<<blocks, 128, 0, st>>>(I, mode, maxNum, d_saOff, d_occOff, numReads, d_kS, d_kO, O, d_num);
        S3_LAUNCHED(1);
        S3_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, t1, d_kS, d_kS, (int)(numReads + 1), st));
        S3_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, t1, d_kO, d_kO, (int)(numReads + 1), st));
        s3_retain_kernel<true><<<blocks, 128, 0, st>>>(I, mode, maxNum, d_saOff, d_occOff, numReads, d_kS, d_kO, O, d_num);
        S3_LAUNCHED(1);
        S3_TRY(cudaGetLastError());
        S3_TRY(cudaMemcpyAsync(outSaOff, d_kS, R8, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaMemcpyAsync(outOccOff, d_kO, R8, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaMemcpyAsync(num, d_num, numReads * 4, cudaMemcpyDeviceToHost, st));
        S3_TRY(cudaStreamSynchronize(st));
        const uint64_t kS = outSaOff[numReads], kO = outOccOff[numReads];
        if (kS) {
            S3_TRY(cudaMemcpyAsync(outSaL, d_oL, kS * 4, cudaMemcpyDeviceToHost, st)); S3_TRY(cudaMemcpyAsync(outSaR, d_oR, kS * 4, cudaMemcpyDeviceToHost, st));
            S3_TRY(cudaMemcpyAsync(outSaFlags, d_oSF, kS * 2, cudaMemcpyDeviceToHost, st));
        }
        if (kO) {
            S3_TRY(cudaMemcpyAsync(outOccPos, d_oOcc, kO * 4, cudaMemcpyDeviceToHost, st));
            S3_TRY(cudaMemcpyAsync(outOccFlags, d_oOF, kO * 2, cudaMemcpyDeviceToHost, st));
        }
        S3_TRY(cudaStreamSynchronize(st));
    }
done:
    if (d_buf) cudaFree(d_buf);
    if (d_tmp) cudaFree(d_tmp);
    return rc;
}
