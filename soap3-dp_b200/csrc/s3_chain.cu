// s3_chain.cu -- a batch of read pairs from queries to alignments without leaving the device.
//
// Replaces, for a batch, what soap3_dp_pair_align does between its GPU calls on the host
// (alignment.cu:1896-2330): all_valid_alignment's round 1 (alignment.cu:855), hostKernel's per-pair work
// (CPUfunctions.cpp:1498-2620: collect_all_answers :1226, the routing of a pair by which mates have hits :2153-2262 and
// :2440-2530, transferAllSAToOcc SAList.cpp:392, PEMappingOccurrences + PEStatsPEOutput PEAlgnmt.cpp:480,777) and the
// default-DP engine's mate rescue (HalfEndOccStream::fetchNextOcc DV-DPfunctions.cu:1900, HalfEndAlgnBatch::pack :2027,
// SemiGlobalAligner::performAlignment :669, the result loop of DP_Space::algnmtCPUThread :2359-2420).  The reference
// ships four padded buffers across PCIe per batch (queries in, 8-word answer slots per read and case out, packed
// windows in, 509-byte traceback patterns out); here the queries go in and a few dozen bytes per pair come out.
//
//   search (round-1 slots, all cases)           s3_search_round1_device
//   collect   per read: its SA ranges over the cases in slot order, capped at MaxOutputPerRead occurrences
//   route     per pair: both mates hit -> pairing; one -> mate rescue (best hits only when there are more than
//             maxHitNumForDP); none -> the both-unaligned list (deep DP is the caller's next stage)
//   locate    the occurrences of every read that is paired or rescues its mate, from the suffix array
//   pair      one stable sort by (read, position); per pair the reference's merge walk: number of valid pairs,
//             the optimal pair, the second-best total
//   route 2   both mates hit but no valid pair: both rescue each other when neither has more than maxHitNumForDP hits
//   windows   per rescuing occurrence the window HalfEndAlgnBatch::pack cuts, in HalfEndOccStream's order
//   DP        s3_dp_align_windows_device
//   results   per window that reached its cutoff: position, score, tie count, CIGAR as (op, length) runs in read order
// Two small device-to-host reads of counts size the later stages (occurrences, windows); everything else is
// stream-ordered.
#include "s3_common.cuh"
#include "s3_pair_walk.cuh"
#include "s3_windows.cuh"
#include "../../include/soap3dp_b200.h"

#include <cub/cub.cuh>
#include <stdlib.h>
#include <string.h>

#define S3_TRYC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { s3_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); return S3_ECUDA; } } while (0)

struct S3Arena {
    char *base; size_t cap, used;
};
static int arena_reserve(S3Arena *a, size_t bytes, cudaStream_t st)
{
    a->used = 0;
    if (bytes <= a->cap) return S3_OK;
    if (a->base) { S3_TRYC(cudaStreamSynchronize(st)); S3_TRYC(cudaFree(a->base)); a->base = NULL; a->cap = 0; }
    bytes += bytes / 4;
    S3_TRYC(cudaMalloc(&a->base, bytes));
    a->cap = bytes;
    return S3_OK;
}
template <typename T> static T *arena_take(S3Arena *a, size_t n)
{
    const size_t off = (a->used + 255) / 256 * 256;
    a->used = off + n * sizeof(T);
    return a->used <= a->cap ? reinterpret_cast<T *>(a->base + off) : NULL;
}
static size_t arena_need(size_t n, size_t sz) { return n * sz + 256; }

struct s3_pe {
    s3_index *ix;
    s3_dp *dp;
    s3_pe_params par;
    uint32_t maxReads, maxReadLength, maxDNALength, maxWindows;
    S3Arena A, B, C;                 // stage buffers: sized by the batch, by the occurrences, by the windows
    void *pinned; size_t pinnedBytes;
    uint32_t *h_counts;              // pinned: what the two mid-chain reads bring back
    float msStages[8];
    cudaEvent_t ev[10];
    int timing;
    // input double buffer of the host entry: the next batch's queries go up on a stream of their own (s3_pe_prefetch)
    // while the current batch is being aligned
    uint32_t *d_in[2]; size_t inBytes[2];
    cudaStream_t copyStream; cudaEvent_t pfDone[2];
    struct { const uint32_t *queries; uint64_t reads; uint32_t wpq; } pf[2];     // what has been prefetched into d_in[k] and not aligned yet (queries NULL: nothing)
    // the batch of the last s3_pe_align[_device] call, still on the device: what s3_pe_deep_dp works on
    struct { const uint32_t *d_q, *d_len; const uint8_t *d_route; uint32_t N, wpq; } last;
};

static int pe_input_buffer(s3_pe *pe, int k, size_t bytes, cudaStream_t st)
{
    if (bytes <= pe->inBytes[k]) return S3_OK;
    if (pe->d_in[k]) { S3_TRYC(cudaStreamSynchronize(st)); S3_TRYC(cudaFree(pe->d_in[k])); pe->d_in[k] = NULL; pe->inBytes[k] = 0; }
    S3_TRYC(cudaMalloc(&pe->d_in[k], bytes));
    pe->inBytes[k] = bytes;
    return S3_OK;
}

// ---- collect (collect_all_answers, CPUfunctions.cpp:1226-1300, round-1 slots only) ------------------------------------
// One thread per read.  COUNT: ranges and occurrences of the read; FILL: the ranges at rangeOff[read].
// A case whose status word is > 0xFFFFFFFD overflowed its slot (isMoreThanSA1): the reference searches such a read
// again (round 2, then the CPU); here the read is flagged and its pair reported as S3_PE_OVERFLOW.
struct S3Collect {
    const uint32_t *answers[S3_MAX_NUM_CASES];
    uint32_t numCases, allowed, wordPerAns, numReads, textLength, maxOutputPerRead;
};
template <bool FILL>
__global__ void s3_pe_collect_kernel(const S3Collect c, uint32_t *__restrict__ nRanges, uint32_t *__restrict__ totOcc, uint8_t *__restrict__ readFlags,
                                     const uint32_t *__restrict__ rangeOff, uint32_t *__restrict__ saL, uint32_t *__restrict__ saR,
                                     uint8_t *__restrict__ saFlags, s3_pe_read_stats *__restrict__ stats = NULL)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= c.numReads) return;
    const size_t off = (size_t)(r >> 5) * 32 * c.wordPerAns + (r & 31);
    uint32_t n = 0, tot = 0;
    uint32_t withError[8] = {0, 0, 0, 0, 0, 0, 0, 0};            // rOutput->WithError: occurrences per mismatch count (:1294)
    bool more = false;
    const uint32_t base = FILL ? rangeOff[r] : 0u;
    for (uint32_t w = 0; w < c.numCases; ++w) {
        const uint32_t *ans = c.answers[w] + off;
        const uint32_t first = ans[0];
        if (first > 0xFFFFFFFDu) { more = true; continue; }
        if (first == 0xFFFFFFFDu) continue;
        for (uint32_t i = 0; i < c.allowed; ++i) {
            const uint32_t a0 = ans[(size_t)(2 * i) * 32], a1 = ans[(size_t)(2 * i + 1) * 32];
            if (a0 >= 0xFFFFFFFDu || a1 >= 0xFFFFFFFDu) break;
            const uint32_t l = a0;
            uint32_t rr = l + (a1 & 0xFFFFFFu);
            if (!(l <= rr && rr <= c.textLength)) break;
            if (tot < c.maxOutputPerRead) {
                if (tot + (rr - l + 1) > c.maxOutputPerRead) rr = l + c.maxOutputPerRead - tot - 1;
                if (FILL) {
                    saL[base + n] = l; saR[base + n] = rr;
                    saFlags[2 * (size_t)(base + n)] = (uint8_t)(((a1 >> 27) & 1u) + 1u);        // strand 1 / 2
                    saFlags[2 * (size_t)(base + n) + 1] = (uint8_t)((a1 >> 24) & 7u);           // mismatches
                }
                ++n; tot += rr - l + 1;
                if (!FILL) withError[(a1 >> 24) & 7u] += rr - l + 1;
            }
            if (tot >= c.maxOutputPerRead) break;
        }
    }
    if (!FILL) {
        nRanges[r] = n; totOcc[r] = tot; readFlags[r] = more ? 1 : 0;
        if (stats) {
            // what hostKernel keeps of a read for MAPQ (CPUfunctions.cpp:2061-2141): the fewest mismatches, X0 = the occurrences with that
            // many, X1 = those with one more
            s3_pe_read_stats o = {0u, 0u, 255u, {0, 0, 0}};
            for (uint32_t m = 0; m < 8; ++m) if (withError[m]) { o.minMismatch = (uint8_t)m; o.x0 = withError[m]; o.x1 = m < 7 ? withError[m + 1] : 0u; break; }
            stats[r] = o;
        }
    }
}

// ---- route (CPUfunctions.cpp:2153-2262) ------------------------------------------------------------------------------
// One thread per pair.  A mate that rescues the other with more than maxHitNumForDP occurrences keeps its best hits only
// (retainAllBest, or retainAllBestAndSecBest when mapping qualities are wanted: SAList.cpp:140-348 on a list without
// occurrences = the ranges with the minimum / minimum + 1 mismatches); still too many -> the new-default-DP list.
// keepRange[g] = 1 for the ranges that stay; locCount[r] = occurrences of read r to locate.
__global__ void s3_pe_route_kernel(uint32_t numPairs, const uint32_t *__restrict__ rangeOff, const uint32_t *__restrict__ saL,
                                   const uint32_t *__restrict__ saR, const uint8_t *__restrict__ saFlags, const uint32_t *__restrict__ totOcc,
                                   const uint8_t *__restrict__ readFlags, uint32_t maxHit, int keepSecondBest,
                                   uint8_t *__restrict__ route, uint8_t *__restrict__ keepRange, uint32_t *__restrict__ locCount,
                                   uint32_t *__restrict__ counters)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= numPairs) return;
    const uint32_t r0 = 2 * p, r1 = 2 * p + 1;
    const uint32_t t0 = totOcc[r0], t1 = totOcc[r1];
    uint8_t rt;
    uint32_t c0 = 0, c1 = 0;
    if (readFlags[r0] || readFlags[r1]) rt = S3_PE_OVERFLOW;
    else if (t0 > 0 && t1 > 0) { rt = S3_PE_BOTH_HIT; c0 = t0; c1 = t1; }
    else if (t0 == 0 && t1 == 0) rt = S3_PE_NONE;
    else {
        const uint32_t r = t0 ? r0 : r1;
        uint32_t tot = t0 ? t0 : t1;
        if (tot > maxHit) {
            int mn = 999;
            for (uint32_t g = rangeOff[r]; g < rangeOff[r + 1]; ++g) mn = min(mn, (int)saFlags[2 * (size_t)g + 1]);
            tot = 0;
            for (uint32_t g = rangeOff[r]; g < rangeOff[r + 1]; ++g) {
                const bool keep = (int)saFlags[2 * (size_t)g + 1] <= mn + (keepSecondBest ? 1 : 0);
                keepRange[g] = keep ? 1 : 0;
                if (keep) tot += saR[g] - saL[g] + 1;
            }
        }
        if (tot <= maxHit) { rt = t0 ? S3_PE_FIRST_RESCUES : S3_PE_SECOND_RESCUES; if (t0) c0 = tot; else c1 = tot; }
        else rt = t0 ? S3_PE_FIRST_TOO_MANY : S3_PE_SECOND_TOO_MANY;
    }
    route[p] = rt;
    locCount[r0] = c0; locCount[r1] = c1;
    atomicAdd(counters + rt, 1u);
}

// ---- locate (transferAllSAToOcc SAList.cpp:392-419; HalfEndOccStream::fetchNextOcc DV-DPfunctions.cu:1900-1960) ------
// One thread per read: the kept ranges in list order, each in suffix-array order.
__global__ void s3_pe_locate_kernel(uint32_t numReads, const uint32_t *__restrict__ sa, const uint32_t *__restrict__ rangeOff,
                                    const uint32_t *__restrict__ saL, const uint32_t *__restrict__ saR, const uint8_t *__restrict__ saFlags,
                                    const uint8_t *__restrict__ keepRange, const uint32_t *__restrict__ locOff,
                                    uint32_t *__restrict__ occPos, uint8_t *__restrict__ occFlags, unsigned long long *__restrict__ occKey,
                                    uint32_t *__restrict__ occVal)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= numReads) return;
    uint32_t o = locOff[r];
    const uint32_t end = locOff[r + 1];
    for (uint32_t g = rangeOff[r]; g < rangeOff[r + 1] && o < end; ++g) {
        if (!keepRange[g]) continue;
        for (uint32_t k = saL[g]; k <= saR[g] && o < end; ++k, ++o) {
            const uint32_t pos = sa[k];
            occPos[o] = pos;
            occFlags[2 * (size_t)o] = saFlags[2 * (size_t)g]; occFlags[2 * (size_t)o + 1] = saFlags[2 * (size_t)g + 1];
            occKey[o] = ((unsigned long long)r << 32) | pos;       // one stable sort orders every read's list by position
            occVal[o] = o;
        }
    }
}

// ---- pairing (PEMappingCore / PEStatsPEPairList, through s3_pair_walk.cuh's walk) --------------------------------------
// One thread per pair with both mates hit: the walk's records are not kept, only what hostKernel reads of them
// (CPUfunctions.cpp:2293-2330): how many valid pairs, the optimal pair, how many pairs share its total, the second total.
typedef s3_pe_pair_result S3PeBest;            // include/soap3dp_b200.h
// The reference walks the two position-ordered lists as one merge: the element that goes first (list 1 on equal
// positions) is a left leg and is tried against the other list from that list's cursor.  That cursor is a function of the
// element alone -- for a list-1 element the first list-2 position >= its own, for a list-2 element the first list-1
// position > its own -- so the elements are independent: ONE WARP per pair, the lanes take the elements in turn (a pair of
// tandem-repeat reads has 10^3 x 10^3 candidates; one thread per pair left a 38 ms tail).  What the walk's order decides is
// rebuilt from an emission key (rank of the left leg in the merge << 32 | partner index): the optimal pair is the first
// record with the smallest (total mismatches, mismatch difference); PEStatsPEPairList's suboptimal pair is the optimal one
// it displaced last, i.e. its total is the smallest total among the records emitted before the first record of the final
// optimal total.
// flags of the occurrences in sorted order, so that a scan reads two arrays front to back instead of chasing val[]
__global__ void s3_pe_sorted_flags_kernel(uint32_t n, const uint32_t *__restrict__ val, const uint8_t *__restrict__ occFlags, uint16_t *__restrict__ sflags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sflags[i] = reinterpret_cast<const uint16_t *>(occFlags)[val[i]];          // strand | mismatches << 8
}

// One group of THREADS threads per pair: a warp for the ordinary pair (HEAVY false: pairs with more than
// S3_PE_HEAVY_ELEMENTS occurrences are put on a list instead), a whole block for the listed ones (HEAVY true).
#define S3_PE_PAIR_WARPS 4
#define S3_PE_HEAVY_ELEMENTS 192u
template <bool HEAVY>
__global__ void __launch_bounds__(S3_PE_PAIR_WARPS * 32)
s3_pe_pair_kernel(uint32_t numPairs, const uint8_t *__restrict__ route, const uint32_t *__restrict__ locOff,
                  const unsigned long long *__restrict__ key, const uint16_t *__restrict__ sflags,
                  const uint32_t *__restrict__ readLengths, S3PairParams P, S3PeBest *__restrict__ best,
                  uint32_t *__restrict__ heavyList, uint32_t *__restrict__ heavyCount)
{
    constexpr uint32_t THREADS = HEAVY ? S3_PE_PAIR_WARPS * 32 : 32, GROUPS = S3_PE_PAIR_WARPS * 32 / THREADS;
    __shared__ uint32_t sStats[GROUPS][16];
    __shared__ unsigned long long sFirst[GROUPS][16];
    __shared__ uint32_t sN[GROUPS], sTD[GROUPS];
    __shared__ unsigned long long sKey[GROUPS];
    const uint32_t grp = HEAVY ? 0u : threadIdx.x >> 5, t = HEAVY ? threadIdx.x : (threadIdx.x & 31);
    uint32_t p;
    if (HEAVY) { if (blockIdx.x >= *heavyCount) return; p = heavyList[blockIdx.x]; }
    else { p = blockIdx.x * GROUPS + grp; if (p >= numPairs) return; }
    S3PeBest b;
    memset(&b, 0, sizeof b);
    b.optimalTotal = b.suboptimalTotal = 127;
    if (route[p] != S3_PE_BOTH_HIT) { if (t == 0) best[p] = b; return; }
    const uint32_t a0 = locOff[2 * p], a1 = locOff[2 * p + 1], b1 = locOff[2 * p + 2];
    if (!HEAVY && b1 - a0 > S3_PE_HEAVY_ELEMENTS) { if (t == 0) heavyList[atomicAdd(heavyCount, 1u)] = p; return; }
    const uint32_t patternLength = readLengths[2 * p + 1];            // pe_in->patternLength = the second read's (CPUfunctions.cpp:2284)
    if (t < 16) { sStats[grp][t] = 0; sFirst[grp][t] = ~0ull; }
    if (t == 0) { sN[grp] = 0; sTD[grp] = 0xFFFFFFFFu; sKey[grp] = ~0ull; }
    if (HEAVY) __syncthreads(); else __syncwarp();
    uint32_t n = 0;
    uint32_t bestTD = 0xFFFFFFFFu;                  // total << 8 | difference of this thread's best record
    unsigned long long bestKey = ~0ull;
    uint32_t bl = 0, br = 0, bgap = 0;
    uint16_t bfl = 0, bfr = 0;
    bool bFirstIsLeft = true;
    for (uint32_t e = a0 + t; e < b1; e += THREADS) {
        const bool firstIsLeft = e < a1;            // an element of list 1 (the first read's) or of list 2
        const uint32_t lpos = (uint32_t)key[e];
        const uint16_t lfl = sflags[e];
        const uint8_t lstrand = (uint8_t)(lfl & 0xFF);
        if (lstrand != P.leftLeg) continue;
        // cursor in the other list, and this element's rank in the merge
        uint32_t lo = firstIsLeft ? a1 : a0, hi = firstIsLeft ? b1 : a1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1, mp = (uint32_t)key[mid];
            if (firstIsLeft ? (mp < lpos) : (mp <= lpos)) lo = mid + 1; else hi = mid;
        }
        const uint32_t end = firstIsLeft ? b1 : a1;
        const unsigned long long rank = (unsigned long long)((e - (firstIsLeft ? a0 : a1)) + (lo - (firstIsLeft ? a1 : a0))) << 32;
        for (uint32_t i = lo; i < end; ++i) {
            const uint32_t rpos = (uint32_t)key[i];
            const uint16_t rfl = sflags[i];
            const uint8_t rstrand = (uint8_t)(rfl & 0xFF);
            const uint32_t rightEnd = rpos + patternLength - 1u, gap = rightEnd - lpos + 1u;
            if (P.lbound <= gap && gap <= P.ubound && rstrand == P.rightLeg) {
                const uint8_t m1 = (uint8_t)((firstIsLeft ? lfl : rfl) >> 8), m2 = (uint8_t)((firstIsLeft ? rfl : lfl) >> 8);
                const int tot = (int8_t)(uint8_t)(m1 + m2);
                int d = (int)(int8_t)m1 - (int)(int8_t)m2;
                if ((int8_t)m2 > (int8_t)m1) d = -d;
                const unsigned long long k = rank | (i - lo);
                if (tot >= 0 && tot < 16) { atomicAdd(&sStats[grp][tot], 1u); atomicMin(&sFirst[grp][tot], k); }
                const uint32_t td = ((uint32_t)(uint8_t)tot << 8) | (uint32_t)(uint8_t)d;
                if (td < bestTD || (td == bestTD && k < bestKey)) {
                    bestTD = td; bestKey = k; bl = lpos; br = rpos; bgap = gap; bfl = lfl; bfr = rfl; bFirstIsLeft = firstIsLeft;
                }
                ++n;
            }
            if (lstrand != rstrand && (uint32_t)(lpos + P.ubound) < rightEnd) break;     // PEIsPairOutOfRange
        }
    }
    // the group's records: count, the best (total, difference) and among those the first emitted
    if (n) atomicAdd(&sN[grp], n);
    if (bestTD != 0xFFFFFFFFu) atomicMin(&sTD[grp], bestTD);
    if (HEAVY) __syncthreads(); else __syncwarp();
    const uint32_t gTD = sTD[grp];
    if (bestTD == gTD && bestTD != 0xFFFFFFFFu) atomicMin(&sKey[grp], bestKey);
    if (HEAVY) __syncthreads(); else __syncwarp();
    const uint32_t nAll = sN[grp];
    if (nAll == 0) { if (t == 0) best[p] = b; return; }
    if (bestTD == gTD && bestKey == sKey[grp]) {
        const int tot = (int)(gTD >> 8);
        const uint16_t f1 = bFirstIsLeft ? bfl : bfr, f2 = bFirstIsLeft ? bfr : bfl;
        b.pos1 = bFirstIsLeft ? bl : br; b.pos2 = bFirstIsLeft ? br : bl; b.insertion = bgap;
        b.strand1 = (uint8_t)(f1 & 0xFF); b.mism1 = (uint8_t)(f1 >> 8);
        b.strand2 = (uint8_t)(f2 & 0xFF); b.mism2 = (uint8_t)(f2 >> 8);
        b.numPairs = nAll;
        b.optimalTotal = (int8_t)tot;
        b.numOptimal = (tot < 16) ? sStats[grp][tot] : 0;
        // the optimal pair displaced last: the smallest total emitted before the first record of the final total
        int sub = 127;
        if (tot < 16) { for (int u = tot + 1; u < 16; ++u) if (sFirst[grp][u] < sFirst[grp][tot]) { sub = u; break; } }
        b.suboptimalTotal = (int8_t)sub;
        b.numSuboptimal = (sub < 16) ? sStats[grp][sub] : 0;
        best[p] = b;
    }
}

// ---- route 2 + windows (CPUfunctions.cpp:2440-2470; HalfEndAlgnBatch::pack DV-DPfunctions.cu:2027-2110) --------------
// Which reads hand their occurrences to the default DP: the hit mate of a *_RESCUES pair; both mates of a pair with hits
// on both sides and no valid pairing, when neither has more than maxHitNumForDP occurrences (more: left to the caller,
// S3_PE_BOTH_NO_PAIR_MANY -- the reference filters them down to their best hits first).
// One thread per read; COUNT: windows of the read, FILL: their descriptors from winOff[read].
struct S3PeWindows {
    uint32_t *alignedOcc;            // index of the occurrence the window hangs on
    uint32_t *readID;                // the read that is aligned by DP (the mate)
    uint8_t *strand;                 // its strand in the window (1 as given, 2 reverse-complemented)
    uint8_t *leftOrRight;            // CandidateInfo.leftOrRight: 1 the DP read is the right end, 0 the left end
    uint32_t *start, *dnaLen, *readLen, *clipLt, *clipRt, *ancL, *ancR;
    int32_t *cutoff;
};
struct S3PeWinParams {
    S3WinParams w;                   // what HalfEndAlgnBatch::pack decides with (csrc/s3_windows.cuh)
    uint32_t maxHit;
};
template <bool FILL>
__global__ void s3_pe_window_kernel(uint32_t numReads, const uint8_t *__restrict__ route, uint8_t *__restrict__ routeFinal, const S3PeBest *__restrict__ best,
                                    const uint32_t *__restrict__ locOff, const uint32_t *__restrict__ occPos, const uint8_t *__restrict__ occFlags,
                                    const uint32_t *__restrict__ readLengths, const S3PeWinParams w, uint32_t *__restrict__ winCount,
                                    const uint32_t *__restrict__ winOff, S3PeWindows out)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= numReads) return;
    const uint32_t p = r >> 1;
    uint8_t rt = route[p];
    if (rt == S3_PE_BOTH_HIT && best[p].numPairs == 0) {
        const uint32_t t0 = locOff[2 * p + 1] - locOff[2 * p], t1 = locOff[2 * p + 2] - locOff[2 * p + 1];
        rt = (t0 <= w.maxHit && t1 <= w.maxHit) ? S3_PE_BOTH_RESCUE : S3_PE_BOTH_NO_PAIR_MANY;
    }
    const bool rescues = rt == S3_PE_BOTH_RESCUE || (rt == S3_PE_FIRST_RESCUES && !(r & 1)) || (rt == S3_PE_SECOND_RESCUES && (r & 1));
    uint32_t n = 0;
    const uint32_t base = FILL ? winOff[r] : 0u;
    if (rescues) {
        const uint32_t alignedLen = readLengths[r], mateLen = readLengths[r ^ 1u];
        for (uint32_t o = locOff[r]; o < locOff[r + 1]; ++o) {
            S3Window x[2];
            const int k = s3_win_half(w.w, r, occPos[o], occFlags[2 * (size_t)o], alignedLen, mateLen, x);
            if (FILL)
                for (int j = 0; j < k; ++j) {
                    const uint32_t t = base + n + j;
                    out.alignedOcc[t] = o; out.readID[t] = x[j].readID; out.strand[t] = x[j].strand; out.leftOrRight[t] = x[j].leftOrRight;
                    out.start[t] = x[j].start; out.dnaLen[t] = x[j].dnaLen; out.readLen[t] = x[j].readLen;
                    out.clipLt[t] = x[j].clipLt; out.clipRt[t] = x[j].clipRt; out.ancL[t] = x[j].ancL; out.ancR[t] = x[j].ancR;
                    out.cutoff[t] = x[j].cutoff;
                }
            n += (uint32_t)k;
        }
    }
    if (!FILL) winCount[r] = n;
    else if (!(r & 1)) routeFinal[p] = rt;
}

// ---- results: CIGAR runs (CigarStringEncoder DV-DPfunctions.h:545-597 as the engines' result loops drive it, ----------
// DV-DPfunctions.cu:2376-2390): the pattern is written right to left, 'V',c repeats the op before it c - 1 more times,
// neighbours of one type merge, runs whose length ends <= 0 separate their neighbours and are dropped; out in read order
// as length << 8 | op.  One thread per window; COUNT / FILL.
template <bool FILL>
__global__ void s3_pe_runs_kernel(uint32_t numWindows, const uint8_t *__restrict__ pattern, uint32_t patternLength,
                                  const int32_t *__restrict__ scores, const int32_t *__restrict__ cutoff,
                                  uint32_t *__restrict__ runCount, const uint32_t *__restrict__ runOff, uint32_t *__restrict__ runs)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numWindows) return;
    uint32_t kept = 0;
    if (scores[t] >= cutoff[t]) {
        const uint8_t *p = pattern + (size_t)t * patternLength, *end = p + patternLength;
        const uint32_t total = FILL ? runOff[t + 1] - runOff[t] : 0u;
        uint32_t *dst = FILL ? runs + runOff[t] : NULL;
        uint8_t last = 'N', curType = 0;
        int curCnt = 0;
        bool have = false;
        for (; p < end && *p != 0; ++p) {
            uint8_t type; int cnt;
            if (*p == 'V') { if (++p >= end) break; type = last; cnt = (int)*p - 1; }
            else { type = last = *p; cnt = 1; }
            if (have && curType == type) curCnt += cnt;
            else {
                if (have && curCnt > 0 && curType != 'N') { if (FILL) dst[total - 1 - kept] = ((uint32_t)curCnt << 8) | curType; ++kept; }
                curType = type; curCnt = cnt; have = true;
            }
        }
        if (have && curCnt > 0 && curType != 'N') { if (FILL) dst[total - 1 - kept] = ((uint32_t)curCnt << 8) | curType; ++kept; }
    }
    if (!FILL) runCount[t] = kept;
}

// one record per window: what DP_Space::algnmtCPUThread puts into AlgnmtDPResult (DV-DPfunctions.cu:2359-2420)
__global__ void s3_pe_result_kernel(uint32_t numWindows, const S3PeWindows win, const uint32_t *__restrict__ occPos,
                                    const uint8_t *__restrict__ occFlags, const int32_t *__restrict__ scores,
                                    const uint32_t *__restrict__ hit, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ runOff,
                                    s3_pe_dp_result *__restrict__ out)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numWindows) return;
    s3_pe_dp_result r;
    memset(&r, 0, sizeof r);
    const uint32_t o = win.alignedOcc[t];
    r.dpReadID = win.readID[t];
    r.alignedPos = occPos[o];
    r.alignedStrand = occFlags[2 * (size_t)o]; r.alignedMismatches = occFlags[2 * (size_t)o + 1];
    r.dpStrand = win.strand[t]; r.leftOrRight = win.leftOrRight[t];
    r.score = scores[t];
    const bool ok = scores[t] >= win.cutoff[t];
    r.dpPos = ok ? win.start[t] + hit[t] : 0xFFFFFFFFu;           // startLocs + hitLocs (:2392), none below the cutoff (:2412)
    r.numSameScore = cnt[t];
    r.runOffset = runOff[t]; r.numRuns = runOff[t + 1] - runOff[t];
    out[t] = r;
}

// =======================================================================================================================
static const uint32_t kAllowed[5] = {2, 4, 4, 2, 1};            // MAX_SA_RANGES_ALLOWED1_* (definitions.h:47-51)
static const uint32_t kCases[5] = {1, 2, 4, 6, 10};             // definitions.h:116-120

extern "C" int s3_pe_create(s3_index *ix, uint32_t maxReads, uint32_t maxReadLength, const s3_pe_params *params, s3_pe **out)
{
    if (!ix || !params || !out || maxReads == 0 || (maxReads & 1) || maxReadLength == 0) { s3_set_error("s3_pe_create: bad argument (maxReads must be even)"); return S3_EINVAL; }
    if (!ix->loc.sa || !ix->d_packedDNA) { s3_set_error("s3_pe_create: the index was uploaded without its suffix array and packed text"); return S3_EINVAL; }
    if (params->numMismatch > 4 || params->insertHigh < params->insertLow || params->insertLow < 0 ||
        (params->strandLeftLeg != 1 && params->strandLeftLeg != 2) || (params->strandRightLeg != 1 && params->strandRightLeg != 2) ||
        params->maxOutputPerRead == 0 || params->maxHitNumForDP == 0) { s3_set_error("s3_pe_create: bad parameters"); return S3_EINVAL; }
    S3_TRYC(cudaSetDevice(ix->device));
    s3_pe *pe = (s3_pe *)calloc(1, sizeof(s3_pe));
    if (!pe) { s3_set_error("out of host memory"); return S3_ENOMEM; }
    pe->ix = ix; pe->par = *params; pe->maxReads = maxReads;
    // the engines' sizes: maxReadLength = (L / 4 + 1) * 4, maxDNALength = insert_high - insert_low + L + 1
    // (DV-DPfunctions.cu:2223-2224 with L = inputMaxReadLength)
    pe->maxReadLength = (maxReadLength / 4 + 1) * 4;
    pe->maxDNALength = (uint32_t)(params->insertHigh - params->insertLow) + pe->maxReadLength + 1;
    pe->maxWindows = params->maxWindows ? params->maxWindows : maxReads;
    int rc = s3_dp_create(pe->maxReadLength, pe->maxDNALength, pe->maxWindows, params->scores, ix->device, &pe->dp);
    if (rc) { free(pe); return rc; }
    s3_dp_set_stream(pe->dp, ix->stream);
    if (cudaMallocHost(&pe->h_counts, 64 * sizeof(uint32_t)) != cudaSuccess) { s3_set_error("s3_pe_create: pinned allocation failed"); s3_pe_free(pe); return S3_ENOMEM; }
    for (int k = 0; k < 10; ++k) if (cudaEventCreate(&pe->ev[k]) != cudaSuccess) { s3_set_error("s3_pe_create: cudaEventCreate failed"); s3_pe_free(pe); return S3_ECUDA; }
    if (cudaStreamCreateWithFlags(&pe->copyStream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&pe->pfDone[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&pe->pfDone[1], cudaEventDisableTiming) != cudaSuccess) {
        s3_set_error("s3_pe_create: copy stream"); s3_pe_free(pe); return S3_ECUDA;
    }
    *out = pe;
    return S3_OK;
}

extern "C" void s3_pe_free(s3_pe *pe)
{
    if (!pe) return;
    cudaSetDevice(pe->ix->device);
    cudaStreamSynchronize(pe->ix->stream);
    if (pe->dp) s3_dp_free(pe->dp);
    if (pe->A.base) cudaFree(pe->A.base);
    if (pe->B.base) cudaFree(pe->B.base);
    if (pe->C.base) cudaFree(pe->C.base);
    if (pe->pinned) cudaFreeHost(pe->pinned);
    if (pe->h_counts) cudaFreeHost(pe->h_counts);
    for (int k = 0; k < 10; ++k) if (pe->ev[k]) cudaEventDestroy(pe->ev[k]);
    if (pe->copyStream) { cudaStreamSynchronize(pe->copyStream); cudaStreamDestroy(pe->copyStream); }
    for (int k = 0; k < 2; ++k) if (pe->pfDone[k]) cudaEventDestroy(pe->pfDone[k]);
    for (int k = 0; k < 2; ++k) if (pe->d_in[k]) cudaFree(pe->d_in[k]);
    free(pe);
}

extern "C" int s3_pe_set_timing(s3_pe *pe, int on) { if (!pe) return S3_EINVAL; pe->timing = on ? 1 : 0; return S3_OK; }
extern "C" int s3_pe_read_timing(s3_pe *pe, float *msPerStage)
{
    if (!pe || !msPerStage) return S3_EINVAL;
    for (int k = 0; k < 8; ++k) msPerStage[k] = pe->msStages[k];
    return S3_OK;
}
extern "C" s3_dp *s3_pe_dp(s3_pe *pe) { return pe ? pe->dp : NULL; }

static int pe_pinned(s3_pe *pe, size_t bytes)
{
    if (bytes <= pe->pinnedBytes) return S3_OK;
    if (pe->pinned) { cudaFreeHost(pe->pinned); pe->pinned = NULL; pe->pinnedBytes = 0; }
    bytes += bytes / 4;
    if (cudaMallocHost(&pe->pinned, bytes) != cudaSuccess) { s3_set_error("s3_pe_align: pinned allocation of %zu bytes failed", bytes); return S3_ENOMEM; }
    pe->pinnedBytes = bytes;
    return S3_OK;
}

#define PE_MARK(k) do { if (pe->timing) S3_TRYC(cudaEventRecord(pe->ev[k], st)); } while (0)

// queriesOnDevice: queries / readLengths are device pointers and the result arrays stay on the device (nothing but the
// counts crosses the link): the device-resident measurement of bench.py
static int pe_run(s3_pe *pe, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads64, uint32_t wordPerQuery,
                  int queriesOnDevice, s3_pe_result *res)
{
    if (!pe || !queries || !readLengths || !res) { s3_set_error("s3_pe_align: NULL argument"); return S3_EINVAL; }
    memset(res, 0, sizeof *res);
    if (numReads64 > pe->maxReads || (numReads64 & 1)) { s3_set_error("s3_pe_align: %llu reads (even, <= %u)", (unsigned long long)numReads64, pe->maxReads); return S3_EINVAL; }
    const uint32_t N = (uint32_t)numReads64, P = N / 2;
    if (N == 0) return S3_OK;
    s3_index *ix = pe->ix;
    S3_TRYC(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const uint32_t k = pe->par.numMismatch, C = kCases[k], allowed = kAllowed[k], wpa = 2 * allowed;
    const size_t up = ((size_t)N + 31) / 32 * 32, maxRanges = (size_t)N * C * allowed;
    int rc;

    // ---- stage 1 buffers
    size_t scanTemp = 0, t2 = 0;
    cub::DeviceScan::ExclusiveSum(NULL, scanTemp, (uint32_t *)NULL, (uint32_t *)NULL, (int)(N + 1), st);
    size_t needA = arena_need(up * wordPerQuery, 4) + arena_need(up, 4) + C * arena_need(up * wpa, 4) + 6 * arena_need(N + 1, 4) + arena_need(N, 1) +
                   2 * arena_need(maxRanges, 4) + arena_need(maxRanges, 2) + arena_need(maxRanges, 1) + 2 * arena_need(P, 1) + arena_need(P, sizeof(S3PeBest)) +
                   arena_need(scanTemp, 1) + arena_need(64, 4) + (pe->par.readStats ? arena_need(N, sizeof(s3_pe_read_stats)) : 0) + 4096;
    if ((rc = arena_reserve(&pe->A, needA, st))) return rc;
    S3Arena *A = &pe->A;
    uint32_t *d_q = const_cast<uint32_t *>(queries), *d_len = const_cast<uint32_t *>(readLengths);
    bool uploaded = false;
    if (!queriesOnDevice) {
        const size_t inBytes = up * wordPerQuery * 4 + up * 4;
        int buf = -1;
        for (int b2 = 0; b2 < 2; ++b2)
            if (pe->pf[b2].queries == queries && pe->pf[b2].reads == numReads64 && pe->pf[b2].wpq == wordPerQuery) buf = b2;
        if (buf >= 0) {
            uploaded = true;
            S3_TRYC(cudaStreamWaitEvent(st, pe->pfDone[buf], 0));
            pe->pf[buf].queries = NULL;
        } else {
            buf = pe->pf[0].queries ? 1 : 0;                     // a buffer no prefetched batch is waiting in
            if (pe->pf[buf].queries) { S3_TRYC(cudaStreamSynchronize(pe->copyStream)); pe->pf[buf].queries = NULL; }     // both taken: drop one
            if ((rc = pe_input_buffer(pe, buf, inBytes, st))) return rc;
        }
        d_q = pe->d_in[buf]; d_len = d_q + up * wordPerQuery;
    }
    uint32_t *d_ans[S3_MAX_NUM_CASES];
    for (uint32_t c = 0; c < C; ++c) d_ans[c] = arena_take<uint32_t>(A, up * wpa);
    uint32_t *d_nRanges = arena_take<uint32_t>(A, N + 1), *d_rangeOff = arena_take<uint32_t>(A, N + 1), *d_totOcc = arena_take<uint32_t>(A, N + 1);
    uint32_t *d_locCount = arena_take<uint32_t>(A, N + 1), *d_locOff = arena_take<uint32_t>(A, N + 1), *d_winCount = arena_take<uint32_t>(A, N + 1);
    uint8_t *d_readFlags = arena_take<uint8_t>(A, N);
    uint32_t *d_saL = arena_take<uint32_t>(A, maxRanges), *d_saR = arena_take<uint32_t>(A, maxRanges);
    uint8_t *d_saFlags = arena_take<uint8_t>(A, 2 * maxRanges), *d_keep = arena_take<uint8_t>(A, maxRanges);
    uint8_t *d_route = arena_take<uint8_t>(A, P), *d_routeFinal = arena_take<uint8_t>(A, P);
    S3PeBest *d_best = arena_take<S3PeBest>(A, P);
    void *d_tmp = arena_take<char>(A, scanTemp);
    uint32_t *d_counters = arena_take<uint32_t>(A, 64);
    s3_pe_read_stats *d_stats = pe->par.readStats ? arena_take<s3_pe_read_stats>(A, N) : NULL;
    if (!d_counters || (pe->par.readStats && !d_stats)) { s3_set_error("s3_pe_align: stage buffer accounting"); return S3_ENOMEM; }

    PE_MARK(0);
    if (!queriesOnDevice && !uploaded) {
        S3_TRYC(cudaMemcpyAsync(d_q, queries, up * wordPerQuery * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_len, readLengths, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    }
    if ((rc = s3_search_round1_device(ix, d_q, d_len, N, wordPerQuery, k, C, allowed, wpa, 0, d_ans, NULL))) return rc;
    PE_MARK(1);
    S3Collect col;
    memset(&col, 0, sizeof col);
    for (uint32_t c = 0; c < C; ++c) col.answers[c] = d_ans[c];
    col.numCases = C; col.allowed = allowed; col.wordPerAns = wpa; col.numReads = N; col.textLength = ix->textLength;
    col.maxOutputPerRead = pe->par.maxOutputPerRead;
    const unsigned nbR = (N + 255) / 256, nbP = (P + 255) / 256;
    S3_TRYC(cudaMemsetAsync(d_nRanges + N, 0, 4, st));
    S3_TRYC(cudaMemsetAsync(d_counters, 0, 64 * 4, st));
    s3_pe_collect_kernel<false><<<nbR, 256, 0, st>>>(col, d_nRanges, d_totOcc, d_readFlags, NULL, NULL, NULL, NULL, d_stats);
    S3_TRYC(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_nRanges, d_rangeOff, (int)(N + 1), st));
    s3_pe_collect_kernel<true><<<nbR, 256, 0, st>>>(col, NULL, NULL, NULL, d_rangeOff, d_saL, d_saR, d_saFlags);
    S3_TRYC(cudaMemsetAsync(d_keep, 1, maxRanges, st));
    S3_TRYC(cudaMemsetAsync(d_locCount + N, 0, 4, st));
    s3_pe_route_kernel<<<nbP, 256, 0, st>>>(P, d_rangeOff, d_saL, d_saR, d_saFlags, d_totOcc, d_readFlags, pe->par.maxHitNumForDP,
                                           pe->par.keepSecondBest, d_route, d_keep, d_locCount, d_counters);
    S3_TRYC(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_locCount, d_locOff, (int)(N + 1), st));
    S3_LAUNCHED(3);
    S3_TRYC(cudaGetLastError());
    // first read of counts: occurrences to locate
    S3_TRYC(cudaMemcpyAsync(pe->h_counts, d_locOff + N, 4, cudaMemcpyDeviceToHost, st));
    S3_TRYC(cudaMemcpyAsync(pe->h_counts + 1, d_rangeOff + N, 4, cudaMemcpyDeviceToHost, st));
    S3_TRYC(cudaStreamSynchronize(st));
    const uint32_t T = pe->h_counts[0];
    res->numRanges = pe->h_counts[1];
    PE_MARK(2);

    // ---- stage 2: locate, pairing, windows
    const size_t Tm = T ? T : 1;
    cub::DeviceRadixSort::SortPairs(NULL, t2, (unsigned long long *)NULL, (unsigned long long *)NULL, (uint32_t *)NULL, (uint32_t *)NULL, (int)Tm, 0, 64, st);
    size_t needB = arena_need(Tm, 4) + arena_need(Tm, 2) + 2 * arena_need(Tm, 8) + 2 * arena_need(Tm, 4) + arena_need(t2, 1) + arena_need(N + 1, 4) + arena_need(Tm, 2) + arena_need(P + 1, 4) + 4096;
    if ((rc = arena_reserve(&pe->B, needB, st))) return rc;
    S3Arena *B = &pe->B;
    uint32_t *d_occPos = arena_take<uint32_t>(B, Tm);
    uint8_t *d_occFlags = arena_take<uint8_t>(B, 2 * Tm);
    unsigned long long *d_keyA = arena_take<unsigned long long>(B, Tm), *d_keyB = arena_take<unsigned long long>(B, Tm);
    uint32_t *d_valA = arena_take<uint32_t>(B, Tm), *d_valB = arena_take<uint32_t>(B, Tm);
    void *d_tmp2 = arena_take<char>(B, t2);
    uint32_t *d_winOff = arena_take<uint32_t>(B, N + 1);
    uint16_t *d_sflags = arena_take<uint16_t>(B, Tm);
    uint32_t *d_heavy = arena_take<uint32_t>(B, P + 1);
    if (!d_heavy) { s3_set_error("s3_pe_align: stage buffer accounting"); return S3_ENOMEM; }
    if (T) {
        s3_pe_locate_kernel<<<nbR, 256, 0, st>>>(N, ix->loc.sa, d_rangeOff, d_saL, d_saR, d_saFlags, d_keep, d_locOff, d_occPos, d_occFlags, d_keyA, d_valA);
        // ordered by (read, position), ties in arrival order: what PERadixSort leaves of each list (PEAlgnmt.cpp:114-199)
        int endBit = 33;
        while (endBit < 64 && (1ull << (endBit - 32)) < (unsigned long long)N) ++endBit;
        S3_TRYC(cub::DeviceRadixSort::SortPairs(d_tmp2, t2, d_keyA, d_keyB, d_valA, d_valB, (int)T, 0, endBit, st));
        s3_pe_sorted_flags_kernel<<<(T + 255) / 256, 256, 0, st>>>(T, d_valB, d_occFlags, d_sflags);
        S3_LAUNCHED(2);
    }
    PE_MARK(3);
    S3PairParams pp = {(uint32_t)pe->par.insertLow, (uint32_t)pe->par.insertHigh, pe->par.strandLeftLeg, pe->par.strandRightLeg, 0};
    // ordinary pairs a warp each; the few with hundreds of occurrences (tandem repeats) a block each afterwards
    S3_TRYC(cudaMemsetAsync(d_heavy + P, 0, 4, st));
    s3_pe_pair_kernel<false><<<(P + S3_PE_PAIR_WARPS - 1) / S3_PE_PAIR_WARPS, S3_PE_PAIR_WARPS * 32, 0, st>>>(P, d_route, d_locOff, d_keyB, d_sflags, d_len, pp, d_best, d_heavy, d_heavy + P);
    s3_pe_pair_kernel<true><<<(unsigned)(Tm / S3_PE_HEAVY_ELEMENTS + 1), S3_PE_PAIR_WARPS * 32, 0, st>>>(P, d_route, d_locOff, d_keyB, d_sflags, d_len, pp, d_best, d_heavy, d_heavy + P);
    S3PeWinParams wp;
    wp.w.insertLow = pe->par.insertLow; wp.w.insertHigh = pe->par.insertHigh; wp.w.leftLeg = pe->par.strandLeftLeg; wp.w.rightLeg = pe->par.strandRightLeg;
    wp.w.softClipLeft = pe->par.softClipLeft; wp.w.softClipRight = pe->par.softClipRight;
    wp.w.cutoff[0] = wp.w.cutoff[1] = pe->par.cutoffThreshold; wp.w.maxDNALength = pe->maxDNALength; wp.w.textLength = ix->textLength;
    wp.maxHit = pe->par.maxHitNumForDP;
    S3PeWindows win;
    memset(&win, 0, sizeof win);
    S3_TRYC(cudaMemsetAsync(d_winCount + N, 0, 4, st));
    s3_pe_window_kernel<false><<<nbR, 256, 0, st>>>(N, d_route, NULL, d_best, d_locOff, d_occPos, d_occFlags, d_len, wp, d_winCount, NULL, win);
    S3_TRYC(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_winCount, d_winOff, (int)(N + 1), st));
    S3_LAUNCHED(2);
    S3_TRYC(cudaGetLastError());
    // second read of counts: windows
    S3_TRYC(cudaMemcpyAsync(pe->h_counts + 2, d_winOff + N, 4, cudaMemcpyDeviceToHost, st));
    S3_TRYC(cudaStreamSynchronize(st));
    const uint32_t M = pe->h_counts[2];
    PE_MARK(4);
    if (M > pe->maxWindows) { s3_set_error("s3_pe_align: %u rescue windows > maxWindows %u of s3_pe_create", M, pe->maxWindows); return S3_EINVAL; }

    // ---- stage 3: windows, DP, results
    const size_t Mm = M ? M : 1, patLen = (size_t)pe->maxReadLength + pe->maxDNALength;
    size_t needC = 9 * arena_need(Mm, 4) + 2 * arena_need(Mm, 1) + 3 * arena_need(Mm, 4) + arena_need(Mm * patLen, 1) + 2 * arena_need(Mm + 1, 4) +
                   arena_need(Mm, sizeof(s3_pe_dp_result)) + arena_need(Mm * (size_t)(pe->maxReadLength + 8), 4) + 4096;
    if ((rc = arena_reserve(&pe->C, needC, st))) return rc;
    S3Arena *Cc = &pe->C;
    win.alignedOcc = arena_take<uint32_t>(Cc, Mm); win.readID = arena_take<uint32_t>(Cc, Mm);
    win.start = arena_take<uint32_t>(Cc, Mm); win.dnaLen = arena_take<uint32_t>(Cc, Mm); win.readLen = arena_take<uint32_t>(Cc, Mm);
    win.clipLt = arena_take<uint32_t>(Cc, Mm); win.clipRt = arena_take<uint32_t>(Cc, Mm); win.ancL = arena_take<uint32_t>(Cc, Mm); win.ancR = arena_take<uint32_t>(Cc, Mm);
    win.strand = arena_take<uint8_t>(Cc, Mm); win.leftOrRight = arena_take<uint8_t>(Cc, Mm);
    win.cutoff = arena_take<int32_t>(Cc, Mm);
    int32_t *d_score = arena_take<int32_t>(Cc, Mm);
    uint32_t *d_hit = arena_take<uint32_t>(Cc, Mm), *d_cnt = arena_take<uint32_t>(Cc, Mm);
    uint8_t *d_pattern = arena_take<uint8_t>(Cc, Mm * patLen);
    uint32_t *d_runCount = arena_take<uint32_t>(Cc, Mm + 1), *d_runOff = arena_take<uint32_t>(Cc, Mm + 1);
    s3_pe_dp_result *d_res = arena_take<s3_pe_dp_result>(Cc, Mm);
    uint32_t *d_runs = arena_take<uint32_t>(Cc, Mm * (size_t)(pe->maxReadLength + 8));
    if (!d_runs) { s3_set_error("s3_pe_align: stage buffer accounting"); return S3_ENOMEM; }
    // the fill pass also writes the final route codes, so it runs even without windows
    s3_pe_window_kernel<true><<<nbR, 256, 0, st>>>(N, d_route, d_routeFinal, d_best, d_locOff, d_occPos, d_occFlags, d_len, wp, NULL, d_winOff, win);
    S3_LAUNCHED(1);
    PE_MARK(5);
    uint32_t totalRuns = 0;
    if (M) {
        if ((rc = s3_dp_align_windows_device(pe->dp, ix, d_q, wordPerQuery, win.readID, win.strand, win.start, win.dnaLen, win.readLen, win.cutoff,
                                             d_score, d_hit, d_cnt, d_pattern, M, win.clipLt, win.clipRt, win.ancL, win.ancR))) return rc;
        PE_MARK(6);
        const unsigned nbM = (M + 127) / 128;
        size_t scanM = 0;
        cub::DeviceScan::ExclusiveSum(NULL, scanM, (uint32_t *)NULL, (uint32_t *)NULL, (int)(M + 1), st);
        if (scanM > scanTemp) { s3_set_error("s3_pe_align: scan scratch"); return S3_ENOMEM; }
        S3_TRYC(cudaMemsetAsync(d_runCount + M, 0, 4, st));
        s3_pe_runs_kernel<false><<<nbM, 128, 0, st>>>(M, d_pattern, (uint32_t)patLen, d_score, win.cutoff, d_runCount, NULL, NULL);
        S3_TRYC(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_runCount, d_runOff, (int)(M + 1), st));
        s3_pe_runs_kernel<true><<<nbM, 128, 0, st>>>(M, d_pattern, (uint32_t)patLen, d_score, win.cutoff, NULL, d_runOff, d_runs);
        s3_pe_result_kernel<<<nbM, 128, 0, st>>>(M, win, d_occPos, d_occFlags, d_score, d_hit, d_cnt, d_runOff, d_res);
        S3_LAUNCHED(3);
        S3_TRYC(cudaGetLastError());
        S3_TRYC(cudaMemcpyAsync(pe->h_counts + 3, d_runOff + M, 4, cudaMemcpyDeviceToHost, st));
    } else PE_MARK(6);
    S3_TRYC(cudaMemcpyAsync(pe->h_counts + 8, d_counters, 16 * 4, cudaMemcpyDeviceToHost, st));
    PE_MARK(7);
    if (queriesOnDevice) {
        S3_TRYC(cudaStreamSynchronize(st));
        totalRuns = M ? pe->h_counts[3] : 0;
        res->d_route = d_routeFinal; res->d_pairs = (s3_pe_pair_result *)d_best; res->d_dp = d_res; res->d_runs = d_runs; res->d_readStats = d_stats;
    } else {
        // the runs' total is needed to size their copy: everything else goes first
        const size_t statBytes = d_stats ? ((size_t)N * sizeof(s3_pe_read_stats) + 255) / 256 * 256 : 0;
        const size_t bytes = (size_t)P + 256 + (size_t)P * sizeof(S3PeBest) + 256 + Mm * sizeof(s3_pe_dp_result) + 256 + statBytes + Mm * (size_t)(pe->maxReadLength + 8) * 4;
        if ((rc = pe_pinned(pe, bytes))) return rc;
        char *h = (char *)pe->pinned;
        res->route = (uint8_t *)h; h += ((size_t)P + 255) / 256 * 256;
        res->pairs = (s3_pe_pair_result *)h; h += ((size_t)P * sizeof(S3PeBest) + 255) / 256 * 256;
        res->dp = (s3_pe_dp_result *)h; h += (Mm * sizeof(s3_pe_dp_result) + 255) / 256 * 256;
        if (d_stats) { res->readStats = (s3_pe_read_stats *)h; h += statBytes; }
        res->runs = (uint32_t *)h;
        S3_TRYC(cudaMemcpyAsync(res->route, d_routeFinal, P, cudaMemcpyDeviceToHost, st));
        S3_TRYC(cudaMemcpyAsync(res->pairs, d_best, (size_t)P * sizeof(S3PeBest), cudaMemcpyDeviceToHost, st));
        if (M) S3_TRYC(cudaMemcpyAsync(res->dp, d_res, (size_t)M * sizeof(s3_pe_dp_result), cudaMemcpyDeviceToHost, st));
        if (d_stats) S3_TRYC(cudaMemcpyAsync(res->readStats, d_stats, (size_t)N * sizeof(s3_pe_read_stats), cudaMemcpyDeviceToHost, st));
        S3_TRYC(cudaStreamSynchronize(st));
        totalRuns = M ? pe->h_counts[3] : 0;
        if (totalRuns) {
            S3_TRYC(cudaMemcpyAsync(res->runs, d_runs, (size_t)totalRuns * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYC(cudaStreamSynchronize(st));
        }
        res->h2dBytes = up * wordPerQuery * 4 + (size_t)N * 4;
        res->d2hBytes = (size_t)P + (size_t)P * sizeof(S3PeBest) + (size_t)M * sizeof(s3_pe_dp_result) + (size_t)totalRuns * 4 + 20 * 4 + (d_stats ? (size_t)N * sizeof(s3_pe_read_stats) : 0);
    }
    res->numPairs = P; res->numOccurrences = T; res->numWindows = M; res->numRuns = totalRuns;
    for (int c = 0; c < 16; ++c) res->routeCounts[c] = pe->h_counts[8 + c];
    pe->last.d_q = d_q; pe->last.d_len = d_len; pe->last.d_route = d_routeFinal; pe->last.N = N; pe->last.wpq = wordPerQuery;
    if (pe->timing) {
        for (int s = 0; s < 7; ++s) { float ms = 0; cudaEventElapsedTime(&ms, pe->ev[s], pe->ev[s + 1]); pe->msStages[s] += ms; }
    }
    return S3_OK;
}

extern "C" int s3_pe_prefetch(s3_pe *pe, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery)
{
    if (!pe || !queries || !readLengths) { s3_set_error("s3_pe_prefetch: NULL argument"); return S3_EINVAL; }
    if (numReads > pe->maxReads || (numReads & 1) || numReads == 0) { s3_set_error("s3_pe_prefetch: %llu reads (even, 1..%u)", (unsigned long long)numReads, pe->maxReads); return S3_EINVAL; }
    S3_TRYC(cudaSetDevice(pe->ix->device));
    const size_t up = ((size_t)numReads + 31) / 32 * 32;
    // a buffer no prefetched batch is waiting in (s3_pe_align is synchronous, so none is in use by a running call); with both
    // taken the call does nothing and that batch is uploaded by its s3_pe_align
    if (pe->pf[0].queries && pe->pf[1].queries) return S3_OK;
    const int buf = pe->pf[0].queries ? 1 : 0;
    int rc;
    if ((rc = pe_input_buffer(pe, buf, up * wordPerQuery * 4 + up * 4, pe->copyStream))) return rc;
    uint32_t *d_q = pe->d_in[buf];
    S3_TRYC(cudaMemcpyAsync(d_q, queries, up * wordPerQuery * 4, cudaMemcpyHostToDevice, pe->copyStream));
    S3_TRYC(cudaMemcpyAsync(d_q + up * wordPerQuery, readLengths, (size_t)numReads * 4, cudaMemcpyHostToDevice, pe->copyStream));
    S3_TRYC(cudaEventRecord(pe->pfDone[buf], pe->copyStream));
    pe->pf[buf].queries = queries; pe->pf[buf].reads = numReads; pe->pf[buf].wpq = wordPerQuery;
    return S3_OK;
}

extern "C" int s3_pe_align(s3_pe *pe, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery, s3_pe_result *out)
{
    return pe_run(pe, queries, readLengths, numReads, wordPerQuery, 0, out);
}
extern "C" int s3_pe_align_device(s3_pe *pe, const uint32_t *d_queries, const uint32_t *d_readLengths, uint64_t numReads, uint32_t wordPerQuery, s3_pe_result *out)
{
    return pe_run(pe, d_queries, d_readLengths, numReads, wordPerQuery, 1, out);
}

// DPForUnalignPairs2 for the both-unaligned pairs of the batch this handle has just aligned (s3_stages.cu)
int s3_stage_deep_dp_of_chain(s3_index *ix, const uint32_t *d_queries, const uint32_t *d_len, uint64_t numReads, uint32_t wordPerQuery, const uint8_t *d_route,
                              uint8_t wantRoute, const s3_stage_params *par, s3_deep_dp_result *out);
extern "C" int s3_pe_deep_dp(s3_pe *pe, const s3_stage_params *par, s3_deep_dp_result *out)
{
    if (!pe || !par || !out) { s3_set_error("s3_pe_deep_dp: NULL argument"); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    if (!pe->last.d_q) { s3_set_error("s3_pe_deep_dp: no batch has been aligned on this handle"); return S3_EINVAL; }
    return s3_stage_deep_dp_of_chain(pe->ix, pe->last.d_q, pe->last.d_len, pe->last.N, pe->last.wpq, pe->last.d_route, S3_PE_NONE, par, out);
}


// =======================================================================================================================
// Single-end batch: search -> collect -> (best hits) -> locate.  The in-memory alignSingleR (soap3-dp-module.cu:62-181:
// soap3_dp_single_align with outputFileName == NULL, hostKernel storing occRec records, CPUfunctions.cpp:1887-1905): per
// read its occurrences in the order collect_all_answers + transferAllSAToOcc leave them (cases ascending, slots in order,
// suffix-array order inside a range), capped at MaxOutputPerRead; reportBest keeps the ranges with the fewest mismatches
// only (retainAllBest, SAList.cpp:140-207, the all-best report type).  Reads whose round-1 slot overflowed in some case
// are flagged (readFlags bit 0), like S3_PE_OVERFLOW.
// =======================================================================================================================
struct s3_se {
    s3_index *ix;
    s3_se_params par;
    uint32_t maxReads;
    S3Arena A, B;
    void *pinned; size_t pinnedBytes;
    uint32_t *h_counts;
};

// one thread per read: which ranges stay, how many occurrences they hold
__global__ void s3_se_select_kernel(uint32_t numReads, const uint32_t *__restrict__ rangeOff, const uint32_t *__restrict__ saL,
                                    const uint32_t *__restrict__ saR, const uint8_t *__restrict__ saFlags, const uint32_t *__restrict__ totOcc,
                                    int reportBest, uint8_t *__restrict__ keepRange, uint32_t *__restrict__ locCount)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= numReads) return;
    uint32_t tot = totOcc[r];
    if (reportBest && tot) {
        int mn = 999;
        for (uint32_t g = rangeOff[r]; g < rangeOff[r + 1]; ++g) mn = min(mn, (int)saFlags[2 * (size_t)g + 1]);
        tot = 0;
        for (uint32_t g = rangeOff[r]; g < rangeOff[r + 1]; ++g) {
            const bool keep = (int)saFlags[2 * (size_t)g + 1] == mn;
            keepRange[g] = keep ? 1 : 0;
            if (keep) tot += saR[g] - saL[g] + 1;
        }
    }
    locCount[r] = tot;
}

__global__ void s3_se_seed_length_kernel(uint32_t up, uint32_t numReads, const uint32_t *__restrict__ readLengths, uint32_t *__restrict__ seedLengths)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= up) return;
    const uint32_t len = r < numReads ? readLengths[r] : 0u;
    seedLengths[r] = len > S3_LONG_READ_LEN ? S3_LONG_READ_SEED_LEN : len;
}

// ---- long reads: the seed alignments of reads longer than 120 bases are extended without gaps over the rest of the read ------
// validateAlignments (CPUfunctions.cpp:1129-1222) as hostKernel calls it (:1812-1842): the search aligned the first 100 bases
// (alignment.cu:2475-2491); an occurrence stays if its seed mismatches + the Hamming distance of the other readLen - 100 bases
// (PE.cpp:28-60, 148-178, 287-325) is within ceil(0.02 * readLen) (twice that when MAPQ is wanted); a reverse-strand occurrence
// moves to the start of the whole read.  onlyKeepBest keeps the running best only (an earlier, worse entry is dropped when a
// better one turns up: the output size resets), the walk stops once maxHitNum entries of minSeedMismatch total are out, and the
// list is cut to maxHitNum.  One thread per read, in place: the kept entries of read r move to the front of its list.
__device__ __forceinline__ uint32_t s3_val_text(const uint32_t *__restrict__ text, uint64_t p) { return (__ldg(text + (p >> 4)) >> (30u - 2u * ((uint32_t)p & 15u))) & 3u; }
__device__ __forceinline__ uint32_t s3_val_read(const uint32_t *__restrict__ q, uint32_t wpq, uint32_t r, uint32_t k)
{
    return (__ldg(q + (size_t)(r / 32) * 32 * wpq + (size_t)(k >> 4) * 32 + r % 32) >> ((k & 15u) << 1)) & 3u;
}

__global__ void s3_validate_kernel(uint32_t numReads, const uint32_t *__restrict__ queries, uint32_t wpq, const uint32_t *__restrict__ readLengths,
                                   const uint32_t *__restrict__ text, uint32_t textLength, const uint32_t *__restrict__ off, uint32_t *__restrict__ pos,
                                   uint8_t *__restrict__ flags, int onlyKeepBest, int minSeedMismatch, int doubleAllowance, int maxHitNum,
                                   uint32_t *__restrict__ outCount)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= numReads) return;
    const uint32_t a = off[r], n = off[r + 1] - a, len = readLengths[r], seed = len > S3_LONG_READ_LEN ? S3_LONG_READ_SEED_LEN : len;
    uint32_t m = n;
    if (n && len > seed) {
        const uint32_t ext = len - seed;
        int maxMismatch = (int)((len + 49u) / 50u);                                   // ceil(0.02 * len)
        if (doubleAllowance) maxMismatch *= 2;
        int pre = maxMismatch;
        m = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t p = pos[a + i], st = flags[2 * (size_t)(a + i)];
            const int mm = flags[2 * (size_t)(a + i) + 1];
            if (mm < minSeedMismatch) continue;
            int mismatch = maxMismatch + 1;
            if (st == 1 && (uint64_t)p + len <= textLength) {
                mismatch = 0;
                for (uint32_t j = 0; j < ext && mismatch + mm <= maxMismatch; ++j) mismatch += s3_val_text(text, (uint64_t)p + seed + j) != s3_val_read(queries, wpq, r, seed + j);
            }
            if (st == 2 && p >= ext) {
                mismatch = 0;
                for (uint32_t j = 0; j < ext && mismatch + mm <= maxMismatch; ++j) mismatch += s3_val_text(text, (uint64_t)p - ext + j) != 3u - s3_val_read(queries, wpq, r, len - 1 - j);
            }
            const int tot = mismatch + mm;
            if (tot > maxMismatch) continue;
            if (onlyKeepBest && tot > pre) continue;
            if (onlyKeepBest && tot < pre) { m = 0; pre = tot; }
            pos[a + m] = st == 1 ? p : p - ext;
            flags[2 * (size_t)(a + m)] = (uint8_t)st;
            flags[2 * (size_t)(a + m) + 1] = (uint8_t)tot;
            ++m;
            if (pre == minSeedMismatch && (int)m >= maxHitNum) break;
        }
    }
    if (n && (int)m > maxHitNum) m = (uint32_t)maxHitNum;
    outCount[r] = m;
}

// kept entries of every read, gathered into lists of their own (offsets = exclusive scan of the counts)
__global__ void s3_validate_gather_kernel(uint32_t numReads, const uint32_t *__restrict__ off, const uint32_t *__restrict__ count, const uint32_t *__restrict__ newOff,
                                          const uint32_t *__restrict__ pos, const uint8_t *__restrict__ flags, uint32_t *__restrict__ outPos, uint8_t *__restrict__ outFlags)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= numReads) return;
    const uint32_t a = off[r], b = newOff[r], m = count[r];
    for (uint32_t i = 0; i < m; ++i) {
        outPos[b + i] = pos[a + i];
        outFlags[2 * (size_t)(b + i)] = flags[2 * (size_t)(a + i)];
        outFlags[2 * (size_t)(b + i) + 1] = flags[2 * (size_t)(a + i) + 1];
    }
}

extern "C" int s3_validate_alignments(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery,
                                      const uint32_t *occOffsets, uint32_t *positions, uint8_t *occFlags, int onlyKeepBest, int minSeedMismatch,
                                      int doubleAllowance, int maxHitNum, uint32_t *outCounts)
{
    if (!ix || !queries || !readLengths || !occOffsets || !positions || !occFlags || !outCounts || maxHitNum < 1) { s3_set_error("s3_validate_alignments: bad argument"); return S3_EINVAL; }
    if (!ix->loc.text) { s3_set_error("s3_validate_alignments: the index was uploaded without its packed text"); return S3_EINVAL; }
    if (numReads == 0) return S3_OK;
    if (numReads > 0xFFFFFFF0ull) { s3_set_error("s3_validate_alignments: too many reads"); return S3_EINVAL; }
    const uint32_t N = (uint32_t)numReads, T = occOffsets[N];
    for (uint32_t r = 0; r < N; ++r) {
        if (occOffsets[r] > occOffsets[r + 1]) { s3_set_error("s3_validate_alignments: offsets not ascending at read %u", r); return S3_EINVAL; }
        if (readLengths[r] > 16u * wordPerQuery) { s3_set_error("s3_validate_alignments: read %u longer than its query words", r); return S3_EINVAL; }
    }
    S3_TRYC(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const size_t up = ((size_t)N + 31) / 32 * 32, Tm = T ? T : 1;
    uint32_t *d_q = NULL, *d_len = NULL, *d_off = NULL, *d_pos = NULL, *d_cnt = NULL;
    uint8_t *d_fl = NULL;
    int rc = S3_OK;
#define S3_VAL(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess && rc == S3_OK) { s3_set_error("s3_validate_alignments: %s", cudaGetErrorString(e_)); rc = S3_ECUDA; } } while (0)
    S3_VAL(cudaMalloc(&d_q, up * wordPerQuery * 4)); S3_VAL(cudaMalloc(&d_len, (size_t)N * 4)); S3_VAL(cudaMalloc(&d_off, ((size_t)N + 1) * 4));
    S3_VAL(cudaMalloc(&d_pos, Tm * 4)); S3_VAL(cudaMalloc(&d_fl, 2 * Tm)); S3_VAL(cudaMalloc(&d_cnt, (size_t)N * 4));
    if (rc == S3_OK) {
        S3_VAL(cudaMemcpyAsync(d_q, queries, up * wordPerQuery * 4, cudaMemcpyHostToDevice, st));
        S3_VAL(cudaMemcpyAsync(d_len, readLengths, (size_t)N * 4, cudaMemcpyHostToDevice, st));
        S3_VAL(cudaMemcpyAsync(d_off, occOffsets, ((size_t)N + 1) * 4, cudaMemcpyHostToDevice, st));
        if (T) { S3_VAL(cudaMemcpyAsync(d_pos, positions, (size_t)T * 4, cudaMemcpyHostToDevice, st)); S3_VAL(cudaMemcpyAsync(d_fl, occFlags, 2 * (size_t)T, cudaMemcpyHostToDevice, st)); }
    }
    if (rc == S3_OK) {
        s3_validate_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, d_q, wordPerQuery, d_len, ix->loc.text, ix->textLength, d_off, d_pos, d_fl, onlyKeepBest, minSeedMismatch,
                                                             doubleAllowance, maxHitNum, d_cnt);
        S3_LAUNCHED(1);
        S3_VAL(cudaGetLastError());
        if (T) { S3_VAL(cudaMemcpyAsync(positions, d_pos, (size_t)T * 4, cudaMemcpyDeviceToHost, st)); S3_VAL(cudaMemcpyAsync(occFlags, d_fl, 2 * (size_t)T, cudaMemcpyDeviceToHost, st)); }
        S3_VAL(cudaMemcpyAsync(outCounts, d_cnt, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
        S3_VAL(cudaStreamSynchronize(st));
    }
#undef S3_VAL
    cudaFree(d_q); cudaFree(d_len); cudaFree(d_off); cudaFree(d_pos); cudaFree(d_fl); cudaFree(d_cnt);
    return rc;
}

extern "C" int s3_se_create(s3_index *ix, uint32_t maxReads, const s3_se_params *params, s3_se **out)
{
    if (!ix || !params || !out || maxReads == 0) { s3_set_error("s3_se_create: bad argument"); return S3_EINVAL; }
    if (!ix->loc.sa) { s3_set_error("s3_se_create: the index was uploaded without its suffix array"); return S3_EINVAL; }
    if (params->numMismatch > 4 || params->maxOutputPerRead == 0) { s3_set_error("s3_se_create: bad parameters"); return S3_EINVAL; }
    S3_TRYC(cudaSetDevice(ix->device));
    s3_se *se = (s3_se *)calloc(1, sizeof(s3_se));
    if (!se) { s3_set_error("out of host memory"); return S3_ENOMEM; }
    se->ix = ix; se->par = *params; se->maxReads = maxReads;
    if (cudaMallocHost(&se->h_counts, 16 * sizeof(uint32_t)) != cudaSuccess) { s3_set_error("s3_se_create: pinned allocation failed"); free(se); return S3_ENOMEM; }
    *out = se;
    return S3_OK;
}

extern "C" void s3_se_free(s3_se *se)
{
    if (!se) return;
    cudaSetDevice(se->ix->device);
    cudaStreamSynchronize(se->ix->stream);
    if (se->A.base) cudaFree(se->A.base);
    if (se->B.base) cudaFree(se->B.base);
    if (se->pinned) cudaFreeHost(se->pinned);
    if (se->h_counts) cudaFreeHost(se->h_counts);
    free(se);
}

static int se_run(s3_se *se, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads64, uint32_t wordPerQuery, int onDevice, s3_se_result *res)
{
    if (!se || !queries || !readLengths || !res) { s3_set_error("s3_se_align: NULL argument"); return S3_EINVAL; }
    memset(res, 0, sizeof *res);
    if (numReads64 > se->maxReads) { s3_set_error("s3_se_align: %llu reads > maxReads %u", (unsigned long long)numReads64, se->maxReads); return S3_EINVAL; }
    const uint32_t N = (uint32_t)numReads64;
    if (N == 0) return S3_OK;
    s3_index *ix = se->ix;
    S3_TRYC(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const uint32_t k = se->par.numMismatch, C = kCases[k], allowed = kAllowed[k], wpa = 2 * allowed;
    const size_t up = ((size_t)N + 31) / 32 * 32, maxRanges = (size_t)N * C * allowed;
    int rc;
    size_t scanTemp = 0;
    cub::DeviceScan::ExclusiveSum(NULL, scanTemp, (uint32_t *)NULL, (uint32_t *)NULL, (int)(N + 1), st);
    const size_t needA = arena_need(up * wordPerQuery, 4) + arena_need(up, 4) + C * arena_need(up * wpa, 4) + 7 * arena_need(N + 1, 4) + arena_need(up, 4) + arena_need(N, 1) +
                         2 * arena_need(maxRanges, 4) + arena_need(maxRanges, 2) + arena_need(maxRanges, 1) + arena_need(scanTemp, 1) + 4096;
    if ((rc = arena_reserve(&se->A, needA, st))) return rc;
    S3Arena *A = &se->A;
    uint32_t *d_q = onDevice ? const_cast<uint32_t *>(queries) : arena_take<uint32_t>(A, up * wordPerQuery);
    uint32_t *d_len = onDevice ? const_cast<uint32_t *>(readLengths) : arena_take<uint32_t>(A, up);
    uint32_t *d_ans[S3_MAX_NUM_CASES];
    for (uint32_t c = 0; c < C; ++c) d_ans[c] = arena_take<uint32_t>(A, up * wpa);
    uint32_t *d_nRanges = arena_take<uint32_t>(A, N + 1), *d_rangeOff = arena_take<uint32_t>(A, N + 1), *d_totOcc = arena_take<uint32_t>(A, N + 1);
    uint32_t *d_locCount = arena_take<uint32_t>(A, N + 1), *d_locOff = arena_take<uint32_t>(A, N + 1);
    uint8_t *d_readFlags = arena_take<uint8_t>(A, N);
    uint32_t *d_seedLen = arena_take<uint32_t>(A, up), *d_valCount = arena_take<uint32_t>(A, N + 1), *d_valOff = arena_take<uint32_t>(A, N + 1);
    uint32_t *d_saL = arena_take<uint32_t>(A, maxRanges), *d_saR = arena_take<uint32_t>(A, maxRanges);
    uint8_t *d_saFlags = arena_take<uint8_t>(A, 2 * maxRanges), *d_keep = arena_take<uint8_t>(A, maxRanges);
    void *d_tmp = arena_take<char>(A, scanTemp);
    if (!d_tmp) { s3_set_error("s3_se_align: stage buffer accounting"); return S3_ENOMEM; }
    if (!onDevice) {
        S3_TRYC(cudaMemcpyAsync(d_q, queries, up * wordPerQuery * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_len, readLengths, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    }
    // long-read mode: the search aligns the first 100 bases of a read longer than 120 (alignment.cu:2475-2491)
    uint32_t *d_searchLen = d_len;
    if (se->par.longReadMode) {
        d_searchLen = d_seedLen;
        s3_se_seed_length_kernel<<<(unsigned)((up + 255) / 256), 256, 0, st>>>((uint32_t)up, N, d_len, d_seedLen);
        S3_LAUNCHED(1);
    }
    if ((rc = s3_search_round1_device(ix, d_q, d_searchLen, N, wordPerQuery, k, C, allowed, wpa, 0, d_ans, NULL))) return rc;
    S3Collect col;
    memset(&col, 0, sizeof col);
    for (uint32_t c = 0; c < C; ++c) col.answers[c] = d_ans[c];
    col.numCases = C; col.allowed = allowed; col.wordPerAns = wpa; col.numReads = N; col.textLength = ix->textLength;
    col.maxOutputPerRead = se->par.maxOutputPerRead;
    const unsigned nbR = (N + 255) / 256;
    S3_TRYC(cudaMemsetAsync(d_nRanges + N, 0, 4, st));
    s3_pe_collect_kernel<false><<<nbR, 256, 0, st>>>(col, d_nRanges, d_totOcc, d_readFlags, NULL, NULL, NULL, NULL);
    S3_TRYC(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_nRanges, d_rangeOff, (int)(N + 1), st));
    s3_pe_collect_kernel<true><<<nbR, 256, 0, st>>>(col, NULL, NULL, NULL, d_rangeOff, d_saL, d_saR, d_saFlags);
    S3_TRYC(cudaMemsetAsync(d_keep, 1, maxRanges, st));
    S3_TRYC(cudaMemsetAsync(d_locCount + N, 0, 4, st));
    s3_se_select_kernel<<<nbR, 256, 0, st>>>(N, d_rangeOff, d_saL, d_saR, d_saFlags, d_totOcc, se->par.reportBest, d_keep, d_locCount);
    S3_TRYC(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_locCount, d_locOff, (int)(N + 1), st));
    S3_LAUNCHED(3);
    S3_TRYC(cudaGetLastError());
    S3_TRYC(cudaMemcpyAsync(se->h_counts, d_locOff + N, 4, cudaMemcpyDeviceToHost, st));
    S3_TRYC(cudaMemcpyAsync(se->h_counts + 1, d_rangeOff + N, 4, cudaMemcpyDeviceToHost, st));
    S3_TRYC(cudaStreamSynchronize(st));
    const uint32_t T = se->h_counts[0];
    const size_t Tm = T ? T : 1;
    if ((rc = arena_reserve(&se->B, arena_need(Tm, 4) + arena_need(Tm, 2) + arena_need(Tm, 8) + arena_need(Tm, 4) + 4096, st))) return rc;
    uint32_t *d_occPos = arena_take<uint32_t>(&se->B, Tm);
    uint8_t *d_occFlags = arena_take<uint8_t>(&se->B, 2 * Tm);
    unsigned long long *d_key = arena_take<unsigned long long>(&se->B, Tm);
    uint32_t *d_val = arena_take<uint32_t>(&se->B, Tm);
    if (!d_val) { s3_set_error("s3_se_align: stage buffer accounting"); return S3_ENOMEM; }
    if (T) {
        s3_pe_locate_kernel<<<nbR, 256, 0, st>>>(N, ix->loc.sa, d_rangeOff, d_saL, d_saR, d_saFlags, d_keep, d_locOff, d_occPos, d_occFlags, d_key, d_val);
        S3_LAUNCHED(1);
        S3_TRYC(cudaGetLastError());
    }
    uint32_t Tout = T;
    if (se->par.longReadMode && T) {
        // validateAlignments on every read's list, then the kept entries gathered into lists of their own
        if (!ix->loc.text) { s3_set_error("s3_se_align: long-read mode needs the packed text on the device"); return S3_EINVAL; }
        S3_TRYC(cudaMemsetAsync(d_valCount + N, 0, 4, st));
        s3_validate_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, d_q, wordPerQuery, d_len, ix->loc.text, ix->textLength, d_locOff, d_occPos, d_occFlags,
                                                             se->par.onlyKeepBest, se->par.minSeedMismatch, se->par.doubleAllowance,
                                                             (int)se->par.maxOutputPerRead, d_valCount);
        S3_TRYC(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_valCount, d_valOff, (int)(N + 1), st));
        s3_validate_gather_kernel<<<nbR, 256, 0, st>>>(N, d_locOff, d_valCount, d_valOff, d_occPos, d_occFlags, (uint32_t *)d_key, (uint8_t *)d_val);
        S3_LAUNCHED(3);
        S3_TRYC(cudaGetLastError());
        S3_TRYC(cudaMemcpyAsync(se->h_counts + 2, d_valOff + N, 4, cudaMemcpyDeviceToHost, st));
        S3_TRYC(cudaStreamSynchronize(st));
        Tout = se->h_counts[2];
        d_locOff = d_valOff; d_occPos = (uint32_t *)d_key; d_occFlags = (uint8_t *)d_val;
    }
    res->numReads = N; res->numRanges = se->h_counts[1]; res->numOccurrences = Tout;
    if (onDevice) {
        S3_TRYC(cudaStreamSynchronize(st));
        res->d_occOffsets = d_locOff; res->d_positions = d_occPos; res->d_occFlags = d_occFlags; res->d_readFlags = d_readFlags;
        return S3_OK;
    }
    const size_t To = Tout ? Tout : 1;
    const size_t bytes = ((size_t)N + 1) * 4 + 256 + To * 4 + 256 + 2 * To + 256 + N + 256;
    if (bytes > se->pinnedBytes) {
        if (se->pinned) { cudaFreeHost(se->pinned); se->pinned = NULL; se->pinnedBytes = 0; }
        if (cudaMallocHost(&se->pinned, bytes + bytes / 4) != cudaSuccess) { s3_set_error("s3_se_align: pinned allocation failed"); return S3_ENOMEM; }
        se->pinnedBytes = bytes + bytes / 4;
    }
    char *h = (char *)se->pinned;
    res->occOffsets = (uint32_t *)h; h += (((size_t)N + 1) * 4 + 255) / 256 * 256;
    res->positions = (uint32_t *)h; h += (To * 4 + 255) / 256 * 256;
    res->occFlags = (uint8_t *)h; h += (2 * To + 255) / 256 * 256;
    res->readFlags = (uint8_t *)h;
    S3_TRYC(cudaMemcpyAsync(res->occOffsets, d_locOff, ((size_t)N + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (Tout) {
        S3_TRYC(cudaMemcpyAsync(res->positions, d_occPos, (size_t)Tout * 4, cudaMemcpyDeviceToHost, st));
        S3_TRYC(cudaMemcpyAsync(res->occFlags, d_occFlags, (size_t)Tout * 2, cudaMemcpyDeviceToHost, st));
    }
    S3_TRYC(cudaMemcpyAsync(res->readFlags, d_readFlags, N, cudaMemcpyDeviceToHost, st));
    S3_TRYC(cudaStreamSynchronize(st));
    res->h2dBytes = up * wordPerQuery * 4 + (size_t)N * 4;
    res->d2hBytes = ((size_t)N + 1) * 4 + (size_t)Tout * 6 + N + 8;
    return S3_OK;
}

extern "C" int s3_se_align(s3_se *se, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery, s3_se_result *out)
{
    return se_run(se, queries, readLengths, numReads, wordPerQuery, 0, out);
}
extern "C" int s3_se_align_device(s3_se *se, const uint32_t *d_queries, const uint32_t *d_readLengths, uint64_t numReads, uint32_t wordPerQuery, s3_se_result *out)
{
    return se_run(se, d_queries, d_readLengths, numReads, wordPerQuery, 1, out);
}


// =======================================================================================================================
// Alignment step of the seeded DP stages (s3_single_dp_align, s3_deep_dp_align): see s3_common.cuh.
// =======================================================================================================================
struct S3StageWs {
    s3_dp *dp; uint32_t maxRead, maxDNA, cap; s3_dp_scores scores;
    uint32_t *d_q; size_t qBytes;                 // query buffer of the current stage call when it came from the host
    const uint32_t *d_qUse;                       // the query buffer the stage works on: d_q, or the caller's device buffer
    S3Arena arena;
    void *pinned[2]; size_t pinnedBytes[2];
};

void s3_stage_ws_free(s3_index *ix)
{
    S3StageWs *ws = (S3StageWs *)ix->stageWs;
    if (!ws) return;
    if (ws->dp) s3_dp_free(ws->dp);
    if (ws->d_q) cudaFree(ws->d_q);
    if (ws->arena.base) cudaFree(ws->arena.base);
    for (int i = 0; i < 2; ++i) if (ws->pinned[i]) cudaFreeHost(ws->pinned[i]);
    free(ws);
    ix->stageWs = NULL;
}

__global__ void s3_stage_readlen_kernel(uint32_t n, const uint32_t *__restrict__ readID, const uint32_t *__restrict__ lengthsByRead, uint32_t *__restrict__ out)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = lengthsByRead[readID[t]];
}

static S3StageWs *stage_ws(s3_index *ix)
{
    S3StageWs *ws = (S3StageWs *)ix->stageWs;
    if (!ws) {
        ws = (S3StageWs *)calloc(1, sizeof(S3StageWs));
        ix->stageWs = ws;
    }
    return ws;
}

const uint32_t *s3_stage_queries(s3_index *ix) { S3StageWs *ws = (S3StageWs *)ix->stageWs; return ws ? ws->d_qUse : NULL; }

// the stage works on a query buffer that is on the device already (the chain's): nothing is copied
int s3_stage_use_queries(s3_index *ix, const uint32_t *d_queries)
{
    S3StageWs *ws = stage_ws(ix);
    if (!ws) { s3_set_error("out of host memory"); return S3_ENOMEM; }
    ws->d_qUse = d_queries;
    return S3_OK;
}

int s3_stage_upload_queries(s3_index *ix, const uint32_t *queries, uint64_t numReads, uint32_t wordPerQuery)
{
    S3StageWs *ws = stage_ws(ix);
    if (!ws) { s3_set_error("out of host memory"); return S3_ENOMEM; }
    S3_TRYC(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const size_t up = ((size_t)numReads + 31) / 32 * 32, qBytes = up * wordPerQuery * 4;
    if (qBytes > ws->qBytes) {
        if (ws->d_q) { S3_TRYC(cudaStreamSynchronize(st)); cudaFree(ws->d_q); ws->d_q = NULL; ws->qBytes = 0; }
        S3_TRYC(cudaMalloc(&ws->d_q, qBytes + qBytes / 4));
        ws->qBytes = qBytes + qBytes / 4;
    }
    S3_TRYC(cudaMemcpyAsync(ws->d_q, queries, qBytes, cudaMemcpyHostToDevice, st));
    ws->d_qUse = ws->d_q;
    return S3_OK;
}

int s3_stage_align(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery, int uploadQueries,
                   uint32_t maxRead, uint32_t maxDNA, s3_dp_scores scores, int slot, uint64_t n64,
                   const uint32_t *readID, const uint8_t *strand, const uint32_t *start, const uint32_t *len, const int32_t *cutoff,
                   const uint32_t *clipLt, const uint32_t *clipRt, const uint32_t *ancL, const uint32_t *ancR,
                   const uint32_t *d_lenByRead, const uint32_t *d_cand, S3StageAligned *out)
{
    memset(out, 0, sizeof *out);
    if (n64 == 0) return S3_OK;
    if (n64 > 0x7FFFFFF0ull) { s3_set_error("s3_stage_align: too many alignments in one call"); return S3_EINVAL; }
    const uint32_t n = (uint32_t)n64;
    const bool dev = d_lenByRead != NULL;
    S3_TRYC(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    S3StageWs *ws = stage_ws(ix);
    if (!ws) { s3_set_error("out of host memory"); return S3_ENOMEM; }
    int rc;
    if (!ws->dp || ws->maxRead != maxRead || ws->maxDNA != maxDNA || ws->cap < n || memcmp(&ws->scores, &scores, sizeof scores)) {
        if (ws->dp) { S3_TRYC(cudaStreamSynchronize(st)); s3_dp_free(ws->dp); ws->dp = NULL; }
        const uint32_t cap = n + n / 4 < 65536u ? 65536u : n + n / 4;
        if ((rc = s3_dp_create(maxRead, maxDNA, cap, scores, ix->device, &ws->dp))) return rc;
        s3_dp_set_stream(ws->dp, st);
        ws->maxRead = maxRead; ws->maxDNA = maxDNA; ws->cap = cap; ws->scores = scores;
    }
    if (uploadQueries || !ws->d_qUse) { if ((rc = s3_stage_upload_queries(ix, queries, numReads, wordPerQuery))) return rc; }
    const size_t patLen = s3_dp_pattern_length(ws->dp);
    size_t scanTemp = 0;
    cub::DeviceScan::ExclusiveSum(NULL, scanTemp, (uint32_t *)NULL, (uint32_t *)NULL, (int)(n + 1), st);
    const size_t need = 13 * arena_need(n, 4) + arena_need(n, 1) + arena_need((size_t)n * patLen, 1) + 2 * arena_need(n + 1, 4) +
                        arena_need((size_t)n * (maxRead + 8), 4) + arena_need(scanTemp, 1) + 4096;
    if ((rc = arena_reserve(&ws->arena, need, st))) return rc;
    S3Arena *A = &ws->arena;
    uint32_t *d_readID = arena_take<uint32_t>(A, n), *d_start = arena_take<uint32_t>(A, n), *d_len = arena_take<uint32_t>(A, n), *d_rl = arena_take<uint32_t>(A, n);
    uint32_t *d_clt = arena_take<uint32_t>(A, n), *d_crt = arena_take<uint32_t>(A, n), *d_al = arena_take<uint32_t>(A, n), *d_ar = arena_take<uint32_t>(A, n);
    int32_t *d_cut = arena_take<int32_t>(A, n), *d_score = arena_take<int32_t>(A, n);
    uint32_t *d_hit = arena_take<uint32_t>(A, n), *d_cnt = arena_take<uint32_t>(A, n), *d_spare = arena_take<uint32_t>(A, n);
    uint8_t *d_strand = arena_take<uint8_t>(A, n);
    uint8_t *d_pattern = arena_take<uint8_t>(A, (size_t)n * patLen);
    uint32_t *d_runCount = arena_take<uint32_t>(A, n + 1), *d_runOff = arena_take<uint32_t>(A, n + 1);
    uint32_t *d_runs = arena_take<uint32_t>(A, (size_t)n * (maxRead + 8));
    void *d_tmp = arena_take<char>(A, scanTemp);
    (void)d_spare;
    if (!d_tmp) { s3_set_error("s3_stage_align: stage buffer accounting"); return S3_ENOMEM; }
    // host side: results of this slot (and, staged there first, the per-alignment read lengths of the host mode)
    const size_t stride = (((size_t)n + 1) * 4 + 255) / 256 * 256;
    const size_t hostBytes = 10 * stride + (size_t)n * (maxRead + 8) * 4 + 256;
    if (hostBytes > ws->pinnedBytes[slot]) {
        if (ws->pinned[slot]) { S3_TRYC(cudaStreamSynchronize(st)); cudaFreeHost(ws->pinned[slot]); ws->pinned[slot] = NULL; ws->pinnedBytes[slot] = 0; }
        if (cudaMallocHost(&ws->pinned[slot], hostBytes + hostBytes / 4) != cudaSuccess) { s3_set_error("s3_stage_align: pinned allocation failed"); return S3_ENOMEM; }
        ws->pinnedBytes[slot] = hostBytes + hostBytes / 4;
    }
    char *h = (char *)ws->pinned[slot];
    int32_t *h_score = (int32_t *)h; uint32_t *h_hit = (uint32_t *)(h + stride), *h_cnt = (uint32_t *)(h + 2 * stride), *h_runOff = (uint32_t *)(h + 3 * stride);
    uint32_t *h_readID = (uint32_t *)(h + 4 * stride), *h_start = (uint32_t *)(h + 5 * stride), *h_cand = (uint32_t *)(h + 6 * stride);
    int32_t *h_cutoff = (int32_t *)(h + 7 * stride);
    uint8_t *h_strand = (uint8_t *)(h + 8 * stride);
    uint32_t *h_runs = (uint32_t *)(h + 10 * stride);
    if (dev) {
        // the DP writes its right clips in place, and the caller's arrays may be reused: work on copies
        d_readID = const_cast<uint32_t *>(readID); d_start = const_cast<uint32_t *>(start); d_len = const_cast<uint32_t *>(len);
        d_cut = const_cast<int32_t *>(cutoff); d_clt = const_cast<uint32_t *>(clipLt); d_al = const_cast<uint32_t *>(ancL); d_ar = const_cast<uint32_t *>(ancR);
        d_strand = const_cast<uint8_t *>(strand);
        S3_TRYC(cudaMemcpyAsync(d_crt, clipRt, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        s3_stage_readlen_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, d_readID, d_lenByRead, d_rl);
        S3_LAUNCHED(1);
    } else {
        for (uint32_t t = 0; t < n; ++t) {
            if (readID[t] >= numReads) { s3_set_error("s3_stage_align: read id %u out of range", readID[t]); return S3_EINVAL; }
            h_runs[t] = readLengths[readID[t]];
        }
        S3_TRYC(cudaMemcpyAsync(d_rl, h_runs, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_readID, readID, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_start, start, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_len, len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_cut, cutoff, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_clt, clipLt, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_crt, clipRt, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_al, ancL, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_ar, ancR, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        S3_TRYC(cudaMemcpyAsync(d_strand, strand, n, cudaMemcpyHostToDevice, st));
    }
    if ((rc = s3_dp_align_windows_device(ws->dp, ix, ws->d_qUse, wordPerQuery, d_readID, d_strand, d_start, d_len, d_rl, d_cut, d_score, d_hit, d_cnt, d_pattern, n,
                                         d_clt, d_crt, d_al, d_ar))) return rc;
    const unsigned nb = (n + 127) / 128;
    S3_TRYC(cudaMemsetAsync(d_runCount + n, 0, 4, st));
    s3_pe_runs_kernel<false><<<nb, 128, 0, st>>>(n, d_pattern, (uint32_t)patLen, d_score, d_cut, d_runCount, NULL, NULL);
    S3_TRYC(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_runCount, d_runOff, (int)(n + 1), st));
    s3_pe_runs_kernel<true><<<nb, 128, 0, st>>>(n, d_pattern, (uint32_t)patLen, d_score, d_cut, NULL, d_runOff, d_runs);
    S3_LAUNCHED(2);
    S3_TRYC(cudaGetLastError());
    S3_TRYC(cudaMemcpyAsync(h_score, d_score, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    S3_TRYC(cudaMemcpyAsync(h_hit, d_hit, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    S3_TRYC(cudaMemcpyAsync(h_cnt, d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    S3_TRYC(cudaMemcpyAsync(h_runOff, d_runOff, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (dev) {
        S3_TRYC(cudaMemcpyAsync(h_readID, d_readID, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        S3_TRYC(cudaMemcpyAsync(h_start, d_start, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        S3_TRYC(cudaMemcpyAsync(h_cutoff, d_cut, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        S3_TRYC(cudaMemcpyAsync(h_strand, d_strand, n, cudaMemcpyDeviceToHost, st));
        if (d_cand) S3_TRYC(cudaMemcpyAsync(h_cand, d_cand, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    }
    S3_TRYC(cudaStreamSynchronize(st));
    const uint32_t total = h_runOff[n];
    if (total) {
        S3_TRYC(cudaMemcpyAsync(h_runs, d_runs, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        S3_TRYC(cudaStreamSynchronize(st));
    }
    out->score = h_score; out->hit = h_hit; out->cnt = h_cnt; out->runOff = h_runOff; out->runs = h_runs; out->numRuns = total;
    out->readID = h_readID; out->start = h_start; out->cutoff = h_cutoff; out->strand = h_strand; out->cand = h_cand;
    out->d_score = d_score; out->d_hit = d_hit;
    return S3_OK;
}
