// Which text window, clips, anchors and cutoff a candidate is aligned with: the decisions of the three DP engines'
// batch packers, shared by s3_dp_make_windows (csrc/s3_windows.cu) and the paired-end chain (csrc/s3_chain.cu).
//
// Replaces the deciding half of SingleEndAlgnBatch::pack (DV-DPfunctions.cu:1425-1468), HalfEndAlgnBatch::pack
// (:2027-2110), PairEndAlgnBatch::packLeft (:3374-3418) and packRight (:3420-3472); the copying half (packRead /
// repackDNA) is s3_dp_pack_kernel.  All arithmetic is the reference's 32-bit unsigned arithmetic: a window start that
// wraps below 0 is caught by the same `>= text length` tests.  Plain integer code with no CUDA in it besides the S3_HD
// qualifier, so the CPU tier compiles this very file with a host compiler (tests/native/windows_harness.cpp).
#pragma once
#include <stdint.h>

#ifndef S3_HD
#ifdef __CUDACC__
#define S3_HD __host__ __device__ __forceinline__
#else
#define S3_HD inline
#endif
#endif

struct S3WinParams {
    int32_t insertLow, insertHigh;       // -v / -u
    int32_t leftLeg, rightLeg;           // peStrandLeftLeg / peStrandRightLeg (1 or 2)
    int32_t softClipLeft, softClipRight; // DPParameters
    int32_t cutoff[2];                   // paramRead[0 / 1].cutoffThreshold (even / odd read of a pair); < 0: ceil(0.3 * read length)
    uint32_t maxDNALength;               // the engine's window capacity ("no anchor" is stored as this value)
    uint32_t textLength;                 // hsp->dnaLength
};

struct S3Window {
    uint32_t readID;                     // the read that is aligned
    uint32_t start, dnaLen, readLen;     // text window [start, start + dnaLen), the read's length
    uint32_t clipLt, clipRt, ancL, ancR;
    int32_t cutoff;
    uint8_t strand;                      // the read as given (1) or reverse-complemented (2)
    uint8_t leftOrRight;                 // half-end only: 1 the read is the right end of the pair, 0 the left end
};

// DPS_MARGIN / DP2_MARGIN (DV-DPfunctions.cu:1005,2549)
S3_HD uint32_t s3_win_margin(uint32_t len) { return len > 100u ? len >> 2 : 25u; }

// (int) ceil(DP_SCORE_THRESHOLD_RATIO * (double) read_length), ratio 0.3 (CPUfunctions.cpp:65,73), unless the ini names a value
S3_HD int32_t s3_win_cutoff(const S3WinParams &w, uint32_t readID, uint32_t len)
{
    const int32_t c = w.cutoff[readID & 1u];
    if (c >= 0) return c;
    const double x = 0.3 * (double)len;
    int32_t i = (int32_t)x;
    return ((double)i < x) ? i + 1 : i;
}

// read +- margin, cut at the text's ends; clips by strand
S3_HD void s3_win_around(const S3WinParams &w, uint32_t pos, uint32_t len, int strand, S3Window &o)
{
    const uint32_t margin = s3_win_margin(len);
    uint32_t start = pos - margin;
    if (start >= w.textLength) start = 0;
    uint32_t n = len + 2u * margin;
    if (start + n > w.textLength) n = w.textLength - start;
    o.start = start; o.dnaLen = n; o.readLen = len; o.strand = (uint8_t)strand;
    o.clipLt = (uint32_t)(strand == 1 ? w.softClipLeft : w.softClipRight);
    o.clipRt = (uint32_t)(strand == 1 ? w.softClipRight : w.softClipLeft);
}

// single-end DP: a seed candidate (read, estimated start, strand)
S3_HD void s3_win_single(const S3WinParams &w, uint32_t readID, uint32_t pos, int strand, uint32_t len, S3Window &o)
{
    s3_win_around(w, pos, len, strand, o);
    o.readID = readID; o.ancL = w.maxDNALength; o.ancR = 0; o.leftOrRight = 0;
    o.cutoff = s3_win_cutoff(w, 0, len);               // one threshold for the batch (SingleEndAlgnBatch: cutoffThreshold)
}

// mate rescue: one occurrence of the aligned read -> 0, 1 or 2 windows for its mate (two only when both legs have one strand)
S3_HD int s3_win_half(const S3WinParams &w, uint32_t alignedReadID, uint32_t pos, int strand, uint32_t alignedLen, uint32_t mateLen, S3Window o[2])
{
    int n = 0;
    const uint32_t mate = alignedReadID ^ 1u;
    for (int side = 0; side < 2; ++side) {
        // side 0: the aligned read is the left end and its mate lies to the right; side 1: the other way round
        if (strand != (side == 0 ? w.leftLeg : w.rightLeg)) continue;
        uint32_t start, stop;
        if (side == 0) {
            stop = pos + (uint32_t)w.insertHigh;
            start = pos + (uint32_t)w.insertLow - mateLen;
            if (start < pos) start = pos;
        } else {
            start = pos + alignedLen - (uint32_t)w.insertHigh;
            stop = pos + alignedLen - (uint32_t)w.insertLow + mateLen;
            if (stop >= pos + alignedLen) stop = pos + alignedLen - 1u;
        }
        if (!(start < w.textLength && stop <= w.textLength)) continue;
        S3Window &x = o[n++];
        const int dpStrand = side == 0 ? w.rightLeg : w.leftLeg;
        x.readID = mate; x.start = start; x.dnaLen = stop - start; x.readLen = mateLen;
        x.strand = (uint8_t)dpStrand; x.leftOrRight = side == 0 ? 1 : 0;
        x.clipLt = (uint32_t)(dpStrand == 1 ? w.softClipLeft : w.softClipRight);
        x.clipRt = (uint32_t)(dpStrand == 1 ? w.softClipRight : w.softClipLeft);
        x.ancL = side == 0 ? w.maxDNALength : (uint32_t)(w.insertHigh - w.insertLow + 1);
        x.ancR = side == 0 ? mateLen : 0u;
        x.cutoff = s3_win_cutoff(w, mate, mateLen);
    }
    return n;
}

// deep DP, left read of a candidate pair: free window, no anchors
S3_HD void s3_win_pair_left(const S3WinParams &w, uint32_t readIDLeft, uint32_t posLeft, uint32_t len, S3Window &o)
{
    s3_win_around(w, posLeft, len, w.leftLeg, o);
    o.readID = readIDLeft; o.ancL = w.maxDNALength; o.ancR = 0; o.leftOrRight = 0;
    o.cutoff = s3_win_cutoff(w, readIDLeft, len);
}

// deep DP, right read, once the left read is aligned at hitPosLeft = its window start + hitLoc: the window ends where the
// largest insert ends, and the alignment must end at or beyond the smallest insert (right anchor)
S3_HD void s3_win_pair_right(const S3WinParams &w, uint32_t readIDLeft, uint32_t posRight, uint32_t hitPosLeft, uint32_t lenRight, S3Window &o)
{
    const uint32_t right = readIDLeft ^ 1u;
    s3_win_around(w, posRight, lenRight, w.rightLeg, o);
    const uint32_t bounded = hitPosLeft + (uint32_t)w.insertHigh - o.start;
    if (bounded < o.dnaLen) o.dnaLen = bounded;
    const int32_t anchor = (int32_t)(hitPosLeft + (uint32_t)w.insertLow - o.start);
    o.readID = right; o.ancL = w.maxDNALength; o.ancR = anchor > 0 ? (uint32_t)anchor : 0u; o.leftOrRight = 0;
    o.cutoff = s3_win_cutoff(w, right, lenRight);
}
