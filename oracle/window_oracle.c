/* TEST INFRASTRUCTURE ONLY: CPU restatement of what the reference's three DP engines decide when they pack a batch -- which
 * text window, which clips, anchors and cutoff each candidate is aligned with.  Follows, statement by
 * statement:
 *   SingleEndAlgnBatch::pack            DV-DPfunctions.cu:1425-1468   window = [pos - margin, + len + 2 margin), margin = DPS_MARGIN :1005
 *   HalfEndAlgnBatch::pack              DV-DPfunctions.cu:2027-2110   mate rescue: window from the aligned mate's position and the insert range
 *   PairEndAlgnBatch::packLeft          DV-DPfunctions.cu:3374-3418   deep DP, left read: free window, no anchors
 *   PairEndAlgnBatch::packRight         DV-DPfunctions.cu:3420-3472   deep DP, right read: window cut at hitPosLeft + insert_high, right anchor
 * Pinned against those bodies cut out of the reference and compiled (oracle/_ref/libref_windows.so): tests/test_cpu_oracle_vs_ref.py.
 */
#include <stdint.h>

#define MARGIN(l) (((l) > 100) ? ((l) >> 2) : 25)

/* -> 1 (always packs) */
int s3o_window_single(uint32_t readID, uint32_t pos, int strand, const uint32_t *readLengths, uint32_t textLength, int clipLeft, int clipRight,
                      uint32_t *start, uint32_t *len, uint32_t *clipLt, uint32_t *clipRt)
{
    uint32_t readLength = readLengths[readID];
    int margin = MARGIN(readLength);
    uint32_t DNAStart = pos - margin;
    if (DNAStart >= textLength) DNAStart = 0;
    uint32_t DNALength = readLength + margin * 2;
    if (DNAStart + DNALength > textLength) DNALength = textLength - DNAStart;
    *clipLt = (strand == 1) ? clipLeft : clipRight;
    *clipRt = (strand == 1) ? clipRight : clipLeft;
    *start = DNAStart; *len = DNALength;
    return 1;
}

/* one occurrence of an aligned read -> 0, 1 or 2 windows for its mate; out arrays take two entries.  Returns the count. */
int s3o_window_half(uint32_t alignedReadID, uint32_t alignedPos, int alignedStrand, const uint32_t *readLengths, uint32_t textLength,
                    int leftLeg, int rightLeg, int insertHigh, int insertLow, int maxDNALength, int clipLeft, int clipRight,
                    int *leftOrRight, uint32_t *start, uint32_t *len, uint32_t *readLen, int *dpStrand, uint32_t *clipLt, uint32_t *clipRt,
                    uint32_t *ancL, uint32_t *ancR)
{
    int n = 0;
    if (alignedStrand != leftLeg && alignedStrand != rightLeg) return 0;
    uint32_t alignedReadLength = readLengths[alignedReadID];
    int unalignedIsReadOrMate = 1 - (alignedReadID & 1);
    uint32_t unalignedReadID = (unalignedIsReadOrMate == 0 ? alignedReadID - 1 : alignedReadID + 1);
    uint32_t unalignedReadLength = readLengths[unalignedReadID];
    if (leftLeg == alignedStrand) {
        uint32_t rightEnd = alignedPos + insertHigh;
        uint32_t rightStart = alignedPos + insertLow - unalignedReadLength;
        if (rightStart < alignedPos) rightStart = alignedPos;
        if (rightStart < textLength && rightEnd <= textLength) {
            leftOrRight[n] = 1; readLen[n] = unalignedReadLength; start[n] = rightStart; len[n] = rightEnd - rightStart;
            ancL[n] = maxDNALength; ancR[n] = unalignedReadLength;
            dpStrand[n] = rightLeg;
            clipLt[n] = (rightLeg == 1) ? clipLeft : clipRight; clipRt[n] = (rightLeg == 1) ? clipRight : clipLeft;
            ++n;
        }
    }
    if (rightLeg == alignedStrand) {
        uint32_t leftStart = alignedPos + alignedReadLength - insertHigh;
        uint32_t leftEnd = alignedPos + alignedReadLength - insertLow + unalignedReadLength;
        if (leftEnd >= alignedPos + alignedReadLength) leftEnd = alignedPos + alignedReadLength - 1;
        if (leftStart < textLength && leftEnd <= textLength) {
            leftOrRight[n] = 0; readLen[n] = unalignedReadLength; start[n] = leftStart; len[n] = leftEnd - leftStart;
            ancL[n] = insertHigh - insertLow + 1; ancR[n] = 0;
            dpStrand[n] = leftLeg;
            clipLt[n] = (leftLeg == 1) ? clipLeft : clipRight; clipRt[n] = (leftLeg == 1) ? clipRight : clipLeft;
            ++n;
        }
    }
    return n;
}

void s3o_window_pair_left(uint32_t readIDLeft, uint32_t posLeft, const uint32_t *readLengths, uint32_t textLength, int leftLeg, int maxDNALength,
                          int clipLeft, int clipRight, uint32_t *start, uint32_t *len, uint32_t *clipLt, uint32_t *clipRt, uint32_t *ancL, uint32_t *ancR)
{
    uint32_t readLength = readLengths[readIDLeft];
    int margin = MARGIN(readLength);
    uint32_t DNAStartLeft = posLeft - margin;
    if (DNAStartLeft >= textLength) DNAStartLeft = 0;
    uint32_t DNALength = readLength + margin * 2;
    if (DNAStartLeft + DNALength > textLength) DNALength = textLength - DNAStartLeft;
    *ancL = maxDNALength; *ancR = 0;
    *clipLt = (leftLeg == 1) ? clipLeft : clipRight; *clipRt = (leftLeg == 1) ? clipRight : clipLeft;
    *start = DNAStartLeft; *len = DNALength;
}

/* the right read of a candidate whose left read reached its cutoff (the caller checks that) */
void s3o_window_pair_right(uint32_t readIDLeft, uint32_t posRight, uint32_t startLeft, uint32_t hitLocLeft, const uint32_t *readLengths, uint32_t textLength,
                           int rightLeg, int insertHigh, int insertLow, int maxDNALength, int clipLeft, int clipRight,
                           uint32_t *readIDRight, uint32_t *start, uint32_t *len, uint32_t *clipLt, uint32_t *clipRt, uint32_t *ancL, uint32_t *ancR)
{
    uint32_t leftIsOdd = readIDLeft & 1;
    uint32_t idRight = leftIsOdd ? (readIDLeft - 1) : (readIDLeft + 1);
    uint32_t readLength = readLengths[idRight];
    uint32_t margin = MARGIN(readLength);
    uint32_t DNAStartRight = posRight - margin;
    if (DNAStartRight >= textLength) DNAStartRight = 0;
    uint32_t DNALength = readLength + margin * 2;
    if (DNAStartRight + DNALength > textLength) DNALength = textLength - DNAStartRight;
    uint32_t hitPosLeft = startLeft + hitLocLeft;
    uint32_t boundedLength = hitPosLeft + insertHigh - DNAStartRight;
    if (boundedLength < DNALength) DNALength = boundedLength;
    *ancL = maxDNALength;
    int rightAnchor = hitPosLeft + insertLow - DNAStartRight;
    *ancR = rightAnchor > 0 ? rightAnchor : 0;
    *clipLt = (rightLeg == 1) ? clipLeft : clipRight; *clipRt = (rightLeg == 1) ? clipRight : clipLeft;
    *readIDRight = idRight; *start = DNAStartRight; *len = DNALength;
}
