/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * The reference's own DP batch packers, run on the host for what they DECIDE (which window, clips, anchors, cutoff):
 *   SingleDP_Space::SingleEndAlignmentEngine::SingleEndAlgnBatch::pack   DV-DPfunctions.cu:1425-1468
 *   DP_Space::HalfEndAlignmentEngine::HalfEndAlgnBatch::pack             DV-DPfunctions.cu:2027-2110
 *   DeepDP_Space::PairEndAlignmentEngine::PairEndAlgnBatch::packLeft     DV-DPfunctions.cu:3374-3418
 *   DeepDP_Space::PairEndAlignmentEngine::PairEndAlgnBatch::packRight    DV-DPfunctions.cu:3420-3472
 * The four bodies are cut out of the reference file by sed at build time (win_*.inc; the qualified method names become
 * free functions, nothing else changes); the batch members they use are the file-level variables of the namespaces below,
 * and the two calls that copy bases (packRead, repackDNA -- what s3_dp_pack_kernel does on the device) are swallowed.
 * Pins oracle/window_oracle.c.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef unsigned int uint;
typedef unsigned char uchar;
#define packRead(...) ((void)0)
#define repackDNA(...) ((void)0)
#define DPS_MARGIN(l) ((l>100) ? (l>>2) : 25)      /* DV-DPfunctions.cu:1005 */
#define DP2_MARGIN(l) ((l>100) ? (l>>2) : 25)      /* DV-DPfunctions.cu:2549 */

namespace single_ns {
struct SeedPos { uint pos; uint readID; int strand; };              /* DV-DPfunctions.h:919-924 */
typedef struct SeedPos CandidateInfo;
static int numOfThreads, batchSize, softClipLeft, softClipRight, cutoffThreshold;
static uint fullDNALength, *upkdLengths, *softClipLtSizes, *softClipRtSizes, *DNALengths, *lengths;
static int *cutoffThresholds;
static CandidateInfo *canInfos;
static void *packedReadSeq, *packedDNASeq, *packedDNA;
#include "win_single.inc"
}

namespace half_ns {
struct SRAOccurrence { uint readID; unsigned long long ambPosition; unsigned char strand, mismatchCount; };    /* the fields pack() reads, 2bwt-flex/SRACore.h:86-95 */
struct CandidateInfo { SRAOccurrence refer; int leftOrRight; };    /* DV-DPfunctions.h:1213-1217 */
static int numOfThreads, batchSize, isDoubleStrand, peStrandLeftLeg, peStrandRightLeg, insert_high, insert_low, maxDNALength, softClipLeft, softClipRight;
static int cutoffThreshold[2];
static uint fullDNALength, *upkdReadLengths, *lengths, *startLocs, *DNALengths, *peLeftAnchorLocs, *peRightAnchorLocs, *softClipLtSizes, *softClipRtSizes;
static int *cutoffThresholds;
static CandidateInfo *canInfo;
static void *packedDNASequence, *packedReadSequence;
struct { struct { struct { void *packedDNA; } *hsp; } *sraIndex; } *index;
#include "win_half.inc"
}

namespace pair_ns {
struct CandidateInfo { uint readIDLeft; uint pos[2]; };             /* DV-DPfunctions.h:1388-1393 */
static int numOfThreads, batchSize, peStrandLeftLeg, peStrandRightLeg, insert_high, insert_low, maxDNALength, softClipLeft, softClipRight;
static int cutoffThreshold[2];
static uint fullDNALength, *upkdLengths, *lengths, *DNALengths, *peLeftAnchorLocs, *peRightAnchorLocs, *softClipLtSizes, *softClipRtSizes;
static int *cutoffThresholds, *scores[2];
static uint *hitLocs[2];
static CandidateInfo *canInfos;
static void *packedDNASeq, *packedReadSeq, *packedDNA;
#include "win_pair.inc"
}

extern "C" {

/* n candidates (readID, pos, strand) -> window start / length, clips, cutoff per candidate; returns how many were packed */
int ref_windows_single(const uint *readID, const uint *pos, const int *strand, int n, uint *readLengths, uint textLength,
                       int clipLeft, int clipRight, int cutoff, uint *outStart, uint *outLen, uint *outClipLt, uint *outClipRt, int *outCutoff)
{
    using namespace single_ns;
    numOfThreads = 0; batchSize = n; softClipLeft = clipLeft; softClipRight = clipRight; cutoffThreshold = cutoff;
    fullDNALength = textLength; upkdLengths = readLengths;
    softClipLtSizes = outClipLt; softClipRtSizes = outClipRt; DNALengths = outLen; cutoffThresholds = outCutoff;
    lengths = (uint *)malloc(sizeof(uint) * (n + 1)); canInfos = (CandidateInfo *)malloc(sizeof(CandidateInfo) * (n + 1));
    for (int i = 0; i < n; ++i) { CandidateInfo c; c.pos = pos[i]; c.readID = readID[i]; c.strand = strand[i]; ref_single_pack(c); }
    for (int i = 0; i < numOfThreads; ++i) outStart[i] = canInfos[i].pos;
    int k = numOfThreads;
    free(lengths); free(canInfos);
    return k;
}

/* n occurrences of aligned reads -> up to 2 n windows; out arrays hold 2 n entries; outOcc[w] = the occurrence a window hangs on */
int ref_windows_half(const uint *readID, const uint *pos, const uchar *strand, int n, uint *readLengths, uint textLength, int leftLeg, int rightLeg,
                     int insHigh, int insLow, int maxDNA, int clipLeft, int clipRight, int cutoff0, int cutoff1,
                     uint *outOcc, int *outLeftOrRight, uint *outStart, uint *outLen, uint *outReadLen, uint *outClipLt, uint *outClipRt,
                     uint *outAncL, uint *outAncR, int *outCutoff)
{
    using namespace half_ns;
    numOfThreads = 0; batchSize = 2 * n + 2; isDoubleStrand = (leftLeg == rightLeg); peStrandLeftLeg = leftLeg; peStrandRightLeg = rightLeg;
    insert_high = insHigh; insert_low = insLow; maxDNALength = maxDNA; softClipLeft = clipLeft; softClipRight = clipRight;
    cutoffThreshold[0] = cutoff0; cutoffThreshold[1] = cutoff1; fullDNALength = textLength; upkdReadLengths = readLengths;
    lengths = outReadLen; startLocs = outStart; DNALengths = outLen; peLeftAnchorLocs = outAncL; peRightAnchorLocs = outAncR;
    softClipLtSizes = outClipLt; softClipRtSizes = outClipRt; cutoffThresholds = outCutoff;
    canInfo = (CandidateInfo *)malloc(sizeof(CandidateInfo) * (2 * n + 2));
    for (int i = 0; i < n; ++i) {
        SRAOccurrence o; memset(&o, 0, sizeof o);
        o.readID = readID[i]; o.ambPosition = pos[i]; o.strand = strand[i]; o.mismatchCount = (unsigned char)(i & 0x7F);     /* carries the occurrence index mod 128 */
        int before = numOfThreads;
        ref_half_pack(o);
        for (int w = before; w < numOfThreads; ++w) outOcc[w] = (uint)i;
    }
    for (int w = 0; w < numOfThreads; ++w) outLeftOrRight[w] = canInfo[w].leftOrRight;
    int k = numOfThreads;
    free(canInfo);
    return k;
}

/* deep DP: left windows of n candidates (readIDLeft, posLeft, posRight), then -- given the left alignments' scores and hit
 * locations -- the right windows (entries whose left score is below its cutoff keep what packLeft wrote, as in the reference) */
int ref_windows_pair(const uint *readIDLeft, const uint *posLeft, const uint *posRight, int n, uint *readLengths, uint textLength, int leftLeg, int rightLeg,
                     int insHigh, int insLow, int maxDNA, int clipLeft, int clipRight, int cutoff0, int cutoff1,
                     uint *outStartL, uint *outLenL, uint *outClipLtL, uint *outClipRtL, uint *outAncLL, uint *outAncRL, int *outCutoffL,
                     const int *scoresLeft, const uint *hitLocsLeft,
                     uint *outStartR, uint *outLenR, uint *outReadLenR, uint *outClipLtR, uint *outClipRtR, uint *outAncLR, uint *outAncRR, int *outCutoffR)
{
    using namespace pair_ns;
    numOfThreads = 0; batchSize = n; peStrandLeftLeg = leftLeg; peStrandRightLeg = rightLeg; insert_high = insHigh; insert_low = insLow;
    maxDNALength = maxDNA; softClipLeft = clipLeft; softClipRight = clipRight; cutoffThreshold[0] = cutoff0; cutoffThreshold[1] = cutoff1;
    fullDNALength = textLength; upkdLengths = readLengths;
    lengths = (uint *)malloc(sizeof(uint) * (n + 1)); canInfos = (CandidateInfo *)malloc(sizeof(CandidateInfo) * (n + 1));
    DNALengths = outLenL; peLeftAnchorLocs = outAncLL; peRightAnchorLocs = outAncRL; softClipLtSizes = outClipLtL; softClipRtSizes = outClipRtL; cutoffThresholds = outCutoffL;
    for (int i = 0; i < n; ++i) { CandidateInfo c; c.readIDLeft = readIDLeft[i]; c.pos[0] = posLeft[i]; c.pos[1] = posRight[i]; ref_pair_packLeft(c); }
    for (int i = 0; i < numOfThreads; ++i) outStartL[i] = canInfos[i].pos[0];
    /* packRight overwrites the batch arrays in place (DV-DPfunctions.cu:3420-3470): give it copies of the left pass's values */
    memcpy(outLenR, outLenL, sizeof(uint) * n); memcpy(outClipLtR, outClipLtL, sizeof(uint) * n); memcpy(outClipRtR, outClipRtL, sizeof(uint) * n);
    memcpy(outAncLR, outAncLL, sizeof(uint) * n); memcpy(outAncRR, outAncRL, sizeof(uint) * n); memcpy(outCutoffR, outCutoffL, sizeof(int) * n);
    memcpy(outReadLenR, lengths, sizeof(uint) * n);
    DNALengths = outLenR; peLeftAnchorLocs = outAncLR; peRightAnchorLocs = outAncRR; softClipLtSizes = outClipLtR; softClipRtSizes = outClipRtR; cutoffThresholds = outCutoffR;
    uint *keepLengths = lengths; lengths = outReadLenR;
    scores[0] = (int *)scoresLeft; hitLocs[0] = (uint *)hitLocsLeft;
    ref_pair_packRight();
    for (int i = 0; i < numOfThreads; ++i) outStartR[i] = canInfos[i].pos[1];
    int k = numOfThreads;
    free(keepLengths); free(canInfos);
    return k;
}

}
