/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * The reference's own paired-end seed-hit join, run on the host: the radix sort macros (DV-DPfunctions.h:60-95),
 * DeepDP_Space::SeedPos / CandidateInfo (:1382-1393), DP2_DIVIDE_GAP (:1413), DP2_MARGIN (DV-DPfunctions.cu:2549),
 * findRevStart (:2626-2653) and pairEndMerge (:2780-2880), cut out of the reference files by sed at build time into
 * seed_pair.inc (the two methods become free functions that read the batch members they use -- readLengths,
 * insert_low, insert_high -- from the file-level variables below; nothing else changes).  The lines of
 * decodePositions / decodeMergePositions around them (two sorts per side, the two merge calls, the final sort,
 * :2963-2999) are spelled here.  It pins oracle/seed_oracle.c's s3o_seed_pair_candidates.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
using std::vector;
typedef unsigned int uint;
static uint *readLengths;
static int insert_low, insert_high;
#include "seed_pair.inc"

extern "C" {

/* hits of the read side and of the mate side, (strandIndex << 31 | readID, pos) in arrival order.  Returns the number
 * of candidates (readIDLeft, posLeft, posRight). */
uint ref_seed_pair_merge(const uint *keyR, const uint *posR, uint nR, const uint *keyM, const uint *posM, uint nM,
                         uint *lengthsByReadID, int insLow, int insHigh, int peStrandLeftLeg, int peStrandRightLeg,
                         uint *outID, uint *outPosL, uint *outPosR)
{
    readLengths = lengthsByReadID; insert_low = insLow; insert_high = insHigh;
    SeedPos *side[2];
    uint len[2];
    const uint *keys[2] = {keyR, keyM}, *poss[2] = {posR, posM};
    const uint ns[2] = {nR, nM};
    for (int s = 0; s < 2; ++s) {
        SeedPos *pos = (SeedPos *)malloc(((size_t)ns[s] + 2) * sizeof(SeedPos)), *auxPos = (SeedPos *)malloc(((size_t)ns[s] + 2) * sizeof(SeedPos));
        for (uint i = 0; i < ns[s]; ++i) { pos[i].strand_readID = keys[s][i]; pos[i].pos = poss[s][i]; }
        pos[ns[s]].strand_readID = 0x7FFFFFFFu; pos[ns[s]].pos = 0xFFFFFFFFu;                 /* array guards, :2961-2962 */
        pos[ns[s] + 1].strand_readID = 0x7FFFFFFFu | (1u << 31); pos[ns[s] + 1].pos = 0xFFFFFFFFu;
        len[s] = ns[s] + 2;
        uint l = len[s];
        MC_RadixSort_32_16 ( pos, pos, auxPos, l );
        MC_RadixSort_32_16 ( pos, strand_readID, auxPos, l );
        free(auxPos);
        side[s] = pos;
    }
    SeedPos *readArr[2], *mateArr[2];
    readArr[0] = side[0]; mateArr[0] = side[1];
    readArr[1] = side[0] + ref_findRevStart(side[0], len[0]);
    mateArr[1] = side[1] + ref_findRevStart(side[1], len[1]);
    vector<CandidateInfo> *canInfo = new vector<CandidateInfo>;
    ref_pairEndMerge(canInfo, readArr[peStrandLeftLeg - 1], mateArr[peStrandRightLeg - 1], 0);
    ref_pairEndMerge(canInfo, mateArr[peStrandLeftLeg - 1], readArr[peStrandRightLeg - 1], 1);
    free(side[0]); free(side[1]);
    vector<CandidateInfo> &candArr = *canInfo;
    uint arrLength = (uint)candArr.size();
    CandidateInfo *auxCandArr = (CandidateInfo *)malloc(((size_t)arrLength + 1) * sizeof(CandidateInfo));
    MC_RadixSort_32_16 ( candArr, readIDLeft, auxCandArr, arrLength );
    free(auxCandArr);
    for (uint i = 0; i < arrLength; ++i) { outID[i] = candArr[i].readIDLeft; outPosL[i] = candArr[i].pos[0]; outPosR[i] = candArr[i].pos[1]; }
    delete canInfo;
    return arrLength;
}

}
