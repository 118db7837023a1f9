/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * The reference's own stage tables on the host: getSeedPositions straight from the reference's unmodified
 * definitions.h (:323-442, included through IniParam.h) and getParameterFor{AllDP,DefaultDP,NewDefaultDP,DeepDP,SingleDP}
 * (CPUfunctions.cpp:46-260), cut by sed at build time into params.inc; the structs they fill (DPParameters PEAlgnmt.h:349,
 * InputOptions / IniParams IniParam.h) are the reference's own.  Pins s3_seed_layout / s3_dp_stage_parameters.
 */
#include <math.h>
#include <string.h>
#include "IniParam.h"
#include "params.inc"

extern "C" {

int ref_seed_positions(int stage, int readLength, int *seedLength, int *seedPositions, int *seedNum)
{
    *seedNum = 0; *seedLength = 0;
    getSeedPositions(stage, readLength, seedLength, seedPositions, seedNum);
    return 0;
}

/* out: softClipLeft, softClipRight, tailTrimLen, singleDPSeedNum, singleDPSeedPos[10], then per read cutoffThreshold,
 * maxHitNum, sampleDist, seedLength (the struct starts zeroed, so fields a stage leaves alone read 0) */
int ref_stage_parameters(int stage, unsigned readLength, unsigned readLength2, int isDefaultThreshold, int dpScoreThreshold,
                         int maxFront, int maxEnd, int *out)
{
    DPParameters dp; InputOptions io; IniParams ini;
    memset(&dp, 0, sizeof dp); memset(&io, 0, sizeof io); memset(&ini, 0, sizeof ini);
    ini.Ini_isDefaultThreshold = isDefaultThreshold; ini.Ini_DPScoreThreshold = dpScoreThreshold;
    ini.Ini_maxFrontLenClipped = maxFront; ini.Ini_maxEndLenClipped = maxEnd;
    switch (stage) {
    case STAGE_SINGLE_DP: getParameterForSingleDP(dp, io, ini, readLength); break;
    case STAGE_DEFAULT_DP: getParameterForDefaultDP(dp, io, ini, readLength, readLength2); break;
    case STAGE_NEW_DEFAULT_DP: getParameterForNewDefaultDP(dp, io, ini, readLength, readLength2); break;
    case STAGE_DEEP_DP_ROUND1: getParameterForDeepDP(dp, io, ini, readLength, readLength2); break;
    case STAGE_DEEP_DP_ROUND2:
        getParameterForDeepDP(dp, io, ini, readLength, readLength2);
        dp.paramRead[0].maxHitNum = MAX_SEED_HITS_DEEP_DP_FOR_NORMAL_READ_2;        /* DV-DPForBothUnalign.cu:138-139 */
        dp.paramRead[1].maxHitNum = MAX_SEED_HITS_DEEP_DP_FOR_NORMAL_READ_2;
        break;
    default: return -1;
    }
    int k = 0;
    out[k++] = dp.softClipLeft; out[k++] = dp.softClipRight; out[k++] = dp.tailTrimLen; out[k++] = dp.singleDPSeedNum;
    for (int i = 0; i < 10; ++i) out[k++] = dp.singleDPSeedPos[i];
    for (int e = 0; e < 2; ++e) { out[k++] = dp.paramRead[e].cutoffThreshold; out[k++] = dp.paramRead[e].maxHitNum; out[k++] = dp.paramRead[e].sampleDist; out[k++] = dp.paramRead[e].seedLength; }
    return k;
}

}
