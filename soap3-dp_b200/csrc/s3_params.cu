// Host-side tables of the DP stages: where the seeds of a read lie, and the per-stage DP parameters.
//
// Replaces getSeedPositions (definitions.h:323-442) and getParameterFor{SingleDP, DefaultDP, NewDefaultDP, DeepDP}
// (CPUfunctions.cpp:59-260; round 2 of deep DP: DV-DPForBothUnalign.cu:138-139).  Pure integer tables (the only floating
// point is the reference's own `(int)(readLength * ratio)` and `ceil(0.3 * readLength)`, evaluated in double like there);
// callers use them to cut seeds for s3_search / s3_seed_candidates and to fill cutoffThresholds / maxPerRange.
#include "s3_common.cuh"
#include "../../include/soap3dp_b200.h"

#include <cmath>
#include <cstring>

namespace {

// seed lengths by read length: single-end / new-default DP (definitions.h:221-240), deep DP rounds 1 and 2 (:186-204)
int single_seed_length(int len) { return len > 300 ? 70 : len > 80 ? 38 : len > 60 ? 32 : len > 40 ? 26 : 22; }
int single_tail_trim(int len)   { return len > 80 ? 10 : len > 60 ? 4 : len > 40 ? 4 : 0; }
int single_max_hits(int len)    { return len > 300 ? 4 : len > 80 ? 10 : len > 60 ? 20 : len > 40 ? 30 : 40; }
int deep_seed_length(int len, int round)
{
    if (round == 1) return len > 150 ? 45 : len > 80 ? 26 : len > 60 ? 24 : len > 40 ? 22 : 20;
    return len > 150 ? 52 : len > 80 ? 30 : len > 60 ? 28 : len > 40 ? 26 : 24;
}
int dp_cutoff(uint32_t len) { return (int)std::ceil(0.3 * (double)len); }     // DP_SCORE_THRESHOLD_RATIO, definitions.h:169

}  // namespace

extern "C" int s3_seed_layout(int stage, int32_t readLength, int32_t *seedLength, int32_t *seedPositions, int32_t capacity, int32_t *seedNum)
{
    if (!seedLength || !seedPositions || !seedNum || capacity < 1) { s3_set_error("s3_seed_layout: NULL argument or no room"); return S3_EINVAL; }
    if (readLength < 1) { s3_set_error("s3_seed_layout: readLength %d", readLength); return S3_EINVAL; }
    int n = 0;
    if (stage == S3_STAGE_SINGLE_DP || stage == S3_STAGE_NEW_DEFAULT_DP) {
        *seedLength = single_seed_length(readLength);
        n = readLength > 120 ? 3 + readLength / 100 : 3;                        // SEED_NUM_SINGLE_DP + one per 100 bases
        if (n > capacity) { s3_set_error("s3_seed_layout: %d seeds, room for %d", n, capacity); return S3_EINVAL; }
        int H = readLength > 300 ? (int)(readLength * 0.15) : 0;
        int X = readLength > 300 ? (int)(readLength * 0.15) : single_tail_trim(readLength);
        int apart = (readLength - X - H) / n;
        for (int i = 0; i < n; ++i) seedPositions[i] = H + i * apart;
        if (seedPositions[n - 1] > readLength - *seedLength - X) seedPositions[n - 1] = readLength - *seedLength - X;
    } else if (stage == S3_STAGE_DEEP_DP_ROUND1 || stage == S3_STAGE_DEEP_DP_ROUND2) {
        *seedLength = deep_seed_length(readLength, stage == S3_STAGE_DEEP_DP_ROUND1 ? 1 : 2);
        int H = 0, T = 0;
        if (readLength > 150) { H = (int)(readLength * 0.1); T = (int)(readLength * 0.2); }
        // from the tail towards the head, half a seed apart; a last seed at the head offset if the walk stopped short
        for (int i = readLength - *seedLength - T; i >= H; i -= *seedLength / 2) {
            if (n >= capacity) { s3_set_error("s3_seed_layout: more than %d seeds", capacity); return S3_EINVAL; }
            seedPositions[n++] = i;
        }
        if (n == 0) {
            // the reference reads seedPositions[-1] here (definitions.h:406, a read shorter than its seed); no seed fits
            *seedNum = 0;
            return S3_OK;
        }
        if (seedPositions[n - 1] > H) {
            if (n >= capacity) { s3_set_error("s3_seed_layout: more than %d seeds", capacity); return S3_EINVAL; }
            seedPositions[n++] = H;
        }
    } else {
        s3_set_error("s3_seed_layout: stage %d has no seeds (definitions.h:317-321: 1 single, 3 new default, 4 / 5 deep DP rounds)", stage);
        return S3_EINVAL;
    }
    *seedNum = n;
    return S3_OK;
}

extern "C" int s3_dp_stage_parameters(int stage, uint32_t readLength, uint32_t readLength2, int isDefaultThreshold,
                                      int32_t dpScoreThreshold, int32_t maxFrontLenClipped, int32_t maxEndLenClipped,
                                      s3_dp_stage_params *out)
{
    if (!out) { s3_set_error("s3_dp_stage_parameters: NULL argument"); return S3_EINVAL; }
    if (stage < S3_STAGE_SINGLE_DP || stage > S3_STAGE_DEEP_DP_ROUND2) { s3_set_error("s3_dp_stage_parameters: stage %d", stage); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    const uint32_t len[2] = {readLength, readLength2};
    const int ends = stage == S3_STAGE_SINGLE_DP ? 1 : 2;
    for (int e = 0; e < ends; ++e)
        out->paramRead[e].cutoffThreshold = isDefaultThreshold == 1 ? dp_cutoff(len[e]) : dpScoreThreshold;
    out->softClipLeft = maxFrontLenClipped;
    out->softClipRight = maxEndLenClipped;
    for (int e = 0; e < ends; ++e) {
        s3_dp_read_params &p = out->paramRead[e];
        const uint32_t l = len[e];
        switch (stage) {
        case S3_STAGE_DEFAULT_DP:                       // CPUfunctions.cpp:59-89
            p.maxHitNum = l > 50 ? 50 : 70;
            break;
        case S3_STAGE_NEW_DEFAULT_DP:                   // :91-132
            p.maxHitNum = l > 50 ? 150 : 200;
            p.seedLength = l > 75 ? 26 : l > 50 ? 24 : 22;
            break;
        case S3_STAGE_DEEP_DP_ROUND1:                   // :135-189
        case S3_STAGE_DEEP_DP_ROUND2:                   // + DV-DPForBothUnalign.cu:138-139: only maxHitNum changes
            p.maxHitNum = stage == S3_STAGE_DEEP_DP_ROUND2 ? 1000 : l > 50 ? 100 : 150;
            p.seedLength = l > 150 ? 45 : l > 80 ? 26 : l > 60 ? 24 : l > 40 ? 22 : 20;
            p.sampleDist = (int)(p.seedLength * 0.5);
            break;
        default:                                        // S3_STAGE_SINGLE_DP, :191-256
            p.maxHitNum = single_max_hits((int)(l > 0x7FFFFFFFu ? 0x7FFFFFFF : l));
            p.seedLength = single_seed_length((int)(l > 0x7FFFFFFFu ? 0x7FFFFFFF : l));
            break;
        }
    }
    if (stage == S3_STAGE_SINGLE_DP) {
        // singleDPSeedNum adds a seed per 100 bases from 101 on (getSeedPositions, which the seeding engines use, from 121
        // on); the three seed offsets are what the reference calls obsolete but still fills, in its mixed int / uint arithmetic
        out->singleDPSeedNum = readLength > 100 ? 3 + (int)(readLength / 100) : 3;
        int X = single_tail_trim((int)(readLength > 0x7FFFFFFFu ? 0x7FFFFFFF : readLength));
        out->singleDPSeedPos[0] = 0;
        out->singleDPSeedPos[2] = (int)((readLength - (uint32_t)X) * 0.5 - 1);
        if ((uint32_t)out->singleDPSeedPos[2] > readLength - (uint32_t)out->paramRead[0].seedLength)
            out->singleDPSeedPos[2] = (int)(readLength - (uint32_t)out->paramRead[0].seedLength);
        out->singleDPSeedPos[1] = (out->singleDPSeedPos[0] + out->singleDPSeedPos[2]) / 2 - 1;
    }
    return S3_OK;
}
