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

// =====================================================================================================================
// Mapping qualities.  Replaces the nine MAPQ functions of the SAM writers (BGS-IO.cpp:2280-2580) with their tables
// (:33, :42, :45) and bwase_initialize (CPUfunctions.cpp:3014-3019).  Scalar host functions: the arithmetic is the
// reference's, in the reference's types (float where it uses float, double where it uses double), so the (int)
// truncations land on the same side.
// =====================================================================================================================
namespace {

const double kMappingScore[6][2] = {{1.0, 1.0}, {0.875, 0.85}, {0.75, 0.7}, {0.625, 0.55}, {0.475, 0.4}, {0.325, 0.25}};   // BGS-IO.cpp:33
const float kPenaltyAvgMisQual[41] = {3, 2.85, 2.71, 2.57, 2.43, 2.3, 2.17, 2.04, 1.92, 1.8, 1.69, 1.58, 1.47, 1.37, 1.27, 1.17, 1.08, 0.99, 0.91,
                                      0.83, 0.75, 0.68, 0.61, 0.54, 0.48, 0.42, 0.37, 0.32, 0.27, 0.23, 0.19, 0.15, 0.12, 0.09, 0.07, 0.05, 0.03,
                                      0.02, 0.01, 0, 0};                                                                      // BGS-IO.cpp:42

// penalty_ratio_x1[x] (BGS-IO.cpp:45) is 1 / (x + 1) to two decimals, halves rounded up
float penalty_ratio_x1(int x) { return (float)(std::floor(100.0 / (x + 1) + 0.5) / 100.0); }

// g_log_n[i] = (int)(4.343 * log(i) + 0.5), i = 1..255 (CPUfunctions.cpp:3018); entry 0 is never set by the reference
int log_n(int i) { return i < 1 ? 0 : (int)(4.343 * std::log((double)i) + 0.5); }

int table_score(int index, int avgMismatchQual, int maxMAPQ, int minMAPQ)
{
    if (index > 5) index = 5;
    int q = (avgMismatchQual - 1) / 20;
    if (q > 1) q = 1; else if (q < 0) q = 0;
    int s = (int)(maxMAPQ * kMappingScore[index][q]);
    return s < minMAPQ ? minMAPQ : s;
}

int bwa_single(int x0, int x1)                                    // bwaLikeSingleQualScore, BGS-IO.cpp:2311
{
    if (x0 > 1) return 0;
    if (x1 == 0) return 37;
    if (x1 > 255) x1 = 255;
    const int n = log_n(x1);
    return 23 < n ? 0 : 23 - n;
}

}  // namespace

extern "C" int32_t s3_mapq_unique(int n, int mismatchNum, int avgMismatchQual, int maxMAPQ, int minMAPQ)       // getMapQualScore :2280
{
    return n == 1 ? table_score(mismatchNum, avgMismatchQual, maxMAPQ, minMAPQ) : minMAPQ;
}

extern "C" int32_t s3_mapq_bwa_single(int x0, int x1) { return bwa_single(x0, x1); }

extern "C" int32_t s3_mapq_single(int mismatchNum, int avgMismatchQual, int x0, int x1, int maxMAPQ, int minMAPQ, int isBWALike)   // :2331
{
    if (isBWALike) return bwa_single(x0, x1);
    if (x0 != 1 || x1 > 0) return minMAPQ;
    return table_score(mismatchNum, avgMismatchQual, maxMAPQ, minMAPQ);
}

extern "C" int32_t s3_mapq_single_dp(int maxDPScore, int avgMismatchQual, int x0, int x1_t1, int x1_t2, int bestDPScore, int secondBestDPScore,
                                     int maxMAPQ, int minMAPQ, int dpThres, int isBWALike)                        // getMapQualScoreForSingleDP :2370
{
    if (isBWALike) return bwa_single(x0, x1_t1 + x1_t2);
    if (x0 > 1 || x1_t1 > 0) return minMAPQ;
    float R1, R2, R3, P;
    if (x1_t2 > 0) R1 = 1.0 - ((float)(secondBestDPScore - dpThres)) / (0.7 * bestDPScore - dpThres);
    else R1 = 1.0;
    const int x1 = x1_t1 + x1_t2;
    R2 = penalty_ratio_x1(x1 > 100 ? 100 : x1);
    R3 = ((float)(bestDPScore - dpThres)) / (maxDPScore - dpThres);
    if (avgMismatchQual < 0) avgMismatchQual = 0; else if (avgMismatchQual > 40) avgMismatchQual = 40;
    P = kPenaltyAvgMisQual[avgMismatchQual];
    int s = (int)(maxMAPQ * R1 * R2 * R3 - P);
    return s < minMAPQ ? minMAPQ : s;
}

extern "C" void s3_mapq_bwa_pair(int x0_0, int x1_0, int x0_1, int x1_1, int op_score, int op_num, int subop_score, int subop_num,
                                 int readlen_0, int readlen_1, int32_t *mapScore0, int32_t *mapScore1)             // bwaLikePairQualScore :2415
{
    int mapq0 = bwa_single(x0_0, x1_0), mapq1 = bwa_single(x0_1, x1_1);
    op_score *= 10; subop_score *= 10;
    int mapq_p = 0;
    if (mapq0 > 0 && mapq1 > 0) {
        mapq_p = mapq0 + mapq1;
        if (mapq_p > 60) mapq_p = 60;
        mapq0 = mapq1 = mapq_p;
    } else {
        if (op_num == 1) {
            if (subop_num == 0) mapq_p = 29;
            else if (op_score - subop_score > (0.3 * ((readlen_0 + readlen_1) / 2))) mapq_p = 23;
            else {
                if (subop_num > 255) subop_num = 255;
                mapq_p = (op_score - subop_score) / 2 - log_n(subop_num);
                if (mapq_p < 0) mapq_p = 0;
            }
        }
        if (mapq0 == 0) mapq0 = (mapq_p + 7 < mapq1) ? mapq_p + 7 : mapq1;
        if (mapq1 == 0) mapq1 = (mapq_p + 7 < mapq0) ? mapq_p + 7 : mapq0;
    }
    if (mapScore0) *mapScore0 = mapq0;
    if (mapScore1) *mapScore1 = mapq1;
}

extern "C" int32_t s3_mapq_pair_end(int mismatchNum, int avgMismatchQual, int x0, int x1, int isBestHit, uint32_t totalNumValidPairs,
                                    int maxMAPQ, int minMAPQ)                                                      // getMapQualScore2 :2465
{
    if (x0 != 1 || totalNumValidPairs != 1) return minMAPQ;
    if (isBestHit == 0 && x1 > 1) return minMAPQ;
    return table_score(mismatchNum, avgMismatchQual, maxMAPQ, minMAPQ);
}

extern "C" int32_t s3_mapq_unique_dp(int n, int dpScore, int maxDPScore, int avgMismatchQual, int maxMAPQ, int minMAPQ)   // getMapQualScoreForDP :2500
{
    if (n != 1) return minMAPQ;
    int idx = 0;
    if (dpScore < maxDPScore) idx = (int)((1.0 - (double)dpScore / maxDPScore) * 100.0 - 1.0) / 5 + 1;
    return table_score(idx, avgMismatchQual, maxMAPQ, minMAPQ);
}

extern "C" int32_t s3_mapq_pair_end_dp(int dpScore, int maxDPScore, int avgMismatchQual, int x0, int x1, int bestDPScore, int secondBestDPScore,
                                       int isBestHit, int totalNumValidPairs, int maxMAPQ, int minMAPQ)           // getMapQualScoreForDP2 :2534
{
    if (x0 != 1 || totalNumValidPairs != 1) return minMAPQ;
    if (isBestHit == 0 && x1 > 1) return minMAPQ;
    int idx = 0;
    if (dpScore < maxDPScore) idx = (int)((1.0 - (double)dpScore / maxDPScore) * 100.0 - 1.0) / 4 + 1;
    int s = table_score(idx, avgMismatchQual, maxMAPQ, minMAPQ);
    if (bestDPScore > secondBestDPScore && ((double)bestDPScore - secondBestDPScore) / maxDPScore < 0.2) s = minMAPQ;
    return s < minMAPQ ? minMAPQ : s;
}

extern "C" int32_t s3_mapq_of_pair(int score1, int score2)                                                        // getMapQualScoreForPair :2577
{
    return score1 > score2 ? (int)(score1 * 0.2 + score2 * 0.8) : (int)(score1 * 0.8 + score2 * 0.2);
}
