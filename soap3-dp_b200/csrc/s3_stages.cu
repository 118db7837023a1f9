// s3_stages.cu -- the two DP stages that start from seeds, for a batch of reads / read pairs.
//
//   s3_single_dp_align   DPForUnalignSingle2 (DV-DPForSingleReads.cu:155; SingleDPWrapper::run :141): seeds of every read
//                        (getSeedPositions, STAGE_SINGLE_DP) -> seeding driver -> candidate positions (decodePositions +
//                        singleMerge) -> window per candidate (SingleEndAlgnBatch::pack) -> DP -> SingleAlgnmtResult
//   s3_deep_dp_align     DPForUnalignPairs2 (DV-DPForBothUnalign.cu:245; DeepDPWrapper::run2 :226, seeding_ext :131-143):
//                        seeds of both mates (STAGE_DEEP_DP_ROUND1; ROUND2 with the larger hit limit for the pairs whose
//                        seeds had too many hits and that found no candidate) -> seeding driver per side -> candidate
//                        position pairs (decodeMergePositions) -> left window, DP, right window cut by the left hit
//                        (packLeft / packRight), DP -> DeepDPAlignResult for candidates whose two reads reach their cutoffs
//
// Both are orchestration, like the reference's wrappers: every step is one of the library's entries (s3_seed_layout,
// s3_dp_stage_parameters, s3_seed_search, s3_seed_candidates / s3_seed_pair_candidates, s3_dp_make_windows,
// s3_dp_align_windows), each verified against its own oracle; host arrays travel between them, as they do between the
// reference's engine threads.  The batches here are the small remainder of a run (reads the search left unaligned).
#include "s3_common.cuh"
#include "../../include/soap3dp_b200.h"

#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

struct SeedSet {                                   // one side's seeding batch (DV-DPfunctions.cu:2655-2680 / 1057-1080)
    std::vector<uint32_t> words, lengths, readIDs, offsets, maxHit;
    uint32_t wordPerSeed = 0;
    uint64_t n = 0;
};

// base k of read r in the query buffer: bits 2 (k % 16) of word k / 16 (QueryParser.cpp:1146)
inline uint32_t read_base(const uint32_t *queries, uint32_t wpq, uint32_t r, uint32_t k)
{
    return (queries[(size_t)(r / 32) * 32 * wpq + (size_t)(k >> 4) * 32 + r % 32] >> ((k & 15u) << 1)) & 3u;
}

void seed_set_reserve(SeedSet &s, uint32_t wordPerSeed, size_t seeds)
{
    s.wordPerSeed = wordPerSeed;
    s.words.assign(((seeds + 31) / 32 * 32 + 32) * wordPerSeed, 0u);
    s.lengths.assign((seeds + 31) / 32 * 32 + 32, 0u);
    s.readIDs.clear(); s.offsets.clear(); s.maxHit.clear();
    s.n = 0;
}

void seed_set_add(SeedSet &s, const uint32_t *queries, uint32_t wpq, uint32_t readID, uint32_t keyID, uint32_t off, uint32_t len, uint32_t maxHit)
{
    const uint64_t id = s.n++;
    uint32_t *dst = s.words.data() + (id / 32) * 32 * s.wordPerSeed + id % 32;
    for (uint32_t i = 0; i < len; ++i) dst[(size_t)(i >> 4) * 32] |= read_base(queries, wpq, readID, off + i) << ((i & 15u) << 1);
    s.lengths[id] = len;
    s.readIDs.push_back(keyID); s.offsets.push_back(off); s.maxHit.push_back(maxHit);
}

// traceback pattern -> CIGAR runs in read order, length << 8 | op (CigarStringEncoder, DV-DPfunctions.h:545-597, as the
// engines' result loops drive it; the same rule as s3_pe_runs_kernel and s3_dp_decode)
void pattern_runs(const uint8_t *p, size_t cap, std::vector<uint32_t> &out)
{
    std::vector<uint32_t> rev;
    uint8_t last = 'N', curType = 0;
    int curCnt = 0;
    bool have = false;
    const uint8_t *end = p + cap;
    for (; p < end && *p != 0; ++p) {
        uint8_t type; int cnt;
        if (*p == 'V') { if (++p >= end) break; type = last; cnt = (int)*p - 1; }
        else { type = last = *p; cnt = 1; }
        if (have && curType == type) curCnt += cnt;
        else {
            if (have && curCnt > 0 && curType != 'N') rev.push_back(((uint32_t)curCnt << 8) | curType);
            curType = type; curCnt = cnt; have = true;
        }
    }
    if (have && curCnt > 0 && curType != 'N') rev.push_back(((uint32_t)curCnt << 8) | curType);
    out.insert(out.end(), rev.rbegin(), rev.rend());
}

struct Aligned {                                   // outputs of one s3_dp_align_windows call
    std::vector<int32_t> score;
    std::vector<uint32_t> hit, cnt;
    std::vector<uint8_t> pattern;
    uint32_t patLen = 0;
};

struct Windows {
    std::vector<uint32_t> cand, readID, start, len, clipLt, clipRt, ancL, ancR;
    std::vector<uint8_t> strand, lor;
    std::vector<int32_t> cutoff;
    uint64_t n = 0;
    void resize(size_t k) { cand.resize(k); readID.resize(k); start.resize(k); len.resize(k); clipLt.resize(k); clipRt.resize(k); ancL.resize(k); ancR.resize(k);
                            strand.resize(k); lor.resize(k); cutoff.resize(k); }
};

int make_windows(s3_index *ix, int mode, const s3_window_params &wp, const uint32_t *readLengths, uint64_t numReads, const uint32_t *ids, const uint32_t *pos,
                 const uint32_t *pos2, const uint8_t *strands, const int32_t *lsc, const uint32_t *lst, const uint32_t *lhit, uint64_t n, Windows &w)
{
    w.resize(n ? n : 1);
    return s3_dp_make_windows(ix, mode, &wp, readLengths, numReads, ids, pos, pos2, strands, lsc, lst, lhit, n, w.cand.data(), w.readID.data(), w.strand.data(),
                              w.lor.data(), w.start.data(), w.len.data(), w.clipLt.data(), w.clipRt.data(), w.ancL.data(), w.ancR.data(), w.cutoff.data(), &w.n);
}

int align_windows(s3_dp *dp, s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wpq, Windows &w, Aligned &a)
{
    a.patLen = s3_dp_pattern_length(dp);
    const size_t n = w.n ? w.n : 1;
    a.score.assign(n, 0); a.hit.assign(n, 0); a.cnt.assign(n, 0); a.pattern.assign(n * a.patLen, 0);
    if (!w.n) return S3_OK;
    return s3_dp_align_windows(dp, ix, queries, readLengths, numReads, wpq, w.readID.data(), w.strand.data(), w.start.data(), w.len.data(), w.cutoff.data(),
                               a.score.data(), a.hit.data(), a.cnt.data(), a.pattern.data(), (uint32_t)w.n, w.clipLt.data(), w.clipRt.data(),
                               w.ancL.data(), w.ancR.data());
}

uint32_t margin_of(uint32_t len) { return len > 100u ? len >> 2 : 25u; }

template <typename T> T *to_malloc(const std::vector<T> &v)
{
    T *p = (T *)malloc((v.size() ? v.size() : 1) * sizeof(T));
    if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

}  // namespace

extern "C" void s3_single_dp_result_free(s3_single_dp_result *r)
{
    if (!r) return;
    free(r->hits); free(r->runs); free(r->unseeded);
    memset(r, 0, sizeof *r);
}
extern "C" void s3_deep_dp_result_free(s3_deep_dp_result *r)
{
    if (!r) return;
    free(r->hits); free(r->runs); free(r->unseeded);
    memset(r, 0, sizeof *r);
}

extern "C" int s3_single_dp_align(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery,
                                  const uint32_t *readIDs, uint64_t n, const s3_stage_params *par, s3_single_dp_result *out)
{
    if (!out) { s3_set_error("s3_single_dp_align: NULL result"); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    if (!ix || !queries || !readLengths || !par || (n && !readIDs)) { s3_set_error("s3_single_dp_align: NULL argument"); return S3_EINVAL; }
    out->numReads = n;
    if (n == 0) return S3_OK;
    uint32_t maxLen = 0;
    for (uint64_t k = 0; k < n; ++k) {
        if (readIDs[k] >= numReads) { s3_set_error("s3_single_dp_align: read id %u out of range", readIDs[k]); return S3_EINVAL; }
        if (readLengths[readIDs[k]] > maxLen) maxLen = readLengths[readIDs[k]];
    }
    int rc;
    // ---- seeds (SingleEndSeedingBatch::packSeeds, DV-DPfunctions.cu:1082-1100)
    SeedSet seeds;
    seed_set_reserve(seeds, (maxLen + 15) / 16, n * 16);
    std::vector<int32_t> seedPos(maxLen + 16);
    for (uint64_t k = 0; k < n; ++k) {
        const uint32_t r = readIDs[k], len = readLengths[r];
        int32_t seedLen = 0, seedNum = 0;
        if ((rc = s3_seed_layout(S3_STAGE_SINGLE_DP, (int32_t)len, &seedLen, seedPos.data(), (int32_t)seedPos.size(), &seedNum))) return rc;
        s3_dp_stage_params sp;
        if ((rc = s3_dp_stage_parameters(S3_STAGE_SINGLE_DP, len, 0, par->isDefaultThreshold, par->dpScoreThreshold, par->softClipLeft, par->softClipRight, &sp))) return rc;
        if (seeds.n + (uint64_t)seedNum > seeds.lengths.size() - 32) { s3_set_error("s3_single_dp_align: more than 16 seeds per read"); return S3_EINVAL; }
        for (int32_t j = 0; j < seedNum; ++j) seed_set_add(seeds, queries, wordPerQuery, r, r, (uint32_t)seedPos[j], (uint32_t)seedLen, (uint32_t)sp.paramRead[0].maxHitNum);
    }
    out->numSeeds = seeds.n;
    // ---- seeding driver, candidate positions (decodePositions + singleMerge, DV-DPfunctions.cu:1101-1219)
    s3_seed_search_result sr;
    if ((rc = s3_seed_search(ix, seeds.words.data(), seeds.lengths.data(), seeds.n, seeds.wordPerSeed, seeds.maxHit.data(), &sr))) return rc;
    std::vector<uint32_t> rid(sr.total), off(sr.total), slen(sr.total), rlen(sr.total);
    std::vector<int32_t> strand(sr.total);
    for (uint64_t s = 0; s < seeds.n; ++s)
        for (uint64_t g = sr.offsets[s]; g < sr.offsets[s + 1]; ++g) {
            rid[g] = seeds.readIDs[s]; off[g] = seeds.offsets[s]; slen[g] = seeds.lengths[s]; rlen[g] = readLengths[seeds.readIDs[s]]; strand[g] = sr.strand[g];
        }
    uint32_t *cR = NULL, *cP = NULL; int32_t *cS = NULL;
    uint64_t nc = 0;
    rc = s3_seed_candidates(ix, sr.saL, sr.saR, strand.data(), rid.data(), off.data(), slen.data(), rlen.data(), sr.total, 0xFFFFFFFFu, &cR, &cP, &cS, &nc);
    s3_seed_search_result_free(&sr);
    if (rc) return rc;
    out->numCandidates = nc;
    // reads without a candidate (alignFlags XOR inputFlags, DV-DPfunctions.cu:1300-1310)
    std::vector<uint8_t> seeded(numReads, 0);
    for (uint64_t c = 0; c < nc; ++c) seeded[cR[c]] = 1;
    std::vector<uint32_t> unseeded;
    for (uint64_t k = 0; k < n; ++k) if (!seeded[readIDs[k]]) unseeded.push_back(readIDs[k]);
    std::vector<s3_dp_hit> hits;
    std::vector<uint32_t> runs;
    if (nc) {
        // ---- windows (SingleEndAlgnBatch::pack), DP, results (SingleDP_Space::algnmtCPUThread, DV-DPfunctions.cu:1699-1733)
        const uint32_t maxRead = (maxLen / 4 + 1) * 4, maxDNA = maxRead + 2 * margin_of(maxRead) + 8;         // DV-DPfunctions.cu:1580-1581
        s3_window_params wp;
        memset(&wp, 0, sizeof wp);
        wp.strandLeftLeg = 1; wp.strandRightLeg = 2; wp.softClipLeft = par->softClipLeft; wp.softClipRight = par->softClipRight;
        wp.cutoffThreshold[0] = wp.cutoffThreshold[1] = par->isDefaultThreshold ? -1 : par->dpScoreThreshold; wp.maxDNALength = maxDNA;
        std::vector<uint8_t> st8(nc);
        for (uint64_t c = 0; c < nc; ++c) st8[c] = (uint8_t)cS[c];
        Windows w;
        Aligned a;
        s3_dp *dp = NULL;
        rc = make_windows(ix, S3_WIN_SINGLE, wp, readLengths, numReads, cR, cP, NULL, st8.data(), NULL, NULL, NULL, nc, w);
        if (rc == S3_OK) rc = s3_dp_create(maxRead, maxDNA, (uint32_t)(nc > 32 ? nc : 32), par->scores, ix->device, &dp);
        if (rc == S3_OK) rc = align_windows(dp, ix, queries, readLengths, numReads, wordPerQuery, w, a);
        if (dp) s3_dp_free(dp);
        if (rc == S3_OK)
            for (uint64_t t = 0; t < w.n; ++t) {
                if (a.score[t] < w.cutoff[t]) continue;
                s3_dp_hit h;
                memset(&h, 0, sizeof h);
                h.readID = w.readID[t]; h.strand = w.strand[t]; h.pos = w.start[t] + a.hit[t]; h.score = a.score[t]; h.numSameScore = a.cnt[t];
                h.runOffset = (uint32_t)runs.size();
                pattern_runs(a.pattern.data() + t * a.patLen, a.patLen, runs);
                h.numRuns = (uint16_t)(runs.size() - h.runOffset);
                hits.push_back(h);
            }
    }
    s3_free(cR); s3_free(cP); s3_free(cS);
    if (rc) return rc;
    out->numHits = hits.size(); out->numRuns = runs.size(); out->numUnseeded = unseeded.size();
    out->hits = to_malloc(hits); out->runs = to_malloc(runs); out->unseeded = to_malloc(unseeded);
    if (!out->hits || !out->runs || !out->unseeded) { s3_single_dp_result_free(out); s3_set_error("s3_single_dp_align: out of host memory"); return S3_ENOMEM; }
    return S3_OK;
}

extern "C" int s3_deep_dp_align(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery,
                                const uint32_t *pairReadIDs, uint64_t n, const s3_stage_params *par, s3_deep_dp_result *out)
{
    if (!out) { s3_set_error("s3_deep_dp_align: NULL result"); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    if (!ix || !queries || !readLengths || !par || (n && !pairReadIDs)) { s3_set_error("s3_deep_dp_align: NULL argument"); return S3_EINVAL; }
    out->numPairs = n;
    if (n == 0) return S3_OK;
    uint32_t maxLen = 0;
    for (uint64_t k = 0; k < n; ++k) {
        const uint32_t e = pairReadIDs[k];
        if ((e & 1u) || (uint64_t)e + 1 >= numReads) { s3_set_error("s3_deep_dp_align: %u is not the even read id of a pair", e); return S3_EINVAL; }
        for (int i = 0; i < 2; ++i) if (readLengths[e + i] > maxLen) maxLen = readLengths[e + i];
    }
    int rc = S3_OK;
    std::vector<uint32_t> candID, candL, candR;                       // readIDLeft, estimated starts: all rounds' candidates
    std::vector<uint32_t> input(pairReadIDs, pairReadIDs + n), next, unseeded;
    std::vector<int32_t> seedPos(maxLen + 16);
    for (int round = 0; round < 2 && !input.empty() && rc == S3_OK; ++round) {
        const int stage = round == 0 ? S3_STAGE_DEEP_DP_ROUND1 : S3_STAGE_DEEP_DP_ROUND2;
        // ---- seeds of both mates (PairEndSeedingBatch::packSeeds, DV-DPfunctions.cu:2682-2706)
        SeedSet side[2];
        for (int i = 0; i < 2; ++i) seed_set_reserve(side[i], (maxLen + 15) / 16, input.size() * 32);
        for (size_t k = 0; k < input.size() && rc == S3_OK; ++k) {
            const uint32_t e = input[k];
            s3_dp_stage_params sp;
            if ((rc = s3_dp_stage_parameters(stage, readLengths[e], readLengths[e + 1], par->isDefaultThreshold, par->dpScoreThreshold, par->softClipLeft, par->softClipRight, &sp))) break;
            for (int i = 0; i < 2; ++i) {
                int32_t seedLen = 0, seedNum = 0;
                if ((rc = s3_seed_layout(stage, (int32_t)readLengths[e + i], &seedLen, seedPos.data(), (int32_t)seedPos.size(), &seedNum))) break;
                if (side[i].n + (uint64_t)seedNum > side[i].lengths.size() - 32) { s3_set_error("s3_deep_dp_align: more than 32 seeds per read"); rc = S3_EINVAL; break; }
                for (int32_t j = 0; j < seedNum; ++j)
                    seed_set_add(side[i], queries, wordPerQuery, e + i, e, (uint32_t)seedPos[j], (uint32_t)seedLen, (uint32_t)sp.paramRead[i].maxHitNum);
            }
        }
        if (rc) break;
        out->numSeeds += side[0].n + side[1].n;
        // ---- seeding driver per side; a pair with a too-many seed on either side is flagged (decodePositions, :2951-2954)
        s3_seed_search_result sr[2];
        memset(sr, 0, sizeof sr);
        std::vector<uint8_t> tooMany(numReads, 0);
        std::vector<uint32_t> rid[2], off[2], slen[2], rlen[2];
        std::vector<int32_t> strand[2];
        for (int i = 0; i < 2 && rc == S3_OK; ++i) {
            rc = s3_seed_search(ix, side[i].words.data(), side[i].lengths.data(), side[i].n, side[i].wordPerSeed, side[i].maxHit.data(), &sr[i]);
            if (rc) break;
            rid[i].resize(sr[i].total); off[i].resize(sr[i].total); slen[i].resize(sr[i].total); rlen[i].resize(sr[i].total); strand[i].resize(sr[i].total);
            for (uint64_t s = 0; s < side[i].n; ++s) {
                if (sr[i].status[s] == 4) tooMany[side[i].readIDs[s]] = 1;
                for (uint64_t g = sr[i].offsets[s]; g < sr[i].offsets[s + 1]; ++g) {
                    rid[i][g] = side[i].readIDs[s]; off[i][g] = side[i].offsets[s]; slen[i][g] = side[i].lengths[s];
                    rlen[i][g] = readLengths[side[i].readIDs[s] + i]; strand[i][g] = sr[i].strand[g];
                }
            }
        }
        // ---- candidate position pairs (decodeMergePositions, DV-DPfunctions.cu:2963-2999)
        uint32_t *cID = NULL, *cL = NULL, *cR = NULL;
        uint64_t nc = 0;
        if (rc == S3_OK)
            rc = s3_seed_pair_candidates(ix, sr[0].saL, sr[0].saR, strand[0].data(), rid[0].data(), off[0].data(), slen[0].data(), rlen[0].data(), sr[0].total,
                                         sr[1].saL, sr[1].saR, strand[1].data(), rid[1].data(), off[1].data(), slen[1].data(), rlen[1].data(), sr[1].total,
                                         0xFFFFFFFFu, readLengths, numReads, par->insertLow, par->insertHigh, par->strandLeftLeg, par->strandRightLeg,
                                         &cID, &cL, &cR, &nc);
        s3_seed_search_result_free(&sr[0]); s3_seed_search_result_free(&sr[1]);
        if (rc) break;
        // ---- seeded / too many / unseeded pairs (performSeeding, DV-DPfunctions.cu:3105-3125)
        std::vector<uint8_t> seeded(numReads, 0);
        for (uint64_t c = 0; c < nc; ++c) { seeded[cID[c] & ~1u] = 1; candID.push_back(cID[c]); candL.push_back(cL[c]); candR.push_back(cR[c]); }
        s3_free(cID); s3_free(cL); s3_free(cR);
        next.clear();
        for (size_t k = 0; k < input.size(); ++k) {
            const uint32_t e = input[k];
            if (seeded[e]) continue;
            if (round == 0 && tooMany[e]) next.push_back(e); else unseeded.push_back(e);
        }
        input.swap(next);
    }
    if (rc) return rc;
    const uint64_t nc = candID.size();
    out->numCandidates = nc;
    std::vector<s3_deep_dp_hit> hits;
    std::vector<uint32_t> runs;
    if (nc) {
        // ---- left read, then the right read inside the window the left hit allows (DP2CPUAlgnThread, DV-DPfunctions.cu:3731-3800)
        const uint32_t maxRead = (maxLen / 4 + 1) * 4, maxDNA = maxRead + 2 * margin_of(maxRead) + 8;
        s3_window_params wp;
        memset(&wp, 0, sizeof wp);
        wp.insertLow = par->insertLow; wp.insertHigh = par->insertHigh; wp.strandLeftLeg = par->strandLeftLeg; wp.strandRightLeg = par->strandRightLeg;
        wp.softClipLeft = par->softClipLeft; wp.softClipRight = par->softClipRight;
        wp.cutoffThreshold[0] = wp.cutoffThreshold[1] = par->isDefaultThreshold ? -1 : par->dpScoreThreshold; wp.maxDNALength = maxDNA;
        Windows wl, wr;
        Aligned al, ar;
        s3_dp *dp = NULL;
        rc = make_windows(ix, S3_WIN_PAIR_LEFT, wp, readLengths, numReads, candID.data(), candL.data(), NULL, NULL, NULL, NULL, NULL, nc, wl);
        if (rc == S3_OK) rc = s3_dp_create(maxRead, maxDNA, (uint32_t)(nc > 32 ? nc : 32), par->scores, ix->device, &dp);
        if (rc == S3_OK) rc = align_windows(dp, ix, queries, readLengths, numReads, wordPerQuery, wl, al);
        if (rc == S3_OK) rc = make_windows(ix, S3_WIN_PAIR_RIGHT, wp, readLengths, numReads, candID.data(), candL.data(), candR.data(), NULL, al.score.data(),
                                           wl.start.data(), al.hit.data(), nc, wr);
        if (rc == S3_OK) rc = align_windows(dp, ix, queries, readLengths, numReads, wordPerQuery, wr, ar);
        if (dp) s3_dp_free(dp);
        if (rc == S3_OK)
            for (uint64_t t = 0; t < wr.n; ++t) {
                if (ar.score[t] < wr.cutoff[t]) continue;           // the left read reached its cutoff or the candidate has no right window
                const uint32_t c = wr.cand[t], left = candID[c], readSide = left & 1u;
                // fields _1 belong to the pair's first read, _2 to its mate, whichever is the left one
                s3_deep_dp_hit h;
                memset(&h, 0, sizeof h);
                h.readID = left - readSide;
                const uint32_t posLeft = wl.start[c] + al.hit[c], posRight = wr.start[t] + ar.hit[t];
                uint32_t roL = (uint32_t)runs.size();
                pattern_runs(al.pattern.data() + (size_t)c * al.patLen, al.patLen, runs);
                uint32_t nL = (uint32_t)runs.size() - roL, roR = (uint32_t)runs.size();
                pattern_runs(ar.pattern.data() + (size_t)t * ar.patLen, ar.patLen, runs);
                uint32_t nR = (uint32_t)runs.size() - roR;
                if (readSide == 0) {
                    h.pos1 = posLeft; h.pos2 = posRight; h.score1 = al.score[c]; h.score2 = ar.score[t]; h.numSame1 = al.cnt[c]; h.numSame2 = ar.cnt[t];
                    h.strand1 = (uint8_t)par->strandLeftLeg; h.strand2 = (uint8_t)par->strandRightLeg;
                    h.runOffset1 = roL; h.numRuns1 = (uint16_t)nL; h.runOffset2 = roR; h.numRuns2 = (uint16_t)nR;
                } else {
                    h.pos1 = posRight; h.pos2 = posLeft; h.score1 = ar.score[t]; h.score2 = al.score[c]; h.numSame1 = ar.cnt[t]; h.numSame2 = al.cnt[c];
                    h.strand1 = (uint8_t)par->strandRightLeg; h.strand2 = (uint8_t)par->strandLeftLeg;
                    h.runOffset1 = roR; h.numRuns1 = (uint16_t)nR; h.runOffset2 = roL; h.numRuns2 = (uint16_t)nL;
                }
                hits.push_back(h);
            }
    }
    if (rc) return rc;
    out->numHits = hits.size(); out->numRuns = runs.size(); out->numUnseeded = unseeded.size();
    out->hits = to_malloc(hits); out->runs = to_malloc(runs); out->unseeded = to_malloc(unseeded);
    if (!out->hits || !out->runs || !out->unseeded) { s3_deep_dp_result_free(out); s3_set_error("s3_deep_dp_align: out of host memory"); return S3_ENOMEM; }
    return S3_OK;
}
