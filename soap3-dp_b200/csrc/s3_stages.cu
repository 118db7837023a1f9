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

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>

namespace {

// S3_STAGE_TIMING=1: wall-clock milliseconds of every step of a stage call on stderr (tuning aid)
struct StageClock {
    bool on;
    std::chrono::steady_clock::time_point t;
    const char *what;
    explicit StageClock(const char *w) : on(getenv("S3_STAGE_TIMING") != NULL), t(std::chrono::steady_clock::now()), what(w) {}
    void lap(const char *step)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[%s] %-28s %8.2f ms\n", what, step, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

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

// `len` bases of read `readID` from base `off` on, 16 per word, as the next seed of the set (word-wise: two source words per word)
void seed_set_add(SeedSet &s, const uint32_t *queries, uint32_t wpq, uint32_t readID, uint32_t keyID, uint32_t off, uint32_t len, uint32_t maxHit)
{
    const uint64_t id = s.n++;
    uint32_t *dst = s.words.data() + (id / 32) * 32 * s.wordPerSeed + id % 32;
    const uint32_t *src = queries + (size_t)(readID / 32) * 32 * wpq + readID % 32;
    for (uint32_t w = 0, done = 0; done < len; ++w, done += 16) {
        const uint32_t k = off + done, sw = k >> 4, sh = (k & 15u) << 1;
        uint32_t v = sw < wpq ? src[(size_t)sw * 32] >> sh : 0u;
        if (sh && sw + 1 < wpq) v |= src[(size_t)(sw + 1) * 32] << (32u - sh);
        const uint32_t rem = len - done;
        if (rem < 16) v &= (1u << (2 * rem)) - 1u;
        dst[(size_t)w * 32] = v;
    }
    s.lengths[id] = len;
    s.readIDs.push_back(keyID); s.offsets.push_back(off); s.maxHit.push_back(maxHit);
}

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

int align_windows(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wpq, int uploadQueries, uint32_t maxRead,
                  uint32_t maxDNA, s3_dp_scores scores, int slot, Windows &w, S3StageAligned &a)
{
    return s3_stage_align(ix, queries, readLengths, numReads, wpq, uploadQueries, maxRead, maxDNA, scores, slot, w.n, w.readID.data(), w.strand.data(),
                          w.start.data(), w.len.data(), w.cutoff.data(), w.clipLt.data(), w.clipRt.data(), w.ancL.data(), w.ancR.data(), &a);
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
    StageClock clk("s3_single_dp_align");
    // ---- seeds (SingleEndSeedingBatch::packSeeds, DV-DPfunctions.cu:1082-1100)
    SeedSet seeds;
    // seed layout and hit limit depend on the read length only: one plan per length that occurs
    struct Plan { int32_t seedLen, seedNum, maxHit; std::vector<int32_t> pos; bool made; };
    std::vector<Plan> plans(maxLen + 1);
    for (auto &p : plans) p.made = false;
    uint64_t totalSeeds = 0;
    uint32_t maxSeedLen = 1;
    for (int pass = 0; pass < 2; ++pass)
    for (uint64_t k = 0; k < n; ++k) {
        const uint32_t r = readIDs[k], len = readLengths[r];
        Plan &p = plans[len];
        if (pass == 0 && p.made) { totalSeeds += (uint64_t)p.seedNum; continue; }
        if (pass == 1 && k == 0) seed_set_reserve(seeds, (maxSeedLen + 15) / 16, (size_t)totalSeeds);
        if (!p.made) {
            p.pos.resize(maxLen + 16);
            if ((rc = s3_seed_layout(S3_STAGE_SINGLE_DP, (int32_t)len, &p.seedLen, p.pos.data(), (int32_t)p.pos.size(), &p.seedNum))) return rc;
            s3_dp_stage_params sp;
            if ((rc = s3_dp_stage_parameters(S3_STAGE_SINGLE_DP, len, 0, par->isDefaultThreshold, par->dpScoreThreshold, par->softClipLeft, par->softClipRight, &sp))) return rc;
            p.maxHit = sp.paramRead[0].maxHitNum;
            p.made = true;
            if ((uint32_t)p.seedLen > maxSeedLen) maxSeedLen = (uint32_t)p.seedLen;
        }
        if (pass == 0) { totalSeeds += (uint64_t)p.seedNum; continue; }
        for (int32_t j = 0; j < p.seedNum; ++j) seed_set_add(seeds, queries, wordPerQuery, r, r, (uint32_t)p.pos[j], (uint32_t)p.seedLen, (uint32_t)p.maxHit);
    }
    out->numSeeds = seeds.n;
    clk.lap("seed packing");
    // ---- seeding driver, candidate positions (decodePositions + singleMerge, DV-DPfunctions.cu:1101-1219)
    s3_seed_search_result sr;
    if ((rc = s3_seed_search(ix, seeds.words.data(), seeds.lengths.data(), seeds.n, seeds.wordPerSeed, seeds.maxHit.data(), &sr))) return rc;
    clk.lap("s3_seed_search");
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
    clk.lap("s3_seed_candidates");
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
        S3StageAligned a;
        rc = make_windows(ix, S3_WIN_SINGLE, wp, readLengths, numReads, cR, cP, NULL, st8.data(), NULL, NULL, NULL, nc, w);
        clk.lap("s3_dp_make_windows");
        if (rc == S3_OK) rc = align_windows(ix, queries, readLengths, numReads, wordPerQuery, 1, maxRead, maxDNA, par->scores, 0, w, a);
        clk.lap("upload + DP + CIGAR runs");
        if (rc == S3_OK && w.n) {
            hits.reserve(w.n);
            runs.reserve(a.numRuns);
            for (uint64_t t = 0; t < w.n; ++t) {
                if (a.score[t] < w.cutoff[t]) continue;
                s3_dp_hit h;
                memset(&h, 0, sizeof h);
                h.readID = w.readID[t]; h.strand = w.strand[t]; h.pos = w.start[t] + a.hit[t]; h.score = a.score[t]; h.numSameScore = a.cnt[t];
                h.runOffset = (uint32_t)runs.size();
                runs.insert(runs.end(), a.runs + a.runOff[t], a.runs + a.runOff[t + 1]);
                h.numRuns = (uint16_t)(runs.size() - h.runOffset);
                hits.push_back(h);
            }
        }
    }
    s3_free(cR); s3_free(cP); s3_free(cS);
    clk.lap("records + CIGAR runs");
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
    StageClock clk("s3_deep_dp_align");
    std::vector<uint32_t> candID, candL, candR;                       // readIDLeft, estimated starts: all rounds' candidates
    std::vector<uint32_t> input(pairReadIDs, pairReadIDs + n), next, unseeded;
    for (int round = 0; round < 2 && !input.empty() && rc == S3_OK; ++round) {
        const int stage = round == 0 ? S3_STAGE_DEEP_DP_ROUND1 : S3_STAGE_DEEP_DP_ROUND2;
        // ---- seeds of both mates (PairEndSeedingBatch::packSeeds, DV-DPfunctions.cu:2682-2706)
        SeedSet side[2];
        // seed layout per read length, hit limits per pair of lengths: made once per distinct value
        struct Layout { int32_t seedLen, seedNum; std::vector<int32_t> pos; bool made; };
        std::vector<Layout> layouts(maxLen + 1);
        for (auto &l : layouts) l.made = false;
        uint64_t sideSeeds[2] = {0, 0};
        uint32_t maxSeedLen = 1;
        for (size_t k = 0; k < input.size() && rc == S3_OK; ++k)
            for (int i = 0; i < 2; ++i) {
                Layout &l = layouts[readLengths[input[k] + i]];
                if (!l.made) {
                    l.pos.resize(maxLen + 16);
                    if ((rc = s3_seed_layout(stage, (int32_t)readLengths[input[k] + i], &l.seedLen, l.pos.data(), (int32_t)l.pos.size(), &l.seedNum))) break;
                    l.made = true;
                    if ((uint32_t)l.seedLen > maxSeedLen) maxSeedLen = (uint32_t)l.seedLen;
                }
                sideSeeds[i] += (uint64_t)l.seedNum;
            }
        if (rc) break;
        for (int i = 0; i < 2; ++i) seed_set_reserve(side[i], (maxSeedLen + 15) / 16, (size_t)sideSeeds[i]);
        uint32_t spLen[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
        s3_dp_stage_params sp;
        for (size_t k = 0; k < input.size() && rc == S3_OK; ++k) {
            const uint32_t e = input[k];
            if (readLengths[e] != spLen[0] || readLengths[e + 1] != spLen[1]) {
                if ((rc = s3_dp_stage_parameters(stage, readLengths[e], readLengths[e + 1], par->isDefaultThreshold, par->dpScoreThreshold, par->softClipLeft, par->softClipRight, &sp))) break;
                spLen[0] = readLengths[e]; spLen[1] = readLengths[e + 1];
            }
            for (int i = 0; i < 2; ++i) {
                Layout &l = layouts[readLengths[e + i]];
                for (int32_t j = 0; j < l.seedNum; ++j)
                    seed_set_add(side[i], queries, wordPerQuery, e + i, e, (uint32_t)l.pos[j], (uint32_t)l.seedLen, (uint32_t)sp.paramRead[i].maxHitNum);
            }
        }
        if (rc) break;
        clk.lap("seed packing");
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
        clk.lap("s3_seed_search x 2");
        // ---- candidate position pairs (decodeMergePositions, DV-DPfunctions.cu:2963-2999)
        uint32_t *cID = NULL, *cL = NULL, *cR = NULL;
        uint64_t nc = 0;
        if (rc == S3_OK)
            rc = s3_seed_pair_candidates(ix, sr[0].saL, sr[0].saR, strand[0].data(), rid[0].data(), off[0].data(), slen[0].data(), rlen[0].data(), sr[0].total,
                                         sr[1].saL, sr[1].saR, strand[1].data(), rid[1].data(), off[1].data(), slen[1].data(), rlen[1].data(), sr[1].total,
                                         0xFFFFFFFFu, readLengths, numReads, par->insertLow, par->insertHigh, par->strandLeftLeg, par->strandRightLeg,
                                         &cID, &cL, &cR, &nc);
        s3_seed_search_result_free(&sr[0]); s3_seed_search_result_free(&sr[1]);
        clk.lap("s3_seed_pair_candidates");
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
        S3StageAligned al, ar;
        clk.lap("round bookkeeping");
        rc = make_windows(ix, S3_WIN_PAIR_LEFT, wp, readLengths, numReads, candID.data(), candL.data(), NULL, NULL, NULL, NULL, NULL, nc, wl);
        clk.lap("windows left");
        if (rc == S3_OK) rc = align_windows(ix, queries, readLengths, numReads, wordPerQuery, 1, maxRead, maxDNA, par->scores, 0, wl, al);
        clk.lap("align left");
        if (rc == S3_OK) rc = make_windows(ix, S3_WIN_PAIR_RIGHT, wp, readLengths, numReads, candID.data(), candL.data(), candR.data(), NULL, al.score,
                                           wl.start.data(), al.hit, nc, wr);
        if (rc == S3_OK) rc = align_windows(ix, queries, readLengths, numReads, wordPerQuery, 0, maxRead, maxDNA, par->scores, 1, wr, ar);
        clk.lap("windows right + align right");
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
                runs.insert(runs.end(), al.runs + al.runOff[c], al.runs + al.runOff[c + 1]);
                uint32_t nL = (uint32_t)runs.size() - roL, roR = (uint32_t)runs.size();
                runs.insert(runs.end(), ar.runs + ar.runOff[t], ar.runs + ar.runOff[t + 1]);
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
    clk.lap("records + CIGAR runs");
    if (rc) return rc;
    out->numHits = hits.size(); out->numRuns = runs.size(); out->numUnseeded = unseeded.size();
    out->hits = to_malloc(hits); out->runs = to_malloc(runs); out->unseeded = to_malloc(unseeded);
    if (!out->hits || !out->runs || !out->unseeded) { s3_deep_dp_result_free(out); s3_set_error("s3_deep_dp_align: out of host memory"); return S3_ENOMEM; }
    return S3_OK;
}
