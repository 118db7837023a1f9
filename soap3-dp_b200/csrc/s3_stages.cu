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
#include "s3_windows.cuh"
#include "../../include/soap3dp_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>

// ---- device side of the stages: seeds cut from the query buffer, windows of the candidates ---------------------------------
#define S3_STAGE_MAX_SEEDS 32
struct S3SeedPlan { int32_t seedLen, seedNum, maxHit, pad; int32_t pos[S3_STAGE_MAX_SEEDS]; };

// one thread per (entry, seed slot): entry k cuts the seeds of read srcRead[k] by plan planIdx[k]; seed ids start at seedStart[k]
__global__ void s3_stage_seed_kernel(uint32_t n, const uint32_t *__restrict__ queries, uint32_t wpq, const uint32_t *__restrict__ entries,
                                     const S3SeedPlan *__restrict__ plans, const uint32_t *__restrict__ lenByRead, uint32_t wordPerSeed,
                                     uint32_t *__restrict__ seeds, uint32_t *__restrict__ seedLen, uint32_t *__restrict__ seedKey,
                                     uint32_t *__restrict__ seedOff, uint32_t *__restrict__ seedMaxHit, uint32_t *__restrict__ seedReadLen)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x, k = idx / S3_STAGE_MAX_SEEDS, j = idx % S3_STAGE_MAX_SEEDS;
    if (k >= n) return;
    const uint32_t r = entries[k], key = entries[n + k], start = entries[3 * (size_t)n + k];
    const S3SeedPlan &p = plans[entries[2 * (size_t)n + k]];
    if ((int32_t)j >= p.seedNum) return;
    const uint32_t id = start + j, off = (uint32_t)p.pos[j], len = (uint32_t)p.seedLen;
    const uint32_t *src = queries + (size_t)(r / 32) * 32 * wpq + r % 32;
    uint32_t *dst = seeds + (size_t)(id / 32) * 32 * wordPerSeed + id % 32;
    for (uint32_t w = 0, done = 0; w < wordPerSeed; ++w, done += 16) {
        uint32_t v = 0;
        if (done < len) {
            const uint32_t b = off + done, sw = b >> 4, sh = (b & 15u) << 1;
            v = sw < wpq ? src[(size_t)sw * 32] >> sh : 0u;
            if (sh && sw + 1 < wpq) v |= src[(size_t)(sw + 1) * 32] << (32u - sh);
            const uint32_t rem = len - done;
            if (rem < 16) v &= (1u << (2 * rem)) - 1u;
        }
        dst[(size_t)w * 32] = v;
    }
    seedLen[id] = len; seedKey[id] = key; seedOff[id] = off; seedMaxHit[id] = (uint32_t)p.maxHit; seedReadLen[id] = lenByRead[r];
}

struct S3StageWin { uint32_t *readID, *start, *dnaLen, *clipLt, *clipRt, *ancL, *ancR, *cand; int32_t *cutoff; uint8_t *strand; };

__device__ __forceinline__ void s3_stage_win_store(const S3StageWin &o, uint32_t k, uint32_t cand, const S3Window &x)
{
    o.cand[k] = cand; o.readID[k] = x.readID; o.start[k] = x.start; o.dnaLen[k] = x.dnaLen; o.clipLt[k] = x.clipLt; o.clipRt[k] = x.clipRt;
    o.ancL[k] = x.ancL; o.ancR[k] = x.ancR; o.strand[k] = x.strand; o.cutoff[k] = x.cutoff;
}

// SingleEndAlgnBatch::pack for every candidate (read id | estimated start | strand)
__global__ void s3_stage_win_single_kernel(S3WinParams w, uint32_t m, const uint32_t *__restrict__ cand, const uint32_t *__restrict__ lenByRead, S3StageWin o)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    S3Window x;
    const uint32_t id = cand[c];
    s3_win_single(w, id, cand[m + c], (int)cand[2 * (size_t)m + c], lenByRead[id], x);
    s3_stage_win_store(o, c, c, x);
}

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
                          w.start.data(), w.len.data(), w.cutoff.data(), w.clipLt.data(), w.clipRt.data(), w.ancL.data(), w.ancR.data(), NULL, NULL, &a);
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

#define S3_TRYS(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { s3_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); rc = S3_ECUDA; goto done; } } while (0)

namespace {

// seed plans of a stage by read length, made once per distinct length; entries (read to cut, key id, plan, first seed id)
struct StagePlans {
    std::vector<S3SeedPlan> table;
    std::vector<int32_t> ofLen;                 // plan index by read length, -1: not made yet
    uint32_t maxSeedLen = 1;
    int get(int stage, uint32_t len, uint32_t len2, int side, const s3_stage_params *par, uint32_t *idx)
    {
        if (len >= ofLen.size()) ofLen.resize(len + 1, -1);
        if (ofLen[len] < 0) {
            S3SeedPlan p;
            memset(&p, 0, sizeof p);
            std::vector<int32_t> pos(len + 64);
            int rc = s3_seed_layout(stage, (int32_t)len, &p.seedLen, pos.data(), (int32_t)pos.size(), &p.seedNum);
            if (rc) return rc;
            if (p.seedNum > S3_STAGE_MAX_SEEDS) { s3_set_error("seeded DP stage: more than %d seeds per read", S3_STAGE_MAX_SEEDS); return S3_EINVAL; }
            for (int32_t j = 0; j < p.seedNum; ++j) p.pos[j] = pos[j];
            s3_dp_stage_params sp;
            if ((rc = s3_dp_stage_parameters(stage, side == 0 ? len : len2, side == 0 ? len2 : len, par->isDefaultThreshold, par->dpScoreThreshold,
                                             par->softClipLeft, par->softClipRight, &sp))) return rc;
            p.maxHit = sp.paramRead[side].maxHitNum;
            if ((uint32_t)p.seedLen > maxSeedLen) maxSeedLen = (uint32_t)p.seedLen;
            ofLen[len] = (int32_t)table.size();
            table.push_back(p);
        }
        *idx = (uint32_t)ofLen[len];
        return S3_OK;
    }
};

// one side's seeding on the device: entries -> seeds -> seeding driver.  Device arrays of the result are freed by the caller.
struct SeedSide {
    uint32_t *d_entries = NULL, *d_seedBuf = NULL;
    S3SeedPlan *d_plans = NULL;
    uint64_t numSeeds = 0;
    S3SeedRangesDev ranges;
    SeedSide() { memset(&ranges, 0, sizeof ranges); }
};

void seed_side_free(s3_index *ix, SeedSide &x)
{
    void *p[] = {x.d_entries, x.d_seedBuf, x.d_plans, x.ranges.d_buf, x.ranges.d_status};
    for (void *q : p) if (q) cudaFreeAsync(q, ix->stream);
    x = SeedSide();
}

// entries: srcRead | key | plan | seedStart (n each, host); the stage's query buffer and d_lenByRead are on the device already
int seed_side_run(s3_index *ix, const std::vector<uint32_t> &entries, size_t n, uint64_t totalSeeds, const StagePlans &plans, uint32_t wpq,
                  const uint32_t *d_lenByRead, SeedSide &x)
{
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    x.numSeeds = totalSeeds;
    if (n == 0 || totalSeeds == 0) return S3_OK;
    if (totalSeeds >= 0x7FFFFFF0ull) { s3_set_error("seeded DP stage: too many seeds in one call"); return S3_EINVAL; }
    const uint32_t S = (uint32_t)totalSeeds, wps = (plans.maxSeedLen + 15) / 16;
    const size_t up = ((size_t)S + 31) / 32 * 32;
    S3_TRYS(cudaMallocAsync((void **)&x.d_entries, 4 * n * 4, st));
    S3_TRYS(cudaMallocAsync((void **)&x.d_plans, plans.table.size() * sizeof(S3SeedPlan), st));
    S3_TRYS(cudaMallocAsync((void **)&x.d_seedBuf, (up * wps + up + 4 * (size_t)S) * 4, st));
    S3_TRYS(cudaMemcpyAsync(x.d_entries, entries.data(), 4 * n * 4, cudaMemcpyHostToDevice, st));
    S3_TRYS(cudaMemcpyAsync(x.d_plans, plans.table.data(), plans.table.size() * sizeof(S3SeedPlan), cudaMemcpyHostToDevice, st));
    {
        uint32_t *d_seeds = x.d_seedBuf, *d_seedLen = d_seeds + up * wps, *d_key = d_seedLen + up, *d_off = d_key + S, *d_maxHit = d_off + S, *d_rl = d_maxHit + S;
        S3_TRYS(cudaMemsetAsync(d_seeds, 0, (up * wps + up) * 4, st));
        const unsigned long long threads = (unsigned long long)n * S3_STAGE_MAX_SEEDS;
        s3_stage_seed_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((uint32_t)n, s3_stage_queries(ix), wpq, x.d_entries, x.d_plans, d_lenByRead, wps,
                                                                                  d_seeds, d_seedLen, d_key, d_off, d_maxHit, d_rl);
        S3_LAUNCHED(1);
        S3_TRYS(cudaGetLastError());
        rc = s3_seed_search_device(ix, d_seeds, d_seedLen, S, wps, d_maxHit, d_key, d_off, d_rl, &x.ranges);
    }
done:
    return rc;
}

}  // namespace

extern "C" int s3_single_dp_align(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery,
                                  const uint32_t *readIDs, uint64_t n, const s3_stage_params *par, s3_single_dp_result *out)
{
    if (!out) { s3_set_error("s3_single_dp_align: NULL result"); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    if (!ix || !queries || !readLengths || !par || (n && !readIDs)) { s3_set_error("s3_single_dp_align: NULL argument"); return S3_EINVAL; }
    if (!ix->loc.sa || !ix->loc.text) { s3_set_error("s3_single_dp_align: the index was uploaded without its suffix array and packed text"); return S3_EINVAL; }
    out->numReads = n;
    if (n == 0) return S3_OK;
    if (numReads >= 0xFFFFFFF0ull) { s3_set_error("s3_single_dp_align: too many reads"); return S3_EINVAL; }
    uint32_t maxLen = 0;
    for (uint64_t k = 0; k < n; ++k) {
        if (readIDs[k] >= numReads) { s3_set_error("s3_single_dp_align: read id %u out of range", readIDs[k]); return S3_EINVAL; }
        if (readLengths[readIDs[k]] > 16u * wordPerQuery) { s3_set_error("s3_single_dp_align: read %u longer than its query words", readIDs[k]); return S3_EINVAL; }
        if (readLengths[readIDs[k]] > maxLen) maxLen = readLengths[readIDs[k]];
    }
    int rc = S3_OK;
    StageClock clk("s3_single_dp_align");
    if (cudaSetDevice(ix->device) != cudaSuccess) { s3_set_error("s3_single_dp_align: cudaSetDevice failed"); return S3_ECUDA; }
    cudaStream_t st = ix->stream;
    // ---- seeds (SingleEndSeedingBatch::packSeeds, DV-DPfunctions.cu:1082-1100): the layout depends on the read length only
    StagePlans plans;
    std::vector<uint32_t> entries(4 * n);
    uint64_t totalSeeds = 0;
    for (uint64_t k = 0; k < n; ++k) {
        const uint32_t r = readIDs[k];
        uint32_t idx;
        if ((rc = plans.get(S3_STAGE_SINGLE_DP, readLengths[r], 0, 0, par, &idx))) return rc;
        entries[k] = r; entries[n + k] = r; entries[2 * n + k] = idx; entries[3 * n + k] = (uint32_t)totalSeeds;
        totalSeeds += (uint64_t)plans.table[idx].seedNum;
    }
    out->numSeeds = totalSeeds;
    uint32_t *d_len = NULL, *d_cand = NULL, *d_win = NULL;
    uint32_t m = 0;
    SeedSide side;
    std::vector<s3_dp_hit> hits;
    std::vector<uint32_t> runs, unseeded;
    std::vector<uint8_t> seeded(numReads, 0);
    S3_TRYS(cudaMallocAsync((void **)&d_len, numReads * 4 + 16, st));
    S3_TRYS(cudaMemcpyAsync(d_len, readLengths, numReads * 4, cudaMemcpyHostToDevice, st));
    if ((rc = s3_stage_upload_queries(ix, queries, numReads, wordPerQuery))) goto done;
    clk.lap("plans + uploads");
    // ---- seeding driver, candidate positions (decodePositions + singleMerge, DV-DPfunctions.cu:1101-1219)
    if ((rc = seed_side_run(ix, entries, n, totalSeeds, plans, wordPerQuery, d_len, side))) goto done;
    clk.lap("seeds + seeding driver");
    if (side.ranges.numRanges) {
        const uint64_t R = side.ranges.numRanges;
        const uint32_t *b = side.ranges.d_buf;
        if ((rc = s3_seed_candidates_device(ix, b, b + R, (const int32_t *)(b + 2 * R), b + 3 * R, b + 4 * R, b + 5 * R, b + 6 * R, R, 0xFFFFFFFFu, &d_cand, &m))) goto done;
    }
    out->numCandidates = m;
    clk.lap("candidates");
    if (m) {
        // ---- windows (SingleEndAlgnBatch::pack), DP, results (SingleDP_Space::algnmtCPUThread, DV-DPfunctions.cu:1699-1733)
        const uint32_t maxRead = (maxLen / 4 + 1) * 4, maxDNA = maxRead + 2 * margin_of(maxRead) + 8;         // DV-DPfunctions.cu:1580-1581
        S3WinParams w;
        memset(&w, 0, sizeof w);
        w.leftLeg = 1; w.rightLeg = 2; w.softClipLeft = par->softClipLeft; w.softClipRight = par->softClipRight;
        w.cutoff[0] = w.cutoff[1] = par->isDefaultThreshold ? -1 : par->dpScoreThreshold; w.maxDNALength = maxDNA; w.textLength = ix->textLength;
        S3_TRYS(cudaMallocAsync((void **)&d_win, (size_t)m * 40 + 64, st));
        S3StageWin o;
        o.readID = d_win; o.start = d_win + m; o.dnaLen = d_win + 2 * (size_t)m; o.clipLt = d_win + 3 * (size_t)m; o.clipRt = d_win + 4 * (size_t)m;
        o.ancL = d_win + 5 * (size_t)m; o.ancR = d_win + 6 * (size_t)m; o.cand = d_win + 7 * (size_t)m; o.cutoff = (int32_t *)(d_win + 8 * (size_t)m);
        o.strand = (uint8_t *)(d_win + 9 * (size_t)m);
        s3_stage_win_single_kernel<<<(m + 255) / 256, 256, 0, st>>>(w, m, d_cand, d_len, o);
        S3_LAUNCHED(1);
        S3_TRYS(cudaGetLastError());
        S3StageAligned a;
        if ((rc = s3_stage_align(ix, queries, readLengths, numReads, wordPerQuery, 0, maxRead, maxDNA, par->scores, 0, m, o.readID, o.strand, o.start, o.dnaLen,
                                 o.cutoff, o.clipLt, o.clipRt, o.ancL, o.ancR, d_len, NULL, &a))) goto done;
        clk.lap("windows + DP + CIGAR runs");
        hits.reserve(m);
        runs.reserve(a.numRuns);
        for (uint32_t t = 0; t < m; ++t) {
            seeded[a.readID[t]] = 1;
            if (a.score[t] < a.cutoff[t]) continue;
            s3_dp_hit h;
            memset(&h, 0, sizeof h);
            h.readID = a.readID[t]; h.strand = a.strand[t]; h.pos = a.start[t] + a.hit[t]; h.score = a.score[t]; h.numSameScore = a.cnt[t];
            h.runOffset = (uint32_t)runs.size();
            runs.insert(runs.end(), a.runs + a.runOff[t], a.runs + a.runOff[t + 1]);
            h.numRuns = (uint16_t)(runs.size() - h.runOffset);
            hits.push_back(h);
        }
    }
    // reads without a candidate (alignFlags XOR inputFlags, DV-DPfunctions.cu:1300-1310)
    for (uint64_t k = 0; k < n; ++k) if (!seeded[readIDs[k]]) unseeded.push_back(readIDs[k]);
    clk.lap("records");
done:
    seed_side_free(ix, side);
    if (d_len) cudaFreeAsync(d_len, st);
    if (d_cand) cudaFreeAsync(d_cand, st);
    if (d_win) cudaFreeAsync(d_win, st);
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
