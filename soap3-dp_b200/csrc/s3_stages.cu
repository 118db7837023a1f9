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
// The stage logic (which reads go on, the rounds, the records) runs on the host like the reference's wrappers; everything
// between stays on the device: the seeds are cut from the query buffer by a kernel, the seeding driver, the seed merges, the
// windows, the DP and the CIGAR runs work on device arrays (s3_seed_search_device, s3_seed_candidates_device,
// s3_seed_pair_candidates_any, s3_stage_align), and what comes back per call is a few small count reads, the candidates'
// read ids and the alignments' scores, positions and runs.  Stream-ordered allocations (cudaMallocAsync) hold the intermediates.
#include "s3_common.cuh"
#include "s3_windows.cuh"
#include <cub/cub.cuh>
#include "../../include/soap3dp_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <map>
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

// packLeft for every candidate (readIDLeft | estimated left start | estimated right start)
__global__ void s3_stage_win_left_kernel(S3WinParams w, uint32_t m, const uint32_t *__restrict__ cand, const uint32_t *__restrict__ lenByRead, S3StageWin o)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    S3Window x;
    const uint32_t id = cand[c];
    s3_win_pair_left(w, id, cand[m + c], lenByRead[id], x);
    s3_stage_win_store(o, c, c, x);
}

// packRight: only for candidates whose left read reached its cutoff; window cut by the left hit.  COUNT / FILL
template <bool FILL>
__global__ void s3_stage_win_right_kernel(S3WinParams w, uint32_t m, const uint32_t *__restrict__ cand, const uint32_t *__restrict__ lenByRead,
                                          const int32_t *__restrict__ leftScore, const uint32_t *__restrict__ leftStart, const uint32_t *__restrict__ leftHit,
                                          uint32_t *__restrict__ count, const uint32_t *__restrict__ off, S3StageWin o)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const uint32_t id = cand[c];
    const bool go = leftScore[c] >= s3_win_cutoff(w, id, lenByRead[id]);
    if (!FILL) { count[c] = go ? 1u : 0u; return; }
    if (!go) return;
    S3Window x;
    s3_win_pair_right(w, id, cand[2 * (size_t)m + c], leftStart[c] + leftHit[c], lenByRead[id ^ 1u], x);
    s3_stage_win_store(o, off[c], c, x);
}

namespace {

// S3_STAGE_TIMING=1: wall-clock milliseconds of every step of a stage call on stderr (tuning aid)
struct StageClock {
    bool on;
    std::chrono::steady_clock::time_point t, t0;
    const char *what;
    explicit StageClock(const char *w) : on(getenv("S3_STAGE_TIMING") != NULL), t(std::chrono::steady_clock::now()), t0(t), what(w) {}
    ~StageClock()
    {
        if (on) fprintf(stderr, "[%s] %-28s %8.2f ms\n", what, "TOTAL since entry checks", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    void lap(const char *step)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[%s] %-28s %8.2f ms\n", what, step, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

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
    std::map<uint64_t, int> made;               // (read length, mate length, side) -> plan index
    uint32_t maxSeedLen = 1;
    int get(int stage, uint32_t len, uint32_t len2, int side, const s3_stage_params *par, uint32_t *idx)
    {
        const uint64_t key = (uint64_t)len | ((uint64_t)len2 << 24) | ((uint64_t)side << 48);
        auto it = made.find(key);
        if (it == made.end()) {
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
            it = made.emplace(key, (int)table.size()).first;
            table.push_back(p);
        }
        *idx = (uint32_t)it->second;
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

namespace {

// the pairs with route `want` of a chain's route array, as even read ids, with the two reads' lengths: id | length | mate's length
__global__ void s3_stage_pick_flag_kernel(uint32_t P, const uint8_t *__restrict__ route, uint8_t want, uint8_t *__restrict__ flags)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) flags[p] = route[p] == want;
}
__global__ void s3_stage_pick_kernel(uint32_t n, const uint32_t *__restrict__ pairs, const uint32_t *__restrict__ lenByRead, uint32_t *__restrict__ out)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t e = 2u * pairs[k];
    out[k] = e; out[n + k] = lenByRead[e]; out[2 * (size_t)n + k] = lenByRead[e + 1];
}

// DPForUnalignPairs2 for the pairs pairIDs[0..n) (even read ids) whose two lengths are pairLens[2k], pairLens[2k + 1].  The query
// buffer is on the device already (s3_stage_queries(ix)) and so are the read lengths by read id.
int deep_dp_core(s3_index *ix, const uint32_t *d_len, uint64_t numReads, uint32_t wordPerQuery, const uint32_t *pairIDs, const uint32_t *pairLens, uint64_t n,
                 const s3_stage_params *par, s3_deep_dp_result *out, StageClock &clk)
{
    int rc = S3_OK;
    cudaStream_t st = ix->stream;
    uint32_t maxLen = 0;
    for (uint64_t k = 0; k < 2 * n; ++k) if (pairLens[k] > maxLen) maxLen = pairLens[k];
    std::vector<uint32_t> candID;                                      // readIDLeft of all rounds' candidates
    std::vector<uint32_t> input(n), next, unseeded;                     // places in pairIDs
    std::vector<uint8_t> flags(numReads, 0);                            // by even read id: 1 a seed with too many hits, 2 a candidate
    std::vector<s3_deep_dp_hit> hits;
    std::vector<uint32_t> runs;
    std::vector<uint32_t> entries;
    std::vector<uint8_t> status;
    uint32_t *d_c = NULL, *d_wl = NULL, *d_wr = NULL, *d_cnt = NULL;
    void *d_tmp = NULL;
    SeedSide side;
    uint32_t *d_round[2] = {NULL, NULL};                                // a round's candidates: ids | left | right (base of s3_seed_pair_candidates_any's allocation)
    uint64_t ncRound[2] = {0, 0};
    for (uint64_t k = 0; k < n; ++k) input[k] = (uint32_t)k;
    for (int round = 0; round < 2 && !input.empty(); ++round) {
        const int stage = round == 0 ? S3_STAGE_DEEP_DP_ROUND1 : S3_STAGE_DEEP_DP_ROUND2;
        // ---- seeds of both mates (PairEndSeedingBatch::packSeeds, DV-DPfunctions.cu:2682-2706): one seeding call for the two sides,
        // the first mates' seeds before the second mates' -- the driver returns ranges in seed order, so each side is a stretch of them
        StagePlans plans;
        const size_t ni = input.size(), ne = 2 * ni;
        entries.resize(4 * ne);
        uint64_t total[2] = {0, 0};
        for (int i = 0; i < 2; ++i) {
            for (size_t k = 0; k < ni; ++k) {
                const uint32_t e = pairIDs[input[k]], len = pairLens[2 * (size_t)input[k] + i], len2 = pairLens[2 * (size_t)input[k] + 1 - i];
                uint32_t idx;
                if ((rc = plans.get(stage, len, len2, i, par, &idx))) goto done;
                const size_t at = (size_t)i * ni + k;
                entries[at] = e + i; entries[ne + at] = e; entries[2 * ne + at] = idx; entries[3 * ne + at] = (uint32_t)(total[0] + total[1]);
                total[i] += (uint64_t)plans.table[idx].seedNum;
            }
        }
        side.ranges.splitSeed = (uint32_t)total[0];
        if ((rc = seed_side_run(ix, entries, ne, total[0] + total[1], plans, wordPerQuery, d_len, side))) goto done;
        out->numSeeds += total[0] + total[1];
        clk.lap("seeds + seeding driver");
        // a pair with a too-many seed on either side is flagged (decodePositions, :2951-2954)
        for (size_t k = 0; k < ni; ++k) flags[pairIDs[input[k]]] = 0;
        if (total[0] + total[1]) {
            status.resize(total[0] + total[1]);
            S3_TRYS(cudaMemcpyAsync(status.data(), side.ranges.d_status, status.size(), cudaMemcpyDeviceToHost, st));
            S3_TRYS(cudaStreamSynchronize(st));
            for (size_t at = 0; at < ne; ++at) {
                const uint32_t s0 = entries[3 * ne + at], s1 = at + 1 < ne ? entries[3 * ne + at + 1] : (uint32_t)status.size();
                for (uint32_t sd = s0; sd < s1; ++sd) if (status[sd] == 4) { flags[entries[ne + at]] |= 1; break; }
            }
        }
        // ---- candidate position pairs (decodeMergePositions, DV-DPfunctions.cu:2963-2999)
        uint32_t *d_id = NULL, *d_l = NULL, *d_r = NULL;
        uint64_t nc = 0;
        {
            const uint64_t R = side.ranges.numRanges, R0 = side.ranges.rangesBeforeSplit;
            const uint32_t *in[2][7];
            for (int a = 0; a < 7; ++a) { in[0][a] = side.ranges.d_buf ? side.ranges.d_buf + a * R : NULL; in[1][a] = side.ranges.d_buf ? side.ranges.d_buf + a * R + R0 : NULL; }
            if ((rc = s3_seed_pair_candidates_any(ix, in[0], R0, in[1], R - R0, 1, 0xFFFFFFFFu, d_len, numReads,
                                                  par->insertLow, par->insertHigh, par->strandLeftLeg, par->strandRightLeg, &d_id, &d_l, &d_r, &nc))) goto done;
            d_round[round] = d_id; ncRound[round] = nc;                     // (d_l = d_id + 3 nc, d_r = d_id + 4 nc: see s3_seed_pair_candidates_any)
            if (nc && (d_l != d_id + 3 * nc || d_r != d_id + 4 * nc)) { s3_set_error("s3_deep_dp_align: candidate layout"); rc = S3_EINVAL; goto done; }
        }
        if (nc) {
            const size_t base = candID.size();
            candID.resize(base + nc);
            S3_TRYS(cudaMemcpyAsync(candID.data() + base, d_id, nc * 4, cudaMemcpyDeviceToHost, st));
            S3_TRYS(cudaStreamSynchronize(st));
            for (size_t c = base; c < base + nc; ++c) flags[candID[c] & ~1u] |= 2;
        }
        seed_side_free(ix, side);
        clk.lap("pair candidates");
        // ---- seeded / too many / unseeded pairs (performSeeding, DV-DPfunctions.cu:3105-3125)
        next.clear();
        for (size_t k = 0; k < ni; ++k) {
            const uint32_t e = pairIDs[input[k]];
            if (flags[e] & 2) continue;
            if (round == 0 && (flags[e] & 1)) next.push_back(input[k]); else unseeded.push_back(e);
        }
        input.swap(next);
    }
    {
        const uint64_t nc64 = candID.size();
        out->numCandidates = nc64;
        if (nc64 >= 0x7FFFFFF0ull) { s3_set_error("s3_deep_dp_align: too many candidates"); rc = S3_EINVAL; goto done; }
        const uint32_t nc = (uint32_t)nc64;
        if (nc) {
            // ---- left read, then the right read inside the window the left hit allows (DP2CPUAlgnThread, DV-DPfunctions.cu:3731-3800)
            const uint32_t maxRead = (maxLen / 4 + 1) * 4, maxDNA = maxRead + 2 * margin_of(maxRead) + 8;
            S3WinParams w;
            memset(&w, 0, sizeof w);
            w.insertLow = par->insertLow; w.insertHigh = par->insertHigh; w.leftLeg = par->strandLeftLeg; w.rightLeg = par->strandRightLeg;
            w.softClipLeft = par->softClipLeft; w.softClipRight = par->softClipRight;
            w.cutoff[0] = w.cutoff[1] = par->isDefaultThreshold ? -1 : par->dpScoreThreshold; w.maxDNALength = maxDNA; w.textLength = ix->textLength;
            auto win_of = [&](uint32_t *b, size_t m) { S3StageWin o; o.readID = b; o.start = b + m; o.dnaLen = b + 2 * m; o.clipLt = b + 3 * m; o.clipRt = b + 4 * m;
                                                       o.ancL = b + 5 * m; o.ancR = b + 6 * m; o.cand = b + 7 * m; o.cutoff = (int32_t *)(b + 8 * m); o.strand = (uint8_t *)(b + 9 * m); return o; };
            size_t scanTemp = 0;
            cub::DeviceScan::ExclusiveSum(NULL, scanTemp, (uint32_t *)NULL, (uint32_t *)NULL, (int)(nc + 1), st);
            S3_TRYS(cudaMallocAsync((void **)&d_c, (size_t)nc * 12, st));
            S3_TRYS(cudaMallocAsync((void **)&d_wl, (size_t)nc * 40 + 64, st));
            S3_TRYS(cudaMallocAsync((void **)&d_wr, (size_t)nc * 40 + 64, st));
            S3_TRYS(cudaMallocAsync((void **)&d_cnt, 2 * ((size_t)nc + 1) * 4, st));
            S3_TRYS(cudaMallocAsync(&d_tmp, scanTemp + 16, st));
            // the rounds' candidates one after the other (they never left the device): ids | left | right
            for (uint64_t r = 0, base = 0; r < 2; base += ncRound[r], ++r) {
                if (!ncRound[r]) continue;
                const uint64_t m = ncRound[r];
                S3_TRYS(cudaMemcpyAsync(d_c + base, d_round[r], m * 4, cudaMemcpyDeviceToDevice, st));
                S3_TRYS(cudaMemcpyAsync(d_c + nc + base, d_round[r] + 3 * m, m * 4, cudaMemcpyDeviceToDevice, st));
                S3_TRYS(cudaMemcpyAsync(d_c + 2 * (size_t)nc + base, d_round[r] + 4 * m, m * 4, cudaMemcpyDeviceToDevice, st));
            }
            const S3StageWin ol = win_of(d_wl, nc), orr = win_of(d_wr, nc);
            const unsigned nb = (nc + 255) / 256;
            s3_stage_win_left_kernel<<<nb, 256, 0, st>>>(w, nc, d_c, d_len, ol);
            S3_LAUNCHED(1);
            S3StageAligned al, ar;
            memset(&ar, 0, sizeof ar);
            if ((rc = s3_stage_align(ix, NULL, NULL, numReads, wordPerQuery, 0, maxRead, maxDNA, par->scores, 0, nc, ol.readID, ol.strand, ol.start, ol.dnaLen,
                                     ol.cutoff, ol.clipLt, ol.clipRt, ol.ancL, ol.ancR, d_len, NULL, &al))) goto done;
            clk.lap("left windows + DP");
            uint32_t *d_count = d_cnt, *d_off = d_cnt + nc + 1;
            S3_TRYS(cudaMemsetAsync(d_count + nc, 0, 4, st));
            s3_stage_win_right_kernel<false><<<nb, 256, 0, st>>>(w, nc, d_c, d_len, al.d_score, ol.start, al.d_hit, d_count, NULL, orr);
            S3_TRYS(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_count, d_off, (int)(nc + 1), st));
            s3_stage_win_right_kernel<true><<<nb, 256, 0, st>>>(w, nc, d_c, d_len, al.d_score, ol.start, al.d_hit, NULL, d_off, orr);
            S3_LAUNCHED(2);
            S3_TRYS(cudaGetLastError());
            uint32_t *h_nr = (uint32_t *)ix->pinnedCount + 12;
            S3_TRYS(cudaMemcpyAsync(h_nr, d_off + nc, 4, cudaMemcpyDeviceToHost, st));
            S3_TRYS(cudaStreamSynchronize(st));
            const uint32_t nr = *h_nr;
            if (nr && (rc = s3_stage_align(ix, NULL, NULL, numReads, wordPerQuery, 0, maxRead, maxDNA, par->scores, 1, nr, orr.readID, orr.strand, orr.start,
                                           orr.dnaLen, orr.cutoff, orr.clipLt, orr.clipRt, orr.ancL, orr.ancR, d_len, orr.cand, &ar))) goto done;
            clk.lap("right windows + DP");
            uint32_t nh = 0;
            uint64_t nruns = 0;
            for (uint32_t t = 0; t < nr; ++t) {
                if (ar.score[t] < ar.cutoff[t]) continue;
                const uint32_t c = ar.cand[t];
                ++nh; nruns += (al.runOff[c + 1] - al.runOff[c]) + (ar.runOff[t + 1] - ar.runOff[t]);
            }
            hits.reserve(nh);
            runs.resize(nruns);
            uint32_t at = 0;
            for (uint32_t t = 0; t < nr; ++t) {
                if (ar.score[t] < ar.cutoff[t]) continue;           // (the left read reached its cutoff: only those have a right window)
                const uint32_t c = ar.cand[t], left = candID[c], readSide = left & 1u;
                // fields _1 belong to the pair's first read, _2 to its mate, whichever is the left one
                s3_deep_dp_hit h;
                memset(&h, 0, sizeof h);
                h.readID = left - readSide;
                const uint32_t posLeft = al.start[c] + al.hit[c], posRight = ar.start[t] + ar.hit[t];
                const uint32_t roL = at, nL = al.runOff[c + 1] - al.runOff[c];
                memcpy(runs.data() + at, al.runs + al.runOff[c], (size_t)nL * 4); at += nL;
                const uint32_t roR = at, nR = ar.runOff[t + 1] - ar.runOff[t];
                memcpy(runs.data() + at, ar.runs + ar.runOff[t], (size_t)nR * 4); at += nR;
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
            clk.lap("records");
        }
    }
done:
    seed_side_free(ix, side);
    {
        void *p[] = {d_round[0], d_round[1], d_c, d_wl, d_wr, d_cnt, d_tmp};
        for (void *q : p) if (q) cudaFreeAsync(q, st);
    }
    if (rc) return rc;
    out->numHits = hits.size(); out->numRuns = runs.size(); out->numUnseeded = unseeded.size();
    out->hits = to_malloc(hits); out->runs = to_malloc(runs); out->unseeded = to_malloc(unseeded);
    if (!out->hits || !out->runs || !out->unseeded) { s3_deep_dp_result_free(out); s3_set_error("s3_deep_dp_align: out of host memory"); return S3_ENOMEM; }
    return S3_OK;
}

}  // namespace

extern "C" int s3_deep_dp_align(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery,
                                const uint32_t *pairReadIDs, uint64_t n, const s3_stage_params *par, s3_deep_dp_result *out)
{
    if (!out) { s3_set_error("s3_deep_dp_align: NULL result"); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    if (!ix || !queries || !readLengths || !par || (n && !pairReadIDs)) { s3_set_error("s3_deep_dp_align: NULL argument"); return S3_EINVAL; }
    if (!ix->loc.sa || !ix->loc.text) { s3_set_error("s3_deep_dp_align: the index was uploaded without its suffix array and packed text"); return S3_EINVAL; }
    out->numPairs = n;
    if (n == 0) return S3_OK;
    if (numReads >= 0xFFFFFFF0ull) { s3_set_error("s3_deep_dp_align: too many reads"); return S3_EINVAL; }
    std::vector<uint32_t> lens(2 * n);
    for (uint64_t k = 0; k < n; ++k) {
        const uint32_t e = pairReadIDs[k];
        if ((e & 1u) || (uint64_t)e + 1 >= numReads) { s3_set_error("s3_deep_dp_align: %u is not the even read id of a pair", e); return S3_EINVAL; }
        for (int i = 0; i < 2; ++i) {
            if (readLengths[e + i] > 16u * wordPerQuery) { s3_set_error("s3_deep_dp_align: read %u longer than its query words", e + i); return S3_EINVAL; }
            lens[2 * k + i] = readLengths[e + i];
        }
    }
    int rc = S3_OK;
    StageClock clk("s3_deep_dp_align");
    if (cudaSetDevice(ix->device) != cudaSuccess) { s3_set_error("s3_deep_dp_align: cudaSetDevice failed"); return S3_ECUDA; }
    cudaStream_t st = ix->stream;
    uint32_t *d_len = NULL;
    S3_TRYS(cudaMallocAsync((void **)&d_len, numReads * 4 + 16, st));
    S3_TRYS(cudaMemcpyAsync(d_len, readLengths, numReads * 4, cudaMemcpyHostToDevice, st));
    if ((rc = s3_stage_upload_queries(ix, queries, numReads, wordPerQuery))) goto done;
    clk.lap("uploads");
    rc = deep_dp_core(ix, d_len, numReads, wordPerQuery, pairReadIDs, lens.data(), n, par, out, clk);
done:
    if (d_len) cudaFreeAsync(d_len, st);
    return rc;
}

// The same stage for the batch a paired-end chain has just aligned: the pairs it left with no occurrence of either read
// (route S3_PE_BOTH_UNALIGNED) are picked on the device, and the stage works on the chain's own query buffer and read lengths.
int s3_stage_deep_dp_of_chain(s3_index *ix, const uint32_t *d_queries, const uint32_t *d_len, uint64_t numReads, uint32_t wordPerQuery, const uint8_t *d_route,
                              uint8_t wantRoute, const s3_stage_params *par, s3_deep_dp_result *out)
{
    memset(out, 0, sizeof *out);
    if (!ix->loc.sa || !ix->loc.text) { s3_set_error("s3_pe_deep_dp: the index was uploaded without its suffix array and packed text"); return S3_EINVAL; }
    if (numReads < 2) return S3_OK;
    int rc = S3_OK;
    StageClock clk("s3_pe_deep_dp");
    if (cudaSetDevice(ix->device) != cudaSuccess) { s3_set_error("s3_pe_deep_dp: cudaSetDevice failed"); return S3_ECUDA; }
    cudaStream_t st = ix->stream;
    const uint32_t P = (uint32_t)(numReads / 2);
    uint8_t *d_flags = NULL;
    uint32_t *d_pairs = NULL, *d_sel = NULL, *d_n = NULL;
    void *d_tmp = NULL;
    size_t t1 = 0;
    uint32_t n = 0;
    std::vector<uint32_t> sel;
    uint32_t *h_n = (uint32_t *)ix->pinnedCount + 13;
    S3_TRYS(cudaMallocAsync((void **)&d_flags, P, st));
    S3_TRYS(cudaMallocAsync((void **)&d_pairs, (size_t)P * 4 + 16, st));
    S3_TRYS(cudaMallocAsync((void **)&d_n, 16, st));
    cub::DeviceSelect::Flagged(NULL, t1, cub::CountingInputIterator<uint32_t>(0), d_flags, d_pairs, d_n, (int)P, st);
    S3_TRYS(cudaMallocAsync(&d_tmp, t1 + 16, st));
    s3_stage_pick_flag_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, d_route, wantRoute, d_flags);
    S3_LAUNCHED(1);
    S3_TRYS(cub::DeviceSelect::Flagged(d_tmp, t1, cub::CountingInputIterator<uint32_t>(0), d_flags, d_pairs, d_n, (int)P, st));
    S3_TRYS(cudaMemcpyAsync(h_n, d_n, 4, cudaMemcpyDeviceToHost, st));
    S3_TRYS(cudaStreamSynchronize(st));
    n = *h_n;
    out->numPairs = n;
    if (n) {
        S3_TRYS(cudaMallocAsync((void **)&d_sel, (size_t)n * 12, st));
        s3_stage_pick_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, d_pairs, d_len, d_sel);
        S3_LAUNCHED(1);
        sel.resize(3 * (size_t)n);
        S3_TRYS(cudaMemcpyAsync(sel.data(), d_sel, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
        S3_TRYS(cudaStreamSynchronize(st));
        std::vector<uint32_t> lens(2 * (size_t)n);
        for (uint32_t k = 0; k < n; ++k) {
            lens[2 * (size_t)k] = sel[n + k]; lens[2 * (size_t)k + 1] = sel[2 * (size_t)n + k];
            if (sel[n + k] > 16u * wordPerQuery || sel[2 * (size_t)n + k] > 16u * wordPerQuery) { s3_set_error("s3_pe_deep_dp: a read longer than its query words"); rc = S3_EINVAL; goto done; }
        }
        if ((rc = s3_stage_use_queries(ix, d_queries))) goto done;
        clk.lap("picking the pairs");
        rc = deep_dp_core(ix, d_len, numReads, wordPerQuery, sel.data(), lens.data(), n, par, out, clk);
    }
done:
    {
        void *p[] = {d_flags, d_pairs, d_sel, d_n, d_tmp};
        for (void *q : p) if (q) cudaFreeAsync(q, st);
    }
    return rc;
}
