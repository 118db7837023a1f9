// s3_common.cuh -- shared device structures of the B200 hot path.
//
// Index layout in HBM (DESIGN.md "data layout"): per direction an array of
// 32-byte, 32-byte-aligned buckets -- ONE DRAM sector -- one per 64 BWT positions:
//     bytes  0..15  cnt[4]        cnt[c] = cumulativeFreq[c] + Occ(c, 64*b + 32)   (count at the bucket MIDDLE)
//     bytes 16..23  hiA loA       bases  0..31 as two bit planes (base k = bit k)
//     bytes 24..31  hiB loB       bases 32..63
// A base's 2-bit code is split into a "hi" and a "lo" plane, so the four symbol counts of
// up to 32 bases cost 3 POPC (hi, lo, hi&lo), and rank'(c, i) counts forward or backward
// from the middle: ONE 256-bit load (LDG.E.256) of ONE sector per rank evaluation.  The
// reference needs a 16 B occ load and a 16 B BWT load from two different cache lines and
// counts 2-bit codes with 64-bit masks (DV-Kernel.cu:27-280).  Positions past the end of
// the text are padded with code 0 and counted as such on both sides of the middle, so
// they cancel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/soap3dp_b200.h"

#define S3_BUCKET_BASES 64u
#define S3_THREADS 128

struct S3Half {
    const uint4 *buckets;     // 2 x uint4 per bucket
    uint32_t inverseSa0;
    uint32_t numBuckets;
};

// Seed tables (DESIGN.md "seed tables"): the interval reached after the first K exact steps of a
// pass, for every K-mer in processing order (first base stepped = most significant digit).
//   fwd0 : steps on the BWT from (0, n)          -- the other-index interval of forward-first programs
//   fwd1 : steps on the BWT from (1, n)          -- backward-only programs (DV-Kernel.cu:3672)
//   rev0 : steps on the reverse BWT from (0, n)  -- forward-first programs (DV-Kernel.cu:3804)
// Entry = (lo, hi); an empty interval is stored as (0xFFFFFFF0 | s, 0), s = the step that emptied it,
// so that the rank-evaluation count of the stepwise search can still be reported.
struct S3Seed {
    const uint2 *fwd0, *fwd1, *rev0;
    uint32_t K;               // 0: no tables
};

// Copy/compute pipeline of the host-pointer entry points: a batch is cut into chunks; chunk c+1 goes up
// on `in` while chunk c computes on the handle's own stream and chunk c-1 comes down on `out`.
#define S3_PIPE_CHUNKS 16
struct S3Pipe {
    cudaStream_t in, out;
    cudaEvent_t up[S3_PIPE_CHUNKS], done[S3_PIPE_CHUNKS];
    int ready;
};
int s3_pipe_init(S3Pipe *p);
void s3_pipe_destroy(S3Pipe *p);

// Text-side arrays for check-and-extend: once an interval holds a single suffix, the
// rest of the read is compared with the text itself instead of being stepped through the index.
struct S3Locate {
    const uint32_t *sa;       // suffix array of the BWT rows (n + 1 entries, row 0 = the '$' suffix), NULL: disabled
    const uint32_t *isa;      // inverse: row of the suffix starting at a text position (n entries)
    const uint32_t *text;     // packed text, 16 bases per word, MSB first (hsp->packedDNA)
};

// Optional per-kernel timing (bench.py's roofline lines): events between the launches of a call, summed per kernel
// slot when read.  Off by default; costs nothing then.
#define S3_TIMING_SLOTS 8
#define S3_TIMING_MARKS 4096        // marks between two reads: ~13 per bench step
struct S3Timing {
    int on, n;
    cudaEvent_t ev[S3_TIMING_MARKS];     // created on first use
    int slot[S3_TIMING_MARKS];           // kernel slot that ENDS at mark i (-1: start of a run of launches)
    int created;
};
void s3_timing_mark(S3Timing *t, cudaStream_t st, int slot);
int s3_timing_read(S3Timing *t, cudaStream_t st, float *msPerSlot, int *launchesPerSlot);
void s3_timing_destroy(S3Timing *t);

struct s3_index {
    int device;
    cudaStream_t stream;
    S3Half fwd, rev;
    S3Seed seed;
    uint2 *d_seed[3];
    S3Locate loc;
    uint32_t *d_isa;
    uint32_t textLength;
    uint4 *d_fwd, *d_rev;
    uint32_t *d_packedDNA;    // optional
    uint32_t *d_sa;           // optional
    size_t bytes;
    // scratch reused across calls (grown on demand)
    void *scratch; size_t scratchBytes;
    void *pinned; size_t pinnedBytes;
    S3Pipe pipe;
    // persistent search launches
    uint32_t *d_workCounter;          // [0] work queue head, [1] number of items the easy kernel left behind, [4..7] S3_HV_* counters
    uint32_t *d_heavy; uint32_t heavyCap, heavyMaxTasks;     // scratch for splitting long enumerations (s3_search.cu)
    S3Timing timing;                  // slots: 0 easy, 1 items, 2 spine, 3 subtree, 4 merge, 5 isBad fixup
    // A second set of everything a search launch writes to, on a stream of its own: two halves of a batch (or two
    // chunks of a host call) are searched side by side, so the latency-bound end of one (its last long items,
    // spines, tasks) runs under the other's bandwidth-bound start.
    struct Side {
        cudaStream_t stream;
        uint32_t *d_workCounter, *d_hardItems; size_t hardCap;
        uint32_t *d_heavy; uint32_t heavyCap, heavyMaxTasks;
        cudaEvent_t fork, join;
        int ready;
    } side;
    int32_t splitBudget;              // steps before an item is split (s3_search_set_split_budget); < 0: never
    uint32_t *d_hardItems; size_t hardCap;
    uint32_t *d_itemStats; size_t itemStatsCap;   // S3_ITEM_STATS builds only (tools/search_tail.py)
    int numSms;
    size_t searchSmem; int searchBlocksPerSm;
    int sharedArrays;                 // s3_index_clone: buckets, seed tables, suffix array, text belong to another handle
    void *stageWs;                    // workspace of the seeded DP stages (s3_stage_align, s3_chain.cu), created on first use
    void *pinnedCount;                // 64 pinned bytes for the small count reads of stream-ordered entries
};

// the seeding driver and the single-end seed merge on device arrays (s3_seed.cu); see the host entries for what they compute
struct S3SeedRangesDev {
    uint32_t *d_buf;             // saL | saR | strand | read id | seed offset | seed length | read length, numRanges words each (cudaFreeAsync)
    uint64_t numRanges;
    uint8_t *d_status;           // per seed: 0 none, 1 kept, 4 too many occurrences (cudaFreeAsync)
    uint32_t splitSeed;          // in: a seed id (0: none); out: rangesBeforeSplit = the ranges of the seeds below it (ranges come in seed order,
    uint64_t rangesBeforeSplit;  // so two groups of seeds searched in one call are two stretches of d_buf's columns)
};
int s3_seed_search_device(s3_index *ix, const uint32_t *d_seeds, const uint32_t *d_seedLengths, uint32_t numSeeds, uint32_t wordPerSeed,
                          const uint32_t *d_maxHit, const uint32_t *d_seedReadID, const uint32_t *d_seedOffset, const uint32_t *d_seedReadLength,
                          S3SeedRangesDev *out);
int s3_seed_candidates_device(s3_index *ix, const uint32_t *d_l, const uint32_t *d_r, const int32_t *d_st, const uint32_t *d_rid,
                              const uint32_t *d_off, const uint32_t *d_sl, const uint32_t *d_rl, uint64_t numRanges, uint32_t maxPerRange,
                              uint32_t **d_out, uint32_t *numCandidates);
int s3_seed_pair_candidates_any(s3_index *ix, const uint32_t *const in0[7], uint64_t n0, const uint32_t *const in1[7], uint64_t n1, int onDevice,
                                uint32_t maxPerRange, const uint32_t *lengthsByReadID, uint64_t numReadIDs,
                                int insertLow, int insertHigh, int peStrandLeftLeg, int peStrandRightLeg,
                                uint32_t **candReadIDLeft, uint32_t **candPosLeft, uint32_t **candPosRight, uint64_t *numCandidates);
int s3_search_csr_device(s3_index *ix, const uint32_t *d_queries, const uint32_t *d_readLengths, uint32_t batchSize, uint32_t wordPerQuery,
                         uint32_t numMismatch, int isExactNumMismatch, unsigned long long *d_starts, uint32_t **d_out, unsigned long long *total);

// Alignment step of the seeded DP stages (s3_stages.cu): windows in host arrays -> scores, hit locations, tie counts and the
// CIGAR runs of the alignments that reach their cutoff, in host arrays the handle owns (two slots: the deep stage keeps the left
// reads' results while the right reads are aligned).  The DP workspace, the device copy of the query buffer and all scratch
// live in the handle's stage workspace and are reused from call to call.
struct S3StageAligned {
    const int32_t *score; const uint32_t *hit, *cnt, *runOff, *runs;
    uint64_t numRuns;
    // device mode only (windows given as device arrays): host copies of what the records need of the windows
    const uint32_t *readID, *start, *cand; const int32_t *cutoff; const uint8_t *strand;
    // device pointers of this call's scores / hit locations (valid until the next call on the handle): the deep stage's right windows read them
    const int32_t *d_score; const uint32_t *d_hit;
};
// d_readLengthsByRead == NULL: the window arrays are host arrays (uploaded here); else they are device arrays, d_cand (may be
// NULL) is copied back with them, and the read length of an alignment is looked up on the device.
int s3_stage_align(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery, int uploadQueries,
                   uint32_t maxRead, uint32_t maxDNA, s3_dp_scores scores, int slot, uint64_t n,
                   const uint32_t *readID, const uint8_t *strand, const uint32_t *start, const uint32_t *len, const int32_t *cutoff,
                   const uint32_t *clipLt, const uint32_t *clipRt, const uint32_t *ancL, const uint32_t *ancR,
                   const uint32_t *d_readLengthsByRead, const uint32_t *d_cand, S3StageAligned *out);
const uint32_t *s3_stage_queries(s3_index *ix);            // device copy of the query buffer of the current stage call (NULL before the first upload)
int s3_stage_upload_queries(s3_index *ix, const uint32_t *queries, uint64_t numReads, uint32_t wordPerQuery);
int s3_stage_use_queries(s3_index *ix, const uint32_t *d_queries);   // the stage works on the caller's device buffer instead
void s3_stage_ws_free(s3_index *ix);

void s3_set_error(const char *fmt, ...);
extern unsigned long long g_s3_launches;
#define S3_LAUNCHED(n) ((void)__atomic_fetch_add(&g_s3_launches, (unsigned long long)(n), __ATOMIC_RELAXED))
#define S3_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            s3_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,             \
                         cudaGetErrorString(e__));                                      \
            return S3_ECUDA;                                                            \
        }                                                                               \
    } while (0)

int s3_scratch(s3_index *ix, size_t bytes, void **out);
int s3_pinned(s3_index *ix, size_t bytes, void **out);

// ---- rank'(., idx) for all four symbols ------------------------------------
// Mathematically the reference's GPUBWTAllOccValue (DV-Kernel.cu:282-299):
// C[c] + #{c in BWT[0, idx)} with the "$ is not stored" shift.
// Two stages so that a caller can put the loads of several rank evaluations in flight
// before the first POPC waits on one of them.
struct S3Bucket {
    uint32_t c0, c1, c2, c3;     // counts at the bucket middle
    uint32_t hiA, loA, hiB, loB; // bit planes of bases 0..31 and 32..63
    uint32_t rem;                // position inside the bucket (0..63)
};

__device__ __forceinline__ S3Bucket s3_rank_load(const uint4 *buckets, uint32_t inverseSa0, uint32_t idx)
{
    S3Bucket k;
    idx -= (idx > inverseSa0);
    k.rem = idx & 63u;
    const uint4 *p = buckets + (size_t)(idx >> 6) * 2;
    // one sector, one instruction (LDG.E.256, sm_100+)
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(k.c0), "=r"(k.c1), "=r"(k.c2), "=r"(k.c3), "=r"(k.hiA), "=r"(k.loA), "=r"(k.hiB), "=r"(k.loB)
        : "l"(p));
    return k;
}

__device__ __forceinline__ void s3_rank_count(const S3Bucket &k, uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3)
{
    const bool isB = k.rem >= 32u;               // first half: count [x,32) and subtract; second: count [0,x) and add
    const uint32_t x = k.rem & 31u;
    const uint32_t m = isB ? ~(0xFFFFFFFFu << x) : (0xFFFFFFFFu << x);
    const uint32_t h = (isB ? k.hiB : k.hiA) & m, l = (isB ? k.loB : k.loA) & m;
    const uint32_t nHi = __popc(h), nLo = __popc(l), nT = __popc(h & l);
    const uint32_t len = isB ? x : 32u - x;
    const uint32_t cA = len - nHi - nLo + nT, cC = nLo - nT, cG = nHi - nT;
    o0 = isB ? k.c0 + cA : k.c0 - cA;
    o1 = isB ? k.c1 + cC : k.c1 - cC;
    o2 = isB ? k.c2 + cG : k.c2 - cG;
    o3 = isB ? k.c3 + nT : k.c3 - nT;
}

__device__ __forceinline__ void s3_rank4(const S3Half &h, uint32_t idx, uint32_t out[4])
{
    const S3Bucket k = s3_rank_load(h.buckets, h.inverseSa0, idx);
    s3_rank_count(k, out[0], out[1], out[2], out[3]);
}
