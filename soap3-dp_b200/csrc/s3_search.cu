// s3_search.cu -- GPU-2BWT exact / <=4-mismatch search for sm_100a.
//
// Replaces the three search kernels of the reference (DV-Kernel.cu:4249 kernel,
// :4505 kernel_4mismatch_1, :4741 kernel_4mismatch_2) and their host drivers
// perform_round{1,2}_alignment (alignment.cu:118-531).
//
// Design (not a port): the reference spells each (case, mismatch set) as its
// own force-inlined recursion (51 blocks) and runs one thread per read, so a
// warp waits for its slowest read and its lanes sit in different recursions.
// Here
//   * a case is a small *program* of phases (direction, read segment, min..max
//     substitutions) interpreted by ONE data-driven loop whose body is "one
//     LF-mapping step on both interval ends" (two rank evaluations);
//   * a lane is a depth-first enumerator with an explicit frame stack; every
//     iteration of the warp's loop first lets each lane make its cheap,
//     memory-free transitions (phase end, report, pop the next substitution
//     branch, strand flip) and then ALL lanes evaluate their two ranks together;
//   * work items (read, case) come from a global queue: a lane that finishes
//     its item takes the next one, so lanes stay busy whatever the length of
//     their own enumeration (persistent warps instead of one thread per read).
// The depth-first order of the reference -- substitutions in ascending symbol
// order *before* following the read's base, phases in program order, the
// second strand appended to the first -- is preserved, so the answer slots are
// bit-identical, including which ranges survive when a slot overflows.
#include "s3_common.cuh"
#include <chrono>
#include "../../include/soap3dp_b200.h"
#include <cub/device/device_select.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#define S3_MAX_PHASES 4
#define S3_MAX_DEPTH 4
#define S3_MAX_WPQ 64          // MAX_READ_LENGTH 1024 / 16 (definitions.h:42)

struct S3Phase { uint32_t dir, start, len, lo, hi; };

__host__ __device__ inline void s3_phase(S3Phase *ph, int &n, uint32_t dir, uint32_t start, uint32_t len,
                                         uint32_t lo, uint32_t hi)
{
    ph[n].dir = dir; ph[n].start = start; ph[n].len = len; ph[n].lo = lo; ph[n].hi = hi; ++n;
}

// Case programs: DV-Kernel.cu:3088-4245.  Region
// sizes are (int)(readLength * ratio) evaluated in double like the reference
// (definitions.h:97-113).  firstL: backward-only programs start from saL = 1,
// bi-directional ones from 0 (DV-Kernel.cu:3672 vs :3804).
__host__ __device__ inline int s3_case_program(uint32_t k, uint32_t cs, uint32_t L, bool exactNum,
                                               S3Phase *ph, uint32_t &firstL)
{
    int n = 0;
    const uint32_t B = 0, F = 1;
    firstL = 0;
    if (k == 0) {
        if (cs == 0) { firstL = 1; s3_phase(ph, n, B, 0, L, 0, 0); }
    } else if (k == 1) {
        const uint32_t X = (uint32_t)(int)(L * 0.5);
        if (cs == 0) { firstL = 1; s3_phase(ph, n, B, X, L - X, 0, 0); s3_phase(ph, n, B, 0, X, exactNum ? 1 : 0, 1); }
        else if (cs == 1) { s3_phase(ph, n, F, 0, X, 0, 0); s3_phase(ph, n, F, X, L - X, 1, 1); }
    } else if (k == 2) {
        const uint32_t X = (uint32_t)(int)(L * .3), Y = (uint32_t)(int)(L * .3), Z = L - X - Y;
        if (cs == 0) { firstL = 1; s3_phase(ph, n, B, X + Y, Z, 0, 0); s3_phase(ph, n, B, 0, X + Y, 0, 2); }
        else if (cs == 1) { s3_phase(ph, n, F, 0, X + Y, 0, 0); s3_phase(ph, n, F, X + Y, Z, 1, 2); }
        else if (cs == 2) { s3_phase(ph, n, F, 0, X, 0, 0); s3_phase(ph, n, F, X, Y, 1, 1); s3_phase(ph, n, F, X + Y, Z, 1, 1); }
        else if (cs == 3) { s3_phase(ph, n, F, X, Y, 0, 0); s3_phase(ph, n, F, X + Y, Z, 1, 1); s3_phase(ph, n, B, 0, X, 1, 1); }
    } else if (k == 3) {
        const uint32_t c1 = (uint32_t)(int)(L * .25), c2 = c1, c3 = c1, c4 = L - c1 - c2 - c3;
        if (cs == 0) { firstL = 1; s3_phase(ph, n, B, c1 + c2, c3 + c4, 0, 0); s3_phase(ph, n, B, 0, c1 + c2, 0, 3); }
        else if (cs == 1) { s3_phase(ph, n, F, 0, c1 + c2, 0, 0); s3_phase(ph, n, F, c1 + c2, c3 + c4, 1, 3); }
        else if (cs == 2) { s3_phase(ph, n, F, 0, c1, 0, 0); s3_phase(ph, n, F, c1, c2, 1, 1); s3_phase(ph, n, F, c1 + c2, c3 + c4, 2, 2); }
        else if (cs == 3) { s3_phase(ph, n, F, c1 + c2, c3, 0, 0); s3_phase(ph, n, F, c1 + c2 + c3, c4, 1, 1); s3_phase(ph, n, B, 0, c1 + c2, 1, 2); }
        else if (cs == 4) { firstL = 1; s3_phase(ph, n, B, c1 + c2 + c3, c4, 0, 0); s3_phase(ph, n, B, c1 + c2, c3, 1, 1); s3_phase(ph, n, B, 0, c1 + c2, 1, 2); }
        else if (cs == 5) { s3_phase(ph, n, F, c1, c2, 0, 0); s3_phase(ph, n, B, 0, c1, 1, 1); s3_phase(ph, n, F, c1 + c2, c3 + c4, 2, 2); }
    } else if (k == 4) {
        const uint32_t c1 = (uint32_t)(int)(L * .2), c2 = c1, c3 = c1, c4 = c1, c5 = L - 4 * c1;
        const uint32_t s3 = c1 + c2, s4 = s3 + c3, s5 = s4 + c4;
        if (cs == 0) { firstL = 1; s3_phase(ph, n, B, s4, c4 + c5, 0, 0); s3_phase(ph, n, B, 0, s4, 0, 4); }
        else if (cs == 1) { s3_phase(ph, n, F, 0, s4, 0, 0); s3_phase(ph, n, F, s4, c4 + c5, 1, 4); }
        else if (cs == 2) { s3_phase(ph, n, F, 0, c1, 0, 0); s3_phase(ph, n, F, c1, c2 + c3, 1, 1); s3_phase(ph, n, F, s4, c4 + c5, 1, 3); }
        else if (cs == 3) { s3_phase(ph, n, F, c1, c2 + c3, 0, 0); s3_phase(ph, n, B, 0, c1, 1, 1); s3_phase(ph, n, F, s4, c4 + c5, 1, 3); }
        else if (cs == 4) { s3_phase(ph, n, F, 0, c1, 0, 0); s3_phase(ph, n, F, c1, c2 + c3, 2, 2); s3_phase(ph, n, F, s4, c4 + c5, 1, 2); }
        else if (cs == 5) { s3_phase(ph, n, F, c1, c2 + c3, 0, 0); s3_phase(ph, n, B, 0, c1, 2, 2); s3_phase(ph, n, F, s4, c4 + c5, 1, 2); }
        else if (cs == 6) { s3_phase(ph, n, F, c1, c2, 0, 0); s3_phase(ph, n, B, 0, c1, 1, 1); s3_phase(ph, n, F, s3, c3, 1, 1); s3_phase(ph, n, F, s4, c4 + c5, 1, 2); }
        else if (cs == 7) { s3_phase(ph, n, F, s3, c3, 0, 0); s3_phase(ph, n, B, c1, c2, 1, 1); s3_phase(ph, n, B, 0, c1, 1, 1); s3_phase(ph, n, F, s4, c4 + c5, 1, 2); }
        else if (cs == 8) { s3_phase(ph, n, F, s4, c4, 0, 0); s3_phase(ph, n, F, s5, c5, 1, 1); s3_phase(ph, n, B, 0, s4, 3, 3); }
        else if (cs == 9) { firstL = 1; s3_phase(ph, n, B, s5, c5, 0, 0); s3_phase(ph, n, B, s4, c4, 1, 1); s3_phase(ph, n, B, 0, s4, 3, 3); }
    }
    return n;
}

// Splitting of long enumerations (see s3_search_kernel): scratch owned by the index handle.
#define S3_TASK_WORDS 16
#define S3_TASK_RESULTS 4          // ranges a task record holds: splitting needs saRangeAllowed <= 4 (every round-1 slot)
enum { S3_HV_ITEMS = 0, S3_HV_SPINE_HEAD = 1, S3_HV_TASKS = 2, S3_HV_TASK_HEAD = 3 };
struct S3Heavy {
    uint32_t *items;               // heavy item ids (case * numQueries + read)
    uint32_t cap;                  // room in items[]; 0: no splitting in this launch
    uint32_t maxTasks;             // task records per unit = (heavy item, strand pass)
    uint32_t *tasks;               // [cap * 2 * maxTasks][S3_TASK_WORDS]
    uint32_t *unitTasks;           // [cap * 2] task records written per unit
    uint32_t *queue;               // task ids (unit * maxTasks + k) waiting for a lane
    uint32_t *counters;            // S3_HV_*: heavy items, spine queue head, tasks queued, task queue head
    int32_t budget;                // LF-mapping steps an item may take in one lane before it is split
};

#define S3_CSR_SLOT 4u                 // ranges per item the counting pass of the capless search keeps
struct S3SearchArgs {
    const uint32_t *queries;
    const uint32_t *readLengths;
    uint32_t numQueries;
    uint32_t wordPerQuery;
    uint32_t *answers[S3_MAX_NUM_CASES];
    uint32_t round, numMismatch, saRangeAllowed, wordPerAnswer;
    uint32_t firstCase, numCases, exactNum;
    uint32_t textLength;
    uint32_t *workCounter;               // zeroed before the launch
    const uint32_t *itemList;            // NULL: every item; else item ids (case * numQueries + read) ...
    const uint32_t *itemCount;           // ... and how many (device memory, written by the easy kernel): [1] taken from the front
                                         // of the list (first phase left a wide interval: likely long), [2] from its back
    uint32_t itemCap;                    // entries the list has room for
    unsigned long long *rankQueries;     // may be NULL
    uint32_t *itemStats;                 // S3_ITEM_STATS builds only: LF-mapping steps spent per item
    S3Heavy heavy;
    // capless search (s3_search): per (read, case) -- read-major -- the number of ranges (first pass) and then
    // where its ranges start (second pass), and the range arrays
    unsigned long long *csrCount;
    uint32_t *csrL, *csrR, *csrInfo;
    uint32_t *csrSlot;                   // counting pass: the first S3_CSR_SLOT ranges of every item (L, R, info), so that only items with more are enumerated twice
};

// DFS frame: a node where substitutions are still allowed, kept in shared memory
// (S3_FRAME_WORDS words per frame, word-major so that lanes never conflict):
//   a[4], b[4]  rank vectors of both interval ends at the node, from which every child interval
//               (and the other index's interval update, DV-Kernel.cu:2673-2690) is derived
//   yhi         upper end of the node's interval in the index NOT being stepped
//   meta        done[0:11] p[11:13] mmp[13:16] mmt[16:19] next[19:22] c[22:24]
#define S3_FRAME_WORDS 10

__device__ __forceinline__ uint32_t s3_meta(uint32_t done, uint32_t p, uint32_t mmp, uint32_t mmt, uint32_t next, uint32_t c)
{
    return done | (p << 11) | (mmp << 13) | (mmt << 16) | (next << 19) | (c << 22);
}

// The read sits in shared memory twice, as the strand it was given in and as its reverse complement
// (what the reference materialises in place between the two strands, DV-Kernel.cu:4351-4395), both in
// the TEXT's packing -- 16 bases per word, first base in the top bits -- so that a stretch of the read
// can be XORed against the packed text.  Word w of a strand lives at sr[w * S3_THREADS].
__device__ __forceinline__ uint32_t s3_base(const uint32_t *sr, uint32_t pos)
{
    return (sr[(pos >> 4) * S3_THREADS] >> (30u - ((pos & 15u) << 1))) & 3u;
}

// 16 bases, first base in the low bits <-> first base in the high bits
__device__ __forceinline__ uint32_t s3_flip16(uint32_t w)
{
    w = __brev(w);
    return ((w & 0x55555555u) << 1) | ((w >> 1) & 0x55555555u);
}

__device__ __forceinline__ uint32_t s3_shr_clamp(uint32_t a, uint32_t s)
{
    uint32_t d;                                   // PTX shr clamps shift amounts > 32 to 32 (result 0)
    asm("shr.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(s));
    return d;
}

// a phase packed in one register: start[0:11] len[11:22] dir[22] lo[23:26] hi[26:29]
__device__ __forceinline__ uint32_t s3_pack_phase(const S3Phase &ph)
{
    return ph.start | (ph.len << 11) | (ph.dir << 22) | (ph.lo << 23) | (ph.hi << 26);
}

// Loads read q into shared memory in both orientations (see s3_base).  Returns its length.
__device__ __forceinline__ uint32_t s3_load_read(const S3SearchArgs &args, uint32_t q, uint32_t *sm0, uint32_t *sm1)
{
    const uint32_t *query = args.queries + (size_t)(q >> 5) * 32 * args.wordPerQuery + (q & 31);
    const uint32_t L = min(args.readLengths[q], 16u * args.wordPerQuery);      // a length the buffer cannot hold is cut, never followed
    const uint32_t nw = (L + 15) >> 4;
    // queries hold base i of a read in bits 2(i%16) of word i/16 (QueryParser.cpp:1146-1152)
    for (uint32_t w = 0; w < nw; ++w) sm0[w * S3_THREADS] = query[w * 32];
    // reverse complement, word j = bases 16j..16j+15 of it = the complemented 32-bit window of the
    // given read's bit stream that ENDS with base L-1-16j: the reversal of the base order is
    // exactly the change from first-base-low to first-base-high packing
    for (uint32_t j = 0; j < nw; ++j) {
        const int ob = 2 * ((int)L - 16 * (int)(j + 1));
        uint32_t win;
        if (ob >= 0) {
            const uint32_t wl = (uint32_t)ob >> 5, sh = (uint32_t)ob & 31u;
            const uint32_t lo = sm0[wl * S3_THREADS], hi = (wl + 1 < nw) ? sm0[(wl + 1) * S3_THREADS] : 0u;
            win = __funnelshift_r(lo, hi, sh);
        } else win = sm0[0] << (uint32_t)(-ob);
        sm1[j * S3_THREADS] = ~win;
    }
    for (uint32_t w = 0; w < nw; ++w) sm0[w * S3_THREADS] = s3_flip16(sm0[w * S3_THREADS]);
    return L;
}

// Check-and-extend.  The caller stands on a node whose interval is ONE suffix (row `row` of the forward
// BWT): every further step could only follow the text at that suffix' position, taking a substitution
// wherever read and text differ (if the phase still allows one) -- so the outcome of the whole remaining
// program is decided by the number of differences in each phase's stretch of the read, counted by XOR
// against the packed text, and the final interval is the row of the suffix that starts where the read
// starts.  Same answer as stepping (the reference's CPU search does the same once an interval is small,
// 2bwt-flex/SRA2BWTCheckAndExtend.c), ~5 sectors instead of one per remaining base.
// State: phases prog[0..nph), standing in phase p (direction pdir) with `done` bases of it consumed and
// mmp substitutions spent in it, mmt in total.  Returns true with the row and the total substitutions
// if the read survives.
__device__ __forceinline__ bool s3_check_extend(const S3Locate &loc, const uint32_t *sr, uint32_t L, uint32_t textLength,
                                                const uint32_t prog[S3_MAX_PHASES], uint32_t nph, uint32_t p, uint32_t pdir,
                                                uint32_t done, uint32_t mmp, uint32_t mmt, uint32_t row,
                                                uint32_t &outRow, uint32_t &outMm)
{
    const uint32_t ta = __ldg(loc.sa + row);                     // text position of the matched block
    // the matched block in read coordinates starts at the smallest position touched so far
    const uint32_t pstart = prog[p] & 0x7FFu, plen = (prog[p] >> 11) & 0x7FFu;
    uint32_t qa = (done > 0) ? (pdir ? pstart : pstart + plen - done) : 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < S3_MAX_PHASES - 1; ++k) if ((uint32_t)k < p) qa = min(qa, prog[k] & 0x7FFu);
    if (ta < qa || (unsigned long long)(ta - qa) + L > textLength) return false;      // hangs off the text
    const uint32_t ps = ta - qa;                                 // text position of read base 0
    // remaining stretches: the rest of the current phase, then the later phases
    uint32_t ra[S3_MAX_PHASES], rb[S3_MAX_PHASES], cnt[S3_MAX_PHASES];
#pragma unroll
    for (int k = 0; k < S3_MAX_PHASES; ++k) {
        const uint32_t st = prog[k] & 0x7FFu, ln = (prog[k] >> 11) & 0x7FFu;
        ra[k] = st; rb[k] = st + ln; cnt[k] = 0;
        if ((uint32_t)k == p) { if (pdir) ra[k] = st + done; else rb[k] = st + ln - done; }
        if ((uint32_t)k < p || (uint32_t)k >= nph) rb[k] = ra[k] = 0;                   // nothing left to check
    }
    const uint32_t *tw = loc.text + (ps >> 4);
    const uint32_t sh = (ps & 15u) << 1;
    const uint32_t nw = (L + 15) >> 4;
    uint32_t t0 = __ldg(tw);
    {
        // Most calls fail, and fail on the total alone: every substitution spent so far is a base where read and
        // text differ, so the stretches still to be checked hold (all differences - mmt) of them, and no more than
        // the phases from p on still allow can be taken.
        uint32_t allowed = 0, total = 0, u0 = t0;
#pragma unroll
        for (int k = 0; k < S3_MAX_PHASES; ++k) if ((uint32_t)k >= p && (uint32_t)k < nph) allowed += (prog[k] >> 26) & 7u;
        for (uint32_t w = 0; w < nw; ++w) {
            const uint32_t u1 = __ldg(tw + w + 1);
            const uint32_t x = sr[w * S3_THREADS] ^ __funnelshift_l(u1, u0, sh);
            uint32_t m = (x | (x >> 1)) & 0x55555555u;
            if (w + 1 == nw && (L & 15u)) m &= ~s3_shr_clamp(0xFFFFFFFFu, 2 * (L & 15u));      // bases past the read's end
            total += __popc(m);
            u0 = u1;
        }
        if (total > allowed - mmp + mmt) return false;
    }
    for (uint32_t w = 0; w < nw; ++w) {
        const uint32_t t1 = __ldg(tw + w + 1);
        const uint32_t x = sr[w * S3_THREADS] ^ __funnelshift_l(t1, t0, sh);             // 16 read bases vs 16 text bases
        const uint32_t m = (x | (x >> 1)) & 0x55555555u;                                // bit 30-2k: base k differs
        t0 = t1;
#pragma unroll
        for (int k = 0; k < S3_MAX_PHASES; ++k) {
            // bases [ra, rb) of the read that fall into this word
            const uint32_t k0 = max(ra[k], 16u * w) - 16u * w, k1 = min(rb[k], 16u * w + 16u);
            if (k1 > 16u * w + k0) {
                const uint32_t mask = s3_shr_clamp(0xFFFFFFFFu, 2 * k0) & ~s3_shr_clamp(0xFFFFFFFFu, 2 * (k1 - 16u * w));
                cnt[k] += __popc(m & mask);
            }
        }
    }
    uint32_t mm = mmt;
#pragma unroll
    for (int k = 0; k < S3_MAX_PHASES; ++k) {
        if ((uint32_t)k >= p && (uint32_t)k < nph) {
            const uint32_t lo = (prog[k] >> 23) & 7u, hi = (prog[k] >> 26) & 7u;
            const uint32_t tot = cnt[k] + (((uint32_t)k == p) ? mmp : 0u);
            if (tot < lo || tot > hi) return false;
            mm += cnt[k];
        }
    }
    outRow = (qa == 0) ? row : __ldg(loc.isa + ps);          // the matched block starts the read: same suffix, same row
    outMm = mm;
    return true;
}

// Packs the phases of (numMismatch, whichCase, L) into prog[]; returns their number.
__device__ __forceinline__ uint32_t s3_pack_program(const S3SearchArgs &args, uint32_t whichCase, uint32_t L,
                                                    uint32_t prog[S3_MAX_PHASES], uint32_t &firstL)
{
    S3Phase ph[S3_MAX_PHASES];
    const uint32_t nph = (uint32_t)s3_case_program(args.numMismatch, whichCase, L, args.exactNum != 0, ph, firstL);
#pragma unroll
    for (int k = 0; k < S3_MAX_PHASES; ++k) prog[k] = (k < (int)nph) ? s3_pack_phase(ph[k]) : 0u;
    return nph;
}

#ifndef S3_WIDE_MIN
#define S3_WIDE_MIN 16         // suffixes left after the easy kernel's walk from which an item counts as likely long
#endif
#ifndef S3_EASY_STEPS
#define S3_EASY_STEPS 16       // LF-mapping steps the easy kernel spends on a pass after its seed lookup
#endif

// First kernel of a search launch: one thread per (read, case), no work queue, no frame stack, every
// thread on the same straight path --
//     seed lookup -> a few exact steps until the interval is one suffix -> check-and-extend
// for the first strand and then the second.  That settles every item whose exact first phase
// pins the read to one place (or to none) -- almost all reads outside repeats.  Anything else (a first
// phase that is still ambiguous after S3_EASY_STEPS steps or at its end, a read too short for the seed
// table) is appended to `hardItems` untouched and enumerated by s3_search_kernel.
#ifndef S3_EASY_MIN_BLOCKS
#define S3_EASY_MIN_BLOCKS 8       // <= 64 registers: more reads in flight (the kernel runs at the random-burst rate of the memory system)
#endif
__global__ void __launch_bounds__(S3_THREADS, S3_EASY_MIN_BLOCKS)
s3_search_easy_kernel(const S3Half fwd, const S3Half rev, const S3Seed seed, const S3Locate loc, const S3SearchArgs args,
                      uint32_t *__restrict__ hardItems, uint32_t *__restrict__ hardCount)
{
    extern __shared__ uint32_t s3_smem[];
    uint32_t *sm0 = s3_smem + threadIdx.x, *sm1 = sm0 + args.wordPerQuery * S3_THREADS;
    const uint32_t item = blockIdx.x * S3_THREADS + threadIdx.x;
    if (item >= args.numQueries * args.numCases) return;
    const uint32_t ci = item / args.numQueries, q = item - ci * args.numQueries;
    const uint32_t whichCase = args.firstCase + ci;
    const uint32_t L = s3_load_read(args, q, sm0, sm1);
    uint32_t prog[S3_MAX_PHASES], firstL;
    const uint32_t nph = s3_pack_program(args, whichCase, L, prog, firstL);
    const uint32_t pstart = prog[0] & 0x7FFu, plen = (prog[0] >> 11) & 0x7FFu, pdir = (prog[0] >> 22) & 1u;
    const uint32_t K = seed.K;
    bool hard = nph == 0 || plen < K || ((prog[0] >> 23) & 63u) != 0;            // phase 0 must be exact and seedable
    uint32_t strand = args.round > 0 ? 0u : (whichCase & 1u);
    uint32_t repRow[2], repMeta[2], nrep = 0;
    bool wide = false;
    for (int pass = 0; pass < 2 && !hard; ++pass, strand ^= 1u) {
        const uint32_t *sr = strand ? sm1 : sm0;
        uint32_t key = 0, rkey = 0;
        for (uint32_t d = 0; d < K; ++d) {
            const uint32_t pos = pdir ? pstart + d : pstart + plen - 1 - d;
            const uint32_t c = s3_base(sr, pos);
            key = (key << 2) | c;
            rkey |= c << (2 * d);
        }
        const uint2 e = __ldg((pdir ? seed.rev0 : seed.fwd1) + key);
        if (e.x >= 0xFFFFFFF0u) continue;                                       // no occurrence of the seed
        uint32_t xlo = e.x, xhi = e.y, ylo = 0, yhi = 0, done = K;
        if (pdir) { ylo = __ldg(seed.fwd0 + rkey).x; yhi = ylo + (xhi - xlo); }
        const uint4 *buckets = pdir ? rev.buckets : fwd.buckets;
        const uint32_t isa0 = pdir ? rev.inverseSa0 : fwd.inverseSa0;
        bool alive = true;
        for (int s = 0; s < S3_EASY_STEPS && alive && xlo != xhi && done < plen; ++s) {
            const S3Bucket ka = s3_rank_load(buckets, isa0, xlo);
            const S3Bucket kb = s3_rank_load(buckets, isa0, xhi + 1);
            const uint32_t pos = pdir ? pstart + done : pstart + plen - 1 - done;
            const uint32_t c = s3_base(sr, pos);
            uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
            s3_rank_count(ka, a0, a1, a2, a3);
            s3_rank_count(kb, b0, b1, b2, b3);
            const uint32_t d1 = b1 - a1, d2 = b2 - a2, d3 = b3 - a3;
            const uint32_t cum = (c < 3 ? d3 : 0) + (c < 2 ? d2 : 0) + (c < 1 ? d1 : 0);
            xlo = (c == 0 ? a0 : c == 1 ? a1 : c == 2 ? a2 : a3) + 1;
            xhi = c == 0 ? b0 : c == 1 ? b1 : c == 2 ? b2 : b3;
            yhi = yhi - cum; ylo = yhi - (xhi - xlo);
            ++done;
            alive = xlo <= xhi;
        }
        if (!alive) continue;
        if (xlo != xhi) { hard = true; wide = xhi - xlo >= S3_WIDE_MIN; break; }    // still several suffixes: enumerate
        uint32_t row, mm;
        if (s3_check_extend(loc, sr, L, args.textLength, prog, nph, 0, pdir, done, 0, 0, pdir ? ylo : xlo, row, mm)) {
            repRow[nrep] = row; repMeta[nrep] = (strand << 27) + (mm << 24); ++nrep;
        }
    }
    if (hard) {
        // The enumerating kernel ends with its longest items: those whose first phase is still spread over many
        // suffixes (repeats) tend to be them, so they go to the front of the list and start first.
        if (wide) hardItems[atomicAdd(hardCount + 1, 1u)] = item;
        else hardItems[args.itemCap - 1u - atomicAdd(hardCount + 2, 1u)] = item;
        return;
    }
    // answer slot (DV-Kernel.cu:355-380,4468-4491); the buffer was filled with 0xFF before the launch
    uint32_t *answer = args.answers[whichCase] + (size_t)(q >> 5) * 32 * args.wordPerAnswer + (q & 31);
    uint32_t saCount = 0;
    for (uint32_t k = 0; k < nrep; ++k) {
        if (saCount < args.saRangeAllowed) { answer[32 * 2 * saCount] = repRow[k]; answer[32 * (2 * saCount + 1)] = repMeta[k]; }
        ++saCount;
    }
    if (saCount == 0) answer[0] = 0xFFFFFFFDu;
    else if (saCount > args.saRangeAllowed) answer[0] = 0xFFFFFFFEu;
}

#ifndef S3_REFILL_MIN
#define S3_REFILL_MIN 4        // idle lanes a warp tolerates before it goes back to the work queue
#endif
#ifndef S3_HEAVY_CAP
#define S3_HEAVY_CAP 4096      // split items per launch, at most (and S3_HEAVY_MB of task records)
#endif
#ifndef S3_HEAVY_MB
#define S3_HEAVY_MB 256
#endif
#ifndef S3_CE_BATCH
#define S3_CE_BATCH 8          // lanes of a warp that wait for each other before they check-and-extend together
#endif
#ifndef S3_SEARCH_MIN_BLOCKS
#define S3_SEARCH_MIN_BLOCKS 6
#endif

// One enumerator, three ways to feed it:
//   S3_MODE_ITEMS    work = (read, case) items, all of them or the list the easy kernel left.  An item that
//                    is still running after `heavy.budget` steps is taken back (its slot is blanked) and
//                    put on the heavy list: a launch would otherwise end with a few lanes walking one
//                    long enumeration each, one dependent DRAM access after the other.
//   S3_MODE_SPINE    work = (heavy item, strand pass).  The lane follows the read's own bases only; where
//                    the item's enumeration would branch into a substitution it writes the child's state
//                    down as a TASK instead, in enumeration order.
//   S3_MODE_SUBTREE  work = tasks.  The ordinary depth-first enumeration below one substitution child,
//                    results into the task record.
// s3_heavy_merge_kernel then strings the tasks' results together in enumeration order and applies the
// slot's cap, so the answer is bit-identical to what the single lane would have written.
#define S3_MODE_ITEMS 0
#define S3_MODE_SPINE 1
#define S3_MODE_SUBTREE 2

// CSR: 0 answer slots (the reference's format); 1 first pass of the capless search (count the ranges of every
// item, write nothing); 2 second pass (write them where the prefix sum of the counts says)
template <bool COUNT, int MODE, int CSR = 0>
__global__ void __launch_bounds__(S3_THREADS, S3_SEARCH_MIN_BLOCKS)
s3_search_kernel(const S3Half fwd, const S3Half rev, const S3Seed seed, const S3Locate loc, const S3SearchArgs args)
{
    extern __shared__ uint32_t s3_smem[];
    uint32_t *fr = s3_smem + threadIdx.x;                                  // frames: fr[(depth*10 + field) * S3_THREADS]
    uint32_t *sm0 = s3_smem + S3_MAX_DEPTH * S3_FRAME_WORDS * S3_THREADS + threadIdx.x;  // read words, given strand
    uint32_t *sm1 = sm0 + args.wordPerQuery * S3_THREADS;                                // reverse complement
    const uint32_t *sr = sm0;                                                            // strand being searched
    const uint32_t lane = threadIdx.x & 31;
    const S3Heavy &hv = args.heavy;
    uint32_t totalItems;                                                                 // < 2^32
    uint32_t *queueHead;
    if (MODE == S3_MODE_ITEMS) {
        // all (read, case) items, or the list the easy kernel left behind
        totalItems = args.itemList ? args.itemCount[1] + args.itemCount[2] : args.numQueries * args.numCases;
        queueHead = args.workCounter;
    } else if (MODE == S3_MODE_SPINE) {
        totalItems = 2 * min(hv.counters[S3_HV_ITEMS], hv.cap);
        queueHead = hv.counters + S3_HV_SPINE_HEAD;
    } else {
        totalItems = hv.counters[S3_HV_TASKS];
        queueHead = hv.counters + S3_HV_TASK_HEAD;
    }
    if (totalItems == 0) return;
    const uint32_t maxRanges = args.saRangeAllowed;
    unsigned long long nrank = 0;
#ifdef S3_ITEM_STATS
    uint32_t statItem = 0, statSteps = 0;
#define S3_STAT_END() do { if (MODE == S3_MODE_ITEMS && args.itemStats) args.itemStats[statItem] = statSteps; } while (0)
#define S3_STAT_STEP() (++statSteps)
#else
#define S3_STAT_END() do { } while (0)
#define S3_STAT_STEP() do { } while (0)
#endif

    // ---- per-lane enumerator state ----
    // (xlo, xhi) is the interval in the index the current phase steps through (BWT for a backward
    // phase, reverse BWT for a forward phase), (ylo, yhi) the interval in the other one.
    bool has = false, dead = false, alive = false;
    uint32_t L = 0, strand = 0, pass = 0, saCount = 0, firstL = 0, nph = 0;
    uint32_t prog[S3_MAX_PHASES] = {0, 0, 0, 0};
    uint32_t pstart = 0, plen = 0, pdir = 0, plo = 0, phi = 0;        // current phase, unpacked
    uint32_t p = 0, done = 0, mmp = 0, mmt = 0, depth = 0;
    uint32_t xlo = 0, xhi = 0, ylo = 0, yhi = 0;
    uint32_t *answer = NULL;
    uint32_t curItem = 0;            // ITEMS: the item id; SPINE: the unit (2 * heavy index + pass); SUBTREE: the task id
    int budget = 0;                  // ITEMS: steps left before the item is split
    uint32_t *task = NULL;           // SPINE: the unit's task block; SUBTREE: the task record

    auto load_phase = [&](uint32_t k) {
        const uint32_t w = (k == 0) ? prog[0] : (k == 1) ? prog[1] : (k == 2) ? prog[2] : prog[3];
        pstart = w & 0x7FF; plen = (w >> 11) & 0x7FF; plo = (w >> 23) & 7; phi = (w >> 26) & 7;
        const uint32_t nd = (w >> 22) & 1;
        if (nd != pdir) {                                             // the other index becomes the stepped one
            uint32_t t = xlo; xlo = ylo; ylo = t;
            t = xhi; xhi = yhi; yhi = t;
            pdir = nd;
        }
    };
    auto start_pass = [&]() {
        depth = 0; p = 0; done = 0; mmp = 0; mmt = 0;
        // backward-only programs start from saL = 1, bi-directional ones from 0 (DV-Kernel.cu:3672 vs :3804)
        pdir = 0; xlo = firstL; xhi = args.textLength; ylo = 0; yhi = args.textLength;
        alive = nph > 0;
        sr = strand ? sm1 : sm0;
        load_phase(0);
        // Seed tables: the first K steps of an exact first phase are one table lookup (the interval after K
        // steps is a function of the K bases alone).  Same intervals as stepping, K dependent loads fewer.
        const uint32_t K = seed.K;
        if (K && alive && plen >= K) {
            uint32_t key = 0, rkey = 0;
            for (uint32_t d = 0; d < K; ++d) {
                const uint32_t pos = pdir ? pstart + d : pstart + plen - 1 - d;
                const uint32_t c = s3_base(sr, pos);
                key = (key << 2) | c;
                rkey |= c << (2 * d);
            }
            const uint2 e = __ldg((pdir ? seed.rev0 : seed.fwd1) + key);
            uint32_t stepsDone = K;
            if (e.x >= 0xFFFFFFF0u) { alive = false; stepsDone = e.x & 15u; }
            else {
                xlo = e.x; xhi = e.y; done = K;
                if (pdir) { ylo = __ldg(seed.fwd0 + rkey).x; yhi = ylo + (xhi - xlo); }
            }
            if (COUNT) nrank += 2 * stepsDone;
        }
    };

    // A task record (S3_TASK_WORDS words): the enumerator state below one substitution child
    //   [0..3] xlo xhi ylo yhi   [4] done[0:11] p[11:13] mmp[13:16] mmt[16:19] pdir[19] strand[20]
    //   [5] number of ranges found (may be cap + 1)   [6..] up to S3_TASK_RESULTS ranges (saL, packed width word)
    auto emit_task = [&](uint32_t a_e, uint32_t b_e, uint32_t cum) {
        uint32_t *t = task + (size_t)saCount * S3_TASK_WORDS;
        const uint32_t tyhi = yhi - cum;
        t[0] = a_e + 1; t[1] = b_e; t[2] = tyhi - (b_e - (a_e + 1)); t[3] = tyhi;
        t[4] = (done + 1) | (p << 11) | ((mmp + 1) << 13) | ((mmt + 1) << 16) | (pdir << 19) | (strand << 20);
        t[5] = 0;
        hv.queue[atomicAdd(hv.counters + S3_HV_TASKS, 1u)] = curItem * hv.maxTasks + saCount;
        ++saCount;
    };

    // report (DV-Kernel.cu:355-380): the interval on the forward BWT; a slot overflow ends the item
    auto report = [&](uint32_t l, uint32_t r, uint32_t mm) {
        const uint32_t packed = (r - l) + (strand << 27) + (mm << 24);
        if (MODE == S3_MODE_ITEMS && CSR != 0) {
            if (CSR == 2) {
                const uint32_t ci = curItem / args.numQueries, q = curItem - ci * args.numQueries;
                const unsigned long long o = args.csrCount[(size_t)q * args.numCases + ci] + saCount;
                args.csrL[o] = l; args.csrR[o] = r;
                args.csrInfo[o] = strand | (mm << 1) | ((args.firstCase + ci) << 4);
            } else if (args.csrSlot && saCount < S3_CSR_SLOT) {
                const uint32_t ci = curItem / args.numQueries, q = curItem - ci * args.numQueries;
                uint32_t *slot = args.csrSlot + ((size_t)q * args.numCases + ci) * (3 * S3_CSR_SLOT) + 3 * saCount;
                slot[0] = l; slot[1] = r; slot[2] = strand | (mm << 1) | ((args.firstCase + ci) << 4);
            }
            ++saCount;                                               // no cap: every range of the item
        } else if (MODE == S3_MODE_ITEMS) {
            if (saCount < maxRanges) {
                answer[32 * 2 * saCount] = l;
                answer[32 * (2 * saCount + 1)] = packed;
            }
            ++saCount;
            if (saCount > maxRanges) { answer[0] = 0xFFFFFFFEu; has = false; S3_STAT_END(); }
        } else if (MODE == S3_MODE_SPINE) {
            // a range found on the spine itself takes its place among the tasks as one that is already done
            uint32_t *t = task + (size_t)saCount * S3_TASK_WORDS;
            t[5] = 1; t[6] = l; t[7] = packed;
            ++saCount;
        } else {
            if (saCount < maxRanges) { task[6 + 2 * saCount] = l; task[7 + 2 * saCount] = packed; }
            ++saCount;
            if (saCount > maxRanges) { task[5] = saCount; has = false; }
        }
    };

    auto check_extend = [&]() {
        uint32_t row, mm;
        alive = false;                                               // this branch ends here, reported or not
        if (s3_check_extend(loc, sr, L, args.textLength, prog, nph, p, pdir, done, mmp, mmt, pdir ? ylo : xlo, row, mm))
            report(row, row, mm);
    };

    while (true) {
        // ---- (R) refill idle lanes from the global queue ----
        const uint32_t want = __ballot_sync(0xFFFFFFFFu, !has && !dead);
        const uint32_t busy = __ballot_sync(0xFFFFFFFFu, has);
        if (!want && !busy) break;
        if (want && ((uint32_t)__popc(want) >= S3_REFILL_MIN || !busy)) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(queueHead, (uint32_t)__popc(want));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (!has && !dead) {
                const uint32_t item = base + __popc(want & ((1u << lane) - 1u));
                // the queue runs case-major so that a full warp refill reads 32 consecutive reads
                if (item >= totalItems || item < base) dead = true;
                else {
                    uint32_t it;                                         // (read, case) of this piece of work
                    if (MODE == S3_MODE_ITEMS) {
                        it = item;
                        if (args.itemList) {                             // front of the list first, then its back
                            const uint32_t nFront = args.itemCount[1];
                            it = item < nFront ? args.itemList[item] : args.itemList[args.itemCap - 1u - (item - nFront)];
                        }
                        curItem = it; budget = hv.budget;
                    }
                    else if (MODE == S3_MODE_SPINE) { curItem = item; it = hv.items[item >> 1]; }
                    else { curItem = hv.queue[item]; it = hv.items[(curItem / hv.maxTasks) >> 1]; }
#ifdef S3_ITEM_STATS
                    statItem = it; statSteps = 0;
#endif
                    const uint32_t ci = it / args.numQueries, q = it - ci * args.numQueries;
                    const uint32_t whichCase = args.firstCase + ci;
                    answer = args.answers[whichCase] + (size_t)(q >> 5) * 32 * args.wordPerAnswer + (q & 31);
                    L = s3_load_read(args, q, sm0, sm1);
                    nph = s3_pack_program(args, whichCase, L, prog, firstL);
                    // round 1: the device read buffer of the reference flips orientation after every
                    // launch, so odd cases meet the reverse strand first (DV-Kernel.cu:4280-4285)
                    strand = args.round > 0 ? 0u : (whichCase & 1u);
                    pass = 0; saCount = 0; has = true;
                    if (MODE == S3_MODE_ITEMS) start_pass();
                    else if (MODE == S3_MODE_SPINE) {
                        task = hv.tasks + (size_t)curItem * hv.maxTasks * S3_TASK_WORDS;
                        strand ^= (curItem & 1u);
                        start_pass();
                    } else {
                        task = hv.tasks + (size_t)curItem * S3_TASK_WORDS;
                        const uint32_t meta = task[4];
                        xlo = task[0]; xhi = task[1]; ylo = task[2]; yhi = task[3];
                        done = meta & 0x7FF; p = (meta >> 11) & 3; mmp = (meta >> 13) & 7; mmt = (meta >> 16) & 7;
                        pdir = (meta >> 19) & 1; strand = (meta >> 20) & 1;
                        sr = strand ? sm1 : sm0;
                        depth = 0; alive = true;
                        load_phase(p);                                   // same direction: nothing is swapped
                    }
                }
            }
        }
        // ---- (A) memory-free transitions until this lane stands on a node that needs its ranks ----
        while (has) {
            if (alive) {
                if (done < plen) break;
                // end of a phase
                if (mmp < plo) alive = false;
                else if (p + 1 == nph) { report(pdir ? ylo : xlo, pdir ? yhi : xhi, mmt); alive = false; }
                else { ++p; done = 0; mmp = 0; load_phase(p); }
            } else if (depth == 0) {
                if (MODE == S3_MODE_SPINE) { hv.unitTasks[curItem] = saCount; has = false; }
                else if (MODE == S3_MODE_SUBTREE) { task[5] = saCount; has = false; }
                // this strand is exhausted
                else if (pass == 0 && nph > 0) { pass = 1; strand ^= 1u; start_pass(); }
                else {
                    // status word (DV-Kernel.cu:4468-4491); the isBad carry between the cases of round 1
                    // is applied by s3_isbad_fixup_kernel because cases run concurrently here
                    if (CSR == 1) {
                        const uint32_t ci = curItem / args.numQueries, q = curItem - ci * args.numQueries;
                        args.csrCount[(size_t)q * args.numCases + ci] = saCount;
                    } else if (CSR == 0 && saCount == 0) answer[0] = 0xFFFFFFFDu;
                    has = false;
                    S3_STAT_END();
                }
            } else {
                // take the next pending branch from the innermost frame
                uint32_t *f = fr + (depth - 1) * S3_FRAME_WORDS * S3_THREADS;
                const uint32_t meta = f[9 * S3_THREADS];
                const uint32_t c = (meta >> 22) & 3;
                uint32_t e = (meta >> 19) & 7;
                const uint32_t fp = (meta >> 11) & 3;
                if (fp != p) { p = fp; load_phase(p); }          // intervals are overwritten below
                done = meta & 0x7FF; mmp = (meta >> 13) & 7; mmt = (meta >> 16) & 7;
                // next substitution symbol with a non-empty interval, ascending
                while (e < 4 && (e == c || f[e * S3_THREADS] + 1 > f[(4 + e) * S3_THREADS])) ++e;
                uint32_t sym;
                if (e < 4) { sym = e; f[9 * S3_THREADS] = (meta & ~(7u << 19)) | ((e + 1) << 19); ++mmp; ++mmt; }
                else { sym = c; --depth; }             // finally the read's own base; frame retired
                uint32_t cum = 0;
                for (uint32_t j = 3; j > sym; --j) cum += f[(4 + j) * S3_THREADS] - f[j * S3_THREADS];
                xlo = f[sym * S3_THREADS] + 1; xhi = f[(4 + sym) * S3_THREADS];
                yhi = f[8 * S3_THREADS] - cum; ylo = yhi - (xhi - xlo);
                ++done;
                alive = (xlo <= xhi);
            }
        }
        __syncwarp();
        // A lane whose interval is one suffix finishes its branch by check-and-extend: three dependent accesses
        // (suffix array, text, inverse) and a few hundred instructions that the rest of the warp sits through.
        // Such lanes wait until S3_CE_BATCH of them can go together, or nobody else has a step to make.
        const bool single = has && !COUNT && loc.sa != NULL && xlo == xhi;
        const uint32_t ceLanes = __ballot_sync(0xFFFFFFFFu, single);
        const uint32_t stepLanes = __ballot_sync(0xFFFFFFFFu, has && !single);
        const bool ceNow = (uint32_t)__popc(ceLanes) >= S3_CE_BATCH || !stepLanes;
        const bool act = has && (!single || ceNow);
        // ---- an item that outlasts its budget goes to the heavy list, if there is room ----
        if (MODE == S3_MODE_ITEMS && !COUNT && act && hv.cap && --budget < 0) {
            const uint32_t slot = atomicAdd(hv.counters + S3_HV_ITEMS, 1u);
            if (slot < hv.cap) {
                hv.items[slot] = curItem;
                for (uint32_t k = 0; k < 2 * min(saCount, maxRanges); ++k) answer[32 * k] = 0xFFFFFFFFu;    // take back what was written
                has = false;
                S3_STAT_END();
            } else budget = 0x7FFFFFFF;                               // no room: this lane finishes it alone
        }
        // ---- (B) one LF-mapping step for every lane that has work: both ranks' loads first ----
        if (has && act) S3_STAT_STEP();
        if (has && single) { if (ceNow) check_extend(); }
        else if (has) {
            const uint4 *buckets = pdir ? rev.buckets : fwd.buckets;
            const uint32_t isa0 = pdir ? rev.inverseSa0 : fwd.inverseSa0;
            const S3Bucket ka = s3_rank_load(buckets, isa0, xlo);
            const S3Bucket kb = s3_rank_load(buckets, isa0, xhi + 1);
            const uint32_t pos = pdir ? pstart + done : pstart + plen - 1 - done;
            const uint32_t c = s3_base(sr, pos);
            uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
            s3_rank_count(ka, a0, a1, a2, a3);
            s3_rank_count(kb, b0, b1, b2, b3);
            if (COUNT) nrank += 2;
            bool pushed = false;
            if (mmp < phi) {
                // does any substitution child survive?  (most do not once the interval is narrow)
                const bool live = (c != 0 && a0 < b0) || (c != 1 && a1 < b1) || (c != 2 && a2 < b2) || (c != 3 && a3 < b3);
                if (live && MODE == S3_MODE_SPINE) {
                    // the children become tasks, ascending like the frame would hand them out; the spine goes on
                    const uint32_t d1 = b1 - a1, d2 = b2 - a2, d3 = b3 - a3;
                    if (c != 0 && a0 < b0) emit_task(a0, b0, d1 + d2 + d3);
                    if (c != 1 && a1 < b1) emit_task(a1, b1, d2 + d3);
                    if (c != 2 && a2 < b2) emit_task(a2, b2, d3);
                    if (c != 3 && a3 < b3) emit_task(a3, b3, 0u);
                } else if (live) {
                    uint32_t *f = fr + depth * S3_FRAME_WORDS * S3_THREADS;
                    f[0 * S3_THREADS] = a0; f[1 * S3_THREADS] = a1; f[2 * S3_THREADS] = a2; f[3 * S3_THREADS] = a3;
                    f[4 * S3_THREADS] = b0; f[5 * S3_THREADS] = b1; f[6 * S3_THREADS] = b2; f[7 * S3_THREADS] = b3;
                    f[8 * S3_THREADS] = yhi;
                    f[9 * S3_THREADS] = s3_meta(done, p, mmp, mmt, 0, c);
                    ++depth;
                    alive = false;     // children are taken from the frame in (A)
                    pushed = true;
                }
            }
            if (!pushed) {
                // follow the read's base
                const uint32_t d1 = b1 - a1, d2 = b2 - a2, d3 = b3 - a3;
                const uint32_t cum = (c < 3 ? d3 : 0) + (c < 2 ? d2 : 0) + (c < 1 ? d1 : 0);
                const uint32_t ac = c == 0 ? a0 : c == 1 ? a1 : c == 2 ? a2 : a3;
                const uint32_t bc = c == 0 ? b0 : c == 1 ? b1 : c == 2 ? b2 : b3;
                xlo = ac + 1; xhi = bc;
                yhi = yhi - cum; ylo = yhi - (xhi - xlo);
                ++done;
                alive = (xlo <= xhi);
            }
        }
    }
    if (COUNT) {
        // warp-aggregate then one atomic per warp
        for (int o = 16; o > 0; o >>= 1) nrank += __shfl_down_sync(0xFFFFFFFFu, nrank, o);
        if (lane == 0 && nrank) atomicAdd(args.rankQueries, nrank);
    }
}

// The split items' answer slots: the tasks of pass 0, then those of pass 1, each in the order the spine
// wrote them = the order one lane would have found their ranges in; the slot's cap and status words as
// in the enumerator's own report (DV-Kernel.cu:355-380,4468-4491).
__global__ void s3_heavy_merge_kernel(const S3SearchArgs args)
{
    // one warp per split item: 32 task records at a time, an exclusive prefix sum of their range counts gives
    // every task its place in the slot
    const S3Heavy &hv = args.heavy;
    const uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (h >= min(hv.counters[S3_HV_ITEMS], hv.cap)) return;
    const uint32_t it = hv.items[h];
    const uint32_t ci = it / args.numQueries, q = it - ci * args.numQueries;
    uint32_t *answer = args.answers[args.firstCase + ci] + (size_t)(q >> 5) * 32 * args.wordPerAnswer + (q & 31);
    const uint32_t maxRanges = args.saRangeAllowed;
    uint32_t total = 0;
    for (uint32_t u = 2 * h; u < 2 * h + 2 && total <= maxRanges; ++u) {
        const uint32_t *tasks = hv.tasks + (size_t)u * hv.maxTasks * S3_TASK_WORDS;
        const uint32_t nt = hv.unitTasks[u];
        for (uint32_t k0 = 0; k0 < nt && total <= maxRanges; k0 += 32) {
            const uint32_t *t = tasks + (size_t)(k0 + lane) * S3_TASK_WORDS;
            const uint32_t cnt = (k0 + lane < nt) ? t[5] : 0u;
            uint32_t before = cnt;                                   // inclusive scan, then shifted
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, before, o); if ((int)lane >= o) before += v; }
            const uint32_t sum = __shfl_sync(0xFFFFFFFFu, before, 31);
            before = total + before - cnt;
            for (uint32_t r = 0; r < cnt && before + r < maxRanges; ++r) {
                answer[32 * 2 * (before + r)] = t[6 + 2 * r];
                answer[32 * (2 * (before + r) + 1)] = t[7 + 2 * r];
            }
            total += sum;
        }
    }
    __syncwarp();
    if (lane == 0) {
        if (total == 0) answer[0] = 0xFFFFFFFDu;
        else if (total > maxRanges) answer[0] = 0xFFFFFFFEu;
    }
}

// Round 1 semantics of isBad (DV-Kernel.cu:4285,4478-4491): once a read overflowed in
// case c, every later case reports a bare overflow slot without being searched.
__global__ void s3_isbad_fixup_kernel(S3SearchArgs args, uint32_t numCases)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= args.numQueries) return;
    const size_t off = (size_t)(q >> 5) * 32 * args.wordPerAnswer + (q & 31);
    bool bad = false;
    for (uint32_t c = 0; c < numCases; ++c) {
        uint32_t *answer = args.answers[c] + off;
        if (bad) {
            for (uint32_t i = 1; i < args.wordPerAnswer; ++i) answer[i * 32] = 0xFFFFFFFFu;
            answer[0] = 0xFFFFFFFEu;
        } else if (answer[0] == 0xFFFFFFFEu) bad = true;
    }
}

// Persistent launch: as many blocks as fit on the device at once (or fewer when the batch is
// small); every warp pulls (read, case) items from args.workCounter.  The caller has filled the
// answer buffers with 0xFF (the "unused" word, DV-Kernel.cu:4268-4276).
static int launch_search(s3_index *ix, S3SearchArgs &a, uint32_t numCases, bool count)
{
    if (a.numQueries == 0) return S3_OK;
    a.numCases = numCases;
    a.workCounter = ix->d_workCounter;
    const size_t smem = (size_t)(2 * a.wordPerQuery + S3_MAX_DEPTH * S3_FRAME_WORDS) * S3_THREADS * sizeof(uint32_t);
    if (smem != ix->searchSmem) {
        S3_CUDA(cudaFuncSetAttribute(s3_search_kernel<true, S3_MODE_ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        S3_CUDA(cudaFuncSetAttribute(s3_search_kernel<false, S3_MODE_ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        S3_CUDA(cudaFuncSetAttribute(s3_search_kernel<false, S3_MODE_SPINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        S3_CUDA(cudaFuncSetAttribute(s3_search_kernel<false, S3_MODE_SUBTREE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int perSm = 0;
        S3_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, s3_search_kernel<false, S3_MODE_ITEMS>, S3_THREADS, smem));
        if (perSm < 1) { s3_set_error("search kernel does not fit on an SM with wordPerQuery %u", a.wordPerQuery); return S3_EINVAL; }
        ix->searchSmem = smem;
        ix->searchBlocksPerSm = perSm;
    }
    // Scratch for splitting long enumerations: every substitution child of a heavy item's spine is a task
    // record, at most 3 per base and one range found on the spine itself per strand pass.
    memset(&a.heavy, 0, sizeof a.heavy);
    if (!count && a.saRangeAllowed <= S3_TASK_RESULTS && ix->splitBudget >= 0) {
        const uint32_t maxTasks = 3 * 16 * a.wordPerQuery + 2;
        if (maxTasks > ix->heavyMaxTasks) {
            if (ix->d_heavy) { S3_CUDA(cudaStreamSynchronize(ix->stream)); S3_CUDA(cudaFree(ix->d_heavy)); ix->d_heavy = NULL; ix->heavyMaxTasks = 0; }
            size_t cap = ((size_t)S3_HEAVY_MB << 20) / ((size_t)2 * maxTasks * S3_TASK_WORDS * 4);
            if (cap > S3_HEAVY_CAP) cap = S3_HEAVY_CAP;
            if (cap < 64) cap = 64;
            // items | unitTasks | queue | tasks
            const size_t words = cap + 2 * cap + 2 * cap * maxTasks + 2 * cap * maxTasks * S3_TASK_WORDS;
            S3_CUDA(cudaMalloc(&ix->d_heavy, words * sizeof(uint32_t)));
            ix->heavyCap = (uint32_t)cap; ix->heavyMaxTasks = maxTasks;
            ix->bytes += words * sizeof(uint32_t);
        }
        a.heavy.cap = ix->heavyCap; a.heavy.maxTasks = ix->heavyMaxTasks;
        a.heavy.items = ix->d_heavy;
        a.heavy.unitTasks = a.heavy.items + ix->heavyCap;
        a.heavy.queue = a.heavy.unitTasks + 2 * (size_t)ix->heavyCap;
        a.heavy.tasks = a.heavy.queue + 2 * (size_t)ix->heavyCap * ix->heavyMaxTasks;
        a.heavy.counters = ix->d_workCounter + 4;
        a.heavy.budget = ix->splitBudget;
    }
    const unsigned long long items = (unsigned long long)a.numQueries * numCases;
    unsigned long long blocks = (items + S3_THREADS - 1) / S3_THREADS;
    const unsigned long long resident = (unsigned long long)ix->numSms * ix->searchBlocksPerSm;
    if (blocks > resident) blocks = resident;
    S3_CUDA(cudaMemsetAsync(ix->d_workCounter, 0, 8 * sizeof(uint32_t), ix->stream));
    a.itemList = NULL; a.itemCount = NULL;
#ifdef S3_ITEM_STATS
    if (items > ix->itemStatsCap) {
        if (ix->d_itemStats) { S3_CUDA(cudaStreamSynchronize(ix->stream)); S3_CUDA(cudaFree(ix->d_itemStats)); ix->d_itemStats = NULL; }
        S3_CUDA(cudaMalloc(&ix->d_itemStats, items * sizeof(uint32_t)));
        ix->itemStatsCap = items;
    }
    S3_CUDA(cudaMemsetAsync(ix->d_itemStats, 0, items * sizeof(uint32_t), ix->stream));
    a.itemStats = ix->d_itemStats;
#endif
    if (!count && ix->loc.sa && ix->seed.K && !getenv("S3_NO_EASY_KERNEL")) {
        // the straight-line kernel settles what it can; the enumerating kernel takes the list it leaves
        if (items > ix->hardCap) {
            if (ix->d_hardItems) { S3_CUDA(cudaStreamSynchronize(ix->stream)); S3_CUDA(cudaFree(ix->d_hardItems)); ix->d_hardItems = NULL; ix->hardCap = 0; }
            S3_CUDA(cudaMalloc(&ix->d_hardItems, items * sizeof(uint32_t)));
            ix->hardCap = items;
        }
        const size_t smemEasy = (size_t)2 * a.wordPerQuery * S3_THREADS * sizeof(uint32_t);
        if (smemEasy > 48 * 1024) S3_CUDA(cudaFuncSetAttribute(s3_search_easy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemEasy));
        a.itemCap = (uint32_t)ix->hardCap;
        s3_timing_mark(&ix->timing, ix->stream, -1);
        s3_search_easy_kernel<<<(unsigned)((items + S3_THREADS - 1) / S3_THREADS), S3_THREADS, smemEasy, ix->stream>>>(
            ix->fwd, ix->rev, ix->seed, ix->loc, a, ix->d_hardItems, ix->d_workCounter + 1);
        s3_timing_mark(&ix->timing, ix->stream, 0);
        S3_LAUNCHED(1);
        S3_CUDA(cudaGetLastError());
        a.itemList = ix->d_hardItems; a.itemCount = ix->d_workCounter + 1;
    }
    s3_timing_mark(&ix->timing, ix->stream, -1);
    if (count) s3_search_kernel<true, S3_MODE_ITEMS><<<(unsigned)blocks, S3_THREADS, smem, ix->stream>>>(ix->fwd, ix->rev, ix->seed, ix->loc, a);
    else s3_search_kernel<false, S3_MODE_ITEMS><<<(unsigned)blocks, S3_THREADS, smem, ix->stream>>>(ix->fwd, ix->rev, ix->seed, ix->loc, a);
    s3_timing_mark(&ix->timing, ix->stream, 1);
    S3_LAUNCHED(1);
    S3_CUDA(cudaGetLastError());
    if (a.heavy.cap) {
        // the items that outlasted their budget: spines, then the tasks they wrote, then the slots.  The
        // three launches find their work counts in device memory and return at once when there is none.
        unsigned long long spineBlocks = (2ull * a.heavy.cap + S3_THREADS - 1) / S3_THREADS;
        if (spineBlocks > blocks) spineBlocks = blocks;
        s3_search_kernel<false, S3_MODE_SPINE><<<(unsigned)spineBlocks, S3_THREADS, smem, ix->stream>>>(ix->fwd, ix->rev, ix->seed, ix->loc, a);
        s3_timing_mark(&ix->timing, ix->stream, 2);
        s3_search_kernel<false, S3_MODE_SUBTREE><<<(unsigned)blocks, S3_THREADS, smem, ix->stream>>>(ix->fwd, ix->rev, ix->seed, ix->loc, a);
        s3_timing_mark(&ix->timing, ix->stream, 3);
        s3_heavy_merge_kernel<<<(a.heavy.cap + 3) / 4, 128, 0, ix->stream>>>(a);
        s3_timing_mark(&ix->timing, ix->stream, 4);
        S3_LAUNCHED(3);
        S3_CUDA(cudaGetLastError());
    }
    return S3_OK;
}

#ifdef S3_ITEM_STATS
// experiment builds only: steps per item of the last launch (0 = settled by the straight-line kernel)
extern "C" int s3_debug_item_stats(s3_index *ix, uint32_t *out, uint64_t n, uint32_t *hardCount)
{
    S3_CUDA(cudaStreamSynchronize(ix->stream));
    if (n > ix->itemStatsCap) n = ix->itemStatsCap;
    S3_CUDA(cudaMemcpy(out, ix->d_itemStats, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    uint32_t hc[2] = {0, 0};
    S3_CUDA(cudaMemcpy(hc, ix->d_workCounter + 2, sizeof hc, cudaMemcpyDeviceToHost));
    *hardCount = hc[0] + hc[1];
    return S3_OK;
}
#endif

extern "C" int s3_search_set_split_budget(s3_index *ix, int32_t steps)
{
    if (!ix) { s3_set_error("s3_search_set_split_budget: NULL index"); return S3_EINVAL; }
    ix->splitBudget = steps;
    return S3_OK;
}

static int check_search_args(const char *fn, s3_index *ix, const void *q, const void *len, uint64_t batchSize,
                             uint32_t wordPerQuery, uint32_t numMismatch, uint32_t numCases, uint32_t saRangeAllowed,
                             uint32_t wordPerAns)
{
    static const uint32_t ncases[5] = {1, 2, 4, 6, 10};
    if (!ix || !q || !len) { s3_set_error("%s: NULL argument", fn); return S3_EINVAL; }
    if (numMismatch > 4 || numCases == 0 || numCases > ncases[numMismatch]) {
        s3_set_error("%s: numMismatch %u / numCases %u out of range (definitions.h:116-120)", fn, numMismatch, numCases);
        return S3_EINVAL;
    }
    if (wordPerQuery == 0 || wordPerQuery > S3_MAX_WPQ) { s3_set_error("%s: wordPerQuery %u out of range", fn, wordPerQuery); return S3_EINVAL; }
    if (wordPerAns < 2 * saRangeAllowed || saRangeAllowed == 0) {
        s3_set_error("%s: wordPerAns %u < 2*saRangeAllowed %u", fn, wordPerAns, saRangeAllowed);
        return S3_EINVAL;
    }
    if (batchSize * numCases > 0xF0000000ull) { s3_set_error("%s: batch too large (batchSize * numCases must stay below 2^32)", fn); return S3_EINVAL; }
    return S3_OK;
}

// ---- the side context (s3_index::Side) ----------------------------------------
static int side_init(s3_index *ix)
{
    if (ix->side.ready) return S3_OK;
    S3_CUDA(cudaStreamCreateWithFlags(&ix->side.stream, cudaStreamNonBlocking));
    S3_CUDA(cudaEventCreateWithFlags(&ix->side.fork, cudaEventDisableTiming));
    S3_CUDA(cudaEventCreateWithFlags(&ix->side.join, cudaEventDisableTiming));
    S3_CUDA(cudaMalloc(&ix->side.d_workCounter, 256));
    ix->side.ready = 1;
    return S3_OK;
}

// launch_search and everything below it use the handle's own fields; the side context is swapped in for a launch
// (host code of one thread: a handle is never used by two threads at once)
static void side_swap(s3_index *ix)
{
    auto sw = [](auto &x, auto &y) { auto t = x; x = y; y = t; };
    sw(ix->stream, ix->side.stream);
    sw(ix->d_workCounter, ix->side.d_workCounter);
    sw(ix->d_hardItems, ix->side.d_hardItems); sw(ix->hardCap, ix->side.hardCap);
    sw(ix->d_heavy, ix->side.d_heavy); sw(ix->heavyCap, ix->side.heavyCap); sw(ix->heavyMaxTasks, ix->side.heavyMaxTasks);
}

static bool side_usable(const s3_index *ix) { return !ix->timing.on && !getenv("S3_NO_SIDE_STREAM"); }

// round 1 of `batchSize` reads on the handle's CURRENT stream and launch resources
static int round1_here(s3_index *ix, const uint32_t *d_queries, const uint32_t *d_readLengths, uint64_t batchSize,
                       uint32_t wordPerQuery, uint32_t numMismatch, uint32_t numCases, uint32_t saRangeAllowed,
                       uint32_t wordPerAns, int isExactNumMismatch, uint32_t *const *d_answers, unsigned long long *d_rankQueries)
{
    S3SearchArgs a;
    memset(&a, 0, sizeof a);
    a.queries = d_queries; a.readLengths = d_readLengths; a.numQueries = (uint32_t)batchSize;
    a.wordPerQuery = wordPerQuery;
    for (uint32_t c = 0; c < numCases; ++c) a.answers[c] = d_answers[c];
    a.round = 0; a.numMismatch = numMismatch; a.saRangeAllowed = saRangeAllowed; a.wordPerAnswer = wordPerAns;
    a.firstCase = 0; a.exactNum = isExactNumMismatch ? 1 : 0; a.textLength = ix->textLength;
    a.rankQueries = d_rankQueries;
    const size_t aBytes = (batchSize + 31) / 32 * 32 * wordPerAns * sizeof(uint32_t);
    for (uint32_t c = 0; c < numCases; ++c) S3_CUDA(cudaMemsetAsync(d_answers[c], 0xFF, aBytes, ix->stream));
    int rc;
    if ((rc = launch_search(ix, a, numCases, d_rankQueries != NULL))) return rc;
    if (numCases > 1 && batchSize > 0) {
        s3_timing_mark(&ix->timing, ix->stream, -1);
        s3_isbad_fixup_kernel<<<(unsigned)((batchSize + 255) / 256), 256, 0, ix->stream>>>(a, numCases);
        s3_timing_mark(&ix->timing, ix->stream, 5);
        S3_LAUNCHED(1);
        S3_CUDA(cudaGetLastError());
    }
    return S3_OK;
}

// The device entry point can search a batch as two halves side by side too (from S3_SIDE_MIN_READS reads on, set in
// the environment).  Off by default: measured on the bench workload it helps the host entry point, whose chunks wait
// for copies anyway (e2e 109.8 -> 113.9 M reads/s), and costs the device-resident pipeline 5 % (profiles/r02f).
static uint64_t side_min_reads()
{
    const char *e = getenv("S3_SIDE_MIN_READS");
    return e ? strtoull(e, NULL, 10) : ~0ull;
}

extern "C" int s3_search_round1_device(s3_index *ix, const uint32_t *d_queries, const uint32_t *d_readLengths,
                                       uint64_t batchSize, uint32_t wordPerQuery, uint32_t numMismatch,
                                       uint32_t numCases, uint32_t saRangeAllowed, uint32_t wordPerAns,
                                       int isExactNumMismatch, uint32_t *const *d_answers,
                                       unsigned long long *d_rankQueries)
{
    int rc = check_search_args("s3_search_round1_device", ix, d_queries, d_readLengths, batchSize, wordPerQuery,
                               numMismatch, numCases, saRangeAllowed, wordPerAns);
    if (rc) return rc;
    S3_CUDA(cudaSetDevice(ix->device));
    if (batchSize < side_min_reads() || d_rankQueries != NULL || !side_usable(ix))
        return round1_here(ix, d_queries, d_readLengths, batchSize, wordPerQuery, numMismatch, numCases, saRangeAllowed,
                           wordPerAns, isExactNumMismatch, d_answers, d_rankQueries);
    // two halves of whole 32-read groups (the interleave unit of the buffers), the second on the side stream
    if ((rc = side_init(ix))) return rc;
    const uint64_t half = ((batchSize / 2 + 31) / 32) * 32;
    uint32_t *ansB[S3_MAX_NUM_CASES];
    for (uint32_t c = 0; c < numCases; ++c) ansB[c] = d_answers[c] + half * wordPerAns;
    S3_CUDA(cudaEventRecord(ix->side.fork, ix->stream));              // whatever the caller queued before this call
    S3_CUDA(cudaStreamWaitEvent(ix->side.stream, ix->side.fork, 0));
    if ((rc = round1_here(ix, d_queries, d_readLengths, half, wordPerQuery, numMismatch, numCases, saRangeAllowed, wordPerAns,
                          isExactNumMismatch, d_answers, NULL))) return rc;
    side_swap(ix);
    rc = round1_here(ix, d_queries + half * wordPerQuery, d_readLengths + half, batchSize - half, wordPerQuery, numMismatch, numCases,
                     saRangeAllowed, wordPerAns, isExactNumMismatch, ansB, NULL);
    cudaError_t e = (rc == S3_OK) ? cudaEventRecord(ix->side.join, ix->stream) : cudaSuccess;
    side_swap(ix);
    if (rc) return rc;
    S3_CUDA(e);
    S3_CUDA(cudaStreamWaitEvent(ix->stream, ix->side.join, 0));
    return S3_OK;
}

extern "C" int s3_search_round1(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths,
                                uint64_t batchSize, uint32_t wordPerQuery, uint32_t numMismatch,
                                uint32_t numCases, uint32_t saRangeAllowed, uint32_t wordPerAns,
                                int isExactNumMismatch, uint32_t *const *answers)
{
    int rc = check_search_args("s3_search_round1", ix, queries, readLengths, batchSize, wordPerQuery,
                               numMismatch, numCases, saRangeAllowed, wordPerAns);
    if (rc) return rc;
    if (batchSize == 0) return S3_OK;
    S3_CUDA(cudaSetDevice(ix->device));
    if ((rc = s3_pipe_init(&ix->pipe))) return rc;
    const size_t roundUp = (batchSize + 31) / 32 * 32;
    const size_t qBytes = roundUp * wordPerQuery * 4, lBytes = roundUp * 4, aBytes = roundUp * wordPerAns * 4;
    char *d;
    if ((rc = s3_scratch(ix, qBytes + lBytes + aBytes * numCases + 256, (void **)&d))) return rc;
    uint32_t *d_q = (uint32_t *)d, *d_l = (uint32_t *)(d + qBytes);
    uint32_t *d_ans[S3_MAX_NUM_CASES];
    for (uint32_t c = 0; c < numCases; ++c) d_ans[c] = (uint32_t *)(d + qBytes + lBytes + aBytes * c);
    // Chunks of whole 32-read groups (the interleave unit of both buffers, so a chunk is a contiguous
    // slice of each): chunk k+1 is copied in while chunk k is searched and chunk k-1 is copied out.
    // (The reference copies, launches per case and copies back strictly in sequence, alignment.cu:158-215.)
    // Two chunks only: a launch needs several items per resident lane to keep the persistent warps busy.
    size_t chunk = (batchSize >= 262144) ? ((roundUp / 2 + 31) / 32) * 32 : roundUp;
    if (getenv("S3_HOST_CHUNKS")) { const int c = atoi(getenv("S3_HOST_CHUNKS")); if (c >= 1 && c <= S3_PIPE_CHUNKS) chunk = ((roundUp / c + 31) / 32) * 32; }   // tuning experiments
    S3Pipe &pp = ix->pipe;
    S3_CUDA(cudaEventRecord(pp.done[0], ix->stream));               // earlier work on the scratch buffer
    S3_CUDA(cudaStreamWaitEvent(pp.in, pp.done[0], 0));
    const bool useSide = chunk < batchSize && side_usable(ix);
    if (useSide && (rc = side_init(ix))) return rc;
    int k = 0;
    for (size_t c0 = 0; c0 < batchSize; c0 += chunk, k = (k + 1) % S3_PIPE_CHUNKS) {
        const size_t cnt = (batchSize - c0 < chunk) ? batchSize - c0 : chunk, cntUp = (cnt + 31) / 32 * 32;
        S3_CUDA(cudaMemcpyAsync(d_q + c0 * wordPerQuery, queries + c0 * wordPerQuery, cntUp * wordPerQuery * 4,
                                cudaMemcpyHostToDevice, pp.in));
        // the reference copies roundUp lengths (alignment.cu:161); only batchSize are meaningful
        S3_CUDA(cudaMemcpyAsync(d_l + c0, readLengths + c0, cnt * 4, cudaMemcpyHostToDevice, pp.in));
        S3_CUDA(cudaEventRecord(pp.up[k], pp.in));
        uint32_t *d_sub[S3_MAX_NUM_CASES];
        for (uint32_t c = 0; c < numCases; ++c) d_sub[c] = d_ans[c] + c0 * wordPerAns;
        // odd chunks on the side stream: the end of one chunk's search runs under the start of the next
        const bool side = (k & 1) && useSide;
        if (side) side_swap(ix);
        cudaError_t e = cudaStreamWaitEvent(ix->stream, pp.up[k], 0);
        rc = (e == cudaSuccess) ? round1_here(ix, d_q + c0 * wordPerQuery, d_l + c0, cnt, wordPerQuery, numMismatch, numCases,
                                              saRangeAllowed, wordPerAns, isExactNumMismatch, d_sub, NULL) : S3_OK;
        if (e == cudaSuccess && rc == S3_OK) e = cudaEventRecord(pp.done[k], ix->stream);
        if (side) side_swap(ix);
        if (rc) return rc;
        S3_CUDA(e);
        S3_CUDA(cudaStreamWaitEvent(pp.out, pp.done[k], 0));
        for (uint32_t c = 0; c < numCases; ++c)
            S3_CUDA(cudaMemcpyAsync(answers[c] + c0 * wordPerAns, d_sub[c], cntUp * wordPerAns * 4, cudaMemcpyDeviceToHost, pp.out));
    }
    S3_CUDA(cudaStreamSynchronize(pp.out));
    S3_CUDA(cudaStreamSynchronize(ix->stream));
    if (useSide) S3_CUDA(cudaStreamSynchronize(ix->side.stream));
    return S3_OK;
}

// ---- round 2 ----------------------------------------------------------------
__global__ void s3_bad_flags_kernel(const uint32_t *__restrict__ answers, uint32_t n, uint32_t wordPerAns,
                                    uint8_t *__restrict__ flags)
{
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    flags[q] = answers[(size_t)(q >> 5) * 32 * wordPerAns + (q & 31)] > 0xFFFFFFFDu;
}

__global__ void s3_gather_bad_kernel(const uint32_t *__restrict__ queries, const uint32_t *__restrict__ readLengths,
                                     const uint32_t *__restrict__ badIdx, uint32_t numBad, uint32_t wordPerQuery,
                                     uint32_t *__restrict__ outQ, uint32_t *__restrict__ outLen)
{
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= numBad) return;
    uint32_t q = badIdx[k];
    const uint32_t *src = queries + (size_t)(q >> 5) * 32 * wordPerQuery + (q & 31);
    uint32_t *dst = outQ + (size_t)(k >> 5) * 32 * wordPerQuery + (k & 31);
    for (uint32_t w = 0; w < wordPerQuery; ++w) dst[w * 32] = src[w * 32];
    outLen[k] = readLengths[q];
}

// ---- capless search ---------------------------------------------------------
// Replaces round 1 + round 2 + the CPU fallback for reads that still overflow (CPUfunctions.cpp:1310-1329,
// 1394-1412) by ONE call without slot caps (SURVEY.md 8b "preferred new path"): every SA range of every read
// for every case of the numMismatch scheme, as CSR.  Two passes of the enumerator over all items: count, prefix
// sum, fill.  Order inside a read: case ascending, inside a case the enumeration order of round 1 (first strand =
// case parity, DV-Kernel.cu:4280-4285), so the reference's slot contents are the first saRangeAllowed entries of a
// case's run and any order-dependent truncation can be replayed by the caller.
// After the counting pass and the prefix sum: an item with at most S3_CSR_SLOT ranges takes them from its slot; the others are listed
// (item id = case * numQueries + read, the enumerator's numbering) for the second pass.
__global__ void s3_csr_scatter_kernel(uint32_t items, uint32_t numQueries, uint32_t numCases, const unsigned long long *__restrict__ starts,
                                      const uint32_t *__restrict__ slot, uint32_t *__restrict__ csrL, uint32_t *__restrict__ csrR, uint32_t *__restrict__ csrInfo,
                                      uint32_t *__restrict__ list)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;             // = read * numCases + case
    if (i >= items) return;
    const unsigned long long a = starts[i], n = starts[i + 1] - a;
    if (n == 0) return;
    if (n <= S3_CSR_SLOT) {
        const uint32_t *src = slot + (size_t)i * (3 * S3_CSR_SLOT);
        for (uint32_t g = 0; g < (uint32_t)n; ++g) { csrL[a + g] = src[3 * g]; csrR[a + g] = src[3 * g + 1]; csrInfo[a + g] = src[3 * g + 2]; }
    } else {
        const uint32_t q = i / numCases, ci = i - q * numCases;
        list[4 + atomicAdd(list + 1, 1u)] = ci * numQueries + q;
    }
}

template <int PASS>
static int launch_search_csr(s3_index *ix, S3SearchArgs &a, uint32_t numCases)
{
    a.numCases = numCases;
    a.workCounter = ix->d_workCounter;
    if (PASS == 1) { a.itemList = NULL; a.itemCount = NULL; }       // (the second pass runs over the list the scatter kernel made)
    memset(&a.heavy, 0, sizeof a.heavy);
    const size_t smem = (size_t)(2 * a.wordPerQuery + S3_MAX_DEPTH * S3_FRAME_WORDS) * S3_THREADS * sizeof(uint32_t);
    S3_CUDA(cudaFuncSetAttribute(s3_search_kernel<false, S3_MODE_ITEMS, PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSm = 0;
    S3_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, s3_search_kernel<false, S3_MODE_ITEMS, PASS>, S3_THREADS, smem));
    if (perSm < 1) { s3_set_error("search kernel does not fit on an SM with wordPerQuery %u", a.wordPerQuery); return S3_EINVAL; }
    const unsigned long long items = (unsigned long long)a.numQueries * numCases;
    unsigned long long blocks = (items + S3_THREADS - 1) / S3_THREADS;
    if (blocks > (unsigned long long)ix->numSms * perSm) blocks = (unsigned long long)ix->numSms * perSm;
    S3_CUDA(cudaMemsetAsync(ix->d_workCounter, 0, 8 * sizeof(uint32_t), ix->stream));
    s3_search_kernel<false, S3_MODE_ITEMS, PASS><<<(unsigned)blocks, S3_THREADS, smem, ix->stream>>>(ix->fwd, ix->rev, ix->seed, ix->loc, a);
    S3_LAUNCHED(1);
    S3_CUDA(cudaGetLastError());
    return S3_OK;
}

extern "C" void s3_search_result_free(s3_search_result *r)
{
    if (!r) return;
    free(r->offsets); free(r->saL); free(r->saR); free(r->info);
    memset(r, 0, sizeof *r);
}

// The capless search on device arrays (the seeded DP stages keep their seeds there; s3_search below is its host-pointer
// form).  d_starts receives batchSize * numCases + 1 run starts (item = read * numCases + case; the last entry is the total),
// *d_out a stream-ordered allocation of 3 * total words (L | R | info) that the caller returns with cudaFreeAsync on the
// index stream (NULL when nothing was found).  One 8-byte read back sizes the second pass.
int s3_search_csr_device(s3_index *ix, const uint32_t *d_queries, const uint32_t *d_readLengths, uint32_t batchSize, uint32_t wordPerQuery,
                         uint32_t numMismatch, int isExactNumMismatch, unsigned long long *d_starts, uint32_t **d_out, unsigned long long *total)
{
    static const uint32_t ncases[5] = {1, 2, 4, 6, 10};
    *d_out = NULL; *total = 0;
    if (batchSize == 0) return S3_OK;
    const uint32_t numCases = ncases[numMismatch];
    const size_t items = (size_t)batchSize * numCases;
    cudaStream_t st = ix->stream;
    size_t scanTemp = 0;
    cub::DeviceScan::ExclusiveSum(NULL, scanTemp, (unsigned long long *)NULL, (unsigned long long *)NULL, (int)(items + 1), st);
    void *d_tmp = NULL;
    // the counting pass keeps the first S3_CSR_SLOT ranges of every item; the second pass enumerates only the items with more
    uint32_t *d_slot = NULL, *d_list = NULL;
    if (cudaMallocAsync(&d_tmp, scanTemp + 16, st) != cudaSuccess || cudaMallocAsync((void **)&d_slot, items * (3 * S3_CSR_SLOT) * 4, st) != cudaSuccess ||
        cudaMallocAsync((void **)&d_list, (items + 4) * 4, st) != cudaSuccess || cudaMemsetAsync(d_starts, 0, (items + 1) * 8, st) != cudaSuccess) {
        s3_set_error("s3_search: scratch for %zu items: %s", items, cudaGetErrorString(cudaGetLastError()));
        if (d_tmp) cudaFreeAsync(d_tmp, st);
        if (d_slot) cudaFreeAsync(d_slot, st);
        if (d_list) cudaFreeAsync(d_list, st);
        return S3_ECUDA;
    }
    S3SearchArgs a;
    memset(&a, 0, sizeof a);
    a.queries = d_queries; a.readLengths = d_readLengths; a.numQueries = batchSize; a.wordPerQuery = wordPerQuery;
    a.round = 0; a.numMismatch = numMismatch; a.saRangeAllowed = 0xFFFFFFFFu; a.wordPerAnswer = 0;
    a.firstCase = 0; a.exactNum = isExactNumMismatch ? 1 : 0; a.textLength = ix->textLength;
    a.csrCount = d_starts; a.csrSlot = d_slot;
    int rc = launch_search_csr<1>(ix, a, numCases);
    if (rc == S3_OK && cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_starts, d_starts, (int)(items + 1), st) != cudaSuccess) { s3_set_error("s3_search: scan failed"); rc = S3_ECUDA; }
    cudaFreeAsync(d_tmp, st);
    unsigned long long *h_total = (unsigned long long *)ix->pinnedCount;
    if (rc == S3_OK && (cudaMemcpyAsync(h_total, d_starts + items, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)) {
        s3_set_error("s3_search: reading the total failed"); rc = S3_ECUDA;
    }
    if (rc == S3_OK) *total = *h_total;
    if (rc == S3_OK && *total >= 0x7FFFFFFFull) { s3_set_error("s3_search: %llu ranges in one call", *total); rc = S3_EINVAL; }
    if (rc == S3_OK && *total) {
        if (cudaMallocAsync((void **)d_out, (size_t)*total * 12, st) != cudaSuccess) { s3_set_error("s3_search: out of device memory"); rc = S3_ECUDA; }
        else {
            a.csrL = *d_out; a.csrR = *d_out + *total; a.csrInfo = *d_out + 2 * *total;
            // list[0..3] = the second pass's item counters (front count at [1], as the enumerator reads them), item ids from list[4] on
            cudaMemsetAsync(d_list, 0, 16, st);
            s3_csr_scatter_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>((uint32_t)items, batchSize, numCases, d_starts, d_slot, a.csrL, a.csrR, a.csrInfo, d_list);
            S3_LAUNCHED(1);
            a.itemList = d_list + 4; a.itemCount = d_list; a.itemCap = (uint32_t)items;
            if ((rc = launch_search_csr<2>(ix, a, numCases))) { cudaFreeAsync(*d_out, st); *d_out = NULL; }
        }
    }
    cudaFreeAsync(d_slot, st); cudaFreeAsync(d_list, st);
    return rc;
}

extern "C" int s3_search(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t batchSize,
                         uint32_t wordPerQuery, uint32_t numMismatch, int isExactNumMismatch, s3_search_result *out)
{
    static const uint32_t ncases[5] = {1, 2, 4, 6, 10};
    if (!out) { s3_set_error("s3_search: NULL result"); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    int rc = check_search_args("s3_search", ix, queries, readLengths, batchSize, wordPerQuery, numMismatch,
                               numMismatch <= 4 ? ncases[numMismatch] : 0, 1, 2);
    if (rc) return rc;
    const uint32_t numCases = ncases[numMismatch];
    out->numReads = batchSize;
    out->offsets = (uint64_t *)calloc(batchSize + 1, sizeof(uint64_t));
    if (!out->offsets) { s3_set_error("s3_search: out of host memory"); return S3_ENOMEM; }
    if (batchSize == 0) return S3_OK;
    S3_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const size_t roundUp = (batchSize + 31) / 32 * 32, items = batchSize * numCases;
    const size_t qBytes = roundUp * wordPerQuery * 4, lBytes = roundUp * 4, cBytes = (items + 1) * 8;
    char *d;
    if ((rc = s3_scratch(ix, qBytes + lBytes + cBytes + 1024, (void **)&d))) { s3_search_result_free(out); return rc; }
    size_t off = 0;
    auto carve = [&](size_t bytes) { char *p = d + off; off += (bytes + 255) / 256 * 256; return p; };
    uint32_t *d_q = (uint32_t *)carve(qBytes), *d_l = (uint32_t *)carve(lBytes);
    unsigned long long *d_cnt = (unsigned long long *)carve(cBytes);
    S3_CUDA(cudaMemcpyAsync(d_q, queries, qBytes, cudaMemcpyHostToDevice, st));
    S3_CUDA(cudaMemsetAsync(d_l, 0, lBytes, st));
    S3_CUDA(cudaMemcpyAsync(d_l, readLengths, batchSize * 4, cudaMemcpyHostToDevice, st));
    uint32_t *d_out = NULL;
    unsigned long long total = 0;
    if ((rc = s3_search_csr_device(ix, d_q, d_l, (uint32_t)batchSize, wordPerQuery, numMismatch, isExactNumMismatch, d_cnt, &d_out, &total))) { s3_search_result_free(out); return rc; }
    // a read's run starts where its first case starts
    S3_CUDA(cudaMemcpy2DAsync(out->offsets, 8, d_cnt, (size_t)numCases * 8, 8, batchSize, cudaMemcpyDeviceToHost, st));
    out->total = total;
    if (total) {
        out->saL = (uint32_t *)malloc((size_t)total * 4); out->saR = (uint32_t *)malloc((size_t)total * 4); out->info = (uint32_t *)malloc((size_t)total * 4);
        if (!out->saL || !out->saR || !out->info) { s3_set_error("s3_search: out of host memory"); rc = S3_ENOMEM; }
        if (rc == S3_OK) {
            cudaError_t e = cudaMemcpyAsync(out->saL, d_out, (size_t)total * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out->saR, d_out + total, (size_t)total * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out->info, d_out + 2 * total, (size_t)total * 4, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) { s3_set_error("s3_search: copying the ranges back failed: %s", cudaGetErrorString(e)); rc = S3_ECUDA; }
        }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == S3_OK) { s3_set_error("s3_search: synchronize failed"); rc = S3_ECUDA; }
    out->offsets[batchSize] = total;
    if (d_out) cudaFreeAsync(d_out, st);
    if (rc != S3_OK) s3_search_result_free(out);
    return rc;
}

// ---- locate -------------------------------------------------------------------
// SA ranges -> text positions from the suffix array in HBM.  Replaces the host loops
// `for k in [l, r]: (*bwt->_bwtSaValue)(bwt, k)` that follow every search in the reference (SAList.cpp:411,
// CPUfunctions.cpp:2915,2975, PEAlgnmt.cpp:1246, DV-DPfunctions.cu:1186,2931).
__global__ void s3_locate_count_kernel(const uint32_t *__restrict__ saL, const uint32_t *__restrict__ saR, uint64_t n,
                                       uint32_t maxPerRange, unsigned long long *__restrict__ cnt)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n) return;
    unsigned long long c = 0;
    if (g < n && saR[g] >= saL[g]) { c = (unsigned long long)(saR[g] - saL[g]) + 1; if (c > maxPerRange) c = maxPerRange; }
    cnt[g] = c;                                                       // cnt[n] = 0: the scan leaves the total there
}

__global__ void s3_locate_fill_kernel(const uint32_t *__restrict__ sa, const uint32_t *__restrict__ saL, const uint32_t *__restrict__ saR,
                                      uint64_t n, uint32_t maxPerRange, const unsigned long long *__restrict__ start,
                                      uint32_t *__restrict__ out)
{
    // one warp per range: consecutive suffix array entries, consecutive outputs
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (g >= n || saR[g] < saL[g]) return;
    unsigned long long c = (unsigned long long)(saR[g] - saL[g]) + 1;
    if (c > maxPerRange) c = maxPerRange;
    for (unsigned long long k = lane; k < c; k += 32) out[start[g] + k] = sa[(size_t)saL[g] + k];
}

extern "C" void s3_free(void *p) { free(p); }

extern "C" int s3_locate(s3_index *ix, const uint32_t *saL, const uint32_t *saR, uint64_t numRanges, uint32_t maxPerRange,
                         uint64_t *offsets, uint32_t **positions, uint64_t *total)
{
    if (!ix || !offsets || !positions || !total || (numRanges && (!saL || !saR))) { s3_set_error("s3_locate: NULL argument"); return S3_EINVAL; }
    if (!ix->loc.sa) { s3_set_error("s3_locate: the index was uploaded without its suffix array"); return S3_EINVAL; }
    if (numRanges >= 0x7FFFFFFFull || maxPerRange == 0) { s3_set_error("s3_locate: numRanges / maxPerRange out of range"); return S3_EINVAL; }
    *positions = NULL; *total = 0; offsets[0] = 0;
    if (numRanges == 0) return S3_OK;
    for (uint64_t g = 0; g < numRanges; ++g)
        if (saR[g] >= saL[g] && saR[g] > ix->textLength) { s3_set_error("s3_locate: range %llu (%u..%u) lies outside the suffix array", (unsigned long long)g, saL[g], saR[g]); return S3_EINVAL; }
    S3_CUDA(cudaSetDevice(ix->device));
    size_t scanTemp = 0;
    cub::DeviceScan::ExclusiveSum(NULL, scanTemp, (unsigned long long *)NULL, (unsigned long long *)NULL, (int)(numRanges + 1), ix->stream);
    const size_t rBytes = numRanges * 4, cBytes = (numRanges + 1) * 8;
    char *d;
    int rc;
    if ((rc = s3_scratch(ix, 2 * rBytes + cBytes + scanTemp + 1024, (void **)&d))) return rc;
    size_t off = 0;
    auto carve = [&](size_t bytes) { char *p = d + off; off += (bytes + 255) / 256 * 256; return p; };
    uint32_t *d_l = (uint32_t *)carve(rBytes), *d_r = (uint32_t *)carve(rBytes);
    unsigned long long *d_cnt = (unsigned long long *)carve(cBytes);
    void *d_tmp = carve(scanTemp);
    S3_CUDA(cudaMemcpyAsync(d_l, saL, rBytes, cudaMemcpyHostToDevice, ix->stream));
    S3_CUDA(cudaMemcpyAsync(d_r, saR, rBytes, cudaMemcpyHostToDevice, ix->stream));
    s3_locate_count_kernel<<<(unsigned)((numRanges + 256) / 256), 256, 0, ix->stream>>>(d_l, d_r, numRanges, maxPerRange, d_cnt);
    S3_LAUNCHED(1);
    S3_CUDA(cudaGetLastError());
    S3_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, scanTemp, d_cnt, d_cnt, (int)(numRanges + 1), ix->stream));
    S3_CUDA(cudaMemcpyAsync(offsets, d_cnt, cBytes, cudaMemcpyDeviceToHost, ix->stream));
    S3_CUDA(cudaStreamSynchronize(ix->stream));
    const uint64_t tot = offsets[numRanges];
    *total = tot;
    if (tot == 0) return S3_OK;
    uint32_t *d_out = NULL;
    S3_CUDA(cudaMalloc(&d_out, tot * 4));
    s3_locate_fill_kernel<<<(unsigned)((numRanges * 32 + 255) / 256), 256, 0, ix->stream>>>(ix->loc.sa, d_l, d_r, numRanges, maxPerRange, d_cnt, d_out);
    S3_LAUNCHED(1);
    uint32_t *h = (uint32_t *)malloc(tot * 4);
    cudaError_t e = cudaGetLastError();
    if (!h) { cudaFree(d_out); s3_set_error("s3_locate: out of host memory"); return S3_ENOMEM; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_out, tot * 4, cudaMemcpyDeviceToHost, ix->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) { free(h); s3_set_error("s3_locate: %s", cudaGetErrorString(e)); return S3_ECUDA; }
    *positions = h;
    return S3_OK;
}

extern "C" int s3_search_round2(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths,
                                uint32_t *const *answers, uint64_t batchSize, uint64_t processedQuery,
                                uint32_t wordPerQuery, uint32_t numMismatch, uint32_t numCases,
                                uint32_t saRangeAllowed2, uint32_t wordPerAns, uint32_t wordPerAns2,
                                int isExactNumMismatch, uint32_t *const *badReadIndices,
                                uint32_t *const *badAnswers, uint64_t *numBad)
{
    int rc = check_search_args("s3_search_round2", ix, queries, readLengths, batchSize, wordPerQuery, numMismatch,
                               numCases, saRangeAllowed2, wordPerAns2);
    if (rc) return rc;
    if (!answers || !badReadIndices || !badAnswers || !numBad) { s3_set_error("s3_search_round2: NULL argument"); return S3_EINVAL; }
    if (processedQuery % 32 != 0) {
        // alignment.cu:258 indexes (processedQuery+readId)/32*32*W + readId%32, which is only a
        // consistent read address when the batch starts on a 32-read boundary (it always does)
        s3_set_error("s3_search_round2: processedQuery must be a multiple of 32");
        return S3_EINVAL;
    }
    for (uint32_t c = 0; c < numCases; ++c) numBad[c] = 0;
    if (batchSize == 0) return S3_OK;
    S3_CUDA(cudaSetDevice(ix->device));
    const uint32_t n = (uint32_t)batchSize;
    const size_t roundUp = (batchSize + 31) / 32 * 32;
    const size_t qBytes = roundUp * wordPerQuery * 4, lBytes = roundUp * 4;
    const size_t aBytes = roundUp * wordPerAns * 4, a2Bytes = roundUp * wordPerAns2 * 4;
    size_t selTemp = 0;
    cub::DeviceSelect::Flagged(NULL, selTemp, cub::CountingInputIterator<uint32_t>(0), (uint8_t *)NULL,
                               (uint32_t *)NULL, (uint32_t *)NULL, (int)n, ix->stream);
    selTemp = (selTemp + 255) / 256 * 256;
    const size_t total = qBytes + lBytes + aBytes + roundUp /*flags*/ + roundUp * 4 /*badIdx*/ + 256 /*count*/ +
                         qBytes + lBytes + a2Bytes + selTemp + 2048;
    char *d;
    if ((rc = s3_scratch(ix, total, (void **)&d))) return rc;
    size_t off = 0;
    auto carve = [&](size_t bytes) { char *p = d + off; off += (bytes + 255) / 256 * 256; return p; };
    uint32_t *d_q = (uint32_t *)carve(qBytes), *d_l = (uint32_t *)carve(lBytes), *d_a = (uint32_t *)carve(aBytes);
    uint8_t *d_flags = (uint8_t *)carve(roundUp);
    uint32_t *d_badIdx = (uint32_t *)carve(roundUp * 4), *d_cnt = (uint32_t *)carve(256);
    uint32_t *d_bq = (uint32_t *)carve(qBytes), *d_bl = (uint32_t *)carve(lBytes), *d_ba = (uint32_t *)carve(a2Bytes);
    void *d_tmp = carve(selTemp);
    const uint32_t *hq = queries + processedQuery * wordPerQuery;   // batch start, 32-aligned
    S3_CUDA(cudaMemcpyAsync(d_q, hq, qBytes, cudaMemcpyHostToDevice, ix->stream));
    S3_CUDA(cudaMemcpyAsync(d_l, readLengths + processedQuery, batchSize * 4, cudaMemcpyHostToDevice, ix->stream));
    for (uint32_t c = 0; c < numCases; ++c) {
        S3_CUDA(cudaMemcpyAsync(d_a, answers[c], aBytes, cudaMemcpyHostToDevice, ix->stream));
        s3_bad_flags_kernel<<<(n + 255) / 256, 256, 0, ix->stream>>>(d_a, n, wordPerAns, d_flags);
        S3_LAUNCHED(1);
        S3_CUDA(cudaGetLastError());
        S3_CUDA(cub::DeviceSelect::Flagged(d_tmp, selTemp, cub::CountingInputIterator<uint32_t>(0), d_flags, d_badIdx,
                                           d_cnt, (int)n, ix->stream));
        uint32_t nb = 0;
        S3_CUDA(cudaMemcpyAsync(&nb, d_cnt, 4, cudaMemcpyDeviceToHost, ix->stream));
        S3_CUDA(cudaStreamSynchronize(ix->stream));
        numBad[c] = nb;
        if (nb == 0) continue;
        const size_t nbUp = ((size_t)nb + 31) / 32 * 32;
        S3_CUDA(cudaMemsetAsync(d_bq, 0, nbUp * wordPerQuery * 4, ix->stream));
        s3_gather_bad_kernel<<<(nb + 255) / 256, 256, 0, ix->stream>>>(d_q, d_l, d_badIdx, nb, wordPerQuery, d_bq, d_bl);
        S3_LAUNCHED(1);
        S3_CUDA(cudaGetLastError());
        S3SearchArgs a;
        memset(&a, 0, sizeof a);
        a.queries = d_bq; a.readLengths = d_bl; a.numQueries = nb; a.wordPerQuery = wordPerQuery;
        a.answers[c] = d_ba;
        a.round = 1; a.numMismatch = numMismatch; a.saRangeAllowed = saRangeAllowed2; a.wordPerAnswer = wordPerAns2;
        a.firstCase = c; a.exactNum = isExactNumMismatch ? 1 : 0; a.textLength = ix->textLength;
        S3_CUDA(cudaMemsetAsync(d_ba, 0xFF, nbUp * wordPerAns2 * 4, ix->stream));
        if ((rc = launch_search(ix, a, 1, false))) return rc;
        S3_CUDA(cudaMemcpyAsync(badReadIndices[c], d_badIdx, (size_t)nb * 4, cudaMemcpyDeviceToHost, ix->stream));
        S3_CUDA(cudaMemcpyAsync(badAnswers[c], d_ba, nbUp * wordPerAns2 * 4, cudaMemcpyDeviceToHost, ix->stream));
        S3_CUDA(cudaStreamSynchronize(ix->stream));
    }
    return S3_OK;
}
