// s3_search.cu -- GPU-2BWT exact / <=4-mismatch search for sm_100a.
//
// Replaces the three search kernels of the reference (DV-Kernel.cu:4249 kernel,
// :4505 kernel_4mismatch_1, :4741 kernel_4mismatch_2) and their host drivers
// perform_round{1,2}_alignment (alignment.cu:118-531).
//
// Design (not a port): the reference spells each (case, mismatch set) as its
// own force-inlined recursion (51 blocks).  Here a case is a small *program* of
// phases (direction, read segment, min..max substitutions) interpreted by ONE
// data-driven loop whose body is "do one LF-mapping step on both interval ends".
// All lanes of a warp execute the same loop body whatever case / depth / strand
// they are in, so the only divergence left is the (rare) push/pop of a
// substitution frame.  The depth-first order of the reference -- substitutions
// in ascending symbol order *before* following the read's base, phases in
// program order -- is preserved, so the answer slots are bit-identical,
// including which ranges survive when a slot overflows.
#include "s3_common.cuh"
#include "../../include/soap3dp_b200.h"
#include <cub/device/device_select.cuh>
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

struct S3SearchArgs {
    const uint32_t *queries;
    const uint32_t *readLengths;
    uint32_t numQueries;
    uint32_t wordPerQuery;
    uint32_t *answers[S3_MAX_NUM_CASES];
    uint32_t round, numMismatch, saRangeAllowed, wordPerAnswer;
    uint32_t firstCase, exactNum;
    uint32_t textLength;
    unsigned long long *rankQueries;     // may be NULL
};

// DFS frame: a node where substitutions are still allowed.  a/b are the rank
// vectors of both interval ends at that node, from which every child interval
// (and the reverse interval update, DV-Kernel.cu:2673-2690) is derived.
struct S3Frame {
    uint32_t a[4], b[4];
    uint32_t r, rr;          // the node's own saR / revSaR
    uint32_t meta;           // done[0:11] p[11:13] mmp[13:16] mmt[16:19] next[19:22] c[22:24]
};

__device__ __forceinline__ uint32_t s3_meta(uint32_t done, uint32_t p, uint32_t mmp, uint32_t mmt, uint32_t next, uint32_t c)
{
    return done | (p << 11) | (mmp << 13) | (mmt << 16) | (next << 19) | (c << 22);
}

// smem read: word w of this thread's read lives at sm[w * S3_THREADS + tid]
__device__ __forceinline__ uint32_t s3_base(const uint32_t *sm, uint32_t pos, uint32_t L, uint32_t strand)
{
    // strand 1 = reverse complement (what the reference materialises in place, DV-Kernel.cu:4351-4395)
    const uint32_t i = strand ? (L - 1 - pos) : pos;
    const uint32_t v = (sm[(i >> 4) * S3_THREADS] >> ((i & 15) << 1)) & 3;
    return strand ? 3 - v : v;
}

template <bool COUNT>
__global__ void __launch_bounds__(S3_THREADS)
s3_search_kernel(const S3Half fwd, const S3Half rev, const S3SearchArgs args)
{
    extern __shared__ uint32_t s3_smem[];
    const uint32_t q = blockIdx.x * S3_THREADS + threadIdx.x;
    const uint32_t whichCase = args.firstCase + blockIdx.y;
    unsigned long long nrank = 0;
    if (q < args.numQueries) {
        const uint32_t lane32 = q & 31;
        const uint32_t *query = args.queries + (size_t)(q >> 5) * 32 * args.wordPerQuery + lane32;
        uint32_t *answer = args.answers[whichCase] + (size_t)(q >> 5) * 32 * args.wordPerAnswer + lane32;
        const uint32_t L = args.readLengths[q];
        uint32_t *sm = s3_smem + threadIdx.x;
        const uint32_t nw = (L + 15) >> 4;
        for (uint32_t w = 0; w < nw; ++w) sm[w * S3_THREADS] = query[w * 32];
        for (uint32_t i = 0; i < args.wordPerAnswer; ++i) answer[i * 32] = 0xFFFFFFFFu;

        S3Phase ph[S3_MAX_PHASES];
        uint32_t firstL;
        const int nph = s3_case_program(args.numMismatch, whichCase, L, args.exactNum != 0, ph, firstL);
        const uint32_t maxRanges = args.saRangeAllowed;
        uint32_t saCount = 0;
        // round 1: the device read buffer of the reference flips orientation after every
        // launch, so odd cases meet the reverse strand first (DV-Kernel.cu:4280-4285)
        uint32_t strand = args.round > 0 ? 0u : (whichCase & 1u);
        S3Frame frames[S3_MAX_DEPTH];

        for (int pass = 0; pass < 2 && nph > 0; ++pass, strand ^= 1u) {
            int depth = 0;
            uint32_t p = 0, done = 0, mmp = 0, mmt = 0;
            uint32_t l = firstL, r = args.textLength, rl = 0, rr = args.textLength;
            bool alive = true;
            while (true) {
                if (saCount > maxRanges) break;
                if (alive && done == ph[p].len) {
                    // end of a phase
                    if (mmp < ph[p].lo) alive = false;
                    else if ((int)p + 1 == nph) {
                        // report (DV-Kernel.cu:355-380)
                        if (saCount < maxRanges) {
                            answer[32 * 2 * saCount] = l;
                            answer[32 * (2 * saCount + 1)] = (r - l) + (strand << 27) + (mmt << 24);
                        }
                        ++saCount;
                        alive = false;
                    } else { ++p; done = 0; mmp = 0; continue; }
                }
                if (alive) {
                    const uint32_t dir = ph[p].dir;
                    const uint32_t pos = dir ? ph[p].start + done : ph[p].start + ph[p].len - 1 - done;
                    const uint32_t c = s3_base(sm, pos, L, strand);
                    uint32_t a[4], b[4];
                    if (dir) { s3_rank4(rev, rl, a); s3_rank4(rev, rr + 1, b); }
                    else     { s3_rank4(fwd, l, a);  s3_rank4(fwd, r + 1, b); }
                    if (COUNT) nrank += 2;
                    if (mmp < ph[p].hi) {
                        // does any substitution child survive?  (most do not once the interval is narrow)
                        uint32_t live = 0;
#pragma unroll
                        for (uint32_t e = 0; e < 4; ++e) live |= (e != c && a[e] + 1 <= b[e]) ? (1u << e) : 0u;
                        if (live) {
                            S3Frame &f = frames[depth++];
#pragma unroll
                            for (int e = 0; e < 4; ++e) { f.a[e] = a[e]; f.b[e] = b[e]; }
                            f.r = r; f.rr = rr;
                            f.meta = s3_meta(done, p, mmp, mmt, 0, c);
                            alive = false;     // children are taken from the frame below
                        }
                    }
                    if (alive) {
                        // follow the read's base
                        const uint32_t cum = (c < 3 ? b[3] - a[3] : 0) + (c < 2 ? b[2] - a[2] : 0) + (c < 1 ? b[1] - a[1] : 0);
                        if (dir) { rl = a[c] + 1; rr = b[c]; r = r - cum; l = r - (rr - rl); }
                        else     { l = a[c] + 1; r = b[c]; rr = rr - cum; rl = rr - (r - l); }
                        ++done;
                        alive = (l <= r);
                    }
                }
                if (!alive) {
                    // take the next pending branch from the innermost frame
                    if (depth == 0) break;
                    S3Frame &f = frames[depth - 1];
                    const uint32_t meta = f.meta;
                    const uint32_t c = (meta >> 22) & 3;
                    uint32_t e = (meta >> 19) & 7;
                    p = (meta >> 11) & 3; done = meta & 0x7FF; mmp = (meta >> 13) & 7; mmt = (meta >> 16) & 7;
                    const uint32_t dir = ph[p].dir;
                    // next substitution symbol with a non-empty interval, ascending
                    while (e < 4 && (e == c || f.a[e] + 1 > f.b[e])) ++e;
                    uint32_t sym;
                    if (e < 4) { sym = e; f.meta = (meta & ~(7u << 19)) | ((e + 1) << 19); ++mmp; ++mmt; }
                    else { sym = c; --depth; }             // finally the read's own base; frame retired
                    uint32_t cum = 0;
                    for (uint32_t j = 3; j > sym; --j) cum += f.b[j] - f.a[j];
                    const uint32_t nlo = f.a[sym] + 1, nhi = f.b[sym];
                    if (dir) { rl = nlo; rr = nhi; r = f.r - cum; l = r - (rr - rl); }
                    else     { l = nlo; r = nhi; rr = f.rr - cum; rl = rr - (r - l); }
                    ++done;
                    alive = (l <= r);
                }
            }
            if (saCount > maxRanges) break;
        }
        // status word (DV-Kernel.cu:4468-4491); the isBad carry between the cases of
        // round 1 is applied by s3_isbad_fixup_kernel because cases run concurrently here
        if (saCount == 0) answer[0] = 0xFFFFFFFDu;
        else if (saCount > maxRanges) answer[0] = 0xFFFFFFFEu;
    }
    if (COUNT) {
        // warp-aggregate then one atomic per warp
        for (int o = 16; o > 0; o >>= 1) nrank += __shfl_down_sync(0xFFFFFFFFu, nrank, o);
        if ((threadIdx.x & 31) == 0 && nrank) atomicAdd(args.rankQueries, nrank);
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

static int launch_search(s3_index *ix, S3SearchArgs &a, uint32_t numCases, bool count)
{
    if (a.numQueries == 0) return S3_OK;
    dim3 grid((a.numQueries + S3_THREADS - 1) / S3_THREADS, numCases);
    size_t smem = (size_t)a.wordPerQuery * S3_THREADS * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        S3_CUDA(cudaFuncSetAttribute(s3_search_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        S3_CUDA(cudaFuncSetAttribute(s3_search_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (count) s3_search_kernel<true><<<grid, S3_THREADS, smem, ix->stream>>>(ix->fwd, ix->rev, a);
    else s3_search_kernel<false><<<grid, S3_THREADS, smem, ix->stream>>>(ix->fwd, ix->rev, a);
    S3_LAUNCHED(1);
    S3_CUDA(cudaGetLastError());
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
    if (batchSize > 0xFFFFFFFFull) { s3_set_error("%s: batch too large", fn); return S3_EINVAL; }
    return S3_OK;
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
    S3SearchArgs a;
    memset(&a, 0, sizeof a);
    a.queries = d_queries; a.readLengths = d_readLengths; a.numQueries = (uint32_t)batchSize;
    a.wordPerQuery = wordPerQuery;
    for (uint32_t c = 0; c < numCases; ++c) a.answers[c] = d_answers[c];
    a.round = 0; a.numMismatch = numMismatch; a.saRangeAllowed = saRangeAllowed; a.wordPerAnswer = wordPerAns;
    a.firstCase = 0; a.exactNum = isExactNumMismatch ? 1 : 0; a.textLength = ix->textLength;
    a.rankQueries = d_rankQueries;
    if ((rc = launch_search(ix, a, numCases, d_rankQueries != NULL))) return rc;
    if (numCases > 1 && batchSize > 0) {
        s3_isbad_fixup_kernel<<<(unsigned)((batchSize + 255) / 256), 256, 0, ix->stream>>>(a, numCases);
        S3_LAUNCHED(1);
        S3_CUDA(cudaGetLastError());
    }
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
    const size_t roundUp = (batchSize + 31) / 32 * 32;
    const size_t qBytes = roundUp * wordPerQuery * 4, lBytes = roundUp * 4, aBytes = roundUp * wordPerAns * 4;
    char *d;
    if ((rc = s3_scratch(ix, qBytes + lBytes + aBytes * numCases + 256, (void **)&d))) return rc;
    uint32_t *d_q = (uint32_t *)d, *d_l = (uint32_t *)(d + qBytes);
    uint32_t *d_ans[S3_MAX_NUM_CASES];
    for (uint32_t c = 0; c < numCases; ++c) d_ans[c] = (uint32_t *)(d + qBytes + lBytes + aBytes * c);
    S3_CUDA(cudaMemcpyAsync(d_q, queries, qBytes, cudaMemcpyHostToDevice, ix->stream));
    // the reference copies roundUp lengths (alignment.cu:161); only batchSize are meaningful
    S3_CUDA(cudaMemcpyAsync(d_l, readLengths, batchSize * 4, cudaMemcpyHostToDevice, ix->stream));
    if ((rc = s3_search_round1_device(ix, d_q, d_l, batchSize, wordPerQuery, numMismatch, numCases, saRangeAllowed,
                                      wordPerAns, isExactNumMismatch, d_ans, NULL))) return rc;
    for (uint32_t c = 0; c < numCases; ++c)
        S3_CUDA(cudaMemcpyAsync(answers[c], d_ans[c], aBytes, cudaMemcpyDeviceToHost, ix->stream));
    S3_CUDA(cudaStreamSynchronize(ix->stream));
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
