// s3_dp.cu -- semi-global affine-gap DP (score pass + traceback) for sm_100a.
//
// Replaces SemiGlobalAligntment / GPUBacktrack (DV-DPfunctions.cu:243,316) and
// SemiGlobalAligner (DV-DPfunctions.cu:520-741), scheme-1 (full table) semantics.
//
// Design (not a port).  The reference runs one thread per alignment and keeps
// the full H and E tables of shorts in global memory (8 B of traffic per cell,
// DV-DPfunctions.cu:57,146-241), then walks them again to trace back.  Here:
//   * one WARP per PAIR of alignments; the two alignments ride in the two 16-bit
//     halves of every register and all arithmetic is DPX 16x2 (VIADD.16x2,
//     VIADDMNMX.S16x2, VIMNMX3.S16x2): the reference's own value range is that of a
//     short clamped at -32000 (DV-DPfunctions.cu:45-48), so nothing is lost.
//   * lane t owns read rows [t*R+1, t*R+R] and sweeps the reference columns as an
//     anti-diagonal wavefront: at step s lane t is in column s-t.  The reference's
//     inner-loop carried registers (upScore, scoreOpenUp, prevScoreUp) simply
//     migrate from lane t-1 to lane t with __shfl_up_sync, so every intermediate
//     value -- including the "unclamped inside a column, clamped to -32000 when
//     stored" rule -- is bit-identical.
//   * H/E of the previous column live in registers (2R per lane).  Per cell only the
//     clamped H (2 bytes) goes to memory, anti-diagonal major so that a warp's store is
//     one contiguous 512-byte line.  The one thing GPUBacktrack reads from the E table
//     ("H == ext + E of the previous column", where an indel starts) is rebuilt by the
//     traceback from the H row, since E is a function of the H values to its left.
//   * best cell: per-lane running best in the reference's scan order, merged
//     across lanes with the (column, row) tie-break; tie counts are summed.
// The traceback kernel is one thread per alignment that reached its cutoff and is
// GPUBacktrack's state machine (DV-DPfunctions.cu:330-508) over the stored H values.
// A 32-bit one-alignment-per-warp variant (byte-plane traceback) remains for read
// lengths above 256 or score parameters that do not fit 16-bit lanes.
#include "s3_common.cuh"
#include "../../include/soap3dp_b200.h"
#include <stdlib.h>
#include <string.h>

#define S3_NEG_INF (-32000)
#define TB_DIAG 0
#define TB_DOPEN 1
#define TB_DEXT 2
#define TB_SMEXIT 3
#define TB_SIEXIT 4
#define TB_IOPEN 5
#define TB_IEXT 6
#define TB_MATCH 8
#define TB_EOPEN 16
#define TB_FOPEN 32
#define TB_FEXIT 64

#define S3_DP_WARPS 4
#define S3_DP_MAX_REF_WORDS 160      // reference window up to 2559 bases

struct s3_dp {
    int device;
    cudaStream_t stream;
    int ownStream;
    uint32_t maxReadLength, maxDNALength, maxBatch;
    s3_dp_scores sc;
    int R;                       // rows per lane
    int narrow;                  // 1: 16x2 kernel (H plane), 0: 32-bit kernels (byte plane)
    int lanes;                   // narrow: lanes that sweep one pair of alignments (16 or 32)
    uint32_t slot;               // wide path: bytes per lane per column in the traceback plane
    uint32_t chunk;              // alignments per traceback-plane chunk
    uint8_t *d_tb;               // wide: chunk * (maxDNALength+1) * 32 * slot bytes
                                 // narrow: chunk/2 pairs * (maxDNALength+lanes+1) steps * lanes * PW words (a pass-2 plane each)
    uint32_t *d_scRight;         // maxBatch
    // narrow path: checkpoints of the sweep (per pair of a chunk), the traceback lists of a chunk (3 x chunk ids:
    // resumed, from column 0, second pass) with their counters, first step of every alignment's second sweep
    uint32_t *d_ckpt, *d_tbList, *d_tbCount, *d_s0;
    uint32_t numCk;
    // staging for the host entry point
    uint32_t *d_dna, *d_read, *d_dnaLen, *d_readLen, *d_hit, *d_cnt, *d_clipLt, *d_clipRt, *d_ancL, *d_ancR;
    int32_t *d_cutoff, *d_score;
    uint8_t *d_pattern;
    unsigned long long *d_cells;
    S3Pipe pipe;
    S3Timing timing;             // slots: 0 score, 1 best cell, 2 traceback
    void *d_stage; size_t stageBytes;    // s3_dp_align_windows: query buffer and window descriptors
};

struct S3DpArgs {
    const uint32_t *dna, *dnaLen, *read, *readLen;
    const uint32_t *clipLt, *clipRt, *ancL, *ancR;
    const int32_t *cutoff;
    int32_t *score;
    uint32_t *hit, *cnt, *scRight;
    uint8_t *tb, *pattern;
    uint32_t first, count;           // alignment range of this chunk
    uint32_t maxReadLength, maxDNALength, dnaWords, readWords;
    uint32_t slot;
    int match, mismatch, open, ext;
    unsigned long long *cells;
    uint32_t *hplane;                // narrow path: H values of a traceback window, [pair][step][lane][PW] words (A low half, B high half)
    uint32_t planeSteps;             // steps per pair in it: pass 1 maxReadLength + slack + checkpoint distance + lanes + 1, pass 2 maxDNALength + lanes + 1
    uint32_t *ckpt; size_t ckStride; uint32_t numCk;      // checkpoints: words per pair of the sweep, checkpoints per pair
    uint32_t *tbList, *tbCount, tbCap, *s0;               // traceback lists (3 x tbCap), their counters, first step of the second sweep
    uint32_t pass;                   // 1: windows, 2: whole tables of what pass 1 could not trace
    // narrow path.  Values travel as v + 32768 in each 16-bit half ("biased"), so that a score parameter is
    // added to both halves by ONE 32-bit add of kX = x * 0x10001 (two's complement for x < 0: no carry or
    // borrow crosses the halves because 0 < biased value +- parameter < 65536) -- an IMAD, which leaves the
    // ALU pipe to the 16x2 max instructions.
    uint32_t kOpen, kExt, kGapInit;
    uint32_t ceh2;                   // biased max(-32000 + open, -32000 + ext): the clamp of H and E, moved into E's max
    uint32_t nego2;                  // biased -32000 + open: the clamp of the diagonal H, in the "+ open" domain
    uint32_t one;                    // 1, opaque to the compiler (mad.lo by it keeps the adds on the FMA pipe)
    uint32_t mism4, delta;           // (mismatch - open) in all four bytes; ((match - open) ^ (mismatch - open)) & 0xFF
    uint32_t colStride;              // narrow path: uint2 entries of one pair's column table (maxDNALength + 1)
};

__device__ __forceinline__ int s3_clamp(int x) { return max(x, S3_NEG_INF); }

template <int R>
__global__ void __launch_bounds__(S3_DP_WARPS * 32)
s3_dp_score_kernel(const S3DpArgs a)
{
    __shared__ uint32_t refWords[S3_DP_WARPS][S3_DP_MAX_REF_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t local = blockIdx.x * S3_DP_WARPS + warp;
    if (local >= a.count) return;                       // whole warp leaves together
    const uint32_t id = a.first + local;
    const uint32_t g = id >> 5, gl = id & 31;
    // lengths past what the packed arrays hold (base i lives in word i >> 4) are cut: never read beyond a slot
    const uint32_t m = min(a.readLen[id], a.readWords * 16u - 1u), n = min(a.dnaLen[id], a.dnaWords * 16u - 1u);
    const uint32_t clipLt = a.clipLt ? a.clipLt[id] : 0u;
    const uint32_t clipRt = a.clipRt ? a.clipRt[id] : 0u;
    const uint32_t anchorLeft = a.ancL ? a.ancL[id] : a.maxDNALength;
    const uint32_t anchorRight = a.ancR ? a.ancR[id] : 0u;
    const int open = a.open, ext = a.ext, gapInit = a.open - a.ext;
    const int clipRtCheck = (int)(m - clipRt);

    // reference window -> shared memory (1-based packing, MSB first; DV-DPfunctions.cu:58)
    const uint32_t *dna = a.dna + (size_t)g * a.dnaWords * 32 + gl;
    const uint32_t nRefWords = (n >> 4) + 1;
    for (uint32_t w = lane; w < nRefWords; w += 32) refWords[warp][w] = dna[(size_t)w * 32];
    // this lane's read bases
    const uint32_t *read = a.read + (size_t)g * a.readWords * 32 + gl;
    const uint32_t i0 = lane * R + 1;                  // first row of this lane (1-based)
    uint32_t rc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t i = i0 + r;
        rc[r] = (i <= m) ? (read[(size_t)(i >> 4) * 32] >> ((15 - (i & 15)) << 1)) & 3 : 4u;   // 4 never matches
    }
    __syncwarp();

    // column 0 (DV-DPfunctions.cu:167-184)
    int Hp[R], Ep[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t i = i0 + r;
        const int h = (i <= clipLt) ? open : gapInit + (int)(i - clipLt) * ext;
        Hp[r] = s3_clamp(h);
        Ep[r] = s3_clamp(h + gapInit);
    }
    int best = S3_NEG_INF;
    uint32_t hitJ = 0, bestI = 0, count = 0;
    int upOut = 0, FOut = 0, diagRawOut = 0;
    uint8_t *tb = a.tb + (size_t)local * (a.maxDNALength + 1) * 32 * a.slot + (size_t)lane * a.slot;
    const uint32_t colStride = 32 * a.slot;

    const uint32_t steps = n + 31;
    for (uint32_t s = 1; s <= steps; ++s) {
        // carried registers of the row loop arrive from the lane above (column j was done there at step s-1)
        int up = __shfl_up_sync(0xFFFFFFFFu, upOut, 1);
        int F = __shfl_up_sync(0xFFFFFFFFu, FOut, 1);
        int diagRaw = __shfl_up_sync(0xFFFFFFFFu, diagRawOut, 1);
        const uint32_t j = s - lane;
        if (j >= 1 && j <= n && i0 <= m) {
            const int init = (j >= anchorLeft) ? S3_NEG_INF : 0;
            const int prevInit = (j >= 2 && j - 1 >= anchorLeft) ? S3_NEG_INF : 0;
            // GPUBacktrack's own idea of the previous column's start value (DV-DPfunctions.cu:352-357,368);
            // differs from prevInit only for j == 1 with anchorLeft == 0
            const int prevInitTb = (j > anchorLeft) ? S3_NEG_INF : 0;
            int diag;
            if (lane == 0) { up = init; F = init + gapInit; diagRaw = prevInit; diag = prevInit; }
            else diag = (i0 - 1 <= clipLt) ? max(prevInit, diagRaw) : diagRaw;
            int upSt = s3_clamp(up);
            const uint32_t refChar = (refWords[warp][j >> 4] >> ((15 - (j & 15)) << 1)) & 3;
            const bool jOk = j >= anchorRight;
            uint32_t packed[(R + 3) / 4];
#pragma unroll
            for (int k = 0; k < (R + 3) / 4; ++k) packed[k] = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t i = i0 + r;
                const bool isMatch = (refChar == rc[r]);
                const int d = isMatch ? a.match : a.mismatch;
                const int left = Hp[r], eL = Ep[r];
                const int e = __viaddmax_s32(left, open, eL + ext);
                F = __viaddmax_s32(up, open, F + ext);
                up = __vimax3_s32(F, e, diag + d);
                const int hSt = s3_clamp(up), eSt = s3_clamp(e);
                const bool nearClip = i <= clipLt + 1;
                uint32_t b = (hSt == d + diagRaw) ? TB_DIAG
                           : (hSt == open + left) ? TB_DOPEN
                           : (hSt == ext + eL) ? TB_DEXT
                           : (nearClip && hSt == prevInitTb + d) ? TB_SMEXIT
                           : (nearClip && hSt == init + open) ? TB_SIEXIT
                           : (hSt == open + upSt) ? TB_IOPEN : TB_IEXT;
                b |= isMatch ? TB_MATCH : 0;
                b |= (eSt == open + left) ? TB_EOPEN : 0;
                diagRaw = left; diag = left;
                if (i <= clipLt) { F = max(init + gapInit, F); diag = max(prevInit, diag); }
                b |= (F == open + upSt) ? TB_FOPEN : 0;
                b |= (nearClip && F == init + open) ? TB_FEXIT : 0;
                packed[r >> 2] |= b << ((r & 3) * 8);
                Hp[r] = hSt; Ep[r] = eSt; upSt = hSt;
                if ((int)i >= clipRtCheck && i <= m && jOk) {
                    if (up > best) { best = up; hitJ = j; bestI = i; count = 1; }
                    else if (up == best) ++count;
                }
            }
            upOut = up; FOut = F; diagRawOut = diagRaw;
            uint8_t *dst = tb + (size_t)j * colStride;
            if (R == 4) *reinterpret_cast<uint32_t *>(dst) = packed[0];
            else if (R == 8) *reinterpret_cast<uint2 *>(dst) = make_uint2(packed[0], packed[(R > 4) ? 1 : 0]);
            else {
#pragma unroll
                for (int k = 0; k < (R + 3) / 4; ++k) reinterpret_cast<uint32_t *>(dst)[k] = packed[k];
            }
        }
    }
    // merge the lanes' bests: highest score, then first in (column, row) scan order
    int gbest = best;
    for (int o = 16; o > 0; o >>= 1) gbest = max(gbest, __shfl_xor_sync(0xFFFFFFFFu, gbest, o));
    unsigned long long key = (best == gbest && count > 0) ? (((unsigned long long)hitJ << 32) | bestI) : ~0ull;
    uint32_t cnt = (best == gbest) ? count : 0u;
    for (int o = 16; o > 0; o >>= 1) {
        key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, o));
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    }
    if (lane == 0) {
        const bool any = key != ~0ull;
        a.score[id] = gbest;
        a.hit[id] = any ? (uint32_t)(key >> 32) : 0u;
        const uint32_t bi = (uint32_t)(key & 0xFFFFFFFFu);
        a.scRight[id] = (any && bi != 0) ? m - bi : 0u;   // bi == 0: only ties with the -32000 start value
        a.cnt[id] = cnt;
        if (a.cells) atomicAdd(a.cells, (unsigned long long)m * n);
    }
}

// One thread per alignment: GPUBacktrack's state machine (DV-DPfunctions.cu:330-508)
// driven by the byte plane.
__global__ void s3_dp_traceback_kernel(const S3DpArgs a, int R)
{
    const uint32_t local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= a.count) return;
    const uint32_t id = a.first + local;
    if (a.score[id] < a.cutoff[id]) return;
    const uint32_t m = min(a.readLen[id], a.readWords * 16u - 1u);
    const uint32_t clipLt = a.clipLt ? a.clipLt[id] : 0u;
    const uint32_t scRight = a.scRight[id];
    const uint8_t *tb = a.tb + (size_t)local * (a.maxDNALength + 1) * 32 * a.slot;
    const uint32_t colStride = 32 * a.slot;
    uint8_t *pat = a.pattern + (size_t)id * (a.maxReadLength + a.maxDNALength);
    uint32_t p = 0;
    if (scRight > 0) { pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)scRight; }
    uint32_t readPos = m - scRight, refIndex = a.hit[id];
    int state = 0;      // 0 NORMAL, 1 I_EXT, 2 D_EXT, 3 SM_EXIT, 4 SI_EXIT
    uint32_t lastCell = 0;
    while (readPos > 0 && refIndex > 0) {
        const uint32_t row = readPos - 1;
        const uint32_t b = tb[(size_t)refIndex * colStride + (row / R) * a.slot + (row % R)];
        lastCell = b;
        if (state == 0) {
            const uint32_t ch = b & 7;
            if (ch == TB_DIAG) { pat[p++] = (b & TB_MATCH) ? 'M' : 'm'; --refIndex; --readPos; }
            else if (ch == TB_DOPEN) { pat[p++] = 'D'; --refIndex; }
            else if (ch == TB_DEXT) { pat[p++] = 'D'; --refIndex; state = 2; }
            else if (ch == TB_SMEXIT) { state = 3; break; }
            else if (ch == TB_SIEXIT) { state = 4; break; }
            else if (ch == TB_IOPEN) { pat[p++] = 'I'; --readPos; }
            else { pat[p++] = 'I'; --readPos; state = 1; }
        } else if (state == 2) {
            pat[p++] = 'D'; --refIndex;
            if (b & TB_EOPEN) state = 0;
        } else {
            if (b & TB_FEXIT) { state = 4; break; }
            pat[p++] = 'I'; --readPos;
            if (b & TB_FOPEN) state = 0;
        }
    }
    if (refIndex == 0) {
        const uint32_t scNum = min(clipLt, readPos);
        if (scNum < readPos) { pat[p++] = 'I'; pat[p++] = 'V'; pat[p++] = (uint8_t)(readPos - scNum); }
        pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)scNum;
    } else if (state == 4) {
        pat[p++] = 'I'; pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)(readPos - 1);
    } else if (state == 3) {
        pat[p++] = (lastCell & TB_MATCH) ? 'M' : 'm';
        pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)(readPos - 1);
        refIndex -= 1;
    }
    pat[p++] = 0;
    a.hit[id] = refIndex;        // start offset inside the window (refOffset == 0 in scheme 1)
}

// =============================================================================
// 16x2 path: two alignments per warp, DPX instructions, H plane + flag plane
// =============================================================================
#define S3_NEG2 0x83008300u          // -32000 in both halves
#define S3_MIN2 0x80008000u          // -32768 in both halves (identity of max)
#define S3_MAX2 0x7FFF7FFFu          // +32767 in both halves (identity of min)

__device__ __forceinline__ uint32_t s3_pk(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ int s3_lo16(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
__device__ __forceinline__ int s3_hi16(uint32_t v) { return (int)(short)(v >> 16); }

// prmt.b32 in its generic form: selector nibble bit 3 replicates the sign of the selected byte
// (__byte_perm masks that bit away)
__device__ __forceinline__ uint32_t s3_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

#define S3_BIAS2 0x80008000u         // the H plane and the score kernel's registers hold v + 32768 per half
#define S3_NEGB2 0x03000300u         // -32000, biased

__device__ __forceinline__ uint32_t s3_fadd(uint32_t x, uint32_t k, uint32_t one)
{
#ifdef S3_DP_FORCE_IMAD
    uint32_t d;                                   // x + k as IMAD (FMA pipe); `one` is 1 at run time
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(k));
    return d;
#else
    (void)one;
    return x + k;                                 // ptxas spreads plain adds over the ALU (IADD3) and FMA (IMAD.IADD) pipes itself
#endif
}
__device__ __forceinline__ uint32_t s3_bpk(int lo, int hi) { return ((uint32_t)(lo + 32768) & 0xFFFFu) | ((uint32_t)(hi + 32768) << 16); }

// ---- the 16x2 path keeps no H plane of the whole table ------------------------------------------------
// (1) s3_dp_sweep16_kernel   sweeps every column once: best cell, tie count, right clip (DV-DPfunctions.cu:225-235)
//     are found inside the sweep; every S3_DP_CK steps each lane leaves a checkpoint of its registers (64 B).
// (2) s3_dp_resweep16_kernel sweeps again, for the alignments that reached their cutoff, only the steps a
//     traceback from the best cell can reach: from the last checkpoint at least readLength + S3_DP_SLACK columns left
//     of the end column, into a plane of that window only.
// (3) s3_dp_traceback16_kernel walks that plane.  A traceback that needs a cell left of its window (more than
//     S3_DP_SLACK deleted bases) puts its alignment on a list; pass 2 re-sweeps those from column 0 into a full-size
//     plane and traces them again.  Results are those of the full table: a resumed sweep continues from the very
//     registers of the first one (the scheme is stated on the CPU by dp_one_resweep of the DP restatement under tests).
#ifndef S3_DP_CK
#define S3_DP_CK 32u                 // steps between two checkpoints (a power of two, >= the lanes of a group)
#endif
#ifndef S3_DP_SLACK
#define S3_DP_SLACK 16u              // columns kept left of (end column - read length)
#endif

template <int R> struct S3DpGeom {
    static constexpr int PW = (R <= 4) ? 4 : (R <= 8) ? 8 : 16; // words per lane and step in a traceback plane: whole 16- / 32- / 64-byte pieces
    static constexpr int CKW = (2 * R + 2 + 7) / 8 * 8;         // words per lane and checkpoint: H + open [R], E [R], F, diagonal
};

// The R rows of one lane at one column: 5 instructions on the ALU pipe per row (PRMT, 3 x VIMNMX3.U16x2,
// VIMNMX.U16x2) and 4 adds the compiler spreads over the ALU and FMA pipes.  HO[r] leaves as H + open of this column.
// (A form whose row-to-row chain is one add and one maximum -- F(i+1) = max(F(i) + ext, X(i) + open, clip) with
// X = max(E, diagonal + score), valid for ext >= open -- costs two more instructions per row and was measured no
// faster: the sweep is bound by issue slots, not by that chain.  Neither was a two-pass form -- E of every row first, which needs
// nothing from the lane above, then the row-to-row chain with the next row's diagonal taken before H is overwritten: 1.074
// against 1.041 ms per 65,536 alignments.)
template <int R>
__device__ __forceinline__ void s3_dp_rows(const uint2 tab, const uint32_t (&sel)[R], uint32_t (&HO)[R], uint32_t (&E)[R],
                                           const uint32_t (&clipIO)[R], const uint32_t (&clipPIO)[R],
                                           uint32_t &upO, uint32_t &F, uint32_t &diagO,
                                           const uint32_t KOPEN, const uint32_t KEXT, const uint32_t CEH2, const uint32_t ONE)
{
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t d = s3_prmt(tab.x, tab.y, sel[r]);                       // substitution score - open, >= 0
        const uint32_t e = __vimax3_u16x2(HO[r], s3_fadd(E[r], KEXT, ONE), CEH2);
        F = __vimax3_u16x2(s3_fadd(F, KEXT, ONE), upO, clipIO[r]);
        const uint32_t dg = __vmaxu2(diagO, clipPIO[r]);
        const uint32_t up = __vimax3_u16x2(F, e, s3_fadd(dg, d, ONE));
        diagO = HO[r];
        upO = s3_fadd(up, KOPEN, ONE);
        HO[r] = upO; E[r] = e;
    }
}

// soft-clip restart operands of a column (DV-DPfunctions.cu:215-219): row i is fed from row i-1 when i-1 <= clipLt
template <int R>
__device__ __forceinline__ void s3_dp_clip_operands(const uint32_t init, const uint32_t prevInit, const uint32_t i0, const uint32_t (&clipLt)[2],
                                                    const uint32_t KOPEN, const uint32_t NEGO2, uint32_t (&clipIO)[R], uint32_t (&clipPIO)[R])
{
    const uint32_t io = init + KOPEN, pio = prevInit + KOPEN;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t i = i0 + r;
        const uint32_t cm = ((i - 1 <= clipLt[0]) ? 0xFFFFu : 0u) | ((i - 1 <= clipLt[1]) ? 0xFFFF0000u : 0u);
        clipIO[r] = io & cm;                            // biased 0 = -32768: the identity of max elsewhere
        clipPIO[r] = (pio & cm) | (NEGO2 & ~cm);
    }
}

// start value of column j of both alignments (biased): 0 left of the left anchor, -32000 from it on
__device__ __forceinline__ uint32_t s3_dp_init2(const uint32_t jA, const uint32_t jB, const uint32_t (&ancL)[2])
{
    return ((jA >= ancL[0]) ? (S3_NEGB2 & 0xFFFFu) : 0x8000u) | ((jB >= ancL[1]) ? (S3_NEGB2 & 0xFFFF0000u) : 0x80000000u);
}

// the two alignments of a pair: what the kernels need of them
struct S3DpPair {
    uint32_t id[2], m[2], n[2], clipLt[2], clipRt[2], ancL[2], ancR[2];
};
__device__ __forceinline__ void s3_dp_load_pair(const S3DpArgs &a, S3DpPair &p)
{
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        // lengths past what the packed arrays hold (base i lives in word i >> 4) are cut: never read beyond a slot
        p.m[x] = min(a.readLen[p.id[x]], a.readWords * 16u - 1u);
        p.n[x] = min(a.dnaLen[p.id[x]], a.dnaWords * 16u - 1u);
        p.clipLt[x] = a.clipLt ? a.clipLt[p.id[x]] : 0u;
        p.clipRt[x] = a.clipRt ? a.clipRt[p.id[x]] : 0u;
        p.ancL[x] = a.ancL ? a.ancL[p.id[x]] : a.maxDNALength;
        p.ancR[x] = a.ancR ? a.ancR[p.id[x]] : 0u;
    }
}

// this lane's read bases become PRMT selectors: the substitution score of a row is looked up in a 4-byte table
// per alignment (byte c = score against reference base c, minus the gap open score: the diagonal H arrives with
// + open on it, see s3_dp_rows)
template <int R>
__device__ __forceinline__ void s3_dp_selectors(const S3DpArgs &a, const S3DpPair &p, const uint32_t i0, uint32_t (&sel)[R])
{
    const uint32_t *readA = a.read + (size_t)(p.id[0] >> 5) * a.readWords * 32 + (p.id[0] & 31);
    const uint32_t *readB = a.read + (size_t)(p.id[1] >> 5) * a.readWords * 32 + (p.id[1] & 31);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t i = i0 + r;
        const uint32_t cA = (i <= p.m[0]) ? (readA[(size_t)(i >> 4) * 32] >> ((15 - (i & 15)) << 1)) & 3 : 0u;
        const uint32_t cB = (i <= p.m[1]) ? (readB[(size_t)(i >> 4) * 32] >> ((15 - (i & 15)) << 1)) & 3 : 0u;
        sel[r] = cA | ((cA | 8u) << 4) | ((cB | 4u) << 8) | ((cB | 12u) << 12);
    }
}

// column 0 (DV-DPfunctions.cu:167-184).  Carried per row, biased: HO = H of the previous column + open and E of
// the previous column, both UNclamped -- the reference's clamp at -32000 when it stores a value is applied
// where the value is used (ceh2 in E's max, nego2 in the diagonal's), which takes the same maximum.
template <int R>
__device__ __forceinline__ void s3_dp_column0(const S3DpPair &p, const uint32_t i0, const int open, const int gapInit, const int ext,
                                              const uint32_t KOPEN, uint32_t (&HO)[R], uint32_t (&E)[R])
{
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t i = i0 + r;
        const int hA = (i <= p.clipLt[0]) ? open : gapInit + (int)(i - p.clipLt[0]) * ext;
        const int hB = (i <= p.clipLt[1]) ? open : gapInit + (int)(i - p.clipLt[1]) * ext;
        HO[r] = s3_bpk(s3_clamp(hA), s3_clamp(hB)) + KOPEN;
        E[r] = s3_bpk(s3_clamp(hA + gapInit), s3_clamp(hB + gapInit));
    }
}

// Sweep of S3_DP_WARPS * (32 / LANES) pairs of alignments per block.  A pair is swept by a group of LANES lanes,
// each owning R consecutive read rows; at step s lane t is in column s - t.
// Best cell (DV-DPfunctions.cu:225-235): the lanes that hold rows i >= m - clipRt ("slot lanes", the last one or two
// of a group unless the right clip is long) leave their R values of every step in a ring in shared memory; when the
// ring holds LANES entries every lane of the group takes one: maximum over its eligible rows, the rows that hold it,
// against a running (best, first cell, tie count) of its own.  The reference's scan -- strictly greater replaces,
// equal counts -- gives the maximum, the number of cells that hold it and the first of them in (column, row) order,
// which is what merging the lanes' triples at the end gives.  The reference compares the UNclamped value of a cell
// with a best that starts at -32000; the values here are unclamped too, so a cell below the clamp neither beats nor
// ties that start value.
// resident blocks per SM the sweeps are compiled for: the more rows a lane holds the more registers
#ifndef S3_DP_MINBLOCKS
#define S3_DP_MINBLOCKS(R) ((R) <= 8 ? 4 : 3)
#endif
template <int R, int LANES>
__global__ void __launch_bounds__(S3_DP_WARPS * 32, S3_DP_MINBLOCKS(R))
s3_dp_sweep16_kernel(const S3DpArgs a)
{
    constexpr int GROUPS = 32 / LANES;                            // pairs per warp
    constexpr int PW = S3DpGeom<R>::PW, CKW = S3DpGeom<R>::CKW;
    // per pair and reference column j: the two substitution tables of that column (byte c of .x / .y = score
    // of alignment A / B's reference base j against read base c), built once, read every step; then the rings
    extern __shared__ uint2 s3_dp_cols[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = lane / LANES, t = lane % LANES;
    const uint32_t gmask = (LANES == 32) ? 0xFFFFFFFFu : (((1u << LANES) - 1u) << (LANES * group));
    const uint32_t numPairs = (a.count + 1) / 2;
    const uint32_t wantPair = (blockIdx.x * S3_DP_WARPS + warp) * GROUPS + group;
    if ((blockIdx.x * S3_DP_WARPS + warp) * GROUPS >= numPairs) return;          // whole warp leaves together
    const bool pairValid = wantPair < numPairs;
    const uint32_t pairLocal = pairValid ? wantPair : numPairs - 1;              // idle groups shadow a real pair, write nothing
    const bool hasB = 2 * pairLocal + 1 < a.count;
    S3DpPair p;
    p.id[0] = a.first + 2 * pairLocal; p.id[1] = a.first + 2 * pairLocal + (hasB ? 1u : 0u);
    s3_dp_load_pair(a, p);
    const uint32_t mMax = max(p.m[0], p.m[1]), nMax = pairValid ? max(p.n[0], p.n[1]) : 0u;
    const int open = a.open, ext = a.ext, gapInit = a.open - a.ext;
    // the constants of the row loop must stay in registers (ptxas otherwise re-reads them from the constant
    // bank every step and the loop waits for them): + 0 from a loaded value it cannot see through
    const uint32_t zero = p.n[0] >> 31;
    const uint32_t KOPEN = a.kOpen + zero, KEXT = a.kExt + zero, CEH2 = a.ceh2 + zero;
    const uint32_t KGAPINIT = a.kGapInit, NEGO2 = a.nego2, ONE = a.one;
    const uint32_t i0 = t * R + 1;                                 // first row of this lane (1-based)

    // reference windows (1-based packing, MSB first; DV-DPfunctions.cu:58) -> column tables in shared memory
    uint2 *cols = s3_dp_cols + (size_t)(warp * GROUPS + group) * a.colStride;
    uint32_t *ring = reinterpret_cast<uint32_t *>(s3_dp_cols + (size_t)S3_DP_WARPS * GROUPS * a.colStride) +
                     (size_t)(warp * GROUPS + group) * (2 * LANES * PW + LANES * R);
    uint32_t *keepRow = ring + 2 * LANES * PW;
    {
        const uint32_t *dnaA = a.dna + (size_t)(p.id[0] >> 5) * a.dnaWords * 32 + (p.id[0] & 31);
        const uint32_t *dnaB = a.dna + (size_t)(p.id[1] >> 5) * a.dnaWords * 32 + (p.id[1] & 31);
#pragma unroll 4
        for (uint32_t j = t; j <= nMax; j += LANES) {
            const uint32_t sh = (15u - (j & 15u)) << 1;
            const uint32_t cA = (dnaA[(size_t)(j >> 4) * 32] >> sh) & 3u, cB = (dnaB[(size_t)(j >> 4) * 32] >> sh) & 3u;
            cols[j] = make_uint2(a.mism4 ^ (a.delta << (cA << 3)), a.mism4 ^ (a.delta << (cB << 3)));
        }
    }
    uint32_t sel[R], HO[R], E[R];
    s3_dp_selectors<R>(a, p, i0, sel);
    s3_dp_column0<R>(p, i0, open, gapInit, ext, KOPEN, HO, E);

    // cells that may end the alignment: rows iLo..m, columns jLo..n
    uint32_t iLo[2], jLo[2];
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        iLo[x] = (p.clipRt[x] >= p.m[x]) ? 1u : max(p.m[x] - p.clipRt[x], 1u);
        jLo[x] = max(p.ancR[x], 1u);
    }
    const uint32_t tLastW = mMax ? (mMax - 1) / R : 0u;
    const uint32_t tiLo = (min(iLo[0], iLo[1]) - 1) / R;
    const uint32_t nSlots = tLastW - tiLo + 1;                    // lanes that hold such rows
    const uint32_t perDrain = (uint32_t)LANES / nSlots;           // steps the ring takes
    const bool slotLane = (uint32_t)t >= tiLo && (uint32_t)t <= tLastW;
    // the ring entry this lane takes when the ring is emptied: step (first buffered + myStep), lane mySlot; which halves
    // of that lane's R words belong to rows that may end the alignment (rows iLo..m)
    const uint32_t myStep = (uint32_t)t / nSlots, mySlot = tiLo + (uint32_t)t % nSlots;
    // (kept in shared memory, [row][lane]: read only when the ring is emptied)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t i = mySlot * R + r + 1;
        keepRow[r * LANES + t] = ((i >= iLo[0] && i <= p.m[0]) ? 0xFFFFu : 0u) | ((i >= iLo[1] && i <= p.m[1]) ? 0xFFFF0000u : 0u);
    }
    const uint32_t jSpan[2] = {p.n[0] - jLo[0], p.n[1] - jLo[1]};     // (wraps when no column is eligible: then jLo > n >= every j)
    const bool anyCol[2] = {jLo[0] <= p.n[0], jLo[1] <= p.n[1]};
    // the ring holds H + open (what the lanes carry); the running best is kept in that domain
    uint32_t best2 = S3_NEGB2 + KOPEN, cnt[2] = {0u, 0u}, key[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};   // key: column << 12 | row
    // the ring has two halves used in turn, so one warp-level sync per emptying is enough: a lane cannot be more than
    // one emptying ahead of another
    auto drain = [&](const uint32_t *buf, const uint32_t sFirst, const uint32_t buffered) {
        __syncwarp(gmask);
        if (myStep < buffered) {
            const uint32_t j = sFirst + myStep - mySlot;           // (wraps below column 1: masked like j > n)
            const uint32_t colKeep = ((anyCol[0] && j - jLo[0] <= jSpan[0]) ? 0xFFFFu : 0u) | ((anyCol[1] && j - jLo[1] <= jSpan[1]) ? 0xFFFF0000u : 0u);
            const uint4 *src = reinterpret_cast<const uint4 *>(buf + (size_t)t * PW);
            uint32_t w[PW];
#pragma unroll
            for (int k = 0; k < PW / 4; ++k) { const uint4 v = src[k]; w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
            uint32_t wm[R];
#pragma unroll
            for (int r = 0; r < R; ++r) wm[r] = w[r] & keepRow[r * LANES + t];     // 0 never wins
            uint32_t mx = wm[0];
#pragma unroll
            for (int r = 1; r < R; ++r) mx = __vmaxu2(mx, wm[r]);
            mx &= colKeep;
            bool gh, gl;
            (void)__vibmax_u16x2(mx, best2, &gh, &gl);
            if (gh || gl) {
                // The slot's rows in ascending order through DV-DPfunctions.cu:225-235 leave: the maximum as the new best
                // if it beats the old one (count = the rows that hold it, position = the first of them), or the rows that
                // tie with the old best added to its count.  Rows equal to the slot maximum, as bit r of each half:
                uint32_t eq = 0;
#pragma unroll
                for (int r = 0; r < R; ++r) eq |= (__vcmpeq2(wm[r], mx) & 0x00010001u) << r;
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    if (x == 0 ? !gl : !gh) continue;
                    const uint32_t rows = (eq >> (16 * x)) & 0xFFFFu, v = (mx >> (16 * x)) & 0xFFFFu, b = (best2 >> (16 * x)) & 0xFFFFu;
                    const uint32_t k = (j << 12) | (mySlot * R + (uint32_t)__ffs(rows));          // 1-based row
                    if (v > b) { cnt[x] = (uint32_t)__popc(rows); key[x] = k; }
                    else { cnt[x] += (uint32_t)__popc(rows); key[x] = min(key[x], k); }
                }
                best2 = __vmaxu2(best2, mx);
            }
        }
    };

    uint32_t upOOut = 0, FOut = 0, diagOOut = 0;
    uint32_t prevInit = S3_BIAS2;                                // start value of the previous column (0, biased)
    uint32_t clipIO[R], clipPIO[R];                              // soft-clip restart operands of the current column
    uint32_t curInit = 0xFFFFFFFFu, curPrev = 0xFFFFFFFFu;       // values clipIO / clipPIO were built from
    const bool laneHasRows = i0 <= mMax && pairValid;
    uint32_t *ck = a.ckpt + (size_t)pairLocal * a.ckStride + (size_t)t * CKW;
    uint32_t pend = 0, sFirst = 1;
    uint32_t *ringNow = ring;
    __syncwarp();

    uint32_t steps = nMax + LANES - 1;
    for (int o = LANES; o < 32; o <<= 1) steps = max(steps, __shfl_xor_sync(0xFFFFFFFFu, steps, o));   // the groups of a warp step together
#ifndef S3_DP_SWEEP_UNROLL
#define S3_DP_SWEEP_UNROLL 1
#endif
    constexpr int SWEEP_UNROLL = S3_DP_SWEEP_UNROLL;
#pragma unroll SWEEP_UNROLL
    for (uint32_t s = 1; s <= steps; ++s) {
        // carried registers of the row loop arrive from the lane above (column j was done there at step s-1):
        // H of the row above + open, F, and H of the previous column's row above + open (the diagonal)
        uint32_t upO = __shfl_up_sync(0xFFFFFFFFu, upOOut, 1, LANES);
        uint32_t F = __shfl_up_sync(0xFFFFFFFFu, FOut, 1, LANES);
        uint32_t diagO = __shfl_up_sync(0xFFFFFFFFu, diagOOut, 1, LANES);
        const uint32_t j = s - t;
        if (j >= 1 && j <= nMax && laneHasRows) {
            const uint32_t init = s3_dp_init2(j, j, p.ancL);
            if (t == 0) { upO = init + KOPEN; F = init + KGAPINIT; diagO = prevInit + KOPEN; }
            if (init != curInit || prevInit != curPrev) {           // rare: first column and anchor crossings
                s3_dp_clip_operands<R>(init, prevInit, i0, p.clipLt, KOPEN, NEGO2, clipIO, clipPIO);
                curInit = init; curPrev = prevInit;
            }
            s3_dp_rows<R>(cols[j], sel, HO, E, clipIO, clipPIO, upO, F, diagO, KOPEN, KEXT, CEH2, ONE);
            upOOut = upO; FOut = F; diagOOut = diagO;
            prevInit = init;
            if (slotLane) {
                uint32_t *dst = ringNow + ((size_t)pend * nSlots + ((uint32_t)t - tiLo)) * PW;
#pragma unroll
                for (int k = 0; k < PW / 4; ++k)
                    reinterpret_cast<uint4 *>(dst)[k] = make_uint4(HO[4 * k < R ? 4 * k : 0], HO[4 * k + 1 < R ? 4 * k + 1 : 0],
                                                                   HO[4 * k + 2 < R ? 4 * k + 2 : 0], HO[4 * k + 3 < R ? 4 * k + 3 : 0]);
            }
        }
        if (++pend == perDrain) { drain(ringNow, sFirst, pend); pend = 0; sFirst = s + 1; ringNow = ring + ((ringNow == ring) ? LANES * PW : 0); }
        if ((s & (S3_DP_CK - 1u)) == 0u && laneHasRows && s / S3_DP_CK <= a.numCk) {
            // a checkpoint: everything a sweep needs to go on from step s + 1 (the row above's H + open is HO[R-1])
            uint32_t *dst = ck + (size_t)(s / S3_DP_CK - 1u) * LANES * CKW;
            uint32_t w[CKW];
#pragma unroll
            for (int r = 0; r < R; ++r) { w[r] = HO[r]; w[R + r] = E[r]; }
            w[2 * R] = FOut; w[2 * R + 1] = diagOOut;
#pragma unroll
            for (int k = 2 * R + 2; k < CKW; ++k) w[k] = 0;
#pragma unroll
            for (int k = 0; k < CKW / 4; ++k) reinterpret_cast<uint4 *>(dst)[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
        }
    }
    if (pend) drain(ringNow, sFirst, pend);
    // merge the lanes' (best, first cell, tie count)
#pragma unroll
    for (int x = 0; x < 2; ++x) {
        const int mine = (int)((best2 >> (16 * x)) & 0xFFFFu) - 32768 - open;
        int g = mine;
        for (int o = LANES / 2; o > 0; o >>= 1) g = max(g, __shfl_xor_sync(0xFFFFFFFFu, g, o));
        // position = the first cell, in (column, row) order, that holds the final best -- if any cell beat the
        // -32000 the reference starts from (cells that only tie with it are counted, hitPos stays 0)
        uint32_t k = (mine == g && g > S3_NEG_INF) ? key[x] : 0xFFFFFFFFu;
        uint32_t c = (mine == g) ? cnt[x] : 0u;
        for (int o = LANES / 2; o > 0; o >>= 1) {
            k = min(k, __shfl_xor_sync(0xFFFFFFFFu, k, o));
            c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        }
        if (t == x && pairValid && (x == 0 || hasB)) {
            const bool any = k != 0xFFFFFFFFu;
            const uint32_t id = p.id[x], hitJ = any ? (k >> 12) : 0u, bi = k & 0xFFFu;
            a.score[id] = g;
            a.cnt[id] = c;
            a.hit[id] = hitJ;
            a.scRight[id] = any ? p.m[x] - bi : 0u;
            if (a.cells) atomicAdd(a.cells, (unsigned long long)p.m[x] * p.n[x]);
            if (g >= a.cutoff[id]) {
                // traced back: from the last checkpoint that leaves readLength + S3_DP_SLACK columns left of the end
                // column to every lane, or from column 0
                const int last = (int)hitJ - (int)p.m[x] - (int)S3_DP_SLACK - 1;
                const uint32_t s0 = (last >= (int)S3_DP_CK) ? ((uint32_t)last / S3_DP_CK) * S3_DP_CK : 0u;
                a.s0[id] = s0;
                const uint32_t cls = s0 ? 0u : 1u;
                a.tbList[(size_t)cls * a.tbCap + atomicAdd(a.tbCount + cls, 1u)] = id;
            }
        }
    }
}

// the pairs of a traceback pass: the resumed alignments two by two, then those that start at column 0
struct S3DpTbPair { uint32_t id[2]; bool valid[2]; bool resumed; };
__device__ __forceinline__ bool s3_dp_tb_pair(const S3DpArgs &a, uint32_t pairIdx, S3DpTbPair &q)
{
    const uint32_t nR = (a.pass == 1) ? a.tbCount[0] : 0u, nS = (a.pass == 1) ? a.tbCount[1] : min(a.tbCount[2], a.tbCap);
    const uint32_t pairsR = (nR + 1) / 2, pairsS = (nS + 1) / 2;
    if (pairIdx >= pairsR + pairsS) return false;
    q.resumed = pairIdx < pairsR;
    const uint32_t k0 = q.resumed ? 2 * pairIdx : 2 * (pairIdx - pairsR), cnt = q.resumed ? nR : nS;
    const uint32_t *list = a.tbList + (size_t)((a.pass == 1) ? (q.resumed ? 0 : 1) : 2) * a.tbCap;
    q.valid[0] = true; q.valid[1] = k0 + 1 < cnt;
    q.id[0] = list[k0]; q.id[1] = q.valid[1] ? list[k0 + 1] : q.id[0];
    return true;
}
__device__ __forceinline__ uint32_t s3_dp_tb_pairs(const S3DpArgs &a)
{
    const uint32_t nR = (a.pass == 1) ? a.tbCount[0] : 0u, nS = (a.pass == 1) ? a.tbCount[1] : min(a.tbCount[2], a.tbCap);
    return (nR + 1) / 2 + (nS + 1) / 2;
}

// Second sweep of the alignments that are traced back, two by two (not the pairs of the first sweep: each half
// has its own first step s0 and end column).  Local step u of half x is step s0[x] + u of its first sweep; the
// lane's R words of a step go to plane[pair][u][lane].
template <int R, int LANES>
__global__ void __launch_bounds__(S3_DP_WARPS * 32, S3_DP_MINBLOCKS(R))
s3_dp_resweep16_kernel(const S3DpArgs a)
{
    constexpr int GROUPS = 32 / LANES;
    constexpr int PW = S3DpGeom<R>::PW, CKW = S3DpGeom<R>::CKW;
    extern __shared__ uint2 s3_dp_cols[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = lane / LANES, t = lane % LANES;
    const uint32_t numPairs = s3_dp_tb_pairs(a);
    const uint32_t wantPair = (blockIdx.x * S3_DP_WARPS + warp) * GROUPS + group;
    if ((blockIdx.x * S3_DP_WARPS + warp) * GROUPS >= numPairs) return;          // whole warp leaves together
    const bool pairValid = wantPair < numPairs;
    const uint32_t pairIdx = pairValid ? wantPair : numPairs - 1;
    S3DpTbPair q;
    s3_dp_tb_pair(a, pairIdx, q);
    S3DpPair p;
    p.id[0] = q.id[0]; p.id[1] = q.id[1];
    s3_dp_load_pair(a, p);
    uint32_t s0[2], hit[2];
#pragma unroll
    for (int x = 0; x < 2; ++x) { s0[x] = q.resumed ? a.s0[p.id[x]] : 0u; hit[x] = a.hit[p.id[x]]; }
    const uint32_t mMax = max(p.m[0], p.m[1]);
    const int open = a.open, ext = a.ext, gapInit = a.open - a.ext;
    const uint32_t zero = p.n[0] >> 31;
    const uint32_t KOPEN = a.kOpen + zero, KEXT = a.kExt + zero, CEH2 = a.ceh2 + zero;
    const uint32_t KGAPINIT = a.kGapInit, NEGO2 = a.nego2, ONE = a.one;
    const uint32_t i0 = t * R + 1;
    const uint32_t tLastW = mMax ? (mMax - 1) / R : 0u;
    // steps until the last lane with rows has done both end columns
    uint32_t steps = pairValid ? min(max(hit[0] + tLastW - s0[0], hit[1] + tLastW - s0[1]), a.planeSteps - 1u) : 0u;

    // column tables of the window: entry c = u - t + LANES - 1 holds column s0[x] + u - t of half x
    uint2 *cols = s3_dp_cols + (size_t)(warp * GROUPS + group) * a.colStride;
    {
        const uint32_t *dnaA = a.dna + (size_t)(p.id[0] >> 5) * a.dnaWords * 32 + (p.id[0] & 31);
        const uint32_t *dnaB = a.dna + (size_t)(p.id[1] >> 5) * a.dnaWords * 32 + (p.id[1] & 31);
        for (uint32_t c = t; c < steps + LANES; c += LANES) {
            const int jA = (int)(s0[0] + c) - (LANES - 1), jB = (int)(s0[1] + c) - (LANES - 1);
            uint32_t cA = 0, cB = 0;
            if (jA >= 1 && jA <= (int)p.n[0]) cA = (dnaA[(size_t)(jA >> 4) * 32] >> ((15u - ((uint32_t)jA & 15u)) << 1)) & 3u;
            if (jB >= 1 && jB <= (int)p.n[1]) cB = (dnaB[(size_t)(jB >> 4) * 32] >> ((15u - ((uint32_t)jB & 15u)) << 1)) & 3u;
            cols[c] = make_uint2(a.mism4 ^ (a.delta << (cA << 3)), a.mism4 ^ (a.delta << (cB << 3)));
        }
    }
    uint32_t sel[R], HO[R], E[R];
    s3_dp_selectors<R>(a, p, i0, sel);
    uint32_t upOOut = 0, FOut = 0, diagOOut = 0, prevInit = S3_BIAS2;
    if (q.resumed) {
        // the registers of the first sweep after step s0[x], each half from the pair it was swept in
        uint32_t w[2][2 * R + 2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            const uint32_t rel = p.id[x] - a.first;
            const uint4 *src = reinterpret_cast<const uint4 *>(a.ckpt + (size_t)(rel >> 1) * a.ckStride +
                                                               ((size_t)(s0[x] / S3_DP_CK - 1u) * LANES + t) * CKW);
            const uint32_t sh = (rel & 1u) * 16u;
#pragma unroll
            for (int k = 0; k < (2 * R + 2 + 3) / 4; ++k) {
                const uint4 v = src[k];
                if (4 * k < 2 * R + 2) w[x][4 * k] = (v.x >> sh) & 0xFFFFu;
                if (4 * k + 1 < 2 * R + 2) w[x][4 * k + 1] = (v.y >> sh) & 0xFFFFu;
                if (4 * k + 2 < 2 * R + 2) w[x][4 * k + 2] = (v.z >> sh) & 0xFFFFu;
                if (4 * k + 3 < 2 * R + 2) w[x][4 * k + 3] = (v.w >> sh) & 0xFFFFu;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) { HO[r] = w[0][r] | (w[1][r] << 16); E[r] = w[0][R + r] | (w[1][R + r] << 16); }
        FOut = w[0][2 * R] | (w[1][2 * R] << 16);
        diagOOut = w[0][2 * R + 1] | (w[1][2 * R + 1] << 16);
        upOOut = HO[R - 1];
        // start value of the column before this lane's first one (>= 1: s0 >= S3_DP_CK >= LANES)
        prevInit = s3_dp_init2(s0[0] - t, s0[1] - t, p.ancL);
    } else {
        s3_dp_column0<R>(p, i0, open, gapInit, ext, KOPEN, HO, E);
    }
    __syncwarp();

    uint32_t clipIO[R], clipPIO[R];
    uint32_t curInit = 0xFFFFFFFFu, curPrev = 0xFFFFFFFFu;
    const bool laneHasRows = i0 <= mMax && pairValid;
    uint32_t *plane = a.hplane + ((size_t)pairIdx * a.planeSteps * LANES + t) * PW;
    for (int o = LANES; o < 32; o <<= 1) steps = max(steps, __shfl_xor_sync(0xFFFFFFFFu, steps, o));   // the groups of a warp step together
    for (uint32_t u = 1; u <= steps; ++u) {
        uint32_t upO = __shfl_up_sync(0xFFFFFFFFu, upOOut, 1, LANES);
        uint32_t F = __shfl_up_sync(0xFFFFFFFFu, FOut, 1, LANES);
        uint32_t diagO = __shfl_up_sync(0xFFFFFFFFu, diagOOut, 1, LANES);
        // a lane that starts at column 0 has nothing to do before its column 1; a resumed lane is past it
        if ((q.resumed || u > (uint32_t)t) && laneHasRows && u < a.planeSteps) {
            const uint32_t init = s3_dp_init2(s0[0] + u - t, s0[1] + u - t, p.ancL);
            if (t == 0) { upO = init + KOPEN; F = init + KGAPINIT; diagO = prevInit + KOPEN; }
            if (init != curInit || prevInit != curPrev) {
                s3_dp_clip_operands<R>(init, prevInit, i0, p.clipLt, KOPEN, NEGO2, clipIO, clipPIO);
                curInit = init; curPrev = prevInit;
            }
            s3_dp_rows<R>(cols[u - t + (LANES - 1)], sel, HO, E, clipIO, clipPIO, upO, F, diagO, KOPEN, KEXT, CEH2, ONE);
            upOOut = upO; FOut = F; diagOOut = diagO;
            prevInit = init;
            // the plane holds what the lanes carry: H + open, unclamped, biased
            uint32_t *hdst = plane + (size_t)u * LANES * PW;
            if (PW >= 8) {
                // 256-bit stores = the lane's whole 32-byte sectors
#pragma unroll
                for (int k = 0; k < PW / 8; ++k)
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(hdst + 8 * k), "r"(HO[8 * k < R ? 8 * k : 0]), "r"(HO[8 * k + 1 < R ? 8 * k + 1 : 0]),
                                 "r"(HO[8 * k + 2 < R ? 8 * k + 2 : 0]), "r"(HO[8 * k + 3 < R ? 8 * k + 3 : 0]), "r"(HO[8 * k + 4 < R ? 8 * k + 4 : 0]),
                                 "r"(HO[8 * k + 5 < R ? 8 * k + 5 : 0]), "r"(HO[8 * k + 6 < R ? 8 * k + 6 : 0]), "r"(HO[8 * k + 7 < R ? 8 * k + 7 : 0]) : "memory");
            } else {
                *reinterpret_cast<uint4 *>(hdst) = make_uint4(HO[0], HO[R > 1 ? 1 : 0], HO[R > 2 ? 2 : 0], HO[R > 3 ? 3 : 0]);
            }
        }
    }
}

// GPUBacktrack (DV-DPfunctions.cu:316-512) over the window plane of one alignment; called by ALL lanes of a
// warp together, each with its own alignment (`active` false: nothing to trace).  The reference reads H of three
// neighbours and, for its third test, E of the previous column.  E is not stored here.  It follows from row i of
// the plane: with a gap extension score <= 0 the clamped recurrence E(c,i) = max(-32000, open + H(c-1,i),
// ext + E(c-1,i)) (DV-DPfunctions.cu:178-183,196-199) unrolls, from any column jc on, to
//     E(j-1,i) = max(-32000, E(jc,i) + (j-1-jc) ext, max over jc <= c <= j-2 of open + H(c,i) + (j-2-c) ext)
// -- a maximum over the row, which the 32 lanes of the warp evaluate together for whichever lane needs it (one
// strided load each per 32 columns and a shuffle reduction).  jc is column 0 (H and E from their defining formulas,
// like the borders H(j,0), H(0,i)) or, for a resumed sweep, the column of the lane's checkpoint, which holds both.
// A cell left of the window sets `fail`; the caller lists the alignment for the second pass.
// Returns the start offset inside the window (the new hitLocs value).
template <int R, int LANES>
__device__ uint32_t s3_dp_traceback16(const S3DpArgs &a, bool active, const uint32_t *plane, const uint32_t *ck, uint32_t ckHalf,
                                      uint32_t s0, uint32_t half, uint32_t id, uint32_t m, uint32_t clipLt, uint32_t anchorLeft,
                                      uint32_t scRight, uint32_t hit, bool &fail)
{
    constexpr int PW = S3DpGeom<R>::PW, CKW = S3DpGeom<R>::CKW;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t *dna = a.dna + (size_t)(id >> 5) * a.dnaWords * 32 + (id & 31);
    const uint32_t *read = a.read + (size_t)(id >> 5) * a.readWords * 32 + (id & 31);
    const int open = a.open, ext = a.ext, gapInit = a.open - a.ext;
    uint8_t *pat = a.pattern + (size_t)id * (a.maxReadLength + a.maxDNALength);
    uint32_t p = 0;
    fail = false;

    auto ckH = [&](uint32_t t, uint32_t r) -> int {      // H at the lane's checkpoint column (kept as H + open, unclamped)
        const uint32_t w = ck[(size_t)t * CKW + r];
        return s3_clamp((int)((ckHalf ? w >> 16 : w) & 0xFFFFu) - 32768 - open);
    };
    auto ckE = [&](uint32_t t, uint32_t r) -> int {
        const uint32_t w = ck[(size_t)t * CKW + R + r];
        return s3_clamp((int)((ckHalf ? w >> 16 : w) & 0xFFFFu) - 32768);
    };
    auto H = [&](uint32_t j, uint32_t i) -> int {
        if (i == 0) return (j == 0) ? 0 : ((j >= anchorLeft) ? S3_NEG_INF : 0);
        if (j == 0) return s3_clamp((i <= clipLt) ? open : gapInit + (int)(i - clipLt) * ext);
        const uint32_t t = (i - 1) / R, r = (i - 1) % R;
        const int u = (int)(j + t) - (int)s0;
        if (u >= 1) {
            const uint32_t w = plane[((size_t)u * LANES + t) * PW + r];
            return s3_clamp((int)((half ? w >> 16 : w) & 0xFFFFu) - 32768 - open);      // the plane keeps H + open, unclamped, biased
        }
        if (u == 0) return ckH(t, r);
        fail = true;
        return 0;
    };
    // E(j-1, i) for every lane that wants it (see above); all lanes of the warp take part
    auto Eprev = [&](bool want, uint32_t j, uint32_t i) -> int {
        int mine = 0;
        const uint32_t t = want ? (i - 1) / R : 0u, r = want ? (i - 1) % R : 0u;
        uint32_t jc = 0;
        int hC = 0, eC = 0;
        if (want) {
            if (s0) {
                jc = s0 - t;
                if (j < jc + 1) { fail = true; want = false; }
                else { hC = ckH(t, r); eC = ckE(t, r); }
            } else {
                const int h0 = (i <= clipLt) ? open : gapInit + (int)(i - clipLt) * ext;
                hC = s3_clamp(h0); eC = s3_clamp(h0 + gapInit);
            }
        }
        uint32_t need = __ballot_sync(0xFFFFFFFFu, want);
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            // the requester's row: column c is at rowBase + (c + t - s0) * LANES * PW
            const unsigned long long rowBase = __shfl_sync(0xFFFFFFFFu, (unsigned long long)(size_t)(plane + (size_t)t * PW + r), src);
            const int uOff = __shfl_sync(0xFFFFFFFFu, (int)t - (int)s0, src);
            const uint32_t jj = __shfl_sync(0xFFFFFFFFu, j, src), hf = __shfl_sync(0xFFFFFFFFu, half, src);
            const uint32_t jcc = __shfl_sync(0xFFFFFFFFu, jc, src);
            const int hCC = __shfl_sync(0xFFFFFFFFu, hC, src), eCC = __shfl_sync(0xFFFFFFFFu, eC, src);
            const uint32_t *row = reinterpret_cast<const uint32_t *>((size_t)rowBase);
            int e = S3_NEG_INF;
            if (lane == 0) {
                e = max(e, eCC + (int)(jj - 1 - jcc) * ext);
                if (jj >= jcc + 2) e = max(e, open + hCC + (int)(jj - 2 - jcc) * ext);
            }
            for (uint32_t c = jcc + 1 + lane; c + 2 <= jj; c += 32) {
                const uint32_t w = row[(size_t)((int)c + uOff) * LANES * PW];
                e = max(e, open + s3_clamp((int)((hf ? w >> 16 : w) & 0xFFFFu) - 32768 - open) + (int)(jj - 2 - c) * ext);
            }
            for (int o = 16; o > 0; o >>= 1) e = max(e, __shfl_xor_sync(0xFFFFFFFFu, e, o));
            if ((int)lane == src) mine = e;
        }
        return mine;
    };
    auto refAt = [&](uint32_t j) -> uint32_t { return (dna[(size_t)(j >> 4) * 32] >> ((15 - (j & 15)) << 1)) & 3; };
    auto readAt = [&](uint32_t i) -> uint32_t { return (read[(size_t)(i >> 4) * 32] >> ((15 - (i & 15)) << 1)) & 3; };

    uint32_t readPos = 0, refIndex = hit, readChar = 0, refChar = 0;
    int cur = 0, next = 0, initScore = 0, prevInitScore = 0;
    int state = 0;      // 0 NORMAL, 1 I_EXT, 2 D_EXT, 3 SM_EXIT, 4 SI_EXIT
    bool run = active;
    if (active) {
        if (scRight > 0) { pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)scRight; }
        readPos = m - scRight;
        readChar = readAt(readPos); refChar = refAt(refIndex);
        cur = H(refIndex, readPos);
        initScore = (refIndex >= anchorLeft) ? S3_NEG_INF : 0;
        prevInitScore = (refIndex > anchorLeft) ? S3_NEG_INF : 0;
    }
#define S3_NEXT_REF() { --refIndex; refChar = refAt(refIndex); initScore = prevInitScore; prevInitScore = (refIndex > anchorLeft) ? S3_NEG_INF : 0; }
#define S3_NEXT_READ() { --readPos; readChar = readAt(readPos); }
    while (true) {
        run = run && !fail && readPos > 0 && refIndex > 0;
        if (!__any_sync(0xFFFFFFFFu, run)) break;
        // the two tests that need no E; the third one does
        int sel = 0, d = 0;         // 1: diagonal, 2: deletion opened here, 3: neither
        if (run && state == 0) {
            d = (refChar == readChar) ? a.match : a.mismatch;
            if (cur == d + (next = H(refIndex - 1, readPos - 1))) sel = 1;
            else if (cur == open + (next = H(refIndex - 1, readPos))) sel = 2;
            else sel = 3;
            if (fail) sel = 0;
        }
        const int ePrev = Eprev(sel == 3, refIndex, readPos);
        if (!run || fail) continue;
        if (state == 0) {
            if (sel == 1) {
                pat[p++] = (refChar == readChar) ? 'M' : 'm';
                S3_NEXT_REF(); S3_NEXT_READ();
                cur = next;
            } else if (sel == 2) {
                pat[p++] = 'D';
                S3_NEXT_REF();
                cur = next;
            } else if (cur == ext + ePrev) {
                pat[p++] = 'D';
                S3_NEXT_REF();
                cur = (short)(cur - ext);
                state = 2;
            } else {
                if (readPos <= clipLt + 1) {
                    if (cur == prevInitScore + d) { state = 3; run = false; continue; }
                    else if (cur == initScore + open) { state = 4; run = false; continue; }
                }
                if (cur == open + (next = H(refIndex, readPos - 1))) {
                    pat[p++] = 'I';
                    S3_NEXT_READ();
                    cur = next;
                } else {
                    pat[p++] = 'I';
                    S3_NEXT_READ();
                    cur = (short)(cur - ext);
                    state = 1;
                }
            }
        } else {
            if (state == 2) {
                pat[p++] = 'D';
                S3_NEXT_REF();
            } else {
                if (readPos <= clipLt + 1 && cur == initScore + open) { state = 4; run = false; continue; }
                pat[p++] = 'I';
                S3_NEXT_READ();
            }
            if (cur == open + (next = H(refIndex, readPos))) { state = 0; cur = next; }
            else cur = (short)(cur - ext);
        }
    }
#undef S3_NEXT_REF
#undef S3_NEXT_READ
    if (!active || fail) return hit;
    if (refIndex == 0) {
        const uint32_t scNum = min(clipLt, readPos);
        if (scNum < readPos) { pat[p++] = 'I'; pat[p++] = 'V'; pat[p++] = (uint8_t)(readPos - scNum); }
        pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)scNum;
    } else if (state == 4) {
        pat[p++] = 'I'; pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)(readPos - 1);
    } else if (state == 3) {
        pat[p++] = (refChar == readChar) ? 'M' : 'm';
        pat[p++] = 'S'; pat[p++] = 'V'; pat[p++] = (uint8_t)(readPos - 1);
        refIndex -= 1;
    }
    pat[p++] = 0;
    return refIndex;             // start offset inside the window (refOffset == 0 in scheme 1)
}

// Traceback, one THREAD per alignment of the pass.  A traceback is a chain of ~readLength dependent loads; what
// hides their latency is having every alignment of the chunk in flight at once, which a thread each (and few
// registers) gives and a lane group per pair does not.
template <int R, int LANES>
__global__ void __launch_bounds__(128)
s3_dp_traceback16_kernel(const S3DpArgs a)
{
    constexpr int PW = S3DpGeom<R>::PW, CKW = S3DpGeom<R>::CKW;
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t numPairs = s3_dp_tb_pairs(a);
    if ((slot & ~31u) >= 2 * numPairs) return;                    // whole warp leaves together
    S3DpTbPair q;
    const bool inRange = s3_dp_tb_pair(a, slot >> 1, q);
    const uint32_t half = slot & 1u;
    const bool active = inRange && q.valid[half];
    const uint32_t id = inRange ? q.id[half] : a.first;
    const uint32_t s0 = (inRange && q.resumed) ? a.s0[id] : 0u;
    const uint32_t m = min(a.readLen[id], a.readWords * 16u - 1u);
    const uint32_t clipLt = a.clipLt ? a.clipLt[id] : 0u;
    const uint32_t ancL = a.ancL ? a.ancL[id] : a.maxDNALength;
    const uint32_t *plane = a.hplane + (size_t)(slot >> 1) * a.planeSteps * LANES * PW;
    const uint32_t rel = id - a.first;
    const uint32_t *ck = a.ckpt + (size_t)(rel >> 1) * a.ckStride + (size_t)(s0 ? s0 / S3_DP_CK - 1u : 0u) * LANES * CKW;
    bool fail = false;
    // every lane of the warp goes in: the lanes help each other with the E rows (s3_dp_traceback16)
    const uint32_t start = s3_dp_traceback16<R, LANES>(a, active, plane, ck, rel & 1u, s0, half, id, m, clipLt, ancL, a.scRight[id], a.hit[id], fail);
    if (!active) return;
    if (fail) {
        // (cannot happen in pass 2: its window is the whole table)
        const uint32_t k = atomicAdd(a.tbCount + 2, 1u);
        if (k < a.tbCap) a.tbList[(size_t)2 * a.tbCap + k] = id;
    } else a.hit[id] = start;
}

static uint32_t dp_plane_words(const s3_dp *dp) { return dp->R <= 4 ? 4 : dp->R <= 8 ? 8 : 16; }      // S3DpGeom<R>::PW
// steps of a pair's traceback plane: pass 1 a window (see s3_dp_sweep16_kernel: at most readLength + slack +
// checkpoint distance + the lanes' stagger), pass 2 the whole table
static uint32_t dp_plane_steps(const s3_dp *dp, int pass)
{
    const uint32_t whole = dp->maxDNALength + dp->lanes + 1, window = dp->maxReadLength + S3_DP_SLACK + S3_DP_CK + dp->lanes + 1;
    return (pass == 1 && window < whole) ? window : whole;
}
// dynamic shared memory: 0 the sweep (column tables of the whole window + rings), 1 / 2 the second sweep of that pass
static size_t dp_narrow_smem(const s3_dp *dp, int which)
{
    const size_t groups = (size_t)S3_DP_WARPS * (32 / dp->lanes);
    if (which == 0) return groups * ((size_t)(dp->maxDNALength + 1) * sizeof(uint2) + ((size_t)2 * dp->lanes * dp_plane_words(dp) + (size_t)dp->lanes * dp->R) * 4);
    return groups * (size_t)(dp_plane_steps(dp, which) + dp->lanes) * sizeof(uint2);
}

// the three kernels of the 16x2 path for (rows per lane, lanes per pair)
struct S3DpNarrowKernels {
    void (*sweep)(const S3DpArgs), (*resweep)(const S3DpArgs), (*traceback)(const S3DpArgs);
};
template <int R, int LANES> static S3DpNarrowKernels dp_narrow_of()
{
    S3DpNarrowKernels k = {s3_dp_sweep16_kernel<R, LANES>, s3_dp_resweep16_kernel<R, LANES>, s3_dp_traceback16_kernel<R, LANES>};
    return k;
}
static S3DpNarrowKernels dp_narrow_kernels(const s3_dp *dp)
{
    if (dp->lanes == 8) switch (dp->R) {
        case 3: return dp_narrow_of<3, 8>(); case 4: return dp_narrow_of<4, 8>(); case 5: return dp_narrow_of<5, 8>();
        case 6: return dp_narrow_of<6, 8>(); case 7: return dp_narrow_of<7, 8>(); case 8: return dp_narrow_of<8, 8>();
        case 9: return dp_narrow_of<9, 8>(); case 10: return dp_narrow_of<10, 8>(); case 11: return dp_narrow_of<11, 8>();
        case 12: return dp_narrow_of<12, 8>(); default: return dp_narrow_of<13, 8>();
    }
    if (dp->lanes == 16) switch (dp->R) {
        case 3: return dp_narrow_of<3, 16>(); case 4: return dp_narrow_of<4, 16>(); case 5: return dp_narrow_of<5, 16>();
        case 6: return dp_narrow_of<6, 16>(); case 7: return dp_narrow_of<7, 16>(); default: return dp_narrow_of<8, 16>();
    }
    switch (dp->R) {
        case 5: return dp_narrow_of<5, 32>(); case 6: return dp_narrow_of<6, 32>(); case 7: return dp_narrow_of<7, 32>();
        case 3: return dp_narrow_of<3, 32>(); case 4: return dp_narrow_of<4, 32>();
        default: return dp_narrow_of<8, 32>();
    }
}
static int dp_narrow_set_smem(const s3_dp *dp)
{
    const S3DpNarrowKernels k = dp_narrow_kernels(dp);
    const size_t s0 = dp_narrow_smem(dp, 0), s1 = dp_narrow_smem(dp, 1), s2 = dp_narrow_smem(dp, 2);
    if (s0 > 48 * 1024) S3_CUDA(cudaFuncSetAttribute((const void *)k.sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s0));
    if ((s1 > s2 ? s1 : s2) > 48 * 1024) S3_CUDA(cudaFuncSetAttribute((const void *)k.resweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(s1 > s2 ? s1 : s2)));
    return S3_OK;
}

static int pick_R(uint32_t maxReadLength)
{
    static const int opts[] = {4, 5, 8, 16, 32};
    for (int k = 0; k < 5; ++k) if ((uint32_t)opts[k] * 32 >= maxReadLength) return opts[k];
    return 0;
}

// inside s3_dp_create, once the workspace exists: a failed call releases what has been allocated so far (the struct is
// calloc'ed, so a partial free is safe)
#define S3_CUDA_DP(call)                                                                \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            s3_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            s3_dp_free(dp);                                                             \
            return S3_ECUDA;                                                            \
        }                                                                               \
    } while (0)
extern "C" int s3_dp_create(uint32_t maxReadLength, uint32_t maxDNALength, uint32_t maxBatch, s3_dp_scores scores,
                            int device, s3_dp **out)
{
    if (!out || maxReadLength == 0 || maxDNALength == 0 || maxBatch == 0) { s3_set_error("s3_dp_create: bad argument"); return S3_EINVAL; }
    const int R = pick_R(maxReadLength);
    if (R == 0) { s3_set_error("s3_dp_create: maxReadLength %u > 1024 (MAX_READ_LENGTH, definitions.h:42)", maxReadLength); return S3_EINVAL; }
    if ((maxDNALength >> 4) + 1 > S3_DP_MAX_REF_WORDS) { s3_set_error("s3_dp_create: maxDNALength %u too large", maxDNALength); return S3_EINVAL; }
    int ndev = s3_device_count();
    if (device < 0 || device >= ndev) { s3_set_error("s3_dp_create: CUDA device %d not available (%d devices); there is no CPU fallback", device, ndev); return S3_ECUDA; }
    S3_CUDA(cudaSetDevice(device));
    s3_dp *dp = (s3_dp *)calloc(1, sizeof(s3_dp));
    if (!dp) { s3_set_error("out of host memory"); return S3_ENOMEM; }
    dp->device = device; dp->maxReadLength = maxReadLength; dp->maxDNALength = maxDNALength; dp->maxBatch = maxBatch;
    dp->sc = scores; dp->R = R; dp->slot = (R == 5) ? 8 : R;
    // 16x2 lanes hold every intermediate value when rows fit 8 per lane, the score parameters are
    // bytes and the best possible score stays below 32000 ("DP value range")
    auto small = [](int v) { return v >= -127 && v <= 127; };
    dp->narrow = maxReadLength <= 256 && small(scores.matchScore) && small(scores.mismatchScore) &&
                 small(scores.gapOpenScore) && small(scores.gapExtendScore) &&
                 (long long)scores.matchScore * maxReadLength <= 32000 && scores.gapExtendScore <= 0 &&
                 scores.matchScore - scores.gapOpenScore >= 0 && scores.matchScore - scores.gapOpenScore <= 127 &&
                 scores.mismatchScore - scores.gapOpenScore >= 0 && scores.mismatchScore - scores.gapOpenScore <= 127 &&
                 !getenv("S3_DP_FORCE_WIDE");      // extension <= 0: the traceback's closed form for E;
                                                   // substitution scores - open in [0, 127]: the score kernel's byte tables
    if (dp->narrow) {
        // lanes per pair x rows per lane: the more rows a lane holds the less each cell pays for a step's fixed work
        // (shuffles, column start values, loop), up to what the registers take: 8 lanes x <= 13 rows for reads <= 104,
        // 16 lanes x <= 8 rows up to 128, 32 x 8 up to 256; as few rows per lane as hold the read
        dp->lanes = (maxReadLength <= 104) ? 8 : (maxReadLength <= 128) ? 16 : 32;
        if (getenv("S3_DP_LANES")) { const int l = atoi(getenv("S3_DP_LANES")); if ((l == 16 && maxReadLength <= 128) || l == 32) dp->lanes = l; }   // tuning experiments
        dp->R = (int)((maxReadLength + dp->lanes - 1) / dp->lanes);
        if (dp->R < 3) dp->R = 3;
        // column tables (and the sweep's rings) in dynamic shared memory
        if (dp_narrow_smem(dp, 2) > 200 * 1024 || dp_narrow_smem(dp, 0) > 200 * 1024) dp->narrow = 0;   // very long windows take the 32-bit path
        else { int rc = dp_narrow_set_smem(dp); if (rc) { s3_dp_free(dp); return rc; } }
        if (!dp->narrow) dp->R = R;
    }
    S3_CUDA_DP(cudaStreamCreateWithFlags(&dp->stream, cudaStreamNonBlocking));
    dp->ownStream = 1;
    // traceback planes are sized per chunk of alignments; narrow: per PAIR a pass-2 plane (the whole table: what a chunk
    // needs should every traceback leave its window), i.e. per alignment half of that
    const size_t perAlign = dp->narrow
        ? (size_t)dp_plane_steps(dp, 2) * dp->lanes * 4 * dp_plane_words(dp) / 2
        : (size_t)(maxDNALength + 1) * 32 * dp->slot;
    size_t freeB = 0, totalB = 0;
    S3_CUDA_DP(cudaMemGetInfo(&freeB, &totalB));
    size_t budget = freeB / 4;
    if (budget > ((size_t)24 << 30)) budget = (size_t)24 << 30;
    size_t chunk = budget / perAlign;
    if (chunk > maxBatch) chunk = maxBatch;
    chunk = (chunk + 1) & ~(size_t)1;                 // whole pairs
    if (chunk < 2) { s3_set_error("s3_dp_create: not enough device memory for one traceback plane"); s3_dp_free(dp); return S3_ENOMEM; }
    dp->chunk = (uint32_t)chunk;
    S3_CUDA_DP(cudaMalloc(&dp->d_tb, chunk * perAlign + 256));
    const size_t up = ((size_t)maxBatch + 31) / 32 * 32;
    const size_t dnaW = (maxDNALength + 15) >> 4, readW = (maxReadLength + 15) >> 4;
    S3_CUDA_DP(cudaMalloc(&dp->d_scRight, up * 4));
    S3_CUDA_DP(cudaMalloc(&dp->d_dna, up * dnaW * 4));
    S3_CUDA_DP(cudaMalloc(&dp->d_read, up * readW * 4));
    S3_CUDA_DP(cudaMalloc(&dp->d_dnaLen, up * 4)); S3_CUDA_DP(cudaMalloc(&dp->d_readLen, up * 4));
    S3_CUDA_DP(cudaMalloc(&dp->d_hit, up * 4)); S3_CUDA_DP(cudaMalloc(&dp->d_cnt, up * 4));
    S3_CUDA_DP(cudaMalloc(&dp->d_clipLt, up * 4)); S3_CUDA_DP(cudaMalloc(&dp->d_clipRt, up * 4));
    S3_CUDA_DP(cudaMalloc(&dp->d_ancL, up * 4)); S3_CUDA_DP(cudaMalloc(&dp->d_ancR, up * 4));
    S3_CUDA_DP(cudaMalloc(&dp->d_cutoff, up * 4)); S3_CUDA_DP(cudaMalloc(&dp->d_score, up * 4));
    S3_CUDA_DP(cudaMalloc(&dp->d_pattern, up * (size_t)(maxReadLength + maxDNALength)));
    S3_CUDA_DP(cudaMalloc(&dp->d_cells, 8));
    if (dp->narrow) {
        dp->numCk = (maxDNALength + dp->lanes - 1) / S3_DP_CK;
        const size_t ckw = (size_t)(2 * dp->R + 2 + 7) / 8 * 8;
        S3_CUDA_DP(cudaMalloc(&dp->d_ckpt, (chunk / 2 + 1) * (size_t)(dp->numCk ? dp->numCk : 1) * dp->lanes * ckw * 4));
        S3_CUDA_DP(cudaMalloc(&dp->d_tbList, 3 * chunk * 4));
        S3_CUDA_DP(cudaMalloc(&dp->d_tbCount, 16));
        S3_CUDA_DP(cudaMalloc(&dp->d_s0, up * 4));
    }
    *out = dp;
    return S3_OK;
}

extern "C" void s3_dp_free(s3_dp *dp)
{
    if (!dp) return;
    cudaSetDevice(dp->device);
    cudaStreamSynchronize(dp->stream);
    void *ptrs[] = {dp->d_tb, dp->d_scRight, dp->d_dna, dp->d_read, dp->d_dnaLen, dp->d_readLen, dp->d_hit, dp->d_cnt,
                    dp->d_clipLt, dp->d_clipRt, dp->d_ancL, dp->d_ancR, dp->d_cutoff, dp->d_score, dp->d_pattern, dp->d_cells,
                    dp->d_ckpt, dp->d_tbList, dp->d_tbCount, dp->d_s0};
    for (size_t i = 0; i < sizeof ptrs / sizeof ptrs[0]; ++i) if (ptrs[i]) cudaFree(ptrs[i]);
    s3_pipe_destroy(&dp->pipe);
    s3_timing_destroy(&dp->timing);
    if (dp->d_stage) cudaFree(dp->d_stage);
    if (dp->ownStream) cudaStreamDestroy(dp->stream);
    free(dp);
}

extern "C" int s3_dp_set_timing(s3_dp *dp, int on)
{
    if (!dp) { s3_set_error("s3_dp_set_timing: NULL workspace"); return S3_EINVAL; }
    dp->timing.on = on ? 1 : 0; dp->timing.n = 0;
    return S3_OK;
}

extern "C" int s3_dp_read_timing(s3_dp *dp, float *msPerSlot, int *launchesPerSlot)
{
    if (!dp || !msPerSlot) { s3_set_error("s3_dp_read_timing: NULL argument"); return S3_EINVAL; }
    S3_CUDA(cudaSetDevice(dp->device));
    return s3_timing_read(&dp->timing, dp->stream, msPerSlot, launchesPerSlot);
}

extern "C" void *s3_dp_stream(const s3_dp *dp) { return dp ? (void *)dp->stream : NULL; }
extern "C" void s3_dp_set_stream(s3_dp *dp, void *stream)
{
    if (!dp) return;
    cudaStreamSynchronize(dp->stream);
    if (dp->ownStream) { cudaStreamDestroy(dp->stream); dp->ownStream = 0; }
    if (stream) { dp->stream = (cudaStream_t)stream; return; }
    // NULL: back to a stream of its own
    cudaSetDevice(dp->device);
    if (cudaStreamCreateWithFlags(&dp->stream, cudaStreamNonBlocking) == cudaSuccess) dp->ownStream = 1;
    else dp->stream = 0;
}
extern "C" uint32_t s3_dp_pattern_length(const s3_dp *dp) { return dp ? dp->maxReadLength + dp->maxDNALength : 0; }

template <int R>
static void launch_score(const S3DpArgs &a, cudaStream_t st)
{
    s3_dp_score_kernel<R><<<(a.count + S3_DP_WARPS - 1) / S3_DP_WARPS, S3_DP_WARPS * 32, 0, st>>>(a);
}

// alignments [begin, end) of the batch arrays in `a`
static int dp_run_device(s3_dp *dp, S3DpArgs a, uint32_t begin, uint32_t end)
{
    a.maxReadLength = dp->maxReadLength; a.maxDNALength = dp->maxDNALength;
    a.dnaWords = (dp->maxDNALength + 15) >> 4; a.readWords = (dp->maxReadLength + 15) >> 4;
    a.slot = dp->slot; a.tb = dp->d_tb; a.scRight = dp->d_scRight;
    a.match = dp->sc.matchScore; a.mismatch = dp->sc.mismatchScore; a.open = dp->sc.gapOpenScore; a.ext = dp->sc.gapExtendScore;
    if (dp->narrow) {
        a.hplane = reinterpret_cast<uint32_t *>(dp->d_tb);
        const int gapInit = dp->sc.gapOpenScore - dp->sc.gapExtendScore;
        auto k32 = [](int v) { return (uint32_t)((long long)v * 0x10001ll); };        // v in both halves of one 32-bit addend
        auto bpk = [](int v) { return (uint32_t)(v + 32768) * 0x10001u; };
        const int open = dp->sc.gapOpenScore, ext = dp->sc.gapExtendScore;
        a.kOpen = k32(open); a.kExt = k32(ext); a.kGapInit = k32(gapInit);
        a.ceh2 = bpk(S3_NEG_INF + (open > ext ? open : ext));
        a.nego2 = bpk(S3_NEG_INF + open);
        a.one = 1;
        a.mism4 = ((uint32_t)(dp->sc.mismatchScore - open) & 0xFFu) * 0x01010101u;
        a.delta = ((uint32_t)(dp->sc.matchScore - open) ^ (uint32_t)(dp->sc.mismatchScore - open)) & 0xFFu;
        a.ckpt = dp->d_ckpt; a.numCk = dp->numCk;
        a.ckStride = (size_t)(dp->numCk ? dp->numCk : 1) * dp->lanes * ((2 * dp->R + 2 + 7) / 8 * 8);
        a.tbList = dp->d_tbList; a.tbCount = dp->d_tbCount; a.tbCap = dp->chunk; a.s0 = dp->d_s0;
    }
    const S3DpNarrowKernels nk = dp->narrow ? dp_narrow_kernels(dp) : S3DpNarrowKernels();
    for (uint32_t first = begin; first < end; first += dp->chunk) {
        a.first = first;
        a.count = (end - first < dp->chunk) ? end - first : dp->chunk;
        if (dp->narrow) {
            const uint32_t perBlock = S3_DP_WARPS * (32 / dp->lanes);
            const uint32_t blocks = ((a.count + 1) / 2 + perBlock - 1) / perBlock;
            // the pairs of a traceback pass: at most every alignment of the chunk, two by two, in two classes
            const uint32_t tbBlocks = (a.count / 2 + 2 + perBlock - 1) / perBlock, tbThreads = 2 * (a.count / 2 + 2);
            S3_CUDA(cudaMemsetAsync(dp->d_tbCount, 0, 16, dp->stream));
            s3_timing_mark(&dp->timing, dp->stream, -1);
            a.pass = 0; a.colStride = dp->maxDNALength + 1;
            nk.sweep<<<blocks, S3_DP_WARPS * 32, dp_narrow_smem(dp, 0), dp->stream>>>(a);
            s3_timing_mark(&dp->timing, dp->stream, 0);
            for (int pass = 1; pass <= 2; ++pass) {
                a.pass = pass; a.planeSteps = dp_plane_steps(dp, pass); a.colStride = a.planeSteps + dp->lanes;
                nk.resweep<<<tbBlocks, S3_DP_WARPS * 32, dp_narrow_smem(dp, pass), dp->stream>>>(a);
                s3_timing_mark(&dp->timing, dp->stream, pass == 1 ? 1 : 3);
                nk.traceback<<<(tbThreads + 127) / 128, 128, 0, dp->stream>>>(a);
                s3_timing_mark(&dp->timing, dp->stream, pass == 1 ? 2 : 3);
            }
            S3_LAUNCHED(5);
            S3_CUDA(cudaGetLastError());
            continue;
        }
        switch (dp->R) {
        case 4: launch_score<4>(a, dp->stream); break;
        case 5: launch_score<5>(a, dp->stream); break;
        case 8: launch_score<8>(a, dp->stream); break;
        case 16: launch_score<16>(a, dp->stream); break;
        default: launch_score<32>(a, dp->stream); break;
        }
        S3_CUDA(cudaGetLastError());
        s3_dp_traceback_kernel<<<(a.count + 127) / 128, 128, 0, dp->stream>>>(a, dp->R);
        S3_LAUNCHED(2);
        S3_CUDA(cudaGetLastError());
    }
    return S3_OK;
}

extern "C" int s3_dp_align_device(s3_dp *dp, const uint32_t *d_dna, const uint32_t *d_dnaLen, const uint32_t *d_read,
                                  const uint32_t *d_readLen, const int32_t *d_cutoff, int32_t *d_scores,
                                  uint32_t *d_hitLocs, uint32_t *d_maxScoreCounts, uint8_t *d_pattern,
                                  uint32_t numOfThreads, const uint32_t *d_clipLt, uint32_t *d_clipRt,
                                  const uint32_t *d_ancL, const uint32_t *d_ancR)
{
    if (!dp || !d_dna || !d_dnaLen || !d_read || !d_readLen || !d_cutoff || !d_scores || !d_hitLocs || !d_maxScoreCounts || !d_pattern) {
        s3_set_error("s3_dp_align_device: NULL argument"); return S3_EINVAL;
    }
    if (numOfThreads > dp->maxBatch) { s3_set_error("s3_dp_align: %u alignments > maxBatch %u", numOfThreads, dp->maxBatch); return S3_EINVAL; }
    if (numOfThreads == 0) return S3_OK;
    S3_CUDA(cudaSetDevice(dp->device));
    S3DpArgs a;
    memset(&a, 0, sizeof a);
    a.dna = d_dna; a.dnaLen = d_dnaLen; a.read = d_read; a.readLen = d_readLen; a.cutoff = d_cutoff;
    a.score = d_scores; a.hit = d_hitLocs; a.cnt = d_maxScoreCounts; a.pattern = d_pattern;
    a.clipLt = d_clipLt; a.clipRt = d_clipRt; a.ancL = d_ancL; a.ancR = d_ancR;
    a.cells = NULL;
    return dp_run_device(dp, a, 0, numOfThreads);
}

// the same for a sub-range of the batch (the host entry point pipelines copies against it)
static int dp_align_device_range(s3_dp *dp, uint32_t begin, uint32_t end, bool clipLt, bool clipRt, bool ancL, bool ancR)
{
    S3DpArgs a;
    memset(&a, 0, sizeof a);
    a.dna = dp->d_dna; a.dnaLen = dp->d_dnaLen; a.read = dp->d_read; a.readLen = dp->d_readLen; a.cutoff = dp->d_cutoff;
    a.score = dp->d_score; a.hit = dp->d_hit; a.cnt = dp->d_cnt; a.pattern = dp->d_pattern;
    a.clipLt = clipLt ? dp->d_clipLt : NULL; a.clipRt = clipRt ? dp->d_clipRt : NULL;
    a.ancL = ancL ? dp->d_ancL : NULL; a.ancR = ancR ? dp->d_ancR : NULL;
    a.cells = NULL;
    return dp_run_device(dp, a, begin, end);
}

// ---- DP batch packing on the device --------------------------------------------
// One thread per (alignment, output word).  Word w of a packed sequence holds bases i = 16w .. 16w+15,
// 1-based, base i in bits 2(15 - (i & 15)); base 0 does not exist (DV-DPfunctions.cu:58,1469-1524).
//   DNA:  base i = text base DNAStart + i - 1 (MC_OldDnaUnpack of hsp->packedDNA, :1512-1524)
//   read: strand 1: base i = read base i - 1;  strand 2: the complement of read base length - i (:1478-1505);
//         read bases come from the query buffer (base k in bits 2(k % 16) of word k / 16, QueryParser.cpp:1146)
__global__ void s3_dp_pack_kernel(const uint32_t *__restrict__ text, const uint32_t *__restrict__ queries, uint32_t wordPerOldQuery,
                                  const uint32_t *__restrict__ readIDs, const uint8_t *__restrict__ strands,
                                  const uint32_t *__restrict__ dnaStarts, const uint32_t *__restrict__ dnaLens,
                                  const uint32_t *__restrict__ readLens, uint32_t first, uint32_t count,
                                  uint32_t dnaWords, uint32_t readWords, uint32_t *__restrict__ outDna, uint32_t *__restrict__ outRead)
{
    const uint32_t perAlign = dnaWords + readWords;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (uint64_t)count * perAlign) return;
    // consecutive threads = consecutive alignments of one word index: the 32-interleaved stores coalesce
    const uint32_t w = (uint32_t)(gid / count), id = first + (uint32_t)(gid % count);
    uint32_t word = 0;
    if (w < dnaWords) {
        const uint32_t start = dnaStarts[id], len = dnaLens[id];
        for (uint32_t i = max(16u * w, 1u); i < 16u * w + 16u && i <= len; ++i) {
            const uint32_t pos = start + i - 1;
            word |= ((text[pos >> 4] >> ((15u - (pos & 15u)) << 1)) & 3u) << ((15u - (i & 15u)) << 1);
        }
        outDna[(size_t)(id >> 5) * 32 * dnaWords + (size_t)w * 32 + (id & 31)] = word;
    } else {
        const uint32_t wr = w - dnaWords, rid = readIDs[id], len = readLens[id];
        const uint32_t *q = queries + (size_t)(rid >> 5) * 32 * wordPerOldQuery + (rid & 31);
        const bool rc = strands[id] == 2;
        for (uint32_t i = max(16u * wr, 1u); i < 16u * wr + 16u && i <= len; ++i) {
            const uint32_t k = rc ? len - i : i - 1;
            uint32_t c = (q[(size_t)(k >> 4) * 32] >> ((k & 15u) << 1)) & 3u;
            if (rc) c = 3u - c;
            word |= c << ((15u - (i & 15u)) << 1);
        }
        outRead[(size_t)(id >> 5) * 32 * readWords + (size_t)wr * 32 + (id & 31)] = word;
    }
}

// device-resident twin of s3_dp_align_windows: every pointer is a device pointer, readLengths are per ALIGNMENT,
// work is enqueued on the workspace's stream and not synchronised
extern "C" int s3_dp_align_windows_device(s3_dp *dp, s3_index *ix, const uint32_t *d_queries, uint32_t wordPerOldQuery,
                                          const uint32_t *d_readIDs, const uint8_t *d_strands, const uint32_t *d_DNAStarts,
                                          const uint32_t *d_DNALengths, const uint32_t *d_readLengths, const int32_t *d_cutoffThresholds,
                                          int32_t *d_scores, uint32_t *d_hitLocs, uint32_t *d_maxScoreCounts, uint8_t *d_pattern,
                                          uint32_t numOfThreads, const uint32_t *d_clipLtSizes, uint32_t *d_clipRtSizes,
                                          const uint32_t *d_anchorLeftLocs, const uint32_t *d_anchorRightLocs)
{
    if (!dp || !ix || !d_queries || !d_readIDs || !d_strands || !d_DNAStarts || !d_DNALengths || !d_readLengths || !d_cutoffThresholds ||
        !d_scores || !d_hitLocs || !d_maxScoreCounts || !d_pattern) { s3_set_error("s3_dp_align_windows_device: NULL argument"); return S3_EINVAL; }
    if (!ix->d_packedDNA) { s3_set_error("s3_dp_align_windows_device: the index was uploaded without the packed text"); return S3_EINVAL; }
    if (ix->device != dp->device) { s3_set_error("s3_dp_align_windows_device: index and workspace live on different devices"); return S3_EINVAL; }
    if (numOfThreads > dp->maxBatch) { s3_set_error("s3_dp_align_windows_device: %u alignments > maxBatch %u", numOfThreads, dp->maxBatch); return S3_EINVAL; }
    if (numOfThreads == 0) return S3_OK;
    S3_CUDA(cudaSetDevice(dp->device));
    const uint32_t n = numOfThreads;
    const size_t dnaW = (dp->maxDNALength + 15) >> 4, readW = (dp->maxReadLength + 15) >> 4;
    const uint64_t threads = (uint64_t)n * (dnaW + readW);
    s3_dp_pack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, dp->stream>>>(ix->d_packedDNA, d_queries, wordPerOldQuery, d_readIDs, d_strands,
                                                                                d_DNAStarts, d_DNALengths, d_readLengths, 0, n, (uint32_t)dnaW,
                                                                                (uint32_t)readW, dp->d_dna, dp->d_read);
    S3_LAUNCHED(1);
    S3_CUDA(cudaGetLastError());
    S3DpArgs a;
    memset(&a, 0, sizeof a);
    a.dna = dp->d_dna; a.dnaLen = d_DNALengths; a.read = dp->d_read; a.readLen = d_readLengths; a.cutoff = d_cutoffThresholds;
    a.score = d_scores; a.hit = d_hitLocs; a.cnt = d_maxScoreCounts; a.pattern = d_pattern;
    a.clipLt = d_clipLtSizes; a.clipRt = d_clipRtSizes; a.ancL = d_anchorLeftLocs; a.ancR = d_anchorRightLocs;
    return dp_run_device(dp, a, 0, n);
}

extern "C" int s3_dp_align_windows(s3_dp *dp, s3_index *ix,
                                   const uint32_t *queries, const uint32_t *queryLengths, uint64_t numQueries, uint32_t wordPerOldQuery,
                                   const uint32_t *readIDs, const uint8_t *strands, const uint32_t *DNAStarts, const uint32_t *DNALengths,
                                   const int32_t *cutoffThresholds, int32_t *scores, uint32_t *hitLocs, uint32_t *maxScoreCounts,
                                   uint8_t *pattern, uint32_t numOfThreads, const uint32_t *clipLtSizes, uint32_t *clipRtSizes,
                                   const uint32_t *anchorLeftLocs, const uint32_t *anchorRightLocs)
{
    if (!dp || !ix || !queries || !queryLengths || !readIDs || !strands || !DNAStarts || !DNALengths || !cutoffThresholds ||
        !scores || !hitLocs || !maxScoreCounts || !pattern) { s3_set_error("s3_dp_align_windows: NULL argument"); return S3_EINVAL; }
    if (!ix->d_packedDNA) { s3_set_error("s3_dp_align_windows: the index was uploaded without the packed text"); return S3_EINVAL; }
    if (ix->device != dp->device) { s3_set_error("s3_dp_align_windows: index and workspace live on different devices"); return S3_EINVAL; }
    if (numOfThreads > dp->maxBatch) { s3_set_error("s3_dp_align_windows: %u alignments > maxBatch %u", numOfThreads, dp->maxBatch); return S3_EINVAL; }
    if (numOfThreads == 0) return S3_OK;
    const uint32_t n = numOfThreads;
    // the caller's windows and reads must fit the workspace and the text (the reference clamps windows the same
    // way before it packs, DV-DPfunctions.cu:1439-1449)
    for (uint32_t t = 0; t < n; ++t) {
        if (readIDs[t] >= numQueries || (strands[t] != 1 && strands[t] != 2) ||
            DNALengths[t] > dp->maxDNALength || queryLengths[readIDs[t]] > dp->maxReadLength ||
            DNALengths[t] >= 16u * (uint32_t)((dp->maxDNALength + 15) >> 4) || queryLengths[readIDs[t]] >= 16u * (uint32_t)((dp->maxReadLength + 15) >> 4) ||
            queryLengths[readIDs[t]] > 16u * wordPerOldQuery ||
            (uint64_t)DNAStarts[t] + DNALengths[t] > ix->textLength) {
            s3_set_error("s3_dp_align_windows: alignment %u is out of range (read %u, window %u+%u)", t, readIDs[t], DNAStarts[t], DNALengths[t]);
            return S3_EINVAL;
        }
    }
    S3_CUDA(cudaSetDevice(dp->device));
    cudaStream_t st = dp->stream;
    const size_t up = ((size_t)n + 31) / 32 * 32, qUp = ((size_t)numQueries + 31) / 32 * 32;
    const size_t dnaW = (dp->maxDNALength + 15) >> 4, readW = (dp->maxReadLength + 15) >> 4;
    const size_t patLen = dp->maxReadLength + dp->maxDNALength;
    // staging: query buffer + per-alignment descriptors (readIDs, strands, starts) next to the workspace's own arrays
    const size_t qBytes = qUp * wordPerOldQuery * 4, need = qBytes + up * 4 * 2 + up + 1024;
    if (need > dp->stageBytes) {
        if (dp->d_stage) { S3_CUDA(cudaStreamSynchronize(st)); S3_CUDA(cudaFree(dp->d_stage)); dp->d_stage = NULL; dp->stageBytes = 0; }
        S3_CUDA(cudaMalloc(&dp->d_stage, need + need / 4));
        dp->stageBytes = need + need / 4;
    }
    uint32_t *d_q = (uint32_t *)dp->d_stage;
    uint32_t *d_rid = (uint32_t *)((char *)dp->d_stage + (qBytes + 255) / 256 * 256), *d_start = d_rid + up;
    uint8_t *d_strand = (uint8_t *)(d_start + up);
    uint32_t *h_len = (uint32_t *)malloc((size_t)n * 4);             // the reads' own lengths, per alignment
    if (!h_len) { s3_set_error("s3_dp_align_windows: out of host memory"); return S3_ENOMEM; }
    for (uint32_t t = 0; t < n; ++t) h_len[t] = queryLengths[readIDs[t]];
    cudaError_t e = cudaMemcpyAsync(d_q, queries, qBytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rid, readIDs, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_start, DNAStarts, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_strand, strands, (size_t)n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dp->d_dnaLen, DNALengths, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dp->d_readLen, h_len, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dp->d_cutoff, cutoffThresholds, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && clipLtSizes) e = cudaMemcpyAsync(dp->d_clipLt, clipLtSizes, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && clipRtSizes) e = cudaMemcpyAsync(dp->d_clipRt, clipRtSizes, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && anchorLeftLocs) e = cudaMemcpyAsync(dp->d_ancL, anchorLeftLocs, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && anchorRightLocs) e = cudaMemcpyAsync(dp->d_ancR, anchorRightLocs, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);            // h_len is read by the copy until here
    free(h_len);
    if (e != cudaSuccess) { s3_set_error("s3_dp_align_windows: staging failed: %s", cudaGetErrorString(e)); return S3_ECUDA; }
    const uint64_t threads = (uint64_t)n * (dnaW + readW);
    s3_dp_pack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(ix->d_packedDNA, d_q, wordPerOldQuery, d_rid, d_strand, d_start,
                                                                        dp->d_dnaLen, dp->d_readLen, 0, n, (uint32_t)dnaW, (uint32_t)readW,
                                                                        dp->d_dna, dp->d_read);
    S3_LAUNCHED(1);
    S3_CUDA(cudaGetLastError());
    int rc = dp_align_device_range(dp, 0, n, clipLtSizes != NULL, clipRtSizes != NULL, anchorLeftLocs != NULL, anchorRightLocs != NULL);
    if (rc) return rc;
    S3_CUDA(cudaMemcpyAsync(scores, dp->d_score, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    S3_CUDA(cudaMemcpyAsync(hitLocs, dp->d_hit, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    S3_CUDA(cudaMemcpyAsync(maxScoreCounts, dp->d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    S3_CUDA(cudaMemcpyAsync(pattern, dp->d_pattern, (size_t)n * patLen, cudaMemcpyDeviceToHost, st));
    S3_CUDA(cudaStreamSynchronize(st));
    return S3_OK;
}

extern "C" int s3_dp_align(s3_dp *dp, const uint32_t *packedDNASequence, const uint32_t *DNALengths,
                           const uint32_t *packedReadSequence, const uint32_t *readLengths,
                           const int32_t *cutoffThresholds, int32_t *scores, uint32_t *hitLocs,
                           uint32_t *maxScoreCounts, uint8_t *pattern, uint32_t numOfThreads,
                           const uint32_t *clipLtSizes, uint32_t *clipRtSizes, const uint32_t *anchorLeftLocs,
                           const uint32_t *anchorRightLocs)
{
    if (!dp || !packedDNASequence || !DNALengths || !packedReadSequence || !readLengths || !cutoffThresholds ||
        !scores || !hitLocs || !maxScoreCounts || !pattern) { s3_set_error("s3_dp_align: NULL argument"); return S3_EINVAL; }
    if (numOfThreads > dp->maxBatch) { s3_set_error("s3_dp_align: %u alignments > maxBatch %u", numOfThreads, dp->maxBatch); return S3_EINVAL; }
    if (numOfThreads == 0) return S3_OK;
    // sequences are 1-based in their slots (base i in word i >> 4, DV-DPfunctions.cu:58): a slot of w words holds 16 w - 1 bases
    for (uint32_t t = 0; t < numOfThreads; ++t)
        if (DNALengths[t] >= 16u * (uint32_t)((dp->maxDNALength + 15) >> 4) || readLengths[t] >= 16u * (uint32_t)((dp->maxReadLength + 15) >> 4)) {
            s3_set_error("s3_dp_align: alignment %u (read %u, window %u bases) does not fit the packed slots of maxReadLength %u / maxDNALength %u "
                         "(a slot of w words holds 16 w - 1 bases)", t, readLengths[t], DNALengths[t], dp->maxReadLength, dp->maxDNALength);
            return S3_EINVAL;
        }
    S3_CUDA(cudaSetDevice(dp->device));
    const size_t n = numOfThreads;
    const size_t dnaW = (dp->maxDNALength + 15) >> 4, readW = (dp->maxReadLength + 15) >> 4;
    const size_t patLen = dp->maxReadLength + dp->maxDNALength;
    cudaStream_t st = dp->stream;
    int rc;
    if ((rc = s3_pipe_init(&dp->pipe))) return rc;
    S3Pipe &pp = dp->pipe;
    // Chunks of whole 32-alignment groups (the interleave unit of the packed sequences): chunk k+1 goes up
    // while chunk k is aligned and chunk k-1 comes down.  Unlike the reference, only the filled part of the
    // fixed-size batch arrays is moved (DV-DPfunctions.cu:678-682 copies batchSize entries whatever the fill).
    // Two chunks only: each must still fill the device several times over.
    size_t chunk = (n >= 65536) ? (((n + 31) / 32 * 32) / 2 + 31) / 32 * 32 : (n + 31) / 32 * 32;
    if (getenv("S3_HOST_CHUNKS")) { const int c = atoi(getenv("S3_HOST_CHUNKS")); if (c >= 1 && c <= S3_PIPE_CHUNKS) chunk = (((n + 31) / 32 * 32) / c + 31) / 32 * 32; }   // tuning experiments
    if (chunk > dp->chunk) chunk = dp->chunk / 32 * 32;
    if (chunk < 32) chunk = 32;
    S3_CUDA(cudaEventRecord(pp.done[0], st));
    S3_CUDA(cudaStreamWaitEvent(pp.in, pp.done[0], 0));
    int k = 0;
    for (size_t c0 = 0; c0 < n; c0 += chunk, k = (k + 1) % S3_PIPE_CHUNKS) {
        const size_t cnt = (n - c0 < chunk) ? n - c0 : chunk, cntUp = (cnt + 31) / 32 * 32;
        S3_CUDA(cudaMemcpyAsync(dp->d_dna + c0 * dnaW, packedDNASequence + c0 * dnaW, cntUp * dnaW * 4, cudaMemcpyHostToDevice, pp.in));
        S3_CUDA(cudaMemcpyAsync(dp->d_read + c0 * readW, packedReadSequence + c0 * readW, cntUp * readW * 4, cudaMemcpyHostToDevice, pp.in));
        S3_CUDA(cudaMemcpyAsync(dp->d_dnaLen + c0, DNALengths + c0, cnt * 4, cudaMemcpyHostToDevice, pp.in));
        S3_CUDA(cudaMemcpyAsync(dp->d_readLen + c0, readLengths + c0, cnt * 4, cudaMemcpyHostToDevice, pp.in));
        S3_CUDA(cudaMemcpyAsync(dp->d_cutoff + c0, cutoffThresholds + c0, cnt * 4, cudaMemcpyHostToDevice, pp.in));
        if (clipLtSizes) S3_CUDA(cudaMemcpyAsync(dp->d_clipLt + c0, clipLtSizes + c0, cnt * 4, cudaMemcpyHostToDevice, pp.in));
        if (clipRtSizes) S3_CUDA(cudaMemcpyAsync(dp->d_clipRt + c0, clipRtSizes + c0, cnt * 4, cudaMemcpyHostToDevice, pp.in));
        if (anchorLeftLocs) S3_CUDA(cudaMemcpyAsync(dp->d_ancL + c0, anchorLeftLocs + c0, cnt * 4, cudaMemcpyHostToDevice, pp.in));
        if (anchorRightLocs) S3_CUDA(cudaMemcpyAsync(dp->d_ancR + c0, anchorRightLocs + c0, cnt * 4, cudaMemcpyHostToDevice, pp.in));
        S3_CUDA(cudaEventRecord(pp.up[k], pp.in));
        S3_CUDA(cudaStreamWaitEvent(st, pp.up[k], 0));
        if ((rc = dp_align_device_range(dp, (uint32_t)c0, (uint32_t)(c0 + cnt), clipLtSizes != NULL, clipRtSizes != NULL,
                                        anchorLeftLocs != NULL, anchorRightLocs != NULL))) return rc;
        S3_CUDA(cudaEventRecord(pp.done[k], st));
        S3_CUDA(cudaStreamWaitEvent(pp.out, pp.done[k], 0));
        S3_CUDA(cudaMemcpyAsync(scores + c0, dp->d_score + c0, cnt * 4, cudaMemcpyDeviceToHost, pp.out));
        S3_CUDA(cudaMemcpyAsync(hitLocs + c0, dp->d_hit + c0, cnt * 4, cudaMemcpyDeviceToHost, pp.out));
        S3_CUDA(cudaMemcpyAsync(maxScoreCounts + c0, dp->d_cnt + c0, cnt * 4, cudaMemcpyDeviceToHost, pp.out));
        S3_CUDA(cudaMemcpyAsync(pattern + c0 * patLen, dp->d_pattern + c0 * patLen, cnt * patLen, cudaMemcpyDeviceToHost, pp.out));
    }
    S3_CUDA(cudaStreamSynchronize(pp.out));
    S3_CUDA(cudaStreamSynchronize(st));
    return S3_OK;
}
