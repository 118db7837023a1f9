/*
 * soap3dp_b200.h -- C ABI of the B200-native GPU alignment hot path of SOAP3-dp.
 *
 * Every entry point names the reference interface it replaces (file:line in
 * aquaskyline/SOAP3-dp).  Plain pointers and sizes only; all functions return
 * 0 on success or a negative S3_E* code, with a message in s3_last_error().
 * Unless a name ends in `_device`, pointers are HOST pointers and the call is
 * synchronous (results are in host memory when it returns), exactly like the
 * reference's perform_round*_alignment / SemiGlobalAligner::performAlignment.
 *
 * There is no CPU fallback: if no CUDA device is usable the calls fail.  A few
 * entries replace code that runs on the host in the reference as well and are
 * host code here too (no device involved): s3_dp_decode, s3_runs_decode,
 * s3_dp_md, s3_seed_layout, s3_dp_stage_parameters, s3_mapq_*, s3_sam_*.
 *
 * Sections, in the order of a batch's way through the aligner: index
 * (s3_index_upload / _load / _clone), search (s3_search_round1 / round2,
 * the capless s3_search, the seeding driver s3_seed_search), locate and the
 * seed-hit merges, DP (s3_dp_*, window selection), decoders, the chains
 * (s3_pe_* = alignPairR with s3_pe_deep_dp, s3_se_* = alignSingleR), the
 * seeded DP stages (s3_single_dp_align, s3_deep_dp_align), stage tables,
 * mapping qualities, measurement hooks, SAM records.
 */
#ifndef SOAP3DP_B200_H
#define SOAP3DP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3_OK          0
#define S3_ECUDA      -1   /* a CUDA runtime call failed            */
#define S3_EINVAL     -2   /* bad argument                          */
#define S3_ENOMEM     -3   /* host or device allocation failed      */

#define S3_MAX_NUM_CASES 10          /* definitions.h:121 MAX_NUM_CASES */

typedef struct s3_index s3_index;    /* device-resident 2BWT index (opaque) */
typedef struct s3_dp    s3_dp;       /* device-resident DP workspace (opaque) */

const char *s3_last_error(void);
int s3_device_count(void);
/* number of kernels this library has launched in this process so far */
unsigned long long s3_launch_count(void);

/* ------------------------------------------------------------------------
 * Index.  Replaces GPUINDEXUpload / GPUINDEXFree (alignment.cu:27-115,
 * alignment.h:41-45).  Inputs are the host arrays of Soap3Index
 * (IndexHandler.h:46-60): bwt->bwtCode / rev_bwt->bwtCode (2 bit/base, 16
 * bases per word MSB first; at least ceil(textLength/16) words are read),
 * gpu_occValue / gpu_revOccValue (numOcc*4 words, BGS-Build.cpp:141-165),
 * and the scalars the kernels receive by value.  The arrays are re-laid out on
 * the device into 32-byte buckets {4 x uint32 running counts, 64 bases as bit planes} and
 * seed tables of the first floor(log4 n) exact steps are built (3 x 8.6 GB at 3.1 Gbp); the
 * caller keeps ownership of the host arrays.  packedDNA (hsp->packedDNA, 16
 * bases/word MSB first) and sa (bwt->saValue, the full suffix array of the
 * n + 1 BWT rows, SaValueFreq == 1 as in soap3-dp-builder.ini:29) are optional:
 * when both are given the search finishes a read by comparing it with the
 * text as soon as its interval is a single suffix ("check and extend", what
 * the reference's CPU search does, 2bwt-flex/SRA2BWTCheckAndExtend.c) instead
 * of stepping through the rest of it; answers are identical either way.
 * ------------------------------------------------------------------------ */
int s3_index_upload(const uint32_t *bwt, const uint32_t *occ,
                    const uint32_t *revBwt, const uint32_t *revOcc,
                    uint32_t numOcc, uint32_t inverseSa0, uint32_t revInverseSa0,
                    uint32_t textLength,
                    const uint32_t *packedDNA, const uint32_t *sa,
                    int device, s3_index **out);
/* the same two optional arrays when they already live in device memory (copied) */
int s3_index_set_locate_device(s3_index *ix, const uint32_t *d_sa, const uint32_t *d_packedDNA);
void s3_index_free(s3_index *ix);
/* The reference's index files straight from disk (INDEXLoad, IndexHandler.cpp:118-330, for the arrays the device needs):
 * <prefix>.bwt, .fmv.gpu, .rev.bwt, .rev.fmv.gpu as soap3-dp-builder + BGS-Build write them (5-word header, payload) and, with
 * withText, .sa (full suffix array: SaValueFreq = 1) and .pac for check-and-extend, locate and the chains.  The files are
 * mapped, page-locked for the copy when the driver allows it, and handed to s3_index_upload; headers are cross-checked. */
int s3_index_load(const char *prefix, int withText, int device, s3_index **out);
/* A second handle on the same device arrays with a stream, work queues and scratch of its own: two batches can then be in
 * flight at once, one host thread per handle (the reference overlaps its search of batch k + 1 with the DP of batch k the
 * same way, alignment.cu:555,1030).  Clones are freed before the handle they were made from. */
int s3_index_clone(s3_index *ix, s3_index **out);
/* bytes of device memory held by the index */
size_t s3_index_device_bytes(const s3_index *ix);
/* the CUDA stream (cudaStream_t) all work of this index is issued on */
void *s3_index_stream(const s3_index *ix);

/* rank'(c, i) = cumulativeFreq[c] + Occ(c, i) probes evaluated on the device
 * with the re-laid-out index; replaces nothing (test hook for GPUBWTOccValue,
 * DV-Kernel.cu:256).  which = 0 forward BWT, 1 reverse BWT.  out[4*i + c]. */
int s3_rank_probe(s3_index *ix, int which, const uint32_t *indices, size_t n, uint32_t *out);

/* ------------------------------------------------------------------------
 * Search, round 1.  Replaces perform_round1_alignment (alignment.cu:118-215,
 * alignment.h:48-51) and its _no_pipeline twin (:329-424): for each of the
 * numCases cases of the numMismatch scheme, both strands, all SA ranges of
 * every read into answers[c] (uint32[ceil32(batchSize) * wordPerAns], the
 * reference's 32-read interleaved slot format, status words 0xFFFFFFFD /
 * 0xFFFFFFFE, and the isBad carry-over between cases; DV-Kernel.cu:4249-4503).
 * queries / readLengths: QueryParser.cpp:1146-1152 layout; not modified.
 * ------------------------------------------------------------------------ */
int s3_search_round1(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths,
                     uint64_t batchSize, uint32_t wordPerQuery,
                     uint32_t numMismatch, uint32_t numCases,
                     uint32_t saRangeAllowed, uint32_t wordPerAns,
                     int isExactNumMismatch, uint32_t *const *answers);

/* Search, round 2.  Replaces perform_round2_alignment (alignment.cu:221-326)
 * and _no_pipeline (:426-531): for each case, the reads whose round-1 status
 * word is > 0xFFFFFFFD are gathered in read order into badReadIndices[c]
 * (numBad[c] entries) and searched again with the larger slot into
 * badAnswers[c] (uint32[ceil32(numBad[c]) * wordPerAns2]).  queries are the
 * whole host query buffer and processedQuery the offset of this batch in it
 * (alignment.cu:258).  badReadIndices[c] / badAnswers[c] must have room for
 * batchSize reads. */
int s3_search_round2(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths,
                     uint32_t *const *answers,
                     uint64_t batchSize, uint64_t processedQuery, uint32_t wordPerQuery,
                     uint32_t numMismatch, uint32_t numCases,
                     uint32_t saRangeAllowed2, uint32_t wordPerAns, uint32_t wordPerAns2,
                     int isExactNumMismatch,
                     uint32_t *const *badReadIndices, uint32_t *const *badAnswers,
                     uint64_t *numBad);

/* Device-resident variant of round 1 used for kernel timing and by callers
 * that keep reads in HBM: d_queries, d_readLengths, d_answers[c] are device
 * pointers; work is enqueued on s3_index_stream() and NOT synchronised.
 * d_rankQueries (device uint64, may be NULL) receives the number of rank
 * evaluations performed (the roofline unit). */
int s3_search_round1_device(s3_index *ix, const uint32_t *d_queries, const uint32_t *d_readLengths,
                            uint64_t batchSize, uint32_t wordPerQuery,
                            uint32_t numMismatch, uint32_t numCases,
                            uint32_t saRangeAllowed, uint32_t wordPerAns,
                            int isExactNumMismatch, uint32_t *const *d_answers,
                            unsigned long long *d_rankQueries);

/* ------------------------------------------------------------------------
 * Capless search.  Replaces round 1 + round 2 + the CPU fallback for reads whose slot still
 * overflows (all_valid_alignment, alignment.cu:855-954; CPUfunctions.cpp:1310-1329,1394-1412)
 * by one call: every SA range of every read for all cases of the numMismatch scheme, no slot
 * cap, as CSR.  Ranges of read q are entries offsets[q] .. offsets[q+1]-1: case ascending, inside
 * a case in the enumeration order of round 1 -- the reference's round-1 slot of (q, case) holds
 * the first saRangeAllowed of them.  saL/saR are inclusive bounds on the forward BWT;
 * info = strand | numMismatches << 1 | case << 4.  The arrays are malloc'ed by the library;
 * release them with s3_search_result_free.
 * ------------------------------------------------------------------------ */
typedef struct {
    uint64_t numReads, total;
    uint64_t *offsets;            /* numReads + 1 */
    uint32_t *saL, *saR, *info;   /* total each */
} s3_search_result;

int s3_search(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths,
              uint64_t batchSize, uint32_t wordPerQuery, uint32_t numMismatch,
              int isExactNumMismatch, s3_search_result *out);
void s3_search_result_free(s3_search_result *r);

/* ------------------------------------------------------------------------
 * Locate.  SA ranges -> text positions from the suffix array in HBM; replaces the host loops
 * `for k in [l, r]: (*bwt->_bwtSaValue)(bwt, k)` that follow every search (SAList.cpp:411,
 * CPUfunctions.cpp:2915,2975, PEAlgnmt.cpp:1246, DV-DPfunctions.cu:1186,2931).  Needs an index
 * uploaded with its suffix array.  Positions of range g: (*positions)[offsets[g] ..
 * offsets[g+1]-1] = SA[saL[g] + k] for k < min(saR[g]-saL[g]+1, maxPerRange), in SA order like
 * the reference's loops; an empty range (saR < saL) gives none.  offsets has numRanges + 1
 * entries; *positions is malloc'ed by the library (s3_free).
 * ------------------------------------------------------------------------ */
int s3_locate(s3_index *ix, const uint32_t *saL, const uint32_t *saR, uint64_t numRanges,
              uint32_t maxPerRange, uint64_t *offsets, uint32_t **positions, uint64_t *total);
void s3_free(void *p);

/* ------------------------------------------------------------------------
 * Seed hits -> candidate positions for single-end DP seeding.  Replaces
 * SingleEndSeedingBatch::decodePositions + singleMerge (DV-DPfunctions.cu:1101-1219): for
 * every SA range g of a seed (strand 1 / 2 as SARecord.strand, the seed's read, its offset in
 * the read, its length, the read's length) the estimated read start of every occurrence
 * (SA[k] - offset, or SA[k] + seedLength + offset - readLength on strand 2, uint arithmetic),
 * ordered by (readID, start) with ties in arrival order -- what the reference's three radix
 * passes leave -- and thinned per read to the hits more than DPS_DIVIDE_GAP = 50 past the last
 * one kept.  maxPerRange caps the occurrences taken from one range (0xFFFFFFFF: all, as the
 * reference).  Needs an index uploaded with its suffix array.  The three output arrays are
 * malloc'ed by the library (s3_free).
 * ------------------------------------------------------------------------ */
int s3_seed_candidates(s3_index *ix, const uint32_t *saL, const uint32_t *saR, const int32_t *strands,
                       const uint32_t *readIDs, const uint32_t *offsets, const uint32_t *seedLengths,
                       const uint32_t *readLengths, uint64_t numRanges, uint32_t maxPerRange,
                       uint32_t **candReadIDs, uint32_t **candPositions, int32_t **candStrands,
                       uint64_t *numCandidates);

/* The same for paired-end DP seeding.  Replaces PairEndSeedingBatch::decodeMergePositions
 * (DV-DPfunctions.cu:2626-2653,2780-2999): the seed ranges of the reads (arrays ...0) and of
 * their mates (arrays ...1; readIDs hold the pair's even read id on both sides, as
 * readIDs[readOrMate][seedID] does), estimated starts, the 50-base thinning of the left
 * group, the insert-size join [insertLow - len - margin, insertHigh - len + margin] with
 * margin = DP2_MARGIN(len) and len = lengthsByReadID[pair id], for read-left/mate-right and
 * mate-left/read-right with the leg strands peStrandLeftLeg / peStrandRightLeg (1 or 2), and
 * the final stable sort by readIDLeft (= pair id, + 1 when the mate is the left end).
 * Output arrays are malloc'ed by the library (s3_free). */
int s3_seed_pair_candidates(s3_index *ix,
                            const uint32_t *saL0, const uint32_t *saR0, const int32_t *strands0, const uint32_t *readIDs0,
                            const uint32_t *offsets0, const uint32_t *seedLengths0, const uint32_t *readLengths0, uint64_t n0,
                            const uint32_t *saL1, const uint32_t *saR1, const int32_t *strands1, const uint32_t *readIDs1,
                            const uint32_t *offsets1, const uint32_t *seedLengths1, const uint32_t *readLengths1, uint64_t n1,
                            uint32_t maxPerRange, const uint32_t *lengthsByReadID, uint64_t numReadIDs,
                            int insertLow, int insertHigh, int peStrandLeftLeg, int peStrandRightLeg,
                            uint32_t **candReadIDLeft, uint32_t **candPosLeft, uint32_t **candPosRight,
                            uint64_t *numCandidates);

/* ------------------------------------------------------------------------
 * The seeding driver of the DP stages.  Replaces single_1_mismatch_alignment2 (alignment.cu:1839-1893) with hostKernelSingle's
 * per-seed bookkeeping (CPUfunctions.cpp:2700-2870): every seed is searched exactly; the seeds without any alignment are
 * searched again with at most 1 mismatch; a seed keeps its SA ranges (SARecord: saLeft, saRight, strand 1 / 2) when they
 * hold at most maxHitNum[seed] occurrences, else none (status 4, too many hits; ProceedDPForTooManyHits = 0).  seeds /
 * seedLengths are a query buffer in the usual layout (QueryParser.cpp:1146-1152) with wordPerSeed words per seed, as the
 * seeding batches build them (DV-DPfunctions.cu:2655-2680).  Ranges of seed s are [offsets[s], offsets[s+1]) in the
 * reference's order: cases ascending, enumeration order inside a case.  status[s]: 0 no hit, 1 ranges kept, 4 too many.
 * The arrays are malloc'ed by the library (s3_seed_search_result_free).
 * ------------------------------------------------------------------------ */
typedef struct {
    uint64_t numSeeds, total;
    uint64_t *offsets;            /* numSeeds + 1 */
    uint32_t *saL, *saR;          /* total each */
    uint8_t *strand;              /* total */
    uint8_t *status;              /* numSeeds */
} s3_seed_search_result;
int s3_seed_search(s3_index *ix, const uint32_t *seeds, const uint32_t *seedLengths, uint64_t numSeeds, uint32_t wordPerSeed,
                   const uint32_t *maxHitNum, s3_seed_search_result *out);
void s3_seed_search_result_free(s3_seed_search_result *r);

/* Tuning knob, answers are identical for every value.  A (read, case) enumeration that is
 * still running in its lane after `steps` LF-mapping steps is split: the substitution children
 * along the read's own path become independent tasks for other lanes and their ranges are merged
 * back in enumeration order under the slot's cap (round-1 slots, saRangeAllowed <= 4).  Default
 * 256; 0 splits every enumerated item; negative: never split (one lane per item, as the
 * reference runs one thread per read, DV-Kernel.cu:4249). */
int s3_search_set_split_budget(s3_index *ix, int32_t steps);

/* ------------------------------------------------------------------------
 * Semi-global affine-gap DP.  Replaces SemiGlobalAligner::{decideConfiguration,
 * init, performAlignment, freeMemory} (DV-DPfunctions.h:120-164,
 * DV-DPfunctions.cu:520-741) and the two kernels SemiGlobalAligntment /
 * GPUBacktrack (DV-DPfunctions.cu:243,316), scheme 1 (full table) semantics.
 * Buffers are the reference's fixed-stride batch arrays (DV-DPfunctions.cu:
 * 57-59,1469-1524): packed sequences 32-interleaved, 1-based, MSB first.
 * NULL clip / anchor arrays mean "none" (DV-DPfunctions.cu:258-263).
 * ------------------------------------------------------------------------ */
typedef struct {
    int32_t matchScore, mismatchScore, gapOpenScore, gapExtendScore;   /* soap3-dp.ini [DP] */
} s3_dp_scores;

/* Sequences are 1-based in their slots (base i in word i >> 4), so a slot of w = ceil(max / 16) words holds 16 w - 1
 * bases: a length equal to a maximum that is a multiple of 16 does not fit (the reference sizes maxReadLength as
 * (L / 4 + 1) * 4 and maxDNALength with + 8 slack, DV-DPfunctions.cu:1580-1581, and never meets the case).  The host
 * entries reject such lengths with S3_EINVAL; the device entries cut them to what the slot holds. */
int s3_dp_create(uint32_t maxReadLength, uint32_t maxDNALength, uint32_t maxBatch,
                 s3_dp_scores scores, int device, s3_dp **out);
void s3_dp_free(s3_dp *dp);
void *s3_dp_stream(const s3_dp *dp);
/* issue this workspace's kernels and copies on `stream` (a cudaStream_t, e.g.
 * s3_index_stream()) so that search and DP of one batch are stream-ordered;
 * NULL gives the workspace a stream of its own again */
void s3_dp_set_stream(s3_dp *dp, void *stream);
/* PatternLength() = maxReadLength + maxDPTableLength bytes per alignment
 * (DV-DPfunctions.cu:54); maxDPTableLength = maxDNALength in scheme 1
 * (decideConfiguration, DV-DPfunctions.cu:592). */
uint32_t s3_dp_pattern_length(const s3_dp *dp);

int s3_dp_align(s3_dp *dp,
                const uint32_t *packedDNASequence, const uint32_t *DNALengths,
                const uint32_t *packedReadSequence, const uint32_t *readLengths,
                const int32_t *cutoffThresholds,
                int32_t *scores, uint32_t *hitLocs, uint32_t *maxScoreCounts,
                uint8_t *pattern, uint32_t numOfThreads,
                const uint32_t *clipLtSizes, uint32_t *clipRtSizes,
                const uint32_t *anchorLeftLocs, const uint32_t *anchorRightLocs);
/* Like the reference, clipRtSizes is only read on the host side: the kernel's
 * actual right clip is reported through the leading 'S','V',n of the pattern
 * (performAlignment copies back scores, hitLocs, pattern and maxScoreCounts
 * only, DV-DPfunctions.cu:719-722).  hitLocs[t] is the 0-based start offset in
 * the window when scores[t] >= cutoffThresholds[t] (pattern written), else the
 * end column of the best cell (pattern untouched), as in the reference. */

/* device-resident twin (pointers are device pointers, stream-ordered, no sync).
 * d_cells (device uint64, may be NULL) receives sum readLength*DNALength. */
int s3_dp_align_device(s3_dp *dp,
                       const uint32_t *d_packedDNASequence, const uint32_t *d_DNALengths,
                       const uint32_t *d_packedReadSequence, const uint32_t *d_readLengths,
                       const int32_t *d_cutoffThresholds,
                       int32_t *d_scores, uint32_t *d_hitLocs, uint32_t *d_maxScoreCounts,
                       uint8_t *d_pattern, uint32_t numOfThreads,
                       const uint32_t *d_clipLtSizes, uint32_t *d_clipRtSizes,
                       const uint32_t *d_anchorLeftLocs, const uint32_t *d_anchorRightLocs);

/* ------------------------------------------------------------------------
 * DP on windows named by the caller: packing on the device + s3_dp_align.  Replaces the
 * base-at-a-time host packers packRead / repackDNA of the three DP engines
 * (SingleEndAlgnBatch, HalfEndAlgnBatch, PairEndAlgnBatch: DV-DPfunctions.cu:1469-1524,
 * 2111-2167, 3474-3529) together with performAlignment: alignment t aligns read readIDs[t]
 * (0-based index into the query buffer `queries`, QueryParser.cpp:1146-1152 layout,
 * wordPerOldQuery words per read; strands[t] = 1 as given, 2 reverse-complemented, as
 * CandidateInfo.strand) against text bases [DNAStarts[t], DNAStarts[t] + DNALengths[t]) of the
 * packed text the index was uploaded with.  What windows to align (margins, insert sizes,
 * anchors, clips) stays the caller's decision, as in the engines' pack() functions.
 * 28 bytes per alignment go to the device instead of the packed batch arrays.  Outputs and
 * the clip / anchor arrays are those of s3_dp_align.
 * ------------------------------------------------------------------------ */
int s3_dp_align_windows(s3_dp *dp, s3_index *ix,
                        const uint32_t *queries, const uint32_t *queryLengths, uint64_t numQueries,
                        uint32_t wordPerOldQuery,
                        const uint32_t *readIDs, const uint8_t *strands,
                        const uint32_t *DNAStarts, const uint32_t *DNALengths,
                        const int32_t *cutoffThresholds,
                        int32_t *scores, uint32_t *hitLocs, uint32_t *maxScoreCounts,
                        uint8_t *pattern, uint32_t numOfThreads,
                        const uint32_t *clipLtSizes, uint32_t *clipRtSizes,
                        const uint32_t *anchorLeftLocs, const uint32_t *anchorRightLocs);

/* device-resident twin: every pointer is a device pointer (d_readLengths per ALIGNMENT, not per query), the work is
 * enqueued on s3_dp_stream() and not synchronised; numOfThreads <= maxBatch of the workspace */
int s3_dp_align_windows_device(s3_dp *dp, s3_index *ix, const uint32_t *d_queries, uint32_t wordPerOldQuery,
                               const uint32_t *d_readIDs, const uint8_t *d_strands,
                               const uint32_t *d_DNAStarts, const uint32_t *d_DNALengths, const uint32_t *d_readLengths,
                               const int32_t *d_cutoffThresholds,
                               int32_t *d_scores, uint32_t *d_hitLocs, uint32_t *d_maxScoreCounts,
                               uint8_t *d_pattern, uint32_t numOfThreads,
                               const uint32_t *d_clipLtSizes, uint32_t *d_clipRtSizes,
                               const uint32_t *d_anchorLeftLocs, const uint32_t *d_anchorRightLocs);

/* ------------------------------------------------------------------------
 * Which windows are aligned.  Replaces the deciding half of the three DP engines' batch packers -- SingleEndAlgnBatch::pack
 * (DV-DPfunctions.cu:1425-1468), HalfEndAlgnBatch::pack (:2027-2110), PairEndAlgnBatch::packLeft (:3374-3418) and packRight
 * (:3420-3472) -- for a batch of candidates, on the device: for every candidate the text window, clips, anchors and cutoff
 * the reference would align it with, in the arrays s3_dp_align_windows takes.  Candidates (numCandidates each):
 *   S3_WIN_SINGLE      readIDs, positions (estimated read start, s3_seed_candidates), strands                 -> 1 window each
 *   S3_WIN_HALF        readIDs / positions / strands of the ALIGNED read's occurrences                        -> 0..2 windows each
 *                      for its mate (readID ^ 1), windows in candidate order (HalfEndOccStream's order)
 *   S3_WIN_PAIR_LEFT   readIDs = readIDLeft, positions = its estimated start (s3_seed_pair_candidates)         -> 1 window each
 *   S3_WIN_PAIR_RIGHT  the same candidates with positions2 = the right read's estimated start and the left    -> 0..1 windows each
 *                      pass's leftScores / leftStarts (its DNAStarts) / leftHitLocs: only candidates whose left
 *                      read reached its cutoff; window cut at hitPosLeft + insertHigh, right anchor at + insertLow
 * readLengths is indexed by read id (numReads entries; reads 2p, 2p + 1 are mates).  Outputs hold 2 * numCandidates
 * entries for S3_WIN_HALF, numCandidates otherwise: outCandidate (index of the candidate a window belongs to), outReadIDs
 * (the read to align), outStrands (1 as given, 2 reverse-complemented), outLeftOrRight (CandidateInfo.leftOrRight of the
 * half-end engine), then the arrays of s3_dp_align_windows.  cutoffThreshold < 0 means DEFAULT = ceil(0.3 * read length).
 * ------------------------------------------------------------------------ */
#define S3_WIN_SINGLE     1
#define S3_WIN_HALF       2
#define S3_WIN_PAIR_LEFT  3
#define S3_WIN_PAIR_RIGHT 4
typedef struct {
    int32_t insertLow, insertHigh, strandLeftLeg, strandRightLeg;
    int32_t softClipLeft, softClipRight;
    int32_t cutoffThreshold[2];          /* paramRead[0 / 1]: the even / odd read of a pair */
    uint32_t maxDNALength;               /* the workspace's maxDNALength: "no left anchor" is stored as this value */
} s3_window_params;
int s3_dp_make_windows(s3_index *ix, int mode, const s3_window_params *par, const uint32_t *readLengths, uint64_t numReads,
                       const uint32_t *readIDs, const uint32_t *positions, const uint32_t *positions2, const uint8_t *strands,
                       const int32_t *leftScores, const uint32_t *leftStarts, const uint32_t *leftHitLocs, uint64_t numCandidates,
                       uint32_t *outCandidate, uint32_t *outReadIDs, uint8_t *outStrands, uint8_t *outLeftOrRight,
                       uint32_t *DNAStarts, uint32_t *DNALengths, uint32_t *clipLtSizes, uint32_t *clipRtSizes,
                       uint32_t *anchorLeftLocs, uint32_t *anchorRightLocs, int32_t *cutoffThresholds, uint64_t *numWindows);

/* ------------------------------------------------------------------------
 * DP result decoding (host work on the arrays s3_dp_align* returned).  Replaces the result loops
 * of the DP engines' CPU threads (SingleDP_Space::algnmtCPUThread DV-DPfunctions.cu:1699-1733,
 * DP_Space::algnmtCPUThread :2359-2400, DeepDP_Space::DP2CPUAlgnThread :3765-3795) with
 * CigarStringEncoder (DV-DPfunctions.h:514-597), and convertToCigarStr (PE.cpp:420-485).  For
 * every alignment t with scores[t] >= cutoffThresholds[t]:
 *   cigars[cigarOffsets[t] .. cigarOffsets[t+1])   the reference's "special" CIGAR in read order
 *                                                  ('M' match and 'm' mismatch kept apart, I, D, S)
 *   samCigars[samOffsets[t] .. samOffsets[t+1])    its SAM form (optional: pass both NULL to skip)
 *   editdist[t]      I + D + (L*match + gapPenalty - score) / (match - mismatch), L = readLength - I - S
 *   refSpanDelta[t]  D - I - S: reference bases covered = readLength + refSpanDelta (the insertSize terms)
 *   opCounts[5t..]   bases in M, m, I, D, S
 * and an empty string, editdist -1, zeros otherwise (the reference emits no result for those).
 * The alignment's text position stays windowStart + hitLocs[t] and num_sameScore stays
 * maxScoreCounts[t].  Strings carry no terminators between alignments; *cigars / *samCigars are
 * malloc'ed by the library (s3_free).  editdist, refSpanDelta, opCounts may be NULL.
 * ------------------------------------------------------------------------ */
int s3_dp_decode(const uint8_t *pattern, uint32_t patternLength, const int32_t *scores,
                 const uint32_t *readLengths, const int32_t *cutoffThresholds, uint32_t numOfThreads,
                 s3_dp_scores scores4,
                 uint64_t *cigarOffsets, char **cigars, uint64_t *samOffsets, char **samCigars,
                 int32_t *editdist, int32_t *refSpanDelta, uint32_t *opCounts);

/* The same for ONE alignment that is already held as (op, count) runs -- what s3_pe_align, s3_single_dp_align and s3_deep_dp_align /
 * s3_pe_deep_dp return (length << 8 | op, read order): the special CIGAR string (NUL-terminated, cigarCapacity bytes of room), the
 * edit distance and the insert-size term D - I - S, i.e. what the SAM writers below take for a DP alignment. */
int s3_runs_decode(const uint32_t *runs, uint32_t numRuns, uint32_t readLength, int32_t score, s3_dp_scores scores4, char *cigar, uint32_t cigarCapacity,
                   uint32_t *cigarLength, int32_t *editdist, int32_t *refSpanDelta);

/* MD:Z and the NM pieces of decoded alignments.  Replaces getMisInfoForDP (PE.cpp:499-666, trim 0)
 * for a batch: from the special CIGARs of s3_dp_decode (cigars / cigarOffsets), the alignments'
 * text positions (windowStart + hitLoc) and the packed text (hsp->packedDNA: 16 bases per word,
 * most significant first), md[mdOffsets[t] .. mdOffsets[t+1]) = the MD string, numMismatch /
 * gapOpen / gapExt as the reference counts them (gapExt counts every gap base), and
 * avgMismatchQual = (int)(sum of qualities at mismatching read offsets / numMismatch), 20 without
 * mismatches or without qualities.  qualities (signed chars as the reference's) of alignment t are
 * qualities[qualityOffsets[t] .. qualityOffsets[t+1]) in read order, or NULL.  An alignment with
 * an empty CIGAR gets an empty MD.  *md is malloc'ed by the library (s3_free); the four count
 * arrays may be NULL. */
int s3_dp_md(const uint32_t *packedDNA, uint64_t textLength, const char *cigars, const uint64_t *cigarOffsets,
             const uint32_t *positions, uint32_t numOfThreads,
             const int8_t *qualities, const uint64_t *qualityOffsets,
             uint64_t *mdOffsets, char **md, int32_t *numMismatch, int32_t *gapOpen, int32_t *gapExt,
             int32_t *avgMismatchQual);

/* ------------------------------------------------------------------------
 * Paired-end pairing of the two reads' occurrence lists, a batch of read pairs per call.
 * Replaces PEMappingOccurrences + PEStatsPEOutput as hostKernel calls them pair by pair once both
 * reads have hits (CPUfunctions.cpp:2281-2310; PEAlgnmt.cpp:114-291,480-637,777-838).  Lists are
 * CSR over read pairs: occurrences of read pair p's first read are pos1 / strand1 / mism1
 * [off1[p], off1[p+1]) in arrival order (text position, SRAOccurrence.strand 1 or 2,
 * mismatchCount), its second read's likewise in ...2; patternLengths[p] = the second read's
 * length (pe_in->patternLength, CPUfunctions.cpp:2284).  Per read pair, in the reference's
 * emission order (both lists stably sorted by position, merge walk, list 1 first on ties), one
 * record per valid pair: (*outPos1)[r], (*outPos2)[r] = algnmt_1 / algnmt_2 (always list 1 /
 * list 2), (*outInsertion)[r], (*outFlags)[4r..] = strand_1, mismatch_1, strand_2, mismatch_2,
 * for r in [pairOffsets[p], pairOffsets[p+1]); the insert-size test is the reference's unsigned
 * insertLbound <= rightPos + patternLength - leftPos <= insertUbound with the left / right leg
 * strands; reportOne = PE_REPORT_ONE.  optimal[p] / suboptimal[p] = index inside p's records of
 * PEStatsPEPairList's two pairs (0xFFFFFFFF: none), mismatchStats[32p + k] = records of p with k
 * mismatches in total.  pairOffsets has numPairs + 1 entries; the four record arrays are
 * malloc'ed by the library (s3_free).  The index handle only names the device and stream.
 * ------------------------------------------------------------------------ */
int s3_pair_occurrences(s3_index *ix,
                        const uint32_t *pos1, const uint8_t *strand1, const uint8_t *mism1, const uint64_t *off1,
                        const uint32_t *pos2, const uint8_t *strand2, const uint8_t *mism2, const uint64_t *off2,
                        const uint32_t *patternLengths, uint64_t numPairs,
                        int32_t insertLbound, int32_t insertUbound, int strandLeftLeg, int strandRightLeg,
                        int reportOne,
                        uint64_t *pairOffsets, uint32_t **outPos1, uint32_t **outPos2, uint32_t **outInsertion,
                        uint8_t **outFlags, uint32_t *optimal, uint32_t *suboptimal, uint32_t *mismatchStats);

/* ------------------------------------------------------------------------
 * Best-hit filters on every read's SA-range list and occurrence list, a batch of reads per call.
 * Replaces retainAllBest / retainAllBestWithCap / retainAllBestAndSecBest (SAList.cpp:140-348)
 * as hostKernel applies them read by read (CPUfunctions.cpp:2170-2255).  Lists are CSR over
 * reads, entries in arrival order: SA ranges saL / saR / saStrand / saMism [saOff[r], saOff[r+1])
 * (PESRAAlignmentResult) and occurrences occPos / occStrand / occMism likewise (SRAOccurrence).
 * Kept entries come back in the same order, CSR by outSaOff / outOccOff (numReads + 1 entries
 * each); out*Flags hold strand, mismatchCount per kept entry; num[r] = the function's return
 * value (occurrences retained).  Mode 1 cuts a range short (outSaR) or skips entries once maxNum
 * occurrences are in, in list order, exactly as the reference's running count does.  The output
 * arrays are the caller's and must hold as many entries as the input lists.
 * ------------------------------------------------------------------------ */
#define S3_RETAIN_ALL_BEST        0
#define S3_RETAIN_BEST_WITH_CAP   1
#define S3_RETAIN_BEST_AND_SECOND 2
int s3_retain_best(s3_index *ix, int mode, int32_t maxNum,
                   const uint32_t *saL, const uint32_t *saR, const uint8_t *saStrand, const uint8_t *saMism, const uint64_t *saOff,
                   const uint32_t *occPos, const uint8_t *occStrand, const uint8_t *occMism, const uint64_t *occOff,
                   uint64_t numReads,
                   uint64_t *outSaOff, uint32_t *outSaL, uint32_t *outSaR, uint8_t *outSaFlags,
                   uint64_t *outOccOff, uint32_t *outOccPos, uint8_t *outOccFlags, uint32_t *num);

/* ------------------------------------------------------------------------
 * A batch of read pairs from queries to alignments without leaving the device.  Replaces what soap3_dp_pair_align does
 * on the host between its GPU calls (alignment.cu:1896-2330) for one batch: round 1 of all_valid_alignment
 * (alignment.cu:855), hostKernel's per-pair work (CPUfunctions.cpp:1498-2620: collect_all_answers :1226, the routing of a
 * pair by which mates have hits :2153-2262 / :2440-2530, transferAllSAToOcc SAList.cpp:392, PEMappingOccurrences +
 * PEStatsPEOutput PEAlgnmt.cpp:480,777) and the default-DP engine's mate rescue (HalfEndOccStream, HalfEndAlgnBatch::pack
 * DV-DPfunctions.cu:1900,2027; performAlignment :669; the result loop :2359-2420) -- the semantics of the empty
 * alignPairR (soap3-dp-module.h:60, soap3-dp-module.cu:183-193) with in-memory results.  Reads 2p and 2p + 1 are the two
 * mates of pair p (the reference's interleaved pair layout).  Queries go in, a few dozen bytes per pair come out:
 *   route[p]   what became of pair p (codes below)
 *   pairs[p]   S3_PE_PAIRED: PEStatsPEPairList's optimal pair (algnmt_1/2, strands, mismatches, insertion), the number
 *              of valid pairs, how many share the optimal / the suboptimal total
 *   dp[t]      one record per rescue window in HalfEndOccStream's order (reads ascending, a read's SA ranges in slot
 *              order, a range in suffix-array order): the aligned mate's occurrence, the DP read, its position
 *              (window start + hitLoc, 0xFFFFFFFF below the cutoff), score, maxScoreCounts, and its CIGAR as
 *              runs[runOffset .. runOffset + numRuns) in read order, each length << 8 | op with op one of M m I D S
 *              (CigarStringEncoder's special CIGAR, DV-DPfunctions.h:545-597)
 * Reads whose round-1 slot overflowed in some case are not searched again here: their pair is S3_PE_OVERFLOW and goes
 * to s3_search_round2 / s3_search like the reference's round 2.  Pairs the reference hands to its later stages are
 * named, not processed: S3_PE_NONE (both-unaligned list: deep DP), *_TOO_MANY (new-default DP), BOTH_NO_PAIR_MANY.
 * The host arrays of s3_pe_align's result are the library's (pinned) and stay valid until the next call on the handle.
 * ------------------------------------------------------------------------ */
#define S3_PE_NONE              0   /* no mate has a hit: addReadIDToBothUnalignedPairs */
#define S3_PE_PAIRED            1   /* both mates hit and a valid pair exists */
#define S3_PE_BOTH_HIT          1   /* (inside the chain: both mates hit, pairing to come) */
#define S3_PE_FIRST_RESCUES     2   /* only the first read hit: its occurrences rescue the mate (dpInput) */
#define S3_PE_SECOND_RESCUES    3
#define S3_PE_BOTH_RESCUE       4   /* both hit, no valid pair, few hits each: each rescues the other (dpInput) */
#define S3_PE_FIRST_TOO_MANY    5   /* only the first read hit, more than maxHitNumForDP best hits: dpInputForNewDefault */
#define S3_PE_SECOND_TOO_MANY   6
#define S3_PE_BOTH_NO_PAIR_MANY 7   /* both hit, no valid pair, one of them with more than maxHitNumForDP hits */
#define S3_PE_OVERFLOW          8   /* a round-1 slot of one of the mates overflowed (isMoreThanSA1) */

typedef struct s3_pe s3_pe;
typedef struct {
    uint32_t numMismatch;                 /* Soap3MisMatchAllow: 2 when DP follows (SOAP3-DP.cu:210-213) */
    int32_t insertLow, insertHigh;        /* -v / -u */
    int32_t strandLeftLeg, strandRightLeg;/* 1 / 2: StrandArrangement +/- */
    uint32_t maxOutputPerRead;            /* soap3-dp.ini MaxOutputPerRead (1000) */
    uint32_t maxHitNumForDP;              /* getParameterForDefaultDP maxHitNum (CPUfunctions.cpp:59-86) */
    int32_t keepSecondBest;               /* needOutputMAPQ: retainAllBestAndSecBest instead of retainAllBest */
    s3_dp_scores scores;                  /* soap3-dp.ini [DP] */
    int32_t cutoffThreshold;              /* DPScoreThreshold; < 0: DEFAULT = ceil(0.3 * read length) */
    int32_t softClipLeft, softClipRight;  /* MaxFrontLenClipped / MaxEndLenClipped */
    uint32_t maxWindows;                  /* rescue windows per batch the DP workspace is sized for; 0: one per read */
    int32_t readStats;                    /* 1: also return, per read, the mismatch statistics hostKernel keeps for MAPQ (readStats below) */
} s3_pe_params;
typedef struct {
    uint32_t pos1, pos2, insertion;       /* PEPairs algnmt_1, algnmt_2, insertion */
    uint8_t strand1, mism1, strand2, mism2;
    uint32_t numPairs;                    /* numPEAlgnmt */
    uint32_t numOptimal, numSuboptimal;   /* num_minMismatch, num_soMinMismatch (CPUfunctions.cpp:2311-2320) */
    int8_t optimalTotal, suboptimalTotal; /* totalMismatchCount of the two pairs of PEStatsPEPairList; 127: none */
    uint16_t pad;
} s3_pe_pair_result;
typedef struct {
    uint32_t x0, x1;                      /* occurrences with the fewest mismatches / with one more: first_X0 / first_X1 of hostKernel */
    uint8_t minMismatch, pad[3];          /* previousMinNumMismatch (CPUfunctions.cpp:2061-2141; rOutput->WithError of collect_all_answers); 255: no occurrence */
} s3_pe_read_stats;
typedef struct {
    uint32_t dpReadID, alignedPos, dpPos;
    int32_t score;
    uint32_t numSameScore, runOffset;
    uint16_t numRuns;
    uint8_t alignedStrand, alignedMismatches, dpStrand, leftOrRight, pad[2];
} s3_pe_dp_result;
typedef struct {
    uint64_t numPairs, numRanges, numOccurrences, numWindows, numRuns;
    uint32_t routeCounts[16];             /* pairs per first-stage route code */
    uint64_t h2dBytes, d2hBytes;          /* what crossed the link for this batch */
    uint8_t *route; s3_pe_pair_result *pairs; s3_pe_dp_result *dp; uint32_t *runs;              /* host (s3_pe_align) */
    uint8_t *d_route; s3_pe_pair_result *d_pairs; s3_pe_dp_result *d_dp; uint32_t *d_runs;      /* device (s3_pe_align_device) */
    s3_pe_read_stats *readStats, *d_readStats;                                                  /* per read, with params.readStats (host / device) */
} s3_pe_result;
/* maxReads (even) reads of up to maxReadLength bases per batch; needs an index uploaded with suffix array and text */
int s3_pe_create(s3_index *ix, uint32_t maxReads, uint32_t maxReadLength, const s3_pe_params *params, s3_pe **out);
void s3_pe_free(s3_pe *pe);
int s3_pe_align(s3_pe *pe, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery,
                s3_pe_result *out);
/* Starts the upload of the NEXT batch's queries on a stream of its own and returns at once; the s3_pe_align call that
 * names the same host arrays finds them on the device.  Called before s3_pe_align of the current batch, the copy runs
 * under that batch's kernels (the reference double-buffers its batches the same way, alignment.cu:555,1030).  The host
 * arrays must stay untouched until that s3_pe_align returns (pinned memory for the copy to be asynchronous). */
int s3_pe_prefetch(s3_pe *pe, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery);
/* queries / readLengths already in device memory, results left there (two 4-byte counts are all that crosses the link) */
int s3_pe_align_device(s3_pe *pe, const uint32_t *d_queries, const uint32_t *d_readLengths, uint64_t numReads,
                       uint32_t wordPerQuery, s3_pe_result *out);
/* milliseconds per stage summed since timing was switched on: 0 search, 1 collect + route, 2 locate + sort,
 * 3 pairing + window count, 4 window fill, 5 DP, 6 CIGAR runs + records */
int s3_pe_set_timing(s3_pe *pe, int on);
int s3_pe_read_timing(s3_pe *pe, float *msPerStage);
s3_dp *s3_pe_dp(s3_pe *pe);               /* the chain's DP workspace (for s3_dp_set_timing / s3_dp_read_timing) */

/* ------------------------------------------------------------------------
 * A batch of single-end reads from queries to occurrences on the device: search (round-1 slots, all cases of the
 * numMismatch scheme) -> collect -> best hits (optional) -> locate.  The in-memory alignSingleR (soap3-dp-module.h:66-74,
 * soap3-dp-module.cu:62-181: soap3_dp_single_align with outputFileName == NULL, hostKernel storing occRec records,
 * CPUfunctions.cpp:1887-1905): occurrences of read r are positions / occFlags [occOffsets[r], occOffsets[r+1]) in the
 * order collect_all_answers (CPUfunctions.cpp:1226-1300) and transferAllSAToOcc (SAList.cpp:392-419) leave them -- cases
 * ascending, slots in order, suffix-array order inside a range -- capped at maxOutputPerRead; occFlags hold strand (1 / 2)
 * and mismatches per occurrence (occRec.strand; occRec.source is 1, occRec.score the mismatch count).  reportBest keeps
 * the ranges with the fewest mismatches only (retainAllBest, SAList.cpp:140-207).  readFlags[r] bit 0: a round-1 slot of
 * the read overflowed in some case (isMoreThanSA1: the reference searches it again, s3_search_round2 / s3_search).
 * Host arrays are the library's (pinned), valid until the next call on the handle.
 * ------------------------------------------------------------------------ */
typedef struct s3_se s3_se;
typedef struct {
    uint32_t numMismatch;                 /* 0..4 */
    uint32_t maxOutputPerRead;            /* soap3-dp.ini MaxOutputPerRead */
    int32_t reportBest;                   /* 0: all valid alignments, 1: all best */
    /* long reads (single_end_alignment, alignment.cu:2475-2491; hostKernel, CPUfunctions.cpp:1812-1842): with longReadMode the
     * search aligns the first S3_LONG_READ_SEED_LEN bases of every read longer than S3_LONG_READ_LEN, and each occurrence is
     * then extended over the rest of the read by s3_validate_alignments' rule; the list is cut to maxOutputPerRead.  The other
     * three fields are validateAlignments' only_keep_best_ans, min_seed_mismatch_allowed and needOutputMAPQ (allowance x 2). */
    int32_t longReadMode, onlyKeepBest, minSeedMismatch, doubleAllowance;
} s3_se_params;
typedef struct {
    uint64_t numReads, numRanges, numOccurrences, h2dBytes, d2hBytes;
    uint32_t *occOffsets, *positions; uint8_t *occFlags, *readFlags;                  /* host (s3_se_align) */
    uint32_t *d_occOffsets, *d_positions; uint8_t *d_occFlags, *d_readFlags;          /* device (s3_se_align_device) */
} s3_se_result;
#define S3_LONG_READ_LEN 120u             /* definitions.h:140 */
#define S3_LONG_READ_SEED_LEN 100u        /* definitions.h:141 */
int s3_se_create(s3_index *ix, uint32_t maxReads, const s3_se_params *params, s3_se **out);
void s3_se_free(s3_se *se);
int s3_se_align(s3_se *se, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery, s3_se_result *out);
int s3_se_align_device(s3_se *se, const uint32_t *d_queries, const uint32_t *d_readLengths, uint64_t numReads, uint32_t wordPerQuery,
                       s3_se_result *out);

/* ------------------------------------------------------------------------
 * Long reads: ungapped extension of seed alignments.  Replaces validateAlignments (CPUfunctions.cpp:1129-1222, with
 * createQueryPackedDNA / createRevQueryPackedDNA / createTargetPackedDNA / numMismatchNew, PE.cpp:28-60,148-178,287-325) as
 * hostKernel applies it to the occurrence list of every read in long-read mode (CPUfunctions.cpp:1812-1842).  Host arrays:
 * the occurrences of read r are positions / occFlags (strand, mismatches of the first seedLen bases) [occOffsets[r],
 * occOffsets[r + 1]); seedLen = S3_LONG_READ_SEED_LEN for reads longer than S3_LONG_READ_LEN, else the read length (nothing to
 * extend: the list is only cut to maxHitNum).  An occurrence stays when seed mismatches + Hamming distance of the remaining
 * bases <= ceil(0.02 * readLen) (twice that with doubleAllowance); a reverse-strand occurrence moves to the start of the
 * whole read; entries with fewer than minSeedMismatch seed mismatches are skipped; onlyKeepBest keeps the running best (the
 * output restarts when a better total turns up) and the walk stops once maxHitNum entries with minSeedMismatch in total are
 * out.  In place: the outCounts[r] kept entries of read r are at the front of its list.  Needs the packed text on the device.
 * ------------------------------------------------------------------------ */
int s3_validate_alignments(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery,
                           const uint32_t *occOffsets, uint32_t *positions, uint8_t *occFlags, int onlyKeepBest, int minSeedMismatch,
                           int doubleAllowance, int maxHitNum, uint32_t *outCounts);

/* ------------------------------------------------------------------------
 * Tables of the DP stages (host, integer).  s3_seed_layout replaces getSeedPositions
 * (definitions.h:323-442): the seed length and the 0-based seed offsets of a read of readLength
 * bases in a seeding stage -- what a caller cuts out of its reads before s3_search and hands to
 * s3_seed_candidates as seedOffsets / seedLengths.  Stages as definitions.h:317-321; stage 2
 * (default DP = mate rescue) has no seeds.  Up to `capacity` offsets are written; a deep-DP read
 * shorter than its seed gets *seedNum = 0 (the reference reads before its array there).
 * s3_dp_stage_parameters replaces getParameterFor{SingleDP,DefaultDP,NewDefaultDP,DeepDP}
 * (CPUfunctions.cpp:59-260; stage 5 = deep DP with round 2's maxHitNum,
 * DV-DPForBothUnalign.cu:138-139): cutoffThreshold = ceil(0.3 * readLength) when
 * isDefaultThreshold == 1 (soap3-dp.ini DPScoreThreshold = DEFAULT), else dpScoreThreshold; the
 * per-read maxHitNum / seedLength / sampleDist; the clip limits passed through.  Fields a stage
 * does not set are 0.  readLength2 is the mate's length (ignored by stage 1).
 * ------------------------------------------------------------------------ */
#define S3_STAGE_SINGLE_DP      1
#define S3_STAGE_DEFAULT_DP     2
#define S3_STAGE_NEW_DEFAULT_DP 3
#define S3_STAGE_DEEP_DP_ROUND1 4
#define S3_STAGE_DEEP_DP_ROUND2 5
int s3_seed_layout(int stage, int32_t readLength, int32_t *seedLength, int32_t *seedPositions,
                   int32_t capacity, int32_t *seedNum);
typedef struct {
    int32_t cutoffThreshold, maxHitNum, sampleDist, seedLength;        /* DPParam, PEAlgnmt.h:341-347 */
} s3_dp_read_params;
typedef struct {
    int32_t softClipLeft, softClipRight, tailTrimLen, singleDPSeedNum, singleDPSeedPos[10];
    s3_dp_read_params paramRead[2];                                    /* DPParameters, PEAlgnmt.h:349-364 */
} s3_dp_stage_params;
int s3_dp_stage_parameters(int stage, uint32_t readLength, uint32_t readLength2, int isDefaultThreshold,
                           int32_t dpScoreThreshold, int32_t maxFrontLenClipped, int32_t maxEndLenClipped,
                           s3_dp_stage_params *out);

/* ------------------------------------------------------------------------
 * The two DP stages that start from seeds, for a batch.  Orchestration of the entries above, like the reference's wrappers:
 *   s3_single_dp_align  replaces DPForUnalignSingle2 (DV-DPForSingleReads.cu:155; SingleDPWrapper::run :141; the engines
 *                       SingleEndSeedingEngine / SingleEndAlignmentEngine, DV-DPfunctions.cu:1026-1780) for the reads
 *                       readIDs[0..n): seeds (s3_seed_layout, stage 1) -> s3_seed_search -> s3_seed_candidates ->
 *                       s3_dp_make_windows (S3_WIN_SINGLE) -> s3_dp_align_windows -> one s3_dp_hit per candidate that
 *                       reaches its cutoff (SingleAlgnmtResult, DV-DPfunctions.cu:1699-1733), in candidate order
 *   s3_deep_dp_align    replaces DPForUnalignPairs2 (DV-DPForBothUnalign.cu:245; DeepDPWrapper::run2 :226, seeding_ext
 *                       :131-143; PairEndSeedingEngine / PairEndAlignmentEngine, DV-DPfunctions.cu:2571-3800) for the pairs
 *                       whose even read ids are pairReadIDs[0..n): seeds of both mates (stage 4; stage 5 with its larger hit
 *                       limit for the pairs that found no candidate and had a seed with too many hits) -> s3_seed_search per
 *                       side -> s3_seed_pair_candidates -> left window, DP, right window cut by the left hit, DP -> one
 *                       s3_deep_dp_hit per candidate whose two reads reach their cutoffs (DeepDPAlignResult, :3755-3795)
 * positions are window start + hitLoc; CIGARs are runs[runOffset .. + numRuns) in read order, each length << 8 | op (the
 * special CIGAR of CigarStringEncoder).  unseeded lists the reads / pairs without any candidate (the reference's
 * unseededIDStream).  cutoff: isDefaultThreshold ? ceil(0.3 * read length) : dpScoreThreshold.  The arrays are malloc'ed by
 * the library (s3_single_dp_result_free / s3_deep_dp_result_free).
 * ------------------------------------------------------------------------ */
typedef struct {
    int32_t insertLow, insertHigh, strandLeftLeg, strandRightLeg;     /* deep DP only */
    s3_dp_scores scores;
    int32_t isDefaultThreshold, dpScoreThreshold;                     /* soap3-dp.ini DPScoreThreshold */
    int32_t softClipLeft, softClipRight;
} s3_stage_params;
typedef struct {
    uint32_t readID, pos;
    int32_t score;
    uint32_t numSameScore, runOffset;
    uint16_t numRuns;
    uint8_t strand, pad;
} s3_dp_hit;
typedef struct {
    uint64_t numReads, numSeeds, numCandidates, numHits, numRuns, numUnseeded;
    s3_dp_hit *hits; uint32_t *runs, *unseeded;
} s3_single_dp_result;
typedef struct {
    uint32_t readID;                      /* the pair's even read id */
    uint32_t pos1, pos2;                  /* algnmt_1 / algnmt_2: of the pair's first read and of its mate */
    int32_t score1, score2;
    uint32_t numSame1, numSame2, runOffset1, runOffset2;
    uint16_t numRuns1, numRuns2;
    uint8_t strand1, strand2, pad[2];
} s3_deep_dp_hit;
typedef struct {
    uint64_t numPairs, numSeeds, numCandidates, numHits, numRuns, numUnseeded;
    s3_deep_dp_hit *hits; uint32_t *runs, *unseeded;
} s3_deep_dp_result;
int s3_single_dp_align(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery,
                       const uint32_t *readIDs, uint64_t n, const s3_stage_params *par, s3_single_dp_result *out);
void s3_single_dp_result_free(s3_single_dp_result *r);
int s3_deep_dp_align(s3_index *ix, const uint32_t *queries, const uint32_t *readLengths, uint64_t numReads, uint32_t wordPerQuery,
                     const uint32_t *pairReadIDs, uint64_t n, const s3_stage_params *par, s3_deep_dp_result *out);
/* The same stage for the batch a paired-end chain has just aligned (the call DPForUnalignPairs2 follows semiGlobalDP2 with,
 * soap3-dp-module.cu / SOAP3-DP.cu main loop): the pairs s3_pe_align[_device] left as S3_PE_NONE are picked on the device and the stage
 * works on the chain's own query buffer and read lengths -- nothing is uploaded again.  To be called before the next s3_pe_prefetch /
 * s3_pe_align on the handle.  Results as s3_deep_dp_align (readIDs are the even read ids of the batch); numPairs = the pairs picked. */
int s3_pe_deep_dp(s3_pe *pe, const s3_stage_params *par, s3_deep_dp_result *out);
void s3_deep_dp_result_free(s3_deep_dp_result *r);

/* ------------------------------------------------------------------------
 * Mapping qualities (host, scalar).  Replace the MAPQ functions of the SAM writers with their
 * tables (BGS-IO.cpp:33-45, 2280-2580; g_log_n of bwase_initialize, CPUfunctions.cpp:3014-3019
 * is built in).  Argument lists are the reference's, without the g_log_n pointer:
 *   s3_mapq_unique       getMapQualScore            :2280     s3_mapq_bwa_single   bwaLikeSingleQualScore :2311
 *   s3_mapq_single       getMapQualScoreSingle      :2331     s3_mapq_single_dp    getMapQualScoreForSingleDP :2370
 *   s3_mapq_bwa_pair     bwaLikePairQualScore       :2415     s3_mapq_pair_end     getMapQualScore2 :2465
 *   s3_mapq_unique_dp    getMapQualScoreForDP       :2500     s3_mapq_pair_end_dp  getMapQualScoreForDP2 :2534
 *   s3_mapq_of_pair      getMapQualScoreForPair     :2577
 * ------------------------------------------------------------------------ */
int32_t s3_mapq_unique(int n, int mismatchNum, int avgMismatchQual, int maxMAPQ, int minMAPQ);
int32_t s3_mapq_bwa_single(int x0, int x1);
int32_t s3_mapq_single(int mismatchNum, int avgMismatchQual, int x0, int x1, int maxMAPQ, int minMAPQ, int isBWALike);
int32_t s3_mapq_single_dp(int maxDPScore, int avgMismatchQual, int x0, int x1_t1, int x1_t2, int bestDPScore,
                          int secondBestDPScore, int maxMAPQ, int minMAPQ, int dpThres, int isBWALike);
void s3_mapq_bwa_pair(int x0_0, int x1_0, int x0_1, int x1_1, int op_score, int op_num, int subop_score, int subop_num,
                      int readlen_0, int readlen_1, int32_t *mapScore0, int32_t *mapScore1);
int32_t s3_mapq_pair_end(int mismatchNum, int avgMismatchQual, int x0, int x1, int isBestHit, uint32_t totalNumValidPairs,
                         int maxMAPQ, int minMAPQ);
int32_t s3_mapq_unique_dp(int n, int dpScore, int maxDPScore, int avgMismatchQual, int maxMAPQ, int minMAPQ);
int32_t s3_mapq_pair_end_dp(int dpScore, int maxDPScore, int avgMismatchQual, int x0, int x1, int bestDPScore,
                            int secondBestDPScore, int isBestHit, int totalNumValidPairs, int maxMAPQ, int minMAPQ);
int32_t s3_mapq_of_pair(int score1, int score2);

/* ------------------------------------------------------------------------
 * Measurement hooks (replace nothing in the reference).  With timing on, CUDA events are
 * recorded between the kernel launches of every call on the handle; read_timing waits for the
 * handle's stream, returns the milliseconds (and launches) per kernel slot since the last read
 * and forgets them.  msPerSlot / launchesPerSlot hold 8 entries.
 *   index slots: 0 straight-line search kernel, 1 enumerating kernel, 2 spine, 3 tasks, 4 merge, 5 isBad carry-over
 *   dp slots:    0 score sweep, 1 best cell, 2 traceback
 * ------------------------------------------------------------------------ */
/* The memory system's random-sector rate: loadsPerThread independent random 32-byte reads per thread over the index's
 * forward bucket array (the search's own access pattern), timed on the index stream.  *numLoads / *ms = sectors per
 * millisecond; what bench.py reports the search launch's executed sectors against. */
int s3_random_sector_probe(s3_index *ix, uint32_t loadsPerThread, float *ms, uint64_t *numLoads);
/* L2 access-policy window over one of the index arrays for the kernels on the index stream (measurement knob; region 0
 * resets): 1 forward buckets, 2 reverse buckets, 3 seed table fwd1, 4 seed table rev0, 5 packed text, 6 suffix array.
 * windowBytes 0 = the largest window the device allows, persistBytes 0 = its largest persisting carve-out.  Answers do
 * not depend on it. */
int s3_index_set_l2_persist(s3_index *ix, int region, size_t windowBytes, size_t persistBytes);
int s3_index_set_timing(s3_index *ix, int on);
int s3_index_read_timing(s3_index *ix, float *msPerSlot, int *launchesPerSlot);
int s3_dp_set_timing(s3_dp *dp, int on);
int s3_dp_read_timing(s3_dp *dp, float *msPerSlot, int *launchesPerSlot);

/* ------------------------------------------------------------------------
 * SAM records (host).  s3_sam_pair_records replaces pairOutputSAMAPI (BGS-IO.cpp:3478-3793) for one read pair: the two
 * records samwrite would receive -- bam1_t.core fields and bam1_t.data (name, CIGAR, 4-bit bases, qualities, tags RG NM X0 X1
 * XM XO XG MD XA) byte for byte -- from the pair's valid pairings (PEPairs in list order, bestIndex = the one reported, < 0:
 * none, both reads come out unmapped), the reads, and the counts hostKernel passes (CPUfunctions.cpp:2281-2380): the optimal /
 * suboptimal total mismatches (PEStatsPEOutput), X0 / X1 of either read, the number of optimal pairings, whether the reported
 * ends are best hits of their reads, the number of valid pairings.  Included: position -> chromosome through the translate
 * table (getChrAndPos, :1746), trimming of an alignment that hangs over a chromosome / segment end (BoundaryCheck, :1779: CIGAR
 * <a>M<b>S or <a>S<b>M, MAPQ 0, mismatches recounted), the MD string and mean mismatch quality (getMdStr, PE.cpp:374), MAPQ
 * (s3_mapq_bwa_pair, or s3_mapq_pair_end + s3_mapq_of_pair; 255 unless the report type is all-valid / all-best), the XA:Z list
 * of the other pairings with the best total (up to peMaxOutputPerPair results).  record.data is malloc'ed
 * (s3_sam_record_free).  alignmentType: OUTPUT_* of definitions.h:127-130.
 * ------------------------------------------------------------------------ */
typedef struct { uint32_t startPos, chrID, correction; } s3_sam_segment;            /* Translate, 2bwt-lib/HSP.h:73-77 */
typedef struct {
    const uint32_t *packedDNA; uint32_t dnaLength;                                    /* hsp->packedDNA, dnaLength */
    const s3_sam_segment *segments; uint32_t numSegments;                             /* hsp->translate, numOfRemovedSegment */
    const uint32_t *ambiguityMap;                                                     /* hsp->ambiguityMap: per 2^18 positions a segment at or after theirs */
    const uint32_t *chrEndPos; uint32_t numChr; const char *const *chrNames;          /* hsp->seqOffset[].endPos; header->target_name */
} s3_sam_genome;
typedef struct {
    int32_t alignmentType, bwaLikeScore, dpMatchScore, dpMisMatchScore, isFastq, maxMAPQ, minMAPQ, isPrintMDNM, outputXAZTag;   /* HSPAux.h */
    uint32_t peMaxOutputPerPair;
    const char *readGroup;
} s3_sam_config;
typedef struct {
    uint32_t algnmt1, algnmt2;                                                        /* PEPairs, PEAlgnmt.h:164-178 */
    uint8_t strand1, mismatch1, strand2, mismatch2;
    int8_t totalMismatchCount; uint8_t pad[3];
} s3_sam_pairing;
typedef struct {
    int32_t tid, pos;                                                                 /* bam1_core_t, samtools-0.1.18/bam.h:169-178 */
    uint16_t bin; uint8_t qual, l_qname; uint16_t flag, n_cigar;
    int32_t l_qseq, mtid, mpos, isize;
    int32_t l_aux, data_len;                                                          /* bam1_t */
    uint8_t *data;
} s3_sam_record;
int s3_sam_pair_records(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_pairing *pairs, uint32_t numPairs, int32_t bestIndex,
                        const uint8_t *query1, const uint8_t *query2, const char *qualities1, const char *qualities2,
                        int32_t readlen1, int32_t readlen2, const char *queryName1, const char *queryName2,
                        int32_t minTotalMismatch, int32_t secMinTotalMismatch, int32_t x0First, int32_t x0Second, int32_t x1First, int32_t x1Second,
                        int32_t numMinMismatchPair, int32_t isBestHit1, int32_t isBestHit2, uint32_t totalNumValidPairs, s3_sam_record out[2]);
void s3_sam_record_free(s3_sam_record *record);
/* OCCOutputSAMAPI (BGS-IO.cpp:5556-5772): a single read's record from its occurrence list (s3_se_align's results): the first
 * occurrence with the fewest mismatches is reported (X0 = how many share that count; one that hangs over a chromosome /
 * segment end gives way to the best one that does not), the others are listed in XA:Z (alignmentType OUTPUT_ALL_BEST: only
 * those with the best count; never one that would be trimmed), X1 = the listed ones with more mismatches, MAPQ =
 * s3_mapq_single; numOcc == 0: the unmapped record.  (The reference leaves the XA buffer of the PREVIOUS read in place when the
 * reported occurrence is trimmed; calls here are independent, the list is empty then.) */
typedef struct { uint32_t ambPosition; uint8_t strand, mismatchCount, pad[2]; } s3_sam_occurrence;      /* SRAOccurrence, SRACore.h:86-95 */
int s3_sam_single_record(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_occurrence *occ, uint32_t numOcc,
                         const uint8_t *query, const char *qualities, int32_t readlen, const char *queryName, s3_sam_record *out);
/* SingleDPOutputSAMAPI (BGS-IO.cpp:5857-6118): a single read's record from its DP alignments (the hits of s3_single_dp_align with
 * the special CIGARs and edit distances of s3_dp_decode): the first alignment with the best score is reported (X0 = how many share
 * it), its CIGAR through convertToCigarStr, MD / XM / XO / XG / NM through getMisInfoForDP (s3_dp_md), MAPQ = s3_mapq_single_dp with
 * x1_t1 / x1_t2 = the listed suboptimal alignments at or above / below 0.7 x the best score; the others go to XA:Z (alignmentType
 * OUTPUT_ALL_VALID / OUTPUT_ALL_BEST; chromosome, strand + position, CIGAR, edit distance).  An alignment with gaps that hangs over
 * a chromosome / segment end is cut there (BoundaryCheckDP, :1807-1976): the longer side stays, the other becomes a soft clip, and
 * such a record carries no XA:Z.  numResult == 0 or ambPosition 0xFFFFFFFF in the first entry: the unmapped record. */
typedef struct { uint32_t ambPosition; uint8_t strand, pad[3]; int32_t score, editdist; const char *cigar; } s3_sam_dp_alignment;   /* SingleAlgnmtResult, PEAlgnmt.h:434-445 */
int s3_sam_single_dp_record(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_dp_alignment *alignments, uint32_t numResult,
                            int32_t singleDPcutoffThreshold, const uint8_t *query, const char *qualities, int32_t readlen, const char *queryName,
                            s3_sam_record *out);
/* pairDeepDPOutputSAMAPI (BGS-IO.cpp:3824-4500): the two records of a read pair from its deep-DP alignments (the hits of
 * s3_deep_dp_align / s3_pe_deep_dp with the special CIGARs and edit distances of s3_dp_decode; entry bestIndex is the one reported,
 * < 0: both reads unmapped): per read the CIGAR, MD and XM / XO / XG / NM of its alignment (cut at a chromosome / segment end like
 * s3_sam_single_dp_record), X0 / X1 from the scores and tie counts of the pair's other alignments and from the counts the search
 * left for the read (x0 / x1 / mismatch = hspaux->x0_array / x1_array / mismatch_array of the two reads), MAPQ = s3_mapq_bwa_pair or
 * s3_mapq_pair_end_dp + s3_mapq_of_pair, XA:Z with the other alignments; a pair whose reads run through each other keeps only the
 * read with fewer mismatches (the other comes out unmapped, MAPQ = s3_mapq_unique_dp).  A reported entry without any alignment
 * belongs to the writer of improperly paired reads (unproperlypairOutputSAMAPI2), which is not built: S3_EINVAL. */
typedef struct {
    int32_t insertSize;                                                               /* DeepDPAlignResult, PEAlgnmt.h:409-429; [0] the pair's first read, [1] its mate */
    uint32_t ambPosition[2]; uint8_t strand[2], pad[2];
    int32_t score[2], editdist[2], numSameScore[2];
    const char *cigar[2];
} s3_sam_deep_alignment;
int s3_sam_deep_dp_records(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_deep_alignment *alignments, uint32_t num, int32_t bestIndex,
                           const uint8_t *query1, const uint8_t *query2, const char *qualities1, const char *qualities2,
                           int32_t readlen1, int32_t readlen2, const char *queryName1, const char *queryName2,
                           const int32_t x0[2], const int32_t x1[2], const int32_t mismatch[2], s3_sam_record out[2]);
/* pairDPOutputSAMAPI (BGS-IO.cpp:4504-5554): the two records of a read pair from its default-DP (mate rescue) results -- the rescue
 * records of s3_pe_align with the special CIGARs and edit distances of s3_dp_decode.  In every entry one read comes from the search
 * (score = its mismatches, ungapped: CIGAR <len>M or the trimmed form, MD by getMdStr, MAPQ leg s3_mapq_pair_end) and the other
 * from DP (score = DP score; CIGAR, MD, XM / XO / XG by getMisInfoForDP, MAPQ leg s3_mapq_pair_end_dp); whichFromDP says which
 * (0 the first read, 1 its mate), and only the entries of the reported entry's kind are counted and listed.  X0 / X1 per read from
 * those entries and the search's counts (x0 / x1 / mismatch), the "second best" pair for s3_mapq_bwa_pair as the reference's scan
 * leaves it, read-through pairs as in s3_sam_deep_dp_records (the single read left: s3_mapq_unique_dp or s3_mapq_unique).  With a
 * reported entry whose DP side missed its cutoff (whichFromDP 2) the other entries of that kind are listed in XA:Z only where they
 * have a position for the read (the reference translates 0xFFFFFFFF through tables that do not reach it there). */
typedef struct {
    uint8_t whichFromDP, strand[2], pad;                                              /* AlgnmtDPResult, PEAlgnmt.h:384-402; [0] the pair's first read, [1] its mate */
    int32_t editdist, insertSize, numSameScore;
    uint32_t ambPosition[2];
    int32_t score[2];
    const char *cigar;                                                                /* special CIGAR of the read that came from DP */
} s3_sam_dp_pairing;
int s3_sam_pair_dp_records(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_dp_pairing *alignments, uint32_t num, int32_t bestIndex,
                           const uint8_t *query1, const uint8_t *query2, const char *qualities1, const char *qualities2,
                           int32_t readlen1, int32_t readlen2, const char *queryName1, const char *queryName2,
                           const int32_t x0[2], const int32_t x1[2], const int32_t mismatch[2], s3_sam_record out[2]);
/* unproperlypairOutputSAMAPI (BGS-IO.cpp:2582-2930): the two records of a read pair whose reads have occurrences but no valid pairing
 * (and of pairs where only one read, or neither, has any): each read reported on its own -- the first occurrence with the fewest
 * mismatches, X0 / X1 = the occurrences with that count / with one more, the others in XA:Z (alignmentType OUTPUT_ALL_VALID /
 * OUTPUT_ALL_BEST, up to peMaxOutputPerRead results), MAPQ = s3_mapq_single >> 1, at least minMAPQ (255 for the unique / random
 * report types; OUTPUT_UNIQUE_BEST reports a read only when its best count is held by one occurrence) -- flags without the
 * proper-pair bit, the mate's chromosome and position in the mate fields, an insert size when both lie on one chromosome. */
int s3_sam_unpaired_records(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_occurrence *occ1, uint32_t numOcc1,
                            const s3_sam_occurrence *occ2, uint32_t numOcc2, uint32_t peMaxOutputPerRead,
                            const uint8_t *query1, const uint8_t *query2, const char *qualities1, const char *qualities2,
                            int32_t readlen1, int32_t readlen2, const char *queryName1, const char *queryName2, s3_sam_record out[2]);
/* unproperlypairDPOutputSAMAPI (BGS-IO.cpp:2932-3447): the same for a pair whose reads went through DP without ending up properly
 * paired -- per read a list of alignments, each from DP (special CIGAR; handled like s3_sam_single_dp_record's) or from the search
 * (isFromDP 0: editdist = its mismatches, cigar = "<len>M"; handled like an occurrence): the first alignment with the best score is
 * reported, X0 = how many share it, X1 = the rest (x1_t1 + x1_t2), MAPQ = s3_mapq_single_dp halved unless BWA-like, at least minMAPQ,
 * 0 when trimmed; XA:Z with CIGARs and edit distances; the insert size discounts a trailing deletion as convertToCigarStr reports it. */
typedef struct { uint32_t ambPosition; uint8_t strand, isFromDP, pad[2]; int32_t score, editdist; const char *cigar; } s3_sam_read_alignment;   /* Algnmt, PEAlgnmt.h:468-479 */
int s3_sam_unpaired_dp_records(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_read_alignment *alignments1, uint32_t num1,
                               const s3_sam_read_alignment *alignments2, uint32_t num2, int32_t singleDPcutoffThreshold,
                               const uint8_t *query1, const uint8_t *query2, const char *qualities1, const char *qualities2,
                               int32_t readlen1, int32_t readlen2, const char *queryName1, const char *queryName2, s3_sam_record out[2]);
/* SingleAnsOutputSAMAPI (BGS-IO.cpp:5774-5827): a single read reported with ONE alignment (the unique-best / random-best report types):
 * MD, XM / NM = its mismatches, X0 = bestHitNum (< 0: no X0 tag), no X1, MAPQ = s3_mapq_unique (255 when bestHitNum <= 0), no trimming.
 * (noAnsOutputSAMAPI, :5829-5855, the read without any alignment, is s3_sam_single_record with numOcc 0.) */
int s3_sam_single_answer_record(const s3_sam_genome *genome, const s3_sam_config *config, uint32_t ambPosition, int32_t strand, int32_t numMismatch, int32_t bestHitNum,
                                const uint8_t *query, const char *qualities, int32_t readlen, const char *queryName, s3_sam_record *out);
/* Which entry of a pair's results the reference reports (bestIndex of the two entries above): the scans of outputDeepDPResult2
 * (OutputDPResult.cpp:590-760: the first entry with the largest score1 + score2) and outputDPResult2 (:263-350: the fewest mismatches of
 * the read that came from the search, then the highest DP score, the first of equals).  -1 for an empty list. */
int32_t s3_sam_pick_deep_dp(const s3_sam_deep_alignment *alignments, uint32_t num);
int32_t s3_sam_pick_pair_dp(const s3_sam_dp_pairing *alignments, uint32_t num);
/* The text line samwrite prints for a record in a SAM file: bam_format1 (samtools-0.1.18/bam.c:243-329) -- QNAME FLAG RNAME POS MAPQ CIGAR
 * RNEXT ('=' on the same chromosome) PNEXT TLEN SEQ QUAL (+ 33) and the tags as TAG:TYPE:VALUE, tab-separated, no newline.  *line is
 * malloc'ed (s3_free). */
int s3_sam_format_line(const s3_sam_record *record, const char *const *chrNames, uint32_t numChr, char **line);
/* Whole batches as SAM text: the loops the reference's output threads run around the record writers, one record per read, printed as
 * s3_sam_format_line prints it, one line (+ '\n') per record in read order; *text is malloc'ed (s3_free), NUL-terminated, *textBytes
 * without the NUL.  Reads are independent: numThreads host threads (0: all of the machine) each take a slice of the batch.
 *   s3_sam_single_batch_text     the results of s3_se_align (occOffsets / positions / occFlags of s3_se_result) -> s3_sam_single_record per
 *                                read, the unmapped record for a read without an occurrence (hostKernel's SAM branch for single reads,
 *                                CPUfunctions.cpp:1862-1941 -> OCCOutputSAMAPI / noAnsOutputSAMAPI; with alignmentType OUTPUT_UNIQUE_BEST /
 *                                OUTPUT_RANDOM_BEST the first occurrence alone through s3_sam_single_answer_record, unique-best only when it
 *                                is the read's one occurrence): BASELINE config 2 from reads to text.  The reference writes the reads without
 *                                an occurrence after its last stage; here every read has its line in input order.
 *   s3_sam_single_dp_batch_text  the results of s3_single_dp_align (hits in candidate order, the candidates of a read next to each
 *                                other; runs) -> s3_runs_decode -> s3_sam_single_dp_record per read that has a hit
 *                                (outputDPSingleResult2, OutputDPResult.cpp:938-1058, with hspaux->singleDPcutoffThreshold,
 *                                alignment.cu:2360); reads without a hit are the caller's (the reference reports them after its last
 *                                stage).  Where the reference's DP batches cut a read's candidates in two, its loop writes the read
 *                                twice; there are no batch borders here. */
typedef struct {
    const uint8_t *bases;                 /* numReads rows of rowBytes: one base code (0..3) per byte, the read as sequenced */
    const char *qualities;                /* numReads rows of rowBytes: Phred values */
    uint32_t rowBytes;
    const uint32_t *readLengths;
    const char *const *names;
} s3_sam_reads;
int s3_sam_single_batch_text(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_reads *reads, uint64_t numReads,
                             const uint32_t *occOffsets, const uint32_t *positions, const uint8_t *occFlags, uint32_t numThreads,
                             char **text, uint64_t *textBytes);
int s3_sam_single_dp_batch_text(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_reads *reads, uint64_t numReads,
                                const s3_dp_hit *hits, uint64_t numHits, const uint32_t *runs, uint64_t numRuns, s3_dp_scores scores,
                                int32_t singleDPcutoffThreshold, uint32_t numThreads, char **text, uint64_t *textBytes);
/* The same for the DP results of read pairs: two lines per pair (the first read's, then its mate's), pairs in the order of their results;
 * readStats (s3_pe_result.readStats of the chain that aligned the batch, indexed by read id; NULL: zeros) gives the x0 / x1 / mismatch
 * counts the writers take from the search (hspaux->x0_array / x1_array / mismatch_array).
 *   s3_sam_deep_dp_batch_text    the hits of s3_deep_dp_align / s3_pe_deep_dp (the hits of a pair next to each other) -> per hit the
 *                                DeepDPAlignResult fields the engine derives (s3_runs_decode; insert size, DV-DPfunctions.cu:3810-3815) ->
 *                                s3_sam_pick_deep_dp -> s3_sam_deep_dp_records (outputDeepDPResult2, OutputDPResult.cpp:590-760)
 *   s3_sam_pair_dp_batch_text    the rescue records of s3_pe_align (dp / runs of s3_pe_result; the records of a pair next to each other)
 *                                -> per record the AlgnmtDPResult the default-DP engine builds (DV-DPfunctions.cu:2355-2440: whichFromDP =
 *                                the DP read's parity, or 2 with an unaligned DP side when it missed its cutoff; insert size :2388-2400) ->
 *                                s3_sam_pick_pair_dp -> s3_sam_pair_dp_records (outputDPResult2, OutputDPResult.cpp:263-420).  A pair none
 *                                of whose rescues succeeded gets no lines: it belongs to the writers of improperly paired reads
 *                                (s3_sam_unpaired_records over the reads' occurrence lists). */
/*   s3_sam_paired_batch_text     the pairs s3_pe_align paired (route S3_PE_PAIRED) with ONE valid pairing -> s3_sam_pair_records with the
 *                                chain's reported pairing, totals and per-read statistics (hostKernel's SAM branch, CPUfunctions.cpp:2281-2380
 *                                -> pairOutputSAMAPI; X0 / X1 per report type as hostKernel passes them: the reads' counts for all-valid /
 *                                all-best, 1 / none for unique-best, none for random-best); readStats is required.  Pairs with more valid pairings need the whole list for XA:Z
 *                                (s3_pair_occurrences) and get no lines here. */
int s3_sam_paired_batch_text(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_reads *reads, uint64_t numReads,
                             const uint8_t *route, const s3_pe_pair_result *pairs, uint64_t numPairs, const s3_pe_read_stats *readStats,
                             uint32_t numThreads, char **text, uint64_t *textBytes);
/*   s3_sam_unpaired_batch_text   pairs without a valid pairing, each read on its own from its occurrence list -> s3_sam_unpaired_records
 *                                (hostKernel's SAM branch, CPUfunctions.cpp:2546-2557 -> unproperlypairOutputSAMAPI, for runs without the DP stages).  The lists are a CSR over
 *                                ALL reads of the batch -- what s3_se_align returns for the same queries -- and pairIDs names the pairs to
 *                                write (pair p = reads 2p and 2p + 1), e.g. those of the chain's routes without a DP result. */
int s3_sam_unpaired_batch_text(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_reads *reads, uint64_t numReads,
                               const uint32_t *occOffsets, const uint32_t *positions, const uint8_t *occFlags,
                               const uint32_t *pairIDs, uint64_t numPairs, uint32_t peMaxOutputPerRead, uint32_t numThreads,
                               char **text, uint64_t *textBytes);
/*   s3_sam_unpaired_dp_batch_text  the pairs no stage paired, after DP: per read ONE list as the reference's AllHits holds it (PEAlgnmt.cpp:1033-1260) --
 *                                the read's hits of s3_single_dp_align when it has any (isFromDP 1; s3_runs_decode), else its occurrences from the
 *                                search (CSR as above; isFromDP 0, score = len x match + mismatches x mismatch score, CIGAR <len>M) -- and
 *                                s3_sam_unpaired_dp_records for the two lists (outputSingleResultForPairEnds, OutputDPResult.cpp:1062-1150). */
int s3_sam_unpaired_dp_batch_text(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_reads *reads, uint64_t numReads,
                                  const uint32_t *occOffsets, const uint32_t *positions, const uint8_t *occFlags,
                                  const s3_dp_hit *hits, uint64_t numHits, const uint32_t *runs, uint64_t numRuns, s3_dp_scores scores,
                                  int32_t singleDPcutoffThreshold, const uint32_t *pairIDs, uint64_t numPairs, uint32_t numThreads,
                                  char **text, uint64_t *textBytes);
int s3_sam_deep_dp_batch_text(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_reads *reads, uint64_t numReads,
                              const s3_deep_dp_hit *hits, uint64_t numHits, const uint32_t *runs, uint64_t numRuns, s3_dp_scores scores,
                              const s3_pe_read_stats *readStats, uint32_t numThreads, char **text, uint64_t *textBytes);
int s3_sam_pair_dp_batch_text(const s3_sam_genome *genome, const s3_sam_config *config, const s3_sam_reads *reads, uint64_t numReads,
                              const s3_pe_dp_result *dp, uint64_t numRecords, const uint32_t *runs, uint64_t numRuns, s3_dp_scores scores,
                              const s3_pe_read_stats *readStats, uint32_t numThreads, char **text, uint64_t *textBytes);

#ifdef __cplusplus
}
#endif
#endif /* SOAP3DP_B200_H */
