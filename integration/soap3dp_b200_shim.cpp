// soap3dp_b200_shim.cpp -- what a SOAP3-dp maintainer links INSTEAD of the device code in
// alignment.cu:27-531 (GPUINDEXUpload .. perform_round2_alignment_no_pipeline) and
// DV-DPfunctions.cu:35-741 (the two DP kernels and SemiGlobalAligner): the reference's own
// function and class definitions, with bodies that call libsoap3dp_b200.so through its C ABI
// (include/soap3dp_b200.h).  It is compiled against the reference's unmodified headers
// (oracle/build_ref.sh does that as a check); nothing else of SOAP3-dp changes.
//
// The four `uint *` the reference threads through its call graph (_bwt, _occ, _revBwt, _revOcc)
// become carriers of one opaque handle: _bwt holds the s3_index pointer, the others stay NULL.
#include "alignment.h"
#include "DV-DPfunctions.h"
#include "soap3dp_b200.h"

static void s3_die(const char *what)
{
    // the reference's own convention: message + exit(1) (alignment.cu:38-42, DV-DPfunctions.h:113-118)
    printf("%s FAILED .. %s\n", what, s3_last_error());
    exit(1);
}

static int s3_current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);            // the reference selects the device with cudaSetDevice (SOAP3-DP.cu:219-220)
    return dev;
}

// ---- alignment.cu:27 / :109 ---------------------------------------------------------------
void GPUINDEXUpload ( Soap3Index * index, uint ** _bwt, uint ** _occ, uint ** _revBwt, uint ** _revOcc )
{
    BWT * bwt = index->sraIndex->bwt;
    BWT * revBwt = index->sraIndex->rev_bwt;
    uint numOcc = ( bwt->textLength + GPU_OCC_INTERVAL - 1 ) / GPU_OCC_INTERVAL + 1;
    s3_index * ix = NULL;
    // With the packed text and the full suffix array (SaValueFreq = 1, soap3-dp-builder.ini:29: bwt->saValue holds
    // textLength + 1 rows, BWT.c:254-285) the search finishes single-suffix intervals by comparing the read with the text
    // (check and extend) and the seed tables' straight-line kernel runs: the path bench.py measures.  A sampled suffix
    // array (saInterval > 1) keeps the stepping search; answers are identical either way.
    HSP * hsp = index->sraIndex->hsp;
    const uint * packedDNA = hsp ? ( const uint * ) hsp->packedDNA : NULL;
    const uint * sa = ( bwt->saInterval == 1 ) ? ( const uint * ) bwt->saValue : NULL;
    if ( !packedDNA || !sa ) { packedDNA = NULL; sa = NULL; }
    if ( s3_index_upload ( bwt->bwtCode, index->gpu_occValue, revBwt->bwtCode, index->gpu_revOccValue,
                           numOcc, bwt->inverseSa0, revBwt->inverseSa0, bwt->textLength,
                           packedDNA, sa, s3_current_device (), &ix ) != S3_OK )
    { s3_die ( "GPUINDEXUpload" ); }
    *_bwt = ( uint * ) ix;
    *_occ = *_revBwt = *_revOcc = NULL;
}

void GPUINDEXFree ( uint * _bwt, uint * _occ, uint * _revBwt, uint * _revOcc )
{
    s3_index_free ( ( s3_index * ) _bwt );
}

// The reference keeps its query and answer buffers in plain malloc'ed memory for the whole run (alignment.cu:689-760) and
// hands the same pointers to every call: they are page-locked the first time they are seen, so that the copies of every
// later call run at the link's rate and overlap the kernels.  A buffer that cannot be registered is used as it is.
#include <map>
static std::map<const void *, size_t> s3_pinned_buffers;
static void s3_pin_once ( const void * p, size_t bytes )
{
    if ( !p || bytes == 0 ) { return; }
    std::map<const void *, size_t>::iterator it = s3_pinned_buffers.find ( p );
    if ( it != s3_pinned_buffers.end () && it->second >= bytes ) { return; }
    if ( it != s3_pinned_buffers.end () ) { cudaHostUnregister ( ( void * ) p ); s3_pinned_buffers.erase ( it ); }
    if ( cudaHostRegister ( ( void * ) p, bytes, cudaHostRegisterPortable ) == cudaSuccess ) { s3_pinned_buffers[p] = bytes; }
    else { cudaGetLastError (); }
}
static void s3_pin_round1 ( uint * nextQuery, uint * nextReadLength, uint ** answers, uint numCases, uint wordPerQuery, uint word_per_ans, ullint batchSize )
{
    const size_t up = ( ( size_t ) batchSize + 31 ) / 32 * 32;
    s3_pin_once ( nextQuery, up * wordPerQuery * sizeof ( uint ) );
    s3_pin_once ( nextReadLength, ( size_t ) batchSize * sizeof ( uint ) );
    for ( uint c = 0; c < numCases; c++ ) { s3_pin_once ( answers[c], up * word_per_ans * sizeof ( uint ) ); }
}

// ---- alignment.cu:118 / :329 ---------------------------------------------------------------
void perform_round1_alignment ( uint * nextQuery, uint * nextReadLength, uint * answers[][MAX_NUM_CASES],
                                uint numMismatch, uint numCases, uint sa_range_allowed, uint wordPerQuery, uint word_per_ans,
                                bool isExactNumMismatch, int doubleBufferIdx, uint blocksNeeded, ullint batchSize,
                                Soap3Index * index, uint * _bwt, uint * _revBwt, uint * _occ, uint * _revOcc )
{
    s3_pin_round1 ( nextQuery, nextReadLength, answers[doubleBufferIdx], numCases, wordPerQuery, word_per_ans, batchSize );
    if ( s3_search_round1 ( ( s3_index * ) _bwt, nextQuery, nextReadLength, batchSize, wordPerQuery,
                            numMismatch, numCases, sa_range_allowed, word_per_ans, isExactNumMismatch,
                            answers[doubleBufferIdx] ) != S3_OK )
    { s3_die ( "perform_round1_alignment" ); }
}

void perform_round1_alignment_no_pipeline ( uint * nextQuery, uint * nextReadLength, uint ** answers,
        uint numMismatch, uint numCases, uint sa_range_allowed, uint wordPerQuery, uint word_per_ans,
        bool isExactNumMismatch, uint blocksNeeded, ullint batchSize,
        Soap3Index * index, uint * _bwt, uint * _revBwt, uint * _occ, uint * _revOcc )
{
    s3_pin_round1 ( nextQuery, nextReadLength, answers, numCases, wordPerQuery, word_per_ans, batchSize );
    if ( s3_search_round1 ( ( s3_index * ) _bwt, nextQuery, nextReadLength, batchSize, wordPerQuery,
                            numMismatch, numCases, sa_range_allowed, word_per_ans, isExactNumMismatch,
                            answers ) != S3_OK )
    { s3_die ( "perform_round1_alignment_no_pipeline" ); }
}

// ---- alignment.cu:221 / :426 ---------------------------------------------------------------
// The reference returns nothing here: the caller re-counts the bad reads from the status words
// (alignment.cu:898-954), so numBad is dropped.
void perform_round2_alignment ( uint * queries, uint * readLengths, uint * answers[][MAX_NUM_CASES],
                                uint numMismatch, uint numCases, uint sa_range_allowed_2, uint wordPerQuery, uint word_per_ans, uint word_per_ans_2,
                                bool isExactNumMismatch, int doubleBufferIdx, uint blocksNeeded, ullint batchSize,
                                Soap3Index * index, uint * _bwt, uint * _revBwt, uint * _occ, uint * _revOcc,
                                uint processedQuery, uint * badReadIndices[][MAX_NUM_CASES],
                                uint * badAnswers[][MAX_NUM_CASES] )
{
    uint64_t numBad[MAX_NUM_CASES];
    if ( s3_search_round2 ( ( s3_index * ) _bwt, queries, readLengths, answers[doubleBufferIdx],
                            batchSize, processedQuery, wordPerQuery, numMismatch, numCases,
                            sa_range_allowed_2, word_per_ans, word_per_ans_2, isExactNumMismatch,
                            badReadIndices[doubleBufferIdx], badAnswers[doubleBufferIdx], numBad ) != S3_OK )
    { s3_die ( "perform_round2_alignment" ); }
}

void perform_round2_alignment_no_pipeline ( uint * queries, uint * readLengths, uint ** answers,
        uint numMismatch, uint numCases, uint sa_range_allowed_2, uint wordPerQuery, uint word_per_ans, uint word_per_ans_2,
        bool isExactNumMismatch, uint blocksNeeded, ullint batchSize,
        Soap3Index * index, uint * _bwt, uint * _revBwt, uint * _occ, uint * _revOcc,
        uint processedQuery, uint ** badReadIndices, uint ** badAnswers )
{
    uint64_t numBad[MAX_NUM_CASES];
    if ( s3_search_round2 ( ( s3_index * ) _bwt, queries, readLengths, answers,
                            batchSize, processedQuery, wordPerQuery, numMismatch, numCases,
                            sa_range_allowed_2, word_per_ans, word_per_ans_2, isExactNumMismatch,
                            badReadIndices, badAnswers, numBad ) != S3_OK )
    { s3_die ( "perform_round2_alignment_no_pipeline" ); }
}

// ---- DV-DPfunctions.cu:520-741 --------------------------------------------------------------
// The class keeps its declaration (DV-DPfunctions.h:120-164); _DPTable carries the s3_dp handle.
SemiGlobalAligner::SemiGlobalAligner () { _DPTable = NULL; }

void SemiGlobalAligner::decideConfiguration ( int maxReadLength, int maxDNALength,
        int & maxDPTableLength, int & numOfBlocks, int & patternLength, DPParameters & dpPara )
{
    // scheme 1 always: a B200 has room for the whole plane and chunks internally
    // (reference: trial cudaMalloc at 64/48/32/16/8/2 blocks, DV-DPfunctions.cu:568-623)
    maxDPTableLength = maxDNALength;
    numOfBlocks = 64;
    patternLength = maxReadLength + maxDPTableLength;
}

void SemiGlobalAligner::init ( int batchSize, int maxReadLength, int maxDNALength, int maxDPTableLength,
                               DPParameters & dpPara )
{
    this->batchSize = batchSize; this->maxReadLength = maxReadLength;
    this->maxDNALength = maxDNALength; this->maxDPTableLength = maxDPTableLength;
    this->dpPara = dpPara;
    s3_dp_scores sc = { dpPara.matchScore, dpPara.mismatchScore, dpPara.openGapScore, dpPara.extendGapScore };
    s3_dp * dp = NULL;
    if ( s3_dp_create ( maxReadLength, maxDNALength, batchSize, sc, s3_current_device (), &dp ) != S3_OK )
    { s3_die ( "SemiGlobalAligner::init" ); }
    _DPTable = dp;
}

void SemiGlobalAligner::performAlignment ( uint * packedDNASequence, uint * DNALengths,
        uint * packedReadSequence, uint * readLengths, int * cutoffThresholds, int * scores, uint * hitLocs,
        uint * maxScoreCounts, uchar * pattern, int numOfThreads,
        uint * clipLtSizes, uint * clipRtSizes, uint * anchorLeftLocs, uint * anchorRightLocs )
{
    if ( s3_dp_align ( ( s3_dp * ) _DPTable, packedDNASequence, DNALengths, packedReadSequence, readLengths,
                       cutoffThresholds, scores, hitLocs, maxScoreCounts, pattern, numOfThreads,
                       clipLtSizes, clipRtSizes, anchorLeftLocs, anchorRightLocs ) != S3_OK )
    { s3_die ( "SemiGlobalAligner::performAlignment" ); }
}

void SemiGlobalAligner::freeMemory ()
{
    s3_dp_free ( ( s3_dp * ) _DPTable );
    _DPTable = NULL;
}

// ---- soap3-dp-module.cu:62-181 ------------------------------------------------------------------
// alignSingleR, the reference's in-memory entry (soap3-dp-module.h:60-74; alignPairR is declared there but "NOT YET IMPLEMENTED" in
// the reference itself, soap3-dp-module.cu:183): every read's alignments as occRec records in algnResultArrays.  The reference
// reaches them through soap3_dp_single_align -> hostKernel -> addOCCToArray (CPUfunctions.cpp:1958-1968: readID, position, strand,
// source 1, score = mismatches; readID = readIDs[r] + accumReadNum with accumReadNum 0 here); this body asks the device-resident
// single-end chain for the same lists (s3_se_align: search, answer collection, best-hit filter, locate, long-read validation).
//   outputOption 1 (all valid) / 2 (all best): the occurrences in hostKernel's order, up to maxHitNum per read
//   outputOption 3 / 4 (unique / random best) are reached in the reference by searching with 0, 1, .. mismatches in turn
//   (best_single_alignment), whose "first" occurrence depends on that schedule: not offered here.
//   enableDP == 1: the reads left without an alignment go through s3_single_dp_align (DPForUnalignSingle2); its alignments are
//   added to the LAST array with source 2 and score = the DP score, as outputDPSingleResult does (OutputDPResult.cpp:1036-1040) --
//   every candidate that reaches the cutoff, without that function's per-read selection.
// The records of the search go to algnArrays[0] (the reference spreads them over its host threads' arrays).
#include "soap3-dp-module.h"
void alignSingleR ( unsigned int * queries, unsigned int * readLengths, unsigned int * readIDs,
                    unsigned int wordPerQuery,
                    unsigned int numQueries, Soap3Index * index,
                    SingleAlignParam * param,
                    unsigned long long & numOfAnswer,
                    unsigned int & numOfAlignedRead,
                    AlgnResultArrays * algnResultArrays )
{
    numOfAnswer = 0; numOfAlignedRead = 0;
    if ( param->outputOption != 1 && param->outputOption != 2 )
    { printf ( "alignSingleR FAILED .. outputOption %d is not offered by the B200 shim (1 = all valid, 2 = all best)\n", param->outputOption ); exit ( 1 ); }
    if ( !algnResultArrays || algnResultArrays->numArrays < 1 ) { printf ( "alignSingleR FAILED .. no result arrays\n" ); exit ( 1 ); }
    uint * _bwt, * _occ, * _revBwt, * _revOcc;
    GPUINDEXUpload ( index, &_bwt, &_occ, &_revBwt, &_revOcc );
    s3_index * ix = ( s3_index * ) _bwt;
    s3_se_params sp;
    memset ( &sp, 0, sizeof ( sp ) );
    sp.numMismatch = ( uint32_t ) param->numMismatch;
    sp.maxOutputPerRead = param->maxHitNum > 0 ? ( uint32_t ) param->maxHitNum : 0xFFFFFFFFu;
    sp.reportBest = param->outputOption == 2;
    sp.longReadMode = param->maxReadLength > LONG_READ_LEN;               // alignment.cu:2475-2491
    s3_se * se = NULL;
    if ( s3_se_create ( ix, numQueries, &sp, &se ) != S3_OK ) { s3_die ( "alignSingleR" ); }
    s3_se_result r;
    if ( s3_se_align ( se, queries, readLengths, numQueries, wordPerQuery, &r ) != S3_OK ) { s3_die ( "alignSingleR" ); }
    AlgnResult * first = algnResultArrays->algnArrays[0];
    std::vector<uint32_t> unaligned;
    for ( unsigned int q = 0; q < numQueries; q++ )
    {
        const uint32_t a = r.occOffsets[q], b = r.occOffsets[q + 1];
        for ( uint32_t k = a; k < b; k++ )
        { addOCCToArray ( first, readIDs[q], r.positions[k], r.occFlags[2 * k], 1, ( char ) r.occFlags[2 * k + 1] ); }
        numOfAnswer += b - a;
        if ( b > a ) { numOfAlignedRead++; }
        else if ( !( r.readFlags[q] & 1 ) ) { unaligned.push_back ( q ); }
    }
    s3_se_free ( se );
    if ( param->enableDP == 1 && !unaligned.empty () )
    {
        s3_stage_params st;
        memset ( &st, 0, sizeof ( st ) );
        st.insertLow = 0; st.insertHigh = 0; st.strandLeftLeg = 1; st.strandRightLeg = 2;
        st.scores.matchScore = param->scoring.matchScore; st.scores.mismatchScore = param->scoring.mismatchScore;
        st.scores.gapOpenScore = param->scoring.openGapScore; st.scores.gapExtendScore = param->scoring.extendGapScore;
        st.isDefaultThreshold = 0; st.dpScoreThreshold = param->scoring.cutoffThreshold;      // setParam, soap3-dp-module.cu:44-46
        st.softClipLeft = 3; st.softClipRight = 8;                                             // setParam, soap3-dp-module.cu:31-33
        s3_single_dp_result d;
        if ( s3_single_dp_align ( ix, queries, readLengths, numQueries, wordPerQuery, unaligned.data (), unaligned.size (), &st, &d ) != S3_OK ) { s3_die ( "alignSingleR (DP)" ); }
        AlgnResult * last = algnResultArrays->algnArrays[algnResultArrays->numArrays - 1];
        uint32_t pre = 0xFFFFFFFFu;
        for ( uint64_t h = 0; h < d.numHits; h++ )
        {
            addOCCToArray ( last, readIDs[d.hits[h].readID], d.hits[h].pos, d.hits[h].strand, 2, ( char ) d.hits[h].score );
            if ( d.hits[h].readID != pre ) { numOfAlignedRead++; pre = d.hits[h].readID; }
        }
        numOfAnswer += d.numHits;
        s3_single_dp_result_free ( &d );
    }
    GPUINDEXFree ( _bwt, _occ, _revBwt, _revOcc );
}

// ---- DV-DPForBothUnalign.cu:245 (DPForUnalignPairs2) / DV-DPfunctions.cu:3731-3826 --------------------------------------------
// The deep-DP stage of a batch as the reference's own result records.  DPForUnalignPairs2 -> DeepDPWrapper::run2 seeds the pairs
// without any alignment, aligns left read then right read, and DeepDP_Space::DP2CPUAlgnThread turns every candidate whose two reads
// reach their cutoffs into a DeepDPAlignResult (PEAlgnmt.h:409-429) that outputDeepDPResult2 (OutputDPResult.cpp:564) writes out.
// deepDPAlignResults runs the stage on the device (s3_deep_dp_align: seeds, seeding driver, candidate pairs, windows, DP, CIGAR runs)
// and builds exactly those records, in the order the reference's engine emits them: a maintainer calls it from DPForUnalignPairs2
// in place of deepDPWrapper.run2 () and hands the array to outputDeepDPResult2 unchanged.  cigarString_1 / _2 are malloc'ed like the
// encoder's (the output code frees them); *unseeded receives the even read ids of the pairs without a candidate (outputUnaligned),
// malloc'ed.  Returns the number of records.
#include "PEAlgnmt.h"
unsigned int deepDPAlignResults ( unsigned int * queries, unsigned int * upkdReadLengths, unsigned int numQueries, unsigned int wordPerQuery,
                                  const unsigned int * pairReadIDs, unsigned int numPairs,
                                  int insert_high, int insert_low, int peStrandLeftLeg, int peStrandRightLeg,
                                  unsigned int * _bwt, DPParameters * dpParameters,
                                  DeepDPAlignResult ** results, unsigned int ** unseeded, unsigned int * numUnseeded )
{
    s3_index * ix = ( s3_index * ) _bwt;
    s3_stage_params st;
    memset ( &st, 0, sizeof ( st ) );
    st.insertLow = insert_low; st.insertHigh = insert_high; st.strandLeftLeg = peStrandLeftLeg; st.strandRightLeg = peStrandRightLeg;
    st.scores.matchScore = dpParameters->matchScore; st.scores.mismatchScore = dpParameters->mismatchScore;
    st.scores.gapOpenScore = dpParameters->openGapScore; st.scores.gapExtendScore = dpParameters->extendGapScore;
    // getParameterForDeepDP (CPUfunctions.cpp:135-148): ceil(0.3 x read length) per read, or the ini's threshold for both reads
    st.isDefaultThreshold = dpParameters->paramRead[0].cutoffThreshold <= 0;
    st.dpScoreThreshold = dpParameters->paramRead[0].cutoffThreshold;
    st.softClipLeft = dpParameters->softClipLeft; st.softClipRight = dpParameters->softClipRight;
    s3_deep_dp_result d;
    if ( s3_deep_dp_align ( ix, queries, upkdReadLengths, numQueries, wordPerQuery, pairReadIDs, numPairs, &st, &d ) != S3_OK ) { s3_die ( "DPForUnalignPairs2" ); }
    DeepDPAlignResult * out = ( DeepDPAlignResult * ) calloc ( d.numHits ? d.numHits : 1, sizeof ( DeepDPAlignResult ) );
    for ( unsigned long long h = 0; h < d.numHits; h++ )
    {
        const s3_deep_dp_hit & x = d.hits[h];
        DeepDPAlignResult & r = out[h];
        int32_t ed[2], dis[2];
        char * cig[2];
        const uint32_t * runs[2] = { d.runs + x.runOffset1, d.runs + x.runOffset2 };
        const uint32_t nruns[2] = { x.numRuns1, x.numRuns2 };
        const int32_t score[2] = { x.score1, x.score2 };
        for ( int k = 0; k < 2; k++ )
        {
            const uint32_t cap = 12 * nruns[k] + 1;
            cig[k] = ( char * ) malloc ( cap );
            if ( s3_runs_decode ( runs[k], nruns[k], upkdReadLengths[x.readID + k], score[k], st.scores, cig[k], cap, NULL, &ed[k], &dis[k] ) != S3_OK ) { s3_die ( "DPForUnalignPairs2 (CIGAR)" ); }
        }
        r.readID = x.readID;
        r.algnmt_1 = x.pos1; r.strand_1 = ( char ) x.strand1; r.score_1 = x.score1; r.editdist_1 = ed[0]; r.cigarString_1 = cig[0]; r.num_sameScore_1 = ( int ) x.numSame1;
        r.algnmt_2 = x.pos2; r.strand_2 = ( char ) x.strand2; r.score_2 = x.score2; r.editdist_2 = ed[1]; r.cigarString_2 = cig[1]; r.num_sameScore_2 = ( int ) x.numSame2;
        // DV-DPfunctions.cu:3810-3815: the length of the engine's batch entry is that of the left read; with reads of one length, as here
        r.insertSize = r.algnmt_1 < r.algnmt_2 ? ( int ) ( r.algnmt_2 - r.algnmt_1 + upkdReadLengths[x.readID + 1] + dis[1] )
                                                : ( int ) ( r.algnmt_1 - r.algnmt_2 + upkdReadLengths[x.readID] + dis[0] );
    }
    if ( unseeded )
    {
        *unseeded = ( unsigned int * ) malloc ( ( d.numUnseeded ? d.numUnseeded : 1 ) * sizeof ( unsigned int ) );
        memcpy ( *unseeded, d.unseeded, d.numUnseeded * sizeof ( unsigned int ) );
    }
    if ( numUnseeded ) { *numUnseeded = ( unsigned int ) d.numUnseeded; }
    const unsigned int n = ( unsigned int ) d.numHits;
    s3_deep_dp_result_free ( &d );
    *results = out;
    return n;
}

// ---- DV-DPForSingleReads.cu:155 (DPForUnalignSingle2) / DV-DPfunctions.cu:1682-1740 -------------------------------------------
// The single-read DP stage of a batch as the reference's own result records: SingleDP_Space::algnmtCPUThread turns every candidate
// that reaches the cutoff into a SingleAlgnmtResult (PEAlgnmt.h:434-445) for outputDPSingleResult2 (OutputDPResult.cpp:937).
// singleDPAlignResults runs the stage on the device (s3_single_dp_align) and builds those records in the engine's order; a
// maintainer calls it from DPForUnalignSingle2 in place of singleDPWrapper.run ().  cigarString is malloc'ed like the encoder's;
// *unseeded receives the ids of the reads without a candidate, malloc'ed.  Returns the number of records.
unsigned int singleDPAlignResults ( unsigned int * queries, unsigned int * upkdReadLengths, unsigned int numQueries, unsigned int wordPerQuery,
                                    const unsigned int * readIDs, unsigned int numReads,
                                    unsigned int * _bwt, DPParameters * dpParameters,
                                    SingleAlgnmtResult ** results, unsigned int ** unseeded, unsigned int * numUnseeded )
{
    s3_index * ix = ( s3_index * ) _bwt;
    s3_stage_params st;
    memset ( &st, 0, sizeof ( st ) );
    st.strandLeftLeg = 1; st.strandRightLeg = 2;
    st.scores.matchScore = dpParameters->matchScore; st.scores.mismatchScore = dpParameters->mismatchScore;
    st.scores.gapOpenScore = dpParameters->openGapScore; st.scores.gapExtendScore = dpParameters->extendGapScore;
    st.isDefaultThreshold = dpParameters->paramRead[0].cutoffThreshold <= 0;                  // getParameterForSingleDP, CPUfunctions.cpp:190-260
    st.dpScoreThreshold = dpParameters->paramRead[0].cutoffThreshold;
    st.softClipLeft = dpParameters->softClipLeft; st.softClipRight = dpParameters->softClipRight;
    s3_single_dp_result d;
    if ( s3_single_dp_align ( ix, queries, upkdReadLengths, numQueries, wordPerQuery, readIDs, numReads, &st, &d ) != S3_OK ) { s3_die ( "DPForUnalignSingle2" ); }
    SingleAlgnmtResult * out = ( SingleAlgnmtResult * ) calloc ( d.numHits ? d.numHits : 1, sizeof ( SingleAlgnmtResult ) );
    for ( unsigned long long h = 0; h < d.numHits; h++ )
    {
        const s3_dp_hit & x = d.hits[h];
        SingleAlgnmtResult & r = out[h];
        int32_t ed = 0;
        const uint32_t cap = 12 * x.numRuns + 1;
        char * cig = ( char * ) malloc ( cap );
        if ( s3_runs_decode ( d.runs + x.runOffset, x.numRuns, upkdReadLengths[x.readID], x.score, st.scores, cig, cap, NULL, &ed, NULL ) != S3_OK ) { s3_die ( "DPForUnalignSingle2 (CIGAR)" ); }
        r.readID = x.readID; r.strand = ( char ) x.strand; r.algnmt = x.pos; r.score = x.score; r.cigarString = cig; r.editdist = ed; r.num_sameScore = ( int ) x.numSameScore;
    }
    if ( unseeded )
    {
        *unseeded = ( unsigned int * ) malloc ( ( d.numUnseeded ? d.numUnseeded : 1 ) * sizeof ( unsigned int ) );
        memcpy ( *unseeded, d.unseeded, d.numUnseeded * sizeof ( unsigned int ) );
    }
    if ( numUnseeded ) { *numUnseeded = ( unsigned int ) d.numUnseeded; }
    const unsigned int n = ( unsigned int ) d.numHits;
    s3_single_dp_result_free ( &d );
    *results = out;
    return n;
}

// ---- DV-SemiDP.cu (semiGlobalDP2) / DV-DPfunctions.cu:2329-2440 (DP_Space::algnmtCPUThread) -------------------------------------
// Mate rescue of a batch of read pairs as the reference's own result records.  semiGlobalDP2 aligns, for every pair of which one
// read (or both, without a valid pairing) has occurrences, the mate inside the window each occurrence allows, and
// DP_Space::algnmtCPUThread turns EVERY window into an AlgnmtDPResult (PEAlgnmt.h:384-402; whichFromDP 2 and position 0xFFFFFFFF when
// the mate missed its cutoff) for outputDPResult2 (OutputDPResult.cpp:240).  rescueDPAlignResults runs the whole paired-end chain on
// the device (s3_pe_align: search, answer collection, routing, locate, pairing, windows, DP, CIGAR runs) and builds those records in
// the order the reference's occurrence stream visits the windows.  cigarString is malloc'ed like the encoder's (NULL for a miss).
unsigned int rescueDPAlignResults ( unsigned int * queries, unsigned int * upkdReadLengths, unsigned int numQueries, unsigned int wordPerQuery,
                                    unsigned int maxReadLength, int insert_high, int insert_low, int peStrandLeftLeg, int peStrandRightLeg,
                                    unsigned int numMismatch, unsigned int maxOutputPerRead, unsigned int maxHitNumForDP,
                                    unsigned int * _bwt, DPParameters * dpParameters, AlgnmtDPResult ** results )
{
    s3_index * ix = ( s3_index * ) _bwt;
    s3_pe_params par;
    memset ( &par, 0, sizeof ( par ) );
    par.numMismatch = numMismatch; par.insertLow = insert_low; par.insertHigh = insert_high;
    par.strandLeftLeg = peStrandLeftLeg; par.strandRightLeg = peStrandRightLeg;
    par.maxOutputPerRead = maxOutputPerRead; par.maxHitNumForDP = maxHitNumForDP; par.keepSecondBest = 0;
    par.scores.matchScore = dpParameters->matchScore; par.scores.mismatchScore = dpParameters->mismatchScore;
    par.scores.gapOpenScore = dpParameters->openGapScore; par.scores.gapExtendScore = dpParameters->extendGapScore;
    par.cutoffThreshold = dpParameters->paramRead[0].cutoffThreshold > 0 ? dpParameters->paramRead[0].cutoffThreshold : -1;      // getParameterForDefaultDP, CPUfunctions.cpp:59-86
    par.softClipLeft = dpParameters->softClipLeft; par.softClipRight = dpParameters->softClipRight;
    s3_pe * pe = NULL;
    if ( s3_pe_create ( ix, numQueries, maxReadLength, &par, &pe ) != S3_OK ) { s3_die ( "semiGlobalDP2" ); }
    s3_pe_result r;
    if ( s3_pe_align ( pe, queries, upkdReadLengths, numQueries, wordPerQuery, &r ) != S3_OK ) { s3_die ( "semiGlobalDP2" ); }
    const unsigned int n = ( unsigned int ) r.numWindows;
    AlgnmtDPResult * out = ( AlgnmtDPResult * ) calloc ( n ? n : 1, sizeof ( AlgnmtDPResult ) );
    for ( unsigned int t = 0; t < n; t++ )
    {
        const s3_pe_dp_result & x = r.dp[t];
        AlgnmtDPResult & a = out[t];
        const unsigned int alignedID = x.dpReadID ^ 1u, alignedIsReadOrMate = alignedID & 1u;
        a.readID = alignedID - alignedIsReadOrMate;
        unsigned int dpPos = 0xFFFFFFFFu;
        if ( x.numRuns )
        {
            int32_t ed = 0, dis = 0;
            const uint32_t cap = 12 * x.numRuns + 1;
            char * cig = ( char * ) malloc ( cap );
            if ( s3_runs_decode ( r.runs + x.runOffset, x.numRuns, upkdReadLengths[x.dpReadID], x.score, par.scores, cig, cap, NULL, &ed, &dis ) != S3_OK ) { s3_die ( "semiGlobalDP2 (CIGAR)" ); }
            a.cigarString = cig; a.editdist = ed; a.whichFromDP = ( char ) ( 1 - alignedIsReadOrMate ); a.num_sameScore = ( int ) x.numSameScore;
            dpPos = x.dpPos;
            a.insertSize = dpPos < x.alignedPos ? ( int ) ( x.alignedPos - dpPos + upkdReadLengths[alignedID] )                     // DV-DPfunctions.cu:2388-2400
                                                : ( int ) ( dpPos - x.alignedPos + upkdReadLengths[x.dpReadID] + dis );
        }
        else { a.cigarString = NULL; a.whichFromDP = 2; }
        const char dpStrand = ( char ) ( x.leftOrRight == 0 ? peStrandLeftLeg : peStrandRightLeg );
        if ( alignedIsReadOrMate == 0 )
        {
            a.algnmt_1 = x.alignedPos; a.algnmt_2 = dpPos; a.score_1 = x.alignedMismatches; a.score_2 = x.score; a.strand_1 = ( char ) x.alignedStrand; a.strand_2 = dpStrand;
        }
        else
        {
            a.algnmt_1 = dpPos; a.algnmt_2 = x.alignedPos; a.score_1 = x.score; a.score_2 = x.alignedMismatches; a.strand_1 = dpStrand; a.strand_2 = ( char ) x.alignedStrand;
        }
    }
    s3_pe_free ( pe );
    *results = out;
    return n;
}
