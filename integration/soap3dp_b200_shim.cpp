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
