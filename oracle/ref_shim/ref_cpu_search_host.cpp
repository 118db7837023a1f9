// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_cpu_search.so, built by oracle/build_ref.sh).
//
// The reference's own CPU implementation of the search path, driven the way hostKernel drives it for a read the GPU
// did not finish (CPUfunctions.cpp:1313-1328): per read and per case ProcessReadDoubleStrand2 on the models
// SRAModelConstruct built for the read length (alignment.cu:745-751), results in an SAList + OCCList.  Everything that
// runs inside the timed region is the reference's unmodified code: ProcessReadDoubleStrand2 (cut by line range into
// cpu_search.inc), BGS-HostAlgnmtAlgo2.cpp, SAList.cpp, 2bwt-flex/SRA2BWTMdl.c, SRA2BWTCheckAndExtend.c and 2bwt-lib/BWT.c
// compiled from where they lie.  This file only builds the in-memory index structs from arrays (what BWTLoad / LTLoad /
// HSPLoad would read from the index files: BWT.c:115-300, LT.c:37-57, LTConstruct.c:46-102) and loops over reads.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <omp.h>

#include "2bwt-lib/BWT.h"
#include "2bwt-lib/HSP.h"
#include "2bwt-lib/DNACount.h"
#include "2bwt-lib/BWTConstruct.h"
#include "2bwt-flex/LT.h"
#include "2bwt-flex/SRACore.h"
#include "2bwt-flex/SRAArguments.h"
#include "2bwt-flex/SRA2BWTMdl.h"
#include "SAList.h"
#include "BGS-HostAlgnmtAlgo2.h"

#include "cpu_search.inc"      // ProcessReadDoubleStrand2, CPUfunctions.cpp:555-619

struct RefCpu
{
    BWT fwd, rev;
    HSP hsp;
    LT lt, rlt;
    SRAIndex index;
    unsigned int * decode;
    unsigned int * pac;
    unsigned int * saCopy;
    unsigned int cum[2][ALPHABET_SIZE + 1];
};

static void fill_bwt ( BWT * b, const unsigned int * words, unsigned long long nwords, unsigned int isa0, unsigned int n, unsigned int * decode,
                       unsigned int * cum )
{
    memset ( b, 0, sizeof ( BWT ) );
    b->textLength = n;
    b->inverseSa0 = isa0;
    b->saInterval = ALL_ONE_MASK;
    b->inverseSaInterval = ALL_ONE_MASK;
    b->decodeTable = decode;
    b->bwtSizeInWord = BWTResidentSizeInWord ( n ) + WORD_BETWEEN_OCC / 2;
    b->bwtCode = ( unsigned int * ) calloc ( b->bwtSizeInWord + 16, sizeof ( unsigned int ) );
    unsigned long long have = BWTFileSizeInWord ( n );
    if ( nwords < have ) { have = nwords; }
    memcpy ( b->bwtCode, words, have * sizeof ( unsigned int ) );
    BWTClearTrailingBwtCode ( b );
    // cumulative frequencies as the .bwt header holds them: characters of the BWT string = characters of the text
    unsigned long long cnt[ALPHABET_SIZE] = {0, 0, 0, 0};
    #pragma omp parallel
    {
        unsigned long long c[ALPHABET_SIZE] = {0, 0, 0, 0};
        #pragma omp for schedule(static)
        for ( long long w = 0; w < ( long long ) have; w++ )
        {
            unsigned int x = b->bwtCode[w];
            unsigned int lo = x & 0x55555555u, hi = ( x >> 1 ) & 0x55555555u;
            unsigned int c3 = __builtin_popcount ( lo & hi ), c2 = __builtin_popcount ( hi & ~lo ), c1 = __builtin_popcount ( lo & ~hi );
            c[3] += c3; c[2] += c2; c[1] += c1; c[0] += 16 - c3 - c2 - c1;
        }
        #pragma omp critical
        for ( int i = 0; i < ALPHABET_SIZE; i++ ) { cnt[i] += c[i]; }
    }
    cnt[0] -= have * 16 - n;                   // cleared trailing positions read as 'A'
    cum[0] = 0;
    for ( int i = 0; i < ALPHABET_SIZE; i++ ) { cum[i + 1] = cum[i] + ( unsigned int ) cnt[i]; }
    b->cumulativeFreq = cum;
    b->occSizeInWord = BWTOccValueMinorSizeInWord ( n );
    b->occMajorSizeInWord = BWTOccValueMajorSizeInWord ( n );
    b->occValue = ( unsigned int * ) calloc ( b->occSizeInWord + 16, sizeof ( unsigned int ) );
    b->occValueMajor = ( unsigned int * ) calloc ( b->occMajorSizeInWord + 16, sizeof ( unsigned int ) );
    BWTGenerateOccValueFromBwt ( b->bwtCode, b->occValue, b->occValueMajor, n, decode );
    b->_bwtSaValue = &BWTFullSaValue;
}

// LTConstruct.c:46-102: counts of every tableSize-mer of the text followed by tableSize - 1 'A's, inclusive prefix sums
static void fill_lt ( LT * lt, const unsigned int * pac, unsigned int n, int reversed )
{
    const int K = LOOKUP_SIZE;
    unsigned long long entries = 1ULL << ( K * LOOKUP_BIT_PER_CHAR );
    unsigned long long mask = entries - 1;
    lt->tableSize = K;
    lt->ltSizeInWord = ( unsigned int ) entries;
    lt->table = ( unsigned int * ) calloc ( entries, sizeof ( unsigned int ) );
    unsigned int * t = lt->table;
    long long total = ( long long ) n + K - 1;      // window ends at text position e = K - 1 .. n + K - 2 (past n: 'A')
    #pragma omp parallel
    {
        int nt = omp_get_num_threads (), id = omp_get_thread_num ();
        long long per = ( total + nt - 1 ) / nt;
        long long e0 = id * per, e1 = e0 + per < total ? e0 + per : total;
        unsigned long long win = 0;
        for ( long long e = e0 - ( K - 1 ) < 0 ? 0 : e0 - ( K - 1 ); e < e1; e++ )
        {
            unsigned int c = 0;
            if ( e < ( long long ) n )
            {
                unsigned long long p = reversed ? ( unsigned long long ) n - 1 - e : ( unsigned long long ) e;
                c = ( pac[p >> 4] >> ( 30 - 2 * ( p & 15 ) ) ) & 3;
            }
            win = ( ( win << 2 ) | c ) & mask;
            if ( e >= e0 && e >= K - 1 ) { __atomic_fetch_add ( &t[win], 1u, __ATOMIC_RELAXED ); }
        }
    }
    for ( unsigned long long i = 1; i < entries; i++ ) { t[i] += t[i - 1]; }
}

extern "C" {

// bwt / rbwt: BWT code words of the text and of the reversed text (.bwt / .rev.bwt payload); pac: 16 bases per word, first base
// in the top bits (hsp->packedDNA); sa: the n + 1 suffix-array values (bwt->saValue with saInterval 1)
void * ref_cpu_create ( const unsigned int * bwt, const unsigned int * rbwt, unsigned long long nwords, unsigned int isa0, unsigned int risa0,
                        unsigned int n, const unsigned int * pac, const unsigned int * sa, int threads )
{
    if ( threads > 0 ) { omp_set_num_threads ( threads ); }
    RefCpu * h = ( RefCpu * ) calloc ( 1, sizeof ( RefCpu ) );
    h->decode = ( unsigned int * ) malloc ( DNA_OCC_CNT_TABLE_SIZE_IN_WORD * sizeof ( unsigned int ) );
    GenerateDNAOccCountTable ( h->decode );
    fill_bwt ( &h->fwd, bwt, nwords, isa0, n, h->decode, h->cum[0] );
    fill_bwt ( &h->rev, rbwt, nwords, risa0, n, h->decode, h->cum[1] );
    h->fwd.saInterval = 1;
    h->fwd.saValue = ( unsigned int * ) sa;
    h->fwd.saValueSizeInWord = n + 1;
    unsigned long long pw = ( ( unsigned long long ) n + 15 ) / 16;
    h->pac = ( unsigned int * ) calloc ( pw + 8, sizeof ( unsigned int ) );      // check-and-extend reads two words past the window
    memcpy ( h->pac, pac, pw * sizeof ( unsigned int ) );
    h->hsp.dnaLength = n;
    h->hsp.packedDNA = h->pac;
    fill_lt ( &h->lt, h->pac, n, 0 );
    fill_lt ( &h->rlt, h->pac, n, 1 );
    h->index.bwt = &h->fwd;
    h->index.rev_bwt = &h->rev;
    h->index.hsp = &h->hsp;
    h->index.hspaux = NULL;
    h->index.highOcc = NULL;
    h->index.lookupTable = &h->lt;
    h->index.rev_lookupTable = &h->rlt;
    return h;
}

void ref_cpu_free ( void * hv )
{
    RefCpu * h = ( RefCpu * ) hv;
    if ( !h ) { return; }
    BWT * b[2] = { &h->fwd, &h->rev };
    for ( int i = 0; i < 2; i++ ) { free ( b[i]->bwtCode ); free ( b[i]->occValue ); free ( b[i]->occValueMajor ); }
    free ( h->lt.table ); free ( h->rlt.table ); free ( h->pac ); free ( h->decode ); free ( h );
}

// The steps of every case of the model the reference builds for this read length (for the record in the bench line / tests)
int ref_cpu_describe ( void * hv, unsigned int readLength, int maxError, int numCases, char * out, int cap )
{
    RefCpu * h = ( RefCpu * ) hv;
    SRASetting s;
    memset ( &s, 0, sizeof ( s ) );
    s.ReadStrand = QUERY_BOTH_STRAND; s.ErrorType = SRA_STEP_ERROR_TYPE_MISMATCH_ONLY; s.OutputType = SRA_REPORT_ALL; s.MaxError = maxError;
    SRAModel * m = SRAModelConstruct ( readLength, QUERY_POS_STRAND, &s, &h->index, SRA_MODEL_16G );
    int len = 0;
    for ( int c = 0; c < numCases && len < cap - 200; c++ )
    {
        len += snprintf ( out + len, cap - len, "case %d type %d:", c, m->cases[c].type );
        for ( int k = 0; k < MAX_NUM_OF_SRA_STEPS && len < cap - 100; k++ )
        {
            SRAStep * st = &m->cases[c].steps[k];
            len += snprintf ( out + len, cap - len, " [t%d %d..%d e%d-%d ce%d]", st->type, st->start, st->end, st->MinError, st->MaxError, st->ceThreshold );
            if ( st->type == SRA_STEP_TYPE_COMPLETE ) { break; }
        }
        len += snprintf ( out + len, cap - len, "\n" );
    }
    SRAModelFree ( m );
    return len;
}

// reads: numReads rows of readLength base codes (0..3).  Per read: every case through ProcessReadDoubleStrand2.
// counts[4 * r + 0..3] = SA ranges, occurrences inside them, check-and-extend occurrences, total occurrences.
// When outCap > 0 the first outCap hits of a read are written as (position, strand, mismatches) to outPos/outStrand/outMism
// (SA ranges expanded in SA order, then the OCCList), outN[r] = how many were written.
// returns the number of occurrences over the batch
unsigned long long ref_cpu_search ( void * hv, const unsigned char * reads, unsigned int numReads, unsigned int readLength, int maxError, int numCases,
                                    unsigned int maxOutputPerRead, int threads, unsigned int * counts,
                                    unsigned int outCap, unsigned int * outPos, unsigned char * outStrand, unsigned char * outMism, unsigned int * outN )
{
    RefCpu * h = ( RefCpu * ) hv;
    if ( threads > 0 ) { omp_set_num_threads ( threads ); }
    unsigned long long total = 0;
    #pragma omp parallel reduction(+ : total)
    {
        // per thread, as hostKernelArguments[threadId] (CPUfunctions.cpp:823-851, alignment.cu:704-705, 745-751)
        SRASetting setting;
        memset ( &setting, 0, sizeof ( setting ) );
        setting.ReadStrand = QUERY_BOTH_STRAND;
        setting.ErrorType = SRA_STEP_ERROR_TYPE_MISMATCH_ONLY;
        setting.OutputType = SRA_REPORT_ALL;
        setting.MaxError = maxError;
        setting.MaxNBMismatch = 0;
        setting.MaxOutputPerRead = maxOutputPerRead;
        SRAIndex index = h->index;
        SRAModel * model = SRAModelConstruct ( readLength, QUERY_POS_STRAND, &setting, &index, SRA_MODEL_16G );
        SRAModel * model_neg = SRAModelConstruct ( readLength, QUERY_NEG_STRAND, &setting, &index, SRA_MODEL_16G );
        SRAQueryResultCount * rOutput = ( SRAQueryResultCount * ) calloc ( 1, sizeof ( SRAQueryResultCount ) );
        char dummyQuality[SRA_MAX_READ_LENGTH];
        memset ( dummyQuality, 1, sizeof ( dummyQuality ) );
        SRAQueryInfo infoPos, infoNeg;
        memset ( &infoPos, 0, sizeof ( infoPos ) );
        memset ( &infoNeg, 0, sizeof ( infoNeg ) );
        SRAQueryInput inPos, inNeg;
        infoPos.ReadStrand = QUERY_POS_STRAND; infoPos.ReadQuality = dummyQuality; infoPos.ReadLength = readLength;
        infoNeg.ReadStrand = QUERY_NEG_STRAND; infoNeg.ReadQuality = dummyQuality; infoNeg.ReadLength = readLength;
        inPos.QueryInfo = &infoPos; inPos.QuerySetting = &setting; inPos.AlgnmtIndex = &index; inPos.QueryOutput = rOutput;
        inNeg.QueryInfo = &infoNeg; inNeg.QuerySetting = &setting; inNeg.AlgnmtIndex = &index; inNeg.QueryOutput = rOutput;
        SAList * sa_list = SAListConstruct ();
        OCCList * occ_list = OCCListConstruct ();
        unsigned char oStrandQuery[SRA_MAX_READ_LENGTH];
        unsigned char thisQuery[SRA_MAX_READ_LENGTH];

        #pragma omp for schedule(dynamic, 64)
        for ( long long r = 0; r < ( long long ) numReads; r++ )
        {
            memcpy ( thisQuery, reads + ( size_t ) r * readLength, readLength );
            for ( unsigned int i = 0; i < readLength; i++ ) { oStrandQuery[i] = soap3DnaComplement[thisQuery[readLength - i - 1]]; }
            infoPos.ReadId = r; infoNeg.ReadId = r;
            infoPos.ReadCode = thisQuery; infoPos.ReportingReadCode = thisQuery;
            infoNeg.ReadCode = oStrandQuery; infoNeg.ReportingReadCode = thisQuery;
            SAListReset ( sa_list );
            OCCListReset ( occ_list );
            rOutput->TotalOccurrences = 0;
            memset ( rOutput->WithError, 0, sizeof ( rOutput->WithError ) );
            for ( int whichCase = 0; whichCase < numCases; whichCase++ )
            {
                ProcessReadDoubleStrand2 ( &inPos, &inNeg, model, model_neg, whichCase, sa_list, occ_list );
            }
            unsigned int inRanges = 0;
            for ( unsigned int i = 0; i < sa_list->curr_size; i++ ) { inRanges += sa_list->sa[i].saIndexRight - sa_list->sa[i].saIndexLeft + 1; }
            if ( counts )
            {
                counts[4 * r + 0] = sa_list->curr_size;
                counts[4 * r + 1] = inRanges;
                counts[4 * r + 2] = occ_list->curr_size;
                counts[4 * r + 3] = rOutput->TotalOccurrences;
            }
            total += inRanges + occ_list->curr_size;
            if ( outCap )
            {
                unsigned int k = 0;
                size_t base = ( size_t ) r * outCap;
                for ( unsigned int i = 0; i < sa_list->curr_size && k < outCap; i++ )
                {
                    for ( unsigned int j = sa_list->sa[i].saIndexLeft; j <= sa_list->sa[i].saIndexRight && k < outCap; j++ )
                    {
                        outPos[base + k] = ( *h->fwd._bwtSaValue ) ( &h->fwd, j );
                        outStrand[base + k] = sa_list->sa[i].strand;
                        outMism[base + k] = sa_list->sa[i].mismatchCount;
                        k++;
                    }
                }
                for ( unsigned int i = 0; i < occ_list->curr_size && k < outCap; i++ )
                {
                    outPos[base + k] = ( unsigned int ) occ_list->occ[i].ambPosition;
                    outStrand[base + k] = occ_list->occ[i].strand;
                    outMism[base + k] = occ_list->occ[i].mismatchCount;
                    k++;
                }
                outN[r] = k;
            }
        }
        SAListFree ( sa_list );
        OCCListFree ( occ_list );
        SRAModelFree ( model );
        SRAModelFree ( model_neg );
        free ( rOutput );
    }
    return total;
}

}   // extern "C"
