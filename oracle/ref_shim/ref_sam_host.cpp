// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_sam.so, built by oracle/build_ref.sh).
// The reference's SAM record of a properly paired read pair: pairOutputSAMAPI (BGS-IO.cpp:3478-3793) with everything it calls --
// getChrAndPosWithBoundaryCheck / BoundaryCheck / getChrAndPos (:1746-2007), getMdStr + mdStr (PE.cpp), the MAPQ functions,
// initializeSAMAlgnmt / initializeSAMAlgnmt2 (:2036-2278), samtools' bam_aux_append -- from BGS-IO.cpp, PE.cpp, SAM.cpp, PEAlgnmt.cpp,
// SAList.cpp and samtools-0.1.18/bam_aux.c compiled whole and unmodified.  samwrite is defined HERE and keeps the record instead of
// writing it; this file only builds the structs the function reads (SRAQueryInput, HSP, HSPAux, OCC, PEOutput, a header with the
// chromosome names) from flat arrays.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#include "BGS-IO.h"
#include "PEAlgnmt.h"
#include "SAM.h"
#include "HSPAux.h"
#include "definitions.h"

int bam_verbose = 0;
#include "sam_nt16.inc"          // bam_nt16_table, samtools-0.1.18/bam_import.c:24-41

void bwase_initialize ( int * g_log_n );     // CPUfunctions.cpp:3014-3019, in mapq-style cut below
#include "sam_bwase.inc"

struct Kept { bam1_core_t core; int l_aux, data_len; std::vector<uint8_t> data; };
static std::vector<Kept> g_kept;

extern "C" int samwrite ( samfile_t * fp, const bam1_t * b )
{
    Kept k;
    k.core = b->core; k.l_aux = b->l_aux; k.data_len = b->data_len;
    k.data.assign ( b->data, b->data + b->data_len );
    g_kept.push_back ( k );
    return b->data_len;
}

extern "C" int ref_sam_pair ( const uint32_t * pac, uint32_t dnaLength, const uint32_t * translate, uint32_t numSeg, const uint32_t * ambiguityMap,
                              const uint32_t * chrEndPos, uint32_t numChr, const char * const * chrNames,
                              int alignmentType, int bwaLike, int dpMatch, int dpMismatch, int isFastq, int maxMAPQ, int minMAPQ, int isPrintMDNM,
                              const char * readGroup, int outputXAZ, uint32_t peMaxOutputPerPair,
                              const uint32_t * pairs, uint32_t numPairs, int bestIdx,
                              const uint8_t * query1, const uint8_t * query2, const char * qual1, const char * qual2, int len1, int len2,
                              const char * name1, const char * name2,
                              int minTot, int secMinTot, int X0f, int X0s, int X1f, int X1s, int numMin, int best1, int best2, uint32_t totalValid,
                              int32_t * core, uint8_t * data, int32_t dataCap, int32_t * dataLen )
{
    HSP hsp;
    memset ( &hsp, 0, sizeof ( hsp ) );
    hsp.dnaLength = dnaLength;
    hsp.packedDNA = ( unsigned int * ) pac;
    hsp.numOfRemovedSegment = numSeg;
    std::vector<Translate> tr ( numSeg );
    for ( uint32_t i = 0; i < numSeg; i++ ) { tr[i].startPos = translate[3 * i]; tr[i].chrID = translate[3 * i + 1]; tr[i].correction = translate[3 * i + 2]; }
    hsp.translate = tr.data ();
    hsp.ambiguityMap = ( unsigned int * ) ambiguityMap;
    std::vector<SeqOffset> so ( numChr );
    for ( uint32_t i = 0; i < numChr; i++ ) { memset ( &so[i], 0, sizeof ( SeqOffset ) ); so[i].endPos = chrEndPos[i]; }
    hsp.seqOffset = so.data ();
    hsp.numOfSeq = numChr;
    HSPAux aux;
    memset ( &aux, 0, sizeof ( aux ) );
    aux.isFastq = isFastq; aux.dpMatchScore = dpMatch; aux.dpMisMatchScore = dpMismatch; aux.alignmentType = alignmentType;
    aux.minMAPQ = minMAPQ; aux.maxMAPQ = maxMAPQ; aux.bwaLikeScore = bwaLike; aux.readGroup = ( char * ) readGroup; aux.isPrintMDNM = isPrintMDNM;
    bwase_initialize ( aux.g_log_n );
    SRAIndex index;
    memset ( &index, 0, sizeof ( index ) );
    index.hsp = &hsp; index.hspaux = &aux;
    OCC occ;
    memset ( &occ, 0, sizeof ( occ ) );
    SAMOccurrenceConstruct ( &occ );
    bam_header_t header;
    memset ( &header, 0, sizeof ( header ) );
    header.n_targets = numChr;
    header.target_name = ( char ** ) chrNames;
    samfile_t sf;
    memset ( &sf, 0, sizeof ( sf ) );
    sf.header = &header;
    SRASetting setting;
    memset ( &setting, 0, sizeof ( setting ) );
    setting.occ = &occ; setting.SAMOutFilePtr = &sf;
    SRAQueryInput in;
    memset ( &in, 0, sizeof ( in ) );
    in.AlgnmtIndex = &index; in.QuerySetting = &setting;
    // the pairs as PEMappingOccurrences leaves them: buckets of PE_MAX_BUCKET_SIZE
    PEOutput out;
    memset ( &out, 0, sizeof ( out ) );
    std::vector<PEPairList *> buckets;
    PEPairs * best = NULL;
    for ( uint32_t p = 0; p < numPairs || buckets.empty (); p++ )
    {
        if ( buckets.empty () || buckets.back ()->pairsCount == PE_MAX_BUCKET_SIZE )
        {
            PEPairList * b = ( PEPairList * ) calloc ( 1, sizeof ( PEPairList ) );
            if ( !buckets.empty () ) { buckets.back ()->next = b; }
            buckets.push_back ( b );
        }
        if ( p >= numPairs ) { break; }
        PEPairs * x = &buckets.back ()->pairs[buckets.back ()->pairsCount++];
        x->algnmt_1 = pairs[7 * p]; x->strand_1 = ( char ) pairs[7 * p + 1]; x->mismatch_1 = ( char ) pairs[7 * p + 2];
        x->algnmt_2 = pairs[7 * p + 3]; x->strand_2 = ( char ) pairs[7 * p + 4]; x->mismatch_2 = ( char ) pairs[7 * p + 5];
        x->totalMismatchCount = ( char ) pairs[7 * p + 6];
        if ( ( int ) p == bestIdx ) { best = x; }
    }
    out.root = buckets[0]; out.tail = buckets.back ();
    DynamicUint8Array * xaz = DynamicUint8ArrayConstruct ();
    g_kept.clear ();
    pairOutputSAMAPI ( &in, &out, best, ( unsigned char * ) query1, ( unsigned char * ) query2, ( char * ) qual1, ( char * ) qual2, len1, len2,
                       ( char * ) name1, ( char * ) name2, ( char ) minTot, ( char ) secMinTot, outputXAZ, xaz, peMaxOutputPerPair,
                       X0f, X0s, X1f, X1s, numMin, ( char ) best1, ( char ) best2, totalValid );
    int n = ( int ) g_kept.size ();
    for ( int r = 0; r < n && r < 2; r++ )
    {
        const Kept & k = g_kept[r];
        int32_t * c = core + 12 * r;
        c[0] = k.core.tid; c[1] = k.core.pos; c[2] = k.core.bin; c[3] = k.core.qual; c[4] = k.core.l_qname; c[5] = k.core.flag; c[6] = k.core.n_cigar;
        c[7] = k.core.l_qseq; c[8] = k.core.mtid; c[9] = k.core.mpos; c[10] = k.core.isize; c[11] = k.l_aux;
        dataLen[r] = k.data_len;
        if ( k.data_len <= dataCap ) { memcpy ( data + ( size_t ) r * dataCap, k.data.data (), k.data_len ); }
    }
    DynamicUint8ArrayFree ( xaz );
    SAMOccurrenceDestruct ( &occ );
    for ( size_t i = 0; i < buckets.size (); i++ ) { free ( buckets[i] ); }
    return n;
}

// OCCOutputSAMAPI (BGS-IO.cpp:5556-5772): the record of a single read from its occurrence list (none: unmapped)
extern "C" int ref_sam_single ( const uint32_t * pac, uint32_t dnaLength, const uint32_t * translate, uint32_t numSeg, const uint32_t * ambiguityMap,
                                const uint32_t * chrEndPos, uint32_t numChr, const char * const * chrNames,
                                int alignmentType, int bwaLike, int isFastq, int maxMAPQ, int minMAPQ, int isPrintMDNM, const char * readGroup,
                                const uint32_t * occ, uint32_t numOcc,
                                const uint8_t * query, const char * qual, int len, const char * name,
                                int32_t * core, uint8_t * data, int32_t dataCap, int32_t * dataLen )
{
    HSP hsp;
    memset ( &hsp, 0, sizeof ( hsp ) );
    hsp.dnaLength = dnaLength;
    hsp.packedDNA = ( unsigned int * ) pac;
    hsp.numOfRemovedSegment = numSeg;
    std::vector<Translate> tr ( numSeg );
    for ( uint32_t i = 0; i < numSeg; i++ ) { tr[i].startPos = translate[3 * i]; tr[i].chrID = translate[3 * i + 1]; tr[i].correction = translate[3 * i + 2]; }
    hsp.translate = tr.data ();
    hsp.ambiguityMap = ( unsigned int * ) ambiguityMap;
    std::vector<SeqOffset> so ( numChr );
    for ( uint32_t i = 0; i < numChr; i++ ) { memset ( &so[i], 0, sizeof ( SeqOffset ) ); so[i].endPos = chrEndPos[i]; }
    hsp.seqOffset = so.data ();
    hsp.numOfSeq = numChr;
    HSPAux aux;
    memset ( &aux, 0, sizeof ( aux ) );
    aux.isFastq = isFastq; aux.alignmentType = alignmentType; aux.minMAPQ = minMAPQ; aux.maxMAPQ = maxMAPQ; aux.bwaLikeScore = bwaLike;
    aux.readGroup = ( char * ) readGroup; aux.isPrintMDNM = isPrintMDNM;
    bwase_initialize ( aux.g_log_n );
    SRAIndex index;
    memset ( &index, 0, sizeof ( index ) );
    index.hsp = &hsp; index.hspaux = &aux;
    OCC occBuf;
    memset ( &occBuf, 0, sizeof ( occBuf ) );
    SAMOccurrenceConstruct ( &occBuf );
    bam_header_t header;
    memset ( &header, 0, sizeof ( header ) );
    header.n_targets = numChr;
    header.target_name = ( char ** ) chrNames;
    samfile_t sf;
    memset ( &sf, 0, sizeof ( sf ) );
    sf.header = &header;
    SRASetting setting;
    memset ( &setting, 0, sizeof ( setting ) );
    setting.occ = &occBuf; setting.SAMOutFilePtr = &sf;
    SRAQueryInfo info;
    memset ( &info, 0, sizeof ( info ) );
    info.ReadCode = ( unsigned char * ) query; info.ReportingReadCode = ( unsigned char * ) query; info.ReadQuality = ( char * ) qual;
    info.ReadLength = len; info.ReadName = ( char * ) name; info.ReadStrand = QUERY_POS_STRAND;
    SRAQueryInput in;
    memset ( &in, 0, sizeof ( in ) );
    in.AlgnmtIndex = &index; in.QuerySetting = &setting; in.QueryInfo = &info;
    OCCList list;
    list.occ = ( SRAOccurrence * ) calloc ( numOcc ? numOcc : 1, sizeof ( SRAOccurrence ) );
    list.curr_size = numOcc; list.available_size = numOcc;
    for ( uint32_t i = 0; i < numOcc; i++ ) { list.occ[i].ambPosition = occ[3 * i]; list.occ[i].strand = ( uint8_t ) occ[3 * i + 1]; list.occ[i].mismatchCount = ( uint8_t ) occ[3 * i + 2]; }
    DynamicUint8Array * xaz = DynamicUint8ArrayConstruct ();
    g_kept.clear ();
    OCCOutputSAMAPI ( &in, &list, xaz, len, alignmentType );
    int n = ( int ) g_kept.size ();
    if ( n >= 1 )
    {
        const Kept & k = g_kept[0];
        core[0] = k.core.tid; core[1] = k.core.pos; core[2] = k.core.bin; core[3] = k.core.qual; core[4] = k.core.l_qname; core[5] = k.core.flag; core[6] = k.core.n_cigar;
        core[7] = k.core.l_qseq; core[8] = k.core.mtid; core[9] = k.core.mpos; core[10] = k.core.isize; core[11] = k.l_aux;
        dataLen[0] = k.data_len;
        if ( k.data_len <= dataCap ) { memcpy ( data, k.data.data (), k.data_len ); }
    }
    DynamicUint8ArrayFree ( xaz );
    SAMOccurrenceDestruct ( &occBuf );
    free ( list.occ );
    return n;
}

// SingleDPOutputSAMAPI (BGS-IO.cpp:5857-6118): the record of a single read from its DP alignments (results[5 * i] = position, strand,
// score, edit distance, index of its special CIGAR in cigars[]; position 0xFFFFFFFF in the first entry: unaligned)
extern "C" int ref_sam_single_dp ( const uint32_t * pac, uint32_t dnaLength, const uint32_t * translate, uint32_t numSeg, const uint32_t * ambiguityMap,
                                   const uint32_t * chrEndPos, uint32_t numChr, const char * const * chrNames,
                                   int alignmentType, int bwaLike, int isFastq, int maxMAPQ, int minMAPQ, int isPrintMDNM, const char * readGroup,
                                   int dpMatch, int singleDPcutoff,
                                   const int32_t * results, uint32_t numResult, const char * const * cigars,
                                   const uint8_t * query, const char * qual, int len, const char * name,
                                   int32_t * core, uint8_t * data, int32_t dataCap, int32_t * dataLen )
{
    HSP hsp;
    memset ( &hsp, 0, sizeof ( hsp ) );
    hsp.dnaLength = dnaLength;
    hsp.packedDNA = ( unsigned int * ) pac;
    hsp.numOfRemovedSegment = numSeg;
    std::vector<Translate> tr ( numSeg );
    for ( uint32_t i = 0; i < numSeg; i++ ) { tr[i].startPos = translate[3 * i]; tr[i].chrID = translate[3 * i + 1]; tr[i].correction = translate[3 * i + 2]; }
    hsp.translate = tr.data ();
    hsp.ambiguityMap = ( unsigned int * ) ambiguityMap;
    std::vector<SeqOffset> so ( numChr );
    for ( uint32_t i = 0; i < numChr; i++ ) { memset ( &so[i], 0, sizeof ( SeqOffset ) ); so[i].endPos = chrEndPos[i]; }
    hsp.seqOffset = so.data ();
    hsp.numOfSeq = numChr;
    HSPAux aux;
    memset ( &aux, 0, sizeof ( aux ) );
    aux.isFastq = isFastq; aux.alignmentType = alignmentType; aux.minMAPQ = minMAPQ; aux.maxMAPQ = maxMAPQ; aux.bwaLikeScore = bwaLike;
    aux.readGroup = ( char * ) readGroup; aux.isPrintMDNM = isPrintMDNM; aux.dpMatchScore = dpMatch; aux.singleDPcutoffThreshold = singleDPcutoff;
    bwase_initialize ( aux.g_log_n );
    SRAIndex index;
    memset ( &index, 0, sizeof ( index ) );
    index.hsp = &hsp; index.hspaux = &aux;
    OCC occBuf;
    memset ( &occBuf, 0, sizeof ( occBuf ) );
    SAMOccurrenceConstruct ( &occBuf );
    bam_header_t header;
    memset ( &header, 0, sizeof ( header ) );
    header.n_targets = numChr;
    header.target_name = ( char ** ) chrNames;
    samfile_t sf;
    memset ( &sf, 0, sizeof ( sf ) );
    sf.header = &header;
    SRASetting setting;
    memset ( &setting, 0, sizeof ( setting ) );
    setting.occ = &occBuf; setting.SAMOutFilePtr = &sf;
    SRAQueryInfo info;
    memset ( &info, 0, sizeof ( info ) );
    info.ReadCode = ( unsigned char * ) query; info.ReportingReadCode = ( unsigned char * ) query; info.ReadQuality = ( char * ) qual;
    info.ReadLength = len; info.ReadName = ( char * ) name; info.ReadStrand = QUERY_POS_STRAND;
    SRAQueryInput in;
    memset ( &in, 0, sizeof ( in ) );
    in.AlgnmtIndex = &index; in.QuerySetting = &setting; in.QueryInfo = &info;
    std::vector<SingleAlgnmtResult> res ( numResult ? numResult : 1 );
    memset ( res.data (), 0, res.size () * sizeof ( SingleAlgnmtResult ) );
    for ( uint32_t i = 0; i < numResult; i++ )
    {
        res[i].readID = 0; res[i].algnmt = ( unsigned int ) results[5 * i]; res[i].strand = ( char ) results[5 * i + 1]; res[i].score = results[5 * i + 2];
        res[i].editdist = results[5 * i + 3]; res[i].cigarString = ( char * ) cigars[results[5 * i + 4]]; res[i].num_sameScore = 1;
    }
    DynamicUint8Array * xaz = DynamicUint8ArrayConstruct ();
    g_kept.clear ();
    SingleDPOutputSAMAPI ( &in, res.data (), 0, numResult, xaz );
    int n = ( int ) g_kept.size ();
    if ( n >= 1 )
    {
        const Kept & k = g_kept[0];
        core[0] = k.core.tid; core[1] = k.core.pos; core[2] = k.core.bin; core[3] = k.core.qual; core[4] = k.core.l_qname; core[5] = k.core.flag; core[6] = k.core.n_cigar;
        core[7] = k.core.l_qseq; core[8] = k.core.mtid; core[9] = k.core.mpos; core[10] = k.core.isize; core[11] = k.l_aux;
        dataLen[0] = k.data_len;
        if ( k.data_len <= dataCap ) { memcpy ( data, k.data.data (), k.data_len ); }
    }
    DynamicUint8ArrayFree ( xaz );
    SAMOccurrenceDestruct ( &occBuf );
    return n;
}

// pairDeepDPOutputSAMAPI (BGS-IO.cpp:3824-4500): the two records of a pair from its deep-DP alignments.  results[13 * i] = insert size,
// then per read position, strand, score, edit distance, tie count, index of its special CIGAR in cigars[]; counts[6] = x0, x1, mismatches
// of the two reads (hspaux->x0_array / x1_array / mismatch_array).  The function writes into the reported entry (a read-through pair
// loses one read): a copy is passed.
extern "C" int ref_sam_deep_dp ( const uint32_t * pac, uint32_t dnaLength, const uint32_t * translate, uint32_t numSeg, const uint32_t * ambiguityMap,
                                 const uint32_t * chrEndPos, uint32_t numChr, const char * const * chrNames,
                                 int alignmentType, int bwaLike, int isFastq, int maxMAPQ, int minMAPQ, int isPrintMDNM, const char * readGroup,
                                 int dpMatch, int dpMismatch,
                                 const int32_t * results, uint32_t num, int bestIdx, const char * const * cigars, const int32_t * counts,
                                 const uint8_t * query1, const uint8_t * query2, const char * qual1, const char * qual2, int len1, int len2,
                                 const char * name1, const char * name2,
                                 int32_t * core, uint8_t * data, int32_t dataCap, int32_t * dataLen )
{
    HSP hsp;
    memset ( &hsp, 0, sizeof ( hsp ) );
    hsp.dnaLength = dnaLength;
    hsp.packedDNA = ( unsigned int * ) pac;
    hsp.numOfRemovedSegment = numSeg;
    std::vector<Translate> tr ( numSeg );
    for ( uint32_t i = 0; i < numSeg; i++ ) { tr[i].startPos = translate[3 * i]; tr[i].chrID = translate[3 * i + 1]; tr[i].correction = translate[3 * i + 2]; }
    hsp.translate = tr.data ();
    hsp.ambiguityMap = ( unsigned int * ) ambiguityMap;
    std::vector<SeqOffset> so ( numChr );
    for ( uint32_t i = 0; i < numChr; i++ ) { memset ( &so[i], 0, sizeof ( SeqOffset ) ); so[i].endPos = chrEndPos[i]; }
    hsp.seqOffset = so.data ();
    hsp.numOfSeq = numChr;
    HSPAux aux;
    memset ( &aux, 0, sizeof ( aux ) );
    aux.isFastq = isFastq; aux.alignmentType = alignmentType; aux.minMAPQ = minMAPQ; aux.maxMAPQ = maxMAPQ; aux.bwaLikeScore = bwaLike;
    aux.readGroup = ( char * ) readGroup; aux.isPrintMDNM = isPrintMDNM; aux.dpMatchScore = dpMatch; aux.dpMisMatchScore = dpMismatch;
    bwase_initialize ( aux.g_log_n );
    // the pair's reads are ids 0 and 1
    int x0a[2] = { counts[0], counts[3] }, x1a[2] = { counts[1], counts[4] }, mma[2] = { counts[2], counts[5] };
    aux.x0_array = x0a; aux.x1_array = x1a; aux.mismatch_array = mma;
    SRAIndex index;
    memset ( &index, 0, sizeof ( index ) );
    index.hsp = &hsp; index.hspaux = &aux;
    OCC occBuf;
    memset ( &occBuf, 0, sizeof ( occBuf ) );
    SAMOccurrenceConstruct ( &occBuf );
    bam_header_t header;
    memset ( &header, 0, sizeof ( header ) );
    header.n_targets = numChr;
    header.target_name = ( char ** ) chrNames;
    samfile_t sf;
    memset ( &sf, 0, sizeof ( sf ) );
    sf.header = &header;
    SRASetting setting;
    memset ( &setting, 0, sizeof ( setting ) );
    setting.occ = &occBuf; setting.SAMOutFilePtr = &sf;
    SRAQueryInput in;
    memset ( &in, 0, sizeof ( in ) );
    in.AlgnmtIndex = &index; in.QuerySetting = &setting;
    std::vector<DeepDPAlignResult> res ( num ? num : 1 );
    memset ( res.data (), 0, res.size () * sizeof ( DeepDPAlignResult ) );
    for ( uint32_t i = 0; i < num; i++ )
    {
        const int32_t * r = results + 13 * i;
        res[i].readID = 0; res[i].insertSize = r[0];
        res[i].algnmt_1 = ( unsigned int ) r[1]; res[i].strand_1 = ( char ) r[2]; res[i].score_1 = r[3]; res[i].editdist_1 = r[4]; res[i].num_sameScore_1 = r[5];
        res[i].cigarString_1 = ( char * ) cigars[r[6]];
        res[i].algnmt_2 = ( unsigned int ) r[7]; res[i].strand_2 = ( char ) r[8]; res[i].score_2 = r[9]; res[i].editdist_2 = r[10]; res[i].num_sameScore_2 = r[11];
        res[i].cigarString_2 = ( char * ) cigars[r[12]];
    }
    DynamicUint8Array * xaz = DynamicUint8ArrayConstruct ();
    g_kept.clear ();
    pairDeepDPOutputSAMAPI ( &in, res.data (), bestIdx >= 0 ? &res[bestIdx] : NULL, 0, num, ( unsigned char * ) query1, ( unsigned char * ) query2,
                             ( char * ) qual1, ( char * ) qual2, len1, len2, ( char * ) name1, ( char * ) name2, xaz );
    int n = ( int ) g_kept.size ();
    for ( int r = 0; r < n && r < 2; r++ )
    {
        const Kept & k = g_kept[r];
        int32_t * c = core + 12 * r;
        c[0] = k.core.tid; c[1] = k.core.pos; c[2] = k.core.bin; c[3] = k.core.qual; c[4] = k.core.l_qname; c[5] = k.core.flag; c[6] = k.core.n_cigar;
        c[7] = k.core.l_qseq; c[8] = k.core.mtid; c[9] = k.core.mpos; c[10] = k.core.isize; c[11] = k.l_aux;
        dataLen[r] = k.data_len;
        if ( k.data_len <= dataCap ) { memcpy ( data + ( size_t ) r * dataCap, k.data.data (), k.data_len ); }
    }
    DynamicUint8ArrayFree ( xaz );
    SAMOccurrenceDestruct ( &occBuf );
    return n;
}

// pairDPOutputSAMAPI (BGS-IO.cpp:4504-5554): the two records of a pair from its default-DP (mate rescue) results.  results[11 * i] =
// whichFromDP, edit distance, insert size, tie count, then per read position, strand, score, and last the index of the entry's special
// CIGAR in cigars[]; counts[6] as ref_sam_deep_dp.
extern "C" int ref_sam_pair_dp ( const uint32_t * pac, uint32_t dnaLength, const uint32_t * translate, uint32_t numSeg, const uint32_t * ambiguityMap,
                                 const uint32_t * chrEndPos, uint32_t numChr, const char * const * chrNames,
                                 int alignmentType, int bwaLike, int isFastq, int maxMAPQ, int minMAPQ, int isPrintMDNM, const char * readGroup,
                                 int dpMatch, int dpMismatch,
                                 const int32_t * results, uint32_t num, int bestIdx, const char * const * cigars, const int32_t * counts,
                                 const uint8_t * query1, const uint8_t * query2, const char * qual1, const char * qual2, int len1, int len2,
                                 const char * name1, const char * name2,
                                 int32_t * core, uint8_t * data, int32_t dataCap, int32_t * dataLen )
{
    HSP hsp;
    memset ( &hsp, 0, sizeof ( hsp ) );
    hsp.dnaLength = dnaLength;
    hsp.packedDNA = ( unsigned int * ) pac;
    hsp.numOfRemovedSegment = numSeg;
    std::vector<Translate> tr ( numSeg );
    for ( uint32_t i = 0; i < numSeg; i++ ) { tr[i].startPos = translate[3 * i]; tr[i].chrID = translate[3 * i + 1]; tr[i].correction = translate[3 * i + 2]; }
    hsp.translate = tr.data ();
    hsp.ambiguityMap = ( unsigned int * ) ambiguityMap;
    std::vector<SeqOffset> so ( numChr );
    for ( uint32_t i = 0; i < numChr; i++ ) { memset ( &so[i], 0, sizeof ( SeqOffset ) ); so[i].endPos = chrEndPos[i]; }
    hsp.seqOffset = so.data ();
    hsp.numOfSeq = numChr;
    HSPAux aux;
    memset ( &aux, 0, sizeof ( aux ) );
    aux.isFastq = isFastq; aux.alignmentType = alignmentType; aux.minMAPQ = minMAPQ; aux.maxMAPQ = maxMAPQ; aux.bwaLikeScore = bwaLike;
    aux.readGroup = ( char * ) readGroup; aux.isPrintMDNM = isPrintMDNM; aux.dpMatchScore = dpMatch; aux.dpMisMatchScore = dpMismatch;
    bwase_initialize ( aux.g_log_n );
    int x0a[2] = { counts[0], counts[3] }, x1a[2] = { counts[1], counts[4] }, mma[2] = { counts[2], counts[5] };
    aux.x0_array = x0a; aux.x1_array = x1a; aux.mismatch_array = mma;
    SRAIndex index;
    memset ( &index, 0, sizeof ( index ) );
    index.hsp = &hsp; index.hspaux = &aux;
    OCC occBuf;
    memset ( &occBuf, 0, sizeof ( occBuf ) );
    SAMOccurrenceConstruct ( &occBuf );
    bam_header_t header;
    memset ( &header, 0, sizeof ( header ) );
    header.n_targets = numChr;
    header.target_name = ( char ** ) chrNames;
    samfile_t sf;
    memset ( &sf, 0, sizeof ( sf ) );
    sf.header = &header;
    SRASetting setting;
    memset ( &setting, 0, sizeof ( setting ) );
    setting.occ = &occBuf; setting.SAMOutFilePtr = &sf;
    SRAQueryInput in;
    memset ( &in, 0, sizeof ( in ) );
    in.AlgnmtIndex = &index; in.QuerySetting = &setting;
    std::vector<AlgnmtDPResult> res ( num ? num : 1 );
    memset ( res.data (), 0, res.size () * sizeof ( AlgnmtDPResult ) );
    for ( uint32_t i = 0; i < num; i++ )
    {
        const int32_t * r = results + 11 * i;
        res[i].readID = 0; res[i].whichFromDP = ( char ) r[0]; res[i].editdist = r[1]; res[i].insertSize = r[2]; res[i].num_sameScore = r[3];
        res[i].algnmt_1 = ( unsigned int ) r[4]; res[i].strand_1 = ( char ) r[5]; res[i].score_1 = r[6];
        res[i].algnmt_2 = ( unsigned int ) r[7]; res[i].strand_2 = ( char ) r[8]; res[i].score_2 = r[9];
        res[i].cigarString = ( char * ) cigars[r[10]];
    }
    DynamicUint8Array * xaz = DynamicUint8ArrayConstruct ();
    g_kept.clear ();
    pairDPOutputSAMAPI ( &in, res.data (), bestIdx >= 0 ? &res[bestIdx] : NULL, 0, num, ( unsigned char * ) query1, ( unsigned char * ) query2,
                         ( char * ) qual1, ( char * ) qual2, len1, len2, ( char * ) name1, ( char * ) name2, xaz );
    int n = ( int ) g_kept.size ();
    for ( int r = 0; r < n && r < 2; r++ )
    {
        const Kept & k = g_kept[r];
        int32_t * c = core + 12 * r;
        c[0] = k.core.tid; c[1] = k.core.pos; c[2] = k.core.bin; c[3] = k.core.qual; c[4] = k.core.l_qname; c[5] = k.core.flag; c[6] = k.core.n_cigar;
        c[7] = k.core.l_qseq; c[8] = k.core.mtid; c[9] = k.core.mpos; c[10] = k.core.isize; c[11] = k.l_aux;
        dataLen[r] = k.data_len;
        if ( k.data_len <= dataCap ) { memcpy ( data + ( size_t ) r * dataCap, k.data.data (), k.data_len ); }
    }
    DynamicUint8ArrayFree ( xaz );
    SAMOccurrenceDestruct ( &occBuf );
    return n;
}

// unproperlypairOutputSAMAPI (BGS-IO.cpp:2582-2930): the two records of a pair without a valid pairing, from the two reads' occurrence lists
extern "C" int ref_sam_unpaired ( const uint32_t * pac, uint32_t dnaLength, const uint32_t * translate, uint32_t numSeg, const uint32_t * ambiguityMap,
                                  const uint32_t * chrEndPos, uint32_t numChr, const char * const * chrNames,
                                  int alignmentType, int bwaLike, int isFastq, int maxMAPQ, int minMAPQ, int isPrintMDNM, const char * readGroup,
                                  const uint32_t * occ1, uint32_t numOcc1, const uint32_t * occ2, uint32_t numOcc2, uint32_t peMaxOutputPerRead,
                                  const uint8_t * query1, const uint8_t * query2, const char * qual1, const char * qual2, int len1, int len2,
                                  const char * name1, const char * name2,
                                  int32_t * core, uint8_t * data, int32_t dataCap, int32_t * dataLen )
{
    HSP hsp;
    memset ( &hsp, 0, sizeof ( hsp ) );
    hsp.dnaLength = dnaLength;
    hsp.packedDNA = ( unsigned int * ) pac;
    hsp.numOfRemovedSegment = numSeg;
    std::vector<Translate> tr ( numSeg );
    for ( uint32_t i = 0; i < numSeg; i++ ) { tr[i].startPos = translate[3 * i]; tr[i].chrID = translate[3 * i + 1]; tr[i].correction = translate[3 * i + 2]; }
    hsp.translate = tr.data ();
    hsp.ambiguityMap = ( unsigned int * ) ambiguityMap;
    std::vector<SeqOffset> so ( numChr );
    for ( uint32_t i = 0; i < numChr; i++ ) { memset ( &so[i], 0, sizeof ( SeqOffset ) ); so[i].endPos = chrEndPos[i]; }
    hsp.seqOffset = so.data ();
    hsp.numOfSeq = numChr;
    HSPAux aux;
    memset ( &aux, 0, sizeof ( aux ) );
    aux.isFastq = isFastq; aux.alignmentType = alignmentType; aux.minMAPQ = minMAPQ; aux.maxMAPQ = maxMAPQ; aux.bwaLikeScore = bwaLike;
    aux.readGroup = ( char * ) readGroup; aux.isPrintMDNM = isPrintMDNM;
    bwase_initialize ( aux.g_log_n );
    SRAIndex index;
    memset ( &index, 0, sizeof ( index ) );
    index.hsp = &hsp; index.hspaux = &aux;
    OCC occBuf;
    memset ( &occBuf, 0, sizeof ( occBuf ) );
    SAMOccurrenceConstruct ( &occBuf );
    bam_header_t header;
    memset ( &header, 0, sizeof ( header ) );
    header.n_targets = numChr;
    header.target_name = ( char ** ) chrNames;
    samfile_t sf;
    memset ( &sf, 0, sizeof ( sf ) );
    sf.header = &header;
    SRASetting setting;
    memset ( &setting, 0, sizeof ( setting ) );
    setting.occ = &occBuf; setting.SAMOutFilePtr = &sf;
    SRAQueryInput in;
    memset ( &in, 0, sizeof ( in ) );
    in.AlgnmtIndex = &index; in.QuerySetting = &setting;
    OCCList list[2];
    const uint32_t * src[2] = { occ1, occ2 };
    const uint32_t cnt[2] = { numOcc1, numOcc2 };
    for ( int k = 0; k < 2; k++ )
    {
        list[k].occ = ( SRAOccurrence * ) calloc ( cnt[k] ? cnt[k] : 1, sizeof ( SRAOccurrence ) );
        list[k].curr_size = cnt[k]; list[k].available_size = cnt[k];
        for ( uint32_t i = 0; i < cnt[k]; i++ ) { list[k].occ[i].ambPosition = src[k][3 * i]; list[k].occ[i].strand = ( uint8_t ) src[k][3 * i + 1]; list[k].occ[i].mismatchCount = ( uint8_t ) src[k][3 * i + 2]; }
    }
    DynamicUint8Array * xaz = DynamicUint8ArrayConstruct ();
    g_kept.clear ();
    unproperlypairOutputSAMAPI ( &in, &list[0], &list[1], ( unsigned char * ) query1, ( unsigned char * ) query2, ( char * ) qual1, ( char * ) qual2, len1, len2,
                                 ( char * ) name1, ( char * ) name2, xaz, peMaxOutputPerRead, alignmentType );
    int n = ( int ) g_kept.size ();
    for ( int r = 0; r < n && r < 2; r++ )
    {
        const Kept & k = g_kept[r];
        int32_t * c = core + 12 * r;
        c[0] = k.core.tid; c[1] = k.core.pos; c[2] = k.core.bin; c[3] = k.core.qual; c[4] = k.core.l_qname; c[5] = k.core.flag; c[6] = k.core.n_cigar;
        c[7] = k.core.l_qseq; c[8] = k.core.mtid; c[9] = k.core.mpos; c[10] = k.core.isize; c[11] = k.l_aux;
        dataLen[r] = k.data_len;
        if ( k.data_len <= dataCap ) { memcpy ( data + ( size_t ) r * dataCap, k.data.data (), k.data_len ); }
    }
    DynamicUint8ArrayFree ( xaz );
    SAMOccurrenceDestruct ( &occBuf );
    free ( list[0].occ ); free ( list[1].occ );
    return n;
}

// unproperlypairDPOutputSAMAPI (BGS-IO.cpp:2932-3447): lists[k][6 * i] = position, strand, score, edit distance, isFromDP, index of the CIGAR in cigars[]
extern "C" int ref_sam_unpaired_dp ( const uint32_t * pac, uint32_t dnaLength, const uint32_t * translate, uint32_t numSeg, const uint32_t * ambiguityMap,
                                     const uint32_t * chrEndPos, uint32_t numChr, const char * const * chrNames,
                                     int alignmentType, int bwaLike, int isFastq, int maxMAPQ, int minMAPQ, int isPrintMDNM, const char * readGroup,
                                     int dpMatch, int singleDPcutoff,
                                     const int32_t * list1, uint32_t num1, const int32_t * list2, uint32_t num2, const char * const * cigars,
                                     const uint8_t * query1, const uint8_t * query2, const char * qual1, const char * qual2, int len1, int len2,
                                     const char * name1, const char * name2,
                                     int32_t * core, uint8_t * data, int32_t dataCap, int32_t * dataLen )
{
    HSP hsp;
    memset ( &hsp, 0, sizeof ( hsp ) );
    hsp.dnaLength = dnaLength;
    hsp.packedDNA = ( unsigned int * ) pac;
    hsp.numOfRemovedSegment = numSeg;
    std::vector<Translate> tr ( numSeg );
    for ( uint32_t i = 0; i < numSeg; i++ ) { tr[i].startPos = translate[3 * i]; tr[i].chrID = translate[3 * i + 1]; tr[i].correction = translate[3 * i + 2]; }
    hsp.translate = tr.data ();
    hsp.ambiguityMap = ( unsigned int * ) ambiguityMap;
    std::vector<SeqOffset> so ( numChr );
    for ( uint32_t i = 0; i < numChr; i++ ) { memset ( &so[i], 0, sizeof ( SeqOffset ) ); so[i].endPos = chrEndPos[i]; }
    hsp.seqOffset = so.data ();
    hsp.numOfSeq = numChr;
    HSPAux aux;
    memset ( &aux, 0, sizeof ( aux ) );
    aux.isFastq = isFastq; aux.alignmentType = alignmentType; aux.minMAPQ = minMAPQ; aux.maxMAPQ = maxMAPQ; aux.bwaLikeScore = bwaLike;
    aux.readGroup = ( char * ) readGroup; aux.isPrintMDNM = isPrintMDNM; aux.dpMatchScore = dpMatch; aux.singleDPcutoffThreshold = singleDPcutoff;
    bwase_initialize ( aux.g_log_n );
    SRAIndex index;
    memset ( &index, 0, sizeof ( index ) );
    index.hsp = &hsp; index.hspaux = &aux;
    OCC occBuf;
    memset ( &occBuf, 0, sizeof ( occBuf ) );
    SAMOccurrenceConstruct ( &occBuf );
    bam_header_t header;
    memset ( &header, 0, sizeof ( header ) );
    header.n_targets = numChr;
    header.target_name = ( char ** ) chrNames;
    samfile_t sf;
    memset ( &sf, 0, sizeof ( sf ) );
    sf.header = &header;
    SRASetting setting;
    memset ( &setting, 0, sizeof ( setting ) );
    setting.occ = &occBuf; setting.SAMOutFilePtr = &sf;
    SRAQueryInput in;
    memset ( &in, 0, sizeof ( in ) );
    in.AlgnmtIndex = &index; in.QuerySetting = &setting;
    std::vector<Algnmt> lists[2];
    const int32_t * src[2] = { list1, list2 };
    const uint32_t cnt[2] = { num1, num2 };
    for ( int k = 0; k < 2; k++ )
    {
        lists[k].resize ( cnt[k] ? cnt[k] : 1 );
        memset ( lists[k].data (), 0, lists[k].size () * sizeof ( Algnmt ) );
        for ( uint32_t i = 0; i < cnt[k]; i++ )
        {
            const int32_t * r = src[k] + 6 * i;
            lists[k][i].algnmt = ( unsigned int ) r[0]; lists[k][i].strand = ( char ) r[1]; lists[k][i].score = r[2]; lists[k][i].editdist = r[3];
            lists[k][i].isFromDP = r[4]; lists[k][i].cigarString = ( char * ) cigars[r[5]]; lists[k][i].num_sameScore = 1;
        }
    }
    DynamicUint8Array * xaz = DynamicUint8ArrayConstruct ();
    g_kept.clear ();
    unproperlypairDPOutputSAMAPI ( &in, lists[0].data (), lists[1].data (), ( int ) num1, ( int ) num2, ( unsigned char * ) query1, ( unsigned char * ) query2,
                                   ( char * ) qual1, ( char * ) qual2, len1, len2, ( char * ) name1, ( char * ) name2, xaz );
    int n = ( int ) g_kept.size ();
    for ( int r = 0; r < n && r < 2; r++ )
    {
        const Kept & k = g_kept[r];
        int32_t * c = core + 12 * r;
        c[0] = k.core.tid; c[1] = k.core.pos; c[2] = k.core.bin; c[3] = k.core.qual; c[4] = k.core.l_qname; c[5] = k.core.flag; c[6] = k.core.n_cigar;
        c[7] = k.core.l_qseq; c[8] = k.core.mtid; c[9] = k.core.mpos; c[10] = k.core.isize; c[11] = k.l_aux;
        dataLen[r] = k.data_len;
        if ( k.data_len <= dataCap ) { memcpy ( data + ( size_t ) r * dataCap, k.data.data (), k.data_len ); }
    }
    DynamicUint8ArrayFree ( xaz );
    SAMOccurrenceDestruct ( &occBuf );
    return n;
}

// SingleAnsOutputSAMAPI (BGS-IO.cpp:5774-5827) and, with ambPosition 0xFFFFFFFF, noAnsOutputSAMAPI (:5829-5855)
extern "C" int ref_sam_single_answer ( const uint32_t * pac, uint32_t dnaLength, const uint32_t * translate, uint32_t numSeg, const uint32_t * ambiguityMap,
                                       const uint32_t * chrEndPos, uint32_t numChr, const char * const * chrNames,
                                       int isFastq, int maxMAPQ, int minMAPQ, int isPrintMDNM, const char * readGroup,
                                       uint32_t ambPosition, int strand, int numMismatch, int bestHitNum,
                                       const uint8_t * query, const char * qual, int len, const char * name,
                                       int32_t * core, uint8_t * data, int32_t dataCap, int32_t * dataLen )
{
    HSP hsp;
    memset ( &hsp, 0, sizeof ( hsp ) );
    hsp.dnaLength = dnaLength;
    hsp.packedDNA = ( unsigned int * ) pac;
    hsp.numOfRemovedSegment = numSeg;
    std::vector<Translate> tr ( numSeg );
    for ( uint32_t i = 0; i < numSeg; i++ ) { tr[i].startPos = translate[3 * i]; tr[i].chrID = translate[3 * i + 1]; tr[i].correction = translate[3 * i + 2]; }
    hsp.translate = tr.data ();
    hsp.ambiguityMap = ( unsigned int * ) ambiguityMap;
    std::vector<SeqOffset> so ( numChr );
    for ( uint32_t i = 0; i < numChr; i++ ) { memset ( &so[i], 0, sizeof ( SeqOffset ) ); so[i].endPos = chrEndPos[i]; }
    hsp.seqOffset = so.data ();
    hsp.numOfSeq = numChr;
    HSPAux aux;
    memset ( &aux, 0, sizeof ( aux ) );
    aux.isFastq = isFastq; aux.minMAPQ = minMAPQ; aux.maxMAPQ = maxMAPQ; aux.readGroup = ( char * ) readGroup; aux.isPrintMDNM = isPrintMDNM;
    bwase_initialize ( aux.g_log_n );
    SRAIndex index;
    memset ( &index, 0, sizeof ( index ) );
    index.hsp = &hsp; index.hspaux = &aux;
    OCC occBuf;
    memset ( &occBuf, 0, sizeof ( occBuf ) );
    SAMOccurrenceConstruct ( &occBuf );
    bam_header_t header;
    memset ( &header, 0, sizeof ( header ) );
    header.n_targets = numChr;
    header.target_name = ( char ** ) chrNames;
    samfile_t sf;
    memset ( &sf, 0, sizeof ( sf ) );
    sf.header = &header;
    SRASetting setting;
    memset ( &setting, 0, sizeof ( setting ) );
    setting.occ = &occBuf; setting.SAMOutFilePtr = &sf;
    SRAQueryInfo info;
    memset ( &info, 0, sizeof ( info ) );
    info.ReadCode = ( unsigned char * ) query; info.ReportingReadCode = ( unsigned char * ) query; info.ReadQuality = ( char * ) qual;
    info.ReadLength = len; info.ReadName = ( char * ) name; info.ReadStrand = QUERY_POS_STRAND;
    SRAQueryInput in;
    memset ( &in, 0, sizeof ( in ) );
    in.AlgnmtIndex = &index; in.QuerySetting = &setting; in.QueryInfo = &info;
    g_kept.clear ();
    if ( ambPosition == 0xFFFFFFFFu ) { noAnsOutputSAMAPI ( &in ); }
    else { SingleAnsOutputSAMAPI ( &in, ( char ) strand, ambPosition, numMismatch, bestHitNum ); }
    int n = ( int ) g_kept.size ();
    if ( n >= 1 )
    {
        const Kept & k = g_kept[0];
        core[0] = k.core.tid; core[1] = k.core.pos; core[2] = k.core.bin; core[3] = k.core.qual; core[4] = k.core.l_qname; core[5] = k.core.flag; core[6] = k.core.n_cigar;
        core[7] = k.core.l_qseq; core[8] = k.core.mtid; core[9] = k.core.mpos; core[10] = k.core.isize; core[11] = k.l_aux;
        dataLen[0] = k.data_len;
        if ( k.data_len <= dataCap ) { memcpy ( data, k.data.data (), k.data_len ); }
    }
    SAMOccurrenceDestruct ( &occBuf );
    return n;
}

// the text line of a record: samtools' own bam_format1_core (samtools-0.1.18/bam.c:243-324, cut by line range; kstring.c compiled whole)
extern "C" {
#include "kstring.h"
}
char * bam_flag2char_table = ( char * ) "pPuUrR12sfd\0\0\0\0\0";        // bam.c:11
char * bam_nt16_rev_table = ( char * ) "=ACMGRSVTWYHKDBN";               // bam_import.c:62
#include "sam_format.inc"
extern "C" int ref_sam_format ( const int32_t * core, const uint8_t * data, int32_t dataLen, const char * const * chrNames, int numChr, char * out, int cap )
{
    bam1_t b;
    memset ( &b, 0, sizeof ( b ) );
    b.core.tid = core[0]; b.core.pos = core[1]; b.core.bin = core[2]; b.core.qual = core[3]; b.core.l_qname = core[4]; b.core.flag = core[5]; b.core.n_cigar = core[6];
    b.core.l_qseq = core[7]; b.core.mtid = core[8]; b.core.mpos = core[9]; b.core.isize = core[10];
    b.l_aux = core[11]; b.data_len = dataLen; b.m_data = dataLen; b.data = ( uint8_t * ) data;
    bam_header_t header;
    memset ( &header, 0, sizeof ( header ) );
    header.n_targets = numChr; header.target_name = ( char ** ) chrNames;
    char * s = bam_format1_core ( &header, &b, BAM_OFDEC );
    int n = ( int ) strlen ( s );
    if ( n < cap ) { memcpy ( out, s, n + 1 ); }
    free ( s );
    return n;
}
