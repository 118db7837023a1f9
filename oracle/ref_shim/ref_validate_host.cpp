// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_validate.so, built by oracle/build_ref.sh).
// The reference's long-read validation -- validateAlignments (CPUfunctions.cpp:1129-1222) with the packers and the popcount
// Hamming distance it calls (PE.cpp:28-60, 148-206, 287-325) -- cut by line range into validate.inc and compiled against the
// reference's own headers.  This file only moves flat arrays in and out of an OCCList.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "2bwt-lib/BWT.h"
#include "2bwt-lib/HSP.h"
#include "2bwt-flex/SRACore.h"
#include "SAList.h"
#include "definitions.h"

#include "validate.inc"

extern "C" uint32_t ref_validate ( const uint32_t * pac, uint32_t dnaLength, const uint8_t * query, uint32_t seed_len, uint32_t read_len,
                                   uint32_t n, uint32_t * pos, uint8_t * strand, uint8_t * mism,
                                   int only_keep_best, int min_seed_mismatch, int max_mismatch, int max_hit_num )
{
    HSP hsp;
    memset ( &hsp, 0, sizeof ( hsp ) );
    hsp.dnaLength = dnaLength;
    hsp.packedDNA = ( unsigned int * ) pac;
    OCCList list;
    list.occ = ( SRAOccurrence * ) calloc ( n ? n : 1, sizeof ( SRAOccurrence ) );
    list.curr_size = n;
    list.available_size = n;
    for ( uint32_t i = 0; i < n; i++ )
    {
        list.occ[i].readID = 7; list.occ[i].ambPosition = pos[i]; list.occ[i].strand = strand[i]; list.occ[i].mismatchCount = mism[i];
    }
    unsigned char q[MAX_READ_LENGTH + 16];
    memcpy ( q, query, read_len );
    validateAlignments ( &list, q, seed_len, read_len, &hsp, only_keep_best != 0, min_seed_mismatch, max_mismatch, max_hit_num, 0 );
    for ( uint32_t i = 0; i < list.curr_size; i++ )
    {
        pos[i] = ( uint32_t ) list.occ[i].ambPosition; strand[i] = list.occ[i].strand; mism[i] = list.occ[i].mismatchCount;
    }
    uint32_t m = list.curr_size;
    free ( list.occ );
    return m;
}
