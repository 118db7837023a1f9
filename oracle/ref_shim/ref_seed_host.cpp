/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * The reference's own seed-hit sort and merge for single-end DP seeding, run on the host: the radix sort macros
 * (DV-DPfunctions.h:60-95), struct SeedPos (:919-926), DPS_DIVIDE_GAP (:944) and the body of
 * SingleEndSeedingEngine::SingleEndSeedingBatch::singleMerge (DV-DPfunctions.cu:1101-1141), cut out of the reference
 * files by sed at build time into seed_merge.inc (the method is renamed to a free function; nothing else changes).
 * It pins oracle/seed_oracle.c: what order the three radix passes of decodePositions (DV-DPfunctions.cu:1214-1216)
 * really leave -- the 8-bit strand pass sorts into the auxiliary array, which is then freed -- and which hits survive.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
using std::vector;
typedef unsigned int uint;
#include "seed_merge.inc"

extern "C" {

/* hits (readID, pos, strand) in arrival order -> the three sorts of decodePositions -> singleMerge.
 * Returns the number of candidates written to out*. */
uint ref_seed_sort_merge(const uint *readID, const uint *pos, const int *strand, uint n, uint *outReadID, uint *outPos, int *outStrand)
{
    SeedPos *p = (SeedPos *)malloc(((size_t)n + 1) * sizeof(SeedPos)), *aux = (SeedPos *)malloc(((size_t)n + 1) * sizeof(SeedPos));
    for (uint i = 0; i < n; ++i) { p[i].readID = readID[i]; p[i].pos = pos[i]; p[i].strand = strand[i]; }
    p[n].readID = 0x7FFFFFFF; p[n].strand = 0; p[n].pos = 0xFFFFFFFF;          /* array guard, DV-DPfunctions.cu:1212 */
    uint len = n + 1;
    SeedPos *auxPos = aux;
    MC_RadixSort_32_16 ( p, pos, auxPos, len );
    MC_RadixSort_32_16 ( p, readID, auxPos, len );
    MC_RadixSort_8_8 ( p, strand, auxPos, len );
    free(aux);
    vector<CandidateInfo> *c = ref_singleMerge(p);
    uint m = (uint)c->size();
    for (uint i = 0; i < m; ++i) { outReadID[i] = (*c)[i].readID; outPos[i] = (*c)[i].pos; outStrand[i] = (*c)[i].strand; }
    delete c;
    free(p);
    return m;
}

}
