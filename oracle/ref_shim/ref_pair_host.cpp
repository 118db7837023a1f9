/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * The reference's own paired-end pairing on the host: PERadixSort, PEMappingCore, PEValidateAndPreparePEInput,
 * PEMappingOccurrences, PEIsPairEndMatch, PEIsPairOutOfRange, PEReportPairResult, the PEInput / PEOutput constructors
 * and PEStatsPEPairList, cut by sed from PEAlgnmt.cpp (:45-57, 114-361, 480-637, 645-711, 777-838) into pair.inc and
 * compiled against the reference's unmodified PEAlgnmt.h.  Nothing in them is edited.  Pins oracle/pair_oracle.c.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "PEAlgnmt.h"
#include "pair.inc"

extern "C" {

/* one read pair: the two occurrence lists as the caller has them (arrival order) -> records in emission order.
 * Returns the number of pairs (PEStatsPEOutput's count); opt / sub = index of the optimal / suboptimal record or -1. */
int ref_pair_occurrences(const unsigned *pos1, const unsigned char *strand1, const unsigned char *mism1, unsigned n1,
                         const unsigned *pos2, const unsigned char *strand2, const unsigned char *mism2, unsigned n2,
                         int patternLength, int insertLbound, int insertUbound, int strandLeftLeg, int strandRightLeg, int outputType,
                         unsigned *outPos1, unsigned *outPos2, unsigned *outInsertion, unsigned char *outFlags, unsigned outCap,
                         int *opt, int *sub, unsigned *mismatchStats /* 30 */)
{
    SRAOccurrence *a = (SRAOccurrence *)calloc(n1 + 1, sizeof(SRAOccurrence)), *b = (SRAOccurrence *)calloc(n2 + 1, sizeof(SRAOccurrence));
    for (unsigned i = 0; i < n1; ++i) { a[i].ambPosition = pos1[i]; a[i].strand = strand1[i]; a[i].mismatchCount = mism1[i]; }
    for (unsigned i = 0; i < n2; ++i) { b[i].ambPosition = pos2[i]; b[i].strand = strand2[i]; b[i].mismatchCount = mism2[i]; }
    PEInput *in = PEInputConstruct((BWT *)1, (HSP *)1);         /* only checked against NULL */
    PEOutput *out = PEOutputConstruct();
    in->OutputType = outputType; in->insertLbound = insertLbound; in->insertUbound = insertUbound;
    in->strandLeftLeg = strandLeftLeg; in->strandRightLeg = strandRightLeg;
    in->patternLength = patternLength;                          /* CPUfunctions.cpp:2284 */
    PEMappingOccurrences(in, out, a, n1, b, n2);
    PEPairs *optimal = NULL, *suboptimal = NULL;
    memset(mismatchStats, 0, 30 * sizeof(unsigned));
    unsigned count = PEStatsPEOutput(out, &optimal, &suboptimal, mismatchStats);
    unsigned k = 0;
    *opt = *sub = -1;
    for (PEPairList *l = out->root; l != NULL && l->pairsCount > 0; l = l->next)
        for (unsigned i = 0; i < l->pairsCount; ++i, ++k) {
            PEPairs *p = &l->pairs[i];
            if (p == optimal) *opt = (int)k;
            if (p == suboptimal) *sub = (int)k;
            if (k < outCap) {
                outPos1[k] = p->algnmt_1; outPos2[k] = p->algnmt_2; outInsertion[k] = (unsigned)p->insertion;
                outFlags[4 * k] = p->strand_1; outFlags[4 * k + 1] = p->mismatch_1; outFlags[4 * k + 2] = p->strand_2; outFlags[4 * k + 3] = p->mismatch_2;
            }
        }
    PEOutputFree(out); PEInputFree(in); free(a); free(b);
    return (int)count;
}

}
