/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * The reference's own MD-string builder on the host: writeNumToStr (PE.cpp:83-110) and getMisInfoForDP (:499-666), cut by
 * sed into md.inc and compiled against the reference's unmodified PE.h (HSP, dnaChar, soap3DnaComplement).  Nothing in
 * them is edited.  Pins s3_dp_md and oracle/decode_oracle.py:md_string.
 */
#include <string.h>
#include <math.h>
#include "PE.h"
#include "md.inc"

extern "C" int ref_md(const unsigned *packedDNA, const unsigned char *query, const char *qualities, unsigned queryLength, unsigned pos, int strand,
                      const char *specialCigar, char *mdOut, int *numMismatch, int *gapOpen, int *gapExt, int *avgQual)
{
    HSP hsp;
    memset(&hsp, 0, sizeof hsp);
    hsp.packedDNA = (unsigned *)packedDNA;
    return getMisInfoForDP(&hsp, (unsigned char *)query, (char *)qualities, queryLength, pos, (char)strand, (char *)specialCigar, mdOut,
                           numMismatch, gapOpen, gapExt, avgQual, 0);
}
