/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * The reference's own DP result decoding, run on the host: CigarStringEncoder (DV-DPfunctions.h:514-529,545-597), the
 * result loop of SingleDP_Space::algnmtCPUThread (DV-DPfunctions.cu:1699-1733), writeNumToStr (PE.cpp:83-110) and
 * convertToCigarStr (PE.cpp:420-485), cut out of the reference files by sed at build time into decode.inc /
 * decode_loop.inc / sam_cigar.inc.  Nothing in them is edited; the loop is pasted into a function whose locals carry
 * the names the loop reads (batch, engine, resultBatch, the four scores, cutoffThreshold).
 * It pins s3_dp_decode (soap3-dp_b200/csrc/s3_decode.cu) and the Python restatement in tests/helpers.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <math.h>
#include <vector>
using std::vector;
typedef unsigned int uint;
typedef unsigned char uchar;
#include "decode.inc"
#include "sam_cigar.inc"

namespace {
/* lookalikes of the members the loop touches (SingleAlgnmtResult PEAlgnmt.h:384-394, CandidateInfo DV-DPfunctions.h:919-926,
 * SingleEndAlgnBatch DV-DPfunctions.h:1010-1040) */
struct SingleAlgnmtResult { uint readID; char strand; uint algnmt; int score; char *cigarString; int editdist; uint num_sameScore; };
typedef vector<SingleAlgnmtResult> SingleDPResultBatch;
struct CanInfo { uint readID; uint pos; char strand; };
struct Batch { int numOfThreads; int *scores; uchar *pattern; CanInfo *canInfos; uint *hitLocs; uint *lengths; uint *maxScoreCounts; };
struct Engine { int patternLength; };
}

extern "C" {

/* Runs the reference loop over the batch; survivors (score >= cutoff) come back compacted, in order:
 * outIndex[k] = the alignment, cigars = the special CIGAR strings joined by '\n'.  Returns their number. */
int ref_dp_decode(const uchar *pattern, int patternLength, const int *scores, const uint *hitLocs, const uint *lengths,
                  const uint *positions, const uint *maxScoreCounts, int n, int cutoffThreshold,
                  int matchScore, int mismatchScore, int openGapScore, int extendGapScore,
                  int *outIndex, uint *outAlgnmt, int *outEditdist, uint *outSameScore, char *cigars, size_t cigarCap)
{
    vector<CanInfo> ci(n);
    for (int i = 0; i < n; ++i) { ci[i].readID = (uint)i; ci[i].pos = positions[i]; ci[i].strand = 1; }
    Batch b = { n, (int *)scores, (uchar *)pattern, ci.data(), (uint *)hitLocs, (uint *)lengths, (uint *)maxScoreCounts };
    Engine e = { patternLength };
    Batch *batch = &b; Engine *engine = &e;
    SingleDPResultBatch *resultBatch = new SingleDPResultBatch;
#include "decode_loop.inc"
    size_t at = 0; int m = (int)resultBatch->size();
    for (int k = 0; k < m; ++k) {
        SingleAlgnmtResult &r = (*resultBatch)[k];
        outIndex[k] = (int)r.readID; outAlgnmt[k] = r.algnmt; outEditdist[k] = r.editdist; outSameScore[k] = r.num_sameScore;
        size_t l = strlen(r.cigarString);
        if (at + l + 2 > cigarCap) { m = -1; break; }
        memcpy(cigars + at, r.cigarString, l); at += l; cigars[at++] = '\n';
        free(r.cigarString);
    }
    cigars[at] = 0;
    delete resultBatch;
    return m;
}

/* convertToCigarStr on one special CIGAR; returns the length written to out (MAX_READ_LENGTH-sized in the reference) */
int ref_convert_cigar(const char *special, char *out)
{
    return convertToCigarStr((char *)special, out, NULL);
}

}
