/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * The reference's own best-hit filters on the host: SAListConstruct / addSAToSAList / SAListFree (SAList.cpp:26-69),
 * retainAllBest, retainAllBestWithCap, retainAllBestAndSecBest (:140-348) and the OCC list helpers (:350-390), cut by sed
 * into retain.inc and compiled against the reference's unmodified SAList.h.  Pins oracle/retain_oracle.c.
 */
#include "SAList.h"
#include "retain.inc"

extern "C" {

/* one read: lists in, filtered lists out (in place in the reference; copied out here).  Returns the function's value. */
unsigned ref_retain_best(int mode, int maxNum, const unsigned *saL, const unsigned *saR, const unsigned char *saStrand, const unsigned char *saMism, unsigned nSa,
                         const unsigned *occPos, const unsigned char *occStrand, const unsigned char *occMism, unsigned nOcc,
                         unsigned *outSaL, unsigned *outSaR, unsigned char *outSaFlags, unsigned *keptSa,
                         unsigned *outOccPos, unsigned char *outOccFlags, unsigned *keptOcc)
{
    SAList *s = SAListConstruct();
    OCCList *o = OCCListConstruct();
    for (unsigned i = 0; i < nSa; ++i) addSAToSAList(s, saL[i], saR[i], saStrand[i], saMism[i]);
    for (unsigned i = 0; i < nOcc; ++i) addToOCCList(o, occPos[i], (char)occStrand[i], (char)occMism[i]);
    unsigned num = mode == 0 ? retainAllBest(s, o) : mode == 1 ? retainAllBestWithCap(s, o, maxNum) : retainAllBestAndSecBest(s, o);
    for (unsigned i = 0; i < s->curr_size; ++i) {
        outSaL[i] = s->sa[i].saIndexLeft; outSaR[i] = s->sa[i].saIndexRight;
        outSaFlags[2 * i] = s->sa[i].strand; outSaFlags[2 * i + 1] = s->sa[i].mismatchCount;
    }
    for (unsigned i = 0; i < o->curr_size; ++i) {
        outOccPos[i] = (unsigned)o->occ[i].ambPosition;
        outOccFlags[2 * i] = o->occ[i].strand; outOccFlags[2 * i + 1] = o->occ[i].mismatchCount;
    }
    *keptSa = s->curr_size; *keptOcc = o->curr_size;
    SAListFree(s); OCCListFree(o);
    return num;
}

}
