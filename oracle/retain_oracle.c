/* TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's best-hit filters on one read's SA-range list and
 * occurrence list, batched over reads, for checking s3_retain_best.  Follows, statement by statement (the running
 * minimum, the resets of both list sizes, the order-dependent cap):
 *   retainAllBest            SAList.cpp:140-207
 *   retainAllBestWithCap     SAList.cpp:209-288
 *   retainAllBestAndSecBest  SAList.cpp:290-348
 * SRAOccurrence.mismatchCount is a uint8_t (2bwt-flex/SRACore.h:86-95; the char version in PEAlgnmt.h is commented out),
 * so counts >= 128 compare as large, not negative.
 * Pinned against those functions compiled from the reference by oracle/build_ref.sh (libref_retain.so):
 * tests/test_cpu_oracle_vs_ref.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* mode 0 / 1 / 2 as above.  Output lists are CSR over reads (outSaOff / outOccOff have numReads + 1 entries); the kept
 * entries of read r are written from outSaOff[r] / outOccOff[r] on; flags = strand, mismatchCount.  num[r] = the
 * function's return value.  Output arrays must hold as many entries as the input lists. */
void s3o_retain_best(int mode, int32_t maxNum,
                     const uint32_t *saL, const uint32_t *saR, const uint8_t *saStrand, const uint8_t *saMism, const uint64_t *saOff,
                     const uint32_t *occPos, const uint8_t *occStrand, const uint8_t *occMism, const uint64_t *occOff, uint64_t numReads,
                     uint64_t *outSaOff, uint32_t *outSaL, uint32_t *outSaR, uint8_t *outSaFlags,
                     uint64_t *outOccOff, uint32_t *outOccPos, uint8_t *outOccFlags, uint32_t *num)
{
    uint64_t sBase = 0, oBase = 0;
    for (uint64_t r = 0; r < numReads; ++r) {
        int minMatch = 999;
        int newSa = 0, newOcc = 0;
        unsigned int n = 0;
        outSaOff[r] = sBase; outOccOff[r] = oBase;
#define KEEP_SA(i, cnt) do { outSaL[sBase + newSa] = saL[i]; outSaR[sBase + newSa] = (uint32_t)(cnt) + saL[i] - 1;           \
                             outSaFlags[2 * (sBase + newSa)] = saStrand[i]; outSaFlags[2 * (sBase + newSa) + 1] = saMism[i]; newSa++; } while (0)
#define KEEP_OCC(i) do { outOccPos[oBase + newOcc] = occPos[i]; outOccFlags[2 * (oBase + newOcc)] = occStrand[i];             \
                         outOccFlags[2 * (oBase + newOcc) + 1] = occMism[i]; newOcc++; } while (0)
        if (mode == 2) {
            for (uint64_t i = saOff[r]; i < saOff[r + 1]; ++i) if (saMism[i] < minMatch) minMatch = saMism[i];
            for (uint64_t i = occOff[r]; i < occOff[r + 1]; ++i) if ((int)occMism[i] < minMatch) minMatch = (int)occMism[i];
            for (uint64_t i = saOff[r]; i < saOff[r + 1]; ++i)
                if (saMism[i] <= minMatch + 1) { int c = saR[i] - saL[i] + 1; KEEP_SA(i, c); n += c; }
            for (uint64_t i = occOff[r]; i < occOff[r + 1]; ++i)
                if ((int)occMism[i] <= minMatch + 1) { KEEP_OCC(i); n++; }
        } else {
            for (uint64_t i = saOff[r]; i < saOff[r + 1]; ++i) {
                if (saMism[i] < minMatch) {
                    minMatch = saMism[i];
                    newSa = 0;
                    int c = saR[i] - saL[i] + 1;
                    if (mode == 1 && c > maxNum) c = maxNum;
                    KEEP_SA(i, c);
                    n = c;
                } else if (saMism[i] == minMatch && (mode == 0 || n < (unsigned int)maxNum)) {
                    int c = saR[i] - saL[i] + 1;
                    if (mode == 1 && n + c > (unsigned int)maxNum) c = maxNum - n;
                    KEEP_SA(i, c);
                    n += c;
                }
            }
            for (uint64_t i = occOff[r]; i < occOff[r + 1]; ++i) {
                int mm = (int)occMism[i];
                if (mm < minMatch) {
                    minMatch = mm;
                    newSa = 0; newOcc = 0;
                    KEEP_OCC(i);
                    n = 1;
                } else if (mm == minMatch && (mode == 0 || n < (unsigned int)maxNum)) {
                    KEEP_OCC(i);
                    n++;
                }
            }
        }
#undef KEEP_SA
#undef KEEP_OCC
        num[r] = n;
        sBase += newSa; oBase += newOcc;
    }
    outSaOff[numReads] = sBase; outOccOff[numReads] = oBase;
}
