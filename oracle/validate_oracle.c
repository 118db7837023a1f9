/* TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's long-read validation of seed alignments, batched over reads,
 * for checking s3_validate_alignments and the long-read mode of s3_se_align.  Follows
 *   validateAlignments                                   CPUfunctions.cpp:1129-1222
 *   createQueryPackedDNA / createRevQueryPackedDNA       PE.cpp:28-60      (the extension, forward / reverse-complemented)
 *   createTargetPackedDNA + numMismatchNew               PE.cpp:287-325, 148-178   (= Hamming distance over the extension)
 *   seed length of a long read, the mismatch allowance   alignment.cu:2475-2491, CPUfunctions.cpp:1812-1818, definitions.h:140-142
 * statement by statement where order matters: the running best, the reset of the output size, the early stop.
 * Pinned against those functions compiled from the reference by oracle/build_ref.sh (libref_validate.so):
 * tests/test_cpu_oracle_vs_ref.py.
 */
#include <math.h>
#include <stdint.h>

static uint32_t text_base(const uint32_t *pac, uint64_t p) { return (pac[p >> 4] >> (30 - 2 * (p & 15))) & 3u; }

/* One read.  query: base codes of the whole read.  Lists are updated in place; returns the new size. */
uint32_t s3o_validate_one(const uint32_t *pac, uint32_t dnaLength, const uint8_t *query, uint32_t seedLen, uint32_t readLen, uint32_t n,
                          uint32_t *pos, uint8_t *strand, uint8_t *mism, int onlyKeepBest, int minSeedMismatch, int maxMismatch, int maxHitNum)
{
    if (readLen <= seedLen) return n;
    const uint32_t ext = readLen - seedLen;
    int pre = maxMismatch;
    uint32_t newSize = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if ((int)mism[i] < minSeedMismatch) continue;
        int mismatch = maxMismatch + 1;
        if (strand[i] == 1 && (uint64_t)pos[i] + readLen <= dnaLength) {
            mismatch = 0;
            for (uint32_t j = 0; j < ext; ++j) mismatch += text_base(pac, (uint64_t)pos[i] + seedLen + j) != query[seedLen + j];
        }
        if (strand[i] == 2 && pos[i] >= ext) {
            mismatch = 0;
            for (uint32_t j = 0; j < ext; ++j) mismatch += text_base(pac, (uint64_t)pos[i] - ext + j) != (uint32_t)(3 - query[readLen - 1 - j]);
        }
        const int tot = mismatch + (int)mism[i];
        if (tot <= maxMismatch) {
            if (onlyKeepBest && tot > pre) continue;
            if (onlyKeepBest && tot < pre) { newSize = 0; pre = tot; }
            const uint32_t p = strand[i] == 1 ? pos[i] : pos[i] - ext;
            strand[newSize] = strand[i];
            mism[newSize] = (uint8_t)tot;
            pos[newSize] = p;
            newSize++;
            if (pre == minSeedMismatch && (int)newSize >= maxHitNum) break;
        }
    }
    return newSize;
}

/* A batch: CSR lists (off has numReads + 1 entries), reads as rows of maxLen base codes.  Long-read mode of hostKernel
 * (CPUfunctions.cpp:1812-1842): seed length 100 for reads longer than 120, allowance ceil(0.02 * readLen) (doubled when MAPQ is
 * wanted), the list cut to maxHitNum afterwards.  outCount[r] = entries of read r kept, written in place from off[r]. */
void s3o_validate_batch(const uint32_t *pac, uint32_t dnaLength, const uint8_t *reads, uint32_t maxLen, const uint32_t *readLengths, uint64_t numReads,
                        const uint32_t *off, uint32_t *pos, uint8_t *strand, uint8_t *mism, int onlyKeepBest, int minSeedMismatch, int doubleAllowance,
                        int maxHitNum, uint32_t *outCount)
{
    for (uint64_t r = 0; r < numReads; ++r) {
        const uint32_t len = readLengths[r], seed = len > 120 ? 100 : len, n = off[r + 1] - off[r];
        int maxMismatch = (int)ceil(0.02 * len);
        if (doubleAllowance) maxMismatch *= 2;
        uint32_t m = n;
        if (n) m = s3o_validate_one(pac, dnaLength, reads + r * maxLen, seed, len, n, pos + off[r], strand + off[r], mism + off[r], onlyKeepBest,
                                    minSeedMismatch, maxMismatch, maxHitNum);
        if (n && (int)m > maxHitNum) m = (uint32_t)maxHitNum;
        outCount[r] = m;
    }
}
