// CPU tier: soap3-dp_b200/csrc/s3_windows.cuh (the window selection both s3_dp_make_windows and the paired-end chain
// run on the device) compiled with the host compiler, one call per candidate like the kernels make them.
#include "s3_windows.cuh"

extern "C" {

static S3WinParams params(int insLow, int insHigh, int leftLeg, int rightLeg, int clipLeft, int clipRight, int cutoff0, int cutoff1, unsigned maxDNA, unsigned textLength)
{
    S3WinParams w;
    w.insertLow = insLow; w.insertHigh = insHigh; w.leftLeg = leftLeg; w.rightLeg = rightLeg; w.softClipLeft = clipLeft; w.softClipRight = clipRight;
    w.cutoff[0] = cutoff0; w.cutoff[1] = cutoff1; w.maxDNALength = maxDNA; w.textLength = textLength;
    return w;
}
static void put(const S3Window &x, unsigned *o) { o[0] = x.readID; o[1] = x.start; o[2] = x.dnaLen; o[3] = x.readLen; o[4] = x.clipLt; o[5] = x.clipRt; o[6] = x.ancL; o[7] = x.ancR; o[8] = (unsigned)x.cutoff; o[9] = x.strand; o[10] = x.leftOrRight; }

// mode 1..4 as S3_WIN_*; out: 11 words per window (two windows for mode 2); returns the number of windows
int harness_window(int mode, int insLow, int insHigh, int leftLeg, int rightLeg, int clipLeft, int clipRight, int cutoff0, int cutoff1, unsigned maxDNA,
                   unsigned textLength, const unsigned *readLengths, unsigned readID, unsigned pos, unsigned pos2, int strand, int leftScore,
                   unsigned leftStart, unsigned leftHit, unsigned *out)
{
    const S3WinParams w = params(insLow, insHigh, leftLeg, rightLeg, clipLeft, clipRight, cutoff0, cutoff1, maxDNA, textLength);
    S3Window x[2];
    int k = 0;
    if (mode == 1) { s3_win_single(w, readID, pos, strand, readLengths[readID], x[0]); k = 1; }
    else if (mode == 2) k = s3_win_half(w, readID, pos, strand, readLengths[readID], readLengths[readID ^ 1u], x);
    else if (mode == 3) { s3_win_pair_left(w, readID, pos, readLengths[readID], x[0]); k = 1; }
    else if (leftScore >= s3_win_cutoff(w, readID, readLengths[readID])) { s3_win_pair_right(w, readID, pos2, leftStart + leftHit, readLengths[readID ^ 1u], x[0]); k = 1; }
    for (int j = 0; j < k; ++j) put(x[j], out + 11 * j);
    return k;
}

}
