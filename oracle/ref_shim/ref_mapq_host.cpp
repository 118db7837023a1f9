/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * The reference's own MAPQ functions on the host: the three tables (BGS-IO.cpp:33,42,45), the nine functions
 * (BGS-IO.cpp:2280-2580) and bwase_initialize (CPUfunctions.cpp:3014-3019), cut by sed into mapq.inc; nothing is edited.
 * Pins the s3_mapq_* entries (soap3-dp_b200/csrc/s3_params.cu).
 */
#include <math.h>
#include <string.h>
#include "mapq.inc"

static int g_log_n[256];
static int g_ready = 0;
static int *logn() { if (!g_ready) { memset(g_log_n, 0, sizeof g_log_n); bwase_initialize(g_log_n); g_ready = 1; } return g_log_n; }

extern "C" {
int ref_mapq_unique(int n, int mm, int q, int mx, int mn) { return getMapQualScore(n, mm, q, mx, mn); }
int ref_mapq_bwa_single(int x0, int x1) { return bwaLikeSingleQualScore(x0, x1, logn()); }
int ref_mapq_single(int mm, int q, int x0, int x1, int mx, int mn, int bwa) { return getMapQualScoreSingle(mm, q, x0, x1, mx, mn, bwa, logn()); }
int ref_mapq_single_dp(int maxDP, int q, int x0, int x1a, int x1b, int best, int second, int mx, int mn, int thres, int bwa)
{ return getMapQualScoreForSingleDP(maxDP, q, x0, x1a, x1b, best, second, mx, mn, thres, bwa, logn()); }
void ref_mapq_bwa_pair(int a, int b, int c, int d, int ops, int opn, int sops, int sopn, int l0, int l1, int *m0, int *m1)
{ bwaLikePairQualScore(a, b, c, d, logn(), ops, opn, sops, sopn, l0, l1, m0, m1); }
int ref_mapq_pair_end(int mm, int q, int x0, int x1, int best, unsigned pairs, int mx, int mn) { return getMapQualScore2(mm, q, x0, x1, (char)best, pairs, mx, mn); }
int ref_mapq_unique_dp(int n, int dp, int maxDP, int q, int mx, int mn) { return getMapQualScoreForDP(n, dp, maxDP, q, mx, mn); }
int ref_mapq_pair_end_dp(int dp, int maxDP, int q, int x0, int x1, int best, int second, int isBest, int pairs, int mx, int mn)
{ return getMapQualScoreForDP2(dp, maxDP, q, x0, x1, best, second, (char)isBest, pairs, mx, mn); }
int ref_mapq_of_pair(int a, int b) { return getMapQualScoreForPair(a, b); }
}
