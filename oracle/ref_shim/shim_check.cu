/* TEST INFRASTRUCTURE ONLY -- see oracle/build_ref.sh.
 * Executes the drop-in shim (integration/soap3dp_b200_shim.cpp) the way SOAP3-dp would: through the reference's own
 * declarations -- Soap3Index / SRAIndex / BWT from its unmodified headers, GPUINDEXUpload, perform_round1_alignment, perform_round2_alignment,
 * SemiGlobalAligner::{init, performAlignment, freeMemory}, GPUINDEXFree -- on the case oracle/make_shim_case.py wrote, and
 * compares every answer word, score, hit location, tie count and traced pattern with the oracle's.  Linked against the shim
 * object and libsoap3dp_b200.so; run on a GPU box as oracle/_ref/shim_check <case dir>.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "alignment.h"
#include "DV-DPfunctions.h"
#include "soap3-dp-module.h"
#include "PEAlgnmt.h"

/* defined by the shim (integration/soap3dp_b200_shim.cpp): the results of the two seeded DP stages as the reference's own records */
unsigned int rescueDPAlignResults ( unsigned int * queries, unsigned int * upkdReadLengths, unsigned int numQueries, unsigned int wordPerQuery,
                                    unsigned int maxReadLength, int insert_high, int insert_low, int peStrandLeftLeg, int peStrandRightLeg,
                                    unsigned int numMismatch, unsigned int maxOutputPerRead, unsigned int maxHitNumForDP,
                                    unsigned int * _bwt, DPParameters * dpParameters, AlgnmtDPResult ** results );
unsigned int singleDPAlignResults ( unsigned int * queries, unsigned int * upkdReadLengths, unsigned int numQueries, unsigned int wordPerQuery,
                                    const unsigned int * readIDs, unsigned int numReads, unsigned int * _bwt, DPParameters * dpParameters,
                                    SingleAlgnmtResult ** results, unsigned int ** unseeded, unsigned int * numUnseeded );
unsigned int deepDPAlignResults ( unsigned int * queries, unsigned int * upkdReadLengths, unsigned int numQueries, unsigned int wordPerQuery,
                                  const unsigned int * pairReadIDs, unsigned int numPairs, int insert_high, int insert_low, int peStrandLeftLeg, int peStrandRightLeg,
                                  unsigned int * _bwt, DPParameters * dpParameters, DeepDPAlignResult ** results, unsigned int ** unseeded, unsigned int * numUnseeded );

template <class T> static std::vector<T> load(const std::string &dir, const char *name)
{
    std::string p = dir + "/" + name + ".bin";
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) { printf("FAIL cannot open %s\n", p.c_str()); exit(2); }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v(sz / sizeof(T));
    if (fread(v.data(), 1, sz, f) != (size_t)sz) { printf("FAIL short read %s\n", p.c_str()); exit(2); }
    fclose(f);
    return v;
}

static size_t pattern_bytes(const uchar *p, size_t cap)
{
    size_t i = 0;
    while (i < cap && p[i] != 0) i += p[i] == 'V' ? 2 : 1;       /* a count byte may be 0 */
    return i < cap ? i + 1 : cap;
}

int main(int argc, char **argv)
{
    std::string dir = argc > 1 ? argv[1] : "oracle/_ref/shim_case";
    unsigned textLength, isa0, risa0, nbwt, nocc, n, wpq, k, numCases, allowed, wpa, m, maxRead, maxDNA, patLen, allowed2, wpa2;
    {
        FILE *f = fopen((dir + "/meta.txt").c_str(), "r");
        if (!f || fscanf(f, "%u %u %u %u %u %u %u %u %u %u %u %u %u %u %u %u %u", &textLength, &isa0, &risa0, &nbwt, &nocc, &n, &wpq, &k, &numCases,
                         &allowed, &wpa, &m, &maxRead, &maxDNA, &patLen, &allowed2, &wpa2) != 17) { printf("FAIL meta.txt\n"); return 2; }
        fclose(f);
    }
    std::vector<uint> bwtCode = load<uint>(dir, "bwt"), occ = load<uint>(dir, "occ"), rbwtCode = load<uint>(dir, "rbwt"), rocc = load<uint>(dir, "rocc");
    /* the reference's own structures, only the fields GPUINDEXUpload reads (alignment.cu:27-107) */
    BWT bwt, revBwt; SRAIndex sra; Soap3Index index;
    memset(&bwt, 0, sizeof bwt); memset(&revBwt, 0, sizeof revBwt); memset(&sra, 0, sizeof sra); memset(&index, 0, sizeof index);
    bwt.bwtCode = bwtCode.data(); bwt.textLength = textLength; bwt.inverseSa0 = isa0;
    revBwt.bwtCode = rbwtCode.data(); revBwt.textLength = textLength; revBwt.inverseSa0 = risa0;
    sra.bwt = &bwt; sra.rev_bwt = &revBwt;
    /* hsp->packedDNA and the full suffix array, as INDEXLoad leaves them with SaValueFreq = 1: the shim then uploads them and
     * the search runs with seed tables' straight-line kernel and check-and-extend (S3_SHIM_NO_TEXT=1: without, the stepping search) */
    std::vector<uint> saValue = load<uint>(dir, "sa"), pac = load<uint>(dir, "pac");
    HSP hsp; memset(&hsp, 0, sizeof hsp);
    if (!getenv("S3_SHIM_NO_TEXT")) {
        hsp.packedDNA = pac.data(); hsp.dnaLength = textLength; sra.hsp = &hsp;
        bwt.saValue = saValue.data(); bwt.saInterval = 1;
    }
    index.sraIndex = &sra; index.gpu_occValue = occ.data(); index.gpu_revOccValue = rocc.data(); index.gpu_numOfOccValue = nocc / 4;
    cudaSetDevice(0);
    uint *_bwt = NULL, *_occ = NULL, *_revBwt = NULL, *_revOcc = NULL;
    GPUINDEXUpload(&index, &_bwt, &_occ, &_revBwt, &_revOcc);
    int fails = 0;

    /* round 1, all cases in one call (alignment.cu:118-215) */
    std::vector<uint> queries = load<uint>(dir, "queries"), lengths = load<uint>(dir, "lengths");
    std::vector<std::vector<uint> > got(numCases, std::vector<uint>((size_t)n * wpa, 0u));
    uint *ans[2][MAX_NUM_CASES];                                  /* the caller's double buffer, alignment.cu:700-760; buffer 1 is used */
    memset(ans, 0, sizeof ans);
    for (unsigned c = 0; c < numCases; ++c) ans[1][c] = got[c].data();
    perform_round1_alignment(queries.data(), lengths.data(), ans, k, numCases, allowed, wpq, wpa, false, 1,
                             (n + 127) / 128, n, &index, _bwt, _revBwt, _occ, _revOcc);
    {
        /* the same call again, timed: the caller's malloc'ed buffers are page-locked by the shim since the first call */
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, 0);
        for (int it = 0; it < 5; ++it)
            perform_round1_alignment(queries.data(), lengths.data(), ans, k, numCases, allowed, wpq, wpa, false, 1,
                                     (n + 127) / 128, n, &index, _bwt, _revBwt, _occ, _revOcc);
        cudaEventRecord(e1, 0); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        printf("INFO perform_round1_alignment through the shim: %.3f ms per call of %u reads x %u cases (%s)\n", ms / 5, n, numCases,
               sra.hsp ? "text + suffix array uploaded: straight-line kernel and check-and-extend" : "no text: stepping search");
    }
    for (unsigned c = 0; c < numCases; ++c) {
        char name[32]; snprintf(name, sizeof name, "answers%u", c);
        std::vector<uint> want = load<uint>(dir, name);
        size_t bad = 0, hits = 0;
        for (size_t i = 0; i < want.size(); ++i) bad += want[i] != got[c][i];
        for (unsigned r = 0; r < n; ++r) hits += want[(size_t)(r / 32) * 32 * wpa + r % 32] < 0xFFFFFFFDu;
        printf("%s perform_round1_alignment case %u: %zu words, %zu reads with hits, %zu words differ\n", bad ? "FAIL" : "PASS", c, want.size(), hits, bad);
        fails += bad != 0;
    }

    /* round 2 on what round 1 left (alignment.cu:221-326); like the reference's caller (alignment.cu:898-954) the number of
     * bad reads per case is re-counted from the round-1 status words, the shim hands none back */
    {
        std::vector<std::vector<uint> > badIdx(numCases, std::vector<uint>(n, 0xDEADBEEFu)), badAns(numCases, std::vector<uint>((size_t)n * wpa2, 0u));
        uint *pIdx[2][MAX_NUM_CASES], *pAns[2][MAX_NUM_CASES];
        memset(pIdx, 0, sizeof pIdx); memset(pAns, 0, sizeof pAns);
        for (unsigned c = 0; c < numCases; ++c) { pIdx[1][c] = badIdx[c].data(); pAns[1][c] = badAns[c].data(); }
        perform_round2_alignment(queries.data(), lengths.data(), ans, k, numCases, allowed2, wpq, wpa, wpa2, false, 1,
                                 (n + 127) / 128, n, &index, _bwt, _revBwt, _occ, _revOcc, 0, pIdx, pAns);
        for (unsigned c = 0; c < numCases; ++c) {
            char name[32];
            snprintf(name, sizeof name, "bad_idx%u", c);
            std::vector<uint> wantIdx = load<uint>(dir, name);
            snprintf(name, sizeof name, "bad_ans%u", c);
            std::vector<uint> wantAns = load<uint>(dir, name);
            size_t nb = 0, bad = 0;
            for (unsigned r = 0; r < n; ++r) nb += got[c][(size_t)(r / 32) * 32 * wpa + r % 32] > 0xFFFFFFFDu;
            bad += nb != wantIdx.size();
            for (size_t i = 0; i < wantIdx.size() && i < nb; ++i) bad += badIdx[c][i] != wantIdx[i];
            /* rows of the nb reads only: the padding lanes of the last group of 32 are not part of the contract */
            for (size_t r = 0; r < nb && r < wantIdx.size(); ++r)
                for (unsigned w = 0; w < wpa2; ++w) {
                    size_t at = (r / 32) * 32 * wpa2 + (size_t)w * 32 + r % 32;
                    bad += badAns[c][at] != wantAns[at];
                }
            printf("%s perform_round2_alignment case %u: %zu reads searched again with %u slots, %zu differences\n", bad ? "FAIL" : "PASS", c, nb, allowed2, bad);
            fails += bad != 0;
        }
    }

    /* DP (DV-DPfunctions.cu:520-741) */
    std::vector<uint> dna = load<uint>(dir, "dp_dna"), dnaLen = load<uint>(dir, "dp_dna_len"), rd = load<uint>(dir, "dp_read"), rdLen = load<uint>(dir, "dp_read_len");
    std::vector<int> cutoff = load<int>(dir, "dp_cutoff"), wScores = load<int>(dir, "dp_scores");
    std::vector<uint> clipLt = load<uint>(dir, "dp_clip_lt"), clipRt = load<uint>(dir, "dp_clip_rt"), ancL = load<uint>(dir, "dp_anchor_l"), ancR = load<uint>(dir, "dp_anchor_r");
    std::vector<uint> wHit = load<uint>(dir, "dp_hit"), wCnt = load<uint>(dir, "dp_cnt");
    std::vector<uchar> wPat = load<uchar>(dir, "dp_pattern");
    DPParameters para; memset(&para, 0, sizeof para);
    para.matchScore = 1; para.mismatchScore = -2; para.openGapScore = -3; para.extendGapScore = -1;       /* soap3-dp.ini */
    SemiGlobalAligner aligner;
    int maxDPTableLength = 0, numOfBlocks = 0, patternLength = 0;
    aligner.decideConfiguration((int)maxRead, (int)maxDNA, maxDPTableLength, numOfBlocks, patternLength, para);
    aligner.init((int)wScores.size(), (int)maxRead, (int)maxDNA, maxDPTableLength, para);
    std::vector<int> scores(wScores.size(), 0);
    std::vector<uint> hit(wHit.size(), 0), cnt(wCnt.size(), 0);
    std::vector<uchar> pat(wPat.size(), 0);
    aligner.performAlignment(dna.data(), dnaLen.data(), rd.data(), rdLen.data(), cutoff.data(), scores.data(), hit.data(), cnt.data(), pat.data(), (int)m,
                             clipLt.data(), clipRt.data(), ancL.data(), ancR.data());
    size_t bs = 0, bh = 0, bc = 0, bp = 0, traced = 0;
    for (unsigned t = 0; t < m; ++t) {
        bs += scores[t] != wScores[t]; bh += hit[t] != wHit[t]; bc += cnt[t] != wCnt[t];
        if (wScores[t] >= cutoff[t]) {
            ++traced;
            const uchar *a = &pat[(size_t)t * patLen], *b = &wPat[(size_t)t * patLen];
            bp += memcmp(a, b, pattern_bytes(b, patLen)) != 0;
        }
    }
    printf("%s SemiGlobalAligner::performAlignment: %u alignments (%zu traced, patternLength %d == %u), differ: scores %zu, hitLocs %zu, counts %zu, patterns %zu\n",
           (bs || bh || bc || bp || patternLength != (int)patLen) ? "FAIL" : "PASS", m, traced, patternLength, patLen, bs, bh, bc, bp);
    fails += (bs || bh || bc || bp || patternLength != (int)patLen);
    aligner.freeMemory();
    GPUINDEXFree(_bwt, _occ, _revBwt, _revOcc);

    /* alignSingleR (soap3-dp-module.h:60-74): all valid alignments of every read into the caller's AlgnResultArrays */
    {
        std::vector<uint> wOff = load<uint>(dir, "single_off"), wPos = load<uint>(dir, "single_pos"), wFlags = load<uint>(dir, "single_flags");
        std::vector<uint> readIDs(n);
        for (unsigned q = 0; q < n; ++q) readIDs[q] = 1000 + q;
        SingleAlignParam par; memset(&par, 0, sizeof par);
        par.maxReadLength = 100; par.numMismatch = (int)k; par.outputOption = 1; par.cpuNumThreads = 1; par.maxHitNum = 1000; par.enableDP = 2;
        AlgnResultArrays *arr = resultArraysConstruct(1);
        unsigned long long numOfAnswer = 0; unsigned int numOfAlignedRead = 0;
        alignSingleR(queries.data(), lengths.data(), readIDs.data(), wpq, n, &index, &par, numOfAnswer, numOfAlignedRead, arr);
        AlgnResult *res = arr->algnArrays[0];
        size_t bad = 0, aligned = 0, overflow = 0;
        /* reads whose round-1 slot overflowed in some case are searched again by the reference (round 2, CPU): the chain reports them
           apart, so they are left out of the comparison -- counted from the round-1 answers */
        std::vector<char> skip(n, 0);
        for (unsigned c = 0; c < numCases; ++c) {
            std::vector<uint> a = load<uint>(dir, (std::string("answers") + char('0' + c)).c_str());
            for (unsigned q = 0; q < n; ++q) if (a[(size_t)(q / 32) * 32 * wpa + q % 32] > 0xFFFFFFFDu) skip[q] = 1;
        }
        size_t g = 0;
        for (unsigned q = 0; q < n; ++q) {
            if (skip[q]) { ++overflow; while (g < res->occTotalNum && res->occ_list[g].readID == readIDs[q]) ++g; continue; }
            aligned += wOff[q + 1] > wOff[q];
            for (uint t = wOff[q]; t < wOff[q + 1]; ++t, ++g) {
                if (g >= res->occTotalNum) { ++bad; continue; }
                const occRec &o = res->occ_list[g];
                bad += o.readID != readIDs[q] || o.ambPosition != wPos[t] || o.strand != (wFlags[t] & 0xFF) || o.source != 1 || o.score != (char)(wFlags[t] >> 8);
            }
        }
        bad += g != res->occTotalNum;
        printf("%s alignSingleR: %u reads (%zu aligned, %zu left to round 2), %u occRec records, numOfAnswer %llu, numOfAlignedRead %u, %zu differences\n",
               bad ? "FAIL" : "PASS", n, aligned, overflow, res->occTotalNum, numOfAnswer, numOfAlignedRead, bad);
        fails += bad != 0;
        resultArraysFree(arr);
    }
    /* deepDPAlignResults (the shim's results of DPForUnalignPairs2, DV-DPForBothUnalign.cu:245): DeepDPAlignResult records as the reference's
       engine builds them, against oracle/seeding_oracle.deep_dp */
    {
        unsigned dn = 0, dpairs = 0, nrec = 0, nuns = 0;
        FILE *f = fopen((dir + "/deep_meta.txt").c_str(), "r");
        if (!f || fscanf(f, "%u %u %u %u", &dn, &dpairs, &nrec, &nuns) != 4) { printf("FAIL deep_meta.txt\n"); return 2; }
        fclose(f);
        std::vector<uint> dq = load<uint>(dir, "deep_queries"), dl = load<uint>(dir, "deep_lengths"), ids = load<uint>(dir, "deep_ids");
        std::vector<int> rec = load<int>(dir, "deep_records");
        std::vector<uchar> cig = load<uchar>(dir, "deep_cigars");
        std::vector<uint> coff = load<uint>(dir, "deep_cigar_off"), wuns = load<uint>(dir, "deep_unseeded");
        uint *_b, *_o, *_rb, *_ro;
        GPUINDEXUpload(&index, &_b, &_o, &_rb, &_ro);
        DPParameters dpp; memset(&dpp, 0, sizeof dpp);
        dpp.matchScore = 1; dpp.mismatchScore = -2; dpp.openGapScore = -3; dpp.extendGapScore = -1; dpp.softClipLeft = 3; dpp.softClipRight = 8;
        dpp.paramRead[0].cutoffThreshold = dpp.paramRead[1].cutoffThreshold = 0;          /* default: ceil(0.3 x read length) */
        DeepDPAlignResult *res = NULL; unsigned int *uns = NULL, gotUns = 0;
        unsigned int got = deepDPAlignResults(dq.data(), dl.data(), dn, wpq, ids.data(), dpairs, 500, 200, 1, 2, _b, &dpp, &res, &uns, &gotUns);
        size_t bad = (got != nrec) + (gotUns != nuns);
        for (unsigned h = 0; h < got && h < nrec; ++h) {
            const int *w = rec.data() + 12 * (size_t)h;
            const DeepDPAlignResult &r = res[h];
            const std::string c1((const char *)cig.data() + coff[2 * h], coff[2 * h + 1] - coff[2 * h]), c2((const char *)cig.data() + coff[2 * h + 1], coff[2 * h + 2] - coff[2 * h + 1]);
            bad += (int)r.readID != w[0] || r.insertSize != w[1] || (int)r.algnmt_1 != w[2] || r.strand_1 != w[3] || r.score_1 != w[4] || r.editdist_1 != w[5] || r.num_sameScore_1 != w[6] ||
                   (int)r.algnmt_2 != w[7] || r.strand_2 != w[8] || r.score_2 != w[9] || r.editdist_2 != w[10] || r.num_sameScore_2 != w[11] ||
                   c1 != r.cigarString_1 || c2 != r.cigarString_2;
        }
        for (unsigned u = 0; u < gotUns && u < nuns; ++u) bad += uns[u] != wuns[u];
        printf("%s deepDPAlignResults (DPForUnalignPairs2): %u pairs, %u DeepDPAlignResult records (position, strand, score, edit distance, ties, CIGAR of both reads, insert size), "
               "%u pairs without a candidate, %zu differences\n", bad ? "FAIL" : "PASS", dpairs, got, gotUns, bad);
        fails += bad != 0;
        for (unsigned h = 0; h < got; ++h) { free(res[h].cigarString_1); free(res[h].cigarString_2); }
        free(res); free(uns);
        GPUINDEXFree(_b, _o, _rb, _ro);
    }
    /* singleDPAlignResults (the shim's results of DPForUnalignSingle2, DV-DPForSingleReads.cu:155): SingleAlgnmtResult records as the reference's
       engine builds them, against oracle/seeding_oracle.single_dp */
    {
        unsigned sn = 0, swpq = 0, nrec = 0, nuns = 0;
        FILE *f = fopen((dir + "/sdp_meta.txt").c_str(), "r");
        if (!f || fscanf(f, "%u %u %u %u", &sn, &swpq, &nrec, &nuns) != 4) { printf("FAIL sdp_meta.txt\n"); return 2; }
        fclose(f);
        std::vector<uint> sq = load<uint>(dir, "sdp_queries"), sl = load<uint>(dir, "sdp_lengths");
        std::vector<int> rec = load<int>(dir, "sdp_records");
        std::vector<uchar> cig = load<uchar>(dir, "sdp_cigars");
        std::vector<uint> coff = load<uint>(dir, "sdp_cigar_off"), wuns = load<uint>(dir, "sdp_unseeded");
        std::vector<uint> ids(sn);
        for (unsigned q = 0; q < sn; ++q) ids[q] = q;
        uint *_b, *_o, *_rb, *_ro;
        GPUINDEXUpload(&index, &_b, &_o, &_rb, &_ro);
        DPParameters dpp; memset(&dpp, 0, sizeof dpp);
        dpp.matchScore = 1; dpp.mismatchScore = -2; dpp.openGapScore = -3; dpp.extendGapScore = -1; dpp.softClipLeft = 3; dpp.softClipRight = 8;
        SingleAlgnmtResult *res = NULL; unsigned int *uns = NULL, gotUns = 0;
        unsigned int got = singleDPAlignResults(sq.data(), sl.data(), sn, swpq, ids.data(), sn, _b, &dpp, &res, &uns, &gotUns);
        size_t bad = (got != nrec) + (gotUns != nuns);
        for (unsigned h = 0; h < got && h < nrec; ++h) {
            const int *w = rec.data() + 6 * (size_t)h;
            const SingleAlgnmtResult &r = res[h];
            const std::string c((const char *)cig.data() + coff[h], coff[h + 1] - coff[h]);
            bad += (int)r.readID != w[0] || r.strand != w[1] || (int)r.algnmt != w[2] || r.score != w[3] || r.editdist != w[4] || r.num_sameScore != w[5] || c != r.cigarString;
        }
        for (unsigned u = 0; u < gotUns && u < nuns; ++u) bad += uns[u] != wuns[u];
        printf("%s singleDPAlignResults (DPForUnalignSingle2): %u reads, %u SingleAlgnmtResult records (position, strand, score, edit distance, ties, CIGAR), "
               "%u reads without a candidate, %zu differences\n", bad ? "FAIL" : "PASS", sn, got, gotUns, bad);
        fails += bad != 0;
        for (unsigned h = 0; h < got; ++h) free(res[h].cigarString);
        free(res); free(uns);
        GPUINDEXFree(_b, _o, _rb, _ro);
    }
    /* rescueDPAlignResults (the shim's results of semiGlobalDP2 through the whole paired-end chain): AlgnmtDPResult records as DP_Space::algnmtCPUThread
       builds them, one per rescue window, against oracle/pe_chain_oracle.pe_chain */
    {
        unsigned rn = 0, nrec = 0, maxHit = 0;
        FILE *f = fopen((dir + "/rescue_meta.txt").c_str(), "r");
        if (!f || fscanf(f, "%u %u %u", &rn, &nrec, &maxHit) != 3) { printf("FAIL rescue_meta.txt\n"); return 2; }
        fclose(f);
        std::vector<uint> rq = load<uint>(dir, "rescue_queries"), rl = load<uint>(dir, "rescue_lengths");
        std::vector<int> rec = load<int>(dir, "rescue_records");
        std::vector<uchar> cig = load<uchar>(dir, "rescue_cigars");
        std::vector<uint> coff = load<uint>(dir, "rescue_cigar_off");
        uint *_b, *_o, *_rb, *_ro;
        GPUINDEXUpload(&index, &_b, &_o, &_rb, &_ro);
        DPParameters dpp; memset(&dpp, 0, sizeof dpp);
        dpp.matchScore = 1; dpp.mismatchScore = -2; dpp.openGapScore = -3; dpp.extendGapScore = -1; dpp.softClipLeft = 3; dpp.softClipRight = 8;
        AlgnmtDPResult *res = NULL;
        unsigned int got = rescueDPAlignResults(rq.data(), rl.data(), rn, wpq, 100, 500, 200, 1, 2, k, 1000, maxHit, _b, &dpp, &res);
        size_t bad = got != nrec, traced = 0;
        for (unsigned t = 0; t < got && t < nrec; ++t) {
            const int *w = rec.data() + 11 * (size_t)t;
            const AlgnmtDPResult &r = res[t];
            bad += (int)r.readID != w[0] || r.whichFromDP != w[1] || (int)r.algnmt_1 != w[2] || (int)r.algnmt_2 != w[3] || r.strand_1 != w[4] || r.strand_2 != w[5] ||
                   r.score_1 != w[6] || r.score_2 != w[7];
            if (w[1] != 2) {       /* (the reference leaves edit distance, insert size and tie count of a miss unset) */
                const std::string c((const char *)cig.data() + coff[t], coff[t + 1] - coff[t]);
                bad += r.editdist != w[8] || r.insertSize != w[9] || r.num_sameScore != w[10] || !r.cigarString || c != r.cigarString;
                ++traced;
            } else bad += r.cigarString != NULL;
        }
        printf("%s rescueDPAlignResults (semiGlobalDP2 through the paired-end chain): %u reads, %u AlgnmtDPResult records (%zu with a CIGAR), %zu differences\n",
               bad ? "FAIL" : "PASS", rn, got, traced, bad);
        fails += bad != 0;
        for (unsigned t = 0; t < got; ++t) free(res[t].cigarString);
        free(res);
        GPUINDEXFree(_b, _o, _rb, _ro);
    }
    printf("%s drop-in shim executed through the reference's declarations\n", fails ? "FAIL" : "PASS");
    return fails ? 1 : 0;
}
