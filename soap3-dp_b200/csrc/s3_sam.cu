// s3_sam.cu -- the SAM record of a properly paired read pair (host code, like the writers it replaces).
//
// s3_sam_pair_records replaces pairOutputSAMAPI (BGS-IO.cpp:3478-3793) with what it calls: position -> (chromosome, offset)
// through the translate table (getChrAndPos, :1746-1776), the boundary check that trims an alignment hanging over a chromosome
// or segment end into <a>M<b>S / <a>S<b>M (BoundaryCheck / getChrAndPosWithBoundaryCheck, :1779-2007), the MD string of a
// gap-free alignment (getMdStr + mdStr, PE.cpp:209-285,374-419), the pair's mapping qualities (s3_mapq_bwa_pair or s3_mapq_pair_end +
// s3_mapq_of_pair), the XA:Z list of the other pairs with the same number of mismatches, flags, mate fields, insert size, and
// the record body of initializeSAMAlgnmt2 (:2136-2278): name, CIGAR, 4-bit bases (reverse-complemented on the reverse strand),
// qualities, tags RG NM X0 X1 XM XO XG MD XA in that order -- byte for byte bam1_t.core + bam1_t.data as samwrite receives them.
// (l_aux is what the reference leaves there: bam_aux_append and initializeSAMAlgnmt2 both add the tags' bytes.)
#include "s3_common.cuh"
#include "../../include/soap3dp_b200.h"

#include <math.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace {

const char kDnaChar[4] = {'A', 'C', 'G', 'T'};
// bam_nt16_table of the four letters (samtools-0.1.18/bam_import.c:24-41): A 1, C 2, G 4, T 8
const uint8_t kNt16[4] = {1, 2, 4, 8};

int write_num(long long num, char *str)                 // writeNumToStr / writeULLToStr (PE.cpp:62-110): no leading zeros, "0" for 0
{
    char tmp[24];
    int n = 0;
    bool neg = num < 0;
    unsigned long long v = neg ? (unsigned long long)(-num) : (unsigned long long)num;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    int k = 0;
    if (neg) str[k++] = '-';
    while (n) str[k++] = tmp[--n];
    return k;
}

// getChrAndPos: -> end of the segment the position lies in
uint32_t chr_and_pos(const s3_sam_genome *g, uint32_t ambPos, unsigned long long *tp, uint32_t *chr)
{
    uint32_t v = g->ambiguityMap[ambPos >> 18];
    while (g->segments[v].startPos > ambPos) --v;
    *tp = (unsigned long long)(uint32_t)(ambPos - g->segments[v].correction);
    *chr = g->segments[v].chrID;
    return v < g->numSegments - 1 ? g->segments[v + 1].startPos - 1 : g->dnaLength;
}

// getChrAndPosWithBoundaryCheck: trimmed bases (> 0 left, < 0 right, 0 none); cigar = the replacement CIGAR when trimmed
int chr_and_pos_checked(const s3_sam_genome *g, uint32_t readLength, uint32_t ambPos, unsigned long long *tp, uint32_t *chr, std::string &cigar)
{
    uint32_t segEnd = chr_and_pos(g, ambPos, tp, chr);
    const uint32_t chrEnd = g->chrEndPos[*chr - 1];
    segEnd = chrEnd < segEnd ? chrEnd : segEnd;
    if (ambPos + readLength <= segEnd + 1) return 0;
    const int aligned = (int)(segEnd - ambPos + 1);
    char buf[40];
    if (aligned >= ((int)readLength + 1) / 2) {
        snprintf(buf, sizeof buf, "%dM%dS", aligned, (int)readLength - aligned);
        cigar = buf;
        return -((int)readLength - aligned);
    }
    snprintf(buf, sizeof buf, "%dS%dM", aligned, (int)readLength - aligned);
    cigar = buf;
    const uint32_t corrected = segEnd + 1;
    if (corrected > ambPos) chr_and_pos(g, corrected, tp, chr);
    return aligned;
}

inline uint32_t text_base(const uint32_t *pac, unsigned long long p) { return (pac[p >> 4] >> (30 - 2 * (p & 15))) & 3u; }

// getMdStr: MD of the (trimmed) read against the text, the mean quality at the mismatches
int md_string(const s3_sam_genome *g, const uint8_t *query, const char *qualities, uint32_t len, uint32_t pos, int strand, int mismatchNum, int trim,
              std::string &md, int *avgQual)
{
    if (strand == 1) {
        if (trim > 0) { query += trim; pos += trim; len -= trim; }
        else if (trim < 0) len -= -trim;
    } else {
        if (trim > 0) { pos += trim; len -= trim; }
        else if (trim < 0) { query += -trim; len -= -trim; }
    }
    *avgQual = 20;                                       // DEFAULT_QUAL_VALUE
    md.clear();
    char num[24];
    if ((char)mismatchNum == 0) { md.assign(num, write_num(len, num)); return (int)md.size(); }
    double sum = 0.0;
    int pre = -1;
    for (uint32_t i = 0; i < len; ++i) {
        const uint32_t q = strand == 2 ? 3u - query[len - 1 - i] : query[i];
        const uint32_t t = text_base(g->packedDNA, (unsigned long long)pos + i);
        if (q == t) continue;
        md.append(num, write_num((int)i - pre - 1, num));
        md.push_back(kDnaChar[t]);
        pre = (int)i;
        sum += qualities[i];                             // the reference indexes the qualities by alignment column, whatever the strand
    }
    md.append(num, write_num((int)len - pre - 1, num));
    *avgQual = (int)(sum / (char)mismatchNum);
    return (int)md.size();
}

void put32(std::vector<uint8_t> &d, uint32_t v) { for (int k = 0; k < 4; ++k) d.push_back((uint8_t)(v >> (8 * k))); }
void put_tag(std::vector<uint8_t> &d, const char *tag, char type, const void *data, size_t len)
{
    d.push_back((uint8_t)tag[0]); d.push_back((uint8_t)tag[1]); d.push_back((uint8_t)type);
    const uint8_t *p = (const uint8_t *)data;
    d.insert(d.end(), p, p + len);
}

// initializeSAMAlgnmt2 (mapped) / initializeSAMAlgnmt (unmapped): the body of the record
void record_body(s3_sam_record &r, std::vector<uint8_t> &d, int readlen, const char *name, const uint8_t *seq, const char *qual, int strand,
                 const std::string &xa, const std::string *cigar, bool unmapped, int mismatchNum, int editDist, int x0, int x1, int gapOpen, int gapExt,
                 const std::string &md, int mapq, const char *readGroup, bool printMDNM)
{
    d.clear();
    d.reserve(2 * (size_t)readlen + xa.size() + md.size() + 192);
    r.bin = 0;                                           // bam_reg2bin(0, 0): the end wraps below 0, no level contains the interval
    r.l_qseq = readlen;
    r.l_qname = (uint8_t)(strlen(name) + 1);
    r.qual = unmapped ? 0 : (uint8_t)mapq;
    d.insert(d.end(), (const uint8_t *)name, (const uint8_t *)name + strlen(name) + 1);
    r.n_cigar = 0;
    if (!unmapped) {
        if (cigar) {
            // AssignCigarStrToSAMIU (BGS-IO.cpp:177-201)
            int num = 0;
            for (char c : *cigar) {
                if (c >= '0' && c <= '9') num = num * 10 + c - '0';
                else if (num > 0) { const uint32_t op = c == 'M' ? 0u : c == 'I' ? 1u : c == 'D' ? 2u : 4u; put32(d, ((uint32_t)num << 4) | op); ++r.n_cigar; num = 0; }
            }
        } else { r.n_cigar = 1; put32(d, (uint32_t)readlen << 4); }
    }
    {
        const size_t at = d.size(), half = ((size_t)readlen + 1) / 2;
        d.resize(at + half + (size_t)readlen);
        uint8_t *p = d.data() + at, *pq = p + half;
        if (strand == 2) {
            if (readlen % 2 == 1) {
                for (int i = (readlen - 1) / 2; i > 0; --i) *p++ = (uint8_t)((kNt16[3 - seq[i * 2]] << 4) | kNt16[3 - seq[i * 2 - 1]]);
                *p++ = (uint8_t)(kNt16[3 - seq[0]] << 4);
            } else {
                for (int i = readlen / 2 - 1; i >= 0; --i) *p++ = (uint8_t)((kNt16[3 - seq[i * 2 + 1]] << 4) | kNt16[3 - seq[i * 2]]);
            }
            for (int i = readlen - 1; i >= 0; --i) *pq++ = (uint8_t)qual[i];
        } else {
            for (int i = 0; i < readlen / 2; ++i) *p++ = (uint8_t)((kNt16[seq[i * 2]] << 4) | kNt16[seq[i * 2 + 1]]);
            if (readlen % 2 == 1) *p++ = (uint8_t)(kNt16[seq[readlen - 1]] << 4);
            memcpy(pq, qual, (size_t)readlen);
        }
    }
    const size_t auxStart = d.size();
    put_tag(d, "RG", 'Z', readGroup, strlen(readGroup) + 1);
    if (!unmapped) {
        if (printMDNM && editDist >= 0) put_tag(d, "NM", 'i', &editDist, 4);
        if (x0 >= 0) put_tag(d, "X0", 'i', &x0, 4);
        if (x1 >= 0) put_tag(d, "X1", 'i', &x1, 4);
        if (mismatchNum >= 0) put_tag(d, "XM", 'i', &mismatchNum, 4);
        if (gapOpen >= 0) put_tag(d, "XO", 'i', &gapOpen, 4);
        if (gapExt >= 0) put_tag(d, "XG", 'i', &gapExt, 4);
        if (printMDNM && !md.empty()) put_tag(d, "MD", 'Z', md.c_str(), md.size() + 1);
        if (!xa.empty()) put_tag(d, "XA", 'Z', xa.c_str(), xa.size() + 1);
    }
    r.l_aux = 2 * (int32_t)(d.size() - auxStart);       // counted by bam_aux_append and again by initializeSAMAlgnmt2 (:2277)
}

int finish(s3_sam_record &r, const std::vector<uint8_t> &d)
{
    r.data_len = (int32_t)d.size();
    r.data = (uint8_t *)malloc(d.size() ? d.size() : 1);
    if (!r.data) return S3_ENOMEM;
    memcpy(r.data, d.data(), d.size());
    return S3_OK;
}

}  // namespace

extern "C" void s3_sam_record_free(s3_sam_record *r)
{
    if (!r) return;
    free(r->data);
    memset(r, 0, sizeof *r);
}

extern "C" int s3_sam_pair_records(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_pairing *pairs, uint32_t numPairs, int32_t bestIndex,
                                   const uint8_t *query1, const uint8_t *query2, const char *qualities1, const char *qualities2,
                                   int32_t readlen1, int32_t readlen2, const char *queryName1, const char *queryName2,
                                   int32_t minTotalMismatch, int32_t secMinTotalMismatch, int32_t x0First, int32_t x0Second, int32_t x1First, int32_t x1Second,
                                   int32_t numMinMismatchPair, int32_t isBestHit1, int32_t isBestHit2, uint32_t totalNumValidPairs, s3_sam_record out[2])
{
    if (!out) { s3_set_error("s3_sam_pair_records: NULL output"); return S3_EINVAL; }
    memset(out, 0, 2 * sizeof(s3_sam_record));
    if (!g || !cfg || !query1 || !query2 || !qualities1 || !qualities2 || !queryName1 || !queryName2 || !cfg->readGroup || (numPairs && !pairs) ||
        readlen1 <= 0 || readlen2 <= 0 || bestIndex >= (int32_t)numPairs) { s3_set_error("s3_sam_pair_records: bad argument"); return S3_EINVAL; }
    if (bestIndex >= 0 && (!g->packedDNA || !g->segments || !g->ambiguityMap || !g->chrEndPos || !g->chrNames || g->numSegments == 0)) {
        s3_set_error("s3_sam_pair_records: incomplete genome description"); return S3_EINVAL;
    }
    std::vector<uint8_t> d;
    const std::string none;
    int rc;
    if (bestIndex < 0) {
        // no pair reported: both reads unmapped (BGS-IO.cpp:3748-3787)
        for (int k = 0; k < 2; ++k) {
            s3_sam_record &r = out[k];
            record_body(r, d, k ? readlen2 : readlen1, k ? queryName2 : queryName1, k ? query2 : query1, k ? qualities2 : qualities1, 1, none, NULL, true,
                        0, 0, 0, 0, 0, 0, none, 0, cfg->readGroup, false);
            r.flag = (uint16_t)(1 | (k ? 128 : 64) | 4 | 8);
            r.tid = r.pos = r.mtid = r.mpos = -1; r.isize = 0;
            if ((rc = finish(r, d))) { s3_sam_record_free(&out[0]); s3_sam_record_free(&out[1]); s3_set_error("s3_sam_pair_records: out of host memory"); return rc; }
        }
        return S3_OK;
    }
    const s3_sam_pairing &best = pairs[bestIndex];
    unsigned long long tp[2];
    uint32_t chr[2];
    std::string newCigar[2], md[2];
    int avgQual[2], bestMismatch[2], mapq[2];
    const int trim1 = chr_and_pos_checked(g, (uint32_t)readlen1, best.algnmt1, &tp[0], &chr[0], newCigar[0]);
    const int trim2 = chr_and_pos_checked(g, (uint32_t)readlen2, best.algnmt2, &tp[1], &chr[1], newCigar[1]);
    md_string(g, query1, qualities1, (uint32_t)readlen1, best.algnmt1, best.strand1, (int8_t)best.mismatch1, trim1, md[0], &avgQual[0]);
    md_string(g, query2, qualities2, (uint32_t)readlen2, best.algnmt2, best.strand2, (int8_t)best.mismatch2, trim2, md[1], &avgQual[1]);
    bestMismatch[0] = (int8_t)best.mismatch1; bestMismatch[1] = (int8_t)best.mismatch2;
    const int trims[2] = {trim1, trim2};
    for (int k = 0; k < 2; ++k)
        if (trims[k] && bestMismatch[k]) { bestMismatch[k] = 0; for (char c : md[k]) bestMismatch[k] += c > '9'; }
    const int8_t minTot = (int8_t)minTotalMismatch, secMinTot = (int8_t)secMinTotalMismatch;
    if (cfg->alignmentType == 1 || cfg->alignmentType == 2) {           // OUTPUT_ALL_VALID / OUTPUT_ALL_BEST
        if (cfg->bwaLikeScore) {
            const int op = (readlen1 + readlen2 - minTot) * cfg->dpMatchScore + minTot * cfg->dpMisMatchScore;
            const int subop = (readlen1 + readlen2 - secMinTot) * cfg->dpMatchScore + secMinTot * cfg->dpMisMatchScore;
            // (unsigned arithmetic like the reference: totalNumValidPairs is unsigned there)
            const int subopNum = (totalNumValidPairs - (uint32_t)numMinMismatchPair) > 0 ? (int)(totalNumValidPairs - (uint32_t)numMinMismatchPair) : 0;
            s3_mapq_bwa_pair(x0First, x1First, x0Second, x1Second, op, numMinMismatchPair, subop, subopNum, readlen1, readlen2, &mapq[0], &mapq[1]);
        } else {
            const int s1 = s3_mapq_pair_end((int8_t)best.mismatch1, cfg->isFastq == 1 ? avgQual[0] : 20, x0First, x1First, (char)isBestHit1, totalNumValidPairs, cfg->maxMAPQ, cfg->minMAPQ);
            const int s2 = s3_mapq_pair_end((int8_t)best.mismatch2, cfg->isFastq == 1 ? avgQual[1] : 20, x0Second, x1Second, (char)isBestHit2, totalNumValidPairs, cfg->maxMAPQ, cfg->minMAPQ);
            mapq[0] = mapq[1] = s3_mapq_of_pair(s1, s2);
        }
        if (trim1) mapq[0] = 0;
        if (trim2) mapq[1] = 0;
    } else mapq[0] = mapq[1] = 255;                                       // SAM_MAPQ_UNAVAILABLE
    const int readlen[2] = {readlen1, readlen2};
    for (int k = 0; k < 2; ++k) {
        // XA:Z: the other pairs with no more mismatches than the best, in list order, up to peMaxOutputPerPair results in all
        std::string xa;
        if (cfg->outputXAZTag == 1) {
            uint32_t total = 1;
            char num[24];
            for (uint32_t i = 0; i < numPairs && total < cfg->peMaxOutputPerPair; ++i) {
                if ((int32_t)i == bestIndex || pairs[i].totalMismatchCount > minTot) continue;
                unsigned long long t;
                uint32_t c;
                chr_and_pos(g, k ? pairs[i].algnmt2 : pairs[i].algnmt1, &t, &c);
                xa += g->chrNames[c - 1];
                xa.push_back(',');
                xa.push_back((k ? pairs[i].strand2 : pairs[i].strand1) == 2 ? '-' : '+');
                xa.append(num, write_num((long long)t, num));
                xa.push_back(',');
                xa.append(num, write_num(readlen[k], num));
                xa += "M,";
                xa.append(num, write_num((int)(int8_t)(k ? pairs[i].mismatch2 : pairs[i].mismatch1), num));
                xa.push_back(';');
                ++total;
            }
        }
        s3_sam_record &r = out[k];
        const int strand = k ? best.strand2 : best.strand1, mateStrand = k ? best.strand1 : best.strand2;
        record_body(r, d, readlen[k], k ? queryName2 : queryName1, k ? query2 : query1, k ? qualities2 : qualities1, strand, xa,
                    newCigar[k].empty() ? NULL : &newCigar[k], false, bestMismatch[k], bestMismatch[k], k ? x0Second : x0First, k ? x1Second : x1First, 0, 0,
                    md[k], mapq[k], cfg->readGroup, cfg->isPrintMDNM != 0);
        r.flag = (uint16_t)(1 | 2 | (k ? 128 : 64) | (strand == 2 ? 16 : 0) | (mateStrand == 2 ? 32 : 0));
        r.tid = (int32_t)chr[k] - 1; r.pos = (int32_t)(tp[k] - 1);
        r.mtid = (int32_t)chr[1 - k] - 1; r.mpos = (int32_t)(tp[1 - k] - 1);
        if (chr[0] == chr[1]) {
            const unsigned long long me = tp[k], mate = tp[1 - k];
            r.isize = mate > me ? (int32_t)(mate + (unsigned)readlen[1 - k] - me) : -(int32_t)(me + (unsigned)readlen[k] - mate);
        } else r.isize = 0;
        if ((rc = finish(r, d))) { s3_sam_record_free(&out[0]); s3_sam_record_free(&out[1]); s3_set_error("s3_sam_pair_records: out of host memory"); return rc; }
    }
    return S3_OK;
}

// OCCOutputSAMAPI (BGS-IO.cpp:5556-5772): a single read's record from its occurrence list -- the occurrence with the fewest
// mismatches (the first of them) is reported, X0 = how many share that count; when it hangs over a chromosome / segment end the
// best occurrence that does not is taken instead (none: the first choice, trimmed); the others go to XA:Z (all of them, or with
// all-best only those with the best count; never one that would be trimmed), X1 = how many of the listed ones have more
// mismatches; MAPQ = s3_mapq_single.  No occurrence: the unmapped record.
extern "C" int s3_sam_single_record(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_occurrence *occ, uint32_t numOcc,
                                    const uint8_t *query, const char *qualities, int32_t readlen, const char *queryName, s3_sam_record *out)
{
    if (!out) { s3_set_error("s3_sam_single_record: NULL output"); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    if (!g || !cfg || !query || !qualities || !queryName || !cfg->readGroup || (numOcc && !occ) || readlen <= 0) { s3_set_error("s3_sam_single_record: bad argument"); return S3_EINVAL; }
    if (numOcc && (!g->packedDNA || !g->segments || !g->ambiguityMap || !g->chrEndPos || !g->chrNames || g->numSegments == 0)) {
        s3_set_error("s3_sam_single_record: incomplete genome description"); return S3_EINVAL;
    }
    std::vector<uint8_t> d;
    std::string xa, md, newCigar, scratch;
    unsigned long long bestTP = 0, tp;
    uint32_t bestChr = 0, chr;
    uint32_t bestMism = 0;
    int best = -1, bestHitNum = 0, secBestHitNum = 0, trim = 0, avgQual = 20, mapq = 0;
    int strand = 1;
    if (numOcc) {
        best = 0; bestMism = occ[0].mismatchCount; bestHitNum = 1;
        for (uint32_t k = 1; k < numOcc; ++k) {
            if (occ[k].mismatchCount < bestMism) { best = (int)k; bestMism = occ[k].mismatchCount; bestHitNum = 1; }
            else if (occ[k].mismatchCount == bestMism) ++bestHitNum;
        }
        strand = occ[best].strand;
        trim = chr_and_pos_checked(g, (uint32_t)readlen, occ[best].ambPosition, &bestTP, &bestChr, scratch);
        if (trim) {
            const int stored = best;
            best = -1; bestMism = 9999; bestHitNum = 0;
            for (uint32_t k = 0; k < numOcc; ++k) {
                if (chr_and_pos_checked(g, (uint32_t)readlen, occ[k].ambPosition, &tp, &chr, scratch)) continue;
                if (occ[k].mismatchCount < bestMism) { best = (int)k; bestMism = occ[k].mismatchCount; bestHitNum = 1; }
                else if (occ[k].mismatchCount == bestMism) ++bestHitNum;
            }
            if (best < 0) { best = stored; bestHitNum = 1; }
            newCigar.clear();
            trim = chr_and_pos_checked(g, (uint32_t)readlen, occ[best].ambPosition, &bestTP, &bestChr, newCigar);
            // (the strand stays that of the first choice, and bestMism what the scan left: 9999 when nothing fits -- as the reference)
        }
    }
    if (numOcc > 1 && !trim) {
        char num[24];
        for (uint32_t k = 0; k < numOcc; ++k) {
            if ((int)k == best) continue;
            if (cfg->alignmentType == 2 && occ[k].mismatchCount > bestMism) continue;
            if (chr_and_pos_checked(g, (uint32_t)readlen, occ[k].ambPosition, &tp, &chr, scratch)) continue;
            xa += g->chrNames[chr - 1];
            xa.push_back(',');
            xa.push_back(occ[k].strand == 2 ? '-' : '+');
            xa.append(num, write_num((long long)tp, num));
            xa.push_back(',');
            xa.append(num, write_num(readlen, num));
            xa += "M,";
            xa.append(num, write_num((int)occ[k].mismatchCount, num));
            xa.push_back(';');
            if (occ[k].mismatchCount > bestMism) ++secBestHitNum;
        }
    }
    if (best >= 0) {
        md_string(g, query, qualities, (uint32_t)readlen, occ[best].ambPosition, strand, (int)bestMism, trim, md, &avgQual);
        if (trim && bestMism) { bestMism = 0; for (char c : md) bestMism += c > '9'; }
        mapq = s3_mapq_single((int)bestMism, cfg->isFastq == 1 ? avgQual : 20, bestHitNum, secBestHitNum, cfg->maxMAPQ, cfg->minMAPQ, cfg->bwaLikeScore);
    }
    const std::string none;
    record_body(*out, d, readlen, queryName, query, qualities, strand, numOcc > 1 ? xa : none, newCigar.empty() ? NULL : &newCigar, numOcc == 0,
                (int)bestMism, (int)bestMism, bestHitNum, secBestHitNum, 0, 0, md, mapq, cfg->readGroup, cfg->isPrintMDNM != 0);
    if (best >= 0) {
        out->flag = (uint16_t)(strand == 2 ? 16 : 0);
        out->tid = (int32_t)bestChr - 1; out->pos = (int32_t)(bestTP - 1);
    } else { out->flag = 4; out->tid = -1; out->pos = -1; }
    out->mtid = -1; out->mpos = -1; out->isize = 0;
    const int rc = finish(*out, d);
    if (rc) { s3_set_error("s3_sam_single_record: out of host memory"); }
    return rc;
}


// ---- SingleDPOutputSAMAPI (BGS-IO.cpp:5857-6118): the record of a single read from its DP alignments ---------------------------
void s3_special_to_sam(const char *sp, size_t len, std::string &out);          // convertToCigarStr (s3_decode.cu)

namespace {

// BoundaryCheckDP + getChrAndPosWithBoundaryCheckDP (BGS-IO.cpp:1807-1976, 2009-2033): an alignment with gaps that hangs over a
// chromosome / segment end is cut there; the longer side stays, the other becomes a soft clip.  Returns the bases trimmed (> 0 on
// the left, < 0 on the right, 0: nothing to trim); newCigar (may be NULL) = the special CIGAR of what stays.
int chr_and_pos_checked_dp(const s3_sam_genome *g, uint32_t readLength, uint32_t ambPos, const char *cigar, unsigned long long *tp, uint32_t *chr,
                           std::string *newCigar)
{
    uint32_t segEnd = chr_and_pos(g, ambPos, tp, chr);
    const uint32_t chrEnd = g->chrEndPos[*chr - 1];
    if (ambPos + readLength * 2u <= chrEnd + 1u && ambPos + readLength * 2u <= segEnd + 1u) return 0;
    segEnd = chrEnd < segEnd ? chrEnd : segEnd;
    std::string left, right;
    char buf[48];
    int leftLen = 0, rightLen = 0, rightOff = 0, leftS = 0, rightS = 0;
    uint32_t refPos = ambPos;
    for (const char *c = cigar; *c;) {
        int num = 0;
        while (*c && *c <= '9') { num = num * 10 + (*c - '0'); ++c; }
        if (!*c) break;
        const char op = *c++;
        if (op == 'S') {
            snprintf(buf, sizeof buf, "%dS", num);
            if (refPos <= segEnd) { left += buf; leftS += num; } else { right += buf; rightS += num; }
        } else if (op == 'M' || op == 'm') {
            if (refPos > segEnd) { rightLen += num; snprintf(buf, sizeof buf, "%d%c", num, op); right += buf; }
            else if (refPos + (uint32_t)num <= segEnd + 1u) { leftLen += num; snprintf(buf, sizeof buf, "%d%c", num, op); left += buf; }
            else {
                const int l = (int)(segEnd - refPos + 1u);
                leftLen += l; snprintf(buf, sizeof buf, "%d%c", l, op); left += buf;
                rightLen += num - l; snprintf(buf, sizeof buf, "%d%c", num - l, op); right += buf;
            }
            refPos += (uint32_t)num;
        } else if (op == 'D') {
            snprintf(buf, sizeof buf, "%dD", num);
            if (refPos > segEnd) { if (right.empty()) rightOff = num; else right += buf; }
            else if (refPos + (uint32_t)num <= segEnd + 1u) left += buf;
            else {
                const int l = (int)(segEnd - refPos + 1u);
                snprintf(buf, sizeof buf, "%dD", l); left += buf;
                snprintf(buf, sizeof buf, "%dD", num - l);
                if (right.empty()) rightOff = num - l; else right += buf;
            }
            refPos += (uint32_t)num;
        } else if (op == 'I') {
            if (refPos == chrEnd + 1u) { /* exactly at the boundary: on neither side */ }
            else if (refPos <= segEnd) { leftLen += num; snprintf(buf, sizeof buf, "%dI", num); left += buf; }
            else { rightLen += num; snprintf(buf, sizeof buf, "%dI", num); right += buf; }
        }
    }
    if (!leftLen || !rightLen) return 0;
    int ret;
    uint32_t corrected;
    if (leftLen >= rightLen) {
        const int clip = (int)readLength - leftLen - leftS;
        if (newCigar) { snprintf(buf, sizeof buf, "%dS", clip); *newCigar = left + buf; }
        corrected = ambPos; ret = -clip;
    } else {
        const int clip = (int)readLength - rightLen - rightS;
        if (newCigar) { snprintf(buf, sizeof buf, "%dS", clip); *newCigar = buf + right; }
        corrected = segEnd + 1u + (uint32_t)rightOff; ret = clip;
    }
    if (ret && corrected > ambPos) chr_and_pos(g, corrected, tp, chr);
    return ret;
}

}  // namespace

extern "C" int s3_sam_single_dp_record(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_dp_alignment *algn, uint32_t numResult,
                                       int32_t singleDPcutoffThreshold, const uint8_t *query, const char *qualities, int32_t readlen, const char *queryName,
                                       s3_sam_record *out)
{
    if (!out) { s3_set_error("s3_sam_single_dp_record: NULL output"); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    if (!g || !cfg || !query || !qualities || !queryName || !cfg->readGroup || (numResult && !algn) || readlen <= 0) { s3_set_error("s3_sam_single_dp_record: bad argument"); return S3_EINVAL; }
    const bool unaligned = numResult == 0 || algn[0].ambPosition == 0xFFFFFFFFu;
    if (!unaligned && (!g->packedDNA || !g->segments || !g->ambiguityMap || !g->chrEndPos || !g->chrNames || g->numSegments == 0)) {
        s3_set_error("s3_sam_single_dp_record: incomplete genome description"); return S3_EINVAL;
    }
    for (uint32_t k = 0; k < numResult && !unaligned; ++k) if (!algn[k].cigar) { s3_set_error("s3_sam_single_dp_record: alignment %u without a CIGAR", k); return S3_EINVAL; }
    std::vector<uint8_t> d;
    std::string xa, md, cigar, newSp;
    unsigned long long bestTP = 0, tp;
    uint32_t bestChr = 0, chr;
    int strand = 1, bestHitNum = 1, secBestHitNum = 0, mism = 0, gapOpen = 0, gapExt = 0, editDist = 0, avgQual = 20, mapq = 0;
    const bool lists = numResult > 1 && (cfg->alignmentType == 1 || cfg->alignmentType == 2);      // OUTPUT_ALL_VALID / OUTPUT_ALL_BEST
    if (!unaligned) {
        // the reference keeps the scores in unsigned ints (BGS-IO.cpp:5877): comparisons are made that way here too
        uint32_t best = 0, bestScore = (uint32_t)algn[0].score, secBestScore = 0;
        for (uint32_t k = 1; k < numResult; ++k) {
            const uint32_t cur = (uint32_t)algn[k].score;
            if (cur >= bestScore) {
                if (cur == bestScore) ++bestHitNum;
                else { secBestScore = bestScore; bestScore = cur; bestHitNum = 1; best = k; }
            } else if (cur >= secBestScore) { if (cur != secBestScore) secBestScore = cur; }
        }
        int trim = chr_and_pos_checked_dp(g, (uint32_t)readlen, algn[best].ambPosition, algn[best].cigar, &bestTP, &bestChr, NULL);
        if (trim) {
            // the best alignment that does not hang over an end, searched from start values of -9998 / -9999 held in unsigned ints:
            // only a score that is negative as an int can replace them (as the reference)
            bestScore = (uint32_t)-9998; secBestScore = (uint32_t)-9999;
            for (uint32_t k = 0; k < numResult; ++k) {
                if (chr_and_pos_checked_dp(g, (uint32_t)readlen, algn[k].ambPosition, algn[k].cigar, &bestTP, &bestChr, NULL)) continue;
                const uint32_t cur = (uint32_t)algn[k].score;
                if (cur >= bestScore) {
                    if (cur == bestScore) ++bestHitNum;
                    else { secBestScore = bestScore; bestScore = cur; bestHitNum = 1; best = k; }
                } else if (cur > secBestScore) secBestScore = cur;
            }
            trim = chr_and_pos_checked_dp(g, (uint32_t)readlen, algn[best].ambPosition, algn[best].cigar, &bestTP, &bestChr, NULL);
        }
        int x1t1 = 0, x1t2 = 0;
        if (lists && !trim) {
            const int thres = (int)(0.7 * bestScore);
            char num[24];
            for (uint32_t k = 0; k < numResult; ++k) {
                if (k == best) continue;
                const uint32_t cur = (uint32_t)algn[k].score;
                if (cfg->alignmentType == 2 && cur < bestScore) continue;
                if (chr_and_pos_checked_dp(g, (uint32_t)readlen, algn[k].ambPosition, algn[k].cigar, &tp, &chr, NULL)) continue;
                xa += g->chrNames[chr - 1];
                xa.push_back(',');
                xa.push_back(algn[k].strand == 2 ? '-' : '+');
                xa.append(num, write_num((long long)tp, num));
                xa.push_back(',');
                s3_special_to_sam(algn[k].cigar, strlen(algn[k].cigar), xa);
                xa.push_back(',');
                xa.append(num, write_num(algn[k].editdist, num));
                xa.push_back(';');
                if (cur < bestScore) { if (cur >= (uint32_t)thres) ++x1t1; else ++x1t2; }
            }
        }
        secBestHitNum = x1t1 + x1t2;
        strand = algn[best].strand;
        trim = chr_and_pos_checked_dp(g, (uint32_t)readlen, algn[best].ambPosition, algn[best].cigar, &bestTP, &bestChr, &newSp);
        const std::string sp = trim ? newSp : std::string(algn[best].cigar);
        s3_special_to_sam(sp.c_str(), sp.size(), cigar);
        // getMisInfoForDP (PE.cpp:499-666): a left trim moves the text position, the qualities stay indexed by read offset
        {
            const uint64_t cigOff[2] = {0, sp.size()}, qOff[2] = {0, (uint64_t)readlen};
            const uint32_t pos = algn[best].ambPosition + (trim > 0 ? (uint32_t)trim : 0u);
            uint64_t mdOff[2];
            char *mdBuf = NULL;
            int32_t nm = 0, go = 0, ge = 0, aq = 20;
            const int rc = s3_dp_md(g->packedDNA, g->dnaLength, sp.c_str(), cigOff, &pos, 1, (const int8_t *)qualities, qOff, mdOff, &mdBuf, &nm, &go, &ge, &aq);
            if (rc) return rc;
            md.assign(mdBuf, (size_t)(mdOff[1] - mdOff[0]));
            free(mdBuf);
            mism = nm; gapOpen = go; gapExt = ge; avgQual = aq;
        }
        editDist = gapExt + mism;
        mapq = s3_mapq_single_dp(readlen * cfg->dpMatchScore, cfg->isFastq == 1 ? avgQual : 20, bestHitNum, x1t1, x1t2, (int)bestScore, (int)secBestScore,
                                 cfg->maxMAPQ, cfg->minMAPQ, singleDPcutoffThreshold, cfg->bwaLikeScore);
    }
    const std::string none;
    record_body(*out, d, readlen, queryName, query, qualities, strand, lists ? xa : none, &cigar, unaligned, mism, editDist, bestHitNum, secBestHitNum,
                gapOpen, gapExt, md, mapq, cfg->readGroup, cfg->isPrintMDNM != 0);
    if (!unaligned) {
        out->flag = (uint16_t)(strand == 2 ? 16 : 0);
        out->tid = (int32_t)bestChr - 1; out->pos = (int32_t)(bestTP - 1);
    } else { out->flag = 4; out->tid = -1; out->pos = -1; }
    out->mtid = -1; out->mpos = -1; out->isize = 0;
    const int rc = finish(*out, d);
    if (rc) { s3_set_error("s3_sam_single_dp_record: out of host memory"); }
    return rc;
}


// ---- pairDeepDPOutputSAMAPI (BGS-IO.cpp:3824-4500): the two records of a read pair from its deep-DP alignments ------------------
namespace {

// readLengthWithCigar (BGS-IO.cpp:3795-3822): text bases an alignment covers (M, m, D; a deletion as the last op is ignored)
int cigar_span(const char *cigar)
{
    int len = 0, x = 0;
    char op = ' ';
    for (const char *p = cigar; *p;) {
        x = 0;
        while (*p && *p <= '9') { x = x * 10 + (*p - '0'); ++p; }
        if (!*p) break;
        op = *p++;
        if (op == 'M' || op == 'm' || op == 'D') len += x;
    }
    if (op == 'D') len -= x;
    return len;
}

struct DpSide {                      // what getMisInfoForDP and convertToCigarStr give for one read's reported alignment
    std::string cigar, md;
    int mism = 0, gapOpen = 0, gapExt = 0, avgQual = 20, trim = 0;
};
int dp_side(const s3_sam_genome *g, const char *special, uint32_t pos, int readlen, const char *qualities, unsigned long long *tp, uint32_t *chr, DpSide &o)
{
    std::string newSp;
    o.trim = chr_and_pos_checked_dp(g, (uint32_t)readlen, pos, special, tp, chr, &newSp);
    const std::string sp = o.trim ? newSp : std::string(special);
    s3_special_to_sam(sp.c_str(), sp.size(), o.cigar);
    const uint64_t cigOff[2] = {0, sp.size()}, qOff[2] = {0, (uint64_t)readlen};
    const uint32_t p = pos + (o.trim > 0 ? (uint32_t)o.trim : 0u);
    uint64_t mdOff[2];
    char *mdBuf = NULL;
    int32_t nm = 0, go = 0, ge = 0, aq = 20;
    const int rc = s3_dp_md(g->packedDNA, g->dnaLength, sp.c_str(), cigOff, &p, 1, (const int8_t *)qualities, qOff, mdOff, &mdBuf, &nm, &go, &ge, &aq);
    if (rc) return rc;
    o.md.assign(mdBuf, (size_t)(mdOff[1] - mdOff[0]));
    free(mdBuf);
    o.mism = nm; o.gapOpen = go; o.gapExt = ge; o.avgQual = aq;
    return S3_OK;
}

}  // namespace

extern "C" int s3_sam_deep_dp_records(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_deep_alignment *algn, uint32_t num, int32_t bestIndex,
                                      const uint8_t *query1, const uint8_t *query2, const char *qualities1, const char *qualities2,
                                      int32_t readlen1, int32_t readlen2, const char *queryName1, const char *queryName2,
                                      const int32_t x0[2], const int32_t x1[2], const int32_t mismatch[2], s3_sam_record out[2])
{
    if (!out) { s3_set_error("s3_sam_deep_dp_records: NULL output"); return S3_EINVAL; }
    memset(out, 0, 2 * sizeof(s3_sam_record));
    if (!g || !cfg || !query1 || !query2 || !qualities1 || !qualities2 || !queryName1 || !queryName2 || !cfg->readGroup || (num && !algn) || !x0 || !x1 || !mismatch ||
        readlen1 <= 0 || readlen2 <= 0 || bestIndex >= (int32_t)num) { s3_set_error("s3_sam_deep_dp_records: bad argument"); return S3_EINVAL; }
    const uint32_t NONE = 0xFFFFFFFFu;
    const uint8_t *query[2] = {query1, query2};
    const char *qual[2] = {qualities1, qualities2}, *name[2] = {queryName1, queryName2};
    const int readlen[2] = {readlen1, readlen2};
    std::vector<uint8_t> d;
    const std::string none;
    int rc;
    if (bestIndex < 0) {
        // nothing to report: both reads unmapped, flags 0x41 / 0x81 (BGS-IO.cpp:4463-4494)
        for (int k = 0; k < 2; ++k) {
            s3_sam_record &r = out[k];
            record_body(r, d, readlen[k], name[k], query[k], qual[k], 1, none, NULL, true, 0, 0, 0, 0, 0, 0, none, 0, cfg->readGroup, false);
            r.flag = (uint16_t)(1 | (k ? 128 : 64));
            r.tid = r.pos = r.mtid = r.mpos = -1; r.isize = 0;
            if ((rc = finish(r, d))) { s3_sam_record_free(&out[0]); s3_sam_record_free(&out[1]); s3_set_error("s3_sam_deep_dp_records: out of host memory"); return rc; }
        }
        return S3_OK;
    }
    if (!g->packedDNA || !g->segments || !g->ambiguityMap || !g->chrEndPos || !g->chrNames || g->numSegments == 0) { s3_set_error("s3_sam_deep_dp_records: incomplete genome description"); return S3_EINVAL; }
    const s3_sam_deep_alignment &best = algn[bestIndex];
    uint32_t bpos[2] = {best.ambPosition[0], best.ambPosition[1]};
    if (bpos[0] == NONE && bpos[1] == NONE) {
        s3_set_error("s3_sam_deep_dp_records: a pair without any alignment goes through the writer of improperly paired reads (unproperlypairOutputSAMAPI2), which is not built");
        return S3_EINVAL;
    }
    for (uint32_t i = 0; i < num; ++i)
        for (int k = 0; k < 2; ++k)
            if (algn[i].ambPosition[k] != NONE && !algn[i].cigar[k]) { s3_set_error("s3_sam_deep_dp_records: alignment %u without a CIGAR", i); return S3_EINVAL; }
    unsigned long long tp[2] = {0, 0};
    uint32_t chr[2] = {0, 0};
    DpSide sd[2];
    int strand[2] = {1, 1}, span[2] = {readlen1, readlen2};
    for (int k = 0; k < 2; ++k) {
        if (bpos[k] == NONE) continue;
        strand[k] = best.strand[k];
        if ((rc = dp_side(g, best.cigar[k], bpos[k], readlen[k], qual[k], &tp[k], &chr[k], sd[k]))) return rc;
        span[k] = cigar_span(best.cigar[k]);
    }
    int bestInsert = (bpos[0] != NONE && bpos[1] != NONE) ? best.insertSize : 0;
    // a pair whose reads run through each other loses the read with more mismatches (:3949-3972)
    if (bpos[0] != NONE && bpos[1] != NONE) {
        const unsigned long long a1 = (unsigned long long)bpos[0] + (sd[0].trim > 0 ? sd[0].trim : 0), a2 = (unsigned long long)bpos[1] + (sd[1].trim > 0 ? sd[1].trim : 0);
        if ((best.strand[0] == 1 && (a1 > a2 || a1 + span[0] > a2 + span[1])) || (best.strand[0] == 2 && (a2 > a1 || a2 + span[1] > a1 + span[0]))) {
            const int drop = sd[0].mism <= sd[1].mism ? 1 : 0;
            bpos[drop] = NONE; tp[drop] = 0; chr[drop] = 0;
            bestInsert = 0;
        }
    }
    const bool both = bpos[0] != NONE && bpos[1] != NONE;
    int bestPairNum = 0, bestPairScore = 0, secBestPairScore = 0, numSimilar = 0;
    if (both) {
        bestPairNum = 1; numSimilar = 1;
        bestPairScore = best.score[0] + best.score[1];
        for (uint32_t i = 0; i < num && num > 1; ++i) {
            if ((int32_t)i == bestIndex) continue;
            const int sum = algn[i].score[0] + algn[i].score[1];
            if (sum == bestPairScore) ++bestPairNum;
            else if (sum > secBestPairScore) secBestPairScore = sum;
            if (algn[i].score[0] >= best.score[0] + cfg->dpMisMatchScore && algn[i].score[1] >= best.score[1] + cfg->dpMisMatchScore) ++numSimilar;
        }
    }
    int bestHitNum[2] = {0, 0}, secBestHitNum[2] = {0, 0}, bestScore[2] = {0, 0}, secBestScore[2] = {0, 0}, isBestHit[2] = {1, 1};
    for (int k = 0; k < 2; ++k) {
        uint32_t bestPos = NONE, secBestPos = NONE;
        if (bpos[k] != NONE) { bestScore[k] = best.score[k]; bestPos = bpos[k]; bestHitNum[k] = 1; }
        if (bpos[k] != NONE && num > 1) {
            for (uint32_t i = 0; i < num; ++i) {
                if ((int32_t)i == bestIndex) continue;
                const int sc = algn[i].score[k], same = algn[i].numSameScore[k];
                const uint32_t pos = algn[i].ambPosition[k];
                if (sc >= bestScore[k]) {
                    if (sc == bestScore[k]) { if (pos != bestPos) bestHitNum[k] += same; }
                    else { secBestScore[k] = bestScore[k]; secBestHitNum[k] = bestHitNum[k]; secBestPos = bestPos; bestScore[k] = sc; bestHitNum[k] = same; bestPos = pos; isBestHit[k] = 0; }
                } else if (sc >= secBestScore[k]) {
                    if (sc == secBestScore[k]) { if (pos != secBestPos) secBestHitNum[k] += same; }
                    else { secBestScore[k] = sc; secBestPos = pos; secBestHitNum[k] = same; }
                }
            }
        }
        // the counts the search left for the read (hspaux->x0_array / x1_array / mismatch_array; the first read's are used from 1 on, the mate's from 2 on)
        if (x0[k] > (k ? 1 : 0)) {
            const int x0Score = mismatch[k] * cfg->dpMisMatchScore + (readlen[k] - mismatch[k]) * cfg->dpMatchScore;
            if (x0Score >= bestScore[k]) {
                if (x0[k] > bestHitNum[k]) bestHitNum[k] = x0[k];
                if (x1[k] > secBestHitNum[k]) secBestHitNum[k] = x1[k];
                if (x0Score > bestScore[k]) isBestHit[k] = 0;
            }
        }
    }
    int mapq[2];
    const bool lists = cfg->alignmentType == 1 || cfg->alignmentType == 2;
    if (lists) {
        if (cfg->bwaLikeScore) {
            s3_mapq_bwa_pair(bestHitNum[0], secBestHitNum[0], bestHitNum[1], secBestHitNum[1], bestPairScore, bestPairNum, secBestPairScore, (int)num - bestPairNum,
                             readlen1, readlen2, &mapq[0], &mapq[1]);
        } else {
            int m[2];
            for (int k = 0; k < 2; ++k)
                m[k] = s3_mapq_pair_end_dp(best.score[k], readlen[k] * cfg->dpMatchScore, cfg->isFastq == 1 ? sd[k].avgQual : 20, bestHitNum[k], secBestHitNum[k],
                                           bestScore[k], secBestScore[k], isBestHit[k], numSimilar, cfg->maxMAPQ, cfg->minMAPQ);
            mapq[0] = mapq[1] = s3_mapq_of_pair(m[0], m[1]);
        }
        for (int k = 0; k < 2; ++k) if (sd[k].trim) mapq[k] = 0;
    } else mapq[0] = mapq[1] = 255;
    for (int k = 0; k < 2; ++k) {
        std::string xa;
        if (bpos[k] != NONE && num > 1) {
            char numBuf[24];
            for (uint32_t i = 0; i < num; ++i) {
                if ((int32_t)i == bestIndex) continue;
                if (cfg->alignmentType == 2 && algn[i].score[0] + algn[i].score[1] < bestPairScore) continue;
                if (algn[i].ambPosition[k] == NONE) continue;             // no position to translate (the stage's hits always have both)
                unsigned long long t;
                uint32_t c;
                chr_and_pos(g, algn[i].ambPosition[k], &t, &c);
                xa += g->chrNames[c - 1];
                xa.push_back(',');
                xa.push_back(algn[i].strand[k] == 2 ? '-' : '+');
                xa.append(numBuf, write_num((long long)t, numBuf));
                xa.push_back(',');
                s3_special_to_sam(algn[i].cigar[k], strlen(algn[i].cigar[k]), xa);
                xa.push_back(',');
                xa.append(numBuf, write_num(algn[i].editdist[k], numBuf));
                xa.push_back(';');
            }
        }
        if (cfg->alignmentType == 4) { bestHitNum[k] = -1; secBestHitNum[k] = -1; }
        else if (!lists) secBestHitNum[k] = -1;
        s3_sam_record &r = out[k];
        if (bpos[k] != NONE) {
            if (bpos[1 - k] == NONE) {
                mapq[k] = s3_mapq_unique_dp(bestHitNum[k], best.score[k], readlen[k] * cfg->dpMatchScore, cfg->isFastq == 1 ? sd[k].avgQual : 20, cfg->maxMAPQ, cfg->minMAPQ);
                if (sd[k].trim) mapq[k] = 0;
            }
            record_body(r, d, readlen[k], name[k], query[k], qual[k], strand[k], xa, &sd[k].cigar, false, sd[k].mism, sd[k].mism + sd[k].gapExt, bestHitNum[k], secBestHitNum[k],
                        sd[k].gapOpen, sd[k].gapExt, sd[k].md, mapq[k], cfg->readGroup, cfg->isPrintMDNM != 0);
        } else {
            record_body(r, d, readlen[k], name[k], query[k], qual[k], strand[k], none, NULL, true, 0, 0, 0, 0, 0, 0, none, 0, cfg->readGroup, false);
        }
        r.flag = (uint16_t)(1 | (both ? 2 : 0) | (k ? 128 : 64) | (bpos[k] == NONE ? 4 : 0) | (bpos[1 - k] == NONE ? 8 : 0) |
                            (bpos[k] != NONE && best.strand[k] == 2 ? 16 : 0) | (bpos[1 - k] != NONE && best.strand[1 - k] == 2 ? 32 : 0));
        const int m = 1 - k;
        r.tid = chr[k] == 0 ? (chr[m] == 0 ? -1 : (int32_t)chr[m] - 1) : (int32_t)chr[k] - 1;
        r.pos = tp[k] == 0 ? (tp[m] == 0 ? -1 : (int32_t)(tp[m] - 1)) : (int32_t)(tp[k] - 1);
        r.mtid = chr[m] == 0 ? (chr[k] == 0 ? -1 : (int32_t)chr[k] - 1) : (int32_t)chr[m] - 1;
        r.mpos = tp[m] == 0 ? (tp[k] == 0 ? -1 : (int32_t)(tp[k] - 1)) : (int32_t)(tp[m] - 1);
        if (bestInsert > 0) r.isize = tp[k] > tp[m] ? -(int32_t)(tp[k] + (unsigned long long)span[k] - tp[m]) : (int32_t)(tp[m] + (unsigned long long)span[m] - tp[k]);
        else r.isize = 0;
        if ((rc = finish(r, d))) { s3_sam_record_free(&out[0]); s3_sam_record_free(&out[1]); s3_set_error("s3_sam_deep_dp_records: out of host memory"); return rc; }
    }
    return S3_OK;
}


// ---- pairDPOutputSAMAPI (BGS-IO.cpp:4504-5554): the two records of a read pair from its default-DP (mate rescue) results ---------
// One read of every entry comes from the search (score = its mismatches, ungapped), the other from DP (score = DP score, special
// CIGAR): whichFromDP says which; only entries of the reported entry's kind count.
extern "C" int s3_sam_pair_dp_records(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_dp_pairing *algn, uint32_t num, int32_t bestIndex,
                                      const uint8_t *query1, const uint8_t *query2, const char *qualities1, const char *qualities2,
                                      int32_t readlen1, int32_t readlen2, const char *queryName1, const char *queryName2,
                                      const int32_t x0[2], const int32_t x1[2], const int32_t mismatch[2], s3_sam_record out[2])
{
    if (!out) { s3_set_error("s3_sam_pair_dp_records: NULL output"); return S3_EINVAL; }
    memset(out, 0, 2 * sizeof(s3_sam_record));
    if (!g || !cfg || !query1 || !query2 || !qualities1 || !qualities2 || !queryName1 || !queryName2 || !cfg->readGroup || (num && !algn) || !x0 || !x1 || !mismatch ||
        readlen1 <= 0 || readlen2 <= 0 || bestIndex >= (int32_t)num) { s3_set_error("s3_sam_pair_dp_records: bad argument"); return S3_EINVAL; }
    const uint32_t NONE = 0xFFFFFFFFu;
    const uint8_t *query[2] = {query1, query2};
    const char *qual[2] = {qualities1, qualities2}, *name[2] = {queryName1, queryName2};
    const int readlen[2] = {readlen1, readlen2};
    std::vector<uint8_t> d;
    const std::string none;
    int rc;
    if (bestIndex < 0) {
        for (int k = 0; k < 2; ++k) {
            s3_sam_record &r = out[k];
            record_body(r, d, readlen[k], name[k], query[k], qual[k], 1, none, NULL, true, 0, 0, 0, 0, 0, 0, none, 0, cfg->readGroup, false);
            r.flag = (uint16_t)(1 | (k ? 128 : 64));
            r.tid = r.pos = r.mtid = r.mpos = -1; r.isize = 0;
            if ((rc = finish(r, d))) { s3_sam_record_free(&out[0]); s3_sam_record_free(&out[1]); s3_set_error("s3_sam_pair_dp_records: out of host memory"); return rc; }
        }
        return S3_OK;
    }
    if (!g->packedDNA || !g->segments || !g->ambiguityMap || !g->chrEndPos || !g->chrNames || g->numSegments == 0) { s3_set_error("s3_sam_pair_dp_records: incomplete genome description"); return S3_EINVAL; }
    const s3_sam_dp_pairing &best = algn[bestIndex];
    const int fromDP = best.whichFromDP;                                  // 0: the first read's alignment is the DP one, 1: the mate's, 2: neither
    for (uint32_t i = 0; i < num; ++i)
        if (algn[i].whichFromDP <= 1 && algn[i].ambPosition[algn[i].whichFromDP] != NONE && !algn[i].cigar) { s3_set_error("s3_sam_pair_dp_records: entry %u without a CIGAR", i); return S3_EINVAL; }
    uint32_t bpos[2] = {best.ambPosition[0], best.ambPosition[1]};
    unsigned long long tp[2] = {0, 0};
    uint32_t chr[2] = {0, 0};
    DpSide sd[2];                                                         // (cigar stays empty for the read that came from the search)
    std::string newCigar[2];
    int strand[2] = {1, 1}, span[2] = {readlen1, readlen2};
    for (int k = 0; k < 2; ++k) {
        if (bpos[k] == NONE) continue;
        strand[k] = best.strand[k];
        if (fromDP == k) {
            if ((rc = dp_side(g, best.cigar, bpos[k], readlen[k], qual[k], &tp[k], &chr[k], sd[k]))) return rc;
            span[k] = cigar_span(best.cigar);
        } else {
            sd[k].trim = chr_and_pos_checked(g, (uint32_t)readlen[k], bpos[k], &tp[k], &chr[k], newCigar[k]);
            md_string(g, query[k], qual[k], (uint32_t)readlen[k], bpos[k], strand[k], best.score[k], sd[k].trim, sd[k].md, &sd[k].avgQual);
            sd[k].mism = best.score[k];
            if (sd[k].trim && sd[k].mism) { sd[k].mism = 0; for (char c : sd[k].md) sd[k].mism += c > '9'; }
        }
    }
    int bestInsert = (bpos[0] != NONE && bpos[1] != NONE) ? best.insertSize : 0;
    if (bpos[0] != NONE && bpos[1] != NONE) {
        const unsigned long long a1 = (unsigned long long)bpos[0] + (sd[0].trim > 0 ? sd[0].trim : 0), a2 = (unsigned long long)bpos[1] + (sd[1].trim > 0 ? sd[1].trim : 0);
        if ((best.strand[0] == 1 && (a1 > a2 || a1 + span[0] > a2 + span[1])) || (best.strand[0] == 2 && (a2 > a1 || a2 + span[1] > a1 + span[0]))) {
            const int drop = sd[0].mism <= sd[1].mism ? 1 : 0;
            bpos[drop] = NONE; tp[drop] = 0; chr[drop] = 0;
            bestInsert = 0;
        }
    }
    const bool both = bpos[0] != NONE && bpos[1] != NONE;
    // pairs with the reported pair's scores, and the "second best" pair as the reference's scan leaves it (:4700-4777)
    int bestPairNum = 0, bestPairScore[2] = {0, 0}, secBestPairNum = 0, secBestPairScore[2] = {0, 0}, numSimilar = 0;
    if (both) {
        bestPairNum = 1; numSimilar = 1;
        bestPairScore[0] = best.score[0]; bestPairScore[1] = best.score[1];
        for (uint32_t i = 0; i < num && num > 1; ++i) {
            if ((int32_t)i == bestIndex || algn[i].whichFromDP != fromDP) continue;
            const int s1 = algn[i].score[0], s2 = algn[i].score[1];
            if (s1 == bestPairScore[0] && s2 == bestPairScore[1]) ++bestPairNum;
            else if (secBestPairNum == 0) { secBestPairScore[0] = s1; secBestPairScore[1] = s2; secBestPairNum = 1; }
            else if (fromDP == 0 || fromDP == 1) {
                const int dp = fromDP, se = 1 - fromDP;                   // the DP read's score: larger is better; the other's mismatches: fewer is better
                const int sdp = algn[i].score[dp], sse = algn[i].score[se];
                if (sse < secBestPairScore[se]) { secBestPairScore[0] = s1; secBestPairScore[1] = s2; secBestPairNum = 1; }
                else if (sse == secBestPairScore[se] && sdp > secBestPairScore[dp]) { secBestPairScore[dp] = sdp; secBestPairNum = 1; }
                else if (sse == secBestPairScore[se] && sdp == secBestPairScore[dp]) ++secBestPairNum;
            }
            // pairs about as good as the reported one (within one mismatch's worth)
        }
        for (uint32_t i = 0; i < num && num > 1; ++i) {
            if ((int32_t)i == bestIndex || algn[i].whichFromDP != fromDP) continue;
            const int s1 = algn[i].score[0], s2 = algn[i].score[1];
            bool similar;
            if (fromDP == 0) similar = s1 >= bestPairScore[0] + cfg->dpMisMatchScore && s2 <= bestPairScore[1] + 1;
            else if (fromDP == 1) similar = s1 <= bestPairScore[0] + 1 && s2 >= bestPairScore[1] + cfg->dpMisMatchScore;
            else similar = s1 <= bestPairScore[0] + 1 && s2 <= bestPairScore[1] + 1;
            if (similar) ++numSimilar;
        }
    }
    int bestHitNum[2] = {0, 0}, secBestHitNum[2] = {0, 0}, bestScore[2] = {0, 0}, secBestScore[2] = {0, 0}, isBestHit[2] = {1, 1};
    for (int k = 0; k < 2; ++k) {
        uint32_t bestPos = NONE, secBestPos = NONE;
        const bool dpRead = fromDP == k;
        if (bpos[k] != NONE) { bestScore[k] = best.score[k]; bestPos = bpos[k]; bestHitNum[k] = 1; secBestScore[k] = dpRead ? 0 : 999; }
        if (bpos[k] != NONE && num > 1) {
            for (uint32_t i = 0; i < num; ++i) {
                if ((int32_t)i == bestIndex || algn[i].whichFromDP != fromDP) continue;
                const int sc = algn[i].score[k];
                const uint32_t pos = algn[i].ambPosition[k];
                if (!dpRead) {                                            // mismatches: smaller is better, every entry counts once
                    if (sc <= bestScore[k]) {
                        if (sc == bestScore[k]) { if (pos != bestPos) ++bestHitNum[k]; }
                        else { secBestScore[k] = bestScore[k]; secBestHitNum[k] = bestHitNum[k]; secBestPos = bestPos; bestScore[k] = sc; bestHitNum[k] = 1; bestPos = pos; isBestHit[k] = 0; }
                    } else if (sc <= secBestScore[k]) {
                        if (sc == secBestScore[k]) { if (pos != secBestPos) ++secBestHitNum[k]; }
                        else { secBestScore[k] = sc; secBestPos = pos; secBestHitNum[k] = 1; }
                    }
                } else {                                                  // DP scores: larger is better, an entry counts with its ties
                    const int same = algn[i].numSameScore;
                    if (sc >= bestScore[k]) {
                        if (sc == bestScore[k]) { if (pos != bestPos) bestHitNum[k] += same; }
                        else { secBestScore[k] = bestScore[k]; secBestHitNum[k] = bestHitNum[k]; secBestPos = bestPos; bestScore[k] = sc; bestPos = pos; bestHitNum[k] = same; isBestHit[k] = 0; }
                    } else if (sc >= secBestScore[k]) {
                        if (sc == secBestScore[k]) { if (pos != secBestPos) secBestHitNum[k] += same; }
                        else { secBestScore[k] = sc; secBestPos = pos; secBestHitNum[k] = same; }
                    }
                }
            }
        }
        if (x0[k] > (k ? 1 : 0)) {
            if (!dpRead) {                                                // the read came from the search: its own counts
                bestHitNum[k] = x0[k]; secBestHitNum[k] = x1[k];
                if (mismatch[k] < best.score[k]) isBestHit[k] = 0;
            } else {
                const int x0Score = mismatch[k] * cfg->dpMisMatchScore + (readlen[k] - mismatch[k]) * cfg->dpMatchScore;
                if (x0Score >= bestScore[k]) {
                    if (x0[k] > bestHitNum[k]) bestHitNum[k] = x0[k];
                    if (x1[k] > secBestHitNum[k]) secBestHitNum[k] = x1[k];
                    if (x0Score > bestScore[k]) isBestHit[k] = 0;
                }
            }
        }
    }
    const bool lists = cfg->alignmentType == 1 || cfg->alignmentType == 2;
    int mapq[2] = {0, 0};
    if (both) {
        if (!lists) mapq[0] = mapq[1] = 255;
        else {
            if (cfg->bwaLikeScore) {
                const int dp = fromDP == 0 ? 0 : 1, se = 1 - dp;          // (whichFromDP 2 takes the second form, like the reference)
                const int op = bestPairScore[se] * cfg->dpMisMatchScore + (readlen[se] - bestPairScore[se]) * cfg->dpMatchScore + bestPairScore[dp];
                const int subop = secBestPairNum > 0 ? secBestPairScore[se] * cfg->dpMisMatchScore + (readlen[se] - secBestPairScore[se]) * cfg->dpMatchScore + secBestPairScore[dp] : 0;
                s3_mapq_bwa_pair(bestHitNum[0], secBestHitNum[0], bestHitNum[1], secBestHitNum[1], op, bestPairNum, subop, secBestPairNum, readlen1, readlen2, &mapq[0], &mapq[1]);
            } else {
                int m[2];
                const int dp = fromDP == 0 ? 0 : 1;
                for (int k = 0; k < 2; ++k) {
                    const int aq = cfg->isFastq == 1 ? sd[k].avgQual : 20;
                    m[k] = k == dp ? s3_mapq_pair_end_dp(best.score[k], readlen[k] * cfg->dpMatchScore, aq, bestHitNum[k], secBestHitNum[k], bestScore[k], secBestScore[k], isBestHit[k],
                                                         numSimilar, cfg->maxMAPQ, cfg->minMAPQ)
                                   : s3_mapq_pair_end(best.score[k], aq, bestHitNum[k], secBestHitNum[k], isBestHit[k], (uint32_t)numSimilar, cfg->maxMAPQ, cfg->minMAPQ);
                }
                mapq[0] = mapq[1] = s3_mapq_of_pair(m[0], m[1]);
            }
            for (int k = 0; k < 2; ++k) if (sd[k].trim) mapq[k] = 0;
        }
    }
    for (int k = 0; k < 2; ++k) {
        std::string xa;
        if (bpos[k] != NONE && num > 1) {
            char nb[24];
            for (uint32_t i = 0; i < num; ++i) {
                if ((int32_t)i == bestIndex || algn[i].whichFromDP != fromDP) continue;
                if (cfg->alignmentType == 2 && (algn[i].score[0] != bestPairScore[0] || algn[i].score[1] != bestPairScore[1])) continue;
                // an entry whose DP side missed its cutoff (whichFromDP 2) has no position for that read: with a reported entry of that kind the
                // reference translates 0xFFFFFFFF through tables that do not reach it; nothing is listed here
                if (algn[i].ambPosition[k] == NONE) continue;
                unsigned long long t;
                uint32_t c;
                chr_and_pos(g, algn[i].ambPosition[k], &t, &c);
                xa += g->chrNames[c - 1];
                xa.push_back(',');
                xa.push_back(algn[i].strand[k] == 2 ? '-' : '+');
                xa.append(nb, write_num((long long)t, nb));
                xa.push_back(',');
                if (algn[i].whichFromDP != k) { xa.append(nb, write_num(readlen[k], nb)); xa += "M,"; xa.append(nb, write_num(algn[i].score[k], nb)); }
                else { s3_special_to_sam(algn[i].cigar, strlen(algn[i].cigar), xa); xa.push_back(','); xa.append(nb, write_num(algn[i].editdist, nb)); }
                xa.push_back(';');
            }
        }
        if (cfg->alignmentType == 4) { bestHitNum[k] = -1; secBestHitNum[k] = -1; }
        else if (!lists) secBestHitNum[k] = -1;
        s3_sam_record &r = out[k];
        if (bpos[k] != NONE) {
            if (bpos[1 - k] == NONE) {
                const int aq = cfg->isFastq == 1 ? sd[k].avgQual : 20;
                mapq[k] = fromDP == k ? s3_mapq_unique_dp(bestHitNum[k], best.score[k], readlen[k] * cfg->dpMatchScore, aq, cfg->maxMAPQ, cfg->minMAPQ)
                                      : s3_mapq_unique(bestHitNum[k], best.score[k], aq, cfg->maxMAPQ, cfg->minMAPQ);
                if (sd[k].trim) mapq[k] = 0;
            }
            const std::string *cig = !newCigar[k].empty() ? &newCigar[k] : (fromDP == k ? &sd[k].cigar : NULL);
            record_body(r, d, readlen[k], name[k], query[k], qual[k], strand[k], xa, cig, false, sd[k].mism, sd[k].mism + sd[k].gapExt, bestHitNum[k], secBestHitNum[k],
                        sd[k].gapOpen, sd[k].gapExt, sd[k].md, mapq[k], cfg->readGroup, cfg->isPrintMDNM != 0);
        } else {
            record_body(r, d, readlen[k], name[k], query[k], qual[k], strand[k], none, NULL, true, 0, 0, 0, 0, 0, 0, none, 0, cfg->readGroup, false);
        }
        r.flag = (uint16_t)(1 | (both ? 2 : 0) | (k ? 128 : 64) | (bpos[k] == NONE ? 4 : 0) | (bpos[1 - k] == NONE ? 8 : 0) |
                            (bpos[k] != NONE && best.strand[k] == 2 ? 16 : 0) | (bpos[1 - k] != NONE && best.strand[1 - k] == 2 ? 32 : 0));
        const int m = 1 - k;
        r.tid = chr[k] == 0 ? (chr[m] == 0 ? -1 : (int32_t)chr[m] - 1) : (int32_t)chr[k] - 1;
        r.pos = tp[k] == 0 ? (tp[m] == 0 ? -1 : (int32_t)(tp[m] - 1)) : (int32_t)(tp[k] - 1);
        r.mtid = chr[m] == 0 ? (chr[k] == 0 ? -1 : (int32_t)chr[k] - 1) : (int32_t)chr[m] - 1;
        r.mpos = tp[m] == 0 ? (tp[k] == 0 ? -1 : (int32_t)(tp[k] - 1)) : (int32_t)(tp[m] - 1);
        if (bestInsert > 0) r.isize = tp[k] > tp[m] ? -(int32_t)(tp[k] + (unsigned long long)span[k] - tp[m]) : (int32_t)(tp[m] + (unsigned long long)span[m] - tp[k]);
        else r.isize = 0;
        if ((rc = finish(r, d))) { s3_sam_record_free(&out[0]); s3_sam_record_free(&out[1]); s3_set_error("s3_sam_pair_dp_records: out of host memory"); return rc; }
    }
    return S3_OK;
}


// ---- unproperlypairOutputSAMAPI (BGS-IO.cpp:2582-2930): the two records of a read pair without a valid pairing ---------------------
// Each read is reported on its own -- the first occurrence with the fewest mismatches, X0 / X1 = the occurrences with that count / one
// more, the others in XA:Z, MAPQ = half of s3_mapq_single (at least minMAPQ) -- with the mate's position in the mate fields.
extern "C" int s3_sam_unpaired_records(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_occurrence *occ1, uint32_t numOcc1,
                                       const s3_sam_occurrence *occ2, uint32_t numOcc2, uint32_t peMaxOutputPerRead,
                                       const uint8_t *query1, const uint8_t *query2, const char *qualities1, const char *qualities2,
                                       int32_t readlen1, int32_t readlen2, const char *queryName1, const char *queryName2, s3_sam_record out[2])
{
    if (!out) { s3_set_error("s3_sam_unpaired_records: NULL output"); return S3_EINVAL; }
    memset(out, 0, 2 * sizeof(s3_sam_record));
    if (!g || !cfg || !query1 || !query2 || !qualities1 || !qualities2 || !queryName1 || !queryName2 || !cfg->readGroup || (numOcc1 && !occ1) || (numOcc2 && !occ2) ||
        readlen1 <= 0 || readlen2 <= 0) { s3_set_error("s3_sam_unpaired_records: bad argument"); return S3_EINVAL; }
    if ((numOcc1 || numOcc2) && (!g->packedDNA || !g->segments || !g->ambiguityMap || !g->chrEndPos || !g->chrNames || g->numSegments == 0)) {
        s3_set_error("s3_sam_unpaired_records: incomplete genome description"); return S3_EINVAL;
    }
    const s3_sam_occurrence *occ[2] = {occ1, occ2};
    const uint32_t numOcc[2] = {numOcc1, numOcc2};
    const uint8_t *query[2] = {query1, query2};
    const char *qual[2] = {qualities1, qualities2}, *name[2] = {queryName1, queryName2};
    const int readlen[2] = {readlen1, readlen2};
    const int type = cfg->alignmentType;
    const bool lists = type == 1 || type == 2;
    int best[2] = {-1, -1}, bestScore[2] = {0, 0}, bestNum[2] = {0, 0}, secNum[2] = {0, 0}, mapq[2] = {0, 0}, avgQual[2] = {20, 20};
    unsigned long long tp[2] = {0, 0};
    uint32_t chr[2] = {0, 0};
    std::string md[2];
    for (int k = 0; k < 2; ++k) {
        if (numOcc[k]) {
            best[k] = 0; bestScore[k] = occ[k][0].mismatchCount; bestNum[k] = 1;
            for (uint32_t i = 1; i < numOcc[k]; ++i) {
                const int mm = occ[k][i].mismatchCount;
                if (mm < bestScore[k]) { secNum[k] = bestScore[k] == mm + 1 ? bestNum[k] : 0; best[k] = (int)i; bestScore[k] = mm; bestNum[k] = 1; }
                else if (mm == bestScore[k]) ++bestNum[k];
                else if (mm == bestScore[k] + 1) ++secNum[k];
            }
        }
        if (best[k] >= 0 && (type != 3 || bestNum[k] == 1)) {
            const s3_sam_occurrence &b = occ[k][best[k]];
            chr_and_pos(g, b.ambPosition, &tp[k], &chr[k]);
            md_string(g, query[k], qual[k], (uint32_t)readlen[k], b.ambPosition, b.strand, b.mismatchCount, 0, md[k], &avgQual[k]);
            if (type == 4 || type == 3) mapq[k] = 255;
            else {
                mapq[k] = s3_mapq_single(b.mismatchCount, cfg->isFastq == 1 ? avgQual[k] : 20, bestNum[k], secNum[k], cfg->maxMAPQ, cfg->minMAPQ, cfg->bwaLikeScore) >> 1;
                if (mapq[k] < cfg->minMAPQ) mapq[k] = cfg->minMAPQ;
            }
        } else best[k] = -1;
    }
    std::vector<uint8_t> d;
    const std::string none;
    int rc;
    for (int k = 0; k < 2; ++k) {
        std::string xa;
        if (lists) {
            uint32_t total = 1;
            char nb[24];
            for (uint32_t i = 0; i < numOcc[k] && total < peMaxOutputPerRead; ++i) {
                if ((int)i == best[k]) continue;
                const int mm = occ[k][i].mismatchCount;
                if (type == 2 && mm > bestScore[k]) continue;
                unsigned long long t;
                uint32_t c;
                chr_and_pos(g, occ[k][i].ambPosition, &t, &c);
                xa += g->chrNames[c - 1];
                xa.push_back(',');
                xa.push_back(occ[k][i].strand == 2 ? '-' : '+');
                xa.append(nb, write_num((long long)t, nb));
                xa.push_back(',');
                xa.append(nb, write_num(readlen[k], nb));
                xa += "M,";
                xa.append(nb, write_num(mm, nb));
                xa.push_back(';');
                ++total;
            }
        }
        s3_sam_record &r = out[k];
        const int m = 1 - k;
        if (best[k] >= 0) {
            const s3_sam_occurrence &b = occ[k][best[k]];
            record_body(r, d, readlen[k], name[k], query[k], qual[k], b.strand, xa, NULL, false, b.mismatchCount, b.mismatchCount, type == 4 ? -1 : bestNum[k],
                        lists ? secNum[k] : -1, 0, 0, md[k], mapq[k], cfg->readGroup, cfg->isPrintMDNM != 0);
        } else {
            // (initializeSAMAlgnmt appends a non-empty XA:Z to an unmapped record too)
            record_body(r, d, readlen[k], name[k], query[k], qual[k], 1, none, NULL, true, 0, 0, 0, 0, 0, 0, none, 0, cfg->readGroup, false);
            if (!xa.empty()) { d.clear(); s3_set_error("s3_sam_unpaired_records: internal: XA:Z of an unmapped read"); return S3_EINVAL; }
        }
        r.flag = (uint16_t)(1 | (best[k] < 0 ? 4 : 0) | (best[m] < 0 ? 8 : 0) | (k ? 128 : 64) | (best[k] >= 0 && occ[k][best[k]].strand == 2 ? 16 : 0) |
                            (best[m] >= 0 && occ[m][best[m]].strand == 2 ? 32 : 0));
        r.tid = chr[k] == 0 ? (chr[m] == 0 ? -1 : (int32_t)chr[m] - 1) : (int32_t)chr[k] - 1;
        r.pos = tp[k] == 0 ? (tp[m] == 0 ? -1 : (int32_t)(tp[m] - 1)) : (int32_t)(tp[k] - 1);
        r.mtid = chr[m] == 0 ? (chr[k] == 0 ? -1 : (int32_t)chr[k] - 1) : (int32_t)chr[m] - 1;
        r.mpos = tp[m] == 0 ? (tp[k] == 0 ? -1 : (int32_t)(tp[k] - 1)) : (int32_t)(tp[m] - 1);
        if (chr[0] > 0 && chr[0] == chr[1]) r.isize = tp[m] > tp[k] ? (int32_t)(tp[m] + (unsigned)readlen[m] - tp[k]) : -(int32_t)(tp[k] + (unsigned)readlen[k] - tp[m]);
        else r.isize = 0;
        if ((rc = finish(r, d))) { s3_sam_record_free(&out[0]); s3_sam_record_free(&out[1]); s3_set_error("s3_sam_unpaired_records: out of host memory"); return rc; }
    }
    return S3_OK;
}


// ---- unproperlypairDPOutputSAMAPI (BGS-IO.cpp:2932-3447): a pair without a valid pairing, the reads' alignment lists after DP ------
namespace {
// what convertToCigarStr (PE.cpp:421-486) reports as deletedEnd: the pending match count when the special CIGAR ends in a deletion
int deleted_end(const char *sp)
{
    const size_t len = strlen(sp);
    int cur = 0, curM = 0, out = 0;
    bool written = false;
    for (size_t i = 0; i < len; ++i) {
        const char c = sp[i];
        if (c >= '0' && c <= '9') { cur = cur * 10 + (c - '0'); continue; }
        switch (c) {
        case 'M': case 'm': curM += cur; cur = 0; break;
        case 'D':
            if ((!written && curM == 0) || i == len - 1) { if (i == len - 1) out = curM; break; }
            /* fall through */
        case 'I': case 'S': written = true; curM = 0; cur = 0; break;
        default: break;
        }
    }
    return out;
}
}  // namespace

extern "C" int s3_sam_unpaired_dp_records(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_read_alignment *algn1, uint32_t num1,
                                          const s3_sam_read_alignment *algn2, uint32_t num2, int32_t singleDPcutoffThreshold,
                                          const uint8_t *query1, const uint8_t *query2, const char *qualities1, const char *qualities2,
                                          int32_t readlen1, int32_t readlen2, const char *queryName1, const char *queryName2, s3_sam_record out[2])
{
    if (!out) { s3_set_error("s3_sam_unpaired_dp_records: NULL output"); return S3_EINVAL; }
    memset(out, 0, 2 * sizeof(s3_sam_record));
    if (!g || !cfg || !query1 || !query2 || !qualities1 || !qualities2 || !queryName1 || !queryName2 || !cfg->readGroup || (num1 && !algn1) || (num2 && !algn2) ||
        readlen1 <= 0 || readlen2 <= 0) { s3_set_error("s3_sam_unpaired_dp_records: bad argument"); return S3_EINVAL; }
    if ((num1 || num2) && (!g->packedDNA || !g->segments || !g->ambiguityMap || !g->chrEndPos || !g->chrNames || g->numSegments == 0)) {
        s3_set_error("s3_sam_unpaired_dp_records: incomplete genome description"); return S3_EINVAL;
    }
    const s3_sam_read_alignment *algn[2] = {algn1, algn2};
    const uint32_t num[2] = {num1, num2};
    for (int k = 0; k < 2; ++k) for (uint32_t i = 0; i < num[k]; ++i) if (!algn[k][i].cigar) { s3_set_error("s3_sam_unpaired_dp_records: alignment %u of read %d without a CIGAR", i, k + 1); return S3_EINVAL; }
    const uint8_t *query[2] = {query1, query2};
    const char *qual[2] = {qualities1, qualities2}, *name[2] = {queryName1, queryName2};
    const int readlen[2] = {readlen1, readlen2};
    const int type = cfg->alignmentType;
    const bool lists = type == 1 || type == 2;
    int best[2] = {-1, -1}, bestScore[2] = {0, 0}, bestNum[2] = {0, 0}, secNum[2] = {0, 0}, mapq[2] = {0, 0}, deletedEnd[2] = {0, 0};
    unsigned long long tp[2] = {0, 0};
    uint32_t chr[2] = {0, 0};
    DpSide sd[2];
    std::string newCigar[2];
    int rc;
    for (int k = 0; k < 2; ++k) {
        int x1t1 = 0, x1t2 = 0, secondBestScore = -9999;
        if (num[k]) {
            best[k] = 0; bestScore[k] = algn[k][0].score; bestNum[k] = 1;
            for (uint32_t i = 1; i < num[k]; ++i) {
                if (algn[k][i].score > bestScore[k]) { best[k] = (int)i; bestScore[k] = algn[k][i].score; bestNum[k] = 1; }
                else if (algn[k][i].score == bestScore[k]) ++bestNum[k];
            }
            const int thres = (int)(0.7 * bestScore[k]);
            for (uint32_t i = 0; i < num[k]; ++i) {
                const int sc = algn[k][i].score;
                if (sc >= bestScore[k]) continue;
                if (sc > secondBestScore) secondBestScore = sc;
                if (sc >= thres) ++x1t1; else ++x1t2;
            }
        }
        secNum[k] = (type == 4 || type == 3) ? -1 : x1t1 + x1t2;
        if (best[k] >= 0 && (type != 3 || bestNum[k] == 1)) {
            const s3_sam_read_alignment &b = algn[k][best[k]];
            if (b.isFromDP == 1) {
                if ((rc = dp_side(g, b.cigar, b.ambPosition, readlen[k], qual[k], &tp[k], &chr[k], sd[k]))) return rc;
                if (!sd[k].trim) deletedEnd[k] = deleted_end(b.cigar);
            } else {
                sd[k].trim = chr_and_pos_checked(g, (uint32_t)readlen[k], b.ambPosition, &tp[k], &chr[k], newCigar[k]);
                md_string(g, query[k], qual[k], (uint32_t)readlen[k], b.ambPosition, b.strand, b.editdist, sd[k].trim, sd[k].md, &sd[k].avgQual);
                sd[k].mism = 0;
                for (char c : sd[k].md) sd[k].mism += c > '9';
            }
            if (type == 4 || type == 3) { mapq[k] = 255; if (k == 0 && sd[k].trim) mapq[k] = 0; }        // (only the first read's trimmed record loses the 255)
            else {
                mapq[k] = s3_mapq_single_dp(readlen[k] * cfg->dpMatchScore, cfg->isFastq == 1 ? sd[k].avgQual : 20, bestNum[k], x1t1, x1t2, bestScore[k], secondBestScore,
                                            cfg->maxMAPQ, cfg->minMAPQ, singleDPcutoffThreshold, cfg->bwaLikeScore);
                if (!cfg->bwaLikeScore) mapq[k] >>= 1;
                if (mapq[k] < cfg->minMAPQ) mapq[k] = cfg->minMAPQ;
                if (sd[k].trim) mapq[k] = 0;
            }
        } else best[k] = -1;
    }
    std::vector<uint8_t> d;
    const std::string none;
    for (int k = 0; k < 2; ++k) {
        std::string xa;
        if (lists) {
            char nb[24];
            for (uint32_t i = 0; i < num[k]; ++i) {
                if ((int)i == best[k]) continue;
                if (type == 2 && algn[k][i].score < bestScore[k]) continue;
                unsigned long long t;
                uint32_t c;
                chr_and_pos(g, algn[k][i].ambPosition, &t, &c);
                xa += g->chrNames[c - 1];
                xa.push_back(',');
                xa.push_back(algn[k][i].strand == 2 ? '-' : '+');
                xa.append(nb, write_num((long long)t, nb));
                xa.push_back(',');
                s3_special_to_sam(algn[k][i].cigar, strlen(algn[k][i].cigar), xa);
                xa.push_back(',');
                xa.append(nb, write_num(algn[k][i].editdist, nb));
                xa.push_back(';');
            }
        }
        s3_sam_record &r = out[k];
        const int m = 1 - k;
        if (best[k] >= 0) {
            const s3_sam_read_alignment &b = algn[k][best[k]];
            const std::string *cig = b.isFromDP ? &sd[k].cigar : (newCigar[k].empty() ? NULL : &newCigar[k]);
            record_body(r, d, readlen[k], name[k], query[k], qual[k], b.strand, xa, cig, false, sd[k].mism, sd[k].mism + sd[k].gapExt, bestNum[k], secNum[k], sd[k].gapOpen, sd[k].gapExt,
                        sd[k].md, mapq[k], cfg->readGroup, cfg->isPrintMDNM != 0);
        } else {
            record_body(r, d, readlen[k], name[k], query[k], qual[k], 1, none, NULL, true, 0, 0, 0, 0, 0, 0, none, 0, cfg->readGroup, false);
            if (!xa.empty()) { s3_set_error("s3_sam_unpaired_dp_records: internal: XA:Z of an unmapped read"); return S3_EINVAL; }
        }
        r.flag = (uint16_t)(1 | (best[k] < 0 ? 4 : 0) | (best[m] < 0 ? 8 : 0) | (k ? 128 : 64) | (best[k] >= 0 && algn[k][best[k]].strand == 2 ? 16 : 0) |
                            (best[m] >= 0 && algn[m][best[m]].strand == 2 ? 32 : 0));
        r.tid = chr[k] == 0 ? (chr[m] == 0 ? -1 : (int32_t)chr[m] - 1) : (int32_t)chr[k] - 1;
        r.pos = tp[k] == 0 ? (tp[m] == 0 ? -1 : (int32_t)(tp[m] - 1)) : (int32_t)(tp[k] - 1);
        r.mtid = chr[m] == 0 ? (chr[k] == 0 ? -1 : (int32_t)chr[k] - 1) : (int32_t)chr[m] - 1;
        r.mpos = tp[m] == 0 ? (tp[k] == 0 ? -1 : (int32_t)(tp[k] - 1)) : (int32_t)(tp[m] - 1);
        if (chr[0] > 0 && chr[0] == chr[1])
            r.isize = tp[m] > tp[k] ? (int32_t)(tp[m] - (unsigned long long)deletedEnd[m] + (unsigned long long)readlen[m] - tp[k])
                                    : -(int32_t)(tp[k] - (unsigned long long)deletedEnd[k] + (unsigned long long)readlen[k] - tp[m]);
        else r.isize = 0;
        if ((rc = finish(r, d))) { s3_sam_record_free(&out[0]); s3_sam_record_free(&out[1]); s3_set_error("s3_sam_unpaired_dp_records: out of host memory"); return rc; }
    }
    return S3_OK;
}


// ---- SingleAnsOutputSAMAPI (BGS-IO.cpp:5774-5827): the record of a single read reported with one alignment (unique / random best) ----
extern "C" int s3_sam_single_answer_record(const s3_sam_genome *g, const s3_sam_config *cfg, uint32_t ambPosition, int32_t strand, int32_t numMismatch, int32_t bestHitNum,
                                           const uint8_t *query, const char *qualities, int32_t readlen, const char *queryName, s3_sam_record *out)
{
    if (!out) { s3_set_error("s3_sam_single_answer_record: NULL output"); return S3_EINVAL; }
    memset(out, 0, sizeof *out);
    if (!g || !cfg || !query || !qualities || !queryName || !cfg->readGroup || readlen <= 0 || !g->packedDNA || !g->segments || !g->ambiguityMap || g->numSegments == 0) {
        s3_set_error("s3_sam_single_answer_record: bad argument"); return S3_EINVAL;
    }
    unsigned long long tp;
    uint32_t chr;
    chr_and_pos(g, ambPosition, &tp, &chr);
    std::string md;
    int avgQual = 20;
    md_string(g, query, qualities, (uint32_t)readlen, ambPosition, strand, numMismatch, 0, md, &avgQual);
    const int mapq = bestHitNum > 0 ? s3_mapq_unique(bestHitNum, numMismatch, cfg->isFastq == 1 ? avgQual : 20, cfg->maxMAPQ, cfg->minMAPQ) : 255;
    std::vector<uint8_t> d;
    const std::string none;
    record_body(*out, d, readlen, queryName, query, qualities, strand, none, NULL, false, numMismatch, numMismatch, bestHitNum, -1, 0, 0, md, mapq, cfg->readGroup, cfg->isPrintMDNM != 0);
    out->flag = (uint16_t)(strand == 2 ? 16 : 0);
    out->tid = (int32_t)chr - 1; out->pos = (int32_t)(tp - 1);
    out->mtid = -1; out->mpos = -1; out->isize = 0;
    const int rc = finish(*out, d);
    if (rc) s3_set_error("s3_sam_single_answer_record: out of host memory");
    return rc;
}

// ---- the SAM text line of a record: bam_format1 (samtools-0.1.18/bam.c:243-329, what samwrite prints for a text file) ----------
static int format_line_into(const s3_sam_record *r, const char *const *chrNames, uint32_t numChr, std::string &s)      // appends the line to s
{
    if (!r || !r->data) { s3_set_error("s3_sam_format_line: NULL argument"); return S3_EINVAL; }
    if ((r->tid >= 0 || r->mtid >= 0) && !chrNames) { s3_set_error("s3_sam_format_line: chromosome names needed"); return S3_EINVAL; }
    if (r->tid >= (int32_t)numChr || r->mtid >= (int32_t)numChr) { s3_set_error("s3_sam_format_line: chromosome id out of range"); return S3_EINVAL; }
    const size_t seqBytes = ((size_t)r->l_qseq + 1) / 2, fixed = (size_t)r->l_qname + 4 * (size_t)r->n_cigar + seqBytes + (size_t)r->l_qseq;
    if (r->l_qname == 0 || fixed > (size_t)r->data_len) { s3_set_error("s3_sam_format_line: the record's data is shorter than its fields"); return S3_EINVAL; }
    static const char nt16[] = "=ACMGRSVTWYHKDBN";                       // bam_nt16_rev_table
    const uint8_t *d = r->data, *cig = d + r->l_qname, *seq = cig + 4 * (size_t)r->n_cigar, *qual = seq + seqBytes, *aux = qual + r->l_qseq, *end = d + r->data_len;
    const size_t room = 2 * (size_t)r->data_len + 96;                       // a hint; growth stays geometric when lines are appended to one text
    if (s.capacity() - s.size() < room) s.reserve(std::max(2 * s.capacity(), s.size() + room));
    char nb[24];
    auto num = [&](long long v) { s.append(nb, write_num(v, nb)); };
    s.append((const char *)d, (size_t)r->l_qname - 1); s.push_back('\t');
    num(r->flag); s.push_back('\t');
    if (r->tid < 0) s += "*\t"; else { s += chrNames[r->tid]; s.push_back('\t'); }
    num((long long)r->pos + 1); s.push_back('\t'); num(r->qual); s.push_back('\t');
    if (r->n_cigar == 0) s.push_back('*');
    else for (uint32_t i = 0; i < r->n_cigar; ++i) {
        uint32_t c;
        memcpy(&c, cig + 4 * (size_t)i, 4);
        num(c >> 4); s.push_back("MIDNSHP=X"[c & 15u]);
    }
    s.push_back('\t');
    if (r->mtid < 0) s += "*\t"; else if (r->mtid == r->tid) s += "=\t"; else { s += chrNames[r->mtid]; s.push_back('\t'); }
    num((long long)r->mpos + 1); s.push_back('\t'); num(r->isize); s.push_back('\t');
    if (r->l_qseq) {
        const size_t at = s.size(), n = (size_t)r->l_qseq;
        const bool noQual = qual[0] == 0xFF;
        s.resize(at + n + 1 + (noQual ? 1 : n));
        char *p = &s[at];
        for (size_t i = 0; i < n; ++i) p[i] = nt16[(seq[i >> 1] >> ((~i & 1) << 2)) & 0xF];
        p[n] = '\t';
        if (noQual) p[n + 1] = '*'; else for (size_t i = 0; i < n; ++i) p[n + 1 + i] = (char)(qual[i] + 33);
    } else s += "*\t*";
    for (const uint8_t *p = aux; p + 3 <= end;) {
        const char type = (char)p[2];
        s.push_back('\t'); s.push_back((char)p[0]); s.push_back((char)p[1]); s.push_back(':');
        p += 3;
        if (type == 'A') { s += "A:"; s.push_back((char)*p); ++p; }
        else if (type == 'C') { s += "i:"; num(*p); ++p; }
        else if (type == 'c') { s += "i:"; num((int8_t)*p); ++p; }
        else if (type == 'S') { uint16_t v; memcpy(&v, p, 2); s += "i:"; num(v); p += 2; }
        else if (type == 's') { int16_t v; memcpy(&v, p, 2); s += "i:"; num(v); p += 2; }
        else if (type == 'I') { uint32_t v; memcpy(&v, p, 4); s += "i:"; num(v); p += 4; }
        else if (type == 'i') { int32_t v; memcpy(&v, p, 4); s += "i:"; num(v); p += 4; }
        else if (type == 'Z' || type == 'H') { s.push_back(type); s.push_back(':'); while (p < end && *p) s.push_back((char)*p++); ++p; }
        else { s3_set_error("s3_sam_format_line: tag type '%c' is not one the record writers produce", type); return S3_EINVAL; }
    }
    return S3_OK;
}

extern "C" int s3_sam_format_line(const s3_sam_record *r, const char *const *chrNames, uint32_t numChr, char **line)
{
    if (!line) { s3_set_error("s3_sam_format_line: NULL argument"); return S3_EINVAL; }
    *line = NULL;
    std::string s;
    const int rc = format_line_into(r, chrNames, numChr, s);
    if (rc) return rc;
    *line = (char *)malloc(s.size() + 1);
    if (!*line) { s3_set_error("s3_sam_format_line: out of host memory"); return S3_ENOMEM; }
    memcpy(*line, s.c_str(), s.size() + 1);
    return S3_OK;
}

// ---- which entry of a pair's results is reported: the scans of outputDeepDPResult2 / outputDPResult2 (OutputDPResult.cpp) ---------
// Deep DP (:590-760): the first entry sets the score to beat -- score1 + score2, or the aligned read's score when the other is unaligned,
// -127 when neither is -- and a later entry takes over when its own is strictly larger.
extern "C" int32_t s3_sam_pick_deep_dp(const s3_sam_deep_alignment *algn, uint32_t num)
{
    if (!algn || num == 0) return -1;
    const uint32_t NONE = 0xFFFFFFFFu;
    auto own = [&](const s3_sam_deep_alignment &a, bool *any) {
        const bool u1 = a.ambPosition[0] == NONE, u2 = a.ambPosition[1] == NONE;
        *any = !(u1 && u2);
        return u1 && !u2 ? a.score[1] : (u2 && !u1 ? a.score[0] : (!u1 && !u2 ? a.score[0] + a.score[1] : -127));
    };
    bool any;
    int32_t best = 0, maxScore = own(algn[0], &any);
    for (uint32_t i = 1; i < num; ++i) {
        const int32_t s = own(algn[i], &any);
        if (any && s > maxScore) { best = (int32_t)i; maxScore = s; }
    }
    return best;
}

// Default DP / mate rescue (:263-350): the fewest mismatches of the read that came from the search, then the highest DP score of the
// other; an entry whose DP read missed its cutoff (position 0xFFFFFFFF) competes with its mismatches alone.  (An entry with neither
// read aligned leaves the reference's running values as they were; here it neither sets nor beats anything.)
extern "C" int32_t s3_sam_pick_pair_dp(const s3_sam_dp_pairing *algn, uint32_t num)
{
    if (!algn || num == 0) return -1;
    const uint32_t NONE = 0xFFFFFFFFu;
    int32_t best = 0, minMismatch = 0x7FFFFFFF, maxScore = -127;
    for (uint32_t i = 0; i < num; ++i) {
        const s3_sam_dp_pairing &a = algn[i];
        const bool u1 = a.ambPosition[0] == NONE, u2 = a.ambPosition[1] == NONE;
        int32_t cm, cs;
        if (u1 && !u2) { cm = a.score[1]; cs = -127; }
        else if (u2 && !u1) { cm = a.score[0]; cs = -127; }
        else if (!u1 && !u2) { cm = a.whichFromDP == 1 ? a.score[0] : a.score[1]; cs = a.whichFromDP == 1 ? a.score[1] : a.score[0]; }
        else continue;
        const bool paired = !u1 && !u2;
        if (i == 0 || cm < minMismatch || (paired && cm == minMismatch && cs > maxScore)) { best = (int32_t)i; minMismatch = cm; maxScore = cs; }
    }
    return best;
}

// ---- whole batches as SAM text -------------------------------------------------------------------------------------------------
// The loops around the record writers that the reference's output threads run per read (hostKernel -> OCCOutputSAMAPI,
// CPUfunctions.cpp:1887-1905; outputDPSingleResult2 -> SingleDPOutputSAMAPI, OutputDPResult.cpp:938-1058): one record per read through
// the entries above, printed by s3_sam_format_line, lines in read order.  Reads are independent, so the batch is cut into contiguous
// slices, one per host thread, and the slices' text is concatenated.
namespace {
struct BatchError { std::atomic<int> rc{0}; char msg[512]; };

void batch_fail(BatchError &e, int rc)
{
    int expected = 0;
    if (e.rc.compare_exchange_strong(expected, rc)) { strncpy(e.msg, s3_last_error(), sizeof e.msg - 1); e.msg[sizeof e.msg - 1] = 0; }
}

// runs `one(r, text)` for r in [0, n) on numThreads host threads and joins the slices' text
template <typename F>
int batch_text(const char *what, uint64_t n, uint32_t numThreads, char **text, uint64_t *textBytes, F one)
{
    if (!text || !textBytes) { s3_set_error("%s: NULL output", what); return S3_EINVAL; }
    *text = NULL; *textBytes = 0;
    uint32_t T = numThreads ? numThreads : std::max(1u, std::thread::hardware_concurrency());
    if ((uint64_t)T > n) T = (uint32_t)std::max<uint64_t>(n, 1);
    std::vector<std::string> part(T);
    BatchError err;
    auto work = [&](uint32_t t) {
        const uint64_t a = n * t / T, b = n * (t + 1) / T;
        try {
            part[t].reserve((size_t)(b - a) * 352 + 4096);                 // about a line of a 100-base read each; growth stays geometric beyond it
            for (uint64_t r = a; r < b && err.rc.load(std::memory_order_relaxed) == 0; ++r) {
                const int rc = one(r, part[t]);
                if (rc) { batch_fail(err, rc); return; }
            }
        } catch (...) { s3_set_error("%s: out of host memory", what); batch_fail(err, S3_ENOMEM); }
    };
    std::vector<std::thread> th;
    uint32_t started = 1;
    try { for (; started < T; ++started) th.emplace_back(work, started); } catch (...) { }
    work(0);
    for (uint32_t t = started; t < T; ++t) work(t);                       // slices whose thread could not be started run here
    for (auto &x : th) x.join();
    if (err.rc) { s3_set_error("%s", err.msg); return err.rc; }
    size_t total = 0;
    for (auto &p : part) total += p.size();
    char *out = (char *)malloc(total + 1);
    if (!out) { s3_set_error("%s: out of host memory", what); return S3_ENOMEM; }
    std::vector<size_t> at(T + 1, 0);
    for (uint32_t t = 0; t < T; ++t) at[t + 1] = at[t] + part[t].size();
    auto copy = [&](uint32_t t) { memcpy(out + at[t], part[t].data(), part[t].size()); std::string().swap(part[t]); };
    th.clear();
    started = 1;
    try { for (; started < T; ++started) th.emplace_back(copy, started); } catch (...) { }
    copy(0);
    for (uint32_t t = started; t < T; ++t) copy(t);
    for (auto &x : th) x.join();
    out[total] = 0;
    *text = out; *textBytes = total;
    return S3_OK;
}

int append_line(const s3_sam_genome *g, s3_sam_record *rec, std::string &text)
{
    const size_t at = text.size();
    const int rc = format_line_into(rec, g->chrNames, g->numChr, text);
    s3_sam_record_free(rec);
    if (rc) { text.resize(at); return rc; }
    text.push_back('\n');
    return S3_OK;
}

bool reads_ok(const s3_sam_reads *rd) { return rd && rd->bases && rd->qualities && rd->readLengths && rd->names && rd->rowBytes; }
}  // namespace

extern "C" int s3_sam_single_batch_text(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_reads *reads, uint64_t numReads,
                                        const uint32_t *occOffsets, const uint32_t *positions, const uint8_t *occFlags, uint32_t numThreads,
                                        char **text, uint64_t *textBytes)
{
    if (text) *text = NULL;
    if (textBytes) *textBytes = 0;
    if (!g || !cfg || !reads_ok(reads) || !occOffsets || (numReads && occOffsets[numReads] && (!positions || !occFlags))) { s3_set_error("s3_sam_single_batch_text: NULL argument"); return S3_EINVAL; }
    for (uint64_t r = 0; r < numReads; ++r) {
        if (occOffsets[r + 1] < occOffsets[r]) { s3_set_error("s3_sam_single_batch_text: occOffsets decrease at read %llu", (unsigned long long)r); return S3_EINVAL; }
        if (reads->readLengths[r] == 0 || reads->readLengths[r] > reads->rowBytes) { s3_set_error("s3_sam_single_batch_text: read %llu has length %u (rows of %u)", (unsigned long long)r, reads->readLengths[r], reads->rowBytes); return S3_EINVAL; }
    }
    return batch_text("s3_sam_single_batch_text", numReads, numThreads, text, textBytes, [&](uint64_t r, std::string &out) {
        const uint32_t a = occOffsets[r], n = occOffsets[r + 1] - a;
        std::vector<s3_sam_occurrence> occ(n);
        for (uint32_t k = 0; k < n; ++k) { occ[k].ambPosition = positions[a + k]; occ[k].strand = occFlags[2 * (size_t)(a + k)]; occ[k].mismatchCount = occFlags[2 * (size_t)(a + k) + 1]; occ[k].pad[0] = occ[k].pad[1] = 0; }
        s3_sam_record rec;
        const uint8_t *q = reads->bases + r * reads->rowBytes;
        const char *ql = reads->qualities + r * reads->rowBytes;
        const int32_t len = (int32_t)reads->readLengths[r];
        int rc;
        // the unique-best / random-best report types write the first occurrence alone (SingleAnsOutputSAMAPI with bestHitNum 1, CPUfunctions.cpp:1862-1925);
        // unique-best reports nothing for a read with more than one occurrence (noAnsOutputSAMAPI)
        if (cfg->alignmentType == 3 || cfg->alignmentType == 4) {
            if (n == 1 || (n > 1 && cfg->alignmentType == 4)) rc = s3_sam_single_answer_record(g, cfg, occ[0].ambPosition, occ[0].strand, occ[0].mismatchCount, 1, q, ql, len, reads->names[r], &rec);
            else rc = s3_sam_single_record(g, cfg, NULL, 0, q, ql, len, reads->names[r], &rec);
        } else rc = s3_sam_single_record(g, cfg, occ.data(), n, q, ql, len, reads->names[r], &rec);
        return rc ? rc : append_line(g, &rec, out);
    });
}

// The hits of s3_single_dp_align are in candidate order, the candidates of a read next to each other; a record is written for every read
// that has a hit (reads without one are reported by the caller's last stage: s3_sam_single_record with no occurrence).  Where the
// reference's DP batches cut a read's candidates in two its writer sees two groups; there are no batch borders here.
extern "C" int s3_sam_single_dp_batch_text(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_reads *reads, uint64_t numReads,
                                           const s3_dp_hit *hits, uint64_t numHits, const uint32_t *runs, uint64_t numRuns, s3_dp_scores scores,
                                           int32_t singleDPcutoffThreshold, uint32_t numThreads, char **text, uint64_t *textBytes)
{
    if (text) *text = NULL;
    if (textBytes) *textBytes = 0;
    if (!g || !cfg || !reads_ok(reads) || (numHits && (!hits || !runs))) { s3_set_error("s3_sam_single_dp_batch_text: NULL argument"); return S3_EINVAL; }
    std::vector<uint64_t> first;                                          // the first hit of every group of equal read ids
    for (uint64_t i = 0; i < numHits; ++i) {
        const s3_dp_hit &h = hits[i];
        if (h.readID >= numReads || (uint64_t)h.runOffset + h.numRuns > numRuns) { s3_set_error("s3_sam_single_dp_batch_text: hit %llu points outside the batch", (unsigned long long)i); return S3_EINVAL; }
        if (reads->readLengths[h.readID] == 0 || reads->readLengths[h.readID] > reads->rowBytes) { s3_set_error("s3_sam_single_dp_batch_text: read %u has length %u (rows of %u)", h.readID, reads->readLengths[h.readID], reads->rowBytes); return S3_EINVAL; }
        if (i == 0 || h.readID != hits[i - 1].readID) first.push_back(i);
    }
    first.push_back(numHits);
    return batch_text("s3_sam_single_dp_batch_text", first.size() - 1, numThreads, text, textBytes, [&](uint64_t grp, std::string &out) {
        const uint64_t a = first[grp], n = first[grp + 1] - a;
        const uint32_t r = hits[a].readID, len = reads->readLengths[r];
        std::vector<s3_sam_dp_alignment> al(n);
        std::vector<std::string> cig(n);
        for (uint64_t k = 0; k < n; ++k) {
            const s3_dp_hit &h = hits[a + k];
            cig[k].resize(10 * (size_t)h.numRuns + 16);
            uint32_t clen = 0; int32_t edit = 0, span = 0;
            const int rc = s3_runs_decode(runs + h.runOffset, h.numRuns, len, h.score, scores, &cig[k][0], (uint32_t)cig[k].size(), &clen, &edit, &span);
            if (rc) return rc;
            memset(&al[k], 0, sizeof al[k]);
            al[k].ambPosition = h.pos; al[k].strand = h.strand; al[k].score = h.score; al[k].editdist = edit; al[k].cigar = cig[k].c_str();
        }
        s3_sam_record rec;
        const int rc = s3_sam_single_dp_record(g, cfg, al.data(), (uint32_t)n, singleDPcutoffThreshold, reads->bases + (size_t)r * reads->rowBytes, reads->qualities + (size_t)r * reads->rowBytes,
                                               (int32_t)len, reads->names[r], &rec);
        return rc ? rc : append_line(g, &rec, out);
    });
}

namespace {
// x0 / x1 / mismatch of the two reads of a pair as the DP writers take them (hspaux->x0_array / x1_array / mismatch_array, filled by
// hostKernel, CPUfunctions.cpp:2061-2141): the chain's per-read statistics, zeros for a read without an occurrence or without statistics
void pair_counts(const s3_pe_read_stats *st, uint32_t evenRead, int32_t x0[2], int32_t x1[2], int32_t mm[2])
{
    for (int k = 0; k < 2; ++k) {
        if (!st) { x0[k] = x1[k] = mm[k] = 0; continue; }
        const s3_pe_read_stats &s = st[evenRead + k];
        x0[k] = (int32_t)s.x0; x1[k] = (int32_t)s.x1; mm[k] = s.x0 ? (int32_t)s.minMismatch : 0;
    }
}

int decode_runs(const uint32_t *runs, uint32_t numRuns, uint32_t readLength, int32_t score, s3_dp_scores sc, std::string &cigar, int32_t *edit, int32_t *span)
{
    cigar.assign(10 * (size_t)numRuns + 16, '\0');
    uint32_t len = 0;
    const int rc = s3_runs_decode(runs, numRuns, readLength, score, sc, &cigar[0], (uint32_t)cigar.size(), &len, edit, span);
    if (rc == S3_OK) cigar.resize(len);
    return rc;
}

int append_pair(const s3_sam_genome *g, s3_sam_record rec[2], std::string &text)
{
    const int a = append_line(g, &rec[0], text), b = append_line(g, &rec[1], text);
    return a ? a : b;
}
}  // namespace

// outputDeepDPResult2 (OutputDPResult.cpp:590-760) over the hits of s3_deep_dp_align / s3_pe_deep_dp: the hits of a pair are next to each
// other; the DeepDPAlignResult fields the engine derives (edit distances, insert size: DV-DPfunctions.cu:3765-3815) come from the runs.
extern "C" int s3_sam_deep_dp_batch_text(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_reads *reads, uint64_t numReads,
                                         const s3_deep_dp_hit *hits, uint64_t numHits, const uint32_t *runs, uint64_t numRuns, s3_dp_scores scores,
                                         const s3_pe_read_stats *readStats, uint32_t numThreads, char **text, uint64_t *textBytes)
{
    if (text) *text = NULL;
    if (textBytes) *textBytes = 0;
    if (!g || !cfg || !reads_ok(reads) || (numHits && (!hits || !runs))) { s3_set_error("s3_sam_deep_dp_batch_text: NULL argument"); return S3_EINVAL; }
    std::vector<uint64_t> first;
    for (uint64_t i = 0; i < numHits; ++i) {
        const s3_deep_dp_hit &h = hits[i];
        if ((h.readID & 1u) || (uint64_t)h.readID + 1 >= numReads || (uint64_t)h.runOffset1 + h.numRuns1 > numRuns || (uint64_t)h.runOffset2 + h.numRuns2 > numRuns) {
            s3_set_error("s3_sam_deep_dp_batch_text: hit %llu points outside the batch", (unsigned long long)i); return S3_EINVAL;
        }
        for (int k = 0; k < 2; ++k) if (reads->readLengths[h.readID + k] == 0 || reads->readLengths[h.readID + k] > reads->rowBytes) {
            s3_set_error("s3_sam_deep_dp_batch_text: read %u has length %u (rows of %u)", h.readID + k, reads->readLengths[h.readID + k], reads->rowBytes); return S3_EINVAL;
        }
        if (i == 0 || h.readID != hits[i - 1].readID) first.push_back(i);
    }
    first.push_back(numHits);
    return batch_text("s3_sam_deep_dp_batch_text", first.size() - 1, numThreads, text, textBytes, [&](uint64_t grp, std::string &out) {
        const uint64_t a = first[grp], n = first[grp + 1] - a;
        const uint32_t r = hits[a].readID, len1 = reads->readLengths[r], len2 = reads->readLengths[r + 1];
        std::vector<s3_sam_deep_alignment> al(n);
        std::vector<std::string> cig(2 * n);
        for (uint64_t k = 0; k < n; ++k) {
            const s3_deep_dp_hit &h = hits[a + k];
            int32_t ed[2], span[2];
            int rc = decode_runs(runs + h.runOffset1, h.numRuns1, len1, h.score1, scores, cig[2 * k], &ed[0], &span[0]);
            if (!rc) rc = decode_runs(runs + h.runOffset2, h.numRuns2, len2, h.score2, scores, cig[2 * k + 1], &ed[1], &span[1]);
            if (rc) return rc;
            s3_sam_deep_alignment &x = al[k];
            memset(&x, 0, sizeof x);
            x.insertSize = h.pos1 < h.pos2 ? (int32_t)(h.pos2 - h.pos1 + len2) + span[1] : (int32_t)(h.pos1 - h.pos2 + len1) + span[0];
            x.ambPosition[0] = h.pos1; x.ambPosition[1] = h.pos2; x.strand[0] = h.strand1; x.strand[1] = h.strand2;
            x.score[0] = h.score1; x.score[1] = h.score2; x.editdist[0] = ed[0]; x.editdist[1] = ed[1];
            x.numSameScore[0] = (int32_t)h.numSame1; x.numSameScore[1] = (int32_t)h.numSame2;
            x.cigar[0] = cig[2 * k].c_str(); x.cigar[1] = cig[2 * k + 1].c_str();
        }
        int32_t x0[2], x1[2], mm[2];
        pair_counts(readStats, r, x0, x1, mm);
        s3_sam_record rec[2];
        const int rc = s3_sam_deep_dp_records(g, cfg, al.data(), (uint32_t)n, s3_sam_pick_deep_dp(al.data(), (uint32_t)n),
                                              reads->bases + (size_t)r * reads->rowBytes, reads->bases + (size_t)(r + 1) * reads->rowBytes,
                                              reads->qualities + (size_t)r * reads->rowBytes, reads->qualities + (size_t)(r + 1) * reads->rowBytes,
                                              (int32_t)len1, (int32_t)len2, reads->names[r], reads->names[r + 1], x0, x1, mm, rec);
        return rc ? rc : append_pair(g, rec, out);
    });
}

// outputDPResult2 (OutputDPResult.cpp:263-420) over the rescue records of s3_pe_align: every record of a pair becomes an AlgnmtDPResult as
// the default-DP engine builds it (DV-DPfunctions.cu:2355-2440: whichFromDP = the DP read's parity, 2 with an unaligned DP side when it
// missed its cutoff; insert size :2388-2400).  A pair none of whose rescues succeeded has no record here: it belongs to the writers of
// improperly paired reads (s3_sam_unpaired_records over the reads' occurrence lists).
extern "C" int s3_sam_pair_dp_batch_text(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_reads *reads, uint64_t numReads,
                                         const s3_pe_dp_result *dp, uint64_t numRecords, const uint32_t *runs, uint64_t numRuns, s3_dp_scores scores,
                                         const s3_pe_read_stats *readStats, uint32_t numThreads, char **text, uint64_t *textBytes)
{
    if (text) *text = NULL;
    if (textBytes) *textBytes = 0;
    if (!g || !cfg || !reads_ok(reads) || (numRecords && !dp)) { s3_set_error("s3_sam_pair_dp_batch_text: NULL argument"); return S3_EINVAL; }
    std::vector<uint64_t> first;
    for (uint64_t i = 0; i < numRecords; ++i) {
        const s3_pe_dp_result &x = dp[i];
        if ((uint64_t)(x.dpReadID | 1u) >= numReads || (x.numRuns && (!runs || (uint64_t)x.runOffset + x.numRuns > numRuns))) {
            s3_set_error("s3_sam_pair_dp_batch_text: record %llu points outside the batch", (unsigned long long)i); return S3_EINVAL;
        }
        for (uint32_t k = x.dpReadID & ~1u; k <= (x.dpReadID | 1u); ++k) if (reads->readLengths[k] == 0 || reads->readLengths[k] > reads->rowBytes) {
            s3_set_error("s3_sam_pair_dp_batch_text: read %u has length %u (rows of %u)", k, reads->readLengths[k], reads->rowBytes); return S3_EINVAL;
        }
        if (i == 0 || (x.dpReadID >> 1) != (dp[i - 1].dpReadID >> 1)) first.push_back(i);
    }
    first.push_back(numRecords);
    return batch_text("s3_sam_pair_dp_batch_text", first.size() - 1, numThreads, text, textBytes, [&](uint64_t grp, std::string &out) {
        const uint64_t a = first[grp], n = first[grp + 1] - a;
        const uint32_t r = dp[a].dpReadID & ~1u;
        std::vector<s3_sam_dp_pairing> al(n);
        std::vector<std::string> cig(n);
        bool any = false;
        for (uint64_t k = 0; k < n; ++k) {
            const s3_pe_dp_result &x = dp[a + k];
            const uint32_t dpSide = x.dpReadID & 1u, alignedID = x.dpReadID ^ 1u;
            s3_sam_dp_pairing &e = al[k];
            memset(&e, 0, sizeof e);
            uint32_t dpPos = 0xFFFFFFFFu;
            if (x.numRuns) {
                int32_t ed = 0, span = 0;
                const int rc = decode_runs(runs + x.runOffset, x.numRuns, reads->readLengths[x.dpReadID], x.score, scores, cig[k], &ed, &span);
                if (rc) return rc;
                dpPos = x.dpPos;
                e.whichFromDP = (uint8_t)dpSide; e.editdist = ed; e.numSameScore = (int32_t)x.numSameScore; e.cigar = cig[k].c_str();
                e.insertSize = dpPos < x.alignedPos ? (int32_t)(x.alignedPos - dpPos + reads->readLengths[alignedID])
                                                    : (int32_t)(dpPos - x.alignedPos + reads->readLengths[x.dpReadID]) + span;
                any = true;
            } else e.whichFromDP = 2;
            e.ambPosition[dpSide] = dpPos; e.strand[dpSide] = x.dpStrand; e.score[dpSide] = x.score;
            e.ambPosition[1 - dpSide] = x.alignedPos; e.strand[1 - dpSide] = x.alignedStrand; e.score[1 - dpSide] = x.alignedMismatches;
        }
        if (!any) return (int)S3_OK;
        int32_t x0[2], x1[2], mm[2];
        pair_counts(readStats, r, x0, x1, mm);
        s3_sam_record rec[2];
        const int rc = s3_sam_pair_dp_records(g, cfg, al.data(), (uint32_t)n, s3_sam_pick_pair_dp(al.data(), (uint32_t)n),
                                              reads->bases + (size_t)r * reads->rowBytes, reads->bases + (size_t)(r + 1) * reads->rowBytes,
                                              reads->qualities + (size_t)r * reads->rowBytes, reads->qualities + (size_t)(r + 1) * reads->rowBytes,
                                              (int32_t)reads->readLengths[r], (int32_t)reads->readLengths[r + 1], reads->names[r], reads->names[r + 1], x0, x1, mm, rec);
        return rc ? rc : append_pair(g, rec, out);
    });
}

// hostKernel's SAM branch for a pair the search paired (CPUfunctions.cpp:2281-2380 -> pairOutputSAMAPI) over the paired-end chain's result:
// the pairs with route S3_PE_PAIRED and ONE valid pairing -- the chain returns the reported pairing, the totals and (params.readStats) the
// per-read counts, which is all the writer takes then.  Pairs with more valid pairings need the whole list for XA:Z (s3_pair_occurrences)
// and are left to the caller, like every other route.
extern "C" int s3_sam_paired_batch_text(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_reads *reads, uint64_t numReads,
                                        const uint8_t *route, const s3_pe_pair_result *pairs, uint64_t numPairs, const s3_pe_read_stats *readStats,
                                        uint32_t numThreads, char **text, uint64_t *textBytes)
{
    if (text) *text = NULL;
    if (textBytes) *textBytes = 0;
    if (!g || !cfg || !reads_ok(reads) || (numPairs && (!route || !pairs || !readStats))) { s3_set_error("s3_sam_paired_batch_text: NULL argument (the chain's readStats are needed)"); return S3_EINVAL; }
    if (2 * numPairs > numReads) { s3_set_error("s3_sam_paired_batch_text: %llu pairs of %llu reads", (unsigned long long)numPairs, (unsigned long long)numReads); return S3_EINVAL; }
    std::vector<uint64_t> pick;
    for (uint64_t p = 0; p < numPairs; ++p) {
        if (route[p] != S3_PE_PAIRED || pairs[p].numPairs != 1) continue;
        for (uint64_t r = 2 * p; r < 2 * p + 2; ++r) if (reads->readLengths[r] == 0 || reads->readLengths[r] > reads->rowBytes) {
            s3_set_error("s3_sam_paired_batch_text: read %llu has length %u (rows of %u)", (unsigned long long)r, reads->readLengths[r], reads->rowBytes); return S3_EINVAL;
        }
        pick.push_back(p);
    }
    return batch_text("s3_sam_paired_batch_text", pick.size(), numThreads, text, textBytes, [&](uint64_t i, std::string &out) {
        const uint64_t p = pick[i], r = 2 * p;
        const s3_pe_pair_result &x = pairs[p];
        const s3_pe_read_stats &s1 = readStats[r], &s2 = readStats[r + 1];
        s3_sam_pairing pr;
        memset(&pr, 0, sizeof pr);
        pr.algnmt1 = x.pos1; pr.algnmt2 = x.pos2; pr.strand1 = x.strand1; pr.mismatch1 = x.mism1; pr.strand2 = x.strand2; pr.mismatch2 = x.mism2; pr.totalMismatchCount = x.optimalTotal;
        // X0 / X1 as hostKernel passes them per report type (CPUfunctions.cpp:2326-2362): the reads' counts for all-valid / all-best, 1 and none
        // for unique-best (one optimal pairing here), none for random-best
        int32_t x0[2] = {(int32_t)s1.x0, (int32_t)s2.x0}, x1[2] = {(int32_t)s1.x1, (int32_t)s2.x1};
        if (cfg->alignmentType == 3) { x0[0] = x0[1] = 1; x1[0] = x1[1] = -1; }
        else if (cfg->alignmentType == 4) { x0[0] = x0[1] = x1[0] = x1[1] = -1; }
        s3_sam_record rec[2];
        const int rc = s3_sam_pair_records(g, cfg, &pr, 1, 0, reads->bases + r * reads->rowBytes, reads->bases + (r + 1) * reads->rowBytes,
                                           reads->qualities + r * reads->rowBytes, reads->qualities + (r + 1) * reads->rowBytes,
                                           (int32_t)reads->readLengths[r], (int32_t)reads->readLengths[r + 1], reads->names[r], reads->names[r + 1],
                                           x.optimalTotal, x.suboptimalTotal, x0[0], x0[1], x1[0], x1[1], (int32_t)x.numOptimal,
                                           s1.minMismatch == x.mism1, s2.minMismatch == x.mism2, x.numPairs, rec);
        return rc ? rc : append_pair(g, rec, out);
    });
}

// hostKernel's SAM branch for a pair without a valid pairing (CPUfunctions.cpp:2546-2557 -> unproperlypairOutputSAMAPI, for runs without the DP stages): each read reported
// on its own from its occurrence list.  The lists are a CSR over all reads of the batch -- what s3_se_align returns for the same queries
// (the paired-end chain keeps its occurrences on the device) -- and pairIDs names the pairs to write (pair p = reads 2p, 2p + 1).
extern "C" int s3_sam_unpaired_batch_text(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_reads *reads, uint64_t numReads,
                                          const uint32_t *occOffsets, const uint32_t *positions, const uint8_t *occFlags,
                                          const uint32_t *pairIDs, uint64_t numPairs, uint32_t peMaxOutputPerRead, uint32_t numThreads,
                                          char **text, uint64_t *textBytes)
{
    if (text) *text = NULL;
    if (textBytes) *textBytes = 0;
    if (!g || !cfg || !reads_ok(reads) || !occOffsets || (numPairs && !pairIDs) || (numReads && occOffsets[numReads] && (!positions || !occFlags))) {
        s3_set_error("s3_sam_unpaired_batch_text: NULL argument"); return S3_EINVAL;
    }
    for (uint64_t i = 0; i < numPairs; ++i) {
        const uint64_t r = 2 * (uint64_t)pairIDs[i];
        if (r + 1 >= numReads) { s3_set_error("s3_sam_unpaired_batch_text: pair %u lies outside the batch", pairIDs[i]); return S3_EINVAL; }
        for (uint64_t k = r; k < r + 2; ++k) {
            if (occOffsets[k + 1] < occOffsets[k]) { s3_set_error("s3_sam_unpaired_batch_text: occOffsets decrease at read %llu", (unsigned long long)k); return S3_EINVAL; }
            if (reads->readLengths[k] == 0 || reads->readLengths[k] > reads->rowBytes) { s3_set_error("s3_sam_unpaired_batch_text: read %llu has length %u (rows of %u)", (unsigned long long)k, reads->readLengths[k], reads->rowBytes); return S3_EINVAL; }
        }
    }
    return batch_text("s3_sam_unpaired_batch_text", numPairs, numThreads, text, textBytes, [&](uint64_t i, std::string &out) {
        const uint64_t r = 2 * (uint64_t)pairIDs[i];
        std::vector<s3_sam_occurrence> occ[2];
        for (int k = 0; k < 2; ++k) {
            const uint32_t a = occOffsets[r + k], n = occOffsets[r + k + 1] - a;
            occ[k].resize(n);
            for (uint32_t j = 0; j < n; ++j) { occ[k][j].ambPosition = positions[a + j]; occ[k][j].strand = occFlags[2 * (size_t)(a + j)]; occ[k][j].mismatchCount = occFlags[2 * (size_t)(a + j) + 1]; occ[k][j].pad[0] = occ[k][j].pad[1] = 0; }
        }
        s3_sam_record rec[2];
        const int rc = s3_sam_unpaired_records(g, cfg, occ[0].data(), (uint32_t)occ[0].size(), occ[1].data(), (uint32_t)occ[1].size(), peMaxOutputPerRead,
                                               reads->bases + r * reads->rowBytes, reads->bases + (r + 1) * reads->rowBytes,
                                               reads->qualities + r * reads->rowBytes, reads->qualities + (r + 1) * reads->rowBytes,
                                               (int32_t)reads->readLengths[r], (int32_t)reads->readLengths[r + 1], reads->names[r], reads->names[r + 1], rec);
        return rc ? rc : append_pair(g, rec, out);
    });
}

// outputSingleResultForPairEnds (OutputDPResult.cpp:1062-1150) over what the stages left for the pairs that never became properly paired: per
// read ONE list (AllHits, PEAlgnmt.cpp:1033-1260) -- the hits of the single-read DP stage when the read has any (inputAlgnmtsToArray: isFromDP 1),
// else its occurrences from the search (inputSoap3AnsToArray: isFromDP 0, score = len x match + mismatches x mismatch score, CIGAR <len>M,
// edit distance = the mismatches), else nothing -- and unproperlypairDPOutputSAMAPI for the two lists.
extern "C" int s3_sam_unpaired_dp_batch_text(const s3_sam_genome *g, const s3_sam_config *cfg, const s3_sam_reads *reads, uint64_t numReads,
                                             const uint32_t *occOffsets, const uint32_t *positions, const uint8_t *occFlags,
                                             const s3_dp_hit *hits, uint64_t numHits, const uint32_t *runs, uint64_t numRuns, s3_dp_scores scores,
                                             int32_t singleDPcutoffThreshold, const uint32_t *pairIDs, uint64_t numPairs, uint32_t numThreads,
                                             char **text, uint64_t *textBytes)
{
    if (text) *text = NULL;
    if (textBytes) *textBytes = 0;
    if (!g || !cfg || !reads_ok(reads) || !occOffsets || (numPairs && !pairIDs) || (numHits && (!hits || !runs)) ||
        (numReads && occOffsets[numReads] && (!positions || !occFlags))) { s3_set_error("s3_sam_unpaired_dp_batch_text: NULL argument"); return S3_EINVAL; }
    // the hits of a read are next to each other (candidate order): first hit and count per read
    std::vector<uint64_t> firstHit(numReads, 0);
    std::vector<uint32_t> hitCount(numReads, 0);
    for (uint64_t i = 0; i < numHits; ++i) {
        const s3_dp_hit &h = hits[i];
        if (h.readID >= numReads || (uint64_t)h.runOffset + h.numRuns > numRuns) { s3_set_error("s3_sam_unpaired_dp_batch_text: hit %llu points outside the batch", (unsigned long long)i); return S3_EINVAL; }
        if (hitCount[h.readID] && hits[i - 1].readID != h.readID) { s3_set_error("s3_sam_unpaired_dp_batch_text: the hits of read %u are not next to each other", h.readID); return S3_EINVAL; }
        if (!hitCount[h.readID]) firstHit[h.readID] = i;
        ++hitCount[h.readID];
    }
    for (uint64_t i = 0; i < numPairs; ++i) {
        const uint64_t r = 2 * (uint64_t)pairIDs[i];
        if (r + 1 >= numReads) { s3_set_error("s3_sam_unpaired_dp_batch_text: pair %u lies outside the batch", pairIDs[i]); return S3_EINVAL; }
        for (uint64_t k = r; k < r + 2; ++k) {
            if (occOffsets[k + 1] < occOffsets[k]) { s3_set_error("s3_sam_unpaired_dp_batch_text: occOffsets decrease at read %llu", (unsigned long long)k); return S3_EINVAL; }
            if (reads->readLengths[k] == 0 || reads->readLengths[k] > reads->rowBytes) { s3_set_error("s3_sam_unpaired_dp_batch_text: read %llu has length %u (rows of %u)", (unsigned long long)k, reads->readLengths[k], reads->rowBytes); return S3_EINVAL; }
        }
    }
    return batch_text("s3_sam_unpaired_dp_batch_text", numPairs, numThreads, text, textBytes, [&](uint64_t i, std::string &out) {
        const uint64_t r = 2 * (uint64_t)pairIDs[i];
        std::vector<s3_sam_read_alignment> al[2];
        std::vector<std::string> cig[2];
        for (int k = 0; k < 2; ++k) {
            const uint64_t id = r + k;
            const uint32_t len = reads->readLengths[id];
            if (hitCount[id]) {
                al[k].resize(hitCount[id]); cig[k].resize(hitCount[id]);
                for (uint32_t j = 0; j < hitCount[id]; ++j) {
                    const s3_dp_hit &h = hits[firstHit[id] + j];
                    int32_t ed = 0, span = 0;
                    const int rc = decode_runs(runs + h.runOffset, h.numRuns, len, h.score, scores, cig[k][j], &ed, &span);
                    if (rc) return rc;
                    s3_sam_read_alignment &a = al[k][j];
                    memset(&a, 0, sizeof a);
                    a.ambPosition = h.pos; a.strand = h.strand; a.isFromDP = 1; a.score = h.score; a.editdist = ed; a.cigar = cig[k][j].c_str();
                }
            } else {
                const uint32_t a0 = occOffsets[id], n = occOffsets[id + 1] - a0;
                al[k].resize(n); cig[k].resize(n ? 1 : 0);
                if (n) { char nb[24]; cig[k][0].assign(nb, write_num(len, nb)); cig[k][0].push_back('M'); }
                for (uint32_t j = 0; j < n; ++j) {
                    s3_sam_read_alignment &a = al[k][j];
                    memset(&a, 0, sizeof a);
                    const int32_t mism = occFlags[2 * (size_t)(a0 + j) + 1];
                    a.ambPosition = positions[a0 + j]; a.strand = occFlags[2 * (size_t)(a0 + j)]; a.isFromDP = 0;
                    a.score = (int32_t)len * cfg->dpMatchScore + mism * cfg->dpMisMatchScore; a.editdist = mism; a.cigar = cig[k][0].c_str();
                }
            }
        }
        s3_sam_record rec[2];
        const int rc = s3_sam_unpaired_dp_records(g, cfg, al[0].data(), (uint32_t)al[0].size(), al[1].data(), (uint32_t)al[1].size(), singleDPcutoffThreshold,
                                                  reads->bases + r * reads->rowBytes, reads->bases + (r + 1) * reads->rowBytes,
                                                  reads->qualities + r * reads->rowBytes, reads->qualities + (r + 1) * reads->rowBytes,
                                                  (int32_t)reads->readLengths[r], (int32_t)reads->readLengths[r + 1], reads->names[r], reads->names[r + 1], rec);
        return rc ? rc : append_pair(g, rec, out);
    });
}
