// DP result decoding on the host: traceback pattern bytes -> CIGAR strings, edit distance, reference-span delta.
//
// Replaces the decode loops of the three DP engines' CPU threads (SingleDP_Space::algnmtCPUThread
// DV-DPfunctions.cu:1696-1730, DP_Space::algnmtCPUThread :2359-2400, DeepDP_Space::DP2CPUAlgnThread :3765-3795) with
// CigarStringEncoder (DV-DPfunctions.h:545-597), and convertToCigarStr (PE.cpp:420-485) for the SAM form.  It is host
// work in the reference and host work here: the inputs are the arrays s3_dp_align* already returned to host memory.
// Alignments are independent, so the batch is cut into contiguous chunks, one host thread each.
#include "s3_common.cuh"
#include "../../include/soap3dp_b200.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Run { uint8_t type; int32_t cnt; };

struct Decoded {
    int32_t editdist, refSpanDelta, gapPenalty;
    uint32_t ops[5];    // M m I D S
};

inline int op_slot(uint8_t t) { return t == 'M' ? 0 : t == 'm' ? 1 : t == 'I' ? 2 : t == 'D' ? 3 : t == 'S' ? 4 : -1; }

inline void put_num(std::string &s, int32_t v)
{
    char buf[16];
    int k = 0;
    uint32_t u = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    do { buf[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) s.push_back('-');
    while (k) s.push_back(buf[--k]);
}

// The pattern is written right-to-left by the traceback; 'V',c repeats the op before it c - 1 more times (c may be 0:
// the count then goes down by one, which is how a soft clip of size 0 disappears).  Adjacent entries of one type merge
// into a run exactly as CigarStringEncoder::append merges them (a run whose count ends <= 0 still separates its
// neighbours and is skipped on output, DV-DPfunctions.h:575).
void pattern_runs(const uint8_t *p, const uint8_t *end, std::vector<Run> &runs)
{
    runs.clear();
    uint8_t last = 'N';
    for (; p < end && *p != 0; ++p) {
        uint8_t type; int32_t cnt;
        if (*p == 'V') { if (++p >= end) break; type = last; cnt = (int32_t)*p - 1; }
        else {
            type = last = *p; cnt = 1;
            while (p + 1 < end && p[1] == type) { ++p; ++cnt; }     // a stretch of one op written byte by byte
        }
        if (!runs.empty() && runs.back().type == type) runs.back().cnt += cnt;
        else runs.push_back({type, cnt});
    }
}

// special CIGAR ('M' match, 'm' mismatch kept apart) in read order + the encoder's counts
void encode_special(const std::vector<Run> &runs, int32_t open, int32_t ext, std::string &out, Decoded &d)
{
    d.gapPenalty = 0;
    for (int k = 0; k < 5; ++k) d.ops[k] = 0;
    for (size_t i = runs.size(); i-- > 0;) {
        const Run &r = runs[i];
        if (r.cnt <= 0 || r.type == 'N') continue;      // the encoder's own sentinel runs are of type 'N' with count 0
        put_num(out, r.cnt);
        out.push_back((char)r.type);
        int s = op_slot(r.type);
        if (s >= 0) d.ops[s] += (uint32_t)r.cnt;
        if (r.type == 'I' || r.type == 'D') d.gapPenalty += open + (r.cnt - 1) * ext;
    }
}

// convertToCigarStr (PE.cpp:420-485): M and m fold into M; a deletion before the first aligned base or as the last op
// is dropped.  The reference leaves the dropped deletion's count in its accumulator, so the digits of the next count
// are appended to it ("3D5M" gives 35M); that is reproduced, since the SAM line is what must be identical.
void special_to_sam(const char *sp, size_t len, std::string &out)
{
    int32_t cur = 0, curM = 0;
    size_t start = out.size();
    for (size_t i = 0; i < len; ++i) {
        char c = sp[i];
        if (c >= '0' && c <= '9') { cur = cur * 10 + (c - '0'); continue; }
        switch (c) {
        case 'M': case 'm': curM += cur; cur = 0; break;
        case 'D':
            if ((out.size() == start && curM == 0) || i == len - 1) break;
            /* fall through */
        case 'I': case 'S':
            if (curM > 0) { put_num(out, curM); out.push_back('M'); curM = 0; }
            put_num(out, cur); out.push_back(c); cur = 0;
            break;
        default: break;
        }
    }
    if (curM > 0) { put_num(out, curM); out.push_back('M'); }
}

struct Chunk {
    std::string cigar, sam;
    std::vector<uint32_t> cigarLen, samLen;
};

}  // namespace

// The (op, count) runs the chains and the seeded stages return for an alignment (length << 8 | op, read order: what CigarStringEncoder
// holds) -> what the SAM writers take: the special CIGAR string, the edit distance and the insert-size term of the result loops
// (DV-DPfunctions.cu:1699-1733,2359-2400,3765-3795), computed as s3_dp_decode computes them from the pattern bytes.
extern "C" int s3_runs_decode(const uint32_t *runs, uint32_t numRuns, uint32_t readLength, int32_t score, s3_dp_scores sc, char *cigar, uint32_t cigarCapacity,
                              uint32_t *cigarLength, int32_t *editdist, int32_t *refSpanDelta)
{
    if ((numRuns && !runs) || !cigar || cigarCapacity == 0) { s3_set_error("s3_runs_decode: NULL argument"); return S3_EINVAL; }
    if (sc.matchScore == sc.mismatchScore) { s3_set_error("s3_runs_decode: match and mismatch scores are equal"); return S3_EINVAL; }
    std::string out;
    uint32_t ops[5] = {0, 0, 0, 0, 0};
    int32_t gapPenalty = 0;
    for (uint32_t i = 0; i < numRuns; ++i) {
        const int32_t cnt = (int32_t)(runs[i] >> 8);
        const uint8_t type = (uint8_t)(runs[i] & 0xFFu);
        const int slot = op_slot(type);
        if (cnt <= 0 || slot < 0) { s3_set_error("s3_runs_decode: run %u is not one of M m I D S with a positive length", i); return S3_EINVAL; }
        put_num(out, cnt);
        out.push_back((char)type);
        ops[slot] += (uint32_t)cnt;
        if (type == 'I' || type == 'D') gapPenalty += sc.gapOpenScore + (cnt - 1) * sc.gapExtendScore;
    }
    if (out.size() + 1 > cigarCapacity) { s3_set_error("s3_runs_decode: the CIGAR needs %zu bytes", out.size() + 1); return S3_EINVAL; }
    memcpy(cigar, out.c_str(), out.size() + 1);
    if (cigarLength) *cigarLength = (uint32_t)out.size();
    const int32_t L = (int32_t)readLength - (int32_t)ops[2] - (int32_t)ops[4];
    const int32_t mism = (L * sc.matchScore + gapPenalty - score) / (sc.matchScore - sc.mismatchScore);
    if (editdist) *editdist = (int32_t)ops[2] + (int32_t)ops[3] + mism;
    if (refSpanDelta) *refSpanDelta = (int32_t)ops[3] - (int32_t)ops[2] - (int32_t)ops[4];
    return S3_OK;
}

// convertToCigarStr for the SAM writers of s3_sam.cu
void s3_special_to_sam(const char *sp, size_t len, std::string &out) { special_to_sam(sp, len, out); }

extern "C" int s3_dp_decode(const uint8_t *pattern, uint32_t patternLength, const int32_t *scores, const uint32_t *readLengths,
                            const int32_t *cutoffThresholds, uint32_t numOfThreads, s3_dp_scores sc,
                            uint64_t *cigarOffsets, char **cigars, uint64_t *samOffsets, char **samCigars,
                            int32_t *editdist, int32_t *refSpanDelta, uint32_t *opCounts)
{
    if (!pattern || !scores || !readLengths || !cutoffThresholds || !cigarOffsets || !cigars || (samOffsets && !samCigars) || (!samOffsets && samCigars)) {
        s3_set_error("s3_dp_decode: NULL argument");
        return S3_EINVAL;
    }
    if (sc.matchScore == sc.mismatchScore) {
        s3_set_error("s3_dp_decode: matchScore == mismatchScore (the edit distance divides by their difference, DV-DPfunctions.cu:1723)");
        return S3_EINVAL;
    }
    *cigars = nullptr;
    if (samCigars) *samCigars = nullptr;
    const uint32_t n = numOfThreads;
    uint32_t nt = std::thread::hardware_concurrency();
    if (const char *e = getenv("S3_DECODE_THREADS")) nt = (uint32_t)atoi(e);
    nt = nt < 1 ? 1 : nt > 32 ? 32 : nt;
    if ((uint64_t)nt * 4096 > n) nt = n / 4096 ? n / 4096 : 1;
    std::vector<Chunk> chunks(nt);
    const bool wantSam = samOffsets != nullptr;
    std::atomic<bool> failed{false};
    auto work = [&](uint32_t c) { try {
        uint32_t lo = (uint32_t)((uint64_t)n * c / nt), hi = (uint32_t)((uint64_t)n * (c + 1) / nt);
        Chunk &ch = chunks[c];
        ch.cigarLen.resize(hi - lo);
        if (wantSam) ch.samLen.resize(hi - lo);
        ch.cigar.reserve((size_t)(hi - lo) * 24);
        if (wantSam) ch.sam.reserve((size_t)(hi - lo) * 12);
        std::vector<Run> runs;
        for (uint32_t t = lo; t < hi; ++t) {
            Decoded d{-1, 0, 0, {0, 0, 0, 0, 0}};
            size_t c0 = ch.cigar.size(), s0 = ch.sam.size();
            if (scores[t] >= cutoffThresholds[t]) {
                const uint8_t *p = pattern + (size_t)t * patternLength;
                pattern_runs(p, p + patternLength, runs);
                encode_special(runs, sc.gapOpenScore, sc.gapExtendScore, ch.cigar, d);
                int32_t L = (int32_t)readLengths[t] - (int32_t)d.ops[2] - (int32_t)d.ops[4];
                int32_t mism = (L * sc.matchScore + d.gapPenalty - scores[t]) / (sc.matchScore - sc.mismatchScore);
                d.editdist = (int32_t)d.ops[2] + (int32_t)d.ops[3] + mism;
                d.refSpanDelta = (int32_t)d.ops[3] - (int32_t)d.ops[2] - (int32_t)d.ops[4];
                if (wantSam) special_to_sam(ch.cigar.data() + c0, ch.cigar.size() - c0, ch.sam);
            }
            ch.cigarLen[t - lo] = (uint32_t)(ch.cigar.size() - c0);
            if (wantSam) ch.samLen[t - lo] = (uint32_t)(ch.sam.size() - s0);
            if (editdist) editdist[t] = d.editdist;
            if (refSpanDelta) refSpanDelta[t] = d.refSpanDelta;
            if (opCounts) for (int k = 0; k < 5; ++k) opCounts[(size_t)t * 5 + k] = d.ops[k];
        }
    } catch (...) { failed = true; } };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (uint32_t c = 0; c < nt; ++c) th.emplace_back(work, c);
        for (auto &t : th) t.join();
    }
    if (failed) { s3_set_error("s3_dp_decode: out of host memory"); return S3_ENOMEM; }
    uint64_t total = 0, totalSam = 0;
    for (auto &ch : chunks) { total += ch.cigar.size(); totalSam += ch.sam.size(); }
    char *out = (char *)malloc(total + 1), *outSam = wantSam ? (char *)malloc(totalSam + 1) : nullptr;
    if (!out || (wantSam && !outSam)) { free(out); free(outSam); s3_set_error("out of host memory"); return S3_ENOMEM; }
    uint64_t at = 0, atSam = 0, pos = 0, posSam = 0;
    uint32_t t = 0;
    for (auto &ch : chunks) {
        memcpy(out + pos, ch.cigar.data(), ch.cigar.size()); pos += ch.cigar.size();
        if (wantSam) { memcpy(outSam + posSam, ch.sam.data(), ch.sam.size()); posSam += ch.sam.size(); }
        for (size_t k = 0; k < ch.cigarLen.size(); ++k, ++t) {
            cigarOffsets[t] = at; at += ch.cigarLen[k];
            if (wantSam) { samOffsets[t] = atSam; atSam += ch.samLen[k]; }
        }
    }
    cigarOffsets[n] = at; out[total] = 0;
    if (wantSam) { samOffsets[n] = atSam; outSam[totalSam] = 0; *samCigars = outSam; }
    *cigars = out;
    return S3_OK;
}


// MD:Z string and the NM pieces of an alignment from its special CIGAR and the packed text.  Replaces getMisInfoForDP
// (PE.cpp:499-666) with trim = 0 (the strand only matters to its trimming): matches accumulate; a mismatch run writes
// the count so far and the text base, further bases of the run as "0" + base; a deletion writes the count, '^' and the
// deleted text bases -- except as the last op, which is ignored; the final count closes the string.
namespace {

inline char text_base(const uint32_t *packed, uint64_t pos)
{
    static const char dna[4] = {'A', 'C', 'G', 'T'};               // dnaChar, 2bwt-lib/HSP.h:278
    return dna[(packed[pos >> 4] >> ((15 - (pos & 15)) * 2)) & 3];
}

struct MdOut { int32_t numMismatch, gapOpen, gapExt, avgQual; };

// returns false when the alignment runs past the text
bool md_one(const uint32_t *packed, uint64_t textLength, const char *cig, size_t len, uint32_t pos, const int8_t *qual, size_t qualLen,
            std::string &md, MdOut &o)
{
    o = MdOut{0, 0, 0, 20};                                          // DEFAULT_QUAL_VALUE, PE.h:27
    int32_t cur = 0, curMatch = 0;
    uint64_t qPos = 0, tPos = pos;
    double sumQual = 0.0;
    for (size_t i = 0; i < len; ++i) {
        const char c = cig[i];
        if (c >= '0' && c <= '9') { cur = cur * 10 + (c - '0'); continue; }
        switch (c) {
        case 'M': curMatch += cur; qPos += cur; tPos += cur; cur = 0; break;
        case 'm':
            if (tPos + (uint64_t)cur > textLength) return false;
            put_num(md, curMatch);
            md.push_back(text_base(packed, tPos));
            if (qual && qPos < qualLen) sumQual += qual[qPos];
            for (int32_t j = 1; j < cur; ++j) {
                md.push_back('0');
                md.push_back(text_base(packed, tPos + j));
                if (qual && qPos + j < qualLen) sumQual += qual[qPos + j];
            }
            qPos += cur; tPos += cur; o.numMismatch += cur; curMatch = 0; cur = 0;
            break;
        case 'I': qPos += cur; o.gapOpen++; o.gapExt += cur; cur = 0; break;
        case 'D':
            if (i == len - 1) break;                                 // last delete, ignored (its count stays, like the reference)
            if (tPos + (uint64_t)cur > textLength) return false;
            put_num(md, curMatch);
            md.push_back('^');
            for (int32_t j = 0; j < cur; ++j) md.push_back(text_base(packed, tPos + j));
            tPos += cur; o.gapOpen++; o.gapExt += cur; curMatch = 0; cur = 0;
            break;
        case 'S': qPos += cur; cur = 0; break;
        default: break;
        }
    }
    put_num(md, curMatch);
    if (o.numMismatch > 0) o.avgQual = (int32_t)(sumQual / o.numMismatch);
    return true;
}

}  // namespace

extern "C" int s3_dp_md(const uint32_t *packedDNA, uint64_t textLength, const char *cigars, const uint64_t *cigarOffsets,
                        const uint32_t *positions, uint32_t numOfThreads, const int8_t *qualities, const uint64_t *qualityOffsets,
                        uint64_t *mdOffsets, char **md, int32_t *numMismatch, int32_t *gapOpen, int32_t *gapExt, int32_t *avgMismatchQual)
{
    if (!packedDNA || !cigarOffsets || !positions || !mdOffsets || !md || (qualities && !qualityOffsets) || (numOfThreads && cigarOffsets[numOfThreads] && !cigars)) {
        s3_set_error("s3_dp_md: NULL argument");
        return S3_EINVAL;
    }
    *md = nullptr;
    const uint32_t n = numOfThreads;
    uint32_t nt = std::thread::hardware_concurrency();
    if (const char *e = getenv("S3_DECODE_THREADS")) nt = (uint32_t)atoi(e);
    nt = nt < 1 ? 1 : nt > 32 ? 32 : nt;
    if ((uint64_t)nt * 4096 > n) nt = n / 4096 ? n / 4096 : 1;
    struct MdChunk { std::string md; std::vector<uint32_t> len; };
    std::vector<MdChunk> chunks(nt);
    std::atomic<bool> failed{false};
    std::atomic<int64_t> pastText{-1};
    auto work = [&](uint32_t c) { try {
        const uint32_t lo = (uint32_t)((uint64_t)n * c / nt), hi = (uint32_t)((uint64_t)n * (c + 1) / nt);
        MdChunk &ch = chunks[c];
        ch.len.resize(hi - lo);
        ch.md.reserve((size_t)(hi - lo) * 12);
        for (uint32_t t = lo; t < hi; ++t) {
            const size_t m0 = ch.md.size();
            MdOut o{0, 0, 0, 20};
            const size_t len = (size_t)(cigarOffsets[t + 1] - cigarOffsets[t]);
            if (len) {                                               // an alignment under its cutoff has no CIGAR and gets no MD
                const int8_t *q = qualities ? qualities + qualityOffsets[t] : nullptr;
                const size_t ql = qualities ? (size_t)(qualityOffsets[t + 1] - qualityOffsets[t]) : 0;
                if (!md_one(packedDNA, textLength, cigars + cigarOffsets[t], len, positions[t], q, ql, ch.md, o)) {
                    int64_t none = -1;
                    pastText.compare_exchange_strong(none, (int64_t)t);
                    return;
                }
            }
            ch.len[t - lo] = (uint32_t)(ch.md.size() - m0);
            if (numMismatch) numMismatch[t] = o.numMismatch;
            if (gapOpen) gapOpen[t] = o.gapOpen;
            if (gapExt) gapExt[t] = o.gapExt;
            if (avgMismatchQual) avgMismatchQual[t] = o.avgQual;
        }
    } catch (...) { failed = true; } };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (uint32_t c = 0; c < nt; ++c) th.emplace_back(work, c);
        for (auto &t : th) t.join();
    }
    if (failed) { s3_set_error("s3_dp_md: out of host memory"); return S3_ENOMEM; }
    if (pastText >= 0) {
        const uint32_t t = (uint32_t)pastText.load();
        s3_set_error("s3_dp_md: alignment %u at %u runs past the text (%llu bases)", t, positions[t], (unsigned long long)textLength);
        return S3_EINVAL;
    }
    uint64_t total = 0;
    for (auto &ch : chunks) total += ch.md.size();
    char *buf = (char *)malloc(total + 1);
    if (!buf) { s3_set_error("s3_dp_md: out of host memory"); return S3_ENOMEM; }
    uint64_t at = 0, posBytes = 0;
    uint32_t t = 0;
    for (auto &ch : chunks) {
        memcpy(buf + posBytes, ch.md.data(), ch.md.size()); posBytes += ch.md.size();
        for (size_t k = 0; k < ch.len.size(); ++k, ++t) { mdOffsets[t] = at; at += ch.len[k]; }
    }
    mdOffsets[n] = at;
    buf[total] = 0;
    *md = buf;
    return S3_OK;
}
