"""End to end: reads -> the CUDA path (chain / seeded stage through the C ABI) -> s3_runs_decode -> the SAM record entries, against
reads -> the composition of the oracles -> the reference's own SAM writers (oracle/_ref/libref_sam.so).  Every piece has its own
parity test; this one checks that what the device entries return is what the record writers need -- bam1_t records equal byte for byte."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch

import helpers
from helpers import HostIndex, ROOT, fmindex, formats
from soap3dp_b200 import api, synth

sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import seeding_oracle  # noqa: E402
from test_cpu_sam import (Config, DeepAlignment, DpAlignment, Genome, I32P, Occurrence, Record, REF, Segment, U8P)  # noqa: E402
from test_stages_gpu import OracleEnv, PAR, mutate  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_sam.so not built")]
SCORES = (1, -2, -3, -1)


@pytest.fixture(scope="module")
def env():
    G = synth.random_genome(400_000, seed=29)
    idx = fmindex.build_index(G, keep_sa=True)
    gi = api.GPUINDEXUpload(idx, device=0, with_text=True, with_sa=True)
    yield G, idx, HostIndex(idx), gi
    api.GPUINDEXFree(gi)


class OneChromosome:
    """the genome description of the SAM entries for a text that is one chromosome (hsp->translate, ambiguityMap, seqOffset)"""

    def __init__(self, G):
        g = G.cpu().numpy().astype(np.uint8)
        self.n = len(g)
        self.pac = helpers.pack_text(g)
        self.translate = np.array([0, 1, 0xFFFFFFFF], np.uint32)            # position p -> 1-based p + 1 on chromosome 1
        self.chr_end = np.array([self.n - 1], np.uint32)
        self.amb = np.zeros((self.n >> 18) + 2, np.uint32)
        self.names = [b"chrSynth"]
        self.segs = (Segment * 1)(Segment(0, 1, 0xFFFFFFFF))
        self.cnames = (C.c_char_p * 1)(*self.names)
        self.gen = Genome(helpers.u32p(self.pac), self.n, self.segs, 1, helpers.u32p(self.amb), helpers.u32p(self.chr_end), 1, self.cnames)

    def ref_args(self):
        return (helpers.u32p(self.pac), self.n, helpers.u32p(self.translate), 1, helpers.u32p(self.amb), helpers.u32p(self.chr_end), 1, self.cnames)


def record_tuple(r):
    return ((r.tid, r.pos, r.bin, r.qual, r.l_qname, r.flag, r.n_cigar, r.l_qseq, r.mtid, r.mpos, r.isize, r.l_aux), bytes(bytearray(r.data[:r.data_len])))


def runs_decode(lib, runs, read_length, score):
    buf = C.create_string_buffer(2048)
    e, s = C.c_int32(), C.c_int32()
    assert lib.s3_runs_decode(np.ascontiguousarray(runs, np.uint32).ctypes.data_as(C.POINTER(C.c_uint32)), len(runs), read_length, score, api.DPScores(*SCORES), buf, 2048,
                              None, C.byref(e), C.byref(s)) == 0
    return buf.value, e.value, s.value


def line_of(lib, one, rec):
    """the record's SAM text line (s3_sam_format_line); the record stays the caller's"""
    lib.s3_sam_format_line.restype = C.c_int
    lib.s3_sam_format_line.argtypes = [C.POINTER(Record), C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_void_p)]
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    line = C.c_void_p()
    assert lib.s3_sam_format_line(C.byref(rec), one.cnames, 1, C.byref(line)) == 0
    text = C.string_at(line.value)
    lib.s3_free(line)
    return text


def batch_reads(reads, quals, lens, names):
    """s3_sam_reads over rows of equal width"""
    from test_cpu_sam import SamReads
    keep = (np.ascontiguousarray(reads, np.uint8), np.ascontiguousarray(quals, np.uint8), np.ascontiguousarray(lens, np.uint32), (C.c_char_p * len(names))(*names))
    assert keep[0].shape == keep[1].shape
    return SamReads(keep[0].ctypes.data_as(U8P), C.cast(keep[1].ctypes.data, C.c_char_p), keep[0].shape[1], keep[2].ctypes.data_as(C.POINTER(C.c_uint32)), keep[3]), keep


def oracle_edit(cigar, read_length, score):
    """edit distance and D - I - S of the reference's result loops (DV-DPfunctions.cu:3788-3794), from the oracle's special CIGAR"""
    ops = {k: 0 for k in "MmIDS"}
    gap = 0
    for k, op in re.findall(r"(\d+)([MmIDS])", cigar):
        ops[op] += int(k)
        if op in "ID":
            gap += SCORES[2] + (int(k) - 1) * SCORES[3]
    L = read_length - ops["I"] - ops["S"]
    mism = int((L * SCORES[0] + gap - score) / (SCORES[0] - SCORES[1]))
    return ops["I"] + ops["D"] + mism, ops["D"] - ops["I"] - ops["S"]


def test_single_end_reads_to_sam_records(env):
    """s3_se_align -> s3_sam_single_record == the oracle chain's occurrences -> OCCOutputSAMAPI"""
    G, idx, hi, gi = env
    ref, lib = C.CDLL(REF), api.load_library()
    ref.ref_sam_single.restype = C.c_int
    lib.s3_sam_single_record.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    one = OneChromosome(G)
    n, L, k = 1500, 100, 2
    rs = synth.simulate_single_end(G, n, L, seed=41, sub_rate=0.015)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    wpq = formats.word_per_query(L)
    reads = rs.reads.numpy().astype(np.uint8)
    q = formats.pack_queries(reads, lens[:n], wpq)
    al = api.SingleAligner(gi, n, num_mismatch=k)
    got = al.align(q, lens, n, wpq)
    al.free()
    # the oracle chain: the search oracle's slots -> collect_all_answers -> transferAllSAToOcc (oracle/pe_chain_oracle.py)
    import pe_chain_oracle
    olib = helpers.load_oracle()
    allowed = formats.SA_RANGES_ROUND1[k]
    wpa = 2 * allowed
    bad = np.zeros(formats.ceil32(n), np.uint8)
    views = []
    for case in range(formats.NUM_CASES[k]):
        a_ = np.zeros(formats.ceil32(n) * wpa, np.uint32)
        helpers.oracle_launch(olib, hi, case, q, lens, n, wpq, a_, bad, 0, k, allowed, wpa)
        views.append(formats.answers_view(a_, n, wpa))
    sa = idx.fwd.sa.cpu().numpy()
    want_occ = [[(int(sa[i]), st, mm) for l, rr, st, mm in ranges for i in range(l, rr + 1)] for ranges, tot, more in pe_chain_oracle.collect(views, allowed, hi.n, 1000)]
    cfg = Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgE2E")
    rng = np.random.default_rng(5)
    mapped = 0
    quals = np.ascontiguousarray(rng.integers(2, 41, (n, L + 1)).astype(np.uint8)); quals[:, -1] = 0
    lines = []
    for r in range(n):
        a, b = int(got["occ_offsets"][r]), int(got["occ_offsets"][r + 1])
        occ = [(int(got["positions"][i]), int(got["occ_flags"][i][0]), int(got["occ_flags"][i][1])) for i in range(a, b)]
        theirs = want_occ[r]
        ql = quals[r]
        name = b"se%d" % r
        qr = np.ascontiguousarray(reads[r])
        arr = (Occurrence * max(len(occ), 1))(*[Occurrence(*o) for o in occ])
        out = Record()
        assert lib.s3_sam_single_record(C.byref(one.gen), C.byref(cfg), arr, len(occ), qr.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name, C.byref(out)) == 0
        mine = record_tuple(out)
        lines.append(line_of(lib, one, out) + b"\n")
        lib.s3_sam_record_free(C.byref(out))
        flat = np.array([x for o in theirs for x in o], np.uint32) if theirs else np.zeros(3, np.uint32)
        core, data, dlen = np.zeros(12, np.int32), np.zeros(8192, np.uint8), np.zeros(1, np.int32)
        assert ref.ref_sam_single(*one.ref_args(), cfg.alignmentType, cfg.bwaLikeScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup,
                                  helpers.u32p(flat), len(theirs), qr.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name,
                                  core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), 8192, dlen.ctypes.data_as(I32P)) == 1
        assert mine == (tuple(int(x) for x in core), bytes(data[:int(dlen[0])])), r
        mapped += bool(occ)
    assert mapped > n // 2
    # the whole batch as SAM text (s3_sam_single_batch_text over the chain's arrays) == those records' lines in read order
    padded = np.zeros((n, L + 1), np.uint8); padded[:, :L] = reads
    rd, keep = batch_reads(padded, quals, lens[:n], [b"se%d" % r for r in range(n)])
    lib.s3_sam_single_batch_text.restype = C.c_int
    text, size = C.c_void_p(), C.c_uint64()
    off, pos, fl = (np.ascontiguousarray(got["occ_offsets"], np.uint32), np.ascontiguousarray(got["positions"], np.uint32), np.ascontiguousarray(got["occ_flags"], np.uint8))
    assert lib.s3_sam_single_batch_text(C.byref(one.gen), C.byref(cfg), C.byref(rd), C.c_uint64(n), helpers.u32p(off), helpers.u32p(pos), fl.ctypes.data_as(U8P), 4,
                                        C.byref(text), C.byref(size)) == 0, lib.s3_last_error()
    assert C.string_at(text.value, size.value) == b"".join(lines)
    lib.s3_free(text)


def test_read_pairs_through_deep_dp_to_sam_records(env):
    """s3_deep_dp_align -> s3_runs_decode -> s3_sam_deep_dp_records == oracle/seeding_oracle.deep_dp -> pairDeepDPOutputSAMAPI, the
    reported entry = the first one with the largest sum of scores (outputDeepDPResult2, OutputDPResult.cpp:700-760)"""
    G, idx, hi, gi = env
    ref, lib = C.CDLL(REF), api.load_library()
    ref.ref_sam_deep_dp.restype = C.c_int
    lib.s3_sam_deep_dp_records.restype = C.c_int
    lib.s3_sam_pick_deep_dp.restype = C.c_int32
    lib.s3_runs_decode.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    one = OneChromosome(G)
    rng = np.random.default_rng(13)
    L, pairs = 100, 240
    m1, m2, _ = synth.simulate_paired_end(G, pairs, L, seed=17, bad_mate_fraction=0.0)
    raw = torch.stack([m1.reads, m2.reads], dim=1).reshape(2 * pairs, L).cpu().numpy()
    reads = [mutate(rng, r, int(rng.integers(3, 8)), int(rng.integers(0, 2))) for r in raw]
    n = 2 * pairs
    wpq = formats.word_per_query(L)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    q = formats.pack_queries(np.stack(reads), lens[:n], wpq)
    ids = 2 * np.arange(pairs, dtype=np.uint32)
    got = api.deep_dp_align(gi, q, lens, n, wpq, ids, api.stage_params())
    want = seeding_oracle.deep_dp(OracleEnv(idx, hi), G.cpu().numpy(), reads, ids.tolist(), PAR)
    assert len(got["hits"]) == len(want["hits"]) > pairs // 3
    by_pair_got, by_pair_want = {}, {}
    for h in got["hits"]:
        by_pair_got.setdefault(int(h["readID"]), []).append(h)
    for w in want["hits"]:
        by_pair_want.setdefault(int(w[0]), []).append(w)
    assert sorted(by_pair_got) == sorted(by_pair_want)
    cfg = Config(1, 0, SCORES[0], SCORES[1], 1, 40, 1, 1, 1, 1000, b"rgDeep")
    zeros = (C.c_int32 * 2)(0, 0)
    counts = np.zeros(6, np.int32)
    quals = np.ascontiguousarray(rng.integers(2, 41, (n, L + 1)).astype(np.uint8)); quals[:, -1] = 0
    lines = []
    for e in sorted(by_pair_got):
        q1, q2 = np.ascontiguousarray(reads[e]).astype(np.uint8), np.ascontiguousarray(reads[e + 1]).astype(np.uint8)
        ql1, ql2 = quals[e], quals[e + 1]
        n1, n2 = b"p%d/1" % e, b"p%d/2" % e
        # ---- the CUDA path's hits -> the writer's inputs
        hs = by_pair_got[e]
        arr = (DeepAlignment * len(hs))()
        keep = []
        for k, h in enumerate(hs):
            d1 = runs_decode(lib, got["runs"][int(h["runOffset1"]):int(h["runOffset1"]) + int(h["numRuns1"])], L, int(h["score1"]))
            d2 = runs_decode(lib, got["runs"][int(h["runOffset2"]):int(h["runOffset2"]) + int(h["numRuns2"])], L, int(h["score2"]))
            keep += [d1[0], d2[0]]
            p1, p2 = int(h["pos1"]), int(h["pos2"])
            arr[k].insertSize = (p2 - p1 + L + d2[2]) if p1 < p2 else (p1 - p2 + L + d1[2])              # DV-DPfunctions.cu:3810-3815
            arr[k].ambPosition[0], arr[k].ambPosition[1] = p1, p2
            arr[k].strand[0], arr[k].strand[1] = int(h["strand1"]), int(h["strand2"])
            arr[k].score[0], arr[k].score[1] = int(h["score1"]), int(h["score2"])
            arr[k].editdist[0], arr[k].editdist[1] = d1[1], d2[1]
            arr[k].numSameScore[0], arr[k].numSameScore[1] = int(h["numSame1"]), int(h["numSame2"])
            arr[k].cigar[0], arr[k].cigar[1] = d1[0], d2[0]
        best = lib.s3_sam_pick_deep_dp(arr, len(hs))                # outputDeepDPResult2's choice
        out = (Record * 2)()
        assert lib.s3_sam_deep_dp_records(C.byref(one.gen), C.byref(cfg), arr, len(hs), best, q1.ctypes.data_as(U8P), q2.ctypes.data_as(U8P), ql1.ctypes.data_as(C.c_char_p),
                                          ql2.ctypes.data_as(C.c_char_p), L, L, n1, n2, zeros, zeros, zeros, out) == 0
        mine = [record_tuple(r) for r in out]
        lines += [line_of(lib, one, out[0]) + b"\n", line_of(lib, one, out[1]) + b"\n"]
        for k in range(2):
            lib.s3_sam_record_free(C.byref(out[k]))
        # ---- the oracle's hits -> the reference's writer
        ws = by_pair_want[e]
        flat, cigs = [], []
        for w in ws:
            _, s1, s2, p1, p2, sc1, sc2, ns1, ns2, c1, c2 = w
            e1, dis1 = oracle_edit(c1, L, sc1)
            e2, dis2 = oracle_edit(c2, L, sc2)
            ins = (p2 - p1 + L + dis2) if p1 < p2 else (p1 - p2 + L + dis1)
            flat += [ins, p1, s1, sc1, e1, ns1, len(cigs), p2, s2, sc2, e2, ns2, len(cigs) + 1]
            cigs += [c1.encode(), c2.encode()]
        wsums = [w[5] + w[6] for w in ws]
        wbest = wsums.index(max(wsums))
        flat = np.array(flat, np.int64).astype(np.int32)
        cig = (C.c_char_p * len(cigs))(*cigs)
        core, data, dlen = np.zeros(24, np.int32), np.zeros(2 * 8192, np.uint8), np.zeros(2, np.int32)
        assert ref.ref_sam_deep_dp(*one.ref_args(), cfg.alignmentType, cfg.bwaLikeScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup,
                                   cfg.dpMatchScore, cfg.dpMisMatchScore, flat.ctypes.data_as(I32P), len(ws), wbest, cig, counts.ctypes.data_as(I32P),
                                   q1.ctypes.data_as(U8P), q2.ctypes.data_as(U8P), ql1.ctypes.data_as(C.c_char_p), ql2.ctypes.data_as(C.c_char_p), L, L, n1, n2,
                                   core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), 8192, dlen.ctypes.data_as(I32P)) == 2
        theirs = [(tuple(int(x) for x in core[12 * r:12 * r + 12]), bytes(data[r * 8192:r * 8192 + int(dlen[r])])) for r in range(2)]
        assert mine == theirs, (e, mine, theirs)
    # the stage's whole result as SAM text (s3_sam_deep_dp_batch_text over hits + runs) == those records' lines, pairs in hit order
    padded = np.zeros((n, L + 1), np.uint8); padded[:, :L] = np.stack(reads)
    rd, keep = batch_reads(padded, quals, lens[:n], [b"p%d/%d" % (r & ~1, 1 + (r & 1)) for r in range(n)])
    lib.s3_sam_deep_dp_batch_text.restype = C.c_int
    text, size = C.c_void_p(), C.c_uint64()
    hits, runs = np.ascontiguousarray(got["hits"]), np.ascontiguousarray(got["runs"], np.uint32)
    assert lib.s3_sam_deep_dp_batch_text(C.byref(one.gen), C.byref(cfg), C.byref(rd), C.c_uint64(n), hits.ctypes.data_as(C.c_void_p), C.c_uint64(len(hits)), helpers.u32p(runs),
                                         C.c_uint64(len(runs)), api.DPScores(*SCORES), None, 4, C.byref(text), C.byref(size)) == 0, lib.s3_last_error()
    assert C.string_at(text.value, size.value) == b"".join(lines) and len(lines) == 2 * len(by_pair_got)
    lib.s3_free(text)


def test_single_reads_through_single_dp_to_sam_records(env):
    """s3_single_dp_align -> s3_runs_decode -> s3_sam_single_dp_record == oracle/seeding_oracle.single_dp -> SingleDPOutputSAMAPI (150-base reads
    with indels, every alignment of a read handed to the writer, which picks the one it reports)"""
    G, idx, hi, gi = env
    ref, lib = C.CDLL(REF), api.load_library()
    ref.ref_sam_single_dp.restype = C.c_int
    lib.s3_sam_single_dp_record.restype = C.c_int
    lib.s3_runs_decode.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    one = OneChromosome(G)
    rng = np.random.default_rng(19)
    L, n = 150, 400
    rs = synth.simulate_single_end(G, n, L, seed=43, sub_rate=0.0)
    reads = [mutate(rng, r, int(rng.integers(3, 9)), int(rng.integers(0, 3))) for r in rs.reads.numpy()]
    wpq = formats.word_per_query(L)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    q = formats.pack_queries(np.stack(reads), lens[:n], wpq)
    ids = np.arange(n, dtype=np.uint32)
    got = api.single_dp_align(gi, q, lens, n, wpq, ids, api.stage_params())
    want = seeding_oracle.single_dp(OracleEnv(idx, hi), G.cpu().numpy(), reads, ids.tolist(), PAR)
    assert len(got["hits"]) == len(want["hits"]) > n // 3
    mine_by, want_by = {}, {}
    for h in got["hits"]:
        mine_by.setdefault(int(h["readID"]), []).append(h)
    for w in want["hits"]:
        want_by.setdefault(int(w[0]), []).append(w)
    assert sorted(mine_by) == sorted(want_by)
    cfg = Config(1, 0, SCORES[0], SCORES[1], 1, 40, 1, 1, 1, 1000, b"rgSDP")
    cutoff = int(np.ceil(0.3 * L))
    quals = np.ascontiguousarray(rng.integers(2, 41, (n, L + 1)).astype(np.uint8)); quals[:, -1] = 0
    lines = []
    for r in sorted(mine_by):
        qr = np.ascontiguousarray(reads[r]).astype(np.uint8)
        ql = quals[r]
        name = b"sdp%d" % r
        hs = mine_by[r]
        arr = (DpAlignment * len(hs))()
        keep = []
        for k, h in enumerate(hs):
            d = runs_decode(lib, got["runs"][int(h["runOffset"]):int(h["runOffset"]) + int(h["numRuns"])], L, int(h["score"]))
            keep.append(d[0])
            arr[k].ambPosition, arr[k].strand, arr[k].score, arr[k].editdist, arr[k].cigar = int(h["pos"]), int(h["strand"]), int(h["score"]), d[1], d[0]
        out = Record()
        assert lib.s3_sam_single_dp_record(C.byref(one.gen), C.byref(cfg), arr, len(hs), cutoff, qr.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name, C.byref(out)) == 0
        mine = record_tuple(out)
        lines.append(line_of(lib, one, out) + b"\n")
        lib.s3_sam_record_free(C.byref(out))
        ws = want_by[r]
        flat, cigs = [], []
        for k, w in enumerate(ws):
            _, st, pos, sc, same, cg = w
            flat += [pos, st, sc, oracle_edit(cg, L, sc)[0], k]
            cigs.append(cg.encode())
        flat = np.array(flat, np.int64).astype(np.int32)
        cig = (C.c_char_p * len(cigs))(*cigs)
        core, data, dlen = np.zeros(12, np.int32), np.zeros(8192, np.uint8), np.zeros(1, np.int32)
        assert ref.ref_sam_single_dp(*one.ref_args(), cfg.alignmentType, cfg.bwaLikeScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup,
                                     cfg.dpMatchScore, cutoff, flat.ctypes.data_as(I32P), len(ws), cig, qr.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name,
                                     core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), 8192, dlen.ctypes.data_as(I32P)) == 1
        assert mine == (tuple(int(x) for x in core), bytes(data[:int(dlen[0])])), r
    # the stage's whole result as SAM text (s3_sam_single_dp_batch_text over hits + runs) == those records' lines, reads in hit order
    padded = np.zeros((n, L + 1), np.uint8); padded[:, :L] = np.stack(reads)
    rd, keep = batch_reads(padded, quals, lens[:n], [b"sdp%d" % r for r in range(n)])
    lib.s3_sam_single_dp_batch_text.restype = C.c_int
    text, size = C.c_void_p(), C.c_uint64()
    hits, runs = np.ascontiguousarray(got["hits"]), np.ascontiguousarray(got["runs"], np.uint32)
    assert lib.s3_sam_single_dp_batch_text(C.byref(one.gen), C.byref(cfg), C.byref(rd), C.c_uint64(n), hits.ctypes.data_as(C.c_void_p), C.c_uint64(len(hits)), helpers.u32p(runs),
                                           C.c_uint64(len(runs)), api.DPScores(*SCORES), cutoff, 4, C.byref(text), C.byref(size)) == 0, lib.s3_last_error()
    assert C.string_at(text.value, size.value) == b"".join(lines)
    lib.s3_free(text)


def test_read_pairs_through_the_chain_to_sam_records(env):
    """s3_pe_align (with params.readStats) -> s3_sam_pair_records == the oracle chain + the restated per-read statistics -> pairOutputSAMAPI,
    for the pairs the chain paired with ONE valid pairing (the chain returns the reported pairing and the counts; with more pairings
    the XA:Z list needs them all: s3_pair_occurrences)"""
    import test_pe_chain_gpu as chain
    from test_cpu_sam import Pairing
    G, idx, hi, gi = env
    ref, lib = C.CDLL(REF), api.load_library()
    ref.ref_sam_pair.restype = C.c_int
    lib.s3_sam_pair_records.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    one = OneChromosome(G)
    L, pairs = 100, 700
    got, want = chain._run_both(env, pairs, L, 51, read_stats=True)
    chain._compare(got, want)
    reads = chain.LAST["reads"]
    ostats = chain.oracle_read_stats(chain.LAST["views"], chain.LAST["allowed"], chain.LAST["text_length"], chain.LAST["max_output"])
    rng = np.random.default_rng(23)
    done = 0
    quals = np.ascontiguousarray(rng.integers(2, 41, (2 * pairs, L + 1)).astype(np.uint8)); quals[:, -1] = 0
    padded = np.zeros((2 * pairs, L + 1), np.uint8); padded[:, :L] = np.asarray(reads)[:2 * pairs]
    rd, keep = batch_reads(padded, quals, np.full(2 * pairs, L, np.uint32), [b"c%d/%d" % (r >> 1, 1 + (r & 1)) for r in range(2 * pairs)])
    lib.s3_sam_paired_batch_text.restype = C.c_int
    for mode in ((1, 0), (2, 1)):                                     # (report type, BWA-like MAPQ)
        cfg = Config(mode[0], mode[1], SCORES[0], SCORES[1], 1, 40, 1, 1, 1, 1000, b"rgPE")
        lines = []
        for p in range(pairs):
            g = got["pairs"][p]
            if int(got["route"][p]) != 1 or int(g["numPairs"]) != 1:
                continue
            w = want["pairs"][p]
            q1, q2 = np.ascontiguousarray(reads[2 * p]).astype(np.uint8), np.ascontiguousarray(reads[2 * p + 1]).astype(np.uint8)
            ql1, ql2 = quals[2 * p], quals[2 * p + 1]
            n1, n2 = b"c%d/1" % p, b"c%d/2" % p
            # ---- the chain's result of the pair -> the writer's inputs
            s1, s2 = got["read_stats"][2 * p], got["read_stats"][2 * p + 1]
            arr = (Pairing * 1)(Pairing(int(g["pos1"]), int(g["pos2"]), int(g["strand1"]), int(g["mism1"]), int(g["strand2"]), int(g["mism2"]), int(g["optimalTotal"])))
            counts = (int(g["optimalTotal"]), int(g["suboptimalTotal"]), int(s1["x0"]), int(s2["x0"]), int(s1["x1"]), int(s2["x1"]), int(g["numOptimal"]),
                      int(int(s1["minMismatch"]) == int(g["mism1"])), int(int(s2["minMismatch"]) == int(g["mism2"])), int(g["numPairs"]))
            out = (Record * 2)()
            assert lib.s3_sam_pair_records(C.byref(one.gen), C.byref(cfg), arr, 1, 0, q1.ctypes.data_as(U8P), q2.ctypes.data_as(U8P), ql1.ctypes.data_as(C.c_char_p),
                                           ql2.ctypes.data_as(C.c_char_p), L, L, n1, n2, *counts, out) == 0
            mine = [record_tuple(r) for r in out]
            lines += [line_of(lib, one, out[0]) + b"\n", line_of(lib, one, out[1]) + b"\n"]
            for k in range(2):
                lib.s3_sam_record_free(C.byref(out[k]))
            # ---- the oracle chain's result -> the reference's writer
            o1, o2 = ostats[2 * p], ostats[2 * p + 1]
            flat = np.array([w["pos1"], w["strand1"], w["mism1"], w["pos2"], w["strand2"], w["mism2"], w["optimalTotal"] & 0xFF], np.uint32)
            wcounts = (w["optimalTotal"], w["suboptimalTotal"], o1[0], o2[0], o1[1], o2[1], w["numOptimal"], int(o1[2] == w["mism1"]), int(o2[2] == w["mism2"]), w["numPairs"])
            core, data, dlen = np.zeros(24, np.int32), np.zeros(2 * 8192, np.uint8), np.zeros(2, np.int32)
            assert ref.ref_sam_pair(*one.ref_args(), cfg.alignmentType, cfg.bwaLikeScore, cfg.dpMatchScore, cfg.dpMisMatchScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ,
                                    cfg.isPrintMDNM, cfg.readGroup, cfg.outputXAZTag, cfg.peMaxOutputPerPair, helpers.u32p(flat), 1, 0,
                                    q1.ctypes.data_as(U8P), q2.ctypes.data_as(U8P), ql1.ctypes.data_as(C.c_char_p), ql2.ctypes.data_as(C.c_char_p), L, L, n1, n2,
                                    *wcounts, core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), 8192, dlen.ctypes.data_as(I32P)) == 2
            theirs = [(tuple(int(x) for x in core[12 * r:12 * r + 12]), bytes(data[r * 8192:r * 8192 + int(dlen[r])])) for r in range(2)]
            assert mine == theirs, (p, mine, theirs)
            done += 1
        # the chain's result of the whole batch as SAM text (s3_sam_paired_batch_text) == those records' lines, pairs in order
        text, size = C.c_void_p(), C.c_uint64()
        rt, pr, stats = np.ascontiguousarray(got["route"], np.uint8), np.ascontiguousarray(got["pairs"]), np.ascontiguousarray(got["read_stats"])
        assert lib.s3_sam_paired_batch_text(C.byref(one.gen), C.byref(cfg), C.byref(rd), C.c_uint64(2 * pairs), rt.ctypes.data_as(U8P), pr.ctypes.data_as(C.c_void_p), C.c_uint64(pairs),
                                            stats.ctypes.data_as(C.c_void_p), 4, C.byref(text), C.byref(size)) == 0, lib.s3_last_error()
        assert C.string_at(text.value, size.value) == b"".join(lines) and len(lines) > pairs // 2
        lib.s3_free(text)
    assert done > pairs // 2


def test_rescued_pairs_through_the_chain_to_sam_records(env):
    """s3_pe_align's rescue records of a pair -> s3_runs_decode -> s3_sam_pair_dp_records == the oracle chain's records -> pairDPOutputSAMAPI.
    Every record of the pair is an AlgnmtDPResult (DV-DPfunctions.cu:2355-2440): whichFromDP = the DP read's parity (2 when it missed its
    cutoff), insert size as computed there; the reported entry = fewest mismatches of the aligned read, then the highest DP score
    (outputDPResult2, OutputDPResult.cpp:296-420)"""
    import test_pe_chain_gpu as chain
    from test_cpu_sam import DpPairing
    G, idx, hi, gi = env
    ref, lib = C.CDLL(REF), api.load_library()
    ref.ref_sam_pair_dp.restype = C.c_int
    lib.s3_sam_pair_dp_records.restype = C.c_int
    lib.s3_sam_pick_pair_dp.restype = C.c_int32
    lib.s3_runs_decode.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    one = OneChromosome(G)
    L, pairs = 100, 900
    got, want = chain._run_both(env, pairs, L, 61, read_stats=True, bad_mate_fraction=0.5)
    chain._compare(got, want)
    reads = chain.LAST["reads"]
    ostats = chain.oracle_read_stats(chain.LAST["views"], chain.LAST["allowed"], chain.LAST["text_length"], chain.LAST["max_output"])
    NONE = 0xFFFFFFFF
    mine_by, want_by = {}, {}
    for t in range(len(got["dp"])):
        mine_by.setdefault(int(got["dp"][t]["dpReadID"]) >> 1, []).append(t)
        want_by.setdefault(int(want["dp"][t]["dpReadID"]) >> 1, []).append(t)
    assert sorted(mine_by) == sorted(want_by) and len(mine_by) > pairs // 5

    def entry(rec, cigar, editdist, dis):
        """one rescue record as the fields of an AlgnmtDPResult: (whichFromDP, strands, editdist, insertSize, numSameScore, positions, scores)"""
        dp_read = int(rec["dpReadID"]) & 1
        al_pos, dp_pos = int(rec["alignedPos"]), int(rec["dpPos"])
        if cigar:
            which = dp_read
            ins = (al_pos - dp_pos + L) if dp_pos < al_pos else (dp_pos - al_pos + L + dis)
        else:
            which, dp_pos, ins, editdist = 2, NONE, 0, 0
        pos, strand, score = [0, 0], [0, 0], [0, 0]
        pos[dp_read], strand[dp_read], score[dp_read] = dp_pos, int(rec["dpStrand"]), int(rec["score"])
        pos[1 - dp_read], strand[1 - dp_read], score[1 - dp_read] = al_pos, int(rec["alignedStrand"]), int(rec["alignedMismatches"])
        return which, strand, editdist, ins, int(rec["numSameScore"]) if cigar else 0, pos, score

    def pick(entries):
        """outputDPResult2: the reported entry of a pair's records"""
        best, mn, mx = None, 0, 0
        for i, (which, strand, ed, ins, same, pos, score) in enumerate(entries):
            un1, un2 = pos[0] == NONE, pos[1] == NONE
            if un1 and not un2:
                cm, cs = score[1], -127
            elif un2 and not un1:
                cm, cs = score[0], -127
            elif not un1 and not un2:
                cm, cs = (score[0], score[1]) if which == 1 else (score[1], score[0])
            else:
                cm, cs = 127, -127
            if best is None or cm < mn or (cm == mn and cs > mx and not (un1 or un2)):
                best, mn, mx = i, cm, cs
        return best

    cfg = Config(1, 0, SCORES[0], SCORES[1], 1, 40, 1, 1, 1, 1000, b"rgRescue")
    rng = np.random.default_rng(31)
    done = 0
    quals = np.ascontiguousarray(rng.integers(2, 41, (2 * pairs, L + 1)).astype(np.uint8)); quals[:, -1] = 0
    lines = []
    for p in sorted(mine_by):
        q1, q2 = np.ascontiguousarray(reads[2 * p]).astype(np.uint8), np.ascontiguousarray(reads[2 * p + 1]).astype(np.uint8)
        ql1, ql2 = quals[2 * p], quals[2 * p + 1]
        n1, n2 = b"r%d/1" % p, b"r%d/2" % p
        # ---- the chain's records -> the writer's inputs
        ents, cigs = [], []
        for t in mine_by[p]:
            rec = got["dp"][t]
            if int(rec["numRuns"]):
                cg, ed, dis = runs_decode(lib, got["runs"][int(rec["runOffset"]):int(rec["runOffset"]) + int(rec["numRuns"])], L, int(rec["score"]))
            else:
                cg, ed, dis = b"", 0, 0
            ents.append(entry(rec, cg, ed, dis))
            cigs.append(cg)
        if all(e[0] == 2 for e in ents):
            continue                                                  # no rescue succeeded: the pair goes to the writers of improperly paired reads
        arr = (DpPairing * len(ents))()
        for k, (which, strand, ed, ins, same, pos, score) in enumerate(ents):
            arr[k].whichFromDP, arr[k].editdist, arr[k].insertSize, arr[k].numSameScore, arr[k].cigar = which, ed, ins, same, cigs[k] or None
            for i in range(2):
                arr[k].strand[i], arr[k].ambPosition[i], arr[k].score[i] = strand[i], pos[i], score[i]
        best = lib.s3_sam_pick_pair_dp(arr, len(ents))              # outputDPResult2's choice
        assert best == pick(ents)
        st = [got["read_stats"][2 * p], got["read_stats"][2 * p + 1]]
        x0 = (C.c_int32 * 2)(*[int(s_["x0"]) for s_ in st]); x1 = (C.c_int32 * 2)(*[int(s_["x1"]) for s_ in st])
        mm = (C.c_int32 * 2)(*[int(s_["minMismatch"]) if int(s_["x0"]) else 0 for s_ in st])
        out = (Record * 2)()
        assert lib.s3_sam_pair_dp_records(C.byref(one.gen), C.byref(cfg), arr, len(ents), best, q1.ctypes.data_as(U8P), q2.ctypes.data_as(U8P), ql1.ctypes.data_as(C.c_char_p),
                                          ql2.ctypes.data_as(C.c_char_p), L, L, n1, n2, x0, x1, mm, out) == 0, api.load_library().s3_last_error()
        mine = [record_tuple(r) for r in out]
        lines += [line_of(lib, one, out[0]) + b"\n", line_of(lib, one, out[1]) + b"\n"]
        for k in range(2):
            lib.s3_sam_record_free(C.byref(out[k]))
        # ---- the oracle chain's records -> the reference's writer
        wents, wcigs = [], []
        for t in want_by[p]:
            w = want["dp"][t]
            cg = w["cigar"]
            ed, dis = oracle_edit(cg, L, w["score"]) if cg else (0, 0)
            wents.append(entry(w, cg, ed, dis))
            wcigs.append(cg.encode())
        wbest = pick(wents)
        flat = []
        for k, (which, strand, ed, ins, same, pos, score) in enumerate(wents):
            flat += [which, ed, ins, same, pos[0] if pos[0] != NONE else -1, strand[0], score[0], pos[1] if pos[1] != NONE else -1, strand[1], score[1], k]
        flat = np.array(flat, np.int64).astype(np.int32)
        cig = (C.c_char_p * len(wcigs))(*wcigs)
        o = [ostats[2 * p], ostats[2 * p + 1]]
        counts = np.array([o[0][0], o[0][1], o[0][2] if o[0][0] else 0, o[1][0], o[1][1], o[1][2] if o[1][0] else 0], np.int32)
        core, data, dlen = np.zeros(24, np.int32), np.zeros(2 * 8192, np.uint8), np.zeros(2, np.int32)
        assert ref.ref_sam_pair_dp(*one.ref_args(), cfg.alignmentType, cfg.bwaLikeScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup,
                                   cfg.dpMatchScore, cfg.dpMisMatchScore, flat.ctypes.data_as(I32P), len(wents), wbest, cig, counts.ctypes.data_as(I32P),
                                   q1.ctypes.data_as(U8P), q2.ctypes.data_as(U8P), ql1.ctypes.data_as(C.c_char_p), ql2.ctypes.data_as(C.c_char_p), L, L, n1, n2,
                                   core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), 8192, dlen.ctypes.data_as(I32P)) == 2
        theirs = [(tuple(int(x) for x in core[12 * r:12 * r + 12]), bytes(data[r * 8192:r * 8192 + int(dlen[r])])) for r in range(2)]
        assert mine == theirs, (p, ents, mine, theirs)
        done += 1
    assert done > pairs // 6
    # the chain's rescue records of the whole batch as SAM text (s3_sam_pair_dp_batch_text) == those records' lines, pairs in record order
    padded = np.zeros((2 * pairs, L + 1), np.uint8); padded[:, :L] = np.asarray(reads)[:2 * pairs]
    rd, keep = batch_reads(padded, quals, np.full(2 * pairs, L, np.uint32), [b"r%d/%d" % (r >> 1, 1 + (r & 1)) for r in range(2 * pairs)])
    lib.s3_sam_pair_dp_batch_text.restype = C.c_int
    text, size = C.c_void_p(), C.c_uint64()
    dp, runs, stats = np.ascontiguousarray(got["dp"]), np.ascontiguousarray(got["runs"], np.uint32), np.ascontiguousarray(got["read_stats"])
    assert lib.s3_sam_pair_dp_batch_text(C.byref(one.gen), C.byref(cfg), C.byref(rd), C.c_uint64(2 * pairs), dp.ctypes.data_as(C.c_void_p), C.c_uint64(len(dp)), helpers.u32p(runs),
                                         C.c_uint64(len(runs)), api.DPScores(*SCORES), stats.ctypes.data_as(C.c_void_p), 4, C.byref(text), C.byref(size)) == 0, lib.s3_last_error()
    assert C.string_at(text.value, size.value) == b"".join(lines) and len(lines) == 2 * done
    lib.s3_free(text)
