"""CPU tier: s3_sam_pair_records (host entry) against the reference's own pairOutputSAMAPI and everything it calls, compiled from
where it lies into oracle/_ref/libref_sam.so with a samwrite that keeps the records (oracle/ref_shim/ref_sam_host.cpp)."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from helpers import ROOT, U32P
from soap3dp_b200 import api

U8P = C.POINTER(C.c_uint8)
I32P = C.POINTER(C.c_int32)
REF = os.path.join(ROOT, "oracle", "_ref", "libref_sam.so")


class Segment(C.Structure):
    _fields_ = [("startPos", C.c_uint32), ("chrID", C.c_uint32), ("correction", C.c_uint32)]


class Genome(C.Structure):
    _fields_ = [("packedDNA", U32P), ("dnaLength", C.c_uint32), ("segments", C.POINTER(Segment)), ("numSegments", C.c_uint32),
                ("ambiguityMap", U32P), ("chrEndPos", U32P), ("numChr", C.c_uint32), ("chrNames", C.POINTER(C.c_char_p))]


class Config(C.Structure):
    _fields_ = [("alignmentType", C.c_int32), ("bwaLikeScore", C.c_int32), ("dpMatchScore", C.c_int32), ("dpMisMatchScore", C.c_int32),
                ("isFastq", C.c_int32), ("maxMAPQ", C.c_int32), ("minMAPQ", C.c_int32), ("isPrintMDNM", C.c_int32), ("outputXAZTag", C.c_int32),
                ("peMaxOutputPerPair", C.c_uint32), ("readGroup", C.c_char_p)]


class Pairing(C.Structure):
    _fields_ = [("algnmt1", C.c_uint32), ("algnmt2", C.c_uint32), ("strand1", C.c_uint8), ("mismatch1", C.c_uint8), ("strand2", C.c_uint8),
                ("mismatch2", C.c_uint8), ("totalMismatchCount", C.c_int8), ("pad", C.c_uint8 * 3)]


class Record(C.Structure):
    _fields_ = [("tid", C.c_int32), ("pos", C.c_int32), ("bin", C.c_uint16), ("qual", C.c_uint8), ("l_qname", C.c_uint8), ("flag", C.c_uint16),
                ("n_cigar", C.c_uint16), ("l_qseq", C.c_int32), ("mtid", C.c_int32), ("mpos", C.c_int32), ("isize", C.c_int32), ("l_aux", C.c_int32),
                ("data_len", C.c_int32), ("data", U8P)]


def product_records(lib, gen, cfg, pairs, best, q1, q2, ql1, ql2, n1, n2, counts):
    arr = (Pairing * max(len(pairs), 1))()
    for k, p in enumerate(pairs):
        arr[k] = Pairing(p[0], p[3], p[1], p[2], p[4], p[5], p[6])
    out = (Record * 2)()
    lib.s3_sam_pair_records.restype = C.c_int
    rc = lib.s3_sam_pair_records(C.byref(gen), C.byref(cfg), arr, len(pairs), best, q1.ctypes.data_as(U8P), q2.ctypes.data_as(U8P),
                                 ql1.ctypes.data_as(C.c_char_p), ql2.ctypes.data_as(C.c_char_p), len(q1), len(q2), n1, n2, *counts, out)
    assert rc == 0, api.last_error() if hasattr(api, "last_error") else rc
    res = []
    for r in out:
        core = (r.tid, r.pos, r.bin, r.qual, r.l_qname, r.flag, r.n_cigar, r.l_qseq, r.mtid, r.mpos, r.isize, r.l_aux)
        res.append((core, bytes(bytearray(r.data[:r.data_len]))))
    lib.s3_sam_record_free.restype = None
    for k in range(2):
        lib.s3_sam_record_free(C.byref(out[k]))
    return res


def reference_records(ref, g, cfg, pairs, best, q1, q2, ql1, ql2, n1, n2, counts):
    flat = np.array([x for p in pairs for x in p], np.uint32) if pairs else np.zeros(7, np.uint32)
    core = np.zeros(24, np.int32)
    cap = 8192
    data = np.zeros(2 * cap, np.uint8)
    dlen = np.zeros(2, np.int32)
    names = (C.c_char_p * len(g["names"]))(*g["names"])
    ref.ref_sam_pair.restype = C.c_int
    n = ref.ref_sam_pair(helpers.u32p(g["pac"]), g["n"], helpers.u32p(g["translate"]), len(g["translate"]) // 3, helpers.u32p(g["amb"]),
                         helpers.u32p(g["chr_end"]), len(g["chr_end"]), names,
                         cfg.alignmentType, cfg.bwaLikeScore, cfg.dpMatchScore, cfg.dpMisMatchScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM,
                         cfg.readGroup, cfg.outputXAZTag, cfg.peMaxOutputPerPair, helpers.u32p(flat), len(pairs), best,
                         q1.ctypes.data_as(U8P), q2.ctypes.data_as(U8P), ql1.ctypes.data_as(C.c_char_p), ql2.ctypes.data_as(C.c_char_p), len(q1), len(q2), n1, n2,
                         *counts, core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), cap, dlen.ctypes.data_as(I32P))
    assert n == 2
    return [(tuple(int(x) for x in core[12 * r:12 * r + 12]), bytes(data[r * cap:r * cap + int(dlen[r])])) for r in range(2)]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_sam.so not built")
def test_sam_pair_records_match_the_reference_writer():
    ref = C.CDLL(REF)
    lib = api.load_library()
    rng = np.random.default_rng(11)
    n = 200_000
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    # three chromosomes; the second one in two segments (500 ambiguous bases were removed between them)
    starts = [0, 70_000, 100_000, 150_000]
    translate = np.array([0, 1, 0xFFFFFFFF,                      # tp = pos + 1
                          70_000, 2, 70_000 - 1,
                          100_000, 2, 70_000 - 1 - 500,
                          150_000, 3, 150_000 - 1], np.uint32)
    chr_end = np.array([69_999, 149_999, 199_999], np.uint32)
    amb = np.full(4, 3, np.uint32)
    names = [b"chr1", b"chrTwo", b"3"]
    g = dict(pac=pac, n=n, translate=translate, amb=amb, chr_end=chr_end, names=names)
    segs = (Segment * 4)(*[Segment(int(translate[3 * i]), int(translate[3 * i + 1]), int(translate[3 * i + 2])) for i in range(4)])
    gen = Genome(helpers.u32p(pac), n, segs, 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, (C.c_char_p * 3)(*names))
    trimmed = unmapped = with_xa = 0
    for trial in range(1500):
        L1, L2 = int(rng.integers(36, 152)), int(rng.integers(36, 152))
        cfg = Config(int(rng.integers(1, 5)), int(rng.integers(0, 2)), 1, -2, int(rng.integers(0, 2)), 40, 1, int(rng.integers(0, 2)), int(rng.integers(0, 2)),
                     int(rng.choice([1, 2, 5, 1000])), b"grp%d" % trial)
        # the reported pairing: positions near ends of chromosomes / segments now and then
        edges = [70_000, 100_000, 150_000, 200_000]
        def place(L):
            if rng.random() < 0.25:
                e = int(rng.choice(edges))
                return max(0, min(n - L, e - int(rng.integers(1, L))))
            return int(rng.integers(0, n - L))
        p1, p2 = place(L1), place(L2)
        if rng.random() < 0.6:
            p2 = max(0, min(n - L2, p1 + int(rng.integers(-400, 400))))
        s1, s2 = int(rng.integers(1, 3)), int(rng.integers(1, 3))

        def read_at(p, L, s, nm):
            r = G[p:p + L].copy()
            for k in rng.choice(L, nm, replace=False):
                r[k] = (r[k] + int(rng.integers(1, 4))) & 3
            return np.ascontiguousarray((3 - r[::-1]) if s == 2 else r).astype(np.uint8)
        m1, m2 = int(rng.integers(0, 4)), int(rng.integers(0, 4))
        q1, q2 = read_at(p1, L1, s1, m1), read_at(p2, L2, s2, m2)
        ql1 = np.ascontiguousarray(rng.integers(2, 41, L1 + 1).astype(np.uint8)); ql1[-1] = 0
        ql2 = np.ascontiguousarray(rng.integers(2, 41, L2 + 1).astype(np.uint8)); ql2[-1] = 0
        pairs = [(p1, s1, m1, p2, s2, m2, m1 + m2)]
        for _ in range(int(rng.integers(0, 6))):
            a, b = int(rng.integers(0, n - L1)), int(rng.integers(0, n - L2))
            x, y = int(rng.integers(0, 4)), int(rng.integers(0, 4))
            pairs.append((a, int(rng.integers(1, 3)), x, b, int(rng.integers(1, 3)), y, x + y))
        order = rng.permutation(len(pairs))
        pairs = [pairs[k] for k in order]
        best = int(np.where(order == 0)[0][0])
        if rng.random() < 0.05:
            best = -1
            unmapped += 1
        tot = sorted(p[6] for p in pairs)
        min_tot = m1 + m2 if rng.random() < 0.8 else tot[0]
        sec = [t for t in tot if t > min_tot]
        counts = (min_tot, sec[0] if sec else 127, int(rng.integers(0, 5)), int(rng.integers(0, 5)), int(rng.integers(0, 300)), int(rng.integers(0, 300)),
                  int(rng.integers(1, 4)), int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(1, 6)))
        if cfg.alignmentType >= 3 and trial % 2:                       # hostKernel's X0 / X1 of the unique-best / random-best report types (CPUfunctions.cpp:2326-2349;
            counts = counts[:2] + ((1, 1, -1, -1) if cfg.alignmentType == 3 else (-1, -1, -1, -1)) + counts[6:]      # with MAPQ computed, -1 indexes before g_log_n there)
        n1, n2 = b"read%d/1" % trial, b"read%d/2" % trial
        got = product_records(lib, gen, cfg, pairs, best, q1, q2, ql1, ql2, n1, n2, counts)
        want = reference_records(ref, g, cfg, pairs, best, q1, q2, ql1, ql2, n1, n2, counts)
        assert got == want, (trial, got, want)
        if best >= 0:
            trimmed += any(b"S" in w[1] or w[0][6] == 2 for w in want)
            with_xa += any(b"XAZ" in w[1] for w in want)
    assert trimmed > 20 and unmapped > 20 and with_xa > 100


class Occurrence(C.Structure):
    _fields_ = [("ambPosition", C.c_uint32), ("strand", C.c_uint8), ("mismatchCount", C.c_uint8), ("pad", C.c_uint8 * 2)]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_sam.so not built")
def test_sam_single_record_matches_the_reference_writer():
    """s3_sam_single_record against OCCOutputSAMAPI: best-hit choice, cross-chromosome handling, XA:Z, X0 / X1, MAPQ, unmapped"""
    ref = C.CDLL(REF)
    lib = api.load_library()
    rng = np.random.default_rng(12)
    n = 200_000
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    translate = np.array([0, 1, 0xFFFFFFFF, 70_000, 2, 70_000 - 1, 100_000, 2, 70_000 - 1 - 500, 150_000, 3, 150_000 - 1], np.uint32)
    chr_end = np.array([69_999, 149_999, 199_999], np.uint32)
    amb = np.full(4, 3, np.uint32)
    names = [b"chr1", b"chrTwo", b"3"]
    segs = (Segment * 4)(*[Segment(int(translate[3 * i]), int(translate[3 * i + 1]), int(translate[3 * i + 2])) for i in range(4)])
    gen = Genome(helpers.u32p(pac), n, segs, 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, (C.c_char_p * 3)(*names))
    cnames = (C.c_char_p * 3)(*names)
    edges = [70_000, 100_000, 150_000, 200_000]
    trimmed = unmapped = with_xa = 0
    for trial in range(1500):
        L = int(rng.integers(36, 152))
        cfg = Config(int(rng.integers(1, 3)), int(rng.integers(0, 2)), 1, -2, int(rng.integers(0, 2)), 40, 1, int(rng.integers(0, 2)), 1, 1000, b"rg%d" % trial)
        m = int(rng.choice([0, 1, 1, 2, 3, 6]))
        occ = []
        for _ in range(m):
            if rng.random() < 0.3:
                e = int(rng.choice(edges))
                p = max(0, min(n - L, e - int(rng.integers(1, L))))
            else:
                p = int(rng.integers(0, n - L))
            occ.append((p, int(rng.integers(1, 3)), int(rng.integers(0, 4))))
        p0, s0 = (occ[0][0], occ[0][1]) if occ else (0, 1)
        r = G[p0:p0 + L].copy()
        for k in rng.choice(L, int(rng.integers(0, 4)), replace=False):
            r[k] = (r[k] + 1) & 3
        q = np.ascontiguousarray((3 - r[::-1]) if s0 == 2 else r).astype(np.uint8)
        ql = np.ascontiguousarray(rng.integers(2, 41, L + 1).astype(np.uint8)); ql[-1] = 0
        name = b"single%d" % trial
        arr = (Occurrence * max(m, 1))(*[Occurrence(*o) for o in occ])
        out = Record()
        lib.s3_sam_single_record.restype = C.c_int
        assert lib.s3_sam_single_record(C.byref(gen), C.byref(cfg), arr, m, q.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name, C.byref(out)) == 0
        got = ((out.tid, out.pos, out.bin, out.qual, out.l_qname, out.flag, out.n_cigar, out.l_qseq, out.mtid, out.mpos, out.isize, out.l_aux),
               bytes(bytearray(out.data[:out.data_len])))
        lib.s3_sam_record_free.restype = None
        lib.s3_sam_record_free(C.byref(out))
        flat = np.array([x for o in occ for x in o], np.uint32) if occ else np.zeros(3, np.uint32)
        core = np.zeros(12, np.int32)
        data = np.zeros(8192, np.uint8)
        dlen = np.zeros(1, np.int32)
        ref.ref_sam_single.restype = C.c_int
        k = ref.ref_sam_single(helpers.u32p(pac), n, helpers.u32p(translate), 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, cnames,
                               cfg.alignmentType, cfg.bwaLikeScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup,
                               helpers.u32p(flat), m, q.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name,
                               core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), 8192, dlen.ctypes.data_as(I32P))
        assert k == 1
        want = (tuple(int(x) for x in core), bytes(data[:int(dlen[0])]))
        assert got == want, (trial, occ, got, want)
        unmapped += m == 0
        trimmed += want[0][6] == 2
        with_xa += b"XAZ" in want[1]
    assert unmapped > 100 and trimmed > 30 and with_xa > 300


class DpAlignment(C.Structure):
    _fields_ = [("ambPosition", C.c_uint32), ("strand", C.c_uint8), ("pad", C.c_uint8 * 3), ("score", C.c_int32), ("editdist", C.c_int32),
                ("cigar", C.c_char_p)]


def random_special_cigar(rng, L):
    """a special CIGAR (M match run, m mismatch run, I, D, S) whose read bases add up to L"""
    s1 = int(rng.integers(1, 6)) if rng.random() < 0.2 else 0
    s2 = int(rng.integers(1, 9)) if rng.random() < 0.2 else 0
    left = L - s1 - s2
    ops = []
    if rng.random() < 0.03:
        ops.append((int(rng.integers(1, 4)), "D"))                  # a leading deletion (convertToCigarStr drops it, its count stays)
    prev = ""
    while left > 0:
        op = "M" if prev != "M" and (not ops or rng.random() < 0.7) else str(rng.choice(["m", "I", "D", "m"]))
        if op == prev:
            op = "M"
        k = int(rng.integers(1, 40)) if op == "M" else int(rng.integers(1, 4))
        if op != "D":
            k = min(k, left)
            left -= k
        ops.append((k, op))
        prev = op
    if rng.random() < 0.03:
        ops.append((int(rng.integers(1, 4)), "D"))                  # a trailing deletion
    body = "".join(f"{k}{op}" for k, op in ops)
    return (f"{s1}S" if s1 else "") + body + (f"{s2}S" if s2 else "")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_sam.so not built")
def test_sam_single_dp_record_matches_the_reference_writer():
    """s3_sam_single_dp_record against SingleDPOutputSAMAPI: best score / X0, XA:Z with CIGARs and edit distances, x1_t1 / x1_t2, MD / XM / XO /
    XG / NM from the special CIGAR, MAPQ, alignments cut at a chromosome or segment end (BoundaryCheckDP), unmapped"""
    ref = C.CDLL(REF)
    lib = api.load_library()
    rng = np.random.default_rng(33)
    n = 200_000
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    translate = np.array([0, 1, 0xFFFFFFFF, 70_000, 2, 70_000 - 1, 100_000, 2, 70_000 - 1 - 500, 150_000, 3, 150_000 - 1], np.uint32)
    chr_end = np.array([69_999, 149_999, 199_999], np.uint32)
    amb = np.full(4, 3, np.uint32)
    names = [b"chr1", b"chrTwo", b"3"]
    segs = (Segment * 4)(*[Segment(int(translate[3 * i]), int(translate[3 * i + 1]), int(translate[3 * i + 2])) for i in range(4)])
    gen = Genome(helpers.u32p(pac), n, segs, 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, (C.c_char_p * 3)(*names))
    cnames = (C.c_char_p * 3)(*names)
    edges = [70_000, 100_000, 150_000]
    trimmed = unmapped = with_xa = 0
    lib.s3_sam_single_dp_record.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    ref.ref_sam_single_dp.restype = C.c_int
    for trial in range(1500):
        L = int(rng.integers(36, 152))
        cfg = Config(int(rng.integers(1, 4)), int(rng.integers(0, 2)), 1, -2, int(rng.integers(0, 2)), 40, 1, int(rng.integers(0, 2)), 1, 1000, b"rg%d" % trial)
        cutoff = int(0.3 * L)
        m = int(rng.choice([0, 1, 1, 2, 3, 6]))
        algn = []
        for _ in range(m):
            if rng.random() < 0.35:
                e = int(rng.choice(edges))
                p = max(0, e - int(rng.integers(1, L + 4)))
            else:
                p = int(rng.integers(0, n - 2 * L - 8))
            algn.append((p, int(rng.integers(1, 3)), int(rng.integers(cutoff, L + 1)) if rng.random() < 0.8 else L - 3, int(rng.integers(0, 9)),
                         random_special_cigar(rng, L).encode()))
        if m and rng.random() < 0.05:
            algn[0] = (0xFFFFFFFF,) + algn[0][1:]                    # the unaligned marker
        q = np.ascontiguousarray(rng.integers(0, 4, L).astype(np.uint8))
        ql = np.ascontiguousarray(rng.integers(2, 41, L + 1).astype(np.uint8)); ql[-1] = 0
        name = b"dp%d" % trial
        arr = (DpAlignment * max(m, 1))()
        for k, a in enumerate(algn):
            arr[k].ambPosition, arr[k].strand, arr[k].score, arr[k].editdist, arr[k].cigar = a
        out = Record()
        rc = lib.s3_sam_single_dp_record(C.byref(gen), C.byref(cfg), arr, m, cutoff, q.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name, C.byref(out))
        assert rc == 0, (trial, algn)
        got = ((out.tid, out.pos, out.bin, out.qual, out.l_qname, out.flag, out.n_cigar, out.l_qseq, out.mtid, out.mpos, out.isize, out.l_aux),
               bytes(bytearray(out.data[:out.data_len])))
        lib.s3_sam_record_free(C.byref(out))
        if m == 0:
            # the reference is never called without a result; its unaligned form is an entry with position 0xFFFFFFFF
            algn_ref = [(0xFFFFFFFF, 1, 0, 0, b"")]
        else:
            algn_ref = algn
        flat = np.array([[np.int64(a[0]).astype(np.int32) if a[0] > 0x7FFFFFFF else a[0], a[1], a[2], a[3], k] for k, a in enumerate(algn_ref)], np.int64).astype(np.int32).ravel()
        cig = (C.c_char_p * len(algn_ref))(*[a[4] for a in algn_ref])
        core = np.zeros(12, np.int32)
        data = np.zeros(8192, np.uint8)
        dlen = np.zeros(1, np.int32)
        k = ref.ref_sam_single_dp(helpers.u32p(pac), n, helpers.u32p(translate), 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, cnames,
                                  cfg.alignmentType, cfg.bwaLikeScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup,
                                  cfg.dpMatchScore, cutoff, flat.ctypes.data_as(I32P), len(algn_ref), cig,
                                  q.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name,
                                  core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), 8192, dlen.ctypes.data_as(I32P))
        assert k == 1
        want = (tuple(int(x) for x in core), bytes(data[:int(dlen[0])]))
        assert got == want, (trial, algn, got, want)
        unmapped += want[0][5] == 4
        trimmed += m == 1 and algn[0][0] != 0xFFFFFFFF and any(algn[0][0] < e <= algn[0][0] + L - 8 for e in edges)      # hangs over an end
        with_xa += b"XAZ" in want[1]
    assert unmapped > 100 and with_xa > 100 and trimmed > 40


class DeepAlignment(C.Structure):
    _fields_ = [("insertSize", C.c_int32), ("ambPosition", C.c_uint32 * 2), ("strand", C.c_uint8 * 2), ("pad", C.c_uint8 * 2),
                ("score", C.c_int32 * 2), ("editdist", C.c_int32 * 2), ("numSameScore", C.c_int32 * 2), ("cigar", C.c_char_p * 2)]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_sam.so not built")
def test_sam_deep_dp_records_match_the_reference_writer():
    """s3_sam_deep_dp_records against pairDeepDPOutputSAMAPI: X0 / X1 from the pair's other alignments and the search's counts, both MAPQ
    modes, XA:Z, read-through pairs (one read comes out unmapped), alignments cut at a chromosome or segment end, insert sizes, no result"""
    ref = C.CDLL(REF)
    lib = api.load_library()
    rng = np.random.default_rng(44)
    n = 200_000
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    translate = np.array([0, 1, 0xFFFFFFFF, 70_000, 2, 70_000 - 1, 100_000, 2, 70_000 - 1 - 500, 150_000, 3, 150_000 - 1], np.uint32)
    chr_end = np.array([69_999, 149_999, 199_999], np.uint32)
    amb = np.full(4, 3, np.uint32)
    names = [b"chr1", b"chrTwo", b"3"]
    segs = (Segment * 4)(*[Segment(int(translate[3 * i]), int(translate[3 * i + 1]), int(translate[3 * i + 2])) for i in range(4)])
    gen = Genome(helpers.u32p(pac), n, segs, 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, (C.c_char_p * 3)(*names))
    cnames = (C.c_char_p * 3)(*names)
    edges = [70_000, 100_000, 150_000]
    lib.s3_sam_deep_dp_records.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    ref.ref_sam_deep_dp.restype = C.c_int
    none = demoted = with_xa = 0
    for trial in range(1500):
        L1, L2 = int(rng.integers(50, 152)), int(rng.integers(50, 152))
        cfg = Config(int(rng.integers(1, 5)), int(rng.integers(0, 2)), 1, -2, int(rng.integers(0, 2)), 40, 1, int(rng.integers(0, 2)), 1, 1000, b"grp%d" % trial)
        m = int(rng.choice([0, 1, 1, 2, 3, 5]))
        algn = []
        for _ in range(m):
            if rng.random() < 0.3:
                e = int(rng.choice(edges))
                p1 = max(0, e - int(rng.integers(1, L1 + 4)))
            else:
                p1 = int(rng.integers(1000, n - 3000))
            gap = int(rng.integers(-40, 400))                        # a negative distance makes the reads run through each other
            s1 = int(rng.integers(1, 3))
            p2 = min(max(p1 + gap if s1 == 1 else p1 - gap, 0), n - 2 * L2 - 8)
            algn.append((int(rng.integers(150, 600)), (p1, p2), (s1, 3 - s1), (int(rng.integers(30, L1 + 1)), int(rng.integers(30, L2 + 1))),
                         (int(rng.integers(0, 9)), int(rng.integers(0, 9))), (int(rng.integers(1, 4)), int(rng.integers(1, 4))),
                         (random_special_cigar(rng, L1).encode(), random_special_cigar(rng, L2).encode())))
        if m > 1 and rng.random() < 0.3:                             # ties in score / position with the first entry
            a = algn[0]
            algn[1] = (algn[1][0], (a[1][0], algn[1][1][1]), algn[1][2], (a[3][0], a[3][1]), algn[1][4], algn[1][5], algn[1][6])
        best = int(rng.integers(0, m)) if m else -1
        counts = np.array([int(rng.integers(0, 4)), int(rng.integers(0, 4)), int(rng.integers(0, 5)),
                           int(rng.integers(0, 4)), int(rng.integers(0, 4)), int(rng.integers(0, 5))], np.int32)
        q1 = np.ascontiguousarray(rng.integers(0, 4, L1).astype(np.uint8)); q2 = np.ascontiguousarray(rng.integers(0, 4, L2).astype(np.uint8))
        ql1 = np.ascontiguousarray(rng.integers(2, 41, L1 + 1).astype(np.uint8)); ql1[-1] = 0
        ql2 = np.ascontiguousarray(rng.integers(2, 41, L2 + 1).astype(np.uint8)); ql2[-1] = 0
        n1, n2 = b"pair%d/1" % trial, b"pair%d/2" % trial
        arr = (DeepAlignment * max(m, 1))()
        for k, a in enumerate(algn):
            arr[k].insertSize = a[0]
            for i in range(2):
                arr[k].ambPosition[i], arr[k].strand[i], arr[k].score[i], arr[k].editdist[i], arr[k].numSameScore[i], arr[k].cigar[i] = a[1][i], a[2][i], a[3][i], a[4][i], a[5][i], a[6][i]
        x0 = (C.c_int32 * 2)(int(counts[0]), int(counts[3])); x1 = (C.c_int32 * 2)(int(counts[1]), int(counts[4])); mm = (C.c_int32 * 2)(int(counts[2]), int(counts[5]))
        out = (Record * 2)()
        rc = lib.s3_sam_deep_dp_records(C.byref(gen), C.byref(cfg), arr, m, best, q1.ctypes.data_as(U8P), q2.ctypes.data_as(U8P), ql1.ctypes.data_as(C.c_char_p),
                                        ql2.ctypes.data_as(C.c_char_p), L1, L2, n1, n2, x0, x1, mm, out)
        assert rc == 0, (trial, algn)
        got = []
        for r in out:
            got.append(((r.tid, r.pos, r.bin, r.qual, r.l_qname, r.flag, r.n_cigar, r.l_qseq, r.mtid, r.mpos, r.isize, r.l_aux), bytes(bytearray(r.data[:r.data_len]))))
        for k in range(2):
            lib.s3_sam_record_free(C.byref(out[k]))
        cigs, flat = [], []
        for a in algn:
            flat += [a[0], a[1][0], a[2][0], a[3][0], a[4][0], a[5][0], len(cigs), a[1][1], a[2][1], a[3][1], a[4][1], a[5][1], len(cigs) + 1]
            cigs += [a[6][0], a[6][1]]
        flat = np.array(flat if flat else [0] * 13, np.int64).astype(np.int32)
        cig = (C.c_char_p * max(len(cigs), 1))(*(cigs or [b""]))
        core = np.zeros(24, np.int32)
        cap = 8192
        data = np.zeros(2 * cap, np.uint8)
        dlen = np.zeros(2, np.int32)
        k = ref.ref_sam_deep_dp(helpers.u32p(pac), n, helpers.u32p(translate), 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, cnames,
                                cfg.alignmentType, cfg.bwaLikeScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup, cfg.dpMatchScore, cfg.dpMisMatchScore,
                                flat.ctypes.data_as(I32P), m, best, cig, counts.ctypes.data_as(I32P),
                                q1.ctypes.data_as(U8P), q2.ctypes.data_as(U8P), ql1.ctypes.data_as(C.c_char_p), ql2.ctypes.data_as(C.c_char_p), L1, L2, n1, n2,
                                core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), cap, dlen.ctypes.data_as(I32P))
        assert k == 2
        want = [(tuple(int(x) for x in core[12 * r:12 * r + 12]), bytes(data[r * cap:r * cap + int(dlen[r])])) for r in range(2)]
        assert got == want, (trial, algn, best, counts.tolist(), got, want)
        none += m == 0
        demoted += m > 0 and bool((want[0][0][5] | want[1][0][5]) & 4)
        with_xa += b"XAZ" in want[0][1]
    assert none > 100 and demoted > 50 and with_xa > 200


class DpPairing(C.Structure):
    _fields_ = [("whichFromDP", C.c_uint8), ("strand", C.c_uint8 * 2), ("pad", C.c_uint8), ("editdist", C.c_int32), ("insertSize", C.c_int32),
                ("numSameScore", C.c_int32), ("ambPosition", C.c_uint32 * 2), ("score", C.c_int32 * 2), ("cigar", C.c_char_p)]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_sam.so not built")
def test_sam_pair_dp_records_match_the_reference_writer():
    """s3_sam_pair_dp_records against pairDPOutputSAMAPI: one read from the search and one from DP per entry, entries of both kinds mixed,
    X0 / X1 by mismatches or by DP score, the second-best pair of the BWA-like MAPQ, both MAPQ modes, XA:Z in both forms, read-through
    pairs, trimmed alignments on either side, no result"""
    ref = C.CDLL(REF)
    lib = api.load_library()
    rng = np.random.default_rng(55)
    n = 200_000
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    translate = np.array([0, 1, 0xFFFFFFFF, 70_000, 2, 70_000 - 1, 100_000, 2, 70_000 - 1 - 500, 150_000, 3, 150_000 - 1], np.uint32)
    chr_end = np.array([69_999, 149_999, 199_999], np.uint32)
    amb = np.full(4, 3, np.uint32)
    names = [b"chr1", b"chrTwo", b"3"]
    segs = (Segment * 4)(*[Segment(int(translate[3 * i]), int(translate[3 * i + 1]), int(translate[3 * i + 2])) for i in range(4)])
    gen = Genome(helpers.u32p(pac), n, segs, 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, (C.c_char_p * 3)(*names))
    cnames = (C.c_char_p * 3)(*names)
    edges = [70_000, 100_000, 150_000]
    lib.s3_sam_pair_dp_records.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    ref.ref_sam_pair_dp.restype = C.c_int
    none = demoted = with_xa = 0
    for trial in range(2000):
        L = (int(rng.integers(50, 152)), int(rng.integers(50, 152)))
        cfg = Config(int(rng.integers(1, 5)), int(rng.integers(0, 2)), 1, -2, int(rng.integers(0, 2)), 40, 1, int(rng.integers(0, 2)), 1, 1000, b"grp%d" % trial)
        m = int(rng.choice([0, 1, 1, 2, 3, 6]))
        main_kind = int(rng.integers(0, 2))
        algn = []
        for _ in range(m):
            kind = main_kind if rng.random() < 0.8 else 1 - main_kind
            if rng.random() < 0.3:
                e = int(rng.choice(edges))
                p1 = max(0, e - int(rng.integers(1, L[0] + 4)))
            else:
                p1 = int(rng.integers(1000, n - 3000))
            gap = int(rng.integers(-40, 400))
            s1 = int(rng.integers(1, 3))
            p2 = min(max(p1 + gap if s1 == 1 else p1 - gap, 0), n - 2 * L[1] - 8)
            sc = [0, 0]
            sc[kind] = int(rng.integers(30, L[kind] + 1))                 # DP score of the read that came from DP
            sc[1 - kind] = int(rng.integers(0, 4))                         # mismatches of the read that came from the search
            algn.append((kind, (s1, 3 - s1), int(rng.integers(0, 9)), int(rng.integers(150, 600)), int(rng.integers(1, 4)), (p1, p2), tuple(sc),
                         random_special_cigar(rng, L[kind]).encode()))
        if m > 1 and rng.random() < 0.4:                                   # ties with the first entry
            a = algn[0]
            algn[1] = (a[0], algn[1][1], algn[1][2], algn[1][3], algn[1][4], (a[5][0], algn[1][5][1]) if rng.random() < 0.5 else algn[1][5], a[6],
                       random_special_cigar(rng, L[a[0]]).encode())
        best = int(rng.integers(0, m)) if m else -1
        counts = np.array([int(rng.integers(0, 4)), int(rng.integers(0, 4)), int(rng.integers(0, 5)),
                           int(rng.integers(0, 4)), int(rng.integers(0, 4)), int(rng.integers(0, 5))], np.int32)
        # the reads: where the reported entry's search-side read lies, with a few substitutions, so that its MD string has content
        q = []
        for k in range(2):
            if m and algn[best][0] != k and algn[best][5][k] + L[k] <= n:
                r = G[algn[best][5][k]:algn[best][5][k] + L[k]].copy()
                for j in rng.choice(L[k], int(rng.integers(0, 4)), replace=False):
                    r[j] = (r[j] + 1) & 3
                q.append(np.ascontiguousarray((3 - r[::-1]) if algn[best][1][k] == 2 else r).astype(np.uint8))
            else:
                q.append(np.ascontiguousarray(rng.integers(0, 4, L[k]).astype(np.uint8)))
        ql = []
        for k in range(2):
            x = np.ascontiguousarray(rng.integers(2, 41, L[k] + 1).astype(np.uint8)); x[-1] = 0
            ql.append(x)
        n1, n2 = b"pair%d/1" % trial, b"pair%d/2" % trial
        arr = (DpPairing * max(m, 1))()
        for k, a in enumerate(algn):
            arr[k].whichFromDP, arr[k].editdist, arr[k].insertSize, arr[k].numSameScore, arr[k].cigar = a[0], a[2], a[3], a[4], a[7]
            for i in range(2):
                arr[k].strand[i], arr[k].ambPosition[i], arr[k].score[i] = a[1][i], a[5][i], a[6][i]
        x0 = (C.c_int32 * 2)(int(counts[0]), int(counts[3])); x1 = (C.c_int32 * 2)(int(counts[1]), int(counts[4])); mm = (C.c_int32 * 2)(int(counts[2]), int(counts[5]))
        out = (Record * 2)()
        rc = lib.s3_sam_pair_dp_records(C.byref(gen), C.byref(cfg), arr, m, best, q[0].ctypes.data_as(U8P), q[1].ctypes.data_as(U8P), ql[0].ctypes.data_as(C.c_char_p),
                                        ql[1].ctypes.data_as(C.c_char_p), L[0], L[1], n1, n2, x0, x1, mm, out)
        assert rc == 0, (trial, algn)
        got = []
        for r in out:
            got.append(((r.tid, r.pos, r.bin, r.qual, r.l_qname, r.flag, r.n_cigar, r.l_qseq, r.mtid, r.mpos, r.isize, r.l_aux), bytes(bytearray(r.data[:r.data_len]))))
        for k in range(2):
            lib.s3_sam_record_free(C.byref(out[k]))
        flat = []
        for k, a in enumerate(algn):
            flat += [a[0], a[2], a[3], a[4], a[5][0], a[1][0], a[6][0], a[5][1], a[1][1], a[6][1], k]
        flat = np.array(flat if flat else [0] * 11, np.int64).astype(np.int32)
        cig = (C.c_char_p * max(m, 1))(*([a[7] for a in algn] or [b""]))
        core = np.zeros(24, np.int32)
        cap = 8192
        data = np.zeros(2 * cap, np.uint8)
        dlen = np.zeros(2, np.int32)
        k = ref.ref_sam_pair_dp(helpers.u32p(pac), n, helpers.u32p(translate), 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, cnames,
                                cfg.alignmentType, cfg.bwaLikeScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup, cfg.dpMatchScore, cfg.dpMisMatchScore,
                                flat.ctypes.data_as(I32P), m, best, cig, counts.ctypes.data_as(I32P),
                                q[0].ctypes.data_as(U8P), q[1].ctypes.data_as(U8P), ql[0].ctypes.data_as(C.c_char_p), ql[1].ctypes.data_as(C.c_char_p), L[0], L[1], n1, n2,
                                core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), cap, dlen.ctypes.data_as(I32P))
        assert k == 2
        want = [(tuple(int(x) for x in core[12 * r:12 * r + 12]), bytes(data[r * cap:r * cap + int(dlen[r])])) for r in range(2)]
        assert got == want, (trial, algn, best, counts.tolist(), got, want)
        none += m == 0
        demoted += m > 0 and bool((want[0][0][5] | want[1][0][5]) & 4)
        with_xa += b"XAZ" in want[0][1]
    assert none > 100 and demoted > 50 and with_xa > 200


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_sam.so not built")
def test_sam_unpaired_records_match_the_reference_writer():
    """s3_sam_unpaired_records against unproperlypairOutputSAMAPI: each read on its own (best occurrence, X0 / X1, XA:Z up to the cap,
    halved MAPQ), all four report types, one or both reads without an occurrence, mate fields and insert sizes"""
    ref = C.CDLL(REF)
    lib = api.load_library()
    rng = np.random.default_rng(66)
    n = 200_000
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    translate = np.array([0, 1, 0xFFFFFFFF, 70_000, 2, 70_000 - 1, 100_000, 2, 70_000 - 1 - 500, 150_000, 3, 150_000 - 1], np.uint32)
    chr_end = np.array([69_999, 149_999, 199_999], np.uint32)
    amb = np.full(4, 3, np.uint32)
    names = [b"chr1", b"chrTwo", b"3"]
    segs = (Segment * 4)(*[Segment(int(translate[3 * i]), int(translate[3 * i + 1]), int(translate[3 * i + 2])) for i in range(4)])
    gen = Genome(helpers.u32p(pac), n, segs, 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, (C.c_char_p * 3)(*names))
    cnames = (C.c_char_p * 3)(*names)
    lib.s3_sam_unpaired_records.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    ref.ref_sam_unpaired.restype = C.c_int
    one_sided = neither = with_xa = 0
    for trial in range(2000):
        L = (int(rng.integers(36, 152)), int(rng.integers(36, 152)))
        cfg = Config(int(rng.integers(1, 5)), int(rng.integers(0, 2)), 1, -2, int(rng.integers(0, 2)), 40, 1, int(rng.integers(0, 2)), 1, 1000, b"grp%d" % trial)
        cap = int(rng.choice([2, 3, 1000]))
        occs, q = [], []
        for k in range(2):
            m = int(rng.choice([0, 1, 1, 2, 4, 7]))
            lst = [(int(rng.integers(0, n - L[k])), int(rng.integers(1, 3)), int(rng.integers(0, 4))) for _ in range(m)]
            occs.append(lst)
            if lst:
                r = G[lst[0][0]:lst[0][0] + L[k]].copy()
                for j in rng.choice(L[k], int(rng.integers(0, 4)), replace=False):
                    r[j] = (r[j] + 1) & 3
                q.append(np.ascontiguousarray((3 - r[::-1]) if lst[0][1] == 2 else r).astype(np.uint8))
            else:
                q.append(np.ascontiguousarray(rng.integers(0, 4, L[k]).astype(np.uint8)))
        ql = []
        for k in range(2):
            x = np.ascontiguousarray(rng.integers(2, 41, L[k] + 1).astype(np.uint8)); x[-1] = 0
            ql.append(x)
        n1, n2 = b"u%d/1" % trial, b"u%d/2" % trial
        arrs = [(Occurrence * max(len(o), 1))(*[Occurrence(*x) for x in o]) for o in occs]
        out = (Record * 2)()
        rc = lib.s3_sam_unpaired_records(C.byref(gen), C.byref(cfg), arrs[0], len(occs[0]), arrs[1], len(occs[1]), cap, q[0].ctypes.data_as(U8P), q[1].ctypes.data_as(U8P),
                                         ql[0].ctypes.data_as(C.c_char_p), ql[1].ctypes.data_as(C.c_char_p), L[0], L[1], n1, n2, out)
        assert rc == 0, (trial, occs)
        got = []
        for r in out:
            got.append(((r.tid, r.pos, r.bin, r.qual, r.l_qname, r.flag, r.n_cigar, r.l_qseq, r.mtid, r.mpos, r.isize, r.l_aux), bytes(bytearray(r.data[:r.data_len]))))
        for k in range(2):
            lib.s3_sam_record_free(C.byref(out[k]))
        flats = [np.array([x for o in lst for x in o], np.uint32) if lst else np.zeros(3, np.uint32) for lst in occs]
        core = np.zeros(24, np.int32)
        dcap = 8192
        data = np.zeros(2 * dcap, np.uint8)
        dlen = np.zeros(2, np.int32)
        k = ref.ref_sam_unpaired(helpers.u32p(pac), n, helpers.u32p(translate), 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, cnames,
                                 cfg.alignmentType, cfg.bwaLikeScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup,
                                 helpers.u32p(flats[0]), len(occs[0]), helpers.u32p(flats[1]), len(occs[1]), cap,
                                 q[0].ctypes.data_as(U8P), q[1].ctypes.data_as(U8P), ql[0].ctypes.data_as(C.c_char_p), ql[1].ctypes.data_as(C.c_char_p), L[0], L[1], n1, n2,
                                 core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), dcap, dlen.ctypes.data_as(I32P))
        assert k == 2
        want = [(tuple(int(x) for x in core[12 * r:12 * r + 12]), bytes(data[r * dcap:r * dcap + int(dlen[r])])) for r in range(2)]
        assert got == want, (trial, occs, cfg.alignmentType, got, want)
        one_sided += bool(occs[0]) != bool(occs[1])
        neither += not occs[0] and not occs[1]
        with_xa += b"XAZ" in want[0][1]
    assert one_sided > 300 and neither > 30 and with_xa > 200


class ReadAlignment(C.Structure):
    _fields_ = [("ambPosition", C.c_uint32), ("strand", C.c_uint8), ("isFromDP", C.c_uint8), ("pad", C.c_uint8 * 2), ("score", C.c_int32), ("editdist", C.c_int32),
                ("cigar", C.c_char_p)]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_sam.so not built")
def test_sam_unpaired_dp_records_match_the_reference_writer():
    """s3_sam_unpaired_dp_records against unproperlypairDPOutputSAMAPI: lists that mix DP alignments and search hits, best score / X0 / X1,
    halved MAPQ (not when BWA-like), trims on either kind, the trailing-deletion term of the insert size, all report types, empty lists"""
    ref = C.CDLL(REF)
    lib = api.load_library()
    rng = np.random.default_rng(77)
    n = 200_000
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    translate = np.array([0, 1, 0xFFFFFFFF, 70_000, 2, 70_000 - 1, 100_000, 2, 70_000 - 1 - 500, 150_000, 3, 150_000 - 1], np.uint32)
    chr_end = np.array([69_999, 149_999, 199_999], np.uint32)
    amb = np.full(4, 3, np.uint32)
    names = [b"chr1", b"chrTwo", b"3"]
    segs = (Segment * 4)(*[Segment(int(translate[3 * i]), int(translate[3 * i + 1]), int(translate[3 * i + 2])) for i in range(4)])
    gen = Genome(helpers.u32p(pac), n, segs, 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, (C.c_char_p * 3)(*names))
    cnames = (C.c_char_p * 3)(*names)
    edges = [70_000, 100_000, 150_000]
    lib.s3_sam_unpaired_dp_records.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    ref.ref_sam_unpaired_dp.restype = C.c_int
    one_sided = with_xa = trailing = 0
    for trial in range(2000):
        L = (int(rng.integers(50, 152)), int(rng.integers(50, 152)))
        cfg = Config(int(rng.integers(1, 5)), int(rng.integers(0, 2)), 1, -2, int(rng.integers(0, 2)), 40, 1, int(rng.integers(0, 2)), 1, 1000, b"grp%d" % trial)
        cutoff = int(0.3 * min(L))
        lists, q = [], []
        near = int(rng.integers(1000, n - 3000))
        for k in range(2):
            m = int(rng.choice([0, 1, 1, 2, 4, 6]))
            lst = []
            for _ in range(m):
                r = rng.random()
                if r < 0.3:
                    p = max(0, int(rng.choice(edges)) - int(rng.integers(1, L[k] + 4)))
                elif r < 0.6:
                    p = min(max(near + int(rng.integers(-300, 300)), 0), n - 2 * L[k] - 8)      # on one chromosome with the mate: insert sizes
                else:
                    p = int(rng.integers(0, n - 2 * L[k] - 8))
                from_dp = int(rng.integers(0, 2))
                if from_dp:
                    cg = random_special_cigar(rng, L[k])
                    if rng.random() < 0.15 and not cg.endswith("D"):
                        cg += "%dD" % int(rng.integers(1, 4))
                    lst.append((p, int(rng.integers(1, 3)), int(rng.integers(cutoff, L[k] + 1)), int(rng.integers(0, 9)), 1, cg.encode()))
                else:
                    mm = int(rng.integers(0, 4))
                    lst.append((p, int(rng.integers(1, 3)), L[k] - 3 * mm, mm, 0, b"%dM" % L[k]))
            if m > 1 and rng.random() < 0.3:
                lst[1] = lst[1][:2] + (lst[0][2],) + lst[1][3:]                      # a tie in score
            lists.append(lst)
            if lst and lst[0][4] == 0 and lst[0][0] + L[k] <= n:
                r = G[lst[0][0]:lst[0][0] + L[k]].copy()
                for j in rng.choice(L[k], int(rng.integers(0, 4)), replace=False):
                    r[j] = (r[j] + 1) & 3
                q.append(np.ascontiguousarray((3 - r[::-1]) if lst[0][1] == 2 else r).astype(np.uint8))
            else:
                q.append(np.ascontiguousarray(rng.integers(0, 4, L[k]).astype(np.uint8)))
        ql = []
        for k in range(2):
            x = np.ascontiguousarray(rng.integers(2, 41, L[k] + 1).astype(np.uint8)); x[-1] = 0
            ql.append(x)
        n1, n2 = b"ud%d/1" % trial, b"ud%d/2" % trial
        arrs = []
        for lst in lists:
            a = (ReadAlignment * max(len(lst), 1))()
            for i, x in enumerate(lst):
                a[i].ambPosition, a[i].strand, a[i].score, a[i].editdist, a[i].isFromDP, a[i].cigar = x
            arrs.append(a)
        out = (Record * 2)()
        rc = lib.s3_sam_unpaired_dp_records(C.byref(gen), C.byref(cfg), arrs[0], len(lists[0]), arrs[1], len(lists[1]), cutoff, q[0].ctypes.data_as(U8P), q[1].ctypes.data_as(U8P),
                                            ql[0].ctypes.data_as(C.c_char_p), ql[1].ctypes.data_as(C.c_char_p), L[0], L[1], n1, n2, out)
        assert rc == 0, (trial, lists)
        got = []
        for r in out:
            got.append(((r.tid, r.pos, r.bin, r.qual, r.l_qname, r.flag, r.n_cigar, r.l_qseq, r.mtid, r.mpos, r.isize, r.l_aux), bytes(bytearray(r.data[:r.data_len]))))
        for k in range(2):
            lib.s3_sam_record_free(C.byref(out[k]))
        cigs, flats = [], []
        for lst in lists:
            f = []
            for x in lst:
                f += [x[0], x[1], x[2], x[3], x[4], len(cigs)]
                cigs.append(x[5])
            flats.append(np.array(f if f else [0] * 6, np.int64).astype(np.int32))
        cig = (C.c_char_p * max(len(cigs), 1))(*(cigs or [b""]))
        core = np.zeros(24, np.int32)
        dcap = 8192
        data = np.zeros(2 * dcap, np.uint8)
        dlen = np.zeros(2, np.int32)
        k = ref.ref_sam_unpaired_dp(helpers.u32p(pac), n, helpers.u32p(translate), 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, cnames,
                                    cfg.alignmentType, cfg.bwaLikeScore, cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup, cfg.dpMatchScore, cutoff,
                                    flats[0].ctypes.data_as(I32P), len(lists[0]), flats[1].ctypes.data_as(I32P), len(lists[1]), cig,
                                    q[0].ctypes.data_as(U8P), q[1].ctypes.data_as(U8P), ql[0].ctypes.data_as(C.c_char_p), ql[1].ctypes.data_as(C.c_char_p), L[0], L[1], n1, n2,
                                    core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), dcap, dlen.ctypes.data_as(I32P))
        assert k == 2
        want = [(tuple(int(x) for x in core[12 * r:12 * r + 12]), bytes(data[r * dcap:r * dcap + int(dlen[r])])) for r in range(2)]
        assert got == want, (trial, lists, cfg.alignmentType, cfg.bwaLikeScore, got, want)
        one_sided += bool(lists[0]) != bool(lists[1])
        with_xa += b"XAZ" in want[0][1]
        trailing += want[0][0][10] != 0 and any(x[5].endswith(b"D") for lst in lists for x in lst)
    assert one_sided > 300 and with_xa > 200 and trailing > 20


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_sam.so not built")
def test_sam_single_answer_and_no_answer_records_match_the_reference_writer():
    """s3_sam_single_answer_record against SingleAnsOutputSAMAPI, and s3_sam_single_record with no occurrence against noAnsOutputSAMAPI"""
    ref = C.CDLL(REF)
    lib = api.load_library()
    rng = np.random.default_rng(88)
    n = 200_000
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    translate = np.array([0, 1, 0xFFFFFFFF, 70_000, 2, 70_000 - 1, 100_000, 2, 70_000 - 1 - 500, 150_000, 3, 150_000 - 1], np.uint32)
    chr_end = np.array([69_999, 149_999, 199_999], np.uint32)
    amb = np.full(4, 3, np.uint32)
    names = [b"chr1", b"chrTwo", b"3"]
    segs = (Segment * 4)(*[Segment(int(translate[3 * i]), int(translate[3 * i + 1]), int(translate[3 * i + 2])) for i in range(4)])
    gen = Genome(helpers.u32p(pac), n, segs, 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, (C.c_char_p * 3)(*names))
    cnames = (C.c_char_p * 3)(*names)
    lib.s3_sam_single_answer_record.restype = C.c_int
    lib.s3_sam_single_record.restype = C.c_int
    lib.s3_sam_record_free.restype = None
    ref.ref_sam_single_answer.restype = C.c_int
    for trial in range(600):
        L = int(rng.integers(36, 152))
        cfg = Config(int(rng.integers(3, 5)), 0, 1, -2, int(rng.integers(0, 2)), 40, 1, int(rng.integers(0, 2)), 1, 1000, b"rg%d" % trial)
        none = trial % 6 == 0
        p, strand, hits = int(rng.integers(0, n - L)), int(rng.integers(1, 3)), int(rng.integers(-1, 4))
        r = G[p:p + L].copy()
        sub = rng.choice(L, int(rng.integers(0, 4)), replace=False)
        for j in sub:
            r[j] = (r[j] + 1) & 3
        q = np.ascontiguousarray((3 - r[::-1]) if strand == 2 else r).astype(np.uint8)
        ql = np.ascontiguousarray(rng.integers(2, 41, L + 1).astype(np.uint8)); ql[-1] = 0
        name = b"one%d" % trial
        out = Record()
        if none:
            rc = lib.s3_sam_single_record(C.byref(gen), C.byref(cfg), (Occurrence * 1)(), 0, q.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name, C.byref(out))
        else:
            rc = lib.s3_sam_single_answer_record(C.byref(gen), C.byref(cfg), p, strand, len(sub), hits, q.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name, C.byref(out))
        assert rc == 0
        got = ((out.tid, out.pos, out.bin, out.qual, out.l_qname, out.flag, out.n_cigar, out.l_qseq, out.mtid, out.mpos, out.isize, out.l_aux),
               bytes(bytearray(out.data[:out.data_len])))
        lib.s3_sam_record_free(C.byref(out))
        core = np.zeros(12, np.int32)
        data = np.zeros(8192, np.uint8)
        dlen = np.zeros(1, np.int32)
        k = ref.ref_sam_single_answer(helpers.u32p(pac), n, helpers.u32p(translate), 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, cnames,
                                      cfg.isFastq, cfg.maxMAPQ, cfg.minMAPQ, cfg.isPrintMDNM, cfg.readGroup, 0xFFFFFFFF if none else p, strand, len(sub), hits,
                                      q.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, name, core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), 8192, dlen.ctypes.data_as(I32P))
        assert k == 1
        want = (tuple(int(x) for x in core), bytes(data[:int(dlen[0])]))
        assert got == want, (trial, none, p, strand, hits, got, want)


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_sam.so not built")
def test_sam_text_lines_match_samtools_formatter():
    """s3_sam_format_line against samtools' own bam_format1 (what samwrite prints into a SAM file), on records of the writers: mapped with
    gapped CIGARs, soft clips, MD / XA:Z tags, reverse strand, mates on the same and on other chromosomes, unmapped"""
    ref = C.CDLL(REF)
    lib = api.load_library()
    rng = np.random.default_rng(99)
    n = 200_000
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    translate = np.array([0, 1, 0xFFFFFFFF, 70_000, 2, 70_000 - 1, 100_000, 2, 70_000 - 1 - 500, 150_000, 3, 150_000 - 1], np.uint32)
    chr_end = np.array([69_999, 149_999, 199_999], np.uint32)
    amb = np.full(4, 3, np.uint32)
    names = [b"chr1", b"chrTwo", b"3"]
    segs = (Segment * 4)(*[Segment(int(translate[3 * i]), int(translate[3 * i + 1]), int(translate[3 * i + 2])) for i in range(4)])
    gen = Genome(helpers.u32p(pac), n, segs, 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, (C.c_char_p * 3)(*names))
    cnames = (C.c_char_p * 3)(*names)
    lib.s3_sam_single_dp_record.restype = C.c_int
    lib.s3_sam_unpaired_records.restype = C.c_int
    lib.s3_sam_format_line.restype = C.c_int
    lib.s3_sam_format_line.argtypes = [C.POINTER(Record), C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_void_p)]
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    lib.s3_sam_record_free.restype = None
    ref.ref_sam_format.restype = C.c_int
    checked = 0

    def check(rec):
        nonlocal checked
        line = C.c_void_p()
        assert lib.s3_sam_format_line(C.byref(rec), cnames, 3, C.byref(line)) == 0
        mine = C.string_at(line.value)
        lib.s3_free(line)
        core = np.array([rec.tid, rec.pos, rec.bin, rec.qual, rec.l_qname, rec.flag, rec.n_cigar, rec.l_qseq, rec.mtid, rec.mpos, rec.isize, rec.l_aux], np.int32)
        data = np.frombuffer(bytes(bytearray(rec.data[:rec.data_len])), np.uint8)
        out = C.create_string_buffer(16384)
        k = ref.ref_sam_format(core.ctypes.data_as(I32P), data.ctypes.data_as(U8P), rec.data_len, cnames, 3, out, 16384)
        assert 0 < k < 16384 and mine == out.value, (mine, out.value)
        assert mine.count(b"\t") >= 11 and mine.split(b"\t")[0] == bytes(bytearray(rec.data[:rec.l_qname - 1]))
        checked += 1

    for trial in range(400):
        L = int(rng.integers(36, 152))
        cfg = Config(int(rng.integers(1, 3)), int(rng.integers(0, 2)), 1, -2, 1, 40, 1, 1, 1, 1000, b"grp%d" % trial)
        q = np.ascontiguousarray(rng.integers(0, 4, L).astype(np.uint8))
        ql = np.ascontiguousarray(rng.integers(2, 41, L + 1).astype(np.uint8)); ql[-1] = 0
        m = int(rng.choice([0, 1, 2, 4]))
        arr = (DpAlignment * max(m, 1))()
        keep = []
        for k in range(m):
            cg = random_special_cigar(rng, L).encode()
            keep.append(cg)
            arr[k].ambPosition, arr[k].strand, arr[k].score, arr[k].editdist, arr[k].cigar = int(rng.integers(0, n - 2 * L - 8)), int(rng.integers(1, 3)), int(rng.integers(30, L + 1)), int(rng.integers(0, 9)), cg
        out = Record()
        assert lib.s3_sam_single_dp_record(C.byref(gen), C.byref(cfg), arr, m, int(0.3 * L), q.ctypes.data_as(U8P), ql.ctypes.data_as(C.c_char_p), L, b"txt%d" % trial, C.byref(out)) == 0
        check(out)
        lib.s3_sam_record_free(C.byref(out))
        # a pair reported read by read: mate fields, '=' / other chromosome / unmapped mate
        occs = []
        for k in range(2):
            occs.append([(int(rng.integers(0, n - L)), int(rng.integers(1, 3)), int(rng.integers(0, 4))) for _ in range(int(rng.choice([0, 1, 3])))])
        arrs = [(Occurrence * max(len(o), 1))(*[Occurrence(*x) for x in o]) for o in occs]
        out2 = (Record * 2)()
        assert lib.s3_sam_unpaired_records(C.byref(gen), C.byref(cfg), arrs[0], len(occs[0]), arrs[1], len(occs[1]), 1000, q.ctypes.data_as(U8P), q.ctypes.data_as(U8P),
                                           ql.ctypes.data_as(C.c_char_p), ql.ctypes.data_as(C.c_char_p), L, L, b"u%d/1" % trial, b"u%d/2" % trial, out2) == 0
        for k in range(2):
            check(out2[k])
            lib.s3_sam_record_free(C.byref(out2[k]))
    assert checked == 1200


def test_sam_pick_reported_entry():
    """s3_sam_pick_deep_dp / s3_sam_pick_pair_dp == the scans of outputDeepDPResult2 / outputDPResult2 restated (OutputDPResult.cpp:263-350,
    590-760; the functions themselves need the whole aligner around them: unpinned restatement, used by tests/test_sam_e2e_gpu.py too)"""
    lib = api.load_library()
    lib.s3_sam_pick_deep_dp.restype = C.c_int32
    lib.s3_sam_pick_pair_dp.restype = C.c_int32
    rng = np.random.default_rng(111)
    NONE = 0xFFFFFFFF
    assert lib.s3_sam_pick_deep_dp(None, 0) == -1 and lib.s3_sam_pick_pair_dp(None, 0) == -1
    for trial in range(2000):
        m = int(rng.integers(1, 8))
        deep = (DeepAlignment * m)()
        want, mx = 0, None
        for k in range(m):
            u = int(rng.choice([0, 0, 0, 1, 2]))
            deep[k].ambPosition[0] = NONE if u == 1 else int(rng.integers(0, 1000))
            deep[k].ambPosition[1] = NONE if u == 2 else int(rng.integers(0, 1000))
            deep[k].score[0], deep[k].score[1] = int(rng.integers(30, 40)), int(rng.integers(30, 40))
            s = deep[k].score[1] if u == 1 else deep[k].score[0] if u == 2 else deep[k].score[0] + deep[k].score[1]
            if mx is None or s > mx:
                want, mx = k, s
        assert lib.s3_sam_pick_deep_dp(deep, m) == want
        pd = (DpPairing * m)()
        want, mn, mx = None, 0, 0
        for k in range(m):
            which = int(rng.integers(0, 2))
            failed = rng.random() < 0.2
            pd[k].whichFromDP = 2 if failed else which
            pd[k].ambPosition[which] = NONE if failed else int(rng.integers(0, 1000))
            pd[k].ambPosition[1 - which] = int(rng.integers(0, 1000))
            pd[k].score[which], pd[k].score[1 - which] = int(rng.integers(30, 36)), int(rng.integers(0, 3))
            cm = pd[k].score[1 - which]
            cs = -127 if failed else pd[k].score[which]
            if want is None or cm < mn or (not failed and cm == mn and cs > mx):
                want, mn, mx = k, cm, cs
        assert lib.s3_sam_pick_pair_dp(pd, m) == want, trial


class SamReads(C.Structure):
    _fields_ = [("bases", U8P), ("qualities", C.c_char_p), ("rowBytes", C.c_uint32), ("readLengths", C.POINTER(C.c_uint32)), ("names", C.POINTER(C.c_char_p))]


def _batch_genome(rng):
    n = 200_000
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    translate = np.array([0, 1, 0xFFFFFFFF, 70_000, 2, 70_000 - 1, 100_000, 2, 70_000 - 1 - 500, 150_000, 3, 150_000 - 1], np.uint32)
    chr_end = np.array([69_999, 149_999, 199_999], np.uint32)
    amb = np.full(4, 3, np.uint32)
    cnames = (C.c_char_p * 3)(b"chr1", b"chrTwo", b"3")
    segs = (Segment * 4)(*[Segment(int(translate[3 * i]), int(translate[3 * i + 1]), int(translate[3 * i + 2])) for i in range(4)])
    gen = Genome(helpers.u32p(pac), n, segs, 4, helpers.u32p(amb), helpers.u32p(chr_end), 3, cnames)
    return n, G, gen, cnames, (pac, chr_end, amb, segs)


def _line_of(lib, rec, cnames):
    line = C.c_void_p()
    assert lib.s3_sam_format_line(C.byref(rec), cnames, 3, C.byref(line)) == 0
    text = C.string_at(line.value)
    lib.s3_free(line)
    lib.s3_sam_record_free(C.byref(rec))
    return text


def _batch_lib():
    lib = api.load_library()
    lib.s3_sam_format_line.restype = C.c_int
    lib.s3_sam_format_line.argtypes = [C.POINTER(Record), C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_void_p)]
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    lib.s3_sam_record_free.restype = None
    for f in (lib.s3_sam_single_record, lib.s3_sam_single_dp_record, lib.s3_sam_single_batch_text, lib.s3_sam_single_dp_batch_text, lib.s3_runs_decode):
        f.restype = C.c_int
    return lib


def _batch_reads(rng, num, row):
    lens = rng.integers(36, 152, num).astype(np.uint32)
    bases = np.ascontiguousarray(rng.integers(0, 4, (num, row)).astype(np.uint8))
    quals = np.ascontiguousarray(rng.integers(2, 41, (num, row)).astype(np.uint8))
    names = (C.c_char_p * num)(*[b"batch%d" % r for r in range(num)])
    rd = SamReads(bases.ctypes.data_as(U8P), C.cast(quals.ctypes.data, C.c_char_p), row, lens.ctypes.data_as(C.POINTER(C.c_uint32)), names)
    return lens, bases, quals, names, rd


def _api_handles(n, keep, cfg, bases, quals, lens, num):
    """the same genome / config / reads as soap3dp_b200.api's objects"""
    pac, chr_end, amb, segs = keep
    g2 = api.SamGenomeDesc(pac, n, [(0, 1, 0xFFFFFFFF), (70_000, 2, 70_000 - 1), (100_000, 2, 70_000 - 1 - 500), (150_000, 3, 150_000 - 1)], amb, chr_end, [b"chr1", b"chrTwo", b"3"])
    c2 = api.SamConfig(*[getattr(cfg, f[0]) for f in Config._fields_])
    r2 = api.SamReads(bases, quals, lens, [b"batch%d" % r for r in range(num)])
    return g2, c2, r2


def test_sam_single_batch_text_is_the_reads_records_in_order():
    """s3_sam_single_batch_text == s3_sam_single_record + s3_sam_format_line per read (each pinned to the reference above), whatever the
    number of host threads; reads without an occurrence come out as unmapped records; bad arguments are refused"""
    lib = _batch_lib()
    rng = np.random.default_rng(314)
    n, G, gen, cnames, keep = _batch_genome(rng)
    num, row = 700, 160
    lens, bases, quals, names, rd = _batch_reads(rng, num, row)
    counts = rng.choice([0, 1, 1, 2, 3, 7], num)
    off = np.zeros(num + 1, np.uint32)
    off[1:] = np.cumsum(counts)
    tot = int(off[-1])
    pos = rng.integers(0, n - 160, tot).astype(np.uint32)
    edge = rng.random(tot) < 0.2
    pos[edge] = (rng.choice([70_000, 100_000, 150_000], int(edge.sum())) - rng.integers(1, 36, int(edge.sum()))).astype(np.uint32)
    flags = np.ascontiguousarray(np.stack([rng.integers(1, 3, tot), rng.integers(0, 4, tot)], 1).astype(np.uint8))
    lib.s3_sam_single_answer_record.restype = C.c_int
    for cfg in (Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA"), Config(2, 1, 1, -2, 0, 40, 1, 0, 1, 1000, b"rgB"), Config(3, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgC"),
                Config(4, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgD")):
        want = []
        for r in range(num):
            a, b = int(off[r]), int(off[r + 1])
            arr = (Occurrence * max(b - a, 1))(*[Occurrence(int(pos[i]), int(flags[i][0]), int(flags[i][1])) for i in range(a, b)])
            out = Record()
            args = (bases[r].ctypes.data_as(U8P), C.cast(quals[r].ctypes.data, C.c_char_p), int(lens[r]), names[r], C.byref(out))
            if cfg.alignmentType in (3, 4):
                # unique best: the read's one occurrence or nothing; random best: the first one (hostKernel, CPUfunctions.cpp:1862-1925)
                if b - a == 1 or (b - a > 1 and cfg.alignmentType == 4):
                    assert lib.s3_sam_single_answer_record(C.byref(gen), C.byref(cfg), int(pos[a]), int(flags[a][0]), int(flags[a][1]), 1, *args) == 0
                else:
                    assert lib.s3_sam_single_record(C.byref(gen), C.byref(cfg), arr, 0, *args) == 0
            else:
                assert lib.s3_sam_single_record(C.byref(gen), C.byref(cfg), arr, b - a, *args) == 0
            want.append(_line_of(lib, out, cnames))
        want = b"".join(x + b"\n" for x in want)
        assert want.count(b"\t4\t*\t0\t0\t*") >= 50                       # unmapped reads are in the batch
        for threads in (1, 5, 0):
            text, size = C.c_void_p(), C.c_uint64()
            rc = lib.s3_sam_single_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), helpers.u32p(off), helpers.u32p(pos), flags.ctypes.data_as(U8P), threads,
                                              C.byref(text), C.byref(size))
            assert rc == 0, lib.s3_last_error()
            got = C.string_at(text.value, size.value)
            lib.s3_free(text)
            assert got == want
    # an empty batch is an empty text; decreasing offsets and a read longer than its row are refused
    cfg = Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA")
    text, size = C.c_void_p(), C.c_uint64(7)
    assert lib.s3_sam_single_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(0), helpers.u32p(off), helpers.u32p(pos), flags.ctypes.data_as(U8P), 3, C.byref(text), C.byref(size)) == 0
    assert size.value == 0 and C.string_at(text.value) == b""
    lib.s3_free(text)
    bad = off.copy(); bad[5] = bad[4] - 1 if bad[4] else 0xFFFFFFFF
    assert lib.s3_sam_single_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), helpers.u32p(bad), helpers.u32p(pos), flags.ctypes.data_as(U8P), 2, C.byref(text), C.byref(size)) != 0
    long_lens = lens.copy(); long_lens[9] = row + 1
    rd2 = SamReads(rd.bases, rd.qualities, row, long_lens.ctypes.data_as(C.POINTER(C.c_uint32)), names)
    assert lib.s3_sam_single_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd2), C.c_uint64(num), helpers.u32p(off), helpers.u32p(pos), flags.ctypes.data_as(U8P), 2, C.byref(text), C.byref(size)) != 0
    assert text.value is None


def test_sam_single_dp_batch_text_groups_the_hits_of_a_read():
    """s3_sam_single_dp_batch_text == per read s3_runs_decode -> s3_sam_single_dp_record -> s3_sam_format_line over the read's hits (the
    grouping of outputDPSingleResult2, OutputDPResult.cpp:938-1058), whatever the number of host threads"""
    import re
    lib = _batch_lib()
    rng = np.random.default_rng(2718)
    n, G, gen, cnames, keep = _batch_genome(rng)
    num, row = 500, 160
    lens, bases, quals, names, rd = _batch_reads(rng, num, row)
    scores = api.DPScores(1, -2, -3, -1)
    hits, runs, per_read = [], [], {}
    for r in range(num):
        if rng.random() < 0.25:
            continue                                                      # a read without a hit: no record
        L = int(lens[r])
        for _ in range(int(rng.choice([1, 1, 2, 4]))):
            cg = random_special_cigar(rng, L)
            mine = [(int(k) << 8) | ord(op) for k, op in re.findall(r"(\d+)([MmIDS])", cg)]
            p = int(rng.choice([70_000, 100_000, 150_000])) - int(rng.integers(1, L)) if rng.random() < 0.2 else int(rng.integers(0, n - 2 * L - 8))
            h = (r, p, int(rng.integers(int(0.3 * L), L + 1)), int(rng.integers(1, 3)), len(runs), len(mine), int(rng.integers(1, 3)), 0)
            runs += mine
            hits.append(h)
            per_read.setdefault(r, []).append(h)
    harr = np.array(hits, api.DP_HIT_DTYPE)
    rarr = np.array(runs, np.uint32)
    cutoff = 30
    for cfg in (Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA"), Config(3, 1, 1, -2, 0, 40, 1, 1, 1, 1000, b"rgB")):
        want = []
        for r in sorted(per_read):
            L = int(lens[r])
            hs = per_read[r]
            arr = (DpAlignment * len(hs))()
            held = []
            for k, h in enumerate(hs):
                buf = C.create_string_buffer(4096)
                e, s = C.c_int32(), C.c_int32()
                sub = np.ascontiguousarray(rarr[h[4]:h[4] + h[5]])
                assert lib.s3_runs_decode(helpers.u32p(sub), h[5], L, h[2], scores, buf, 4096, None, C.byref(e), C.byref(s)) == 0
                held.append(buf.value)
                arr[k].ambPosition, arr[k].strand, arr[k].score, arr[k].editdist, arr[k].cigar = h[1], h[6], h[2], e.value, held[-1]
            out = Record()
            assert lib.s3_sam_single_dp_record(C.byref(gen), C.byref(cfg), arr, len(hs), cutoff, bases[r].ctypes.data_as(U8P), C.cast(quals[r].ctypes.data, C.c_char_p), L, names[r],
                                               C.byref(out)) == 0
            want.append(_line_of(lib, out, cnames))
        want = b"".join(x + b"\n" for x in want)
        for threads in (1, 4, 0):
            text, size = C.c_void_p(), C.c_uint64()
            rc = lib.s3_sam_single_dp_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), harr.ctypes.data_as(C.c_void_p), C.c_uint64(len(harr)),
                                                 helpers.u32p(rarr), C.c_uint64(len(rarr)), scores, cutoff, threads, C.byref(text), C.byref(size))
            assert rc == 0, lib.s3_last_error()
            got = C.string_at(text.value, size.value)
            lib.s3_free(text)
            assert got == want and got.count(b"\n") == len(per_read)
        g2, c2, r2 = _api_handles(n, keep, cfg, bases, quals, lens, num)
        assert api.sam_single_dp_batch_text(g2, c2, r2, harr, rarr, scores, cutoff, num_threads=3) == want
    # a hit whose runs lie outside the run array is refused
    cfg = Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA")
    bad = harr.copy(); bad[3]["runOffset"] = len(rarr)
    text, size = C.c_void_p(), C.c_uint64()
    assert lib.s3_sam_single_dp_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), bad.ctypes.data_as(C.c_void_p), C.c_uint64(len(bad)),
                                           helpers.u32p(rarr), C.c_uint64(len(rarr)), scores, cutoff, 2, C.byref(text), C.byref(size)) != 0


def _cigar_runs(cg):
    import re
    return [(int(k) << 8) | ord(op) for k, op in re.findall(r"(\d+)([MmIDS])", cg)]


def _decode(lib, runs, L, score, scores):
    buf = C.create_string_buffer(4096)
    e, s = C.c_int32(), C.c_int32()
    sub = np.ascontiguousarray(runs, np.uint32)
    assert lib.s3_runs_decode(helpers.u32p(sub), len(sub), L, score, scores, buf, 4096, None, C.byref(e), C.byref(s)) == 0
    return buf.value, e.value, s.value


def _stats(rng, num):
    st = np.zeros(num, api.PE_READ_STATS_DTYPE)
    st["x0"] = rng.integers(0, 4, num)
    st["x1"] = rng.integers(0, 4, num)
    st["minMismatch"] = np.where(st["x0"] > 0, rng.integers(0, 4, num), 255)
    return st


def _counts(st, r):
    x0 = (C.c_int32 * 2)(*[int(st[r + k]["x0"]) for k in range(2)])
    x1 = (C.c_int32 * 2)(*[int(st[r + k]["x1"]) for k in range(2)])
    mm = (C.c_int32 * 2)(*[int(st[r + k]["minMismatch"]) if int(st[r + k]["x0"]) else 0 for k in range(2)])
    return x0, x1, mm


def test_sam_deep_dp_batch_text_is_the_pairs_records_in_order():
    """s3_sam_deep_dp_batch_text == per pair: the hits' runs -> s3_runs_decode, insert size (DV-DPfunctions.cu:3810-3815), s3_sam_pick_deep_dp,
    s3_sam_deep_dp_records, s3_sam_format_line -- with and without the search's per-read statistics, whatever the number of host threads"""
    lib = _batch_lib()
    lib.s3_sam_deep_dp_records.restype = C.c_int
    lib.s3_sam_deep_dp_batch_text.restype = C.c_int
    lib.s3_sam_pick_deep_dp.restype = C.c_int32
    rng = np.random.default_rng(1618)
    n, G, gen, cnames, keep = _batch_genome(rng)
    num, row = 600, 160
    lens, bases, quals, names, rd = _batch_reads(rng, num, row)
    scores = api.DPScores(1, -2, -3, -1)
    st = _stats(rng, num)
    hits, runs, per_pair = [], [], {}
    for r in range(0, num, 2):
        if rng.random() < 0.3:
            continue
        L1, L2 = int(lens[r]), int(lens[r + 1])
        for _ in range(int(rng.choice([1, 1, 2, 4]))):
            r1, r2 = _cigar_runs(random_special_cigar(rng, L1)), _cigar_runs(random_special_cigar(rng, L2))
            p1 = int(rng.choice([70_000, 100_000, 150_000])) - int(rng.integers(1, L1)) if rng.random() < 0.2 else int(rng.integers(1000, n - 3000))
            s1 = int(rng.integers(1, 3))
            gap = int(rng.integers(-30, 400))
            p2 = min(max(p1 + gap if s1 == 1 else p1 - gap, 0), n - 2 * L2 - 8)
            h = (r, p1, p2, int(rng.integers(30, L1 + 1)), int(rng.integers(30, L2 + 1)), int(rng.integers(1, 3)), int(rng.integers(1, 3)), len(runs), len(runs) + len(r1),
                 len(r1), len(r2), s1, 3 - s1, (0, 0))
            runs += r1 + r2
            hits.append(h)
            per_pair.setdefault(r, []).append(h)
    harr = np.array(hits, api.DEEP_HIT_DTYPE)
    rarr = np.array(runs, np.uint32)
    for cfg, stats in ((Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA"), st), (Config(2, 1, 1, -2, 0, 40, 1, 1, 1, 1000, b"rgB"), None)):
        want = []
        for r in sorted(per_pair):
            L1, L2 = int(lens[r]), int(lens[r + 1])
            hs = per_pair[r]
            arr = (DeepAlignment * len(hs))()
            held = []
            for k, h in enumerate(hs):
                d1 = _decode(lib, rarr[h[7]:h[7] + h[9]], L1, h[3], scores)
                d2 = _decode(lib, rarr[h[8]:h[8] + h[10]], L2, h[4], scores)
                held += [d1[0], d2[0]]
                arr[k].insertSize = (h[2] - h[1] + L2 + d2[2]) if h[1] < h[2] else (h[1] - h[2] + L1 + d1[2])
                arr[k].ambPosition[0], arr[k].ambPosition[1] = h[1], h[2]
                arr[k].strand[0], arr[k].strand[1] = h[11], h[12]
                arr[k].score[0], arr[k].score[1] = h[3], h[4]
                arr[k].editdist[0], arr[k].editdist[1] = d1[1], d2[1]
                arr[k].numSameScore[0], arr[k].numSameScore[1] = h[5], h[6]
                arr[k].cigar[0], arr[k].cigar[1] = d1[0], d2[0]
            x0, x1, mm = _counts(stats, r) if stats is not None else ((C.c_int32 * 2)(0, 0),) * 3
            out = (Record * 2)()
            assert lib.s3_sam_deep_dp_records(C.byref(gen), C.byref(cfg), arr, len(hs), lib.s3_sam_pick_deep_dp(arr, len(hs)), bases[r].ctypes.data_as(U8P), bases[r + 1].ctypes.data_as(U8P),
                                              C.cast(quals[r].ctypes.data, C.c_char_p), C.cast(quals[r + 1].ctypes.data, C.c_char_p), L1, L2, names[r], names[r + 1], x0, x1, mm, out) == 0
            want += [_line_of(lib, out[0], cnames), _line_of(lib, out[1], cnames)]
        want = b"".join(x + b"\n" for x in want)
        for threads in (1, 6, 0):
            text, size = C.c_void_p(), C.c_uint64()
            rc = lib.s3_sam_deep_dp_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), harr.ctypes.data_as(C.c_void_p), C.c_uint64(len(harr)), helpers.u32p(rarr),
                                               C.c_uint64(len(rarr)), scores, stats.ctypes.data_as(C.c_void_p) if stats is not None else None, threads, C.byref(text), C.byref(size))
            assert rc == 0, lib.s3_last_error()
            got = C.string_at(text.value, size.value)
            lib.s3_free(text)
            assert got == want and got.count(b"\n") == 2 * len(per_pair)
        g2, c2, r2 = _api_handles(n, keep, cfg, bases, quals, lens, num)
        assert api.sam_deep_dp_batch_text(g2, c2, r2, harr, rarr, scores, stats, num_threads=3) == want
    bad = harr.copy(); bad[2]["readID"] += 1                                # an odd read id is not a pair's
    text, size = C.c_void_p(), C.c_uint64()
    assert lib.s3_sam_deep_dp_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), bad.ctypes.data_as(C.c_void_p), C.c_uint64(len(bad)), helpers.u32p(rarr),
                                         C.c_uint64(len(rarr)), scores, None, 2, C.byref(text), C.byref(size)) != 0


def test_sam_pair_dp_batch_text_is_the_rescued_pairs_records_in_order():
    """s3_sam_pair_dp_batch_text == per pair: every rescue record as the AlgnmtDPResult of the default-DP engine (whichFromDP, insert size,
    the unaligned form of a DP side under its cutoff), s3_sam_pick_pair_dp, s3_sam_pair_dp_records, s3_sam_format_line; pairs none of whose
    rescues succeeded get no lines"""
    lib = _batch_lib()
    lib.s3_sam_pair_dp_records.restype = C.c_int
    lib.s3_sam_pair_dp_batch_text.restype = C.c_int
    lib.s3_sam_pick_pair_dp.restype = C.c_int32
    rng = np.random.default_rng(5772)
    n, G, gen, cnames, keep = _batch_genome(rng)
    num, row = 600, 160
    lens, bases, quals, names, rd = _batch_reads(rng, num, row)
    scores = api.DPScores(1, -2, -3, -1)
    st = _stats(rng, num)
    NONE = 0xFFFFFFFF
    recs, runs, per_pair = [], [], {}
    for r in range(0, num, 2):
        if rng.random() < 0.3:
            continue
        main = int(rng.integers(0, 2))
        for _ in range(int(rng.choice([1, 1, 2, 4]))):
            dp_side = main if rng.random() < 0.8 else 1 - main
            Ld = int(lens[r + dp_side])
            ok = rng.random() < 0.75
            mine = _cigar_runs(random_special_cigar(rng, Ld)) if ok else []
            pa = int(rng.choice([70_000, 100_000, 150_000])) - int(rng.integers(1, 36)) if rng.random() < 0.15 else int(rng.integers(1000, n - 3000))
            sa = int(rng.integers(1, 3))
            gap = int(rng.integers(-30, 400))
            pd = min(max(pa + gap if sa == 1 else pa - gap, 0), n - 2 * Ld - 8)
            x = (r + dp_side, pa, pd if ok else 0, int(rng.integers(30, Ld + 1)) if ok else int(rng.integers(0, 29)), int(rng.integers(1, 3)), len(runs), len(mine),
                 sa, int(rng.integers(0, 4)), 3 - sa, int(rng.integers(0, 2)), (0, 0))
            runs += mine
            recs.append(x)
            per_pair.setdefault(r, []).append(x)
    darr = np.array(recs, api.PE_DP_DTYPE)
    rarr = np.array(runs, np.uint32)
    silent = 0
    for cfg in (Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA"), Config(2, 1, 1, -2, 0, 40, 1, 1, 1, 1000, b"rgB")):
        want = []
        silent = 0
        for r in sorted(per_pair):
            xs = per_pair[r]
            arr = (DpPairing * len(xs))()
            held = []
            for k, x in enumerate(xs):
                side = x[0] & 1
                if x[6]:
                    cg, ed, dis = _decode(lib, rarr[x[5]:x[5] + x[6]], int(lens[x[0]]), x[3], scores)
                    held.append(cg)
                    arr[k].whichFromDP, arr[k].editdist, arr[k].numSameScore, arr[k].cigar = side, ed, x[4], cg
                    arr[k].insertSize = (x[1] - x[2] + int(lens[x[0] ^ 1])) if x[2] < x[1] else (x[2] - x[1] + int(lens[x[0]]) + dis)
                    dp_pos = x[2]
                else:
                    arr[k].whichFromDP, dp_pos = 2, NONE
                arr[k].ambPosition[side], arr[k].strand[side], arr[k].score[side] = dp_pos, x[9], x[3]
                arr[k].ambPosition[1 - side], arr[k].strand[1 - side], arr[k].score[1 - side] = x[1], x[7], x[8]
            if all(a.whichFromDP == 2 for a in arr):
                silent += 1
                continue
            x0, x1, mm = _counts(st, r)
            out = (Record * 2)()
            assert lib.s3_sam_pair_dp_records(C.byref(gen), C.byref(cfg), arr, len(xs), lib.s3_sam_pick_pair_dp(arr, len(xs)), bases[r].ctypes.data_as(U8P), bases[r + 1].ctypes.data_as(U8P),
                                              C.cast(quals[r].ctypes.data, C.c_char_p), C.cast(quals[r + 1].ctypes.data, C.c_char_p), int(lens[r]), int(lens[r + 1]), names[r], names[r + 1],
                                              x0, x1, mm, out) == 0, lib.s3_last_error()
            want += [_line_of(lib, out[0], cnames), _line_of(lib, out[1], cnames)]
        want = b"".join(x + b"\n" for x in want)
        for threads in (1, 6, 0):
            text, size = C.c_void_p(), C.c_uint64()
            rc = lib.s3_sam_pair_dp_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), darr.ctypes.data_as(C.c_void_p), C.c_uint64(len(darr)), helpers.u32p(rarr),
                                               C.c_uint64(len(rarr)), scores, st.ctypes.data_as(C.c_void_p), threads, C.byref(text), C.byref(size))
            assert rc == 0, lib.s3_last_error()
            got = C.string_at(text.value, size.value)
            lib.s3_free(text)
            assert got == want
        g2, c2, r2 = _api_handles(n, keep, cfg, bases, quals, lens, num)
        assert api.sam_pair_dp_batch_text(g2, c2, r2, darr, rarr, scores, st, num_threads=3) == want
    assert silent > 10


def test_sam_paired_batch_text_is_the_paired_reads_records_in_order():
    """s3_sam_paired_batch_text == s3_sam_pair_records + s3_sam_format_line for the pairs with route 1 and one valid pairing (the chain's
    reported pairing, totals and per-read statistics as the writer's counts); other routes and pairs with more pairings get no lines"""
    lib = _batch_lib()
    lib.s3_sam_pair_records.restype = C.c_int
    lib.s3_sam_paired_batch_text.restype = C.c_int
    rng = np.random.default_rng(8128)
    n, G, gen, cnames, keep = _batch_genome(rng)
    num, row = 800, 160
    lens, bases, quals, names, rd = _batch_reads(rng, num, row)
    P = num // 2
    st = _stats(rng, num)
    st["x0"] = np.maximum(st["x0"], 1)
    st["minMismatch"] = rng.integers(0, 3, num)
    route = rng.choice([1, 1, 1, 2, 4, 0], P).astype(np.uint8)
    pr = np.zeros(P, api.PE_PAIR_DTYPE)
    for p in range(P):
        L1, L2 = int(lens[2 * p]), int(lens[2 * p + 1])
        p1 = int(rng.choice([70_000, 100_000, 150_000])) - int(rng.integers(1, L1)) if rng.random() < 0.15 else int(rng.integers(1000, n - 3000))
        s1 = int(rng.integers(1, 3))
        gap = int(rng.integers(50, 400))
        p2 = min(max(p1 + gap if s1 == 1 else p1 - gap, 0), n - L2)
        # the reads: the text at the pairing's positions with a few substitutions, so that MD and the mismatch counts have content
        for k, (pos, sd, L) in enumerate(((p1, s1, L1), (p2, 3 - s1, L2))):
            r = G[pos:pos + L].copy()
            for j in rng.choice(L, int(rng.integers(0, 3)), replace=False):
                r[j] = (r[j] + 1) & 3
            bases[2 * p + k, :L] = (3 - r[::-1]) if sd == 2 else r
        m1, m2 = int(rng.integers(0, 3)), int(rng.integers(0, 3))
        pr[p] = (p1, p2, abs(p2 - p1) + L2, s1, m1, 3 - s1, m2, int(rng.choice([1, 1, 1, 2])), int(rng.integers(1, 3)), int(rng.integers(0, 3)), m1 + m2,
                 int(rng.choice([127, m1 + m2 + 1])), 0)
    for cfg in (Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA"), Config(2, 1, 1, -2, 0, 40, 1, 1, 1, 1000, b"rgB"), Config(3, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgC"),
                Config(4, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgD")):
        want, done = [], 0
        for p in range(P):
            x = pr[p]
            if int(route[p]) != 1 or int(x["numPairs"]) != 1:
                continue
            arr = (Pairing * 1)(Pairing(int(x["pos1"]), int(x["pos2"]), int(x["strand1"]), int(x["mism1"]), int(x["strand2"]), int(x["mism2"]), int(x["optimalTotal"])))
            s1, s2 = st[2 * p], st[2 * p + 1]
            # X0 / X1 per report type as hostKernel passes them (CPUfunctions.cpp:2326-2362)
            xs = {3: (1, 1, -1, -1), 4: (-1, -1, -1, -1)}.get(cfg.alignmentType, (int(s1["x0"]), int(s2["x0"]), int(s1["x1"]), int(s2["x1"])))
            counts = (int(x["optimalTotal"]), int(x["suboptimalTotal"]), *xs, int(x["numOptimal"]),
                      int(int(s1["minMismatch"]) == int(x["mism1"])), int(int(s2["minMismatch"]) == int(x["mism2"])), int(x["numPairs"]))
            out = (Record * 2)()
            assert lib.s3_sam_pair_records(C.byref(gen), C.byref(cfg), arr, 1, 0, bases[2 * p].ctypes.data_as(U8P), bases[2 * p + 1].ctypes.data_as(U8P),
                                           C.cast(quals[2 * p].ctypes.data, C.c_char_p), C.cast(quals[2 * p + 1].ctypes.data, C.c_char_p), int(lens[2 * p]), int(lens[2 * p + 1]),
                                           names[2 * p], names[2 * p + 1], *counts, out) == 0, lib.s3_last_error()
            want += [_line_of(lib, out[0], cnames), _line_of(lib, out[1], cnames)]
            done += 1
        want = b"".join(x + b"\n" for x in want)
        assert 100 < done < P
        for threads in (1, 6, 0):
            text, size = C.c_void_p(), C.c_uint64()
            rc = lib.s3_sam_paired_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), route.ctypes.data_as(U8P), pr.ctypes.data_as(C.c_void_p), C.c_uint64(P),
                                              st.ctypes.data_as(C.c_void_p), threads, C.byref(text), C.byref(size))
            assert rc == 0, lib.s3_last_error()
            got = C.string_at(text.value, size.value)
            lib.s3_free(text)
            assert got == want
        g2, c2, r2 = _api_handles(n, keep, cfg, bases, quals, lens, num)
        assert api.sam_paired_batch_text(g2, c2, r2, route, pr, st, num_threads=3) == want
    text, size = C.c_void_p(), C.c_uint64()
    assert lib.s3_sam_paired_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), route.ctypes.data_as(U8P), pr.ctypes.data_as(C.c_void_p), C.c_uint64(P),
                                        None, 2, C.byref(text), C.byref(size)) != 0                   # no statistics, no counts for the writer


def test_sam_unpaired_batch_text_is_the_named_pairs_records_in_order():
    """s3_sam_unpaired_batch_text == s3_sam_unpaired_records + s3_sam_format_line for the pairs named, from a CSR of occurrences over all
    reads of the batch (s3_se_align's arrays); pairs with one or both reads without an occurrence included"""
    lib = _batch_lib()
    lib.s3_sam_unpaired_records.restype = C.c_int
    lib.s3_sam_unpaired_batch_text.restype = C.c_int
    rng = np.random.default_rng(4669)
    n, G, gen, cnames, keep = _batch_genome(rng)
    num, row = 600, 160
    lens, bases, quals, names, rd = _batch_reads(rng, num, row)
    counts = rng.choice([0, 1, 1, 2, 4, 7], num)
    off = np.zeros(num + 1, np.uint32)
    off[1:] = np.cumsum(counts)
    tot = int(off[-1])
    pos = rng.integers(0, n - 160, tot).astype(np.uint32)
    flags = np.ascontiguousarray(np.stack([rng.integers(1, 3, tot), rng.integers(0, 4, tot)], 1).astype(np.uint8))
    ids = np.sort(rng.choice(num // 2, 180, replace=False)).astype(np.uint32)
    for cfg, cap in ((Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA"), 1000), (Config(2, 1, 1, -2, 0, 40, 1, 1, 1, 1000, b"rgB"), 3), (Config(4, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgC"), 1000)):
        want = []
        for p in ids.tolist():
            arrs, ns = [], []
            for r in (2 * p, 2 * p + 1):
                a, b = int(off[r]), int(off[r + 1])
                arrs.append((Occurrence * max(b - a, 1))(*[Occurrence(int(pos[i]), int(flags[i][0]), int(flags[i][1])) for i in range(a, b)]))
                ns.append(b - a)
            out = (Record * 2)()
            assert lib.s3_sam_unpaired_records(C.byref(gen), C.byref(cfg), arrs[0], ns[0], arrs[1], ns[1], cap, bases[2 * p].ctypes.data_as(U8P), bases[2 * p + 1].ctypes.data_as(U8P),
                                               C.cast(quals[2 * p].ctypes.data, C.c_char_p), C.cast(quals[2 * p + 1].ctypes.data, C.c_char_p), int(lens[2 * p]), int(lens[2 * p + 1]),
                                               names[2 * p], names[2 * p + 1], out) == 0, lib.s3_last_error()
            want += [_line_of(lib, out[0], cnames), _line_of(lib, out[1], cnames)]
        want = b"".join(x + b"\n" for x in want)
        for threads in (1, 5, 0):
            text, size = C.c_void_p(), C.c_uint64()
            rc = lib.s3_sam_unpaired_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), helpers.u32p(off), helpers.u32p(pos), flags.ctypes.data_as(U8P),
                                                helpers.u32p(ids), C.c_uint64(len(ids)), cap, threads, C.byref(text), C.byref(size))
            assert rc == 0, lib.s3_last_error()
            got = C.string_at(text.value, size.value)
            lib.s3_free(text)
            assert got == want and got.count(b"\n") == 2 * len(ids)
    bad = ids.copy(); bad[0] = num // 2                                      # a pair beyond the batch
    text, size = C.c_void_p(), C.c_uint64()
    assert lib.s3_sam_unpaired_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), helpers.u32p(off), helpers.u32p(pos), flags.ctypes.data_as(U8P),
                                          helpers.u32p(bad), C.c_uint64(len(bad)), 1000, 2, C.byref(text), C.byref(size)) != 0


def test_sam_unpaired_dp_batch_text_builds_one_list_per_read():
    """s3_sam_unpaired_dp_batch_text == per named pair: each read's list as the reference's AllHits holds it (its single-read DP hits when it
    has any, else its occurrences from the search with score = len x match + mismatches x mismatch score and CIGAR <len>M, else nothing)
    -> s3_sam_unpaired_dp_records -> s3_sam_format_line"""
    lib = _batch_lib()
    lib.s3_sam_unpaired_dp_records.restype = C.c_int
    lib.s3_sam_unpaired_dp_batch_text.restype = C.c_int
    rng = np.random.default_rng(1729)
    n, G, gen, cnames, keep = _batch_genome(rng)
    num, row = 600, 160
    lens, bases, quals, names, rd = _batch_reads(rng, num, row)
    scores = api.DPScores(1, -2, -3, -1)
    kind = rng.choice([0, 1, 2], num, p=[0.2, 0.4, 0.4])                      # nothing / search occurrences / DP hits (a few of these also have occurrences: DP wins)
    counts = np.where(kind == 1, rng.choice([1, 1, 2, 5], num), np.where((kind == 2) & (rng.random(num) < 0.2), 1, 0))
    off = np.zeros(num + 1, np.uint32)
    off[1:] = np.cumsum(counts)
    tot = int(off[-1])
    near = rng.integers(1000, n - 3000, num // 2)
    pos = rng.integers(0, n - 400, tot).astype(np.uint32)
    flags = np.ascontiguousarray(np.stack([rng.integers(1, 3, tot), rng.integers(0, 4, tot)], 1).astype(np.uint8))
    hits, runs, per_read = [], [], {}
    for r in range(num):
        if kind[r] != 2:
            continue
        L = int(lens[r])
        for _ in range(int(rng.choice([1, 1, 2, 4]))):
            cg = random_special_cigar(rng, L)
            mine = _cigar_runs(cg)
            p = int(near[r // 2]) + int(rng.integers(-300, 300)) if rng.random() < 0.5 else int(rng.integers(0, n - 2 * L - 8))
            h = (r, p, int(rng.integers(int(0.3 * L), L + 1)), int(rng.integers(1, 3)), len(runs), len(mine), int(rng.integers(1, 3)), 0)
            runs += mine
            hits.append(h)
            per_read.setdefault(r, []).append(h)
    # mates of reads with DP hits often lie near them (insert sizes on one chromosome)
    for r in range(num):
        if kind[r] == 1 and rng.random() < 0.5:
            pos[int(off[r])] = int(near[r // 2]) + int(rng.integers(-300, 300))
    harr = np.array(hits, api.DP_HIT_DTYPE)
    rarr = np.array(runs, np.uint32)
    ids = np.sort(rng.choice(num // 2, 200, replace=False)).astype(np.uint32)
    cutoff = 30
    for cfg in (Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA"), Config(2, 1, 1, -2, 0, 40, 1, 1, 1, 1000, b"rgB"), Config(3, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgC")):
        want = []
        mixed = 0
        for p in ids.tolist():
            arrs, ns, held = [], [], []
            for r in (2 * p, 2 * p + 1):
                L = int(lens[r])
                if r in per_read:
                    hs = per_read[r]
                    a = (ReadAlignment * len(hs))()
                    for k, h in enumerate(hs):
                        cg, ed, _ = _decode(lib, rarr[h[4]:h[4] + h[5]], L, h[2], scores)
                        held.append(cg)
                        a[k].ambPosition, a[k].strand, a[k].score, a[k].editdist, a[k].isFromDP, a[k].cigar = h[1], h[6], h[2], ed, 1, cg
                    arrs.append(a); ns.append(len(hs))
                else:
                    lo, hi = int(off[r]), int(off[r + 1])
                    a = (ReadAlignment * max(hi - lo, 1))()
                    cg = b"%dM" % L
                    held.append(cg)
                    for k, i in enumerate(range(lo, hi)):
                        mm = int(flags[i][1])
                        a[k].ambPosition, a[k].strand, a[k].score, a[k].editdist, a[k].isFromDP, a[k].cigar = int(pos[i]), int(flags[i][0]), L * cfg.dpMatchScore + mm * cfg.dpMisMatchScore, mm, 0, cg
                    arrs.append(a); ns.append(hi - lo)
            mixed += (2 * p in per_read) != (2 * p + 1 in per_read)
            out = (Record * 2)()
            assert lib.s3_sam_unpaired_dp_records(C.byref(gen), C.byref(cfg), arrs[0], ns[0], arrs[1], ns[1], cutoff, bases[2 * p].ctypes.data_as(U8P), bases[2 * p + 1].ctypes.data_as(U8P),
                                                  C.cast(quals[2 * p].ctypes.data, C.c_char_p), C.cast(quals[2 * p + 1].ctypes.data, C.c_char_p), int(lens[2 * p]), int(lens[2 * p + 1]),
                                                  names[2 * p], names[2 * p + 1], out) == 0, lib.s3_last_error()
            want += [_line_of(lib, out[0], cnames), _line_of(lib, out[1], cnames)]
        want = b"".join(x + b"\n" for x in want)
        assert mixed > 40
        for threads in (1, 5, 0):
            text, size = C.c_void_p(), C.c_uint64()
            rc = lib.s3_sam_unpaired_dp_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), helpers.u32p(off), helpers.u32p(pos), flags.ctypes.data_as(U8P),
                                                   harr.ctypes.data_as(C.c_void_p), C.c_uint64(len(harr)), helpers.u32p(rarr), C.c_uint64(len(rarr)), scores, cutoff,
                                                   helpers.u32p(ids), C.c_uint64(len(ids)), threads, C.byref(text), C.byref(size))
            assert rc == 0, lib.s3_last_error()
            got = C.string_at(text.value, size.value)
            lib.s3_free(text)
            assert got == want and got.count(b"\n") == 2 * len(ids)
        g2, c2, r2 = _api_handles(n, keep, cfg, bases, quals, lens, num)
        assert api.sam_unpaired_dp_batch_text(g2, c2, r2, off, pos, flags, harr, rarr, scores, cutoff, ids, num_threads=3) == want
    # hits of a read that are not next to each other are refused
    bad = np.concatenate([harr, harr[:1]])
    text, size = C.c_void_p(), C.c_uint64()
    assert lib.s3_sam_unpaired_dp_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), helpers.u32p(off), helpers.u32p(pos), flags.ctypes.data_as(U8P),
                                             bad.ctypes.data_as(C.c_void_p), C.c_uint64(len(bad)), helpers.u32p(rarr), C.c_uint64(len(rarr)), scores, cutoff,
                                             helpers.u32p(ids), C.c_uint64(len(ids)), 2, C.byref(text), C.byref(size)) != 0


def test_api_mirror_of_the_batch_text_entries():
    """soap3dp_b200.api's wrappers of the batch entries give the text the raw C calls give (single reads and unpaired pairs; the other
    wrappers share their marshalling)"""
    lib = _batch_lib()
    rng = np.random.default_rng(999)
    n, G, gen, cnames, (pac, chr_end, amb, segs) = _batch_genome(rng)
    num, row = 300, 160
    lens, bases, quals, names, rd = _batch_reads(rng, num, row)
    counts = rng.choice([0, 1, 2, 5], num)
    off = np.zeros(num + 1, np.uint32)
    off[1:] = np.cumsum(counts)
    tot = int(off[-1])
    pos = rng.integers(0, n - 160, tot).astype(np.uint32)
    flags = np.ascontiguousarray(np.stack([rng.integers(1, 3, tot), rng.integers(0, 4, tot)], 1).astype(np.uint8))
    cfg = Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA")
    text, size = C.c_void_p(), C.c_uint64()
    assert lib.s3_sam_single_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), helpers.u32p(off), helpers.u32p(pos), flags.ctypes.data_as(U8P), 3,
                                        C.byref(text), C.byref(size)) == 0
    want_single = C.string_at(text.value, size.value)
    lib.s3_free(text)
    ids = np.arange(0, num // 2, 3, dtype=np.uint32)
    assert lib.s3_sam_unpaired_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), helpers.u32p(off), helpers.u32p(pos), flags.ctypes.data_as(U8P),
                                          helpers.u32p(ids), C.c_uint64(len(ids)), 1000, 3, C.byref(text), C.byref(size)) == 0
    want_unpaired = C.string_at(text.value, size.value)
    lib.s3_free(text)
    g2 = api.SamGenomeDesc(pac, n, [(0, 1, 0xFFFFFFFF), (70_000, 2, 70_000 - 1), (100_000, 2, 70_000 - 1 - 500), (150_000, 3, 150_000 - 1)], amb, chr_end, [b"chr1", "chrTwo", b"3"])
    c2 = api.SamConfig(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgA")
    r2 = api.SamReads(bases, quals, lens, [b"batch%d" % r for r in range(num)])
    assert api.sam_single_batch_text(g2, c2, r2, off, pos, flags, num_threads=2) == want_single
    assert api.sam_unpaired_batch_text(g2, c2, r2, off, pos, flags, ids) == want_unpaired
    with pytest.raises(api.S3Error):
        api.sam_paired_batch_text(g2, c2, r2, np.ones(num // 2, np.uint8), np.zeros(num // 2, api.PE_PAIR_DTYPE), None)      # the chain's statistics are required
    with pytest.raises(ValueError):
        api.SamReads(bases, quals[:, :10], lens, [b"x"] * num)


def test_sam_batch_text_edge_sizes():
    """more host threads than items, one item, no item: the same text as one thread gives; empty results give an empty text"""
    lib = _batch_lib()
    rng = np.random.default_rng(77)
    n, G, gen, cnames, keep = _batch_genome(rng)
    num, row = 6, 160
    lens, bases, quals, names, rd = _batch_reads(rng, num, row)
    off = np.array([0, 1, 1, 3, 3, 4, 6], np.uint32)
    pos = rng.integers(0, n - 160, 6).astype(np.uint32)
    flags = np.ascontiguousarray(np.stack([rng.integers(1, 3, 6), rng.integers(0, 3, 6)], 1).astype(np.uint8))
    g2, c2, r2 = _api_handles(n, keep, Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rgE"), bases, quals, lens, num)
    one = api.sam_single_batch_text(g2, c2, r2, off, pos, flags, num_threads=1)
    assert one.count(b"\n") == num
    for t in (2, 6, 7, 64):
        assert api.sam_single_batch_text(g2, c2, r2, off, pos, flags, num_threads=t) == one
    ids = np.array([1], np.uint32)
    a = api.sam_unpaired_batch_text(g2, c2, r2, off, pos, flags, ids, num_threads=1)
    assert a.count(b"\n") == 2 and api.sam_unpaired_batch_text(g2, c2, r2, off, pos, flags, ids, num_threads=32) == a
    scores = api.DPScores(1, -2, -3, -1)
    none_hits, none_runs = np.zeros(0, api.DP_HIT_DTYPE), np.zeros(0, np.uint32)
    assert api.sam_single_dp_batch_text(g2, c2, r2, none_hits, none_runs, scores, 30, num_threads=8) == b""
    assert api.sam_deep_dp_batch_text(g2, c2, r2, np.zeros(0, api.DEEP_HIT_DTYPE), none_runs, scores) == b""
    assert api.sam_pair_dp_batch_text(g2, c2, r2, np.zeros(0, api.PE_DP_DTYPE), none_runs, scores) == b""
    assert api.sam_unpaired_batch_text(g2, c2, r2, off, pos, flags, np.zeros(0, np.uint32)) == b""
    # without DP hits the after-DP writer of unpaired reads works from the occurrences alone
    b = api.sam_unpaired_dp_batch_text(g2, c2, r2, off, pos, flags, none_hits, none_runs, scores, 30, ids, num_threads=3)
    assert b.count(b"\n") == 2
