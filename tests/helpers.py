"""Shared test helpers: synthetic data + ctypes bindings of the oracles.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch oracle/.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

s3 = _pkg.load()
from soap3dp_b200 import fmindex, formats  # noqa: E402

U32P = C.POINTER(C.c_uint32)


def u32p(a):
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    return a.ctypes.data_as(U32P)


def load_oracle():
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    lib = C.CDLL(path)
    lib.s3o_search_launch.restype = C.c_ulonglong
    lib.s3o_search_launch.argtypes = [C.c_uint32, U32P, U32P, C.c_uint32, C.c_uint32,
                                      U32P, U32P, C.c_uint32, U32P, U32P, C.c_uint32, C.c_uint32,
                                      U32P, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    lib.s3o_rank.restype = C.c_uint32
    lib.s3o_rank.argtypes = [U32P, U32P, C.c_uint32, C.c_int, C.c_uint32]
    return lib


def load_ref_search():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_search.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_search_launch.restype = C.c_ulonglong
    lib.ref_search_launch.argtypes = [C.c_uint32, U32P, U32P, C.c_uint32, C.c_uint32,
                                      U32P, U32P, C.c_uint32, U32P, U32P, C.c_uint32, C.c_uint32,
                                      U32P, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.c_int, C.c_int]
    lib.ref_rank.restype = C.c_uint32
    lib.ref_rank.argtypes = [U32P, U32P, C.c_uint32, C.c_int, C.c_uint32]
    return lib


class HostIndex:
    """numpy uint32 views of a Soap3IndexArrays for ctypes calls."""

    def __init__(self, idx):
        self.idx = idx
        self.n = idx.text_length
        self.bwt = idx.fwd.bwt_words.cpu().numpy().view(np.uint32)
        self.occ = idx.fwd.occ.cpu().numpy().view(np.uint32)
        self.rbwt = idx.rev.bwt_words.cpu().numpy().view(np.uint32)
        self.rocc = idx.rev.occ.cpu().numpy().view(np.uint32)
        self.isa0 = idx.fwd.inverse_sa0
        self.risa0 = idx.rev.inverse_sa0


def oracle_launch(lib, hi, case, queries, lengths, n, wpq, answers, is_bad, rnd, k, sa_allowed, wpa, exact=0):
    return lib.s3o_search_launch(case, u32p(queries), u32p(lengths), n, wpq, u32p(hi.bwt), u32p(hi.occ), hi.isa0,
                                 u32p(hi.rbwt), u32p(hi.rocc), hi.risa0, hi.n, u32p(answers),
                                 is_bad.ctypes.data_as(C.POINTER(C.c_uint8)), rnd, k, sa_allowed, wpa, exact)


def ref_launch(lib, hi, case, queries, lengths, n, wpq, answers, is_bad, rnd, k, sa_allowed, wpa, exact=0, nthreads=0):
    return lib.ref_search_launch(case, u32p(queries), u32p(lengths), n, wpq, u32p(hi.bwt), u32p(hi.occ), hi.isa0,
                                 u32p(hi.rbwt), u32p(hi.rocc), hi.risa0, hi.n, u32p(answers),
                                 is_bad.ctypes.data_as(C.POINTER(C.c_uint8)), rnd, k, sa_allowed, wpa, exact, nthreads)


# ---------------------------------------------------------------------------
# DP
# ---------------------------------------------------------------------------
I32P = C.POINTER(C.c_int32)
U8P = C.POINTER(C.c_uint8)

_DP_ARGS = [U32P, U32P, C.c_uint32, U32P, U32P, C.c_uint32, I32P, I32P, U32P, U32P, U8P, C.c_uint32,
            U32P, U32P, U32P, U32P, C.c_int, C.c_int, C.c_int, C.c_int]


def load_oracle_dp():
    lib = load_oracle()
    lib.s3o_dp_align.restype = C.c_ulonglong
    lib.s3o_dp_align.argtypes = _DP_ARGS
    return lib


def load_ref_dp():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_dp.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_dp_align.restype = C.c_int
    lib.ref_dp_align.argtypes = [U32P, U32P, C.c_uint32, C.c_uint32, U32P, U32P, C.c_uint32, I32P, I32P, U32P, U32P, U8P,
                                 C.c_uint32, U32P, U32P, U32P, U32P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    return lib


def _opt(a):
    return u32p(a) if a is not None else None


class DPBatch:
    """A batch in the reference's SemiGlobalAligner host format."""

    def __init__(self, dna, dna_len, read, read_len, max_dna, max_read, cutoff, clip_lt=None, clip_rt=None,
                 anchor_l=None, anchor_r=None):
        self.n = len(dna_len)
        self.max_dna, self.max_read = max_dna, max_read
        self.dna = formats.pack_dp_sequences(dna, max_dna)
        self.read = formats.pack_dp_sequences(read, max_read)
        up = formats.ceil32(self.n)

        def padu(a):
            if a is None:
                return None
            o = np.zeros(up, np.uint32)
            o[:self.n] = a
            return o
        self.dna_len, self.read_len = padu(dna_len), padu(read_len)
        self.cutoff = np.zeros(up, np.int32)
        self.cutoff[:self.n] = cutoff
        self.clip_lt, self.clip_rt, self.anchor_l, self.anchor_r = padu(clip_lt), padu(clip_rt), padu(anchor_l), padu(anchor_r)
        self.pat_len = max_read + max_dna

    def outputs(self):
        up = formats.ceil32(self.n)
        return (np.zeros(up, np.int32), np.zeros(up, np.uint32), np.zeros(up, np.uint32),
                np.zeros(up * self.pat_len, np.uint8))


def oracle_dp(lib, b, scores=(1, -2, -3, -1)):
    sc, hit, cnt, pat = b.outputs()
    cells = lib.s3o_dp_align(u32p(b.dna), u32p(b.dna_len), b.max_dna, u32p(b.read), u32p(b.read_len), b.max_read,
                             b.cutoff.ctypes.data_as(I32P), sc.ctypes.data_as(I32P), u32p(hit), u32p(cnt),
                             pat.ctypes.data_as(U8P), b.n, _opt(b.clip_lt), _opt(b.clip_rt), _opt(b.anchor_l),
                             _opt(b.anchor_r), *scores)
    return sc, hit, cnt, pat, cells


def ref_dp(lib, b, scores=(1, -2, -3, -1), nthreads=0):
    sc, hit, cnt, pat = b.outputs()
    lib.ref_dp_align(u32p(b.dna), u32p(b.dna_len), b.max_dna, b.max_dna, u32p(b.read), u32p(b.read_len), b.max_read,
                     b.cutoff.ctypes.data_as(I32P), sc.ctypes.data_as(I32P), u32p(hit), u32p(cnt),
                     pat.ctypes.data_as(U8P), b.n, _opt(b.clip_lt), _opt(b.clip_rt), _opt(b.anchor_l),
                     _opt(b.anchor_r), *scores, 1, nthreads)
    return sc, hit, cnt, pat


def load_ref_dp_cuda():
    """the reference's DP kernels compiled for sm_100a (oracle/build_ref.sh); None when not built"""
    path = os.path.join(ROOT, "oracle", "_ref", "libref_dp_cuda.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_dp_cuda_align.restype = C.c_int
    lib.ref_dp_cuda_align.argtypes = [U32P, U32P, C.c_uint32, C.c_uint32, U32P, U32P, C.c_uint32, I32P, I32P, U32P, U32P, U8P,
                                      C.c_uint32, U32P, U32P, U32P, U32P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_float)]
    return lib


def ref_dp_cuda(lib, b, scores=(1, -2, -3, -1), num_blocks=64):
    """-> ((scores, hitLocs, counts, pattern), kernel milliseconds): the reference's kernels on the GPU, launched
    numOfBlocks x 128 alignments at a time like SemiGlobalAligner::performAlignment (DV-DPfunctions.cu:669-725)"""
    sc, hit, cnt, pat = b.outputs()
    ms = C.c_float(0)
    rc = lib.ref_dp_cuda_align(u32p(b.dna), u32p(b.dna_len), b.max_dna, b.max_dna, u32p(b.read), u32p(b.read_len), b.max_read,
                               b.cutoff.ctypes.data_as(I32P), sc.ctypes.data_as(I32P), u32p(hit), u32p(cnt),
                               pat.ctypes.data_as(U8P), b.n, _opt(b.clip_lt), _opt(b.clip_rt), _opt(b.anchor_l),
                               _opt(b.anchor_r), *scores, num_blocks, C.byref(ms))
    if rc != 0:
        raise RuntimeError("ref_dp_cuda_align failed")
    return (sc, hit, cnt, pat), float(ms.value)


def load_ref_search_cuda():
    """the reference's search kernels compiled for sm_100a (oracle/build_ref_search_cuda.sh); None when not built"""
    path = os.path.join(ROOT, "oracle", "_ref", "libref_search_cuda.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_search_cuda_upload.restype = C.c_int
    lib.ref_search_cuda_upload.argtypes = [U32P, U32P, C.c_size_t, U32P, U32P, C.c_size_t]
    lib.ref_search_cuda_free.restype = None
    lib.ref_search_cuda_round1.restype = C.c_int
    lib.ref_search_cuda_round1.argtypes = [U32P, U32P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                           C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                           C.POINTER(C.c_float)]
    return lib


def ref_search_cuda_round1(lib, hi, queries, lengths, n, wpq, k, sa_allowed, wpa, max_launch=1 << 20):
    """-> (answers per case, kernel milliseconds): perform_round1_alignment with the reference's kernels on the GPU,
    at most 1,048,576 reads per launch (NUM_BLOCKS * THREADS_PER_BLOCK, definitions.h:75-77).  The index must have been
    uploaded with lib.ref_search_cuda_upload."""
    ncases = formats.NUM_CASES[k]
    up = formats.ceil32(n)
    answers = [np.zeros(up * wpa, np.uint32) for _ in range(ncases)]
    total = 0.0
    for first in range(0, n, max_launch):
        cnt = min(max_launch, n - first)
        ptrs = (C.c_void_p * ncases)(*[a[first * wpa:].ctypes.data for a in answers])
        ms = C.c_float(0)
        rc = lib.ref_search_cuda_round1(u32p(queries[first * wpq:]), u32p(lengths[first:]), cnt, wpq, hi.isa0, hi.risa0, hi.n,
                                        ptrs, k, ncases, sa_allowed, wpa, 0, C.byref(ms))
        if rc != 0:
            raise RuntimeError("ref_search_cuda_round1 failed")
        total += float(ms.value)
    return answers, total


def pattern_end(w):
    """index of the 0 terminator; a count byte after 'V' may legitimately be 0"""
    i = 0
    while True:
        if w[i] == ord('V'):
            i += 2
            continue
        if w[i] == 0:
            return i
        i += 1


def compare_dp(b, got, want, what=""):
    """scores / hitLocs / counts for every alignment; pattern bytes (through the 0
    terminator) for those reaching the cutoff -- the only ones the reference writes."""
    gs, gh, gc, gp = got[:4]
    ws, wh, wc, wp = want[:4]
    n = b.n
    assert np.array_equal(gs[:n], ws[:n]), f"{what} scores differ at {np.nonzero(gs[:n] != ws[:n])[0][:5]}"
    assert np.array_equal(gc[:n], wc[:n]), f"{what} maxScoreCounts differ at {np.nonzero(gc[:n] != wc[:n])[0][:5]}"
    assert np.array_equal(gh[:n], wh[:n]), f"{what} hitLocs differ at {np.nonzero(gh[:n] != wh[:n])[0][:5]}"
    npass = 0
    for t in range(n):
        if ws[t] >= b.cutoff[t]:
            npass += 1
            w = wp[t * b.pat_len:(t + 1) * b.pat_len]
            g = gp[t * b.pat_len:(t + 1) * b.pat_len]
            i = pattern_end(w)
            assert np.array_equal(g[:i + 1], w[:i + 1]), \
                f"{what} pattern differs for alignment {t}: {bytes(g[:i + 1])} vs {bytes(w[:i + 1])}"
    return npass


def make_dp_batch(genome, n, read_len, mode, seed, indel_rate=0.004, sub_rate=0.02, insert=(200, 500), max_read=None):
    """Synthetic DP batches shaped like the engines' packers:
    mode 'single' : window = read +- margin, clips by strand, no anchors (DV-DPfunctions.cu:1425)
    mode 'rescue' : mate-rescue windows with anchors (DV-DPfunctions.cu:2027)."""
    from soap3dp_b200 import synth
    G = genome
    rs = synth.simulate_single_end(G, n, read_len, seed=seed, sub_rate=sub_rate, indel_rate=indel_rate, margin=1000)
    rng = np.random.default_rng(seed)
    reads = rs.reads.numpy()
    pos = rs.pos.numpy()
    strand = rs.strand.numpy()
    # DP aligns the read in reference orientation (reverse-strand reads are
    # reverse-complemented by the packers, DV-DPfunctions.cu:1497-1505)
    fw = np.where(strand[:, None] == 1, 3 - reads[:, ::-1], reads)
    Gn = G.numpy()
    if max_read is None:
        max_read = (read_len // 4 + 1) * 4
    if mode == "single":
        margin = read_len // 4 if read_len > 100 else 25
        wlen = read_len + 2 * margin
        max_dna = max_read + 2 * margin + 8
        start = pos - margin + rng.integers(-5, 6, n)
        dna = Gn[start[:, None] + np.arange(wlen)[None, :]]
        dna_len = np.full(n, wlen, np.uint32)
        clip_lt = np.where(strand == 0, 3, 8).astype(np.uint32)
        clip_rt = np.where(strand == 0, 8, 3).astype(np.uint32)
        anchor_l = anchor_r = None
    else:
        lo, hi = insert
        wlen = hi - lo + read_len
        max_dna = hi - lo + max_read + 1
        left_side = rng.integers(0, 2, n).astype(bool)
        off = rng.integers(0, hi - lo, n)
        start = pos - off
        dna = Gn[start[:, None] + np.arange(wlen)[None, :]]
        dna_len = np.full(n, wlen, np.uint32)
        clip_lt = rng.integers(0, 9, n).astype(np.uint32)
        clip_rt = rng.integers(0, 9, n).astype(np.uint32)
        anchor_l = np.where(left_side, max_dna, hi - lo + 1).astype(np.uint32)
        anchor_r = np.where(left_side, read_len, 0).astype(np.uint32)
    rl = np.full(n, read_len, np.uint32)
    rl[::9] = read_len - 3
    dna_len[::13] -= 7
    cutoff = np.ceil(0.3 * rl).astype(np.int32)
    return DPBatch(dna.astype(np.uint8), dna_len, fw.astype(np.uint8), rl, max_dna, max_read, cutoff,
                   clip_lt, clip_rt, anchor_l, anchor_r)


# ---- DP result decoding (pattern bytes -> CIGAR, edit distance) ----------------------------------------------------
def load_decode_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import decode_oracle
    return decode_oracle


def load_ref_decode():
    """the reference's CigarStringEncoder + result loop + convertToCigarStr (oracle/build_ref.sh); None when not built"""
    path = os.path.join(ROOT, "oracle", "_ref", "libref_decode.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_dp_decode.restype = C.c_int
    lib.ref_dp_decode.argtypes = [U8P, C.c_int, I32P, U32P, U32P, U32P, U32P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, I32P, U32P, I32P, U32P, C.c_char_p, C.c_size_t]
    lib.ref_convert_cigar.restype = C.c_int
    lib.ref_convert_cigar.argtypes = [C.c_char_p, C.c_char_p]
    return lib


def ref_decode(lib, pattern, pat_len, scores, hit_locs, lengths, positions, counts, cutoff, scores4):
    """-> list of (index, algnmt, special cigar, sam cigar, editdist, num_sameScore) for the alignments the reference keeps"""
    n = len(scores)
    idx, alg, ed, same = np.zeros(n, np.int32), np.zeros(n, np.uint32), np.zeros(n, np.int32), np.zeros(n, np.uint32)
    cap = n * 1100 + 16
    buf = C.create_string_buffer(cap)
    m = lib.ref_dp_decode(np.ascontiguousarray(pattern, np.uint8).ctypes.data_as(U8P), pat_len,
                          np.ascontiguousarray(scores, np.int32).ctypes.data_as(I32P), u32p(np.ascontiguousarray(hit_locs, np.uint32)),
                          u32p(np.ascontiguousarray(lengths, np.uint32)), u32p(np.ascontiguousarray(positions, np.uint32)),
                          u32p(np.ascontiguousarray(counts, np.uint32)), n, cutoff, *scores4,
                          idx.ctypes.data_as(I32P), u32p(alg), ed.ctypes.data_as(I32P), u32p(same), buf, cap)
    assert m >= 0
    cigars = buf.value.decode("ascii").split("\n")[:m]
    out = []
    sam = C.create_string_buffer(4096)
    for k in range(m):
        lib.ref_convert_cigar(cigars[k].encode("ascii"), sam)
        out.append((int(idx[k]), int(alg[k]), cigars[k], sam.value.decode("ascii"), int(ed[k]), int(same[k])))
    return out


def synthetic_patterns(rng, n, pat_len, max_ops=40):
    """pattern bytes as GPUBacktrack may write them and some it never writes (a count byte of 0 or 255, an escape
    first, deletions at either end): right-to-left op bytes, 'V',count escapes, 0-terminated"""
    pat = np.zeros(n * pat_len, np.uint8)
    for t in range(n):
        out = []
        if rng.random() < 0.5:
            out += [ord('S'), ord('V'), int(rng.integers(0, 12))]          # the right clip is always written this way
        elif rng.random() < 0.05:
            out += [ord('V'), int(rng.integers(0, 5))]
        for _ in range(int(rng.integers(1, max_ops))):
            op = "MMMMMMMMmmIDS"[int(rng.integers(0, 13))]
            out.append(ord(op))
            r = rng.random()
            if r < 0.25:
                out += [ord('V'), int(rng.integers(0, 9))]
            elif r < 0.27:
                out += [ord('V'), 255]
        out = out[:pat_len - 3]
        if out and out[-1] == ord('V'):
            out.append(2)
        pat[t * pat_len:t * pat_len + len(out)] = out
    return pat


# ---- paired-end pairing of the two reads' occurrence lists (PEAlgnmt.cpp) ------------------------------------------
U64P = C.POINTER(C.c_uint64)


def make_occurrence_lists(rng, num_pairs, max_occ=12, span=1 << 20, near_edges=False):
    """per read pair two occurrence lists in arrival order: clusters of a left and a right hit an insert apart,
    duplicates of a position, both strands, 0..4 mismatches, empty and single-element lists"""
    p1, s1, m1, o1, p2, s2, m2, o2 = [], [], [], [0], [], [], [], [0]
    for _ in range(num_pairs):
        n1, n2 = int(rng.integers(0, max_occ)), int(rng.integers(0, max_occ))
        base = int(rng.integers(0, 600)) if near_edges and rng.random() < 0.5 else \
            (0xFFFFFFFF - int(rng.integers(0, 900)) if near_edges else int(rng.integers(1000, span)))
        a = (base + rng.integers(-300, 300, n1)) & 0xFFFFFFFF
        b = (base + rng.integers(-700, 700, n2)) & 0xFFFFFFFF
        if n1 > 2 and rng.random() < 0.5:
            a[1] = a[0]
        if n2 > 2 and rng.random() < 0.5:
            b[2] = b[0]
        p1 += list(a); p2 += list(b)
        s1 += list(rng.integers(1, 3, n1)); s2 += list(rng.integers(1, 3, n2))
        m1 += list(rng.integers(0, 5, n1)); m2 += list(rng.integers(0, 5, n2))
        o1.append(len(p1)); o2.append(len(p2))
    f = lambda x, t: np.ascontiguousarray(np.array(x, dtype=np.int64).astype(t))
    return (f(p1, np.uint32), f(s1, np.uint8), f(m1, np.uint8), f(o1, np.uint64),
            f(p2, np.uint32), f(s2, np.uint8), f(m2, np.uint8), f(o2, np.uint64))


def oracle_pair_occurrences(lists, pattern_lengths, lbound, ubound, left_leg, right_leg, report_one):
    """oracle/pair_oracle.c -> dict(offsets, pos1, pos2, insertion, flags[n,4], optimal, suboptimal, stats[numPairs,32])"""
    lib = load_oracle()
    U8 = C.POINTER(C.c_uint8)
    lib.s3o_pair_occurrences.restype = C.c_uint64
    lib.s3o_pair_occurrences.argtypes = [U32P, U8, U8, U64P, U32P, U8, U8, U64P, U32P, C.c_uint64, C.c_int32, C.c_int32, C.c_int, C.c_int,
                                         C.c_int, U64P, U32P, U32P, U32P, U8, C.c_uint64, U32P, U32P, U32P]
    p1, s1, m1, o1, p2, s2, m2, o2 = lists
    npairs = len(o1) - 1
    pl = np.ascontiguousarray(pattern_lengths, np.uint32)
    offs = np.zeros(npairs + 1, np.uint64)
    opt, sub, stats = np.zeros(npairs, np.uint32), np.zeros(npairs, np.uint32), np.zeros((npairs, 32), np.uint32)
    args = (u32p(p1), s1.ctypes.data_as(U8), m1.ctypes.data_as(U8), o1.ctypes.data_as(U64P), u32p(p2), s2.ctypes.data_as(U8),
            m2.ctypes.data_as(U8), o2.ctypes.data_as(U64P), u32p(pl), npairs, lbound, ubound, left_leg, right_leg, int(report_one))
    total = lib.s3o_pair_occurrences(*args, offs.ctypes.data_as(U64P), None, None, None, None, 0, None, None, None)
    a, b, ins, fl = np.zeros(total, np.uint32), np.zeros(total, np.uint32), np.zeros(total, np.uint32), np.zeros((total, 4), np.uint8)
    lib.s3o_pair_occurrences(*args, offs.ctypes.data_as(U64P), u32p(a), u32p(b), u32p(ins), fl.ctypes.data_as(U8), total,
                             u32p(opt), u32p(sub), u32p(stats))
    return dict(offsets=offs, pos1=a, pos2=b, insertion=ins, flags=fl, optimal=opt, suboptimal=sub, stats=stats)


def load_ref_pair():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_pair.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    U8 = C.POINTER(C.c_uint8)
    I = C.POINTER(C.c_int)
    lib.ref_pair_occurrences.restype = C.c_int
    lib.ref_pair_occurrences.argtypes = [U32P, U8, U8, C.c_uint, U32P, U8, U8, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         U32P, U32P, U32P, U8, C.c_uint, I, I, U32P]
    return lib


def ref_pair_occurrences(lib, lists, pattern_lengths, lbound, ubound, left_leg, right_leg, report_one):
    """the reference's PEMappingOccurrences + PEStatsPEOutput, read pair by read pair -> the oracle's dict"""
    U8 = C.POINTER(C.c_uint8)
    p1, s1, m1, o1, p2, s2, m2, o2 = lists
    npairs = len(o1) - 1
    offs = np.zeros(npairs + 1, np.uint64)
    opt, sub, stats = np.zeros(npairs, np.uint32), np.zeros(npairs, np.uint32), np.zeros((npairs, 32), np.uint32)
    A, B, INS, FL = [], [], [], []
    for p in range(npairs):
        a0, a1, b0, b1 = int(o1[p]), int(o1[p + 1]), int(o2[p]), int(o2[p + 1])
        cap = (a1 - a0) * (b1 - b0) * 2 + 1
        a, b, ins, fl = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros((cap, 4), np.uint8)
        x1, y1, z1 = (np.ascontiguousarray(v[a0:a1]) for v in (p1, s1, m1))
        x2, y2, z2 = (np.ascontiguousarray(v[b0:b1]) for v in (p2, s2, m2))
        io, isub = C.c_int(-1), C.c_int(-1)
        st = np.zeros(30, np.uint32)
        n = lib.ref_pair_occurrences(u32p(x1), y1.ctypes.data_as(U8), z1.ctypes.data_as(U8), a1 - a0, u32p(x2), y2.ctypes.data_as(U8),
                                     z2.ctypes.data_as(U8), b1 - b0, int(pattern_lengths[p]), lbound, ubound, left_leg, right_leg,
                                     1 if report_one else 0, u32p(a), u32p(b), u32p(ins), fl.ctypes.data_as(U8), cap,
                                     C.byref(io), C.byref(isub), u32p(st))
        assert n <= cap
        A.append(a[:n]); B.append(b[:n]); INS.append(ins[:n]); FL.append(fl[:n])
        offs[p + 1] = offs[p] + n
        opt[p], sub[p] = io.value & 0xFFFFFFFF, isub.value & 0xFFFFFFFF
        stats[p, :30] = st
    cat = lambda v, shape: np.concatenate(v) if v else np.zeros(shape, np.uint32)
    return dict(offsets=offs, pos1=cat(A, 0), pos2=cat(B, 0), insertion=cat(INS, 0),
                flags=np.concatenate(FL) if FL else np.zeros((0, 4), np.uint8), optimal=opt, suboptimal=sub, stats=stats)


def same_pairing(x, y):
    return all(np.array_equal(x[k], y[k]) for k in ("offsets", "pos1", "pos2", "insertion", "flags", "optimal", "suboptimal", "stats"))


# ---- best-hit filters on a read's SA-range list and occurrence list (SAList.cpp retainAllBest family) ---------------
def make_hit_lists(rng, num_reads, max_sa=7, max_occ=7, high_counts=False):
    """per read an SA-range list and an occurrence list in arrival order; either may be empty, mismatch counts 0..4 with
    the minimum on either side or on both, ranges of 1..40 suffixes"""
    n_sa, n_occ = rng.integers(0, max_sa, num_reads), rng.integers(0, max_occ, num_reads)
    sa_off, occ_off = np.zeros(num_reads + 1, np.uint64), np.zeros(num_reads + 1, np.uint64)
    sa_off[1:], occ_off[1:] = np.cumsum(n_sa), np.cumsum(n_occ)
    ts, to = int(sa_off[-1]), int(occ_off[-1])
    sa_l = rng.integers(0, 1 << 31, ts).astype(np.uint32)
    sa_r = (sa_l + rng.integers(0, 40, ts)).astype(np.uint32)
    occ_m = rng.integers(0, 5, to).astype(np.uint8)
    if high_counts:                                        # DP-derived lists may carry other values in the field: counts >= 128 are large, not negative
        sel = rng.random(to) < 0.3
        occ_m[sel] = rng.integers(126, 256, int(sel.sum())).astype(np.uint8)
    return (sa_l, sa_r, rng.integers(1, 3, ts).astype(np.uint8), rng.integers(0, 5, ts).astype(np.uint8), sa_off,
            rng.integers(0, 1 << 32, to).astype(np.uint32), rng.integers(1, 3, to).astype(np.uint8), occ_m, occ_off)


_RETAIN_ARGS = None


def _retain_argtypes():
    U8 = C.POINTER(C.c_uint8)
    return [C.c_int, C.c_int32, U32P, U32P, U8, U8, U64P, U32P, U8, U8, U64P, C.c_uint64,
            U64P, U32P, U32P, U8, U64P, U32P, U8, U32P]


def run_retain(fn, lists, mode, max_num):
    """fn = s3o_retain_best or the harness twin -> dict(sa_off, sa_l, sa_r, sa_flags, occ_off, occ_pos, occ_flags, num)"""
    U8 = C.POINTER(C.c_uint8)
    fn.restype = None
    fn.argtypes = _retain_argtypes()
    sa_l, sa_r, sa_s, sa_m, sa_off, occ_p, occ_s, occ_m, occ_off = lists
    n = len(sa_off) - 1
    o_sa_off, o_occ_off = np.zeros(n + 1, np.uint64), np.zeros(n + 1, np.uint64)
    o_l, o_r, o_sf = np.zeros(len(sa_l), np.uint32), np.zeros(len(sa_l), np.uint32), np.zeros((len(sa_l), 2), np.uint8)
    o_p, o_of = np.zeros(len(occ_p), np.uint32), np.zeros((len(occ_p), 2), np.uint8)
    num = np.zeros(n, np.uint32)
    b8 = lambda x: x.ctypes.data_as(U8)
    fn(mode, max_num, u32p(sa_l), u32p(sa_r), b8(sa_s), b8(sa_m), sa_off.ctypes.data_as(U64P), u32p(occ_p), b8(occ_s), b8(occ_m),
       occ_off.ctypes.data_as(U64P), n, o_sa_off.ctypes.data_as(U64P), u32p(o_l), u32p(o_r), b8(o_sf), o_occ_off.ctypes.data_as(U64P),
       u32p(o_p), b8(o_of), u32p(num))
    ks, ko = int(o_sa_off[-1]), int(o_occ_off[-1])
    return dict(sa_off=o_sa_off, sa_l=o_l[:ks], sa_r=o_r[:ks], sa_flags=o_sf[:ks], occ_off=o_occ_off, occ_pos=o_p[:ko], occ_flags=o_of[:ko], num=num)


def oracle_retain_best(lists, mode, max_num=0):
    return run_retain(load_oracle().s3o_retain_best, lists, mode, max_num)


def load_ref_retain():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_retain.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    U8 = C.POINTER(C.c_uint8)
    lib.ref_retain_best.restype = C.c_uint
    lib.ref_retain_best.argtypes = [C.c_int, C.c_int, U32P, U32P, U8, U8, C.c_uint, U32P, U8, U8, C.c_uint, U32P, U32P, U8, U32P, U32P, U8, U32P]
    return lib


def ref_retain_best(lib, lists, mode, max_num=0):
    """the reference's functions read by read -> the oracle's dict"""
    U8 = C.POINTER(C.c_uint8)
    sa_l, sa_r, sa_s, sa_m, sa_off, occ_p, occ_s, occ_m, occ_off = lists
    n = len(sa_off) - 1
    out = dict(sa_off=np.zeros(n + 1, np.uint64), occ_off=np.zeros(n + 1, np.uint64), num=np.zeros(n, np.uint32))
    L, R, SF, P, OF = [], [], [], [], []
    b8 = lambda x: x.ctypes.data_as(U8)
    for r in range(n):
        s0, s1, o0, o1 = int(sa_off[r]), int(sa_off[r + 1]), int(occ_off[r]), int(occ_off[r + 1])
        a = [np.ascontiguousarray(v[s0:s1]) for v in (sa_l, sa_r, sa_s, sa_m)]
        b = [np.ascontiguousarray(v[o0:o1]) for v in (occ_p, occ_s, occ_m)]
        ol, orr, osf = np.zeros(s1 - s0 + 1, np.uint32), np.zeros(s1 - s0 + 1, np.uint32), np.zeros((s1 - s0 + 1, 2), np.uint8)
        op, oof = np.zeros(o1 - o0 + 1, np.uint32), np.zeros((o1 - o0 + 1, 2), np.uint8)
        ks, ko = C.c_uint32(0), C.c_uint32(0)
        out["num"][r] = lib.ref_retain_best(mode, max_num, u32p(a[0]), u32p(a[1]), b8(a[2]), b8(a[3]), s1 - s0, u32p(b[0]), b8(b[1]), b8(b[2]), o1 - o0,
                                            u32p(ol), u32p(orr), b8(osf), C.byref(ks), u32p(op), b8(oof), C.byref(ko))
        L.append(ol[:ks.value]); R.append(orr[:ks.value]); SF.append(osf[:ks.value]); P.append(op[:ko.value]); OF.append(oof[:ko.value])
        out["sa_off"][r + 1] = out["sa_off"][r] + ks.value
        out["occ_off"][r + 1] = out["occ_off"][r] + ko.value
    out.update(sa_l=np.concatenate(L), sa_r=np.concatenate(R), sa_flags=np.concatenate(SF), occ_pos=np.concatenate(P), occ_flags=np.concatenate(OF))
    return out


def same_retained(x, y):
    return all(np.array_equal(x[k], y[k]) for k in ("sa_off", "sa_l", "sa_r", "sa_flags", "occ_off", "occ_pos", "occ_flags", "num"))


# ---- which windows the DP engines align (oracle/window_oracle.c, oracle/_ref/libref_windows.so) -----------------------
def oracle_windows_single(lib, rid, pos, strand, lens, text, clip_l, clip_r):
    """-> [n, 4] start, length, clipLt, clipRt"""
    lib.s3o_window_single.restype = C.c_int
    lib.s3o_window_single.argtypes = [C.c_uint32, C.c_uint32, C.c_int, U32P, C.c_uint32, C.c_int, C.c_int, U32P, U32P, U32P, U32P]
    out = np.zeros((len(rid), 4), np.uint32)
    v = [C.c_uint32() for _ in range(4)]
    for c in range(len(rid)):
        lib.s3o_window_single(int(rid[c]), int(pos[c]), int(strand[c]), u32p(lens), text, clip_l, clip_r, *[C.byref(x) for x in v])
        out[c] = [x.value for x in v]
    return out


def oracle_windows_half(lib, rid, pos, strand, lens, text, P):
    """-> candidate index per window, [m, 9] leftOrRight, start, length, readLen, dpStrand, clipLt, clipRt, ancL, ancR"""
    I32 = C.POINTER(C.c_int)
    lib.s3o_window_half.restype = C.c_int
    lib.s3o_window_half.argtypes = [C.c_uint32, C.c_uint32, C.c_int, U32P, C.c_uint32] + [C.c_int] * 7 + [I32, U32P, U32P, U32P, I32, U32P, U32P, U32P, U32P]
    lor, dps = (C.c_int * 2)(), (C.c_int * 2)()
    a = [(C.c_uint32 * 2)() for _ in range(7)]
    cand, rows = [], []
    for c in range(len(rid)):
        k = lib.s3o_window_half(int(rid[c]), int(pos[c]), int(strand[c]), u32p(lens), text, P["left"], P["right"], P["ins_high"], P["ins_low"], P["max_dna"],
                                P["clip_l"], P["clip_r"], lor, a[0], a[1], a[2], dps, a[3], a[4], a[5], a[6])
        for j in range(k):
            cand.append(c)
            rows.append([lor[j], a[0][j], a[1][j], a[2][j], dps[j], a[3][j], a[4][j], a[5][j], a[6][j]])
    return np.array(cand, np.uint32), (np.array(rows, np.uint32) if rows else np.zeros((0, 9), np.uint32))


def oracle_windows_pair_left(lib, rid, pos, lens, text, P):
    """-> [n, 6] start, length, clipLt, clipRt, ancL, ancR"""
    lib.s3o_window_pair_left.restype = None
    lib.s3o_window_pair_left.argtypes = [C.c_uint32, C.c_uint32, U32P, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int] + [U32P] * 6
    v = [C.c_uint32() for _ in range(6)]
    out = np.zeros((len(rid), 6), np.uint32)
    for c in range(len(rid)):
        lib.s3o_window_pair_left(int(rid[c]), int(pos[c]), u32p(lens), text, P["left"], P["max_dna"], P["clip_l"], P["clip_r"], *[C.byref(x) for x in v])
        out[c] = [x.value for x in v]
    return out


def oracle_windows_pair_right(lib, rid, pos2, lstart, lhit, lens, text, P):
    """-> [n, 7] right read id, start, length, clipLt, clipRt, ancL, ancR"""
    lib.s3o_window_pair_right.restype = None
    lib.s3o_window_pair_right.argtypes = [C.c_uint32] * 4 + [U32P, C.c_uint32] + [C.c_int] * 6 + [U32P] * 7
    v = [C.c_uint32() for _ in range(7)]
    out = np.zeros((len(rid), 7), np.uint32)
    for c in range(len(rid)):
        lib.s3o_window_pair_right(int(rid[c]), int(pos2[c]), int(lstart[c]), int(lhit[c]), u32p(lens), text, P["right"], P["ins_high"], P["ins_low"],
                                  P["max_dna"], P["clip_l"], P["clip_r"], *[C.byref(x) for x in v])
        out[c] = [x.value for x in v]
    return out


def load_ref_windows():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_windows.so")
    return C.CDLL(path) if os.path.exists(path) else None


def ref_windows_single(lib, rid, pos, strand, lens, text, clip_l, clip_r, cutoff):
    n = len(rid)
    I32 = C.POINTER(C.c_int)
    st = np.ascontiguousarray(strand, np.int32)
    o = [np.zeros(n, np.uint32) for _ in range(4)]
    oc = np.zeros(n, np.int32)
    lens = np.ascontiguousarray(lens, np.uint32)
    k = lib.ref_windows_single(u32p(np.ascontiguousarray(rid)), u32p(np.ascontiguousarray(pos)), st.ctypes.data_as(I32), n, u32p(lens), text, clip_l, clip_r, cutoff,
                               u32p(o[0]), u32p(o[1]), u32p(o[2]), u32p(o[3]), oc.ctypes.data_as(I32))
    assert k == n and (oc == cutoff).all()
    return np.stack(o, axis=1)


def ref_windows_half(lib, rid, pos, strand, lens, text, P, cut0, cut1):
    n = len(rid)
    I32 = C.POINTER(C.c_int)
    U8 = C.POINTER(C.c_uint8)
    u = [np.zeros(2 * n + 2, np.uint32) for _ in range(8)]
    lor, cut = np.zeros(2 * n + 2, np.int32), np.zeros(2 * n + 2, np.int32)
    st = np.ascontiguousarray(strand, np.uint8)
    lens = np.ascontiguousarray(lens, np.uint32)
    k = lib.ref_windows_half(u32p(np.ascontiguousarray(rid)), u32p(np.ascontiguousarray(pos)), st.ctypes.data_as(U8), n, u32p(lens), text, P["left"], P["right"],
                             P["ins_high"], P["ins_low"], P["max_dna"], P["clip_l"], P["clip_r"], cut0, cut1,
                             u32p(u[0]), lor.ctypes.data_as(I32), u32p(u[1]), u32p(u[2]), u32p(u[3]), u32p(u[4]), u32p(u[5]), u32p(u[6]), u32p(u[7]),
                             cut.ctypes.data_as(I32))
    # columns like oracle_windows_half: leftOrRight, start, length, readLen, (dpStrand: not recorded by the packer), clipLt, clipRt, ancL, ancR
    rows = np.stack([lor[:k].astype(np.uint32), u[1][:k], u[2][:k], u[3][:k], np.zeros(k, np.uint32), u[4][:k], u[5][:k], u[6][:k], u[7][:k]], axis=1)
    return u[0][:k].copy(), rows, cut[:k].copy()


def ref_windows_pair(lib, rid, pos, pos2, lens, text, P, lsc, lhit):
    """-> left [n, 6] start, length, clipLt, clipRt, ancL, ancR; right [n, 7] start, length, readLen, clipLt, clipRt, ancL, ancR"""
    n = len(rid)
    I32 = C.POINTER(C.c_int)
    L = [np.zeros(n, np.uint32) for _ in range(6)]
    R = [np.zeros(n, np.uint32) for _ in range(7)]
    cl, cr = np.zeros(n, np.int32), np.zeros(n, np.int32)
    lens = np.ascontiguousarray(lens, np.uint32)
    lsc = np.ascontiguousarray(lsc, np.int32)
    k = lib.ref_windows_pair(u32p(np.ascontiguousarray(rid)), u32p(np.ascontiguousarray(pos)), u32p(np.ascontiguousarray(pos2)), n, u32p(lens), text,
                             P["left"], P["right"], P["ins_high"], P["ins_low"], P["max_dna"], P["clip_l"], P["clip_r"], P["cut"][0], P["cut"][1],
                             u32p(L[0]), u32p(L[1]), u32p(L[2]), u32p(L[3]), u32p(L[4]), u32p(L[5]), cl.ctypes.data_as(I32),
                             lsc.ctypes.data_as(I32), u32p(np.ascontiguousarray(lhit, np.uint32)),
                             u32p(R[0]), u32p(R[1]), u32p(R[2]), u32p(R[3]), u32p(R[4]), u32p(R[5]), u32p(R[6]), cr.ctypes.data_as(I32))
    assert k == n
    return np.stack(L, axis=1), np.stack(R, axis=1)


# ---------------------------------------------------------------------------
# the reference's own CPU search (ProcessReadDoubleStrand2 on SRAModelConstruct's models), oracle/_ref/libref_cpu_search.so
# ---------------------------------------------------------------------------
def load_ref_cpu_search():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_cpu_search.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_cpu_create.restype = C.c_void_p
    lib.ref_cpu_create.argtypes = [U32P, U32P, C.c_ulonglong, C.c_uint32, C.c_uint32, C.c_uint32, U32P, U32P, C.c_int]
    lib.ref_cpu_free.argtypes = [C.c_void_p]
    lib.ref_cpu_search.restype = C.c_ulonglong
    lib.ref_cpu_search.argtypes = [C.c_void_p, U8P, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_int, U32P,
                                   C.c_uint32, U32P, U8P, U8P, U32P]
    lib.ref_cpu_describe.restype = C.c_int
    lib.ref_cpu_describe.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_char_p, C.c_int]
    return lib


class RefCpuSearch:
    """The reference's CPU index structs built from the arrays of an index (BWT code words of the text and of the reversed text,
    packed text, suffix array), and its CPU search over a batch of reads."""

    def __init__(self, lib, bwt, rbwt, isa0, risa0, n, pac, sa, threads=0):
        self.lib = lib
        self.keep = (bwt, rbwt, pac, sa)                   # the suffix array is used in place
        self.h = lib.ref_cpu_create(u32p(bwt), u32p(rbwt), len(bwt), isa0, risa0, n, u32p(pac), u32p(sa), threads)
        if not self.h:
            raise RuntimeError("ref_cpu_create failed")

    def describe(self, read_length, k, num_cases):
        buf = C.create_string_buffer(16384)
        self.lib.ref_cpu_describe(self.h, read_length, k, num_cases, buf, len(buf))
        return buf.value.decode()

    def search(self, reads, k, num_cases, max_output_per_read=0xFFFFFFFF, threads=0, out_cap=0):
        """reads: [n, L] uint8 base codes -> dict(total, counts [n, 4] = ranges, occurrences in ranges, check-and-extend occurrences,
        total; and with out_cap: hits = per read a list of (position, strand, mismatches))"""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n, L = reads.shape
        counts = np.zeros(4 * n, np.uint32)
        if out_cap:
            pos, st, mm, on = np.zeros(n * out_cap, np.uint32), np.zeros(n * out_cap, np.uint8), np.zeros(n * out_cap, np.uint8), np.zeros(n, np.uint32)
            args = (out_cap, u32p(pos), st.ctypes.data_as(U8P), mm.ctypes.data_as(U8P), u32p(on))
        else:
            args = (0, None, None, None, None)
        total = self.lib.ref_cpu_search(self.h, reads.ctypes.data_as(U8P), n, L, k, num_cases, max_output_per_read, threads, u32p(counts), *args)
        out = dict(total=int(total), counts=counts.reshape(n, 4))
        if out_cap:
            out["hits"] = [[(int(pos[r * out_cap + i]), int(st[r * out_cap + i]), int(mm[r * out_cap + i])) for i in range(int(on[r]))] for r in range(n)]
        return out

    def free(self):
        if self.h:
            self.lib.ref_cpu_free(self.h)
            self.h = None


# ---------------------------------------------------------------------------
# long reads: validation of seed alignments (validateAlignments)
# ---------------------------------------------------------------------------
_VAL_ONE = [U32P, C.c_uint32, U8P, C.c_uint32, C.c_uint32, C.c_uint32, U32P, U8P, U8P, C.c_int, C.c_int, C.c_int, C.c_int]


def load_ref_validate():
    path = os.path.join(ROOT, "oracle", "_ref", "libref_validate.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_validate.restype = C.c_uint32
    lib.ref_validate.argtypes = _VAL_ONE
    return lib


def validate_one(fn, pac, n_text, read, seed_len, pos, strand, mism, keep_best, min_seed, max_mism, max_hit):
    """fn = liboracle's s3o_validate_one or libref_validate's ref_validate -> kept (pos, strand, mism) lists"""
    fn.restype = C.c_uint32
    fn.argtypes = _VAL_ONE
    p, s, m = np.array(pos, np.uint32), np.array(strand, np.uint8), np.array(mism, np.uint8)
    if len(p) == 0:
        p, s, m = np.zeros(1, np.uint32), np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    read = np.ascontiguousarray(read, np.uint8)
    k = fn(u32p(pac), n_text, read.ctypes.data_as(U8P), seed_len, len(read), len(pos), u32p(p), s.ctypes.data_as(U8P), m.ctypes.data_as(U8P),
           int(keep_best), min_seed, max_mism, max_hit)
    return p[:k].tolist(), s[:k].tolist(), m[:k].tolist()


def oracle_validate_batch(lib, pac, n_text, reads, read_lengths, off, pos, flags, keep_best, min_seed, double_allowance, max_hit):
    """reads: [n, maxLen] uint8; CSR lists; flags [T, 2] = strand, mismatches -> (counts, pos, flags) with the kept entries in front"""
    lib.s3o_validate_batch.restype = None
    lib.s3o_validate_batch.argtypes = [U32P, C.c_uint32, U8P, C.c_uint32, U32P, C.c_uint64, U32P, U32P, U8P, U8P, C.c_int, C.c_int, C.c_int, C.c_int, U32P]
    reads = np.ascontiguousarray(reads, np.uint8)
    n = reads.shape[0]
    p = np.ascontiguousarray(pos, np.uint32).copy()
    fl = np.ascontiguousarray(flags, np.uint8).reshape(-1, 2)
    s, m = np.ascontiguousarray(fl[:, 0]).copy(), np.ascontiguousarray(fl[:, 1]).copy()
    if len(p) == 0:
        p, s, m = np.zeros(1, np.uint32), np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    cnt = np.zeros(max(n, 1), np.uint32)
    lens = np.ascontiguousarray(read_lengths, np.uint32)
    off = np.ascontiguousarray(off, np.uint32)
    lib.s3o_validate_batch(u32p(pac), n_text, reads.ctypes.data_as(U8P), reads.shape[1], u32p(lens), n, u32p(off), u32p(p), s.ctypes.data_as(U8P),
                           m.ctypes.data_as(U8P), int(keep_best), min_seed, int(double_allowance), max_hit, u32p(cnt))
    return cnt[:n], p, np.stack([s, m], axis=1)


def pack_text(codes):
    """base codes -> hsp->packedDNA words (16 bases per word, first base in the top bits), four words of padding"""
    codes = np.asarray(codes, np.uint8)
    n = len(codes)
    pac = np.zeros((n + 15) // 16 + 4, np.uint32)
    for k in range(16):
        v = codes[k::16].astype(np.uint32)
        pac[:len(v)] |= v << np.uint32(30 - 2 * k)
    return pac


def validation_cases(rng, genome, trials):
    """random reads with occurrence lists that hold true hits on both strands, near-misses, text edges, random places"""
    n = len(genome)
    for _ in range(trials):
        L = int(rng.integers(101, 260))
        seed = int(rng.choice([100, 100, 100, L, 37]))
        ext = L - seed if L > seed else 0
        p0 = int(rng.integers(0, n - L))
        read = genome[p0:p0 + L].copy()
        for _ in range(int(rng.integers(0, 6))):
            read[int(rng.integers(0, L))] = rng.integers(0, 4)
        rev = rng.random() < 0.4
        if rev:
            read = (3 - read[::-1]).astype(np.uint8)
        m = int(rng.integers(0, 12))
        pos, st, mm = [], [], []
        for _ in range(m):
            s = int(rng.integers(1, 3))
            c = rng.random()
            if s == 1:
                p = p0 if (c < 0.6 and not rev) else (n - int(rng.integers(0, L)) if c < 0.7 else int(rng.integers(0, n)))
            else:
                p = p0 + ext if (c < 0.6 and rev) else (int(rng.integers(0, L)) if c < 0.7 else int(rng.integers(0, n)))
            pos.append(p); st.append(s); mm.append(int(rng.integers(0, 3)))
        yield read, seed, pos, st, mm, int(rng.integers(0, 2)), int(rng.integers(0, 3)), int(rng.integers(0, 8)), int(rng.integers(1, 6))


# ---- the reference's index files (s3_index_load) --------------------------------------------------------------------------
def write_reference_files(prefix, idx):
    """the on-disk formats (2bwt-lib/BWT.c:170-285, BGS-Build.cpp:139-160, TextConverter.c:666-720) from in-memory arrays"""
    n = idx.text_length
    for half, tag in ((idx.fwd, ""), (idx.rev, ".rev")):
        hdr = np.array([half.inverse_sa0] + list(half.cum_freq[1:]), np.uint32)
        np.concatenate([hdr, half.bwt_words.cpu().numpy().view(np.uint32)[:(n + 15) // 16]]).tofile(prefix + tag + ".bwt")
        np.concatenate([hdr, half.occ.cpu().numpy().view(np.uint32)]).tofile(prefix + tag + ".fmv.gpu")
    hdr = np.array([idx.fwd.inverse_sa0] + list(idx.fwd.cum_freq[1:]), np.uint32)
    np.concatenate([hdr, np.array([1], np.uint32), idx.fwd.sa.cpu().numpy().astype(np.uint32)]).tofile(prefix + ".sa")
    words = idx.packed_text.cpu().numpy().view(np.uint32)[:(n + 15) // 16]
    # n / 4 + 1 data bytes (the last one holds the n % 4 bases left over, none when n is a multiple of four), then n % 4
    data = (words.astype(">u4").tobytes() + b"\0")[:n // 4 + 1]
    last = n % 4
    open(prefix + ".pac", "wb").write(data + bytes([last]))


def run_reference_builders(G, tmp):
    """soap3-dp-builder + BGS-Build (oracle/_ref, the reference's own, built by oracle/build_ref.sh) on genome G; returns the index prefix"""
    import subprocess
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    fa = os.path.join(tmp, "g.fa")
    seq = "".join(np.array(list("ACGT"))[G.numpy()])
    with open(fa, "w") as f:
        f.write(">chr1\n")
        for i in range(0, len(seq), 60):
            f.write(seq[i:i + 60] + "\n")
    subprocess.check_call([os.path.join(ref_dir, "soap3-dp-builder"), fa], stdout=subprocess.DEVNULL, cwd=ref_dir)
    subprocess.check_call([os.path.join(ref_dir, "BGS-Build"), fa + ".index"], stdout=subprocess.DEVNULL, cwd=ref_dir)
    return fa + ".index"
