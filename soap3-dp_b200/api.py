"""Host-side mirror of the reference's GPU boundary over the C ABI.

Same names, argument meaning and error behaviour as the reference functions the
C ABI replaces (include/soap3dp_b200.h cites file:line):

    GPUINDEXUpload / GPUINDEXFree                      alignment.cu:27-115
    perform_round1_alignment / perform_round2_alignment alignment.cu:118-326
    SemiGlobalAligner                                   DV-DPfunctions.h:120-164

Everything here is ctypes over ``libsoap3dp_b200.so``.  There is NO CPU
fallback: if the library is missing or no CUDA device is present the calls
raise (the reference prints "CUDA ... FAILED" and exit(1)s, alignment.cu:38-42).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import formats

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsoap3dp_b200.so")

U32P = C.POINTER(C.c_uint32)
I32P = C.POINTER(C.c_int32)
U8P = C.POINTER(C.c_uint8)
U64P = C.POINTER(C.c_uint64)

EXPORTS = [
    "s3_last_error", "s3_device_count", "s3_launch_count", "s3_dp_set_stream", "s3_index_upload", "s3_index_free", "s3_index_device_bytes",
    "s3_index_set_locate_device", "s3_search_set_split_budget",
    "s3_index_set_timing", "s3_index_read_timing", "s3_dp_set_timing", "s3_dp_read_timing",
    "s3_search", "s3_search_result_free", "s3_locate", "s3_free", "s3_dp_align_windows", "s3_dp_decode", "s3_dp_md", "s3_seed_layout", "s3_dp_stage_parameters", "s3_pair_occurrences", "s3_retain_best",
    "s3_mapq_unique", "s3_mapq_bwa_single", "s3_mapq_single", "s3_mapq_single_dp", "s3_mapq_bwa_pair", "s3_mapq_pair_end",
    "s3_mapq_unique_dp", "s3_mapq_pair_end_dp", "s3_mapq_of_pair", "s3_seed_candidates", "s3_seed_pair_candidates",
    "s3_index_stream", "s3_rank_probe", "s3_search_round1", "s3_search_round2", "s3_search_round1_device",
    "s3_dp_create", "s3_dp_free", "s3_dp_stream", "s3_dp_pattern_length", "s3_dp_align", "s3_dp_align_device",
    "s3_dp_align_windows_device", "s3_random_sector_probe", "s3_dp_make_windows", "s3_index_set_l2_persist", "s3_index_clone",
    "s3_pe_create", "s3_pe_free", "s3_pe_prefetch", "s3_pe_align", "s3_pe_align_device", "s3_pe_set_timing", "s3_pe_read_timing", "s3_pe_dp",
    "s3_se_create", "s3_se_free", "s3_se_align", "s3_se_align_device", "s3_validate_alignments", "s3_sam_pair_records", "s3_sam_single_record", "s3_sam_single_dp_record", "s3_sam_deep_dp_records", "s3_sam_pair_dp_records", "s3_sam_unpaired_records", "s3_sam_unpaired_dp_records", "s3_sam_single_answer_record", "s3_sam_format_line", "s3_sam_pick_deep_dp", "s3_sam_pick_pair_dp", "s3_sam_single_batch_text", "s3_sam_single_dp_batch_text", "s3_sam_deep_dp_batch_text", "s3_sam_pair_dp_batch_text", "s3_sam_paired_batch_text", "s3_sam_unpaired_batch_text", "s3_sam_unpaired_dp_batch_text", "s3_runs_decode", "s3_sam_record_free", "s3_index_load", "s3_pe_deep_dp", "s3_seed_search", "s3_seed_search_result_free",
    "s3_single_dp_align", "s3_single_dp_result_free", "s3_deep_dp_align", "s3_deep_dp_result_free",
]


class S3Error(RuntimeError):
    pass


class DPScores(C.Structure):
    _fields_ = [("matchScore", C.c_int32), ("mismatchScore", C.c_int32),
                ("gapOpenScore", C.c_int32), ("gapExtendScore", C.c_int32)]


_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA library; fails loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("S3_LIB_PATH", LIB_PATH)      # tuning experiments load a variant build of the same library
    if not os.path.exists(path):
        raise S3Error(f"{path} is not built (run `python __graft_entry__.py build`); "
                      "soap3dp_b200 has no CPU fallback")
    lib = C.CDLL(path)
    lib.s3_last_error.restype = C.c_char_p
    lib.s3_device_count.restype = C.c_int
    lib.s3_launch_count.restype = C.c_ulonglong
    lib.s3_dp_set_stream.restype = None
    lib.s3_dp_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.s3_index_upload.restype = C.c_int
    lib.s3_index_upload.argtypes = [U32P, U32P, U32P, U32P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    U32P, U32P, C.c_int, C.POINTER(C.c_void_p)]
    lib.s3_index_set_locate_device.restype = C.c_int
    lib.s3_index_set_locate_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.s3_search_set_split_budget.restype = C.c_int
    lib.s3_search_set_split_budget.argtypes = [C.c_void_p, C.c_int32]
    lib.s3_index_free.restype = None
    lib.s3_index_free.argtypes = [C.c_void_p]
    lib.s3_index_device_bytes.restype = C.c_size_t
    lib.s3_index_device_bytes.argtypes = [C.c_void_p]
    lib.s3_index_stream.restype = C.c_void_p
    lib.s3_index_stream.argtypes = [C.c_void_p]
    lib.s3_rank_probe.restype = C.c_int
    lib.s3_rank_probe.argtypes = [C.c_void_p, C.c_int, U32P, C.c_size_t, U32P]
    PP = C.POINTER(C.c_void_p)
    lib.s3_search_round1.restype = C.c_int
    lib.s3_search_round1.argtypes = [C.c_void_p, U32P, U32P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_uint32, C.c_int, PP]
    lib.s3_search_round2.restype = C.c_int
    lib.s3_search_round2.argtypes = [C.c_void_p, U32P, U32P, PP, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, PP, PP, U64P]
    lib.s3_search_round1_device.restype = C.c_int
    lib.s3_search_round1_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                            C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, PP, C.c_void_p]
    lib.s3_dp_create.restype = C.c_int
    lib.s3_dp_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, DPScores, C.c_int, C.POINTER(C.c_void_p)]
    lib.s3_dp_free.restype = None
    lib.s3_dp_free.argtypes = [C.c_void_p]
    lib.s3_dp_stream.restype = C.c_void_p
    lib.s3_dp_stream.argtypes = [C.c_void_p]
    lib.s3_dp_pattern_length.restype = C.c_uint32
    lib.s3_dp_pattern_length.argtypes = [C.c_void_p]
    lib.s3_dp_align.restype = C.c_int
    lib.s3_dp_align.argtypes = [C.c_void_p, U32P, U32P, U32P, U32P, I32P, I32P, U32P, U32P, U8P, C.c_uint32,
                                U32P, U32P, U32P, U32P]
    lib.s3_dp_align_device.restype = C.c_int
    lib.s3_dp_align_device.argtypes = [C.c_void_p] + [C.c_void_p] * 9 + [C.c_uint32] + [C.c_void_p] * 4
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        raise S3Error(f"{what} failed ({rc}): {load_library().s3_last_error().decode()}")


def _u32(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags.c_contiguous, "need contiguous uint32"
    return a.ctypes.data_as(U32P)


def _ptr_array(arrs: Sequence[np.ndarray]):
    arr = (C.c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        arr[i] = a.ctypes.data
    return arr


class GpuIndex:
    """Opaque device index handle (what the reference keeps as _bwt/_occ/_revBwt/_revOcc)."""

    def __init__(self, handle: int, text_length: int):
        self.handle = C.c_void_p(handle)
        self.text_length = text_length

    @property
    def stream(self) -> int:
        return load_library().s3_index_stream(self.handle)

    @property
    def device_bytes(self) -> int:
        return load_library().s3_index_device_bytes(self.handle)


def _np_u32(t) -> np.ndarray:
    """torch int32 / numpy array -> contiguous numpy uint32 view"""
    if hasattr(t, "detach"):
        t = t.detach().cpu().contiguous().numpy()
    return np.ascontiguousarray(t).view(np.uint32)


def GPUINDEXUpload(index, device: int = 0, with_text: bool = False, with_sa: bool = False) -> GpuIndex:
    """alignment.cu:27: copy the index to device memory (and re-lay it out).
    ``index`` is a fmindex.Soap3IndexArrays (the Soap3Index stand-in)."""
    lib = load_library()
    bwt, occ = _np_u32(index.fwd.bwt_words), _np_u32(index.fwd.occ)
    rbwt, rocc = _np_u32(index.rev.bwt_words), _np_u32(index.rev.occ)
    pac = sa = None
    if with_text:
        if index.packed_text is None:
            raise S3Error("GPUINDEXUpload: index has no packed text")
        pac = _np_u32(index.packed_text)
    if with_sa:
        if index.fwd.sa is None:
            raise S3Error("GPUINDEXUpload: index has no suffix array")
        s = index.fwd.sa
        s = s.detach().cpu().numpy() if hasattr(s, "detach") else np.asarray(s)
        sa = np.ascontiguousarray(s.astype(np.uint32))
    out = C.c_void_p()
    rc = lib.s3_index_upload(_u32(bwt), _u32(occ), _u32(rbwt), _u32(rocc), index.fwd.num_occ,
                             index.fwd.inverse_sa0, index.rev.inverse_sa0, index.text_length,
                             _u32(pac) if pac is not None else None, _u32(sa) if sa is not None else None,
                             device, C.byref(out))
    _check(rc, "GPUINDEXUpload")
    return GpuIndex(out.value, index.text_length)


def set_locate_device(gpu_index: GpuIndex, d_sa: int, d_packed_text: int):
    """s3_index_set_locate_device: suffix array (uint32[n+1]) and packed text already on the device
    (raw device pointers); enables check-and-extend."""
    _check(load_library().s3_index_set_locate_device(gpu_index.handle, C.c_void_p(d_sa), C.c_void_p(d_packed_text)),
           "s3_index_set_locate_device")


class SearchResult(C.Structure):
    _fields_ = [("numReads", C.c_uint64), ("total", C.c_uint64), ("offsets", U64P), ("saL", U32P), ("saR", U32P), ("info", U32P)]


def search(gpu_index: GpuIndex, queries: np.ndarray, read_lengths: np.ndarray, batch_size: int, word_per_query: int,
           num_mismatch: int, is_exact_num_mismatch: bool = False):
    """Capless search (include/soap3dp_b200.h s3_search): -> (offsets[N+1] uint64, saL, saR, info) numpy copies."""
    lib = load_library()
    lib.s3_search.restype = C.c_int
    lib.s3_search.argtypes = [C.c_void_p, U32P, U32P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(SearchResult)]
    lib.s3_search_result_free.restype = None
    lib.s3_search_result_free.argtypes = [C.POINTER(SearchResult)]
    res = SearchResult()
    _check(lib.s3_search(gpu_index.handle, _u32(queries), _u32(read_lengths), batch_size, word_per_query, num_mismatch,
                         1 if is_exact_num_mismatch else 0, C.byref(res)), "s3_search")
    try:
        n, tot = int(res.numReads), int(res.total)
        offsets = np.ctypeslib.as_array(res.offsets, shape=(n + 1,)).copy()
        if tot:
            sa_l = np.ctypeslib.as_array(res.saL, shape=(tot,)).copy()
            sa_r = np.ctypeslib.as_array(res.saR, shape=(tot,)).copy()
            info = np.ctypeslib.as_array(res.info, shape=(tot,)).copy()
        else:
            sa_l = sa_r = info = np.zeros(0, np.uint32)
    finally:
        lib.s3_search_result_free(C.byref(res))
    return offsets, sa_l, sa_r, info


def locate(gpu_index: GpuIndex, sa_l: np.ndarray, sa_r: np.ndarray, max_per_range: int = 0xFFFFFFFF):
    """SA ranges -> text positions (include/soap3dp_b200.h s3_locate): -> (offsets[n+1] uint64, positions uint32)."""
    lib = load_library()
    lib.s3_locate.restype = C.c_int
    lib.s3_locate.argtypes = [C.c_void_p, U32P, U32P, C.c_uint64, C.c_uint32, U64P, C.POINTER(U32P), U64P]
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    n = len(sa_l)
    sa_l = np.ascontiguousarray(sa_l, np.uint32)
    sa_r = np.ascontiguousarray(sa_r, np.uint32)
    offsets = np.zeros(n + 1, np.uint64)
    pos, total = U32P(), C.c_uint64(0)
    _check(lib.s3_locate(gpu_index.handle, _u32(sa_l), _u32(sa_r), n, max_per_range,
                         offsets.ctypes.data_as(U64P), C.byref(pos), C.byref(total)), "s3_locate")
    try:
        out = np.ctypeslib.as_array(pos, shape=(int(total.value),)).copy() if total.value else np.zeros(0, np.uint32)
    finally:
        if total.value:
            lib.s3_free(pos)
    return offsets, out


def seed_candidates(gpu_index: GpuIndex, sa_l, sa_r, strands, read_ids, offsets, seed_lengths, read_lengths,
                    max_per_range: int = 0xFFFFFFFF):
    """s3_seed_candidates (SingleEndSeedingBatch::decodePositions + singleMerge): -> (readIDs, positions, strands)."""
    lib = load_library()
    lib.s3_seed_candidates.restype = C.c_int
    lib.s3_seed_candidates.argtypes = [C.c_void_p, U32P, U32P, I32P, U32P, U32P, U32P, U32P, C.c_uint64, C.c_uint32,
                                       C.POINTER(U32P), C.POINTER(U32P), C.POINTER(I32P), U64P]
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    u = [np.ascontiguousarray(a, np.uint32) for a in (sa_l, sa_r, read_ids, offsets, seed_lengths, read_lengths)]
    st = np.ascontiguousarray(strands, np.int32)
    r, p, s, m = U32P(), U32P(), I32P(), C.c_uint64(0)
    _check(lib.s3_seed_candidates(gpu_index.handle, _u32(u[0]), _u32(u[1]), st.ctypes.data_as(I32P), _u32(u[2]), _u32(u[3]),
                                  _u32(u[4]), _u32(u[5]), len(st), max_per_range, C.byref(r), C.byref(p), C.byref(s), C.byref(m)),
           "s3_seed_candidates")
    n = int(m.value)
    if n == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.int32)
    try:
        return (np.ctypeslib.as_array(r, shape=(n,)).copy(), np.ctypeslib.as_array(p, shape=(n,)).copy(),
                np.ctypeslib.as_array(s, shape=(n,)).copy())
    finally:
        lib.s3_free(r); lib.s3_free(p); lib.s3_free(s)


def seed_pair_candidates(gpu_index: GpuIndex, side0, side1, lengths_by_read_id, insert_low: int, insert_high: int,
                         left_leg: int = 1, right_leg: int = 2, max_per_range: int = 0xFFFFFFFF):
    """s3_seed_pair_candidates (PairEndSeedingBatch::decodeMergePositions).  side0 / side1 = (saL, saR, strands, readIDs,
    offsets, seedLengths, readLengths) of the reads' and the mates' seed ranges -> (readIDLeft, posLeft, posRight)."""
    lib = load_library()
    side_t = [U32P, U32P, I32P, U32P, U32P, U32P, U32P, C.c_uint64]
    lib.s3_seed_pair_candidates.restype = C.c_int
    lib.s3_seed_pair_candidates.argtypes = [C.c_void_p] + side_t + side_t + [C.c_uint32, U32P, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                                                            C.c_int, C.POINTER(U32P), C.POINTER(U32P), C.POINTER(U32P), U64P]
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    keep, args = [], []
    for sd in (side0, side1):
        a = [np.ascontiguousarray(sd[0], np.uint32), np.ascontiguousarray(sd[1], np.uint32), np.ascontiguousarray(sd[2], np.int32)] + \
            [np.ascontiguousarray(x, np.uint32) for x in sd[3:7]]
        keep.append(a)
        args += [_u32(a[0]), _u32(a[1]), a[2].ctypes.data_as(I32P), _u32(a[3]), _u32(a[4]), _u32(a[5]), _u32(a[6]), len(a[0])]
    lens = np.ascontiguousarray(lengths_by_read_id, np.uint32)
    r, pl, pr, m = U32P(), U32P(), U32P(), C.c_uint64(0)
    _check(lib.s3_seed_pair_candidates(gpu_index.handle, *args, max_per_range, _u32(lens), len(lens), insert_low, insert_high,
                                       left_leg, right_leg, C.byref(r), C.byref(pl), C.byref(pr), C.byref(m)), "s3_seed_pair_candidates")
    n = int(m.value)
    if n == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint32)
    try:
        return tuple(np.ctypeslib.as_array(x, shape=(n,)).copy() for x in (r, pl, pr))
    finally:
        lib.s3_free(r); lib.s3_free(pl); lib.s3_free(pr)


def decode_alignments(pattern: np.ndarray, pattern_length: int, scores, read_lengths, cutoffs, dp_scores: "DPScores",
                      sam: bool = True, split: bool = True):
    """s3_dp_decode (the result loops of the DP engines' CPU threads + CigarStringEncoder, DV-DPfunctions.cu:1699-1733,
    DV-DPfunctions.h:514-597; convertToCigarStr PE.cpp:420-485): -> dict(cigar=[str], sam=[str] | None, editdist,
    ref_span_delta, op_counts[n,5] as M m I D S).  Alignments under their cutoff give '' and editdist -1.
    split=False leaves the strings as the library returns them: cigar / sam = (offsets[n+1], bytes)."""
    lib = load_library()
    lib.s3_dp_decode.restype = C.c_int
    lib.s3_dp_decode.argtypes = [U8P, C.c_uint32, I32P, U32P, I32P, C.c_uint32, DPScores, U64P, C.POINTER(C.c_char_p),
                                 U64P, C.POINTER(C.c_char_p), I32P, I32P, U32P]
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    n = len(scores)
    pattern = np.ascontiguousarray(pattern, np.uint8)
    assert pattern.size >= n * pattern_length
    scores = np.ascontiguousarray(scores, np.int32)
    read_lengths = np.ascontiguousarray(read_lengths, np.uint32)
    cutoffs = np.ascontiguousarray(cutoffs, np.int32)
    off, soff = np.zeros(n + 1, np.uint64), np.zeros(n + 1, np.uint64)
    ed, span, ops = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros((n, 5), np.uint32)
    cig, scig = C.c_char_p(), C.c_char_p()
    _check(lib.s3_dp_decode(pattern.ctypes.data_as(U8P), pattern_length, scores.ctypes.data_as(I32P), _u32(read_lengths),
                            cutoffs.ctypes.data_as(I32P), n, dp_scores, off.ctypes.data_as(U64P), C.byref(cig),
                            soff.ctypes.data_as(U64P) if sam else None, C.byref(scig) if sam else None,
                            ed.ctypes.data_as(I32P), span.ctypes.data_as(I32P), ops.ctypes.data_as(U32P)), "s3_dp_decode")
    try:
        text = C.string_at(cig, int(off[n]))
        cigars = (off, text)
        sams = (soff, C.string_at(scig, int(soff[n]))) if sam else None
        if split:
            text = text.decode("ascii")
            cigars = [text[int(off[t]):int(off[t + 1])] for t in range(n)]
            if sam:
                text = sams[1].decode("ascii")
                sams = [text[int(soff[t]):int(soff[t + 1])] for t in range(n)]
    finally:
        lib.s3_free(cig)
        if sam:
            lib.s3_free(scig)
    return dict(cigar=cigars, sam=sams, editdist=ed, ref_span_delta=span, op_counts=ops)


def pair_occurrences(gpu_index: GpuIndex, pos1, strand1, mism1, off1, pos2, strand2, mism2, off2, pattern_lengths,
                     insert_lbound: int, insert_ubound: int, strand_left_leg: int = 1, strand_right_leg: int = 2,
                     report_one: bool = False):
    """s3_pair_occurrences (PEMappingOccurrences + PEStatsPEOutput, PEAlgnmt.cpp:480-637,777-838, batched over read pairs):
    -> dict(offsets[numPairs+1], pos1, pos2, insertion, flags[n,4] = strand_1 mismatch_1 strand_2 mismatch_2, optimal,
    suboptimal, stats[numPairs,32])."""
    lib = load_library()
    U8PP = C.POINTER(U8P)
    lib.s3_pair_occurrences.restype = C.c_int
    lib.s3_pair_occurrences.argtypes = [C.c_void_p, U32P, U8P, U8P, U64P, U32P, U8P, U8P, U64P, U32P, C.c_uint64, C.c_int32, C.c_int32,
                                        C.c_int, C.c_int, C.c_int, U64P, C.POINTER(U32P), C.POINTER(U32P), C.POINTER(U32P), U8PP,
                                        U32P, U32P, U32P]
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    a32 = lambda x: np.ascontiguousarray(x, np.uint32)
    a8 = lambda x: np.ascontiguousarray(x, np.uint8)
    a64 = lambda x: np.ascontiguousarray(x, np.uint64)
    pos1, pos2, pl = a32(pos1), a32(pos2), a32(pattern_lengths)
    strand1, mism1, strand2, mism2 = a8(strand1), a8(mism1), a8(strand2), a8(mism2)
    off1, off2 = a64(off1), a64(off2)
    npairs = len(off1) - 1
    assert len(off2) == npairs + 1 and len(pl) >= npairs
    offs = np.zeros(npairs + 1, np.uint64)
    opt, sub, stats = np.zeros(npairs, np.uint32), np.zeros(npairs, np.uint32), np.zeros((npairs, 32), np.uint32)
    o1, o2, oi, of = U32P(), U32P(), U32P(), U8P()
    b8 = lambda x: x.ctypes.data_as(U8P)
    _check(lib.s3_pair_occurrences(gpu_index.handle, _u32(pos1), b8(strand1), b8(mism1), off1.ctypes.data_as(U64P),
                                   _u32(pos2), b8(strand2), b8(mism2), off2.ctypes.data_as(U64P), _u32(pl), npairs,
                                   insert_lbound, insert_ubound, strand_left_leg, strand_right_leg, 1 if report_one else 0,
                                   offs.ctypes.data_as(U64P), C.byref(o1), C.byref(o2), C.byref(oi), C.byref(of),
                                   _u32(opt), _u32(sub), _u32(stats)), "s3_pair_occurrences")
    n = int(offs[npairs])
    try:
        if n:
            a, b, ins = (np.ctypeslib.as_array(x, shape=(n,)).copy() for x in (o1, o2, oi))
            fl = np.ctypeslib.as_array(of, shape=(n, 4)).copy()
        else:
            a = b = ins = np.zeros(0, np.uint32)
            fl = np.zeros((0, 4), np.uint8)
    finally:
        if n:
            lib.s3_free(o1); lib.s3_free(o2); lib.s3_free(oi); lib.s3_free(of)
    return dict(offsets=offs, pos1=a, pos2=b, insertion=ins, flags=fl, optimal=opt, suboptimal=sub, stats=stats)


def md_strings(packed_text: np.ndarray, text_length: int, cigars, positions, qualities=None):
    """s3_dp_md (getMisInfoForDP, PE.cpp:499-666): special CIGARs (list of str) + text positions ->
    dict(md=[str], num_mismatch, gap_open, gap_ext, avg_mismatch_qual).  qualities: list of int8 arrays in read order."""
    lib = load_library()
    I8P = C.POINTER(C.c_int8)
    lib.s3_dp_md.restype = C.c_int
    lib.s3_dp_md.argtypes = [U32P, C.c_uint64, C.c_char_p, U64P, U32P, C.c_uint32, I8P, U64P, U64P, C.POINTER(C.c_char_p), I32P, I32P, I32P, I32P]
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    n = len(cigars)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c in cigars])
    text = "".join(cigars).encode("ascii")
    pos = np.ascontiguousarray(positions, np.uint32)
    packed = np.ascontiguousarray(packed_text, np.uint32)
    q = qoff = None
    if qualities is not None:
        qoff = np.zeros(n + 1, np.uint64)
        qoff[1:] = np.cumsum([len(x) for x in qualities])
        q = np.ascontiguousarray(np.concatenate([np.asarray(x, np.int8) for x in qualities]) if n else np.zeros(0, np.int8))
    moff = np.zeros(n + 1, np.uint64)
    nm, go, ge, aq = (np.zeros(n, np.int32) for _ in range(4))
    out = C.c_char_p()
    i32 = lambda a: a.ctypes.data_as(I32P)
    _check(lib.s3_dp_md(_u32(packed), text_length, text, off.ctypes.data_as(U64P), _u32(pos), n,
                        q.ctypes.data_as(I8P) if q is not None else None, qoff.ctypes.data_as(U64P) if qoff is not None else None,
                        moff.ctypes.data_as(U64P), C.byref(out), i32(nm), i32(go), i32(ge), i32(aq)), "s3_dp_md")
    try:
        md = C.string_at(out, int(moff[n])).decode("ascii")
    finally:
        lib.s3_free(out)
    return dict(md=[md[int(moff[t]):int(moff[t + 1])] for t in range(n)], num_mismatch=nm, gap_open=go, gap_ext=ge, avg_mismatch_qual=aq)


RETAIN_ALL_BEST, RETAIN_BEST_WITH_CAP, RETAIN_BEST_AND_SECOND = 0, 1, 2


def retain_best(gpu_index: GpuIndex, mode: int, sa_l, sa_r, sa_strand, sa_mism, sa_off, occ_pos, occ_strand, occ_mism, occ_off,
                max_num: int = 0):
    """s3_retain_best (retainAllBest / retainAllBestWithCap / retainAllBestAndSecBest, SAList.cpp:140-348, a batch of reads):
    -> dict(sa_off, sa_l, sa_r, sa_flags[n,2] = strand mismatchCount, occ_off, occ_pos, occ_flags[n,2], num)."""
    lib = load_library()
    lib.s3_retain_best.restype = C.c_int
    lib.s3_retain_best.argtypes = [C.c_void_p, C.c_int, C.c_int32, U32P, U32P, U8P, U8P, U64P, U32P, U8P, U8P, U64P, C.c_uint64,
                                   U64P, U32P, U32P, U8P, U64P, U32P, U8P, U32P]
    a32 = lambda x: np.ascontiguousarray(x, np.uint32)
    a8 = lambda x: np.ascontiguousarray(x, np.uint8)
    sa_l, sa_r, occ_pos = a32(sa_l), a32(sa_r), a32(occ_pos)
    sa_strand, sa_mism, occ_strand, occ_mism = a8(sa_strand), a8(sa_mism), a8(occ_strand), a8(occ_mism)
    sa_off, occ_off = np.ascontiguousarray(sa_off, np.uint64), np.ascontiguousarray(occ_off, np.uint64)
    n = len(sa_off) - 1
    assert len(occ_off) == n + 1
    o_sa_off, o_occ_off = np.zeros(n + 1, np.uint64), np.zeros(n + 1, np.uint64)
    o_l, o_r, o_sf = np.zeros(len(sa_l), np.uint32), np.zeros(len(sa_l), np.uint32), np.zeros((len(sa_l), 2), np.uint8)
    o_p, o_of = np.zeros(len(occ_pos), np.uint32), np.zeros((len(occ_pos), 2), np.uint8)
    num = np.zeros(n, np.uint32)
    b8 = lambda x: x.ctypes.data_as(U8P)
    _check(lib.s3_retain_best(gpu_index.handle, mode, max_num, _u32(sa_l), _u32(sa_r), b8(sa_strand), b8(sa_mism), sa_off.ctypes.data_as(U64P),
                              _u32(occ_pos), b8(occ_strand), b8(occ_mism), occ_off.ctypes.data_as(U64P), n,
                              o_sa_off.ctypes.data_as(U64P), _u32(o_l), _u32(o_r), b8(o_sf), o_occ_off.ctypes.data_as(U64P), _u32(o_p), b8(o_of),
                              _u32(num)), "s3_retain_best")
    ks, ko = int(o_sa_off[-1]), int(o_occ_off[-1])
    return dict(sa_off=o_sa_off, sa_l=o_l[:ks], sa_r=o_r[:ks], sa_flags=o_sf[:ks], occ_off=o_occ_off, occ_pos=o_p[:ko], occ_flags=o_of[:ko], num=num)


STAGE_SINGLE_DP, STAGE_DEFAULT_DP, STAGE_NEW_DEFAULT_DP, STAGE_DEEP_DP_ROUND1, STAGE_DEEP_DP_ROUND2 = 1, 2, 3, 4, 5   # definitions.h:317-321


class DPReadParams(C.Structure):
    _fields_ = [("cutoffThreshold", C.c_int32), ("maxHitNum", C.c_int32), ("sampleDist", C.c_int32), ("seedLength", C.c_int32)]


class DPStageParams(C.Structure):
    _fields_ = [("softClipLeft", C.c_int32), ("softClipRight", C.c_int32), ("tailTrimLen", C.c_int32),
                ("singleDPSeedNum", C.c_int32), ("singleDPSeedPos", C.c_int32 * 10), ("paramRead", DPReadParams * 2)]


def getSeedPositions(stage: int, read_length: int, capacity: int = 256):
    """s3_seed_layout (getSeedPositions, definitions.h:323-442): -> (seedLength, [seed offsets])."""
    lib = load_library()
    lib.s3_seed_layout.restype = C.c_int
    lib.s3_seed_layout.argtypes = [C.c_int, C.c_int32, I32P, I32P, C.c_int32, I32P]
    sl, num = C.c_int32(0), C.c_int32(0)
    pos = (C.c_int32 * capacity)()
    _check(lib.s3_seed_layout(stage, read_length, C.byref(sl), pos, capacity, C.byref(num)), "s3_seed_layout")
    return int(sl.value), [int(pos[i]) for i in range(num.value)]


def getParameterForDP(stage: int, read_length: int, read_length2: int = 0, is_default_threshold: bool = True,
                      dp_score_threshold: int = 0, max_front_clipped: int = 0, max_end_clipped: int = 0) -> DPStageParams:
    """s3_dp_stage_parameters (getParameterFor{SingleDP,DefaultDP,NewDefaultDP,DeepDP}, CPUfunctions.cpp:59-260)."""
    lib = load_library()
    lib.s3_dp_stage_parameters.restype = C.c_int
    lib.s3_dp_stage_parameters.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int32, C.c_int32, C.c_int32,
                                           C.POINTER(DPStageParams)]
    out = DPStageParams()
    _check(lib.s3_dp_stage_parameters(stage, read_length, read_length2, 1 if is_default_threshold else 0, dp_score_threshold,
                                      max_front_clipped, max_end_clipped, C.byref(out)), "s3_dp_stage_parameters")
    return out


def set_timing(handle: int, on: bool, dp: bool = False):
    """Per-kernel timing hooks of include/soap3dp_b200.h (handle: GpuIndex.handle or SemiGlobalAligner.handle)."""
    lib = load_library()
    fn = lib.s3_dp_set_timing if dp else lib.s3_index_set_timing
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int]
    _check(fn(handle, 1 if on else 0), "set_timing")


def read_timing(handle: int, dp: bool = False):
    """-> (ms per kernel slot, launches per kernel slot) since the last read."""
    lib = load_library()
    fn = lib.s3_dp_read_timing if dp else lib.s3_index_read_timing
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    ms, cnt = (C.c_float * 8)(), (C.c_int * 8)()
    _check(fn(handle, ms, cnt), "read_timing")
    return list(ms), list(cnt)


def set_split_budget(gpu_index: GpuIndex, steps: int):
    """Tuning knob of the search (include/soap3dp_b200.h): steps before a long enumeration is split."""
    _check(load_library().s3_search_set_split_budget(gpu_index.handle, steps), "s3_search_set_split_budget")


def GPUINDEXFree(gpu_index: GpuIndex):
    """alignment.cu:109"""
    if gpu_index.handle:
        load_library().s3_index_free(gpu_index.handle)
        gpu_index.handle = C.c_void_p(0)


def rank_probe(gpu_index: GpuIndex, which: int, indices: np.ndarray) -> np.ndarray:
    idx = np.ascontiguousarray(indices, dtype=np.uint32)
    out = np.empty((idx.size, 4), dtype=np.uint32)
    _check(load_library().s3_rank_probe(gpu_index.handle, which, _u32(idx), idx.size, _u32(out.reshape(-1))), "s3_rank_probe")
    return out


def perform_round1_alignment(gpu_index: GpuIndex, queries: np.ndarray, read_lengths: np.ndarray, batch_size: int,
                             word_per_query: int, num_mismatch: int, num_cases: Optional[int] = None,
                             sa_range_allowed: Optional[int] = None, word_per_ans: Optional[int] = None,
                             is_exact_num_mismatch: bool = False) -> List[np.ndarray]:
    """alignment.cu:118.  Returns answers[case] (uint32[ceil32(batch)*word_per_ans])."""
    if num_cases is None:
        num_cases = formats.NUM_CASES[num_mismatch]
    if sa_range_allowed is None:
        sa_range_allowed = formats.SA_RANGES_ROUND1[num_mismatch]
    if word_per_ans is None:
        word_per_ans = 2 * sa_range_allowed
    up = formats.ceil32(batch_size)
    assert queries.size >= up * word_per_query, "queries must cover ceil32(batchSize) reads (alignment.cu:157)"
    answers = [np.empty(up * word_per_ans, dtype=np.uint32) for _ in range(num_cases)]
    rc = load_library().s3_search_round1(gpu_index.handle, _u32(queries), _u32(read_lengths), batch_size,
                                         word_per_query, num_mismatch, num_cases, sa_range_allowed, word_per_ans,
                                         int(is_exact_num_mismatch), _ptr_array(answers))
    _check(rc, "perform_round1_alignment")
    return answers


def perform_round2_alignment(gpu_index: GpuIndex, queries: np.ndarray, read_lengths: np.ndarray,
                             answers: Sequence[np.ndarray], batch_size: int, word_per_query: int, num_mismatch: int,
                             word_per_ans: int, sa_range_allowed_2: Optional[int] = None,
                             word_per_ans_2: Optional[int] = None, processed_query: int = 0,
                             is_exact_num_mismatch: bool = False):
    """alignment.cu:221.  Returns (badReadIndices[case], badAnswers[case])."""
    num_cases = len(answers)
    if sa_range_allowed_2 is None:
        sa_range_allowed_2 = formats.SA_RANGES_ROUND2[num_mismatch]
    if word_per_ans_2 is None:
        word_per_ans_2 = 2 * sa_range_allowed_2
    up = formats.ceil32(batch_size)
    bad_idx = [np.empty(max(batch_size, 1), dtype=np.uint32) for _ in range(num_cases)]
    bad_ans = [np.empty(max(up * word_per_ans_2, 1), dtype=np.uint32) for _ in range(num_cases)]
    num_bad = (C.c_uint64 * num_cases)()
    rc = load_library().s3_search_round2(gpu_index.handle, _u32(queries), _u32(read_lengths), _ptr_array(answers),
                                         batch_size, processed_query, word_per_query, num_mismatch, num_cases,
                                         sa_range_allowed_2, word_per_ans, word_per_ans_2,
                                         int(is_exact_num_mismatch), _ptr_array(bad_idx), _ptr_array(bad_ans), num_bad)
    _check(rc, "perform_round2_alignment")
    out_idx, out_ans = [], []
    for c in range(num_cases):
        nb = int(num_bad[c])
        out_idx.append(bad_idx[c][:nb].copy())
        out_ans.append(bad_ans[c][:formats.ceil32(nb) * word_per_ans_2].copy())
    return out_idx, out_ans


class SemiGlobalAligner:
    """DV-DPfunctions.h:120-164.  decideConfiguration/init collapse into the
    constructor (a B200 always has room for scheme 1, the full table);
    performAlignment keeps the reference's argument list."""

    def __init__(self, max_read_length: int, max_dna_length: int, batch_size: int,
                 match: int = 1, mismatch: int = -2, gap_open: int = -3, gap_extend: int = -1, device: int = 0):
        lib = load_library()
        self.max_read_length, self.max_dna_length, self.batch_size = max_read_length, max_dna_length, batch_size
        out = C.c_void_p()
        _check(lib.s3_dp_create(max_read_length, max_dna_length, batch_size,
                                DPScores(match, mismatch, gap_open, gap_extend), device, C.byref(out)),
               "SemiGlobalAligner.init")
        self.handle = out
        self.max_dp_table_length = max_dna_length               # scheme 1 (DV-DPfunctions.cu:592)
        self.pattern_length = lib.s3_dp_pattern_length(self.handle)

    @property
    def stream(self) -> int:
        return load_library().s3_dp_stream(self.handle)

    def performAlignment(self, packedDNASequence, DNALengths, packedReadSequence, readLengths, cutoffThresholds,
                         numOfThreads, clipLtSizes=None, clipRtSizes=None, anchorLeftLocs=None, anchorRightLocs=None):
        """DV-DPfunctions.cu:669.  Returns (scores, hitLocs, maxScoreCounts, pattern)."""
        up = formats.ceil32(max(numOfThreads, 1))
        scores = np.zeros(up, np.int32)
        hit = np.zeros(up, np.uint32)
        cnt = np.zeros(up, np.uint32)
        pat = np.zeros(up * self.pattern_length, np.uint8)

        def opt(a):
            return _u32(a) if a is not None else None
        rc = load_library().s3_dp_align(self.handle, _u32(packedDNASequence), _u32(DNALengths), _u32(packedReadSequence),
                                        _u32(readLengths), cutoffThresholds.ctypes.data_as(I32P),
                                        scores.ctypes.data_as(I32P), _u32(hit), _u32(cnt), pat.ctypes.data_as(U8P),
                                        numOfThreads, opt(clipLtSizes), opt(clipRtSizes), opt(anchorLeftLocs),
                                        opt(anchorRightLocs))
        _check(rc, "SemiGlobalAligner.performAlignment")
        return scores, hit, cnt, pat

    def performAlignmentOnWindows(self, gpu_index, queries, queryLengths, numQueries, wordPerOldQuery, readIDs, strands,
                                  DNAStarts, DNALengths, cutoffThresholds, numOfThreads, clipLtSizes=None, clipRtSizes=None,
                                  anchorLeftLocs=None, anchorRightLocs=None):
        """s3_dp_align_windows: the batch is packed on the device from the index's text and the query buffer
        (the engines' packRead / repackDNA, DV-DPfunctions.cu:1469-1524).  Returns (scores, hitLocs, maxScoreCounts, pattern)."""
        lib = load_library()
        lib.s3_dp_align_windows.restype = C.c_int
        lib.s3_dp_align_windows.argtypes = [C.c_void_p, C.c_void_p, U32P, U32P, C.c_uint64, C.c_uint32, U32P, U8P, U32P, U32P,
                                            I32P, I32P, U32P, U32P, U8P, C.c_uint32, U32P, U32P, U32P, U32P]
        up = formats.ceil32(max(numOfThreads, 1))
        scores = np.zeros(up, np.int32)
        hit = np.zeros(up, np.uint32)
        cnt = np.zeros(up, np.uint32)
        pat = np.zeros(up * self.pattern_length, np.uint8)

        def opt(a):
            return _u32(a) if a is not None else None
        strands = np.ascontiguousarray(strands, np.uint8)
        rc = lib.s3_dp_align_windows(self.handle, gpu_index.handle, _u32(queries), _u32(queryLengths), numQueries, wordPerOldQuery,
                                     _u32(readIDs), strands.ctypes.data_as(U8P), _u32(DNAStarts), _u32(DNALengths),
                                     cutoffThresholds.ctypes.data_as(I32P), scores.ctypes.data_as(I32P), _u32(hit), _u32(cnt),
                                     pat.ctypes.data_as(U8P), numOfThreads, opt(clipLtSizes), opt(clipRtSizes),
                                     opt(anchorLeftLocs), opt(anchorRightLocs))
        _check(rc, "SemiGlobalAligner.performAlignmentOnWindows")
        return scores, hit, cnt, pat

    def freeMemory(self):
        if self.handle:
            load_library().s3_dp_free(self.handle)
            self.handle = C.c_void_p(0)

    def set_stream(self, stream: int):
        load_library().s3_dp_set_stream(self.handle, C.c_void_p(stream))

    def align_device(self, d_dna: int, d_dna_len: int, d_read: int, d_read_len: int, d_cutoff: int, d_scores: int,
                     d_hit: int, d_cnt: int, d_pattern: int, n: int, d_clip_lt: int = 0, d_clip_rt: int = 0,
                     d_anchor_l: int = 0, d_anchor_r: int = 0):
        """s3_dp_align_device: raw device pointers (ints), stream-ordered, not synchronised."""
        p = [C.c_void_p(x or None) for x in (d_dna, d_dna_len, d_read, d_read_len, d_cutoff, d_scores, d_hit, d_cnt,
                                             d_pattern)]
        q = [C.c_void_p(x or None) for x in (d_clip_lt, d_clip_rt, d_anchor_l, d_anchor_r)]
        _check(load_library().s3_dp_align_device(self.handle, *p, n, *q), "s3_dp_align_device")


def search_round1_device(gpu_index: GpuIndex, d_queries: int, d_read_lengths: int, batch_size: int,
                         word_per_query: int, num_mismatch: int, num_cases: int, sa_range_allowed: int,
                         word_per_ans: int, d_answers: Sequence[int], d_rank_queries: int = 0,
                         is_exact_num_mismatch: bool = False):
    """s3_search_round1_device: raw device pointers (ints), enqueued on the index stream."""
    arr = (C.c_void_p * len(d_answers))(*[C.c_void_p(a) for a in d_answers])
    _check(load_library().s3_search_round1_device(gpu_index.handle, C.c_void_p(d_queries), C.c_void_p(d_read_lengths),
                                                  batch_size, word_per_query, num_mismatch, num_cases,
                                                  sa_range_allowed, word_per_ans, int(is_exact_num_mismatch), arr,
                                                  C.c_void_p(d_rank_queries or None)), "s3_search_round1_device")


def launch_count() -> int:
    return int(load_library().s3_launch_count())


# ---------------------------------------------------------------------------------------------------------------------
# paired-end batch on the device (s3_pe_*): the in-memory alignPairR the reference declares and leaves empty
# (soap3-dp-module.h:60, soap3-dp-module.cu:183-193)
# ---------------------------------------------------------------------------------------------------------------------
PE_NONE, PE_PAIRED, PE_FIRST_RESCUES, PE_SECOND_RESCUES, PE_BOTH_RESCUE = 0, 1, 2, 3, 4
PE_FIRST_TOO_MANY, PE_SECOND_TOO_MANY, PE_BOTH_NO_PAIR_MANY, PE_OVERFLOW = 5, 6, 7, 8


class PEParams(C.Structure):
    _fields_ = [("numMismatch", C.c_uint32), ("insertLow", C.c_int32), ("insertHigh", C.c_int32),
                ("strandLeftLeg", C.c_int32), ("strandRightLeg", C.c_int32), ("maxOutputPerRead", C.c_uint32),
                ("maxHitNumForDP", C.c_uint32), ("keepSecondBest", C.c_int32), ("scores", DPScores),
                ("cutoffThreshold", C.c_int32), ("softClipLeft", C.c_int32), ("softClipRight", C.c_int32),
                ("maxWindows", C.c_uint32), ("readStats", C.c_int32)]


class PEResult(C.Structure):
    _fields_ = [("numPairs", C.c_uint64), ("numRanges", C.c_uint64), ("numOccurrences", C.c_uint64), ("numWindows", C.c_uint64),
                ("numRuns", C.c_uint64), ("routeCounts", C.c_uint32 * 16), ("h2dBytes", C.c_uint64), ("d2hBytes", C.c_uint64),
                ("route", C.c_void_p), ("pairs", C.c_void_p), ("dp", C.c_void_p), ("runs", C.c_void_p),
                ("d_route", C.c_void_p), ("d_pairs", C.c_void_p), ("d_dp", C.c_void_p), ("d_runs", C.c_void_p),
                ("readStats", C.c_void_p), ("d_readStats", C.c_void_p)]


PE_PAIR_DTYPE = np.dtype([("pos1", np.uint32), ("pos2", np.uint32), ("insertion", np.uint32), ("strand1", np.uint8), ("mism1", np.uint8),
                          ("strand2", np.uint8), ("mism2", np.uint8), ("numPairs", np.uint32), ("numOptimal", np.uint32),
                          ("numSuboptimal", np.uint32), ("optimalTotal", np.int8), ("suboptimalTotal", np.int8), ("pad", np.uint16)])
PE_READ_STATS_DTYPE = np.dtype([("x0", np.uint32), ("x1", np.uint32), ("minMismatch", np.uint8), ("pad", np.uint8, (3,))])
PE_DP_DTYPE = np.dtype([("dpReadID", np.uint32), ("alignedPos", np.uint32), ("dpPos", np.uint32), ("score", np.int32),
                        ("numSameScore", np.uint32), ("runOffset", np.uint32), ("numRuns", np.uint16), ("alignedStrand", np.uint8),
                        ("alignedMismatches", np.uint8), ("dpStrand", np.uint8), ("leftOrRight", np.uint8), ("pad", np.uint8, (2,))])


def pe_params(num_mismatch=2, insert_low=200, insert_high=500, left_leg=1, right_leg=2, max_output_per_read=1000,
              max_hit_num_for_dp=None, keep_second_best=False, scores=(1, -2, -3, -1), cutoff=-1, soft_clip_left=3,
              soft_clip_right=8, max_windows=0, read_length=100, read_stats=False) -> PEParams:
    """Defaults as soap3_dp_pair_align sees them: Soap3MisMatchAllow 2 with DP (SOAP3-DP.cu:210-213), the ini's MaxOutputPerRead,
    getParameterForDefaultDP's maxHitNum for the read length, clips 3 / 8 (soap3-dp-module.cu:14)."""
    if max_hit_num_for_dp is None:
        max_hit_num_for_dp = getParameterForDP(2, read_length, read_length).paramRead[0].maxHitNum
    return PEParams(num_mismatch, insert_low, insert_high, left_leg, right_leg, max_output_per_read, max_hit_num_for_dp,
                    int(keep_second_best), DPScores(*scores), cutoff, soft_clip_left, soft_clip_right, max_windows, int(read_stats))


class PairAligner:
    """s3_pe_create / s3_pe_align: a batch of read pairs (reads 2p, 2p + 1 = the mates of pair p) from queries to
    pairings and mate-rescue alignments on the device."""

    def __init__(self, gpu_index: GpuIndex, max_reads: int, max_read_length: int, params: PEParams):
        lib = load_library()
        lib.s3_pe_create.restype = C.c_int
        lib.s3_pe_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(PEParams), C.POINTER(C.c_void_p)]
        lib.s3_pe_free.restype = None
        lib.s3_pe_free.argtypes = [C.c_void_p]
        for fn in (lib.s3_pe_align, lib.s3_pe_align_device):
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(PEResult)]
        lib.s3_pe_set_timing.restype = C.c_int
        lib.s3_pe_set_timing.argtypes = [C.c_void_p, C.c_int]
        lib.s3_pe_read_timing.restype = C.c_int
        lib.s3_pe_read_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        lib.s3_pe_dp.restype = C.c_void_p
        lib.s3_pe_dp.argtypes = [C.c_void_p]
        self.params = params
        out = C.c_void_p()
        _check(lib.s3_pe_create(gpu_index.handle, max_reads, max_read_length, C.byref(params), C.byref(out)), "s3_pe_create")
        self.handle = out

    def deep_dp(self, params, counts_only: bool = False):
        """s3_pe_deep_dp: DPForUnalignPairs2 for the both-unaligned pairs of the batch this handle has just aligned (on the device)"""
        lib = load_library()
        lib.s3_pe_deep_dp.restype = C.c_int
        lib.s3_pe_deep_dp.argtypes = [C.c_void_p, C.POINTER(StageParams), C.POINTER(_StageResult)]
        lib.s3_deep_dp_result_free.restype = None
        lib.s3_deep_dp_result_free.argtypes = [C.POINTER(_StageResult)]
        res = _StageResult()
        _check(lib.s3_pe_deep_dp(self.handle, C.byref(params), C.byref(res)), "s3_pe_deep_dp")
        return _stage_unpack(res, DEEP_HIT_DTYPE, counts_only, lib.s3_deep_dp_result_free)

    def align(self, queries, read_lengths, num_reads: int, word_per_query: int, copy: bool = True):
        """queries / read_lengths: host uint32 arrays (or raw host addresses).  -> dict of numpy arrays (copies unless copy=False)"""
        res = PEResult()
        q = queries.ctypes.data if hasattr(queries, "ctypes") else int(queries)
        l = read_lengths.ctypes.data if hasattr(read_lengths, "ctypes") else int(read_lengths)
        _check(load_library().s3_pe_align(self.handle, C.c_void_p(q), C.c_void_p(l), num_reads, word_per_query, C.byref(res)), "s3_pe_align")
        return self._unpack(res, copy)

    def prefetch(self, queries, read_lengths, num_reads: int, word_per_query: int):
        """s3_pe_prefetch: start the upload of the next batch (pinned host arrays or raw host addresses)"""
        lib = load_library()
        lib.s3_pe_prefetch.restype = C.c_int
        lib.s3_pe_prefetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        q = queries.ctypes.data if hasattr(queries, "ctypes") else int(queries)
        l = read_lengths.ctypes.data if hasattr(read_lengths, "ctypes") else int(read_lengths)
        _check(lib.s3_pe_prefetch(self.handle, C.c_void_p(q), C.c_void_p(l), num_reads, word_per_query), "s3_pe_prefetch")

    def align_device(self, d_queries: int, d_read_lengths: int, num_reads: int, word_per_query: int) -> PEResult:
        res = PEResult()
        _check(load_library().s3_pe_align_device(self.handle, C.c_void_p(d_queries), C.c_void_p(d_read_lengths), num_reads, word_per_query,
                                                 C.byref(res)), "s3_pe_align_device")
        return res

    @staticmethod
    def _unpack(res: PEResult, copy: bool):
        def view(ptr, dtype, n):
            if not ptr or n == 0:
                return np.zeros(0, dtype)
            buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
            a = np.frombuffer(buf, dtype=dtype, count=n)
            return a.copy() if copy else a
        P, M, R = int(res.numPairs), int(res.numWindows), int(res.numRuns)
        return {"route": view(res.route, np.uint8, P), "pairs": view(res.pairs, PE_PAIR_DTYPE, P), "dp": view(res.dp, PE_DP_DTYPE, M),
                "runs": view(res.runs, np.uint32, R), "num_ranges": int(res.numRanges), "num_occurrences": int(res.numOccurrences),
                "route_counts": list(res.routeCounts), "h2d_bytes": int(res.h2dBytes), "d2h_bytes": int(res.d2hBytes),
                "read_stats": view(res.readStats, PE_READ_STATS_DTYPE, 2 * P)}

    def set_timing(self, on: bool):
        _check(load_library().s3_pe_set_timing(self.handle, int(on)), "s3_pe_set_timing")

    def read_timing(self):
        ms = (C.c_float * 8)()
        _check(load_library().s3_pe_read_timing(self.handle, ms), "s3_pe_read_timing")
        return list(ms)

    @property
    def dp_handle(self):
        return C.c_void_p(load_library().s3_pe_dp(self.handle))

    def free(self):
        if self.handle:
            load_library().s3_pe_free(self.handle)
            self.handle = C.c_void_p(0)


def runs_to_cigar(runs: np.ndarray) -> str:
    """(length << 8 | op) runs of s3_pe_align -> the special CIGAR string the reference's encoder writes"""
    return "".join(f"{int(r) >> 8}{chr(int(r) & 0xFF)}" for r in runs)


# ---------------------------------------------------------------------------------------------------------------------
# which windows are aligned (s3_dp_make_windows: the deciding half of the three DP engines' pack() functions)
# ---------------------------------------------------------------------------------------------------------------------
WIN_SINGLE, WIN_HALF, WIN_PAIR_LEFT, WIN_PAIR_RIGHT = 1, 2, 3, 4


class WindowParams(C.Structure):
    _fields_ = [("insertLow", C.c_int32), ("insertHigh", C.c_int32), ("strandLeftLeg", C.c_int32), ("strandRightLeg", C.c_int32),
                ("softClipLeft", C.c_int32), ("softClipRight", C.c_int32), ("cutoffThreshold", C.c_int32 * 2), ("maxDNALength", C.c_uint32)]


def make_windows(gpu_index: GpuIndex, mode: int, params: WindowParams, read_lengths, read_ids, positions, strands=None, positions2=None,
                 left_scores=None, left_starts=None, left_hit_locs=None):
    """s3_dp_make_windows -> dict(candidate, read_ids, strands, left_or_right, dna_starts, dna_lengths, clip_lt, clip_rt, anchor_l, anchor_r, cutoffs)"""
    lib = load_library()
    lib.s3_dp_make_windows.restype = C.c_int
    lib.s3_dp_make_windows.argtypes = [C.c_void_p, C.c_int, C.POINTER(WindowParams), U32P, C.c_uint64, U32P, U32P, U32P, U8P, I32P, U32P, U32P, C.c_uint64,
                                       U32P, U32P, U8P, U8P, U32P, U32P, U32P, U32P, U32P, U32P, I32P, U64P]
    n = len(read_ids)
    cap = max(2 * n if mode == WIN_HALF else n, 1)
    u = lambda a: _u32(np.ascontiguousarray(a, np.uint32)) if a is not None else None
    keep = [np.ascontiguousarray(x, np.uint32) if x is not None else None for x in (read_lengths, read_ids, positions, positions2, left_starts, left_hit_locs)]
    st = np.ascontiguousarray(strands, np.uint8) if strands is not None else None
    sc = np.ascontiguousarray(left_scores, np.int32) if left_scores is not None else None
    o32 = [np.zeros(cap, np.uint32) for _ in range(8)]
    o8 = [np.zeros(cap, np.uint8) for _ in range(2)]
    cut = np.zeros(cap, np.int32)
    m = C.c_uint64()
    p = lambda a: _u32(a) if a is not None else None
    _check(lib.s3_dp_make_windows(gpu_index.handle, mode, C.byref(params), p(keep[0]), len(keep[0]), p(keep[1]), p(keep[2]), p(keep[3]),
                                  st.ctypes.data_as(U8P) if st is not None else None, sc.ctypes.data_as(I32P) if sc is not None else None,
                                  p(keep[4]), p(keep[5]), n, _u32(o32[0]), _u32(o32[1]), o8[0].ctypes.data_as(U8P), o8[1].ctypes.data_as(U8P),
                                  _u32(o32[2]), _u32(o32[3]), _u32(o32[4]), _u32(o32[5]), _u32(o32[6]), _u32(o32[7]), cut.ctypes.data_as(I32P),
                                  C.byref(m)), "s3_dp_make_windows")
    k = int(m.value)
    names = ("candidate", "read_ids", "dna_starts", "dna_lengths", "clip_lt", "clip_rt", "anchor_l", "anchor_r")
    out = {nm: a[:k].copy() for nm, a in zip(names, o32)}
    out.update(strands=o8[0][:k].copy(), left_or_right=o8[1][:k].copy(), cutoffs=cut[:k].copy())
    return out


# ---------------------------------------------------------------------------------------------------------------------
# single-end batch on the device (s3_se_*): the in-memory alignSingleR (soap3-dp-module.cu:62)
# ---------------------------------------------------------------------------------------------------------------------
class SEParams(C.Structure):
    _fields_ = [("numMismatch", C.c_uint32), ("maxOutputPerRead", C.c_uint32), ("reportBest", C.c_int32),
                ("longReadMode", C.c_int32), ("onlyKeepBest", C.c_int32), ("minSeedMismatch", C.c_int32), ("doubleAllowance", C.c_int32)]


class SEResult(C.Structure):
    _fields_ = [("numReads", C.c_uint64), ("numRanges", C.c_uint64), ("numOccurrences", C.c_uint64), ("h2dBytes", C.c_uint64), ("d2hBytes", C.c_uint64),
                ("occOffsets", C.c_void_p), ("positions", C.c_void_p), ("occFlags", C.c_void_p), ("readFlags", C.c_void_p),
                ("d_occOffsets", C.c_void_p), ("d_positions", C.c_void_p), ("d_occFlags", C.c_void_p), ("d_readFlags", C.c_void_p)]


class SingleAligner:
    """s3_se_create / s3_se_align: alignSingleR's results (occurrences per read) for a batch, on the device."""

    def __init__(self, gpu_index: GpuIndex, max_reads: int, num_mismatch: int = 2, max_output_per_read: int = 1000, report_best: bool = False,
                 long_read_mode: bool = False, only_keep_best: bool = False, min_seed_mismatch: int = 0, double_allowance: bool = False):
        lib = load_library()
        lib.s3_se_create.restype = C.c_int
        lib.s3_se_create.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(SEParams), C.POINTER(C.c_void_p)]
        lib.s3_se_free.restype = None
        lib.s3_se_free.argtypes = [C.c_void_p]
        for fn in (lib.s3_se_align, lib.s3_se_align_device):
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(SEResult)]
        par = SEParams(num_mismatch, max_output_per_read, int(report_best), int(long_read_mode), int(only_keep_best), min_seed_mismatch, int(double_allowance))
        out = C.c_void_p()
        _check(lib.s3_se_create(gpu_index.handle, max_reads, C.byref(par), C.byref(out)), "s3_se_create")
        self.handle = out

    def align(self, queries, read_lengths, num_reads: int, word_per_query: int, copy: bool = True):
        """-> dict of numpy arrays (copies; with copy=False views of the handle's host buffers, valid until the next call)"""
        res = SEResult()
        q = queries.ctypes.data if hasattr(queries, "ctypes") else int(queries)
        l = read_lengths.ctypes.data if hasattr(read_lengths, "ctypes") else int(read_lengths)
        _check(load_library().s3_se_align(self.handle, C.c_void_p(q), C.c_void_p(l), num_reads, word_per_query, C.byref(res)), "s3_se_align")

        def view(ptr, dtype, n):
            if not ptr or n == 0:
                return np.zeros(0, dtype)
            buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
            a = np.frombuffer(buf, dtype=dtype, count=n)
            return a.copy() if copy else a
        n, t = int(res.numReads), int(res.numOccurrences)
        return {"occ_offsets": view(res.occOffsets, np.uint32, n + 1), "positions": view(res.positions, np.uint32, t),
                "occ_flags": view(res.occFlags, np.uint8, 2 * t).reshape(-1, 2), "read_flags": view(res.readFlags, np.uint8, n),
                "num_ranges": int(res.numRanges), "h2d_bytes": int(res.h2dBytes), "d2h_bytes": int(res.d2hBytes)}

    def align_device(self, d_queries: int, d_read_lengths: int, num_reads: int, word_per_query: int) -> SEResult:
        res = SEResult()
        _check(load_library().s3_se_align_device(self.handle, C.c_void_p(d_queries), C.c_void_p(d_read_lengths), num_reads, word_per_query,
                                                 C.byref(res)), "s3_se_align_device")
        return res

    def free(self):
        if self.handle:
            load_library().s3_se_free(self.handle)
            self.handle = C.c_void_p(0)


def validate_alignments(gpu_index: GpuIndex, queries, read_lengths, num_reads: int, word_per_query: int, occ_offsets, positions, occ_flags,
                        only_keep_best=False, min_seed_mismatch=0, double_allowance=False, max_hit_num=1000):
    """s3_validate_alignments (validateAlignments, CPUfunctions.cpp:1129): -> (counts[num_reads], positions, occ_flags updated copies)"""
    lib = load_library()
    lib.s3_validate_alignments.restype = C.c_int
    lib.s3_validate_alignments.argtypes = [C.c_void_p, U32P, U32P, C.c_uint64, C.c_uint32, U32P, U32P, U8P, C.c_int, C.c_int, C.c_int, C.c_int, U32P]
    pos = np.ascontiguousarray(positions, np.uint32).copy()
    fl = np.ascontiguousarray(occ_flags, np.uint8).copy()
    off = np.ascontiguousarray(occ_offsets, np.uint32)
    counts = np.zeros(max(num_reads, 1), np.uint32)
    if len(pos) == 0:
        pos, fl = np.zeros(1, np.uint32), np.zeros(2, np.uint8)
    _check(lib.s3_validate_alignments(gpu_index.handle, _u32(queries), _u32(read_lengths), num_reads, word_per_query, _u32(off), _u32(pos),
                                      fl.ctypes.data_as(U8P), int(only_keep_best), min_seed_mismatch, int(double_allowance), max_hit_num, _u32(counts)),
           "s3_validate_alignments")
    return counts[:num_reads], pos, fl


class SeedSearchResult(C.Structure):
    _fields_ = [("numSeeds", C.c_uint64), ("total", C.c_uint64), ("offsets", U64P), ("saL", U32P), ("saR", U32P), ("strand", U8P), ("status", U8P)]


def seed_search(gpu_index: GpuIndex, seeds: np.ndarray, seed_lengths: np.ndarray, num_seeds: int, word_per_seed: int, max_hit_num):
    """s3_seed_search (single_1_mismatch_alignment2, alignment.cu:1839): -> offsets[numSeeds+1], saL, saR, strand, status[numSeeds]"""
    lib = load_library()
    lib.s3_seed_search.restype = C.c_int
    lib.s3_seed_search.argtypes = [C.c_void_p, U32P, U32P, C.c_uint64, C.c_uint32, U32P, C.POINTER(SeedSearchResult)]
    lib.s3_seed_search_result_free.restype = None
    lib.s3_seed_search_result_free.argtypes = [C.POINTER(SeedSearchResult)]
    mh = np.ascontiguousarray(np.broadcast_to(np.asarray(max_hit_num, np.uint32), (max(num_seeds, 1),)))
    res = SeedSearchResult()
    _check(lib.s3_seed_search(gpu_index.handle, _u32(seeds), _u32(seed_lengths), num_seeds, word_per_seed, _u32(mh), C.byref(res)), "s3_seed_search")
    t = int(res.total)
    cp = lambda p, n, dt: np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)
    out = (cp(res.offsets, num_seeds + 1, np.uint64), cp(res.saL, t, np.uint32), cp(res.saR, t, np.uint32), cp(res.strand, t, np.uint8),
           cp(res.status, num_seeds, np.uint8))
    lib.s3_seed_search_result_free(C.byref(res))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the DP stages that start from seeds (DPForUnalignSingle2 / DPForUnalignPairs2)
# ---------------------------------------------------------------------------------------------------------------------
class StageParams(C.Structure):
    _fields_ = [("insertLow", C.c_int32), ("insertHigh", C.c_int32), ("strandLeftLeg", C.c_int32), ("strandRightLeg", C.c_int32), ("scores", DPScores),
                ("isDefaultThreshold", C.c_int32), ("dpScoreThreshold", C.c_int32), ("softClipLeft", C.c_int32), ("softClipRight", C.c_int32)]


DP_HIT_DTYPE = np.dtype([("readID", np.uint32), ("pos", np.uint32), ("score", np.int32), ("numSameScore", np.uint32), ("runOffset", np.uint32),
                         ("numRuns", np.uint16), ("strand", np.uint8), ("pad", np.uint8)])
DEEP_HIT_DTYPE = np.dtype([("readID", np.uint32), ("pos1", np.uint32), ("pos2", np.uint32), ("score1", np.int32), ("score2", np.int32),
                           ("numSame1", np.uint32), ("numSame2", np.uint32), ("runOffset1", np.uint32), ("runOffset2", np.uint32),
                           ("numRuns1", np.uint16), ("numRuns2", np.uint16), ("strand1", np.uint8), ("strand2", np.uint8), ("pad", np.uint8, (2,))])


class _StageResult(C.Structure):
    _fields_ = [("numIn", C.c_uint64), ("numSeeds", C.c_uint64), ("numCandidates", C.c_uint64), ("numHits", C.c_uint64), ("numRuns", C.c_uint64),
                ("numUnseeded", C.c_uint64), ("hits", C.c_void_p), ("runs", C.c_void_p), ("unseeded", C.c_void_p)]


def stage_params(insert_low=200, insert_high=500, left_leg=1, right_leg=2, scores=(1, -2, -3, -1), default_threshold=True, threshold=0,
                 soft_clip_left=3, soft_clip_right=8) -> StageParams:
    return StageParams(insert_low, insert_high, left_leg, right_leg, DPScores(*scores), int(default_threshold), threshold, soft_clip_left, soft_clip_right)


def _stage_align(fn_name, free_name, dtype, gpu_index, queries, read_lengths, num_reads, word_per_query, ids, params, counts_only=False):
    lib = load_library()
    fn, fr = getattr(lib, fn_name), getattr(lib, free_name)
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, U32P, U32P, C.c_uint64, C.c_uint32, U32P, C.c_uint64, C.POINTER(StageParams), C.POINTER(_StageResult)]
    fr.restype = None
    fr.argtypes = [C.POINTER(_StageResult)]
    ids = np.ascontiguousarray(ids, np.uint32)
    res = _StageResult()
    _check(fn(gpu_index.handle, _u32(queries), _u32(read_lengths), num_reads, word_per_query, _u32(ids), len(ids), C.byref(params), C.byref(res)), fn_name)
    return _stage_unpack(res, dtype, counts_only, fr)


def _stage_unpack(res, dtype, counts_only, fr):
    def view(ptr, dt, n):
        if not ptr or n == 0:
            return np.zeros(0, dt)
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt, count=n).copy()
    if counts_only:                    # (timing loops: the C entry has done all its work, the arrays are not copied into numpy)
        out = {"num_hits": int(res.numHits), "num_runs": int(res.numRuns), "num_unseeded": int(res.numUnseeded), "num_seeds": int(res.numSeeds),
               "num_candidates": int(res.numCandidates), "num_input": int(res.numIn)}
    else:
        out = {"hits": view(res.hits, dtype, int(res.numHits)), "runs": view(res.runs, np.uint32, int(res.numRuns)),
               "unseeded": view(res.unseeded, np.uint32, int(res.numUnseeded)), "num_seeds": int(res.numSeeds), "num_candidates": int(res.numCandidates),
               "num_input": int(res.numIn)}
    fr(C.byref(res))
    return out


def single_dp_align(gpu_index: GpuIndex, queries, read_lengths, num_reads: int, word_per_query: int, read_ids, params: StageParams, counts_only=False):
    """s3_single_dp_align (DPForUnalignSingle2, DV-DPForSingleReads.cu:155)"""
    return _stage_align("s3_single_dp_align", "s3_single_dp_result_free", DP_HIT_DTYPE, gpu_index, queries, read_lengths, num_reads, word_per_query, read_ids, params,
                        counts_only)


def deep_dp_align(gpu_index: GpuIndex, queries, read_lengths, num_reads: int, word_per_query: int, pair_read_ids, params: StageParams, counts_only=False):
    """s3_deep_dp_align (DPForUnalignPairs2, DV-DPForBothUnalign.cu:245)"""
    return _stage_align("s3_deep_dp_align", "s3_deep_dp_result_free", DEEP_HIT_DTYPE, gpu_index, queries, read_lengths, num_reads, word_per_query, pair_read_ids, params,
                        counts_only)


def set_l2_persist(gpu_index: GpuIndex, region: int, window_bytes: int = 0, persist_bytes: int = 0):
    """s3_index_set_l2_persist (measurement knob): access-policy window over one of the index arrays; region 0 resets"""
    lib = load_library()
    lib.s3_index_set_l2_persist.restype = C.c_int
    lib.s3_index_set_l2_persist.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t]
    _check(lib.s3_index_set_l2_persist(gpu_index.handle, region, window_bytes, persist_bytes), "s3_index_set_l2_persist")


def index_load(prefix: str, with_text: bool = True, device: int = 0) -> GpuIndex:
    """s3_index_load: the reference's index files (<prefix>.bwt / .fmv.gpu / .rev.bwt / .rev.fmv.gpu [/ .sa / .pac]) straight from disk"""
    lib = load_library()
    lib.s3_index_load.restype = C.c_int
    lib.s3_index_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    out = C.c_void_p()
    _check(lib.s3_index_load(prefix.encode(), int(with_text), device, C.byref(out)), "s3_index_load")
    n = int(np.fromfile(prefix + ".bwt", dtype=np.uint32, count=5)[4])
    return GpuIndex(out.value, n)


def index_clone(gpu_index: GpuIndex) -> GpuIndex:
    """s3_index_clone: a second handle on the same device arrays with its own stream and scratch (one host thread per handle)"""
    lib = load_library()
    lib.s3_index_clone.restype = C.c_int
    lib.s3_index_clone.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    out = C.c_void_p()
    _check(lib.s3_index_clone(gpu_index.handle, C.byref(out)), "s3_index_clone")
    return GpuIndex(out.value, gpu_index.text_length)


# ---- SAM text of whole batches (s3_sam_*_batch_text): the structs of include/soap3dp_b200.h and one wrapper per entry ----------------
class SamSegment(C.Structure):
    _fields_ = [("startPos", C.c_uint32), ("chrID", C.c_uint32), ("correction", C.c_uint32)]


class SamGenome(C.Structure):
    _fields_ = [("packedDNA", C.POINTER(C.c_uint32)), ("dnaLength", C.c_uint32), ("segments", C.POINTER(SamSegment)), ("numSegments", C.c_uint32),
                ("ambiguityMap", C.POINTER(C.c_uint32)), ("chrEndPos", C.POINTER(C.c_uint32)), ("numChr", C.c_uint32), ("chrNames", C.POINTER(C.c_char_p))]


class SamConfig(C.Structure):
    _fields_ = [("alignmentType", C.c_int32), ("bwaLikeScore", C.c_int32), ("dpMatchScore", C.c_int32), ("dpMisMatchScore", C.c_int32),
                ("isFastq", C.c_int32), ("maxMAPQ", C.c_int32), ("minMAPQ", C.c_int32), ("isPrintMDNM", C.c_int32), ("outputXAZTag", C.c_int32),
                ("peMaxOutputPerPair", C.c_uint32), ("readGroup", C.c_char_p)]


class SamReadsStruct(C.Structure):
    _fields_ = [("bases", C.POINTER(C.c_uint8)), ("qualities", C.c_char_p), ("rowBytes", C.c_uint32), ("readLengths", C.POINTER(C.c_uint32)),
                ("names", C.POINTER(C.c_char_p))]


def _as_u32(a):
    return np.ascontiguousarray(a, np.uint32)


class SamGenomeDesc:
    """s3_sam_genome over numpy arrays (kept alive here): packed text (hsp->packedDNA), the translate table as rows (startPos, chrID, correction),
    the ambiguity map, the chromosomes' last positions and names"""

    def __init__(self, packed_dna, dna_length: int, segments, ambiguity_map, chr_end_pos, chr_names):
        self.pac, self.amb, self.end = _as_u32(packed_dna), _as_u32(ambiguity_map), _as_u32(chr_end_pos)
        seg = np.asarray(segments, np.int64).reshape(-1, 3)
        self.segs = (SamSegment * len(seg))(*[SamSegment(int(a) & 0xFFFFFFFF, int(b), int(c) & 0xFFFFFFFF) for a, b, c in seg])
        self.names = (C.c_char_p * len(chr_names))(*[n if isinstance(n, bytes) else n.encode() for n in chr_names])
        p = C.POINTER(C.c_uint32)
        self.struct = SamGenome(self.pac.ctypes.data_as(p), dna_length, self.segs, len(seg), self.amb.ctypes.data_as(p), self.end.ctypes.data_as(p), len(chr_names), self.names)


class SamReads:
    """s3_sam_reads: bases (one code per byte) and Phred qualities as (numReads, rowBytes) uint8 arrays, read lengths, names"""

    def __init__(self, bases, qualities, read_lengths, names):
        self.bases, self.quals, self.lens = np.ascontiguousarray(bases, np.uint8), np.ascontiguousarray(qualities, np.uint8), _as_u32(read_lengths)
        if self.bases.ndim != 2 or self.bases.shape != self.quals.shape or len(self.lens) != len(self.bases) or len(names) != len(self.bases):
            raise ValueError("SamReads: bases and qualities must be (numReads, rowBytes) arrays with one length and one name per read")
        self.names = (C.c_char_p * len(names))(*[n if isinstance(n, bytes) else n.encode() for n in names])
        self.num = len(self.bases)
        self.struct = SamReadsStruct(self.bases.ctypes.data_as(C.POINTER(C.c_uint8)), C.cast(self.quals.ctypes.data, C.c_char_p), self.bases.shape[1],
                                     self.lens.ctypes.data_as(C.POINTER(C.c_uint32)), self.names)


def _sam_text(name: str, genome: SamGenomeDesc, config: SamConfig, reads: SamReads, args, num_threads: int) -> bytes:
    lib = load_library()
    fn = getattr(lib, name)
    fn.restype = C.c_int
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    text, size = C.c_void_p(), C.c_uint64()
    _check(fn(C.byref(genome.struct), C.byref(config), C.byref(reads.struct), C.c_uint64(reads.num), *args, C.c_uint32(num_threads), C.byref(text), C.byref(size)), name)
    out = C.string_at(text.value, size.value)
    lib.s3_free(text)
    return out


def _csr(occ_offsets, positions, occ_flags):
    off, pos, fl = _as_u32(occ_offsets), _as_u32(positions), np.ascontiguousarray(occ_flags, np.uint8)
    p = C.POINTER(C.c_uint32)
    return (off, pos, fl), [off.ctypes.data_as(p), pos.ctypes.data_as(p), fl.ctypes.data_as(C.POINTER(C.c_uint8))]


def _hits(hits, dtype, runs):
    h, r = np.ascontiguousarray(hits, dtype), _as_u32(runs)
    return (h, r), [h.ctypes.data_as(C.c_void_p), C.c_uint64(len(h)), r.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(len(r))]


def _stats(read_stats):
    if read_stats is None:
        return None, None
    s = np.ascontiguousarray(read_stats, PE_READ_STATS_DTYPE)
    return s, s.ctypes.data_as(C.c_void_p)


def sam_single_batch_text(genome, config, reads, occ_offsets, positions, occ_flags, num_threads=0) -> bytes:
    """s3_sam_single_batch_text over the arrays SingleAligner.align returns: one SAM line per read, unmapped ones included"""
    keep, a = _csr(occ_offsets, positions, occ_flags)
    return _sam_text("s3_sam_single_batch_text", genome, config, reads, a, num_threads)


def sam_single_dp_batch_text(genome, config, reads, hits, runs, scores: DPScores, cutoff: int, num_threads=0) -> bytes:
    """s3_sam_single_dp_batch_text over single_dp_align's hits + runs: one line per read that has a hit"""
    keep, a = _hits(hits, DP_HIT_DTYPE, runs)
    return _sam_text("s3_sam_single_dp_batch_text", genome, config, reads, a + [scores, C.c_int32(cutoff)], num_threads)


def sam_paired_batch_text(genome, config, reads, route, pairs, read_stats, num_threads=0) -> bytes:
    """s3_sam_paired_batch_text over PairAligner.align's route / pairs / read_stats: two lines per paired pair with one valid pairing"""
    rt, pr = np.ascontiguousarray(route, np.uint8), np.ascontiguousarray(pairs, PE_PAIR_DTYPE)
    keep, st = _stats(read_stats)
    return _sam_text("s3_sam_paired_batch_text", genome, config, reads, [rt.ctypes.data_as(C.POINTER(C.c_uint8)), pr.ctypes.data_as(C.c_void_p), C.c_uint64(len(pr)), st], num_threads)


def sam_pair_dp_batch_text(genome, config, reads, dp, runs, scores: DPScores, read_stats=None, num_threads=0) -> bytes:
    """s3_sam_pair_dp_batch_text over PairAligner.align's rescue records + runs: two lines per pair with a successful rescue"""
    keep, a = _hits(dp, PE_DP_DTYPE, runs)
    keep2, st = _stats(read_stats)
    return _sam_text("s3_sam_pair_dp_batch_text", genome, config, reads, a + [scores, st], num_threads)


def sam_deep_dp_batch_text(genome, config, reads, hits, runs, scores: DPScores, read_stats=None, num_threads=0) -> bytes:
    """s3_sam_deep_dp_batch_text over deep_dp_align / PairAligner.deep_dp hits + runs: two lines per pair with a hit"""
    keep, a = _hits(hits, DEEP_HIT_DTYPE, runs)
    keep2, st = _stats(read_stats)
    return _sam_text("s3_sam_deep_dp_batch_text", genome, config, reads, a + [scores, st], num_threads)


def sam_unpaired_batch_text(genome, config, reads, occ_offsets, positions, occ_flags, pair_ids, max_output_per_read=1000, num_threads=0) -> bytes:
    """s3_sam_unpaired_batch_text: the named pairs, each read on its own from a CSR of occurrences over all reads of the batch"""
    keep, a = _csr(occ_offsets, positions, occ_flags)
    ids = _as_u32(pair_ids)
    return _sam_text("s3_sam_unpaired_batch_text", genome, config, reads, a + [ids.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(len(ids)), C.c_uint32(max_output_per_read)], num_threads)


def sam_unpaired_dp_batch_text(genome, config, reads, occ_offsets, positions, occ_flags, hits, runs, scores: DPScores, cutoff: int, pair_ids, num_threads=0) -> bytes:
    """s3_sam_unpaired_dp_batch_text: the named pairs after DP, per read its single-read DP hits when it has any, else its occurrences"""
    keep, a = _csr(occ_offsets, positions, occ_flags)
    keep2, b = _hits(hits, DP_HIT_DTYPE, runs)
    ids = _as_u32(pair_ids)
    return _sam_text("s3_sam_unpaired_dp_batch_text", genome, config, reads, a + b + [scores, C.c_int32(cutoff), ids.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(len(ids))], num_threads)
