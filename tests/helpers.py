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
