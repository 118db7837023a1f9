"""CPU tier: the walk that both kernels of s3_pair_occurrences run (soap3-dp_b200/csrc/s3_pair_walk.cuh) compiled with
the host compiler and driven like csrc/s3_pair.cu drives it (tests/native/pair_walk_harness.cpp), against the pairing
oracle -- which the CPU tier pins against the reference's PEMappingOccurrences (test_cpu_oracle_vs_ref.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers

NATIVE = os.path.join(os.path.dirname(__file__), "native")


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(NATIVE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpair_walk_harness.so")
    src = os.path.join(NATIVE, "pair_walk_harness.cpp")
    hdr = os.path.join(helpers.ROOT, "soap3-dp_b200", "csrc", "s3_pair_walk.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-x", "c++", src,
                               "-I", os.path.dirname(hdr), "-o", so])
    lib = C.CDLL(so)
    U8 = C.POINTER(C.c_uint8)
    U32P, U64P = helpers.U32P, helpers.U64P
    lib.harness_pair_occurrences.restype = C.c_uint64
    lib.harness_pair_occurrences.argtypes = [U32P, U8, U8, U64P, U32P, U8, U8, U64P, U32P, C.c_uint64, C.c_int32, C.c_int32, C.c_int, C.c_int,
                                             C.c_int, U64P, U32P, U32P, U32P, U8, C.c_uint64, U32P, U32P, U32P]
    return lib


def run_harness(lib, lists, pl, lb, ub, left_leg, right_leg, report_one):
    U8 = C.POINTER(C.c_uint8)
    u32p, U64P = helpers.u32p, helpers.U64P
    p1, s1, m1, o1, p2, s2, m2, o2 = lists
    npairs = len(o1) - 1
    pl = np.ascontiguousarray(pl, np.uint32)
    offs = np.zeros(npairs + 1, np.uint64)
    opt, sub, stats = np.zeros(npairs, np.uint32), np.zeros(npairs, np.uint32), np.zeros((npairs, 32), np.uint32)
    args = (u32p(p1), s1.ctypes.data_as(U8), m1.ctypes.data_as(U8), o1.ctypes.data_as(U64P), u32p(p2), s2.ctypes.data_as(U8),
            m2.ctypes.data_as(U8), o2.ctypes.data_as(U64P), u32p(pl), npairs, lb, ub, left_leg, right_leg, int(report_one))
    total = lib.harness_pair_occurrences(*args, offs.ctypes.data_as(U64P), None, None, None, None, 0, None, None, None)
    a, b, ins, fl = np.zeros(total, np.uint32), np.zeros(total, np.uint32), np.zeros(total, np.uint32), np.zeros((total, 4), np.uint8)
    lib.harness_pair_occurrences(*args, offs.ctypes.data_as(U64P), u32p(a), u32p(b), u32p(ins), fl.ctypes.data_as(U8), total,
                                 u32p(opt), u32p(sub), u32p(stats))
    return dict(offsets=offs, pos1=a, pos2=b, insertion=ins, flags=fl, optimal=opt, suboptimal=sub, stats=stats)


def test_the_kernels_walk_gives_the_oracles_pairs(harness):
    rng = np.random.default_rng(17)
    found = 0
    for legs in ((1, 2), (2, 1), (1, 1), (2, 2)):
        for report_one in (False, True):
            for near_edges in (False, True):
                lists = helpers.make_occurrence_lists(rng, 500, max_occ=16, near_edges=near_edges)
                pl = rng.integers(60, 151, 500).astype(np.uint32)
                for lb, ub in ((200, 500), (1, 300)):
                    got = run_harness(harness, lists, pl, lb, ub, *legs, report_one)
                    want = helpers.oracle_pair_occurrences(lists, pl, lb, ub, *legs, report_one)
                    assert helpers.same_pairing(got, want), (legs, report_one, near_edges, lb, ub)
                    found += len(want["pos1"])
    assert found > 10000


def test_the_kernels_walk_on_empty_batches_and_lists(harness):
    z32, z8, z64 = np.zeros(0, np.uint32), np.zeros(0, np.uint8), np.zeros(4, np.uint64)
    lists = (z32, z8, z8, z64, z32, z8, z8, z64)                   # three read pairs, no occurrences at all
    got = run_harness(harness, lists, np.full(3, 100, np.uint32), 200, 500, 1, 2, False)
    assert got["offsets"].tolist() == [0, 0, 0, 0] and (got["optimal"] == 0xFFFFFFFF).all() and (got["suboptimal"] == 0xFFFFFFFF).all()
    want = helpers.oracle_pair_occurrences(lists, np.full(3, 100, np.uint32), 200, 500, 1, 2, False)
    assert helpers.same_pairing(got, want)


def test_the_retain_kernels_step_gives_the_oracles_lists(harness):
    """csrc/s3_retain_walk.cuh (closed form: minima first, then one pass) against the oracle's statement-by-statement
    restatement of the reference's running-minimum filters"""
    rng = np.random.default_rng(41)
    for mode, caps in ((0, (0,)), (1, (1, 2, 3, 7, 40, 10000)), (2, (0,))):
        for cap in caps:
            for max_sa, max_occ in ((7, 7), (1, 9), (9, 1), (3, 3)):
                lists = helpers.make_hit_lists(rng, 3000, max_sa=max_sa, max_occ=max_occ)
                got = helpers.run_retain(harness.harness_retain_best, lists, mode, cap)
                want = helpers.oracle_retain_best(lists, mode, cap)
                assert helpers.same_retained(got, want), (mode, cap, max_sa, max_occ)
