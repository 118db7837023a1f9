"""CPU tier: the mapping-quality entries (s3_mapq_*, host scalars) against the reference's own nine MAPQ functions and
tables compiled into oracle/_ref/libref_mapq.so, over random and exhaustive small argument ranges, and against a committed
fixture generated from that library (tests/golden/mapq_golden.json)."""
import ctypes as C
import itertools
import json
import os

import numpy as np
import pytest

import helpers
from soap3dp_b200 import api

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "mapq_golden.json")
NAMES = ("unique", "bwa_single", "single", "single_dp", "pair_end", "unique_dp", "pair_end_dp", "of_pair")


def load_ref():
    path = os.path.join(helpers.ROOT, "oracle", "_ref", "libref_mapq.so")
    return C.CDLL(path) if os.path.exists(path) else None


def cases(rng, n=4000):
    """name -> list of argument tuples (ints); ranges cover every table row, both branches of every test"""
    r = lambda lo, hi: int(rng.integers(lo, hi))
    out = {k: [] for k in NAMES + ("bwa_pair",)}
    for _ in range(n):
        mx, mn = r(30, 61), r(0, 4)
        out["unique"].append((r(0, 3), r(0, 9), r(-5, 70), mx, mn))
        out["bwa_single"].append((r(0, 4), r(0, 400)))
        out["single"].append((r(0, 9), r(-5, 70), r(0, 3), r(0, 3), mx, mn, r(0, 2)))
        best = r(31, 151)
        out["single_dp"].append((r(best, 201), r(-5, 70), r(0, 3), r(0, 2), r(0, 140), best, r(31, best + 1), mx, mn, 30, r(0, 2)))
        out["bwa_pair"].append((r(0, 3), r(0, 300), r(0, 3), r(0, 300), r(30, 200), r(0, 3), r(0, 200), r(0, 400), r(50, 151), r(50, 151)))
        out["pair_end"].append((r(0, 9), r(-5, 70), r(0, 3), r(0, 4), r(0, 2), r(0, 3), mx, mn))
        maxdp = r(50, 201)
        out["unique_dp"].append((r(0, 3), r(20, maxdp + 1), maxdp, r(-5, 70), mx, mn))
        b2 = r(20, maxdp + 1)
        out["pair_end_dp"].append((r(20, maxdp + 1), maxdp, r(-5, 70), r(0, 3), r(0, 4), b2, r(0, b2 + 1), r(0, 2), r(0, 3), mx, mn))
        out["of_pair"].append((r(0, 61), r(0, 61)))
    # every x1 of the penalty table, every quality of the penalty row
    for x1, q in itertools.product(range(0, 130, 1), (0, 7, 19, 40, 55)):
        out["single_dp"].append((150, q, 1, 0, x1, 120, 90, 40, 1, 30, 0))
    for q in range(-2, 45):
        out["single_dp"].append((100, q, 1, 0, 0, 88, 50, 40, 1, 30, 0))
    return out


def product(name, args):
    lib = api.load_library()
    if name == "bwa_pair":
        a, b = C.c_int32(0), C.c_int32(0)
        lib.s3_mapq_bwa_pair(*args, C.byref(a), C.byref(b))
        return [a.value, b.value]
    return getattr(lib, "s3_mapq_" + name)(*args)


def reference(ref, name, args):
    if name == "bwa_pair":
        a, b = C.c_int(0), C.c_int(0)
        ref.ref_mapq_bwa_pair(*args, C.byref(a), C.byref(b))
        return [a.value, b.value]
    return getattr(ref, "ref_mapq_" + name)(*args)


@pytest.mark.skipif(load_ref() is None, reason="oracle/_ref/libref_mapq.so not built")
def test_mapq_entries_match_the_reference_functions():
    ref = load_ref()
    for name, argl in cases(np.random.default_rng(12)).items():
        seen = set()
        for args in argl:
            want = reference(ref, name, args)
            assert product(name, args) == want, (name, args)
            seen.add(tuple(want) if isinstance(want, list) else want)
        assert len(seen) > 2, name                      # not one constant answer


def test_mapq_entries_match_the_golden_fixture():
    g = json.load(open(GOLDEN))
    total = 0
    for name, rows in g["cases"].items():
        for args, want in rows:
            assert product(name, tuple(args)) == want, (name, args)
            total += 1
    assert total > 1000
