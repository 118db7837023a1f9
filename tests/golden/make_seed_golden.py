"""Generates tests/golden/seed_golden.json from the REFERENCE's own seed-merge code (oracle/_ref/libref_seed.so and
libref_seed_pair.so, built by oracle/build_ref.sh from DV-DPfunctions.h:60-95 and DV-DPfunctions.cu:1101-1141,2626-2653,
2780-2880): seeded hit sets -> sha256 of the candidate arrays (and the first rows spelled out).  Run in the build container.

Usage:  python tests/golden/make_seed_golden.py
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
U, I = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def single_case(seed, n, nreads, span):
    rng = np.random.default_rng(seed)
    rid = rng.integers(0, nreads, n).astype(np.uint32)
    x = rng.integers(0, span, n).astype(np.uint32)
    st = rng.integers(1, 3, n).astype(np.int32)
    off = rng.integers(0, 60, n).astype(np.uint32)
    return rid, x, st, off, np.full(n, 28, np.uint32), np.full(n, 100, np.uint32)


def pair_case(seed, n0, n1, npairs, span):
    rng = np.random.default_rng(seed)
    sides = []
    for n in (n0, n1):
        sides.append(((rng.integers(0, npairs, n) * 2).astype(np.uint32), rng.integers(0, span, n).astype(np.uint32),
                      rng.integers(1, 3, n).astype(np.int32), rng.integers(0, 60, n).astype(np.uint32),
                      np.full(n, 28, np.uint32), rng.choice([100, 100, 150, 75], n).astype(np.uint32)))
    lens = rng.choice([100, 100, 150, 75], npairs * 2 + 2).astype(np.uint32)
    return sides, lens


SINGLE = [(1, 5000, 40, 3000), (2, 300, 3, 200), (3, 20000, 2000, 100000)]
PAIR = [(11, 3000, 3000, 60, 5000), (12, 500, 800, 5, 1500), (13, 20000, 15000, 1500, 200000)]


def main():
    ref1 = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_seed.so"))
    ref2 = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_seed_pair.so"))
    ref1.ref_seed_sort_merge.restype = C.c_uint32
    ref1.ref_seed_sort_merge.argtypes = [U, U, I, C.c_uint32, U, U, I]
    ref2.ref_seed_pair_merge.restype = C.c_uint32
    ref2.ref_seed_pair_merge.argtypes = [U, U, C.c_uint32, U, U, C.c_uint32, U, C.c_int, C.c_int, C.c_int, C.c_int, U, U, U]

    def u(a):
        return a.ctypes.data_as(U)
    out = {"single": [], "pair": []}
    for seed, n, nreads, span in SINGLE:
        rid, x, st, off, sl, rl = single_case(seed, n, nreads, span)
        est = np.where(st == 1, x - off, x + sl + off - rl).astype(np.uint32)
        o = [np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.int32)]
        m = ref1.ref_seed_sort_merge(u(rid), u(est), st.ctypes.data_as(I), n, u(o[0]), u(o[1]), o[2].ctypes.data_as(I))
        out["single"].append({"case": [seed, n, nreads, span], "candidates": int(m), "readIDs": sha(o[0][:m]), "positions": sha(o[1][:m]),
                              "strands": sha(o[2][:m]), "first": [[int(a), int(b), int(c)] for a, b, c in zip(o[0][:8], o[1][:8], o[2][:8])]})
    for seed, n0, n1, npairs, span in PAIR:
        sides, lens = pair_case(seed, n0, n1, npairs, span)
        keys, poss = [], []
        for rid, x, st, off, sl, rl in sides:
            si = (st - 1).astype(np.uint32)
            poss.append(np.where(si == 0, x - off, x + sl + off - rl).astype(np.uint32))
            keys.append((rid | (si << 31)).astype(np.uint32))
        for legs in ((1, 2), (2, 1), (1, 1), (2, 2)):
            cap = (n0 + n1) * 50 + 10
            o = [np.zeros(cap, np.uint32) for _ in range(3)]
            m = ref2.ref_seed_pair_merge(u(keys[0]), u(poss[0]), n0, u(keys[1]), u(poss[1]), n1, u(lens), 200, 500, legs[0], legs[1],
                                         u(o[0]), u(o[1]), u(o[2]))
            out["pair"].append({"case": [seed, n0, n1, npairs, span], "legs": list(legs), "candidates": int(m), "readIDLeft": sha(o[0][:m]),
                                "posLeft": sha(o[1][:m]), "posRight": sha(o[2][:m]),
                                "first": [[int(a), int(b), int(c)] for a, b, c in zip(o[0][:6], o[1][:6], o[2][:6])]})
    json.dump(out, open(os.path.join(HERE, "seed_golden.json"), "w"), indent=1)
    print("wrote seed_golden.json:", len(out["single"]), "single-end cases,", len(out["pair"]), "paired-end cases")


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    main()
