"""CPU tier: which windows the DP engines align.  Three implementations of the same decisions:
the reference's own packers cut out of DV-DPfunctions.cu and compiled (oracle/_ref/libref_windows.so), the restatement
(oracle/window_oracle.c) and the product's s3_windows.cuh compiled for the host (tests/native/windows_harness.cpp) --
candidates in the middle of the text, at both of its ends, reads of 36..250 bases, both strands, FR / FF legs."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import helpers

NATIVE = os.path.join(os.path.dirname(__file__), "native")
U32P, I32P = helpers.U32P, C.POINTER(C.c_int32)
U8P = C.POINTER(C.c_uint8)


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(NATIVE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libwindows_harness.so")
    src = os.path.join(NATIVE, "windows_harness.cpp")
    hdr = os.path.join(helpers.ROOT, "soap3-dp_b200", "csrc", "s3_windows.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-x", "c++", src, "-I", os.path.dirname(hdr), "-o", so])
    lib = C.CDLL(so)
    lib.harness_window.restype = C.c_int
    lib.harness_window.argtypes = [C.c_int] * 9 + [C.c_uint, C.c_uint, U32P, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_uint, C.c_uint, U32P]
    return lib


def candidates(rng, n, text_length, lens):
    num_reads = len(lens)
    rid = rng.integers(0, num_reads, n).astype(np.uint32)
    pos = rng.integers(0, text_length, n).astype(np.int64)
    edge = rng.random(n)
    pos = np.where(edge < 0.15, rng.integers(0, 700, n), pos)                         # at the start of the text: windows wrap below 0
    pos = np.where(edge > 0.85, text_length - rng.integers(1, 700, n), pos)           # at its end
    return rid, pos.astype(np.uint32), rng.integers(1, 3, n).astype(np.uint8)


def cutoff_of(par_cut, length):
    return par_cut if par_cut >= 0 else int(math.ceil(0.3 * float(length)))


def run_product(lib, mode, P, lens, rid, pos, pos2=None, strand=None, lsc=None, lst=None, lhit=None):
    out = np.zeros(22, np.uint32)
    rows, cand = [], []
    for c in range(len(rid)):
        k = lib.harness_window(mode, P["ins_low"], P["ins_high"], P["left"], P["right"], P["clip_l"], P["clip_r"], P["cut"][0], P["cut"][1], P["max_dna"],
                               P["text"], helpers.u32p(lens), int(rid[c]), int(pos[c]), int(pos2[c]) if pos2 is not None else 0,
                               int(strand[c]) if strand is not None else 0, int(lsc[c]) if lsc is not None else 0,
                               int(lst[c]) if lst is not None else 0, int(lhit[c]) if lhit is not None else 0, helpers.u32p(out))
        for j in range(k):
            rows.append(out[11 * j:11 * j + 11].copy()); cand.append(c)
    return np.array(cand, np.uint32), (np.stack(rows) if rows else np.zeros((0, 11), np.uint32))


PARAMS = [dict(ins_low=200, ins_high=500, left=1, right=2, clip_l=3, clip_r=8, cut=(-1, -1)),
          dict(ins_low=150, ins_high=650, left=2, right=1, clip_l=49, clip_r=49, cut=(40, 35)),
          dict(ins_low=100, ins_high=400, left=1, right=1, clip_l=0, clip_r=5, cut=(-1, -1))]


@pytest.mark.parametrize("pi", range(len(PARAMS)))
def test_windows_three_ways(harness, pi):
    ref = helpers.load_ref_windows()
    olib = helpers.load_oracle()
    rng = np.random.default_rng(40 + pi)
    text = 5_000_000
    lens = rng.choice([36, 50, 75, 100, 101, 150, 250], 400).astype(np.uint32)
    P = dict(PARAMS[pi], text=text, max_dna=PARAMS[pi]["ins_high"] - PARAMS[pi]["ins_low"] + 256 + 1)
    n = 3000
    rid, pos, strand = candidates(rng, n, text, lens)
    # ---- single-end: one threshold for the whole batch
    Ps = dict(P, cut=(P["cut"][0], P["cut"][0]))
    cand, got = run_product(harness, 1, Ps, lens, rid, pos, strand=strand)
    want = helpers.oracle_windows_single(olib, rid, pos, strand, lens, text, P["clip_l"], P["clip_r"])
    assert np.array_equal(got[:, [1, 2, 4, 5]], want)
    assert np.array_equal(got[:, 8].astype(np.int32), np.array([cutoff_of(Ps["cut"][0], lens[r]) for r in rid], np.int32))
    if ref is not None:
        r = helpers.ref_windows_single(ref, rid, pos, strand, lens, text, P["clip_l"], P["clip_r"], 33)
        assert np.array_equal(r, want)
    # ---- half-end
    cand, got = run_product(harness, 2, P, lens, rid, pos, strand=strand)
    ocand, want = helpers.oracle_windows_half(olib, rid, pos, strand, lens, text, P)
    assert np.array_equal(cand, ocand) and np.array_equal(got[:, [10, 1, 2, 3, 9, 4, 5, 6, 7]], want)
    assert np.array_equal(got[:, 0], rid[cand] ^ 1)
    assert np.array_equal(got[:, 8].astype(np.int32), np.array([cutoff_of(P["cut"][(r ^ 1) & 1], lens[r ^ 1]) for r in rid[cand]], np.int32))
    assert len(cand) > n // 3
    if ref is not None:
        rcand, r, rcut = helpers.ref_windows_half(ref, rid, pos, strand, lens, text, P, cut0=77, cut1=88)
        assert np.array_equal(rcand, ocand) and np.array_equal(r[:, [0, 1, 2, 3, 5, 6, 7, 8]], want[:, [0, 1, 2, 3, 5, 6, 7, 8]])
        assert np.array_equal(rcut, np.where((rid[rcand] ^ 1) & 1, 88, 77))          # cutoffThreshold[unalignedIsReadOrMate]
    # ---- deep DP: left windows, then right windows given the left alignments
    pos2 = (pos.astype(np.int64) + rng.integers(-100, 600, n)).clip(0, text - 1).astype(np.uint32)
    cand, gotL = run_product(harness, 3, P, lens, rid, pos)
    wantL = helpers.oracle_windows_pair_left(olib, rid, pos, lens, text, P)
    assert np.array_equal(gotL[:, [1, 2, 4, 5, 6, 7]], wantL)
    lsc = rng.integers(0, 80, n).astype(np.int32)
    lhit = rng.integers(0, 60, n).astype(np.uint32)
    cand, gotR = run_product(harness, 4, P, lens, rid, pos, pos2=pos2, lsc=lsc, lst=gotL[:, 1], lhit=lhit)
    passed = np.array([lsc[c] >= cutoff_of(P["cut"][rid[c] & 1], lens[rid[c]]) for c in range(n)])
    assert np.array_equal(cand, np.nonzero(passed)[0])
    wantR = helpers.oracle_windows_pair_right(olib, rid[cand], pos2[cand], gotL[cand, 1], lhit[cand], lens, text, P)
    assert np.array_equal(gotR[:, [0, 1, 2, 4, 5, 6, 7]], wantR)
    if ref is not None and P["cut"][0] >= 0:
        rL, rR = helpers.ref_windows_pair(ref, rid, pos, pos2, lens, text, P, lsc, lhit)
        assert np.array_equal(rL, wantL)
        assert np.array_equal(rR[cand][:, [0, 1, 3, 4, 5, 6]], wantR[:, [1, 2, 3, 4, 5, 6]])
