"""CPU tier: the oracle against the committed golden fixtures, which were produced by
the reference's own code (tests/golden/make_golden.py) -- this is what pins the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import (ROOT, HostIndex, fmindex, formats, load_oracle, load_oracle_dp, make_dp_batch, oracle_dp,
                     oracle_launch, pattern_end)
from soap3dp_b200 import synth

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def genome_and_index():
    dig = json.load(open(os.path.join(GOLD, "index_digests.json")))
    G = synth.random_genome(dig["genome"]["n"], seed=dig["genome"]["seed"])
    assert sha(G.numpy()) == dig["genome"]["sha256"], "synthetic genome generator drifted from the fixture"
    return G, fmindex.build_index(G), dig


def test_index_builder_matches_reference_builders(genome_and_index):
    """our torch builder == soap3-dp-builder + BGS-Build output, bit for bit"""
    G, idx, dig = genome_and_index
    n = idx.text_length
    nw = (n + 15) // 16
    for half, tag in ((idx.fwd, ""), (idx.rev, "rev.")):
        hdr = [half.inverse_sa0] + list(half.cum_freq[1:])
        assert hdr == dig[tag + "bwt"]["header"]
        assert hdr == dig[tag + "fmv.gpu"]["header"]
        assert sha(half.bwt_words.numpy().view(np.uint32)[:nw]) == dig[tag + "bwt"]["payload_sha256"]
        assert sha(half.occ.numpy().view(np.uint32)) == dig[tag + "fmv.gpu"]["payload_sha256"]
    # .sa payload: [saInterval=1] then the n+1 SA values, sa[0] = n (2bwt-lib/BWT.c:225-285)
    sa = idx.fwd.sa.numpy().astype(np.uint32)
    assert sha(np.concatenate([np.array([1], np.uint32), sa])) == dig["sa"]["payload_sha256"]


def test_bucketed_suffix_array_equals_single_bucket(genome_and_index):
    G, idx, _ = genome_and_index
    sa2 = fmindex.build_suffix_array(G, max_bucket=1 << 14)
    assert bool((sa2 == idx.fwd.sa).all())


def test_search_oracle_matches_reference_kernels_golden(genome_and_index):
    G, idx, _ = genome_and_index
    gold = json.load(open(os.path.join(GOLD, "search_golden.json")))
    hi = HostIndex(idx)
    olib = load_oracle()
    for entry in gold["sets"]:
        L, n, seed = entry["L"], entry["n"], entry["seed"]
        rs = synth.simulate_single_end(G, n, L, seed=seed, sub_rate=0.015)
        lens = rs.lengths.numpy().astype(np.uint32)
        lens[::5] = L - 1
        lens[3::11] = L - 7
        wpq = formats.word_per_query(L)
        q = formats.pack_queries(rs.reads.numpy(), lens, wpq)
        lens_up = np.zeros(formats.ceil32(n), np.uint32)
        lens_up[:n] = lens
        it = iter(entry["launches"])
        for k in range(5):
            for rnd, allowed in ((0, formats.SA_RANGES_ROUND1[k]), (1, formats.SA_RANGES_ROUND2[k])):
                wpa = 2 * allowed
                bad = np.zeros(formats.ceil32(n), np.uint8)
                for case in range(formats.NUM_CASES[k]):
                    g = next(it)
                    assert (g["k"], g["round"], g["case"]) == (k, rnd, case)
                    a = np.zeros(formats.ceil32(n) * wpa, np.uint32)
                    oracle_launch(olib, hi, case, q, lens_up, n, wpq, a, bad, rnd, k, allowed, wpa)
                    v = formats.answers_view(a, n, wpa)
                    for i, row in g["rows"].items():
                        assert [int(x) for x in v[int(i)]] == row, (L, k, rnd, case, i)
                    assert sha(v) == g["sha256"], (L, k, rnd, case)


def test_dp_oracle_matches_reference_kernels_golden(genome_and_index):
    G, _, _ = genome_and_index
    gold = json.load(open(os.path.join(GOLD, "dp_golden.json")))
    olib = load_oracle_dp()
    for g in gold["batches"]:
        b = make_dp_batch(G, g["n"], g["L"], g["mode"], seed=g["L"] + len(g["mode"]), indel_rate=0.006)
        sc, hit, cnt, pat, _ = oracle_dp(olib, b, tuple(g["scores"]))
        assert sha(sc[:b.n]) == g["scores_sha256"], g
        h = hashlib.sha256()
        for t in range(b.n):
            if sc[t] >= b.cutoff[t]:
                w = pat[t * b.pat_len:(t + 1) * b.pat_len]
                h.update(w[:pattern_end(w) + 1].tobytes())
            else:
                h.update(b"-")
        assert sha(hit[:b.n]) == g["hitLocs"], g
        assert sha(cnt[:b.n]) == g["maxScoreCounts"], g
        assert h.hexdigest() == g["patterns"], g


def test_seed_oracles_match_the_golden_fixtures():
    """oracle/seed_oracle.c against tests/golden/seed_golden.json, which tests/golden/make_seed_golden.py produced with the
    reference's own sort macros, singleMerge, findRevStart and pairEndMerge (works without oracle/_ref)."""
    import ctypes as C
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_seed_golden", os.path.join(GOLD, "make_seed_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    gold = json.load(open(os.path.join(GOLD, "seed_golden.json")))
    o = load_oracle()
    U, I = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    o.s3o_seed_candidates.restype = C.c_uint64
    o.s3o_seed_candidates.argtypes = [U, U, U, I, U, U, U, U, C.c_uint64, C.c_uint32, U, U, I, C.c_uint64]
    o.s3o_seed_pair_candidates.restype = C.c_uint64
    o.s3o_seed_pair_candidates.argtypes = [U] + [U, U, I, U, U, U, U, C.c_uint64] * 2 + [C.c_uint32, U, C.c_int, C.c_int, C.c_int, C.c_int,
                                                                                         U, U, U, C.c_uint64]

    def u(a):
        return a.ctypes.data_as(U)
    for g in gold["single"]:
        seed, n, nreads, span = g["case"]
        rid, x, st, off, sl, rl = mk.single_case(seed, n, nreads, span)
        sa = np.arange(span + 64, dtype=np.uint32)                 # identity suffix array: a range [x, x] is the hit x
        out = [np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.int32)]
        m = o.s3o_seed_candidates(u(sa), u(x), u(x), st.ctypes.data_as(I), u(rid), u(off), u(sl), u(rl), n, 0xFFFFFFFF,
                                  u(out[0]), u(out[1]), out[2].ctypes.data_as(I), n)
        assert m == g["candidates"]
        assert [sha(out[0][:m]), sha(out[1][:m]), sha(out[2][:m])] == [g["readIDs"], g["positions"], g["strands"]]
        assert [[int(a), int(b), int(c)] for a, b, c in zip(out[0][:8], out[1][:8], out[2][:8])] == g["first"]
    for g in gold["pair"]:
        seed, n0, n1, npairs, span = g["case"]
        sides, lens = mk.pair_case(seed, n0, n1, npairs, span)
        sa = np.arange(span + 200, dtype=np.uint32)
        a = []
        for rid, x, st, off, sl, rl in sides:
            a += [u(x), u(x), st.ctypes.data_as(I), u(rid), u(off), u(sl), u(rl), len(x)]
        cap = (n0 + n1) * 50 + 10
        out = [np.zeros(cap, np.uint32) for _ in range(3)]
        m = o.s3o_seed_pair_candidates(u(sa), *a, 0xFFFFFFFF, u(lens), 200, 500, g["legs"][0], g["legs"][1], u(out[0]), u(out[1]), u(out[2]), cap)
        assert m == g["candidates"], g["legs"]
        assert [sha(out[0][:m]), sha(out[1][:m]), sha(out[2][:m])] == [g["readIDLeft"], g["posLeft"], g["posRight"]], g["legs"]


def test_pair_oracle_matches_the_golden_fixture():
    """oracle/pair_oracle.c against records generated from the reference's PEMappingOccurrences + PEStatsPEOutput"""
    import json
    import os
    import helpers
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pair_golden.json")))
    total = 0
    for case in g["cases"]:
        types = (np.uint32, np.uint8, np.uint8, np.uint64) * 2
        lists = tuple(np.ascontiguousarray(np.array(v, dtype=t)) for v, t in zip(case["lists"], types))
        got = helpers.oracle_pair_occurrences(lists, np.array(case["pattern_lengths"], np.uint32), *case["bounds"], *case["legs"],
                                              case["report_one"])
        for k, v in case["want"].items():
            assert np.array_equal(np.asarray(got[k]).astype(np.int64), np.array(v, dtype=np.int64).reshape(np.asarray(got[k]).shape)), k
        stats = np.zeros_like(got["stats"])
        for p, k, c in case["stats_nonzero"]:
            stats[p, k] = c
        assert np.array_equal(stats, got["stats"])
        total += len(case["want"]["pos1"])
    assert total > 200
