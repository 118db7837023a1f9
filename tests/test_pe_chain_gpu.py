"""Parity of the device-resident paired-end chain (s3_pe_align: search -> collect -> route -> locate -> pairing ->
rescue windows -> DP -> CIGAR runs) with the composition of the oracles on the host (oracle/pe_chain_oracle.py):
route code of every pair, the optimal pairing and its counts, every rescue window's record and CIGAR -- bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import (DPBatch, HostIndex, ROOT, fmindex, formats, load_oracle, load_oracle_dp, oracle_dp, oracle_launch,
                     oracle_pair_occurrences)
from soap3dp_b200 import api, synth

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import decode_oracle  # noqa: E402
import pe_chain_oracle  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    G = synth.random_genome(500_000, seed=17)
    idx = fmindex.build_index(G, keep_sa=True)
    gi = api.GPUINDEXUpload(idx, device=0, with_text=True, with_sa=True)
    yield G, idx, HostIndex(idx), gi
    api.GPUINDEXFree(gi)


def _dp_fn(dna, dna_len, rd, rl, max_dna, max_read, cutoff, clip_lt, clip_rt, anc_l, anc_r, scores):
    b = DPBatch(dna, dna_len, rd, rl, max_dna, max_read, cutoff, clip_lt, clip_rt, anc_l, anc_r)
    sc, hit, cnt, pat, _ = oracle_dp(load_oracle_dp(), b, scores)
    return sc, hit, cnt, pat, b.pat_len


def _decode_fn(pat, score, read_length, scores):
    return decode_oracle.decode_one(pat, score, read_length, scores)[0]


LAST = {}     # the inputs of the last _run_both call (reads, the oracle's answer slots), for tests that go on from its results


def _run_both(env, pairs, L, seed, k=2, max_hit=None, keep_second=False, max_output=1000, bad_mate_fraction=0.2, insert=(200, 500), read_stats=False):
    G, idx, hi, gi = env
    m1, m2, _ = synth.simulate_paired_end(G, pairs, L, seed=seed, insert_lo=insert[0], insert_hi=insert[1], bad_mate_fraction=bad_mate_fraction)
    reads = torch.stack([m1.reads, m2.reads], dim=1).reshape(2 * pairs, L).cpu().numpy()
    n = 2 * pairs
    # a few pairs with both mates unfindable, and a few with a mate from a different place (both hit, no valid pair)
    rng = np.random.default_rng(seed)
    for p in rng.choice(pairs, max(pairs // 20, 1), replace=False):
        reads[2 * p] = rng.integers(0, 4, L)
        reads[2 * p + 1] = rng.integers(0, 4, L)
    gen = G.cpu().numpy()
    for p in rng.choice(pairs, max(pairs // 20, 1), replace=False):
        o = int(rng.integers(1000, len(gen) - 1000))
        reads[2 * p + 1] = gen[o:o + L]
    wpq = formats.word_per_query(L)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    q = formats.pack_queries(reads, lens[:n], wpq)
    par = api.pe_params(num_mismatch=k, insert_low=insert[0], insert_high=insert[1], max_output_per_read=max_output,
                        max_hit_num_for_dp=max_hit, keep_second_best=keep_second, read_length=L, read_stats=read_stats)
    al = api.PairAligner(gi, n, L, par)
    try:
        got = al.align(q, lens, n, wpq)
    finally:
        al.free()
    # the oracles
    olib = load_oracle()
    allowed = formats.SA_RANGES_ROUND1[k]
    wpa = 2 * allowed
    bad = np.zeros(formats.ceil32(n), np.uint8)
    views = []
    for case in range(formats.NUM_CASES[k]):
        a = np.zeros(formats.ceil32(n) * wpa, np.uint32)
        oracle_launch(olib, hi, case, q, lens, n, wpq, a, bad, 0, k, allowed, wpa)
        views.append(formats.answers_view(a, n, wpa))
    sa = idx.fwd.sa.cpu().numpy()
    max_read = (L // 4 + 1) * 4
    opar = dict(insert_low=insert[0], insert_high=insert[1], left_leg=1, right_leg=2, max_output_per_read=max_output, max_hit=par.maxHitNumForDP,
                keep_second_best=keep_second, cutoff=-1, soft_clip_left=3, soft_clip_right=8, max_read=max_read,
                max_dna=insert[1] - insert[0] + max_read + 1, scores=(1, -2, -3, -1))
    want = pe_chain_oracle.pe_chain(views, allowed, lens[:n], sa, gen, list(reads), opar, oracle_pair_occurrences, _dp_fn, _decode_fn)
    LAST.update(reads=reads, views=views, allowed=allowed, max_output=max_output, text_length=hi.n)
    return got, want


def _compare(got, want):
    assert np.array_equal(got["route"], want["route"]), np.nonzero(got["route"] != want["route"])[0][:10]
    for p, w in enumerate(want["pairs"]):
        g = got["pairs"][p]
        if w is None:
            assert g["numPairs"] == 0
            continue
        assert g["numPairs"] == w["numPairs"], p
        if w["numPairs"]:
            for k in ("pos1", "pos2", "insertion", "strand1", "mism1", "strand2", "mism2", "optimalTotal", "numOptimal", "suboptimalTotal", "numSuboptimal"):
                assert int(g[k]) == w[k], (p, k, int(g[k]), w[k])
    assert len(got["dp"]) == len(want["dp"])
    traced = 0
    for t, w in enumerate(want["dp"]):
        g = got["dp"][t]
        for k in ("dpReadID", "alignedPos", "alignedStrand", "alignedMismatches", "dpStrand", "leftOrRight", "score", "numSameScore", "dpPos"):
            assert int(g[k]) == w[k], (t, k, int(g[k]), w[k])
        cig = api.runs_to_cigar(got["runs"][int(g["runOffset"]):int(g["runOffset"]) + int(g["numRuns"])])
        assert cig == w["cigar"], (t, cig, w["cigar"])
        traced += w["cigar"] != ""
    return traced


def oracle_read_stats(views, allowed, text_length, max_output):
    """what hostKernel keeps of a read for MAPQ (CPUfunctions.cpp:2061-2141) from rOutput->WithError of collect_all_answers: (X0, X1, fewest mismatches)"""
    out = []
    for ranges, tot, more in pe_chain_oracle.collect(views, allowed, text_length, max_output):
        with_error = [0] * 9
        for l, r, st, mm in ranges:
            with_error[mm] += r - l + 1
        m = next((i for i in range(8) if with_error[i]), None)
        out.append((0, 0, 255) if m is None else (with_error[m], with_error[m + 1], m))
    return out


def test_pe_chain_read_stats(env):
    """params.readStats: X0 / X1 / the fewest mismatches of every read == the restatement over the oracle's answer slots"""
    got, want = _run_both(env, 1200, 100, 8, read_stats=True)
    _compare(got, want)
    stats = oracle_read_stats(LAST["views"], LAST["allowed"], LAST["text_length"], LAST["max_output"])
    assert len(got["read_stats"]) == len(stats) == 2400
    mine = [(int(x["x0"]), int(x["x1"]), int(x["minMismatch"])) for x in got["read_stats"]]
    assert mine == stats
    assert sum(1 for s_ in stats if s_[2] == 255) > 50 and sum(1 for s_ in stats if s_[0] > 1) > 10
    plain, _ = _run_both(env, 300, 100, 8)
    assert len(plain["read_stats"]) == 0


@pytest.mark.parametrize("L,pairs,seed", [(100, 1500, 3), (75, 800, 4), (50, 600, 5)])
def test_pe_chain_bit_exact(env, L, pairs, seed):
    got, want = _run_both(env, pairs, L, seed)
    traced = _compare(got, want)
    routes = np.bincount(want["route"], minlength=9)
    assert routes[pe_chain_oracle.PAIRED] > pairs // 3 and routes[pe_chain_oracle.NONE] > 0
    assert routes[pe_chain_oracle.FIRST_RESCUES] + routes[pe_chain_oracle.SECOND_RESCUES] > pairs // 20 and traced > pairs // 20
    assert routes[pe_chain_oracle.BOTH_RESCUE] > 0
    assert got["d2h_bytes"] < 200 * pairs              # a few dozen bytes per pair come back


def test_pe_chain_few_hits_allowed(env):
    """maxHitNumForDP 1 and a cap of 3 occurrences per read: the best-hit filter, the too-many routes and the truncation of
    a read's list all come into play (the genome has repeats)"""
    for keep_second in (False, True):
        got, want = _run_both(env, 1200, 100, 7, max_hit=1, keep_second=keep_second, max_output=3)
        _compare(got, want)


def test_pe_chain_empty_and_bad_args(env):
    G, idx, hi, gi = env
    al = api.PairAligner(gi, 64, 100, api.pe_params())
    z = np.zeros(64 * 8, np.uint32)
    out = al.align(z, z[:64], 0, 8)
    assert len(out["route"]) == 0
    with pytest.raises(api.S3Error):
        al.align(z, z[:64], 63, 8)                      # odd number of reads
    with pytest.raises(api.S3Error):
        al.align(z, z[:64], 128, 8)                     # more than maxReads
    al.free()


def test_pe_prefetch_gives_the_same_batches(env):
    """s3_pe_prefetch uploads the next batch while the current one is aligned: results of three batches are those of plain calls"""
    G, idx, hi, gi = env
    L, pairs = 100, 700
    wpq = formats.word_per_query(L)
    n = 2 * pairs
    sets = []
    for seed in (21, 22, 23):
        m1, m2, _ = synth.simulate_paired_end(G, pairs, L, seed=seed, bad_mate_fraction=0.2)
        reads = torch.stack([m1.reads, m2.reads], dim=1).reshape(n, L).cpu().numpy()
        lens = np.zeros(formats.ceil32(n), np.uint32)
        lens[:n] = L
        sets.append((formats.pack_queries(reads, lens[:n], wpq), lens))
    al = api.PairAligner(gi, n, L, api.pe_params())
    plain = [al.align(q, l, n, wpq) for q, l in sets]
    got = []
    for k, (q, l) in enumerate(sets):
        if k + 1 < len(sets):
            al.prefetch(sets[k + 1][0], sets[k + 1][1], n, wpq)
        got.append(al.align(q, l, n, wpq))
    al.free()
    for a, b in zip(plain, got):
        for key in ("route", "pairs", "dp", "runs"):
            assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4])
def test_se_chain_bit_exact(env, k):
    """s3_se_align (alignSingleR's results): every read's occurrences == collect_all_answers + transferAllSAToOcc restated on the oracle's slots"""
    G, idx, hi, gi = env
    L, n = 100, 3000
    rs = synth.simulate_single_end(G, n, L, seed=30 + k, sub_rate=0.02)
    reads = rs.reads.cpu().numpy()
    wpq = formats.word_per_query(L)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    q = formats.pack_queries(reads, lens[:n], wpq)
    olib = load_oracle()
    allowed = formats.SA_RANGES_ROUND1[k]
    wpa = 2 * allowed
    bad = np.zeros(formats.ceil32(n), np.uint8)
    views = []
    for case in range(formats.NUM_CASES[k]):
        a = np.zeros(formats.ceil32(n) * wpa, np.uint32)
        oracle_launch(olib, hi, case, q, lens, n, wpq, a, bad, 0, k, allowed, wpa)
        views.append(formats.answers_view(a, n, wpa))
    sa = idx.fwd.sa.cpu().numpy()
    for best, cap in ((False, 1000), (True, 1000), (False, 5)):
        col = pe_chain_oracle.collect(views, allowed, hi.n, cap)
        al = api.SingleAligner(gi, n, num_mismatch=k, max_output_per_read=cap, report_best=best)
        got = al.align(q, lens, n, wpq)
        al.free()
        off, hits = 0, 0
        for r, (ranges, tot, more) in enumerate(col):
            if best and ranges:
                ranges, tot = pe_chain_oracle.retain(ranges, False)
            want = [(int(sa[i]), st, mm) for l, rr, st, mm in ranges for i in range(l, rr + 1)]
            a, b = int(got["occ_offsets"][r]), int(got["occ_offsets"][r + 1])
            assert a == off and b - a == len(want), (k, best, cap, r)
            have = [(int(p), int(f[0]), int(f[1])) for p, f in zip(got["positions"][a:b], got["occ_flags"][a:b])]
            assert have == want, (k, best, cap, r)
            assert int(got["read_flags"][r]) == int(more)
            off = b
            hits += bool(want)
        assert hits > n // 20
