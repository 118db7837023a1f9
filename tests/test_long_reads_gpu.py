"""Long reads (> 120 bases): s3_validate_alignments against the oracle restatement of validateAlignments, and the long-read
mode of s3_se_align (search of the first 100 bases -> collect -> locate -> validation) against the composition of the oracles."""
import os
import sys

import numpy as np
import pytest

import helpers
from helpers import HostIndex, ROOT, fmindex, formats, load_oracle, oracle_launch
from soap3dp_b200 import api, synth

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pe_chain_oracle  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    G = synth.random_genome(400_000, seed=41)
    idx = fmindex.build_index(G, keep_sa=True)
    gi = api.GPUINDEXUpload(idx, device=0, with_text=True, with_sa=True)
    yield G, idx, HostIndex(idx), gi
    api.GPUINDEXFree(gi)


def test_validate_alignments_bit_exact(env):
    G, idx, hi, gi = env
    gen = G.cpu().numpy().astype(np.uint8)
    pac = idx.packed_text.cpu().numpy().view(np.uint32)
    pac = np.concatenate([pac, np.zeros(4, np.uint32)])
    olib = load_oracle()
    rng = np.random.default_rng(7)
    for keep, mins, dbl, mh in ((0, 0, 0, 1000), (1, 0, 0, 1000), (1, 1, 0, 3), (0, 0, 1, 2), (1, 2, 1, 1)):
        cases = list(helpers.validation_cases(rng, gen, 1500))
        n = len(cases)
        maxL = max(len(c[0]) for c in cases)
        reads = np.zeros((n, maxL), np.uint8)
        lens = np.zeros(formats.ceil32(n), np.uint32)
        off = [0]
        pos, fl = [], []
        for k, (read, seed, p, st, mm, *_rest) in enumerate(cases):
            reads[k, :len(read)] = read
            lens[k] = len(read)
            pos += p
            fl += [[s, m] for s, m in zip(st, mm)]
            off.append(len(pos))
        wpq = formats.word_per_query(maxL)
        q = formats.pack_queries(reads, lens[:n], wpq)
        fl = np.array(fl, np.uint8).reshape(-1, 2)
        cnt, gp, gf = api.validate_alignments(gi, q, lens, n, wpq, np.array(off, np.uint32), np.array(pos, np.uint32), fl, bool(keep), mins, bool(dbl), mh)
        wc, wp, wf = helpers.oracle_validate_batch(olib, pac, hi.n, reads, lens[:n], off, pos, fl, keep, mins, dbl, mh)
        assert np.array_equal(cnt, wc)
        gf = gf.reshape(-1, 2)
        for r in range(n):
            a, m = off[r], int(wc[r])
            assert np.array_equal(gp[a:a + m], wp[a:a + m]) and np.array_equal(gf[a:a + m], wf[a:a + m]), (keep, mins, dbl, mh, r)
        assert 0 < int(wc.sum()) < len(pos)


def test_validate_alignments_bad_args(env):
    G, idx, hi, gi = env
    z = np.zeros(32 * 8, np.uint32)
    with pytest.raises(api.S3Error):
        api.validate_alignments(gi, z, z[:32], 2, 8, np.array([0, 2, 1], np.uint32), np.zeros(2, np.uint32), np.zeros((2, 2), np.uint8))
    with pytest.raises(api.S3Error):
        api.validate_alignments(gi, z, z[:32], 1, 8, np.array([0, 0], np.uint32), np.zeros(1, np.uint32), np.zeros((1, 2), np.uint8), max_hit_num=0)
    cnt, _, _ = api.validate_alignments(gi, z, z[:32], 0, 8, np.array([0], np.uint32), np.zeros(0, np.uint32), np.zeros((0, 2), np.uint8))
    assert len(cnt) == 0


@pytest.mark.parametrize("L,k", [(150, 2), (200, 1), (110, 2)])
def test_se_long_read_mode_bit_exact(env, L, k):
    """reads of 150 / 200 bases (seed = first 100) and 110 bases (not long: the mode changes nothing but the cap)"""
    G, idx, hi, gi = env
    gen = G.cpu().numpy().astype(np.uint8)
    pac = np.concatenate([idx.packed_text.cpu().numpy().view(np.uint32), np.zeros(4, np.uint32)])
    n = 2500
    rs = synth.simulate_single_end(G, n, L, seed=50 + L, sub_rate=0.012)
    reads = rs.reads.cpu().numpy().astype(np.uint8)
    wpq = formats.word_per_query(L)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    lens[5:n:9] = L - 7                                      # mixed lengths inside the batch
    for r in range(5, n, 9):
        reads[r, L - 7:] = 0
    q = formats.pack_queries(reads, lens[:n], wpq)
    seed_lens = np.where(lens > 120, 100, lens).astype(np.uint32)
    olib = load_oracle()
    allowed = formats.SA_RANGES_ROUND1[k]
    wpa = 2 * allowed
    bad = np.zeros(formats.ceil32(n), np.uint8)
    views = []
    for case in range(formats.NUM_CASES[k]):
        a = np.zeros(formats.ceil32(n) * wpa, np.uint32)
        oracle_launch(olib, hi, case, q, seed_lens, n, wpq, a, bad, 0, k, allowed, wpa)
        views.append(formats.answers_view(a, n, wpa))
    sa = idx.fwd.sa.cpu().numpy()
    for keep, mins, dbl, cap in ((False, 0, False, 1000), (True, 0, False, 1000), (False, 0, True, 4)):
        col = pe_chain_oracle.collect(views, allowed, hi.n, cap)
        off, pos, fl = [0], [], []
        for ranges, tot, more in col:
            for l, rr, st, mm in ranges:
                for i in range(l, rr + 1):
                    pos.append(int(sa[i]))
                    fl.append([st, mm])
            off.append(len(pos))
        fl = np.array(fl, np.uint8).reshape(-1, 2)
        wc, wp, wf = helpers.oracle_validate_batch(olib, pac, hi.n, reads, lens[:n], off, pos, fl, keep, mins, dbl, cap)
        al = api.SingleAligner(gi, n, num_mismatch=k, max_output_per_read=cap, long_read_mode=True, only_keep_best=keep, min_seed_mismatch=mins,
                               double_allowance=dbl)
        got = al.align(q, lens, n, wpq)
        al.free()
        assert np.array_equal(np.diff(got["occ_offsets"].astype(np.int64)), wc.astype(np.int64))
        aligned = 0
        for r in range(n):
            a, m, b = off[r], int(wc[r]), int(got["occ_offsets"][r])
            assert np.array_equal(got["positions"][b:b + m], wp[a:a + m]) and np.array_equal(got["occ_flags"][b:b + m], wf[a:a + m]), (keep, cap, r)
            aligned += m > 0
        assert aligned > n // 4
        if L > 120:
            assert int(wc.sum()) < len(pos) or keep is False
