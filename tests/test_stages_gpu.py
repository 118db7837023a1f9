"""The DP stages that start from seeds (s3_single_dp_align = DPForUnalignSingle2, s3_deep_dp_align = DPForUnalignPairs2) against
the composition of the oracles on the host (oracle/seeding_oracle.py): seeds, seeding driver, candidate positions, windows,
alignments and CIGARs -- bit-exact."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

import helpers
from helpers import DPBatch, HostIndex, ROOT, fmindex, formats, load_oracle, load_oracle_dp, oracle_dp, oracle_launch, u32p
from soap3dp_b200 import api, synth

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import decode_oracle  # noqa: E402
import params_oracle  # noqa: E402
import seeding_oracle  # noqa: E402

pytestmark = pytest.mark.gpu
U, I = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def env():
    G = synth.random_genome(400_000, seed=23)
    idx = fmindex.build_index(G, keep_sa=True)
    gi = api.GPUINDEXUpload(idx, device=0, with_text=True, with_sa=True)
    yield G, idx, HostIndex(idx), gi
    api.GPUINDEXFree(gi)


class OracleEnv:
    def __init__(self, idx, hi, sa=None):
        self.olib, self.hi = load_oracle(), hi
        self.sa = sa if sa is not None else np.ascontiguousarray(idx.fwd.sa.cpu().numpy().astype(np.uint32))
        o = self.olib
        o.s3o_seed_candidates.restype = C.c_uint64
        o.s3o_seed_candidates.argtypes = [U, U, U, I, U, U, U, U, C.c_uint64, C.c_uint32, U, U, I, C.c_uint64]
        o.s3o_seed_pair_candidates.restype = C.c_uint64
        o.s3o_seed_pair_candidates.argtypes = [U] + [U, U, I, U, U, U, U, C.c_uint64] * 2 + [C.c_uint32, U, C.c_int, C.c_int, C.c_int, C.c_int, U, U, U, C.c_uint64]

    def search(self, seeds):
        n = len(seeds)
        L = max((len(s) for s in seeds), default=1)
        wps = formats.word_per_query(L)
        lens = np.zeros(formats.ceil32(max(n, 1)), np.uint32)
        lens[:n] = [len(s) for s in seeds]
        rows = np.zeros((max(n, 1), L), np.uint8)
        for k, s in enumerate(seeds):
            rows[k, :len(s)] = s
        q = formats.pack_queries(rows[:n] if n else rows[:0], lens[:n], wps) if n else np.zeros(32 * wps, np.uint32)

        def launcher(qq, ll, m):
            def run(case, k, allowed, wpa):
                a = np.zeros(formats.ceil32(max(m, 1)) * wpa, np.uint32)
                oracle_launch(self.olib, self.hi, case, qq, ll, m, wps, a, np.zeros(formats.ceil32(max(m, 1)), np.uint8), 0, k, allowed, wpa)
                return formats.answers_view(a, m, wpa)
            return run

        def launch_one(ids):
            ll = np.zeros(formats.ceil32(max(len(ids), 1)), np.uint32)
            ll[:len(ids)] = lens[ids]
            return launcher(formats.pack_queries(rows[ids], lens[ids], wps), ll, len(ids))
        return launcher(q, lens, n), launch_one

    @staticmethod
    def _cols(rng):
        a = np.array(rng, np.int64).reshape(-1, 7)
        f = lambda c, t=np.uint32: np.ascontiguousarray(a[:, c].astype(t))
        return f(0), f(1), f(2, np.int32), f(3), f(4), f(5), f(6)

    def seed_candidates(self, rng):
        if not rng:
            return []
        l, r, st, rid, off, sl, rl = self._cols(rng)
        cap = int((r.astype(np.int64) - l + 1).sum()) + 1
        o = [np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.int32)]
        m = self.olib.s3o_seed_candidates(u32p(self.sa), u32p(l), u32p(r), st.ctypes.data_as(I), u32p(rid), u32p(off), u32p(sl), u32p(rl), len(l), 0xFFFFFFFF,
                                          u32p(o[0]), u32p(o[1]), o[2].ctypes.data_as(I), cap)
        return [(int(o[0][k]), int(o[1][k]), int(o[2][k])) for k in range(m)]

    def seed_pair_candidates(self, r0, r1, lens, ins_low, ins_high, left, right):
        z = [np.zeros(1, np.uint32)] * 2 + [np.zeros(1, np.int32)] + [np.zeros(1, np.uint32)] * 4
        a0, a1 = (self._cols(r0) if r0 else z), (self._cols(r1) if r1 else z)
        cap = 4 * (sum(r - l + 1 for l, r, *_ in r0) + sum(r - l + 1 for l, r, *_ in r1)) + 64
        o = [np.zeros(cap, np.uint32) for _ in range(3)]

        def args(a, n):
            return [u32p(a[0]), u32p(a[1]), a[2].ctypes.data_as(I), u32p(a[3]), u32p(a[4]), u32p(a[5]), u32p(a[6]), n]
        lens = np.ascontiguousarray(lens, np.uint32)
        m = self.olib.s3o_seed_pair_candidates(u32p(self.sa), *args(a0, len(r0)), *args(a1, len(r1)), 0xFFFFFFFF, u32p(lens), ins_low, ins_high, left, right,
                                               u32p(o[0]), u32p(o[1]), u32p(o[2]), cap)
        assert m <= cap
        return [(int(o[0][k]), int(o[1][k]), int(o[2][k])) for k in range(m)]

    def window_single(self, r, pos, strand, lens, text, clip_l, clip_r):
        return [int(x) for x in helpers.oracle_windows_single(self.olib, [r], [pos], [strand], lens, text, clip_l, clip_r)[0]]

    def window_pair_left(self, left, pos, lens, text, P):
        return [int(x) for x in helpers.oracle_windows_pair_left(self.olib, [left], [pos], lens, text, dict(P, clip_l=P["clip_l"], clip_r=P["clip_r"]))[0]]

    def window_pair_right(self, left, pos2, lstart, lhit, lens, text, P):
        return [int(x) for x in helpers.oracle_windows_pair_right(self.olib, [left], [pos2], [lstart], [lhit], lens, text, P)[0]]

    def dp(self, dna, dna_len, rd, rl, max_dna, max_read, cutoff, clip_lt, clip_rt, anc_l, anc_r, scores):
        b = DPBatch(dna, dna_len, rd, rl, max_dna, max_read, cutoff, clip_lt, clip_rt, anc_l, anc_r)
        sc, hit, cnt, pat, _ = oracle_dp(load_oracle_dp(), b, scores)
        return sc, hit, cnt, pat, b.pat_len

    def decode(self, pat, score, read_length, scores):
        return decode_oracle.decode_one(pat, score, read_length, scores)[0]

    def seed_positions(self, stage, length):
        return params_oracle.seed_positions(stage, length)

    def max_hit(self, stage, l1, l2):
        p = params_oracle.stage_parameters(stage, l1, l2)
        return p["reads"][0]["maxHitNum"], p["reads"][1]["maxHitNum"]


def mutate(rng, read, n_sub, n_indel):
    r = list(read)
    for _ in range(n_sub):
        k = int(rng.integers(0, len(r)))
        r[k] = (r[k] + int(rng.integers(1, 4))) & 3
    for _ in range(n_indel):
        k = int(rng.integers(10, len(r) - 10))
        if rng.random() < 0.5:
            del r[k]
            r.append(int(rng.integers(0, 4)))
        else:
            r.insert(k, int(rng.integers(0, 4)))
            r.pop()
    return np.array(r, np.uint8)


PAR = dict(ins_low=200, ins_high=500, left=1, right=2, clip_l=3, clip_r=8, cut=(-1, -1), default_threshold=True, threshold=0, scores=(1, -2, -3, -1))


@pytest.mark.parametrize("L", [100, 150])
def test_single_dp_stage(env, L):
    G, idx, hi, gi = env
    rng = np.random.default_rng(L)
    n = 300
    rs = synth.simulate_single_end(G, n, L, seed=L + 1, sub_rate=0.01)
    reads = [mutate(rng, r, int(rng.integers(3, 9)), int(rng.integers(0, 3))) for r in rs.reads.cpu().numpy()]
    for k in range(0, n, 25):
        reads[k] = rng.integers(0, 4, L).astype(np.uint8)                       # no seed anywhere
    wpq = formats.word_per_query(L)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    q = formats.pack_queries(np.stack(reads), lens[:n], wpq)
    ids = np.arange(0, n, dtype=np.uint32)[rng.random(n) < 0.8]
    got = api.single_dp_align(gi, q, lens, n, wpq, ids, api.stage_params())
    want = seeding_oracle.single_dp(OracleEnv(idx, hi), G.cpu().numpy(), reads, ids.tolist(), PAR)
    assert got["num_seeds"] == want["seeds"] and got["num_candidates"] == want["candidates"]
    assert got["unseeded"].tolist() == want["unseeded"]
    assert len(got["hits"]) == len(want["hits"]) > n // 3
    for h, w in zip(got["hits"], want["hits"]):
        cig = api.runs_to_cigar(got["runs"][int(h["runOffset"]):int(h["runOffset"]) + int(h["numRuns"])])
        assert (int(h["readID"]), int(h["strand"]), int(h["pos"]), int(h["score"]), int(h["numSameScore"]), cig) == w


def test_deep_dp_stage(env):
    G, idx, hi, gi = env
    rng = np.random.default_rng(8)
    L, pairs = 100, 160
    m1, m2, _ = synth.simulate_paired_end(G, pairs, L, seed=12, bad_mate_fraction=0.0)
    raw = torch.stack([m1.reads, m2.reads], dim=1).reshape(2 * pairs, L).cpu().numpy()
    reads = [mutate(rng, r, int(rng.integers(3, 8)), int(rng.integers(0, 2))) for r in raw]
    for p in range(0, pairs, 20):
        reads[2 * p] = rng.integers(0, 4, L).astype(np.uint8)
        reads[2 * p + 1] = rng.integers(0, 4, L).astype(np.uint8)
    n = 2 * pairs
    wpq = formats.word_per_query(L)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    q = formats.pack_queries(np.stack(reads), lens[:n], wpq)
    ids = (2 * np.arange(pairs, dtype=np.uint32))[rng.random(pairs) < 0.9]
    got = api.deep_dp_align(gi, q, lens, n, wpq, ids, api.stage_params())
    want = seeding_oracle.deep_dp(OracleEnv(idx, hi), G.cpu().numpy(), reads, ids.tolist(), PAR)
    assert got["num_seeds"] == want["seeds"] and got["num_candidates"] == want["candidates"]
    assert got["unseeded"].tolist() == want["unseeded"]
    assert len(got["hits"]) == len(want["hits"]) > pairs // 3
    for h, w in zip(got["hits"], want["hits"]):
        c1 = api.runs_to_cigar(got["runs"][int(h["runOffset1"]):int(h["runOffset1"]) + int(h["numRuns1"])])
        c2 = api.runs_to_cigar(got["runs"][int(h["runOffset2"]):int(h["runOffset2"]) + int(h["numRuns2"])])
        assert (int(h["readID"]), int(h["strand1"]), int(h["strand2"]), int(h["pos1"]), int(h["pos2"]), int(h["score1"]), int(h["score2"]),
                int(h["numSame1"]), int(h["numSame2"]), c1, c2) == w


def deep_dp_of_the_chain(env, pairs=400, seed=21):
    """s3_pe_deep_dp (the both-unaligned pairs picked on the device, the chain's own query buffer) == s3_deep_dp_align of the pairs the
    chain routed as S3_PE_NONE == the seeding oracle, after s3_pe_align and after s3_pe_align_device; -> (pairs picked, paired alignments)"""
    G, idx, hi, gi = env
    rng = np.random.default_rng(seed)
    L = 100
    m1, m2, _ = synth.simulate_paired_end(G, pairs, L, seed=31, bad_mate_fraction=0.0)
    raw = torch.stack([m1.reads, m2.reads], dim=1).reshape(2 * pairs, L).cpu().numpy()
    reads = [r.copy() for r in raw]
    for p in range(0, pairs, 3):                                    # a third of the pairs: both mates beyond the search, within the seeds' reach
        for i in range(2):
            reads[2 * p + i] = mutate(rng, raw[2 * p + i], int(rng.integers(4, 8)), int(rng.integers(0, 2)))
    for p in range(0, pairs, 30):
        reads[2 * p] = rng.integers(0, 4, L).astype(np.uint8)
        reads[2 * p + 1] = rng.integers(0, 4, L).astype(np.uint8)
    n = 2 * pairs
    wpq = formats.word_per_query(L)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    q = formats.pack_queries(np.stack(reads), lens[:n], wpq)
    sp = api.stage_params()
    pe = api.PairAligner(gi, n, L, api.pe_params(num_mismatch=2, insert_low=200, insert_high=500, scores=(1, -2, -3, -1), read_length=L, max_windows=n))
    try:
        got_chain = pe.align(q, lens, n, wpq)
        ids = (2 * np.nonzero(got_chain["route"] == 0)[0]).astype(np.uint32)
        assert len(ids) > pairs // 5
        a = pe.deep_dp(sp)
        qd, ld = torch.from_numpy(q.view(np.int32)).cuda(), torch.from_numpy(lens.view(np.int32)).cuda()
        pe.align_device(qd.data_ptr(), ld.data_ptr(), n, wpq)
        torch.cuda.synchronize()
        b = pe.deep_dp(sp)
    finally:
        pe.free()
    want = api.deep_dp_align(gi, q, lens, n, wpq, ids, sp)
    ora = seeding_oracle.deep_dp(OracleEnv(idx, hi), G.cpu().numpy(), reads, ids.tolist(), PAR)
    assert want["num_seeds"] == ora["seeds"] and want["num_candidates"] == ora["candidates"] and len(want["hits"]) == len(ora["hits"]) > len(ids) // 3
    for got in (a, b):
        assert got["num_input"] == len(ids)
        assert got["num_seeds"] == want["num_seeds"] and got["num_candidates"] == want["num_candidates"]
        assert got["unseeded"].tolist() == want["unseeded"].tolist()
        assert got["hits"].tobytes() == want["hits"].tobytes() and np.array_equal(got["runs"], want["runs"])
    return len(ids), len(want["hits"])


def test_deep_dp_of_the_chain(env):
    deep_dp_of_the_chain(env)


def test_stages_empty_and_bad_args(env):
    G, idx, hi, gi = env
    z = np.zeros(32 * 8, np.uint32)
    out = api.single_dp_align(gi, z, z[:32], 32, 8, np.zeros(0, np.uint32), api.stage_params())
    assert len(out["hits"]) == 0
    with pytest.raises(api.S3Error):
        api.deep_dp_align(gi, z, z[:32], 32, 8, np.array([3], np.uint32), api.stage_params())      # odd id
    with pytest.raises(api.S3Error):
        api.single_dp_align(gi, z, z[:32], 32, 8, np.array([40], np.uint32), api.stage_params())
    pe = api.PairAligner(gi, 64, 100, api.pe_params(read_length=100))
    try:
        with pytest.raises(api.S3Error):
            pe.deep_dp(api.stage_params())                             # no batch has been aligned on the handle
    finally:
        pe.free()
