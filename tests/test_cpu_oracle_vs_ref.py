"""CPU tier: the oracle restatements against the reference's own kernels compiled for the
host (oracle/_ref, built by oracle/build_ref.sh where /root/reference exists).  Skipped
on boxes where oracle/_ref was not shipped; the golden-fixture tests cover those."""
import os
import numpy as np
import pytest

from helpers import (DPBatch, HostIndex, compare_dp, fmindex, formats, load_oracle, load_oracle_dp, load_ref_dp,
                     load_ref_search, make_dp_batch, oracle_dp, oracle_launch, ref_dp, ref_launch, u32p)
import helpers
from soap3dp_b200 import synth

rlib = load_ref_search()
dlib = load_ref_dp()


@pytest.fixture(scope="module")
def env():
    G = synth.random_genome(400_000, seed=31)
    return G, HostIndex(fmindex.build_index(G))


@pytest.mark.skipif(rlib is None, reason="oracle/_ref/libref_search.so not built")
def test_rank_matches_reference_rank(env):
    G, hi = env
    olib = load_oracle()
    rng = np.random.default_rng(0)
    idxs = np.concatenate([rng.integers(0, hi.n + 2, 50000), [0, 1, hi.n, hi.n + 1, hi.isa0, hi.isa0 + 1]])
    for i in idxs:
        for c in range(4):
            assert olib.s3o_rank(u32p(hi.bwt), u32p(hi.occ), int(i), c, hi.isa0) == \
                rlib.ref_rank(u32p(hi.bwt), u32p(hi.occ), int(i), c, hi.isa0)


@pytest.mark.skipif(rlib is None, reason="oracle/_ref/libref_search.so not built")
@pytest.mark.parametrize("L", [100, 51, 200])
def test_search_all_cases_match_reference_kernels(env, L):
    """every mismatch level, case and round; lengths where (int)(L*ratio) truncates differently"""
    G, hi = env
    olib = load_oracle()
    n = 800
    rs = synth.simulate_single_end(G, n, L, seed=L, sub_rate=0.02)
    lens = rs.lengths.numpy().astype(np.uint32)
    lens[::3] = L - 2
    lens[1::7] = L - 9
    wpq = formats.word_per_query(L)
    q = formats.pack_queries(rs.reads.numpy(), lens, wpq)
    lens_up = np.zeros(formats.ceil32(n), np.uint32)
    lens_up[:n] = lens
    for k in range(5):
        for rnd, allowed in ((0, formats.SA_RANGES_ROUND1[k]), (1, formats.SA_RANGES_ROUND2[k])):
            wpa = 2 * allowed
            bo = np.zeros(formats.ceil32(n), np.uint8)
            br = bo.copy()
            qr = q.copy()
            for case in range(formats.NUM_CASES[k]):
                if rnd == 1:
                    qr = q.copy()
                ao = np.zeros(formats.ceil32(n) * wpa, np.uint32)
                ar = ao.copy()
                no = oracle_launch(olib, hi, case, q, lens_up, n, wpq, ao, bo, rnd, k, allowed, wpa)
                nr = ref_launch(rlib, hi, case, qr, lens_up, n, wpq, ar, br, rnd, k, allowed, wpa, nthreads=2)
                assert np.array_equal(formats.answers_view(ao, n, wpa), formats.answers_view(ar, n, wpa)), (L, k, rnd, case)
                assert np.array_equal(bo, br)
                view = formats.answers_view(ao, n, wpa)
                if not (view[:, 0] == formats.ANSWER_OVERFLOW).any():
                    assert no == nr, "rank-query counts must agree when no slot overflowed"


@pytest.mark.skipif(dlib is None, reason="oracle/_ref/libref_dp.so not built")
@pytest.mark.parametrize("mode", ["single", "rescue"])
def test_dp_matches_reference_kernels(env, mode):
    G, _ = env
    olib = load_oracle_dp()
    for L in (100, 150):
        for scores in ((1, -2, -3, -1), (2, -3, -5, -2)):
            b = make_dp_batch(G, 400, L, mode, seed=3 * L, indel_rate=0.006)
            assert compare_dp(b, oracle_dp(olib, b, scores), ref_dp(dlib, b, scores), f"{mode} {L}") > 350


@pytest.mark.skipif(dlib is None, reason="oracle/_ref/libref_dp.so not built")
def test_dp_adversarial_matches_reference_kernels():
    """random read/window pairs, big clips, random anchors (>= 1, as every reference caller
    produces: DV-DPfunctions.cu:2073,2099,3401,3455), cutoffs >= 0"""
    olib = load_oracle_dp()
    for seed in range(4):
        rng = np.random.default_rng(seed)
        n, L = 600, int(rng.integers(20, 130))
        W = L + int(rng.integers(5, 200))
        dna = rng.integers(0, 4, (n, W)).astype(np.uint8)
        read = rng.integers(0, 4, (n, L)).astype(np.uint8)
        for t in range(0, n, 2):
            o = rng.integers(0, W - L)
            dna[t, o:o + L] = read[t]
            for _ in range(int(rng.integers(0, 6))):
                dna[t, o + rng.integers(0, L)] = rng.integers(0, 4)
        b = DPBatch(dna, rng.integers(W - 3, W + 1, n).astype(np.uint32), read,
                    rng.integers(max(L - 8, 1), L + 1, n).astype(np.uint32), W + 14, (L // 4 + 1) * 4,
                    rng.integers(0, 30, n).astype(np.int32), rng.integers(0, L, n).astype(np.uint32),
                    rng.integers(0, L, n).astype(np.uint32), rng.integers(1, W + 14, n).astype(np.uint32),
                    rng.integers(0, W, n).astype(np.uint32))
        for scores in ((1, -2, -3, -1), (1, -1, -2, -1)):
            compare_dp(b, oracle_dp(olib, b, scores), ref_dp(dlib, b, scores), f"adv {seed}")


def test_seed_oracle_matches_the_reference_sort_and_merge():
    """oracle/seed_oracle.c == the reference's own radix-sort macros + singleMerge body (oracle/_ref/libref_seed.so):
    ties, more than 255 hits (the reference's 8-bit strand pass overflows its counters and sorts into a freed array),
    estimated starts that wrapped below zero."""
    import ctypes as C
    import os
    import helpers
    path = os.path.join(helpers.ROOT, "oracle", "_ref", "libref_seed.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_seed.so not built")
    ref = C.CDLL(path)
    U, I = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    ref.ref_seed_sort_merge.restype = C.c_uint32
    ref.ref_seed_sort_merge.argtypes = [U, U, I, C.c_uint32, U, U, I]
    olib = helpers.load_oracle()
    olib.s3o_seed_candidates.restype = C.c_uint64
    olib.s3o_seed_candidates.argtypes = [U, U, U, I, U, U, U, U, C.c_uint64, C.c_uint32, U, U, I, C.c_uint64]
    rng = np.random.default_rng(12)
    for n, nreads, span in ((5000, 40, 3000), (300, 3, 200), (1, 1, 10), (20000, 2000, 100000)):
        # as ranges of one position each over an identity "suffix array", so that both sides see the same hits
        sa = np.arange(span + 64, dtype=np.uint32)
        rid = rng.integers(0, nreads, n).astype(np.uint32)
        x = rng.integers(0, span, n).astype(np.uint32)
        st = rng.integers(1, 3, n).astype(np.int32)
        off = rng.integers(0, 60, n).astype(np.uint32)
        sl = np.full(n, 28, np.uint32)
        rl = np.full(n, 100, np.uint32)
        est = np.where(st == 1, x - off, x + sl + off - rl).astype(np.uint32)       # wraps for small x
        o = [np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.int32)]
        m = ref.ref_seed_sort_merge(rid.ctypes.data_as(U), est.ctypes.data_as(U), st.ctypes.data_as(I), n,
                                    o[0].ctypes.data_as(U), o[1].ctypes.data_as(U), o[2].ctypes.data_as(I))
        g = [np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.int32)]
        m2 = olib.s3o_seed_candidates(sa.ctypes.data_as(U), x.ctypes.data_as(U), x.ctypes.data_as(U), st.ctypes.data_as(I),
                                      rid.ctypes.data_as(U), off.ctypes.data_as(U), sl.ctypes.data_as(U), rl.ctypes.data_as(U),
                                      n, 0xFFFFFFFF, g[0].ctypes.data_as(U), g[1].ctypes.data_as(U), g[2].ctypes.data_as(I), n)
        assert m == m2
        for a, b in zip(o, g):
            assert np.array_equal(a[:m], b[:m])


def test_seed_pair_oracle_matches_the_reference_join():
    """s3o_seed_pair_candidates == the reference's own findRevStart + pairEndMerge + sorts (oracle/_ref/libref_seed_pair.so),
    for every combination of leg strands -- with equal legs the reference's second merge call joins against the array its
    first call thinned in place."""
    import ctypes as C
    import os
    import helpers
    path = os.path.join(helpers.ROOT, "oracle", "_ref", "libref_seed_pair.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_seed_pair.so not built")
    ref = C.CDLL(path)
    o = helpers.load_oracle()
    U, I = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    ref.ref_seed_pair_merge.restype = C.c_uint32
    ref.ref_seed_pair_merge.argtypes = [U, U, C.c_uint32, U, U, C.c_uint32, U, C.c_int, C.c_int, C.c_int, C.c_int, U, U, U]
    o.s3o_seed_pair_candidates.restype = C.c_uint64
    o.s3o_seed_pair_candidates.argtypes = [U] + [U, U, I, U, U, U, U, C.c_uint64] * 2 + [C.c_uint32, U, C.c_int, C.c_int, C.c_int, C.c_int,
                                                                                         U, U, U, C.c_uint64]
    rng = np.random.default_rng(5)

    def u(a):
        return a.ctypes.data_as(U)
    for n0, n1, npairs, span in ((3000, 3000, 60, 5000), (500, 800, 5, 1500), (1, 1, 1, 100), (20000, 15000, 1500, 200000), (4000, 10, 40, 3000)):
        for legs in ((1, 2), (2, 1), (1, 1), (2, 2)):
            sa = np.arange(span + 200, dtype=np.uint32)
            sides = []
            for n in (n0, n1):
                rid = (rng.integers(0, npairs, n) * 2).astype(np.uint32)
                x = rng.integers(0, span, n).astype(np.uint32)
                st = rng.integers(1, 3, n).astype(np.int32)
                off = rng.integers(0, 60, n).astype(np.uint32)
                sl = np.full(n, 28, np.uint32)
                rl = rng.choice([100, 100, 150, 75], n).astype(np.uint32)
                sides.append((x, st, rid, off, sl, rl))
            lens = rng.choice([100, 100, 150, 75], npairs * 2 + 2).astype(np.uint32)
            keys, poss = [], []
            for x, st, rid, off, sl, rl in sides:
                si = (st - 1).astype(np.uint32)
                poss.append(np.where(si == 0, x - off, x + sl + off - rl).astype(np.uint32))
                keys.append((rid | (si << 31)).astype(np.uint32))
            cap = (n0 + n1) * 50 + 10
            want = [np.zeros(cap, np.uint32) for _ in range(3)]
            m = ref.ref_seed_pair_merge(u(keys[0]), u(poss[0]), n0, u(keys[1]), u(poss[1]), n1, u(lens), 200, 500, legs[0], legs[1],
                                        u(want[0]), u(want[1]), u(want[2]))
            got = [np.zeros(cap, np.uint32) for _ in range(3)]
            a = []
            for x, st, rid, off, sl, rl in sides:
                a += [u(x), u(x), st.ctypes.data_as(I), u(rid), u(off), u(sl), u(rl), len(x)]
            m2 = o.s3o_seed_pair_candidates(u(sa), *a, 0xFFFFFFFF, u(lens), 200, 500, legs[0], legs[1], u(got[0]), u(got[1]), u(got[2]), cap)
            assert m == m2, (legs, m, m2)
            for w, g in zip(want, got):
                assert np.array_equal(w[:m], g[:m]), legs


def test_pair_oracle_matches_the_reference_pairing():
    """oracle/pair_oracle.c == the reference's PEMappingOccurrences + PEStatsPEOutput (oracle/_ref/libref_pair.so): every
    leg-strand combination, report-all and report-one, ties in position, empty lists, positions at both ends of the
    32-bit range (the predicates wrap there like the reference's unsigned arithmetic)."""
    import helpers
    ref = helpers.load_ref_pair()
    if ref is None:
        pytest.skip("oracle/_ref/libref_pair.so not built")
    rng = np.random.default_rng(5)
    found = 0
    for legs in ((1, 2), (2, 1), (1, 1), (2, 2)):
        for report_one in (False, True):
            for near_edges in (False, True):
                lists = helpers.make_occurrence_lists(rng, 300, near_edges=near_edges)
                pl = rng.integers(60, 151, 300).astype(np.uint32)
                for lb, ub in ((200, 500), (1, 300)):
                    want = helpers.ref_pair_occurrences(ref, lists, pl, lb, ub, *legs, report_one)
                    got = helpers.oracle_pair_occurrences(lists, pl, lb, ub, *legs, report_one)
                    assert helpers.same_pairing(got, want), (legs, report_one, near_edges, lb, ub)
                    found += len(want["pos1"])
    assert found > 5000


def test_retain_oracle_matches_the_reference_filters():
    """oracle/retain_oracle.c == the reference's retainAllBest / retainAllBestWithCap / retainAllBestAndSecBest
    (oracle/_ref/libref_retain.so): empty lists on either side, the minimum on the SA side, on the occurrence side, on both,
    caps that cut a range short, that fall between entries, that the first best occurrence overrides"""
    import helpers
    ref = helpers.load_ref_retain()
    if ref is None:
        pytest.skip("oracle/_ref/libref_retain.so not built")
    rng = np.random.default_rng(31)
    for mode, caps in ((0, (0,)), (1, (1, 2, 5, 30, 1000)), (2, (0,))):
        for cap in caps:
            lists = helpers.make_hit_lists(rng, 1500)
            want = helpers.ref_retain_best(ref, lists, mode, cap)
            got = helpers.oracle_retain_best(lists, mode, cap)
            assert helpers.same_retained(got, want), (mode, cap)
            assert int(want["sa_off"][-1]) > 300 and int(want["occ_off"][-1]) > 300
            lists = helpers.make_hit_lists(rng, 800, high_counts=True)          # mismatchCount is a uint8_t: 128..255 are large
            assert helpers.same_retained(helpers.oracle_retain_best(lists, mode, cap), helpers.ref_retain_best(ref, lists, mode, cap)), (mode, cap, "high")


def test_dp_without_the_full_table_gives_the_same_alignments():
    """design check for the next DP kernel: a score pass that keeps (H, E) of every 32nd column, then a
    second sweep of only the columns the traceback can reach (restarting from an earlier kept column when it reaches
    further) -- same scores, hit locations, tie counts and patterns as the full-table oracle, on engine-shaped batches
    and on adversarial ones (random pairs, big clips, random anchors, cutoff 0)"""
    import ctypes as C
    import helpers
    from helpers import I32P, U8P, U32P, _opt
    lib = load_oracle_dp()
    lib.s3o_dp_align_resweep.restype = C.c_ulonglong
    lib.s3o_dp_align_resweep.argtypes = helpers._DP_ARGS + [C.c_uint32, C.c_uint32, C.POINTER(C.c_ulonglong)]

    def resweep(b, scores, every, slack, stats):
        sc, hit, cnt, pat = b.outputs()
        lib.s3o_dp_align_resweep(u32p(b.dna), u32p(b.dna_len), b.max_dna, u32p(b.read), u32p(b.read_len), b.max_read,
                                 b.cutoff.ctypes.data_as(I32P), sc.ctypes.data_as(I32P), u32p(hit), u32p(cnt), pat.ctypes.data_as(U8P), b.n,
                                 _opt(b.clip_lt), _opt(b.clip_rt), _opt(b.anchor_l), _opt(b.anchor_r), *scores, every, slack, stats)
        return sc, hit, cnt, pat

    G = synth.random_genome(300_000, seed=9)
    cost = {}
    for mode, L in (("rescue", 100), ("single", 100), ("rescue", 150)):
        b = make_dp_batch(G, 600, L, mode, seed=L, indel_rate=0.01)
        for every, slack in ((32, 16), (16, 0), (64, 40)):
            stats = (C.c_ulonglong * 3)(0, 0, 0)
            assert compare_dp(b, resweep(b, (1, -2, -3, -1), every, slack, stats), oracle_dp(lib, b), f"resweep {mode} {L} {every}") > 500
            cost[(mode, L, every, slack)] = (stats[0] / stats[2], stats[1])
    # mate-rescue windows are 4 x the read: with 32-column checkpoints and 16 columns of slack about a third of the columns
    # are swept again and restarts are rare
    assert cost[("rescue", 100, 32, 16)][0] < 0.45 and cost[("rescue", 100, 32, 16)][1] <= 6
    rng = np.random.default_rng(3)
    adv = (C.c_ulonglong * 3)(0, 0, 0)
    for seed in range(3):
        n, L = 500, int(rng.integers(20, 130))
        W = L + int(rng.integers(5, 300))
        dna = rng.integers(0, 4, (n, W)).astype(np.uint8)
        read = rng.integers(0, 4, (n, L)).astype(np.uint8)
        for t in range(0, n, 2):
            o = rng.integers(0, W - L)
            dna[t, o:o + L] = read[t]
        b = DPBatch(dna, rng.integers(W - 3, W + 1, n).astype(np.uint32), read, rng.integers(max(L - 8, 1), L + 1, n).astype(np.uint32),
                    W + 14, (L // 4 + 1) * 4, np.zeros(n, np.int32), rng.integers(0, L, n).astype(np.uint32),
                    rng.integers(0, L, n).astype(np.uint32), rng.integers(1, W + 14, n).astype(np.uint32), rng.integers(0, W, n).astype(np.uint32))
        compare_dp(b, resweep(b, (1, -2, -3, -1), 32, 8, adv), oracle_dp(lib, b), f"resweep adversarial {seed}")
        compare_dp(b, resweep(b, (1, -1, -2, -1), 8, 0, adv), oracle_dp(lib, b, (1, -1, -2, -1)), f"resweep adversarial {seed} b")
    print("resweep cost (fraction of columns swept again, restarts per 600):", {k: (round(v[0], 3), v[1]) for k, v in cost.items()},
          f"; adversarial {adv[0]} of {adv[2]} columns, {adv[1]} restarts")


clib = helpers.load_ref_cpu_search()


@pytest.mark.skipif(clib is None, reason="oracle/_ref/libref_cpu_search.so not built")
@pytest.mark.parametrize("L", [100, 75])
def test_search_oracle_finds_what_the_reference_cpu_search_finds(L):
    """The reference's own CPU search -- ProcessReadDoubleStrand2 per case on the models SRAModelConstruct builds, with its 13-mer
    lookup tables and check-and-extend, as hostKernel runs it for a read the GPU left over (CPUfunctions.cpp:1313-1328) --
    reports, for every read and mismatch level, exactly the (position, strand, mismatches) set the answer slots of the
    search oracle expand to."""
    G = synth.random_genome(300_000, seed=5)
    idx = fmindex.build_index(G, keep_sa=True)
    hi = HostIndex(idx)
    sa = idx.fwd.sa.cpu().numpy().astype(np.uint32)
    pac = idx.packed_text.cpu().numpy().view(np.uint32)
    ref = helpers.RefCpuSearch(clib, hi.bwt, hi.rbwt, hi.isa0, hi.risa0, hi.n, pac, sa, threads=2)
    olib = load_oracle()
    n = 1500
    try:
        for k in range(5):
            assert ref.describe(L, k, formats.NUM_CASES[k]).count("case") == formats.NUM_CASES[k]
            rs = synth.simulate_single_end(G, n, L, seed=30 + k, sub_rate=0.02)
            reads = np.ascontiguousarray(rs.reads.cpu().numpy().astype(np.uint8))
            got = ref.search(reads, k, formats.NUM_CASES[k], threads=2, out_cap=2048)
            wpq = formats.word_per_query(L)
            lens = np.zeros(formats.ceil32(n), np.uint32)
            lens[:n] = L
            q = formats.pack_queries(reads, lens[:n], wpq)
            allowed = 1024
            wpa = 2 * allowed
            bad = np.zeros(formats.ceil32(n), np.uint8)
            want = [set() for _ in range(n)]
            for case in range(formats.NUM_CASES[k]):
                a = np.zeros(formats.ceil32(n) * wpa, np.uint32)
                oracle_launch(olib, hi, case, q, lens, n, wpq, a, bad, 0, k, allowed, wpa)
                v = formats.answers_view(a, n, wpa)
                for r in range(n):
                    w = v[r]
                    assert int(w[0]) <= 0xFFFFFFFD
                    if int(w[0]) == 0xFFFFFFFD:
                        continue
                    for i in range(allowed):
                        a0, a1 = int(w[2 * i]), int(w[2 * i + 1])
                        if a0 >= 0xFFFFFFFD or a1 >= 0xFFFFFFFD:
                            break
                        for j in range(a0, a0 + (a1 & 0xFFFFFF) + 1):
                            want[r].add((int(sa[j]), ((a1 >> 27) & 1) + 1, (a1 >> 24) & 7))
            found = 0
            for r in range(n):
                assert len(got["hits"][r]) == len(set(got["hits"][r])), (k, r)
                assert set(got["hits"][r]) == want[r], (k, r)
                assert int(got["counts"][r, 3]) == len(want[r])
                found += bool(want[r])
            assert found > n // 10
    finally:
        ref.free()


vlib = helpers.load_ref_validate()


@pytest.mark.skipif(vlib is None, reason="oracle/_ref/libref_validate.so not built")
def test_validate_oracle_matches_the_reference_validation():
    """validateAlignments + the packers / popcount distance it calls, cut from the reference, against oracle/validate_oracle.c"""
    rng = np.random.default_rng(1)
    G = rng.integers(0, 4, 60_000).astype(np.uint8)
    pac = helpers.pack_text(G)
    olib = load_oracle()
    kept = changed = 0
    for read, seed, pos, st, mm, keep, mins, mx, mh in helpers.validation_cases(rng, G, 4000):
        a = helpers.validate_one(olib.s3o_validate_one, pac, len(G), read, seed, pos, st, mm, keep, mins, mx, mh)
        b = helpers.validate_one(vlib.ref_validate, pac, len(G), read, seed, pos, st, mm, keep, mins, mx, mh)
        assert a == b, (read.tolist(), seed, pos, st, mm, keep, mins, mx, mh)
        kept += len(a[0])
        changed += a[0] != pos
    assert kept > 1000 and changed > 1000


def test_index_file_formats_equal_reference_builders(tmp_path):
    """the six files s3_index_load maps (.bwt, .fmv.gpu, .rev.bwt, .rev.fmv.gpu, .sa, .pac) written from our builder's arrays ==
    the files soap3-dp-builder + BGS-Build write for the same genome, byte for byte (headers, payloads, the .pac tail byte)"""
    import tempfile
    from helpers import run_reference_builders, write_reference_files
    from soap3dp_b200 import synth as _synth
    if not os.path.exists(os.path.join(helpers.ROOT, "oracle", "_ref", "soap3-dp-builder")):
        pytest.skip("oracle/_ref builders not built")
    for n in (250_003, 100_000):                       # a length that leaves 3 bases in the last .pac byte, and one that fills it
        G = _synth.random_genome(n, seed=77)
        idx = fmindex.build_index(G, keep_sa=True)
        with tempfile.TemporaryDirectory() as tmp:
            ref = run_reference_builders(G, tmp)
            ours = os.path.join(tmp, "ours.index")
            write_reference_files(ours, idx)
            for e in (".bwt", ".fmv.gpu", ".rev.bwt", ".rev.fmv.gpu", ".sa", ".pac"):
                assert open(ref + e, "rb").read() == open(ours + e, "rb").read(), (n, e)
