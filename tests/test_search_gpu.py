"""Parity of the CUDA search path (through the C ABI) with the oracle.

Bit-exact comparison of every answer word, for every mismatch level, case,
round-1 and round-2 slot size, ragged read lengths, and the isBad carry-over.
"""
import numpy as np
import pytest
import torch

from helpers import HostIndex, fmindex, formats, load_oracle, oracle_launch, s3
from soap3dp_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["stepping", "check_and_extend", "stepping+split", "check_and_extend+split",
                                        "check_and_extend+nosplit"])
def env(request):
    """The same index in every way the search can run: BWT/occ only (every base is an LF-mapping step)
    or with the suffix array and the packed text as well (single-suffix intervals are finished by
    check-and-extend); long enumerations split after the default 256 steps (rare at this size), after
    3 steps (almost every enumerated item goes through spine + tasks + merge, and the heavy list
    overflows so some stay with their lane) or never.  All must give the oracle's answer slots bit for bit."""
    G = synth.random_genome(600_000, seed=11)
    idx = fmindex.build_index(G)
    ce = request.param.startswith("check_and_extend")
    gi = api.GPUINDEXUpload(idx, device=0, with_text=ce, with_sa=ce)
    if request.param.endswith("+split"):
        api.set_split_budget(gi, 3)
    elif request.param.endswith("+nosplit"):
        api.set_split_budget(gi, -1)
    yield G, idx, HostIndex(idx), gi
    api.GPUINDEXFree(gi)


def _oracle_round1(olib, hi, q, lens, n, wpq, k, allowed, wpa, ncases):
    bad = np.zeros(formats.ceil32(n), np.uint8)
    out = []
    for case in range(ncases):
        a = np.zeros(formats.ceil32(n) * wpa, np.uint32)
        oracle_launch(olib, hi, case, q, lens, n, wpq, a, bad, 0, k, allowed, wpa)
        out.append(a)
    return out


def test_rank_probe(env):
    G, idx, hi, gi = env
    olib = load_oracle()
    rng = np.random.default_rng(3)
    n = hi.n
    probes = np.concatenate([rng.integers(0, n + 2, 20000), np.arange(0, 400), np.arange(n - 400, n + 2),
                             np.arange(hi.isa0 - 3, hi.isa0 + 4), np.arange(hi.risa0 - 3, hi.risa0 + 4)]).astype(np.uint32)
    probes = probes[probes <= n + 1]
    from helpers import u32p
    for which, (bwt, occ, isa0) in enumerate(((hi.bwt, hi.occ, hi.isa0), (hi.rbwt, hi.rocc, hi.risa0))):
        got = api.rank_probe(gi, which, probes)
        sel = rng.choice(len(probes), 3000, replace=False)
        for i in list(sel) + list(range(len(probes) - 420, len(probes))):
            for c in range(4):
                assert got[i, c] == olib.s3o_rank(u32p(bwt), u32p(occ), int(probes[i]), c, isa0), (which, i, c)


@pytest.mark.parametrize("L", [100, 36, 150])
@pytest.mark.parametrize("k", [0, 1, 2, 3, 4])
def test_round1_and_round2_bit_exact(env, L, k):
    G, idx, hi, gi = env
    olib = load_oracle()
    n = 4099                                   # ragged: not a multiple of 32
    rs = synth.simulate_single_end(G, n, L, seed=100 + L + k, sub_rate=0.015)
    reads = rs.reads.numpy()
    lens = rs.lengths.numpy().astype(np.uint32)
    lens[::5] = L - 1
    lens[3::11] = L - 7
    wpq = formats.word_per_query(L)
    q = formats.pack_queries(reads, lens, wpq)
    lens_up = np.zeros(formats.ceil32(n), np.uint32)
    lens_up[:n] = lens
    allowed = formats.SA_RANGES_ROUND1[k]
    wpa = 2 * allowed
    ncases = formats.NUM_CASES[k]
    got = api.perform_round1_alignment(gi, q, lens_up, n, wpq, k)
    want = _oracle_round1(olib, hi, q, lens_up, n, wpq, k, allowed, wpa, ncases)
    for c in range(ncases):
        gv, wv = formats.answers_view(got[c], n, wpa), formats.answers_view(want[c], n, wpa)
        assert np.array_equal(gv, wv), f"k={k} case={c}: {np.nonzero((gv != wv).any(1))[0][:5]}"
    # round 2 on the overflowing reads
    allowed2 = formats.SA_RANGES_ROUND2[k]
    wpa2 = 2 * allowed2
    bad_idx, bad_ans = api.perform_round2_alignment(gi, q, lens_up, got, n, wpq, k, wpa)
    for c in range(ncases):
        view = formats.answers_view(want[c], n, wpa)
        exp_idx = np.nonzero(view[:, 0] > 0xFFFFFFFD)[0].astype(np.uint32)
        assert np.array_equal(bad_idx[c], exp_idx)
        nb = len(exp_idx)
        if nb == 0:
            continue
        bq = formats.pack_queries(reads[exp_idx], lens[exp_idx], wpq)
        bl = np.zeros(formats.ceil32(nb), np.uint32)
        bl[:nb] = lens[exp_idx]
        a = np.zeros(formats.ceil32(nb) * wpa2, np.uint32)
        oracle_launch(olib, hi, c, bq, bl, nb, wpq, a, np.zeros(formats.ceil32(nb), np.uint8), 1, k, allowed2, wpa2)
        assert np.array_equal(formats.answers_view(bad_ans[c], nb, wpa2), formats.answers_view(a, nb, wpa2)), (k, c)


def test_reads_at_the_text_ends(env):
    """Reads cut from the first and last bases of the text, and reads hanging off either end:
    the l = 1 start of the backward-only cases and check-and-extend's bounds."""
    G, idx, hi, gi = env
    olib = load_oracle()
    g = G.numpy()
    L, k = 60, 2
    n_text = len(g)
    rng = np.random.default_rng(5)
    reads = []
    for off in list(range(0, 6)) + list(range(n_text - L - 5, n_text - L + 1)):
        r = g[off:off + L].copy()
        reads.append(r)
        r2 = r.copy(); r2[rng.integers(0, L)] ^= 1                     # one substitution
        reads.append(r2)
        reads.append((3 - r[::-1]).astype(np.uint8))                  # reverse strand
    for cut in (1, 3, 17):                                            # hang off the ends
        reads.append(np.concatenate([rng.integers(0, 4, cut).astype(np.uint8), g[:L - cut]]))
        reads.append(np.concatenate([g[n_text - (L - cut):], rng.integers(0, 4, cut).astype(np.uint8)]))
    reads = np.stack(reads).astype(np.uint8)
    n = len(reads)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    wpq = formats.word_per_query(L)
    q = formats.pack_queries(reads, lens[:n], wpq)
    for k in (0, 1, 2, 3):
        allowed = formats.SA_RANGES_ROUND1[k]
        wpa = 2 * allowed
        got = api.perform_round1_alignment(gi, q, lens, n, wpq, k)
        want = _oracle_round1(olib, hi, q, lens, n, wpq, k, allowed, wpa, formats.NUM_CASES[k])
        for c in range(formats.NUM_CASES[k]):
            gv, wv = formats.answers_view(got[c], n, wpa), formats.answers_view(want[c], n, wpa)
            assert np.array_equal(gv, wv), f"k={k} case={c}: reads {np.nonzero((gv != wv).any(1))[0][:8]}"


def test_exact_num_mismatch_flag(env):
    G, idx, hi, gi = env
    olib = load_oracle()
    n, L, k = 1000, 100, 1
    rs = synth.simulate_single_end(G, n, L, seed=77, sub_rate=0.01)
    lens = np.full(formats.ceil32(n), L, np.uint32)
    wpq = formats.word_per_query(L)
    q = formats.pack_queries(rs.reads.numpy(), lens[:n], wpq)
    got = api.perform_round1_alignment(gi, q, lens, n, wpq, k, is_exact_num_mismatch=True)
    bad = np.zeros(formats.ceil32(n), np.uint8)
    for case in range(2):
        a = np.zeros(formats.ceil32(n) * 8, np.uint32)
        oracle_launch(olib, hi, case, q, lens, n, wpq, a, bad, 0, k, 4, 8, exact=1)
        assert np.array_equal(formats.answers_view(got[case], n, 8), formats.answers_view(a, n, 8))


def test_empty_batch_and_bad_args(env):
    G, idx, hi, gi = env
    q = np.zeros(32 * 8, np.uint32)
    lens = np.zeros(32, np.uint32)
    out = api.perform_round1_alignment(gi, q, lens, 0, 8, 2)
    assert len(out) == 4
    with pytest.raises(api.S3Error):
        api.perform_round1_alignment(gi, q, lens, 1, 8, 5, num_cases=1, sa_range_allowed=1, word_per_ans=2)
    with pytest.raises(api.S3Error):
        api.perform_round1_alignment(gi, q, lens, 1, 8, 2, num_cases=7, sa_range_allowed=4, word_per_ans=8)


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4])
def test_capless_search_is_the_uncapped_slot_sequence(env, k):
    """s3_search (CSR, no caps) == for every read, the cases' slot contents in order when the oracle is given
    slots nothing overflows (1024 ranges per case, no isBad carry-over between the cases)."""
    G, idx, hi, gi = env
    olib = load_oracle()
    n, L = (1501, 100) if k < 4 else (701, 100)
    rs = synth.simulate_single_end(G, n, L, seed=300 + k, sub_rate=0.02)
    reads = rs.reads.numpy()
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    lens[7:n:13] = L - 3
    wpq = formats.word_per_query(L)
    q = formats.pack_queries(reads, lens[:n], wpq)
    offsets, sa_l, sa_r, info = api.search(gi, q, lens, n, wpq, k)
    assert offsets[0] == 0 and offsets[-1] == len(sa_l) == len(sa_r) == len(info)
    allowed, wpa = (1024, 2048) if k < 4 else (4096, 8192)
    want = []
    for case in range(formats.NUM_CASES[k]):
        a = np.zeros(formats.ceil32(n) * wpa, np.uint32)
        oracle_launch(olib, hi, case, q, lens, n, wpq, a, np.zeros(formats.ceil32(n), np.uint8), 0, k, allowed, wpa)
        want.append(formats.answers_view(a, n, wpa))
    total = 0
    for r in range(n):
        exp = []
        for case in range(formats.NUM_CASES[k]):
            row = want[case][r]
            assert row[0] != 0xFFFFFFFE, "oracle slot overflowed; enlarge it"
            if row[0] == 0xFFFFFFFD:
                continue
            for s in range(allowed):
                if row[2 * s] == 0xFFFFFFFF and row[2 * s + 1] == 0xFFFFFFFF:
                    break
                w = int(row[2 * s + 1])
                exp.append((int(row[2 * s]), int(row[2 * s]) + (w & 0xFFFFFF), ((w >> 27) & 1) | (((w >> 24) & 7) << 1) | (case << 4)))
        lo, hi_ = int(offsets[r]), int(offsets[r + 1])
        got = list(zip(sa_l[lo:hi_].tolist(), sa_r[lo:hi_].tolist(), info[lo:hi_].tolist()))
        assert got == exp, f"k={k} read {r}: {got[:4]} != {exp[:4]}"
        total += len(exp)
    assert total == offsets[-1] and total > 0


def test_capless_search_empty_batch(env):
    G, idx, hi, gi = env
    offsets, sa_l, sa_r, info = api.search(gi, np.zeros(32 * 8, np.uint32), np.zeros(32, np.uint32), 0, 8, 2)
    assert offsets.tolist() == [0] and len(sa_l) == 0


def test_locate_is_the_suffix_array_gather(env, request):
    """s3_locate == SA[l .. r] for every range (capped per range), in SA order, empty ranges skipped; and the
    positions of a capless 1-mismatch search really are occurrences of the reads within 1 substitution."""
    G, idx, hi, gi = env
    if "check_and_extend" not in request.node.name:
        with pytest.raises(api.S3Error):
            api.locate(gi, np.array([1], np.uint32), np.array([2], np.uint32))
        return
    sa = idx.fwd.sa.numpy().astype(np.uint32) if hasattr(idx.fwd.sa, "numpy") else np.asarray(idx.fwd.sa, np.uint32)
    rng = np.random.default_rng(9)
    n = 5000
    l = rng.integers(0, hi.n - 40, n).astype(np.uint32)
    r = (l + rng.integers(0, 40, n)).astype(np.uint32)
    r[::17] = l[::17] - 1                                   # empty ranges
    l[0], r[0] = 0, 0
    for cap in (0xFFFFFFFF, 5):
        offsets, pos = api.locate(gi, l, r, cap)
        exp = [sa[int(a):int(a) + min(int(b) - int(a) + 1, cap)] if b >= a else sa[:0] for a, b in zip(l.astype(np.int64), r.astype(np.int64))]
        assert offsets[-1] == len(pos) == sum(len(e) for e in exp)
        assert np.array_equal(pos, np.concatenate(exp))
        assert np.array_equal(offsets, np.concatenate([[0], np.cumsum([len(e) for e in exp])]).astype(np.uint64))
    # end to end: search -> locate -> the text at the position is within 1 substitution of the read
    nr, L = 300, 64
    rs = synth.simulate_single_end(G, nr, L, seed=41, sub_rate=0.01)
    lens = np.zeros(formats.ceil32(nr), np.uint32)
    lens[:nr] = L
    wpq = formats.word_per_query(L)
    offs, sa_l, sa_r, info = api.search(gi, formats.pack_queries(rs.reads.numpy(), lens[:nr], wpq), lens, nr, wpq, 1)
    o2, pos = api.locate(gi, sa_l, sa_r, 8)
    g = G.numpy()
    reads = rs.reads.numpy()
    checked = 0
    for q in range(nr):
        for e in range(int(offs[q]), int(offs[q + 1])):
            rd = reads[q] if (info[e] & 1) == 0 else (3 - reads[q][::-1])
            for p in pos[int(o2[e]):int(o2[e + 1])]:
                assert int((g[int(p):int(p) + L] != rd).sum()) == ((int(info[e]) >> 1) & 7) <= 1
                checked += 1
    assert checked > nr // 2


def test_two_halves_side_by_side_give_the_same_slots(env, monkeypatch):
    """Large batches are searched as two halves on two streams (device entry point) and the host entry point puts
    every other chunk on the side stream; forced here at test size.  Slots must stay the oracle's."""
    G, idx, hi, gi = env
    olib = load_oracle()
    n, L, k = 4099, 100, 2
    rs = synth.simulate_single_end(G, n, L, seed=907, sub_rate=0.015)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    lens[3:n:11] = L - 5
    wpq = formats.word_per_query(L)
    q = formats.pack_queries(rs.reads.numpy(), lens[:n], wpq)
    allowed = formats.SA_RANGES_ROUND1[k]
    wpa, ncases = 2 * allowed, formats.NUM_CASES[k]
    want = _oracle_round1(olib, hi, q, lens, n, wpq, k, allowed, wpa, ncases)
    # host entry point, two chunks
    monkeypatch.setenv("S3_HOST_CHUNKS", "2")
    got = api.perform_round1_alignment(gi, q, lens, n, wpq, k)
    for c in range(ncases):
        assert np.array_equal(formats.answers_view(got[c], n, wpa), formats.answers_view(want[c], n, wpa)), f"host chunks, case {c}"
    # device entry point, two halves
    monkeypatch.setenv("S3_SIDE_MIN_READS", "64")
    dq = torch.from_numpy(q.view(np.int32)).cuda()
    dl = torch.from_numpy(lens.view(np.int32)).cuda()
    da = [torch.empty(formats.ceil32(n) * wpa, dtype=torch.int32, device="cuda") for _ in range(ncases)]
    torch.cuda.synchronize()
    for _ in range(2):                                                    # twice: the side context is reused
        api.search_round1_device(gi, dq.data_ptr(), dl.data_ptr(), n, wpq, k, ncases, allowed, wpa, [a.data_ptr() for a in da])
    torch.cuda.ExternalStream(gi.stream).synchronize()
    for c in range(ncases):
        gv = formats.answers_view(da[c].cpu().numpy().view(np.uint32), n, wpa)
        assert np.array_equal(gv, formats.answers_view(want[c], n, wpa)), f"device halves, case {c}"


def test_reference_cuda_search_kernels_give_the_same_slots(env):
    """Second opinion: the reference's own search kernels compiled for sm_100a (oracle/_ref/libref_search_cuda.so, built
    by oracle/build_ref_search_cuda.sh where the reference is present) run on this GPU == our slots, k = 1..3."""
    from helpers import load_ref_search_cuda, ref_search_cuda_round1, u32p
    lib = load_ref_search_cuda()
    if lib is None:
        pytest.skip("oracle/_ref/libref_search_cuda.so not built")
    G, idx, hi, gi = env
    assert lib.ref_search_cuda_upload(u32p(hi.bwt), u32p(hi.rbwt), len(hi.bwt), u32p(hi.occ), u32p(hi.rocc), len(hi.occ)) == 0
    try:
        n, L = 3000, 100
        for k in (1, 2, 3):
            rs = synth.simulate_single_end(G, n, L, seed=600 + k, sub_rate=0.015)
            lens = np.zeros(formats.ceil32(n), np.uint32)
            lens[:n] = L
            wpq = formats.word_per_query(L)
            q = formats.pack_queries(rs.reads.numpy(), lens[:n], wpq)
            allowed = formats.SA_RANGES_ROUND1[k]
            wpa = 2 * allowed
            ref, ms = ref_search_cuda_round1(lib, hi, q.copy(), lens, n, wpq, k, allowed, wpa)
            got = api.perform_round1_alignment(gi, q, lens, n, wpq, k)
            for c in range(formats.NUM_CASES[k]):
                assert np.array_equal(formats.answers_view(got[c], n, wpa), formats.answers_view(ref[c], n, wpa)), (k, c)
    finally:
        lib.ref_search_cuda_free()


def test_seed_candidates_match_the_oracle(env, request):
    """s3_seed_candidates (positions from the suffix array in HBM, one stable radix sort, merge walk) == the oracle's
    restatement of decodePositions + singleMerge (pinned against the reference's own sort macros and merge body in the
    CPU tier): real seed hits of a 1-mismatch capless search, both strands, with and without a cap per range."""
    import ctypes as C
    G, idx, hi, gi = env
    if "check_and_extend" not in request.node.name:
        with pytest.raises(api.S3Error):
            api.seed_candidates(gi, [1], [2], [1], [0], [0], [20], [100])
        return
    sa = idx.fwd.sa.numpy().astype(np.uint32) if hasattr(idx.fwd.sa, "numpy") else np.asarray(idx.fwd.sa, np.uint32)
    nr, L, seed_len = 600, 100, 22
    rs = synth.simulate_single_end(G, nr, L, seed=71, sub_rate=0.01)
    reads = rs.reads.numpy()
    # three seeds per read (offsets 0, 39, 78), searched exactly: the ranges a seeding round hands to decodePositions
    offs = np.array([0, 39, 78], np.uint32)
    seeds = np.stack([reads[:, o:o + seed_len] for o in offs], axis=1).reshape(-1, seed_len)
    ns = len(seeds)
    lens = np.zeros(formats.ceil32(ns), np.uint32)
    lens[:ns] = seed_len
    wps = formats.word_per_query(seed_len)
    offsets, sa_l, sa_r, info = api.search(gi, formats.pack_queries(seeds, lens[:ns], wps), lens, ns, wps, 1)
    seed_of = np.repeat(np.arange(ns), np.diff(offsets).astype(np.int64))
    strands = ((info & 1) + 1).astype(np.int32)                           # SARecord.strand: 1 as given, 2 reverse complement
    read_ids = (seed_of // 3).astype(np.uint32)
    off = offs[seed_of % 3]
    sl = np.full(len(sa_l), seed_len, np.uint32)
    rl = np.full(len(sa_l), L, np.uint32)
    assert len(sa_l) > nr
    olib = load_oracle()
    U, I = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    olib.s3o_seed_candidates.restype = C.c_uint64
    olib.s3o_seed_candidates.argtypes = [U, U, U, I, U, U, U, U, C.c_uint64, C.c_uint32, U, U, I, C.c_uint64]
    for cap in (0xFFFFFFFF, 3):
        got = api.seed_candidates(gi, sa_l, sa_r, strands, read_ids, off, sl, rl, cap)
        total = int((np.minimum(sa_r.astype(np.int64) - sa_l + 1, cap)).sum())
        w = [np.zeros(total, np.uint32), np.zeros(total, np.uint32), np.zeros(total, np.int32)]
        m = olib.s3o_seed_candidates(sa.ctypes.data_as(U), sa_l.ctypes.data_as(U), sa_r.ctypes.data_as(U), strands.ctypes.data_as(I),
                                     read_ids.ctypes.data_as(U), off.ctypes.data_as(U), sl.ctypes.data_as(U), rl.ctypes.data_as(U),
                                     len(sa_l), cap, w[0].ctypes.data_as(U), w[1].ctypes.data_as(U), w[2].ctypes.data_as(I), total)
        assert m == len(got[0]) and m >= nr // 2
        for a, b in zip(got, w):
            assert np.array_equal(a, b[:m])
        # most reads: one candidate at the read's true start (its three seeds agree)
        true_pos = rs.pos.numpy()
        hit = sum(1 for r, p in zip(got[0], got[1]) if abs(int(p) - int(true_pos[r])) <= 2)
        assert hit > nr * 0.8


@pytest.mark.parametrize("legs", [(1, 2), (2, 1), (1, 1), (2, 2)])
def test_seed_pair_candidates_match_the_oracle(env, request, legs):
    """s3_seed_pair_candidates == the oracle's restatement of decodePositions x 2 + pairEndMerge x 2 + the final sort
    (pinned against the reference's own functions in the CPU tier): real seed hits of simulated pairs, and a dense
    synthetic set (many hits per read, thinning, equal legs: the second call joins against the array the first thinned)."""
    import ctypes as C
    G, idx, hi, gi = env
    if "check_and_extend" not in request.node.name:
        one = ([1], [2], [1], [0], [0], [20], [100])
        with pytest.raises(api.S3Error):                 # no suffix array on the device: the call must say so
            api.seed_pair_candidates(gi, one, one, np.array([100, 100], np.uint32), 200, 500, legs[0], legs[1])
        return
    sa = idx.fwd.sa.numpy().astype(np.uint32) if hasattr(idx.fwd.sa, "numpy") else np.asarray(idx.fwd.sa, np.uint32)
    olib = load_oracle()
    U, I = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    olib.s3o_seed_pair_candidates.restype = C.c_uint64
    olib.s3o_seed_pair_candidates.argtypes = [U] + [U, U, I, U, U, U, U, C.c_uint64] * 2 + [C.c_uint32, U, C.c_int, C.c_int, C.c_int, C.c_int,
                                                                                           U, U, U, C.c_uint64]

    def check(side0, side1, lens, cap_per_range, min_expected):
        got = api.seed_pair_candidates(gi, side0, side1, lens, 200, 500, legs[0], legs[1], cap_per_range)
        a = []
        keep = []
        for sd in (side0, side1):
            arrs = [np.ascontiguousarray(sd[0], np.uint32), np.ascontiguousarray(sd[1], np.uint32), np.ascontiguousarray(sd[2], np.int32)] + \
                   [np.ascontiguousarray(x, np.uint32) for x in sd[3:7]]
            keep.append(arrs)
            a += [arrs[0].ctypes.data_as(U), arrs[1].ctypes.data_as(U), arrs[2].ctypes.data_as(I)] + [x.ctypes.data_as(U) for x in arrs[3:]] + [len(arrs[0])]
        cap = 4 * (len(side0[0]) + len(side1[0])) * 64 + 16
        w = [np.zeros(cap, np.uint32) for _ in range(3)]
        m = olib.s3o_seed_pair_candidates(sa.ctypes.data_as(U), *a, cap_per_range, lens.ctypes.data_as(U), 200, 500, legs[0], legs[1],
                                          w[0].ctypes.data_as(U), w[1].ctypes.data_as(U), w[2].ctypes.data_as(U), cap)
        assert m == len(got[0]) and m <= cap and m >= min_expected, (m, len(got[0]))
        for x, y in zip(got, w):
            assert np.array_equal(x, y[:m])
        return got

    # --- real pairs: three exact-or-1-mismatch seeds per end
    npairs, L, seed_len = 300, 100, 22
    m1, m2, _ = synth.simulate_paired_end(G, npairs, L, seed=91, insert_lo=200, insert_hi=500)
    offs = np.array([0, 39, 78], np.uint32)
    sides = []
    for mate in (m1, m2):
        reads = mate.reads.numpy()
        seeds = np.stack([reads[:, o:o + seed_len] for o in offs], axis=1).reshape(-1, seed_len)
        ns = len(seeds)
        lens_s = np.zeros(formats.ceil32(ns), np.uint32)
        lens_s[:ns] = seed_len
        wps = formats.word_per_query(seed_len)
        offsets, sa_l, sa_r, info = api.search(gi, formats.pack_queries(seeds, lens_s[:ns], wps), lens_s, ns, wps, 1)
        seed_of = np.repeat(np.arange(ns), np.diff(offsets).astype(np.int64))
        sides.append((sa_l, sa_r, ((info & 1) + 1).astype(np.int32), (2 * (seed_of // 3)).astype(np.uint32), offs[seed_of % 3],
                      np.full(len(sa_l), seed_len, np.uint32), np.full(len(sa_l), L, np.uint32)))
    lens = np.full(2 * npairs, L, np.uint32)
    got = check(sides[0], sides[1], lens, 0xFFFFFFFF, npairs // 2 if legs == (1, 2) else 0)
    if legs == (1, 2):                                   # FR pairs: nearly every pair has its candidate at the true starts
        assert len(set(got[0].tolist())) > npairs * 0.8
    # --- dense synthetic ranges over the same suffix array: wide ranges, many per read, capped and not
    rng = np.random.default_rng(17)
    dense = []
    for n in (4000, 3500):
        l = rng.integers(0, hi.n - 64, n).astype(np.uint32)
        dense.append((l, (l + rng.integers(0, 12, n)).astype(np.uint32), rng.integers(1, 3, n).astype(np.int32),
                      (2 * rng.integers(0, 25, n)).astype(np.uint32), rng.integers(0, 70, n).astype(np.uint32),
                      np.full(n, 25, np.uint32), rng.choice([100, 150], n).astype(np.uint32)))
    lens2 = rng.choice([100, 150, 75], 60).astype(np.uint32)
    check(dense[0], dense[1], lens2, 0xFFFFFFFF, 0)
    check(dense[0], dense[1], lens2, 4, 0)


def test_longest_reads(env):
    """MAX_READ_LENGTH = 1024 (definitions.h:42): 1000-base reads, 64 words per query -- the largest shared-memory
    footprint of the search kernels and the largest task blocks of the split path."""
    G, idx, hi, gi = env
    olib = load_oracle()
    n, L = 150, 1000
    rs = synth.simulate_single_end(G, n, L, seed=5150, sub_rate=0.002)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    lens[1:n:4] = L - 9
    wpq = formats.word_per_query(L)
    assert wpq == 64
    q = formats.pack_queries(rs.reads.numpy(), lens[:n], wpq)
    for k in (1, 2):
        allowed = formats.SA_RANGES_ROUND1[k]
        wpa, ncases = 2 * allowed, formats.NUM_CASES[k]
        got = api.perform_round1_alignment(gi, q, lens, n, wpq, k)
        want = _oracle_round1(olib, hi, q, lens, n, wpq, k, allowed, wpa, ncases)
        for c in range(ncases):
            gv, wv = formats.answers_view(got[c], n, wpa), formats.answers_view(want[c], n, wpa)
            assert np.array_equal(gv, wv), f"k={k} case={c}: {np.nonzero((gv != wv).any(1))[0][:5]}"
        assert sum(int((formats.answers_view(w, n, wpa)[:, 0] < 0xFFFFFFFD).sum()) for w in want) > 20


def test_seed_search_is_the_seeding_driver(env, request):
    """s3_seed_search == single_1_mismatch_alignment2 restated on the oracle's slots: exact first, 1 mismatch for the seeds
    without an alignment, ranges dropped beyond maxHitNum occurrences"""
    import os
    import sys
    from helpers import ROOT
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import seeding_oracle
    G, idx, hi, gi = env
    olib = load_oracle()
    n, L = 1800, 26
    rs = synth.simulate_single_end(G, n, L, seed=77, sub_rate=0.03)
    reads = rs.reads.numpy()
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    lens[5:n:11] = 22
    wps = formats.word_per_query(L)
    q = formats.pack_queries(reads, lens[:n], wps)

    def launcher(qq, ll, m):
        def run(case, k, allowed, wpa):
            a = np.zeros(formats.ceil32(max(m, 1)) * wpa, np.uint32)
            oracle_launch(olib, hi, case, qq, ll, m, wps, a, np.zeros(formats.ceil32(max(m, 1)), np.uint8), 0, k, allowed, wpa)
            return formats.answers_view(a, m, wpa)
        return run

    def launch_one(ids):
        ll = np.zeros(formats.ceil32(max(len(ids), 1)), np.uint32)
        ll[:len(ids)] = lens[ids]
        return launcher(formats.pack_queries(reads[ids], lens[ids], wps), ll, len(ids))
    for max_hit in (2, 40):
        want = seeding_oracle.seeding_driver(launcher(q, lens, n), launch_one, n, max_hit)
        offsets, sa_l, sa_r, strand, status = api.seed_search(gi, q, lens, n, wps, max_hit)
        assert status.tolist() == [w[0] for w in want]
        for s, (st, ranges) in enumerate(want):
            a, b = int(offsets[s]), int(offsets[s + 1])
            assert list(zip(sa_l[a:b].tolist(), sa_r[a:b].tolist(), strand[a:b].tolist())) == ranges, s
        assert (status == 1).sum() > n // 3 and (status == 4).sum() > 0
