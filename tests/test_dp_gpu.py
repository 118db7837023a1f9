"""Parity of the CUDA DP path (through the C ABI) with the DP oracle:
scores, hitLocs, tie counts for every alignment and the pattern bytes of every
alignment that reaches its cutoff -- bit-exact."""
import numpy as np
import pytest

from helpers import compare_dp, load_oracle_dp, make_dp_batch, oracle_dp, DPBatch
from soap3dp_b200 import api, formats, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def genome():
    return synth.random_genome(1_000_000, seed=21)


def _run(b, scores=(1, -2, -3, -1)):
    al = api.SemiGlobalAligner(b.max_read, b.max_dna, max(b.n, 1), *scores)
    try:
        return al.performAlignment(b.dna, b.dna_len, b.read, b.read_len, b.cutoff, b.n, b.clip_lt, b.clip_rt,
                                   b.anchor_l, b.anchor_r)
    finally:
        al.freeMemory()


@pytest.mark.parametrize("mode", ["single", "rescue"])
@pytest.mark.parametrize("L", [100, 150, 75, 36, 250])
def test_dp_bit_exact(genome, mode, L):
    olib = load_oracle_dp()
    for scores in ((1, -2, -3, -1), (2, -3, -5, -2)):
        b = make_dp_batch(genome, 1030, L, mode, seed=L + len(mode), indel_rate=0.006)
        got = _run(b, scores)
        want = oracle_dp(olib, b, scores)
        npass = compare_dp(b, got, want, f"{mode} L={L} {scores}")
        assert npass > 900


def test_dp_null_clip_and_anchor_arrays(genome):
    olib = load_oracle_dp()
    b = make_dp_batch(genome, 500, 100, "single", seed=5)
    b.clip_lt = b.clip_rt = None
    compare_dp(b, _run(b), oracle_dp(olib, b), "null clips")
    b = make_dp_batch(genome, 500, 100, "rescue", seed=6)
    b.anchor_l = b.anchor_r = None
    compare_dp(b, _run(b), oracle_dp(olib, b), "null anchors")


def test_dp_random_sequences_low_cutoff(genome):
    """Unrelated read/window pairs with cutoff 0 and large clips: exercises the soft-clip
    exits, gap chains and tie counting far from the easy diagonal."""
    olib = load_oracle_dp()
    rng = np.random.default_rng(9)
    n, L, W = 2000, 60, 90
    dna = rng.integers(0, 4, (n, W)).astype(np.uint8)
    read = rng.integers(0, 4, (n, L)).astype(np.uint8)
    # half of them: plant the read with a few edits so that real alignments exist too
    for t in range(0, n, 2):
        o = rng.integers(0, W - L)
        dna[t, o:o + L] = read[t]
        for _ in range(3):
            dna[t, o + rng.integers(0, L)] = rng.integers(0, 4)
    b = DPBatch(dna, rng.integers(W - 10, W + 1, n).astype(np.uint32), read, rng.integers(L - 8, L + 1, n).astype(np.uint32),
                W + 14, 64, np.zeros(n, np.int32), rng.integers(0, 30, n).astype(np.uint32),
                rng.integers(0, 30, n).astype(np.uint32), rng.integers(1, W + 14, n).astype(np.uint32),
                rng.integers(0, 40, n).astype(np.uint32))
    for scores in ((1, -2, -3, -1), (1, -1, -2, -1), (3, -1, -4, -1)):
        compare_dp(b, _run(b, scores), oracle_dp(olib, b, scores), f"random {scores}")


def test_dp_empty_and_bad_args(genome):
    al = api.SemiGlobalAligner(104, 162, 64)
    z = np.zeros(32 * 11, np.uint32)
    out = al.performAlignment(z, z[:32], z[:32 * 7], z[:32], np.zeros(32, np.int32), 0)
    assert out[0].shape[0] == 32
    with pytest.raises(api.S3Error):
        al.performAlignment(z, z[:32], z[:32 * 7], z[:32], np.zeros(32, np.int32), 65)
    al.freeMemory()
    with pytest.raises(api.S3Error):
        api.SemiGlobalAligner(2000, 162, 64)


@pytest.mark.parametrize("L", [100, 150, 61])
def test_windows_packed_on_the_device_give_the_same_alignments(L):
    """s3_dp_align_windows (read ids + window starts; batch arrays built on the device from the index's text and the
    query buffer) == s3_dp_align on the arrays the host packers build (formats.pack_dp_sequences = packRead /
    repackDNA, DV-DPfunctions.cu:1469-1524), both strands, ragged lengths, windows at the very ends of the text."""
    from soap3dp_b200 import fmindex
    G = synth.random_genome(200_000, seed=21)
    idx = fmindex.build_index(G)
    gi = api.GPUINDEXUpload(idx, device=0, with_text=True, with_sa=True)
    g = G.numpy()
    rng = np.random.default_rng(L)
    nq, n = 700, 1500
    rs = synth.simulate_single_end(G, nq, L, seed=50 + L, sub_rate=0.02, indel_rate=0.004, margin=400)
    reads = rs.reads.numpy()
    qlen = np.zeros(formats.ceil32(nq), np.uint32)
    qlen[:nq] = L
    qlen[5:nq:7] = L - 4
    wpq = formats.word_per_query(L)
    queries = formats.pack_queries(reads, qlen[:nq], wpq)
    rid = rng.integers(0, nq, n).astype(np.uint32)
    pos = rs.pos.numpy()[rid]
    strand = (rs.strand.numpy()[rid] + 1).astype(np.uint8)            # 1 as given, 2 reverse-complemented
    strand[::10] = 3 - strand[::10]                                   # some against the wrong strand: low scores
    margin = 25
    max_read = (L // 4 + 1) * 4
    max_dna = max_read + 2 * margin + 8
    start = np.clip(pos - margin + rng.integers(-5, 6, n), 0, len(g) - 1).astype(np.uint32)
    dlen = np.minimum(L + 2 * margin, len(g) - start).astype(np.uint32)
    start[0], dlen[0] = 0, L + 2 * margin
    start[1] = len(g) - (L + 2 * margin); dlen[1] = L + 2 * margin
    dlen[2::13] -= 7
    rlen = qlen[rid]
    cutoff = np.ceil(0.3 * rlen).astype(np.int32)
    clip_lt = np.where(strand == 1, 3, 8).astype(np.uint32)
    clip_rt = np.where(strand == 1, 8, 3).astype(np.uint32)
    # the host packers' view of the same batch
    dna = np.zeros((n, int(dlen.max())), np.uint8)
    for t in range(n):
        dna[t, :dlen[t]] = g[start[t]:start[t] + dlen[t]]
    rd = np.zeros((n, L), np.uint8)
    for t in range(n):
        r = reads[rid[t], :rlen[t]]
        rd[t, :rlen[t]] = r if strand[t] == 1 else 3 - r[::-1]
    al = api.SemiGlobalAligner(max_read, max_dna, n)
    up = formats.ceil32(n)

    def padded(a, dtype=np.uint32):
        o = np.zeros(up, dtype)
        o[:n] = a
        return o
    want = al.performAlignment(formats.pack_dp_sequences(dna, max_dna), padded(dlen), formats.pack_dp_sequences(rd, max_read),
                               padded(rlen), padded(cutoff, np.int32), n, padded(clip_lt), padded(clip_rt))
    got = al.performAlignmentOnWindows(gi, queries, qlen, nq, wpq, padded(rid), padded(strand, np.uint8), padded(start), padded(dlen),
                                       padded(cutoff, np.int32), n, padded(clip_lt), padded(clip_rt))
    for a, b, name in zip(got, want, ("scores", "hitLocs", "maxScoreCounts", "pattern")):
        if name == "pattern":
            a = a.reshape(up, -1)[:n]; b = b.reshape(up, -1)[:n]
            ok = want[0][:n] >= cutoff                                   # patterns exist for traced alignments only
            for t in np.nonzero(ok)[0]:
                ea, eb = a[t], b[t]
                la = int(np.argmax(ea == 0)); lb = int(np.argmax(eb == 0))
                assert la == lb and np.array_equal(ea[:la], eb[:lb]), f"pattern {t}"
        else:
            assert np.array_equal(a[:n], b[:n]), name
    assert int((want[0][:n] >= cutoff).sum()) > n // 2
    with pytest.raises(api.S3Error):                                      # window past the end of the text
        bad = padded(start).copy(); bad[3] = len(g) - 5
        al.performAlignmentOnWindows(gi, queries, qlen, nq, wpq, padded(rid), padded(strand, np.uint8), bad, padded(dlen),
                                     padded(cutoff, np.int32), n)
    al.freeMemory()
    api.GPUINDEXFree(gi)


def test_reference_cuda_kernels_give_the_same_alignments():
    """Second opinion: the reference's own DP kernels compiled for sm_100a (oracle/_ref/libref_dp_cuda.so, built by
    oracle/build_ref.sh where the reference is present) run on this GPU == our kernels == the oracle."""
    from helpers import load_ref_dp_cuda, ref_dp_cuda
    lib = load_ref_dp_cuda()
    if lib is None:
        pytest.skip("oracle/_ref/libref_dp_cuda.so not built")
    G = synth.random_genome(300_000, seed=77)
    for mode, L in (("rescue", 100), ("single", 100)):
        b = make_dp_batch(G, 2000, L, mode, seed=13)
        ref_out, ms = ref_dp_cuda(lib, b)
        al = api.SemiGlobalAligner(b.max_read, b.max_dna, b.n)
        ours = al.performAlignment(b.dna, b.dna_len, b.read, b.read_len, b.cutoff, b.n, b.clip_lt, b.clip_rt, b.anchor_l, b.anchor_r)
        al.freeMemory()
        assert compare_dp(b, ours, ref_out, f"ours vs reference CUDA kernels ({mode})") > 1000
        assert compare_dp(b, ref_out, oracle_dp(load_oracle_dp(), b), f"reference CUDA kernels vs oracle ({mode})") > 1000
        assert ms > 0


@pytest.mark.parametrize("mode,L,scores", [("single", 100, (1, -4, -3, -1)),      # mismatch below gap open: not the 16x2 tables
                                           ("rescue", 100, (1, -4, -3, -1)),
                                           ("single", 300, (1, -2, -3, -1)),      # reads above 256 bases
                                           ("rescue", 100, (1, -2, -3, 0))])      # free extension
def test_dp_32bit_path_bit_exact(genome, mode, L, scores):
    """Score parameters or read lengths outside what 16-bit lanes hold take the 32-bit kernels (one alignment per warp,
    byte-plane traceback); same oracle."""
    olib = load_oracle_dp()
    b = make_dp_batch(genome, 700, L, mode, seed=3 * L + len(mode), indel_rate=0.006)
    npass = compare_dp(b, _run(b, scores), oracle_dp(olib, b, scores), f"32-bit path {mode} L={L} {scores}")
    assert npass > 300


def test_dp_forced_32bit_path(genome, monkeypatch):
    monkeypatch.setenv("S3_DP_FORCE_WIDE", "1")
    olib = load_oracle_dp()
    b = make_dp_batch(genome, 600, 100, "rescue", seed=44, indel_rate=0.006)
    assert compare_dp(b, _run(b), oracle_dp(olib, b), "forced 32-bit path") > 400


def test_dp_long_windows(genome):
    """Windows of ~2000 reference bases (deep DP for a wide insert range): the largest column tables the 16x2 kernels keep
    in shared memory, and a plane of 2000+ steps per pair."""
    olib = load_oracle_dp()
    rng = np.random.default_rng(23)
    g = genome.numpy()
    n, L, W = 260, 100, 2000
    start = rng.integers(0, len(g) - W - 1, n)
    dna = g[start[:, None] + np.arange(W)[None, :]].astype(np.uint8)
    read = np.zeros((n, L), np.uint8)
    for t in range(n):
        o = rng.integers(0, W - L)
        read[t] = dna[t, o:o + L]
        for _ in range(3):
            read[t, rng.integers(0, L)] = rng.integers(0, 4)
    dlen = rng.integers(W - 40, W + 1, n).astype(np.uint32)
    rlen = rng.integers(L - 5, L + 1, n).astype(np.uint32)
    b = DPBatch(dna, dlen, read, rlen, W + 8, 104, np.ceil(0.3 * rlen).astype(np.int32),
                rng.integers(0, 9, n).astype(np.uint32), rng.integers(0, 9, n).astype(np.uint32),
                rng.integers(1, W + 8, n).astype(np.uint32), rng.integers(0, 300, n).astype(np.uint32))
    assert compare_dp(b, _run(b), oracle_dp(olib, b), "long windows") > n // 2
    with pytest.raises(api.S3Error):
        api.SemiGlobalAligner(104, 2600, 64)              # beyond what a window may hold (2559 bases)


@pytest.mark.parametrize("mode", ["single", "rescue"])
def test_alignments_decode_to_the_reference_cigars(genome, mode):
    """s3_dp_align -> s3_dp_decode against the restatement of the engines' result loops run on the DP oracle's
    outputs (DV-DPfunctions.cu:1699-1733, DV-DPfunctions.h:514-597, PE.cpp:420-485)"""
    from helpers import load_decode_oracle
    orc = load_decode_oracle()
    scores = (1, -2, -3, -1)
    b = make_dp_batch(genome, 700, 100, mode, seed=44, indel_rate=0.01)
    sc, hit, cnt, pat = _run(b, scores)[:4]
    d = api.decode_alignments(pat, b.pat_len, sc[:b.n], b.read_len, b.cutoff, api.DPScores(*scores))
    wsc, whit, wcnt, wpat = oracle_dp(load_oracle_dp(), b, scores)[:4]
    want = orc.decode_batch(wpat, b.pat_len, wsc[:b.n], b.read_len, b.cutoff, scores)
    got = [(d["cigar"][t], d["sam"][t], int(d["editdist"][t]), int(d["ref_span_delta"][t]),
            tuple(int(x) for x in d["op_counts"][t])) for t in range(b.n)]
    assert got == want
    assert sum(1 for w in want if w[0]) > 500
